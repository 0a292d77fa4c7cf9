"""Drop-in for the reference's ``autoencoder/autoencoder.py``: same class names, constructor signatures, sub-module
names and ``state_dict`` keys, same default initialisation and RNG consumption order -- but the two top-level models
(``ReconstructionAutoencoder``, ``SegmentationAutoencoder``) run as ONE fused CUDA pass each on the launch-plan engine
(``image_segmentation_b200/engine.py``), on the same kernels as the U-Net:

* ``EncoderBlock``  (autoencoder.py:15-33): conv3x3 (bias=False) + BN + ReLU, twice, then MaxPool2d(2,2); the second
  BatchNorm-apply pass writes the skip activation AND the pooled map (and the 2-bit arg-max for the backward routing).
* ``DecoderBlockWithSkips`` (:69-93): ConvTranspose2d(k2,s2) whose epilogue stores straight into the FIRST channels of the
  concat buffer, the encoder's skip activation having been written into the LAST channels (``torch.cat([up, skip], 1)``,
  :91 -- the opposite order of unet/unet.py:63).
* ``DecoderBlockNoSkips`` (:128-147): ConvTranspose2d + the same double conv, no concatenation.
* ``ReconstructionAutoencoder`` (:182-203): encoder -> no-skip decoder -> Conv3x3(64 -> dout) + Sigmoid.
* ``SegmentationEncoder`` / ``SegmentationAutoencoder`` (:218-306): optional pre-trained, optionally FROZEN encoder
  (no weight gradients and no data gradients are computed below the first trainable layer) -> decoder with skips ->
  1x1 classifier fused with the last BatchNorm.

The nn.Conv2d / nn.BatchNorm2d / nn.ConvTranspose2d objects are parameter holders (checkpoints of the reference load
unchanged and any optimizer works); sub-blocks are not callable on their own.
"""
import torch
import torch.nn as nn

from ..engine import Engine, EngineModule as _EngineModule, NetPlan


def _holder_forward(self, *a, **k):
    raise RuntimeError("sub-blocks of the autoencoder family are parameter holders; call the enclosing "
                       "ReconstructionAutoencoder / SegmentationAutoencoder (the whole network runs as one fused CUDA pass)")


class EncoderBlock(nn.Module):
    """Two conv3x3 (bias=False) + BatchNorm + ReLU, then MaxPool2d(2,2); returns (pooled, skip) in the reference."""

    def __init__(self, din, dout):
        super().__init__()
        self.conv1 = nn.Conv2d(din, dout, kernel_size=3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(dout)
        self.relu1 = nn.ReLU()
        self.conv2 = nn.Conv2d(dout, dout, kernel_size=3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(dout)
        self.relu2 = nn.ReLU(inplace=True)
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2)

    forward = _holder_forward


class Encoder(nn.Module):
    """Three encoder blocks: base, 2*base, 4*base channels."""

    def __init__(self, din, base_channels):
        super().__init__()
        self.encoderPart1 = EncoderBlock(din, base_channels)
        self.encoderPart2 = EncoderBlock(base_channels, base_channels * 2)
        self.encoderPart3 = EncoderBlock(base_channels * 2, base_channels * 4)

    forward = _holder_forward


def _double_conv_nobias(cin, cout):
    return nn.Sequential(
        nn.Conv2d(cin, cout, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
        nn.Conv2d(cout, cout, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class DecoderBlockWithSkips(nn.Module):
    """ConvTranspose2d(din_up -> dout) + cat([up, skip]) + two conv3x3/BN/ReLU."""

    def __init__(self, din_up, din_skip, dout):
        super().__init__()
        self.up = nn.ConvTranspose2d(din_up, dout, kernel_size=2, stride=2)
        self.convs = _double_conv_nobias(dout + din_skip, dout)

    forward = _holder_forward


class DecoderWithSkips(nn.Module):
    def __init__(self, base_channels):
        super().__init__()
        b = base_channels
        self.decoderBlock1 = DecoderBlockWithSkips(din_up=b * 4, din_skip=b * 4, dout=b * 2)
        self.decoderBlock2 = DecoderBlockWithSkips(din_up=b * 2, din_skip=b * 2, dout=b)
        self.decoderBlock3 = DecoderBlockWithSkips(din_up=b, din_skip=b, dout=b)

    forward = _holder_forward


class DecoderBlockNoSkips(nn.Module):
    """ConvTranspose2d(din_up -> dout) + two conv3x3/BN/ReLU on the up-sampled map only."""

    def __init__(self, din_up, dout):
        super().__init__()
        self.up = nn.ConvTranspose2d(din_up, dout, kernel_size=2, stride=2)
        self.convs = _double_conv_nobias(dout, dout)

    forward = _holder_forward


class DecoderNoSkips(nn.Module):
    def __init__(self, base_channels):
        super().__init__()
        b = base_channels
        self.decoderBlock1 = DecoderBlockNoSkips(din_up=b * 4, dout=b * 2)
        self.decoderBlock2 = DecoderBlockNoSkips(din_up=b * 2, dout=b)
        self.decoderBlock3 = DecoderBlockNoSkips(din_up=b, dout=b)

    forward = _holder_forward


# ---------------------------------------------------------------------------------------------------------------------
# plan builders
# ---------------------------------------------------------------------------------------------------------------------
def _check_image(model_name, din):
    def check(x):
        if x.dim() != 4 or x.shape[1] != din:
            raise RuntimeError(f"{model_name}: expected input [N,{din},H,W], got {tuple(x.shape)}")
        if x.shape[2] % 8 or x.shape[3] % 8:
            # the reference floors in MaxPool2d and centre-crops the skips (autoencoder.py:83-88); the accelerated path
            # covers the sizes it is trained on
            raise ValueError(f"{model_name}: input height/width must be multiples of 8, got {x.shape[2]}x{x.shape[3]}")
    return check


def _build_encoder(plan: NetPlan, enc: Encoder, x, skip_views=(None, None, None)):
    """Encoder (autoencoder.py:44-54).  ``skip_views[i]`` is where block i+1 writes its skip activation (a slice of the
    decoder's concat buffer) or None when nothing reads it.  Returns the bottleneck Act."""
    n, din, h, w = x.shape
    prev = plan.image_input(din, h, w)
    for i, blk in enumerate((enc.encoderPart1, enc.encoderPart2, enc.encoderPart3)):
        name = f"encoder.part{i + 1}"
        l1 = plan.conv_bn_relu(name + ".c1", blk.conv1, blk.bn1, prev, end_block=True)
        l2 = plan.conv_bn_relu(name + ".c2", blk.conv2, blk.bn2, l1.out, out=skip_views[i], pool=True)
        prev = l2.pooled
    return prev


def _decoder_convs(plan, name, convs, src):
    l1 = plan.conv_bn_relu(name + ".c1", convs[0], convs[1], src)
    l2 = plan.conv_bn_relu(name + ".c2", convs[3], convs[4], l1.out)
    return l2.out


class ReconstructionAutoencoder(_EngineModule):
    """Encoder -> DecoderNoSkips -> Conv3x3(base -> dout) + Sigmoid (autoencoder.py:171-203).

    ``forward(x)``: x is [N, din, H, W] fp32 on a CUDA device (H, W multiples of 8) -> reconstruction [N, dout, H, W]."""

    def __init__(self, din, dout=3, base_channels=64):
        super().__init__()
        self.encoder = Encoder(din, base_channels)
        self.decoder = DecoderNoSkips(base_channels)
        self.decoderOut = nn.Sequential(nn.Conv2d(base_channels, dout, kernel_size=3, padding=1), nn.Sigmoid())
        self.din = din
        self.precision = "bf16"
        self.conv_algo = "auto"
        self._engine = None

    def _build(self, plan: NetPlan, x):
        from ..engine import ConvSigmoidOut
        prev = _build_encoder(plan, self.encoder, x)
        for i, blk in enumerate((self.decoder.decoderBlock1, self.decoder.decoderBlock2, self.decoder.decoderBlock3)):
            name = f"decoder.block{i + 1}"
            ct = plan.conv_transpose(name + ".up", blk.up, prev, end_block=True)
            prev = _decoder_convs(plan, name, blk.convs, ct.out)
        plan.add(ConvSigmoidOut(plan, "decoderOut", self.decoderOut[0], prev))

    def forward(self, x):
        if self._engine is None:
            self._engine = Engine(self, self._build, _check_image("ReconstructionAutoencoder", self.din))
        return self._engine.run(x)



class SegmentationEncoder(nn.Module):
    """Encoder with optional pre-trained weights and freezing (autoencoder.py:206-268); parameter holder."""

    def __init__(self, din, base_channels, pretrained_encoder_path=None, freeze_encoder=True):
        super().__init__()
        self.encoder = Encoder(din, base_channels)
        if pretrained_encoder_path:
            try:
                full_state_dict = torch.load(pretrained_encoder_path, weights_only=False,
                                             map_location=lambda storage, loc: storage)
                if "model_state_dict" in full_state_dict:
                    model_state_dict = full_state_dict["model_state_dict"]
                elif "state_dict" in full_state_dict:
                    model_state_dict = full_state_dict["state_dict"]
                else:
                    model_state_dict = full_state_dict
                encoder_state_dict = {k[len("encoder."):]: v for k, v in model_state_dict.items() if k.startswith("encoder.")}
                if not encoder_state_dict:
                    print("Warning: Could not extract encoder state dict. Checkpoint might be empty or incompatible.")
                else:
                    load_result = self.encoder.load_state_dict(encoder_state_dict, strict=True)
                    print("Loaded encoder weights. Load result:")
                    if load_result.missing_keys:
                        print("  Missing keys:", load_result.missing_keys)
                    if load_result.unexpected_keys:
                        print("  Unexpected keys:", load_result.unexpected_keys)
                    if not load_result.missing_keys and not load_result.unexpected_keys:
                        print("  All keys matched successfully.")
            except FileNotFoundError:
                print(f"Warning: Pre-trained encoder file not found: {pretrained_encoder_path}. Using random weights.")
            except Exception as e:
                print(f"Warning: Error loading weights: {e}. Check compatibility. Using random weights.")
        if freeze_encoder:
            if not pretrained_encoder_path:
                print("Warning: Freezing encoder, but no pre-trained weights were loaded.")
            for param in self.encoder.parameters():
                param.requires_grad = False
            print("Encoder parameters frozen.")
        else:
            print("Encoder parameters are trainable.")

    forward = _holder_forward


class SegmentationAutoencoder(_EngineModule):
    """(Frozen) encoder -> DecoderWithSkips -> 1x1 classifier (autoencoder.py:271-306).

    ``forward(x)``: x is [N, din, H, W] fp32 on a CUDA device (H, W multiples of 8) -> logits [N, num_classes, H, W]."""

    def __init__(self, din, base_channels=64, num_classes=4, pretrained_encoder_path=None, freeze_encoder=True):
        super().__init__()
        self.num_classes = num_classes
        self.encoder = SegmentationEncoder(din, base_channels, pretrained_encoder_path=pretrained_encoder_path,
                                           freeze_encoder=freeze_encoder)
        self.decoder = DecoderWithSkips(base_channels)
        self.finalConv = nn.Conv2d(base_channels, num_classes, kernel_size=1)
        self.din = din
        self.base_channels = base_channels
        self.precision = "bf16"
        self.conv_algo = "auto"
        self._engine = None

    def _build(self, plan: NetPlan, x):
        n, _, h, w = x.shape
        b = self.base_channels
        dec = self.decoder
        # concat buffers in the reference's order [up, skip] (autoencoder.py:91): decoderBlock3 works at full resolution
        cat3 = plan.cat(h, w, [b, b], name="cat.block3")                   # [up(b) | skip1(b)]
        cat2 = plan.cat(h // 2, w // 2, [b, 2 * b], name="cat.block2")     # [up(b) | skip2(2b)]
        cat1 = plan.cat(h // 4, w // 4, [2 * b, 4 * b], name="cat.block1")  # [up(2b) | skip3(4b)]
        prev = _build_encoder(plan, self.encoder.encoder, x, (cat3.parts[1], cat2.parts[1], cat1.parts[1]))
        for i, (blk, cat) in enumerate(((dec.decoderBlock1, cat1), (dec.decoderBlock2, cat2), (dec.decoderBlock3, cat3))):
            name = f"decoder.block{i + 1}"
            plan.conv_transpose(name + ".up", blk.up, prev, out=cat.parts[0], end_block=True)
            prev = _decoder_convs(plan, name, blk.convs, cat)
        plan.head_1x1("finalConv", self.finalConv, prev)

    def forward(self, x):
        if self._engine is None:
            self._engine = Engine(self, self._build, _check_image("SegmentationAutoencoder", self.din))
        return self._engine.run(x)
