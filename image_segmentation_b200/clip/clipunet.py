"""Drop-in for the reference's ``clip/clipunet.py``: frozen CLIP ViT-B/16 encoder + U-Net-style decoder.

Same class names, constructor signatures, sub-module names and ``state_dict`` keys.  The ViT stays what it is in the
reference -- Hugging Face ``CLIPVisionModel`` run by PyTorch (third-party, frozen, clip/clipunet.py:25-30) -- and the
DECODER (clip/clipunet.py:68-188), which holds every trainable parameter, runs as one fused CUDA pass on the launch-plan
engine, on the same kernels as the U-Net:

* the ViT's patch tokens ``hidden_state[:, 1:, :]`` ARE the NHWC feature map [N,14,14,768]; the reference's
  ``reshape(...).permute(0,3,1,2).contiguous()`` (:48-50,:60-62) is never materialised;
* ``init_conv`` (1x1, 768 -> 1024, :125) and every block's ``skip_conv`` (1x1, 768 -> C/2, :84) are plain tensor-core
  GEMMs over all N*196 tokens with the bias in the epilogue;
* ``F.interpolate(skip, size, 'bilinear', align_corners=False)`` (:99-100) writes straight into the second half of the
  block's concat buffer, the ConvTranspose epilogue into the first half (``torch.cat([x, skip], 1)``, :102);
* conv3x3 (bias=False) + BN + ReLU blocks and the 1x1 classifier as in the U-Net (28^2 ... 224^2 maps: partial tiles).
"""
import torch
import torch.nn as nn

from ..autoencoder.autoencoder import _EngineModule, _holder_forward
from ..engine import BilinearUp, Conv1x1, Engine, NetPlan


class ClipViTEncoder(nn.Module):
    """CLIP vision transformer; ``forward`` returns (bottleneck [N,768,14,14], list of skip maps) like the reference.

    ``clip_vit`` (extra, optional): an already constructed ``CLIPVisionModel`` -- e.g. random-init weights where
    ``from_pretrained`` has no network (BASELINE.json config 4); by default the reference's ``from_pretrained`` calls run."""

    def __init__(self, model_name="openai/clip-vit-base-patch16", freeze_encoder=True, skip_indices=[3, 5, 7, 9], clip_vit=None):
        super().__init__()
        self.skip_indices = sorted(skip_indices)
        if clip_vit is not None:
            self.config = clip_vit.config
            self.clip_vit = clip_vit
        else:
            from transformers import CLIPVisionConfig, CLIPVisionModel
            self.config = CLIPVisionConfig.from_pretrained(model_name)
            self.clip_vit = CLIPVisionModel.from_pretrained(model_name)
        if freeze_encoder:
            for param in self.clip_vit.parameters():
                param.requires_grad = False
        self.grid_size = self.config.image_size // self.config.patch_size
        self.hidden_dim = self.config.hidden_size

    def tokens(self, x):
        """[last_hidden_state] + [hidden_states[i] for i in skip_indices], each [N, 1 + grid^2, hidden]."""
        if x.shape[2] != self.config.image_size or x.shape[3] != self.config.image_size:
            print(f"Input image size ({x.shape[2]}x{x.shape[3]}) doesn't match "
                  f"CLIP expected size ({self.config.image_size}x{self.config.image_size}). "
                  f"Behavior may be unexpected. Consider resizing input.")
        outputs = self.clip_vit(pixel_values=x, output_hidden_states=True)
        return [outputs.last_hidden_state] + [outputs.hidden_states[i] for i in self.skip_indices]

    def forward(self, x):
        toks = self.tokens(x)

        def to_map(t):
            return t[:, 1:, :].reshape(x.shape[0], self.grid_size, self.grid_size, self.hidden_dim).permute(0, 3, 1, 2).contiguous()
        return to_map(toks[0]), [to_map(t) for t in toks[1:]]


class DecoderBlock(nn.Module):
    """ConvTranspose2d(C -> C/2) | skip_conv 1x1 (768 -> C/2) + bilinear resize -> cat -> two conv3x3/BN/ReLU."""

    def __init__(self, in_channels, in_channels_skip, out_channels):
        super().__init__()
        self.upsample = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
        self.skip_conv = nn.Conv2d(in_channels_skip, in_channels // 2, kernel_size=1)
        self.conv_block = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True))

    forward = _holder_forward


class UNetDecoder(nn.Module):
    def __init__(self, encoder_hidden_dim, decoder_channels):
        super().__init__()
        self.init_conv = nn.Conv2d(encoder_hidden_dim, decoder_channels[0], kernel_size=1)
        self.decoder_blocks = nn.ModuleList()
        in_channels = decoder_channels[0]
        for i in range(len(decoder_channels) - 1):
            out_ch = decoder_channels[i + 1]
            self.decoder_blocks.append(DecoderBlock(in_channels=in_channels, in_channels_skip=encoder_hidden_dim,
                                                    out_channels=out_ch))
            in_channels = out_ch

    forward = _holder_forward


class _DecoderHost(_EngineModule):
    """The trainable part of ClipUNet as one engine-backed module (decoder + output layer share one plan)."""


class ClipUNet(_EngineModule):
    """``forward(x)``: x [N,3,224,224] fp32 on a CUDA device -> logits [N, num_classes, 224, 224] fp32."""

    def __init__(self, num_classes=4, decoder_channels=[1024, 512, 256, 128, 64], freeze_encoder=True,
                 model_name="openai/clip-vit-base-patch16", skip_indices=[3, 5, 7, 9], clip_vit=None):
        super().__init__()
        self.encoder = ClipViTEncoder(model_name=model_name, freeze_encoder=freeze_encoder, skip_indices=skip_indices,
                                      clip_vit=clip_vit)
        self.decoder = UNetDecoder(encoder_hidden_dim=self.encoder.hidden_dim, decoder_channels=decoder_channels)
        self.output_layer = nn.Conv2d(decoder_channels[-1], num_classes, kernel_size=1)
        self.precision = "bf16"
        self.conv_algo = "auto"
        self.vit_autocast_dtype = None     # e.g. torch.bfloat16 to run the frozen ViT under autocast (the reference: fp32)
        self._engine = None

    def _build(self, plan: NetPlan, *tokens):
        grid = self.encoder.grid_size
        blocks = list(self.decoder.decoder_blocks)
        if len(tokens) - 1 < len(blocks):
            raise ValueError(f"{len(blocks)} decoder blocks need {len(blocks)} skip maps, got {len(tokens) - 1}")
        acts = [plan.token_input(i, t, grid) for i, t in enumerate(tokens)]
        x = plan.add(Conv1x1(plan, "decoder.init_conv", self.decoder.init_conv, acts[0])).out
        skips = list(reversed(acts[1:]))                      # zip(decoder_blocks, reversed(skips)), clipunet.py:141
        size = grid
        for bi, blk in enumerate(blocks):
            name = f"decoder.block{bi}"
            size *= 2
            half = blk.upsample.out_channels
            cat = plan.cat(size, size, [half, blk.skip_conv.out_channels], name=name + ".cat")
            plan.conv_transpose(name + ".upsample", blk.upsample, x, out=cat.parts[0], end_block=True)
            sk = plan.add(Conv1x1(plan, name + ".skip_conv", blk.skip_conv, skips[bi])).out
            plan.add(BilinearUp(name + ".interp", sk, cat.parts[1]))
            cb = blk.conv_block
            l1 = plan.conv_bn_relu(name + ".c1", cb[0], cb[1], cat)
            l2 = plan.conv_bn_relu(name + ".c2", cb[3], cb[4], l1.out)
            x = l2.out
        plan.head_1x1("output_layer", self.output_layer, x)

    def forward(self, x):
        from .. import _lib as L
        L.require_cuda(x)
        if any(p.requires_grad for p in self.encoder.clip_vit.parameters()):
            raise NotImplementedError("freeze_encoder=False (gradients into the ViT) is not on the accelerated path")
        with torch.no_grad():
            if self.vit_autocast_dtype is not None:
                with torch.autocast("cuda", dtype=self.vit_autocast_dtype):
                    toks = self.encoder.tokens(x)
            else:
                toks = self.encoder.tokens(x)
        toks = [t.float() for t in toks]
        if self._engine is None:
            self._engine = Engine(self, self._build)
        return self._engine.run(*toks)
