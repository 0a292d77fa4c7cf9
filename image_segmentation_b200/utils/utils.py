"""Batch helpers used by the train / eval loops (reference: utils/utils.py:13-115).

Only what the U-Net path touches is mirrored: aspect-preserving resize + zero padding to a square
``target_size`` (host-side input preparation, exactly where the reference does it) and its inverse.  At the
training resolution (inputs already ``target_size x target_size``) the forward transform is the identity and is
skipped entirely, which removes the reference's per-image Python loop from the hot loop.  The inverse
(``process_batch_reverse``: crop + resize of the network output back to every image's original size) runs on the
GPU as ONE launch of ``unetk_crop_resize`` for the whole ragged batch; it has no CPU path.
"""
from typing import List

import torch
import torch.nn.functional as F

from .. import _lib as L


def _interp_name(interpolation) -> str:
    name = getattr(interpolation, "value", interpolation)
    return str(name).lower()


def resize_with_padding(image, target_size=512, interpolation="bilinear"):
    """(C,H,W) -> (C,target,target): longer side scaled to target_size, centred, zero padded."""
    _, orig_h, orig_w = image.shape
    scale = min(target_size / orig_w, target_size / orig_h)
    new_w, new_h = int(round(orig_w * scale)), int(round(orig_h * scale))
    if (new_h, new_w) == (orig_h, orig_w):
        resized = image
    else:
        mode = _interp_name(interpolation)
        src = image.unsqueeze(0)
        was_int = not src.is_floating_point()
        if mode == "nearest":
            resized = F.interpolate(src.float(), size=(new_h, new_w), mode="nearest")
        else:
            resized = F.interpolate(src.float(), size=(new_h, new_w), mode=mode, align_corners=False, antialias=True)
        if was_int:
            resized = resized.round().to(image.dtype)
        resized = resized.squeeze(0)
    pad_w, pad_h = target_size - new_w, target_size - new_h
    left, top = pad_w // 2, pad_h // 2
    padded = F.pad(resized, (left, pad_w - left, top, pad_h - top), value=0)
    meta = {"original_size": (orig_h, orig_w), "new_size": (new_h, new_w),
            "pad": (left, top, pad_w - left, pad_h - top), "scale": scale}
    return padded, meta


def reverse_resize_and_padding(image, meta, interpolation="bilinear"):
    """Crop the padding away and resize (C,target,target) back to the original size (utils/utils.py:51-75)."""
    return process_batch_reverse(image.unsqueeze(0), [meta], interpolation)[0]


def process_batch_forward(batch_images, target_size=512, interpolation="bilinear"):
    """Batch (tensor or list of (C,H,W)) -> (N,C,target,target) + per-image metadata."""
    if torch.is_tensor(batch_images) and batch_images.dim() == 4 and batch_images.shape[1] != 4 \
            and batch_images.shape[-2:] == (target_size, target_size):
        n = batch_images.shape[0]
        meta = {"original_size": (target_size, target_size), "new_size": (target_size, target_size),
                "pad": (0, 0, 0, 0), "scale": 1.0}
        return batch_images, [meta] * n          # identity at the training resolution
    out, metas = [], []
    for image in batch_images:
        if image.ndim == 3 and image.shape[0] == 4:
            image = image[:3, ...]               # RGBA -> RGB as in the reference
        r, m = resize_with_padding(image, target_size, interpolation)
        out.append(r)
        metas.append(m)
    return torch.stack(out), metas


_MODES = {"bilinear": 0, "nearest": 1}


def process_batch_reverse(batch_outputs, meta_list, interpolation="bilinear") -> List[torch.Tensor]:
    """(N,C,target,target) network outputs -> list of (C,orig_h,orig_w) tensors (utils/utils.py:101-115).

    One CUDA launch for the whole batch; the returned tensors are views of one packed buffer."""
    L.require_cuda(batch_outputs)
    mode = _interp_name(interpolation)
    if mode not in _MODES:
        raise NotImplementedError(f"interpolation {interpolation!r}: only 'bilinear' and 'nearest' are implemented")
    if batch_outputs.dim() != 4 or len(meta_list) != batch_outputs.shape[0]:
        raise ValueError(f"expected (N,C,H,W) outputs with one meta per image, got {tuple(batch_outputs.shape)}")
    src = batch_outputs.detach()
    src = src if (src.dtype == torch.float32 and src.is_contiguous()) else src.float().contiguous()
    n, c, th, tw = src.shape
    for m in meta_list:
        left, top, _, _ = m["pad"]
        nh, nw = m["new_size"]
        if not (0 <= top and 0 <= left and nh > 0 and nw > 0 and top + nh <= th and left + nw <= tw):
            raise ValueError(f"meta {m} does not fit a {th}x{tw} output")
    with torch.cuda.device(src.device):
        table, total, mx = L.eval_image_table(meta_list, src.device)
        out = torch.empty(total * c, dtype=torch.float32, device=src.device)
        L.crop_resize(src, table, mx, _MODES[mode], out)
    res, off = [], 0
    for m in meta_list:
        oh, ow = m["original_size"]
        res.append(out[off * c:(off + oh * ow) * c].view(c, oh, ow))
        off += oh * ow
    return res
