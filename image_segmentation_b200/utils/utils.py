"""Host-side batch helpers used by the train / eval loops (reference: utils/utils.py:13-115).

Only what the U-Net training path touches is mirrored: aspect-preserving resize + zero padding to a
square ``target_size`` and its inverse.  At the training resolution (inputs already
``target_size x target_size``) the forward transform is the identity and is skipped entirely, which
removes the reference's per-image Python loop from the hot loop.
"""
from typing import List

import torch
import torch.nn.functional as F


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
    """Crop the padding away and resize (C,target,target) back to the original size."""
    left, top, _, _ = meta["pad"]
    new_h, new_w = meta["new_size"]
    cropped = image[..., top: top + new_h, left: left + new_w]
    orig_h, orig_w = meta["original_size"]
    if (orig_h, orig_w) == (new_h, new_w):
        return cropped
    return F.interpolate(cropped.unsqueeze(0), size=(orig_h, orig_w), mode=interpolation,
                         align_corners=False if interpolation != "nearest" else None).squeeze(0)


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


def process_batch_reverse(batch_outputs, meta_list, interpolation="bilinear") -> List[torch.Tensor]:
    return [reverse_resize_and_padding(o, m, interpolation) for o, m in zip(batch_outputs, meta_list)]
