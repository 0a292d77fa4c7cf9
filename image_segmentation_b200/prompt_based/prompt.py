"""Drop-in for the reference's ``prompt_based/prompt.py``: ``PromptModel`` = frozen ``ClipUNet`` + ``unet(4, 1)``
selection network + probability composition (prompt.py:33-56).

Both networks run on the launch-plan engine; the tail -- softmax over the four CLIP classes, sigmoid of the mask logit,
``final = [1 - m, m p0 + m p3, m p1, m p2]`` (:36-56, six ATen ops and two [N,4,H,W] temporaries in the reference) -- is
one kernel forward and one backward (``unetk_prompt_compose_*``).  The CLIP branch is frozen (:30-31), so the backward
pass returns a gradient for the mask logits only.
"""
import torch
from torch import nn

from .. import _lib as L
from ..clip.clipunet import ClipUNet
from ..unet.unet import unet


class _PromptCompose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, clip_logit, mask_logit):
        out = torch.empty_like(clip_logit)
        with torch.cuda.device(clip_logit.device):
            L.prompt_compose_fwd(clip_logit, mask_logit, out)
        ctx.save_for_backward(clip_logit, mask_logit)
        return out

    @staticmethod
    def backward(ctx, g):
        clip_logit, mask_logit = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("the CLIP branch of PromptModel is frozen (prompt_based/prompt.py:30-31)")
        g = g.contiguous()
        dmask = torch.empty_like(mask_logit)
        with torch.cuda.device(g.device):
            L.prompt_compose_bwd(clip_logit, mask_logit, g, dmask)
        return None, dmask


class PromptModel(nn.Module):
    """``forward(x, heatmap)``: x [N,3,H,W], heatmap [N,1,H,W] (point-prompt Gaussian) -> class probabilities [N,4,H,W]
    (0 = not selected, 1 = background(+boundary), 2 = cat, 3 = dog).

    Args:
        path (str, optional): checkpoint of a trained ClipUNet (``checkpoint["model_state_dict"]``).
        clip (extra, optional): an already constructed ``ClipUNet`` (e.g. with random-init ViT weights where
            ``from_pretrained`` has no network); by default ``ClipUNet()`` is built as in the reference."""

    def __init__(self, path=None, clip=None):
        super().__init__()
        self.clip = clip if clip is not None else ClipUNet()
        self.mask = unet(4, 1)
        self.softmax = nn.Softmax(dim=1)
        self.sigmoid = nn.Sigmoid()
        if path is not None:
            try:
                checkpoint = torch.load(path, weights_only=False, map_location=lambda storage, loc: storage)
                self.clip.load_state_dict(checkpoint["model_state_dict"])
            except Exception as e:
                print(f"Error loading checkpoint: {str(e)[:200]}")
                raise
        for param in self.clip.parameters():
            param.requires_grad = False

    def forward(self, x, heatmap):
        L.require_cuda(x, heatmap)
        clip_logit = self.clip(x).detach()
        mask_logit = self.mask(torch.concat([x, heatmap], dim=1))
        if clip_logit.shape[1] != 4 or mask_logit.shape[1] != 1:
            raise RuntimeError("PromptModel composes 4 CLIP classes with a single mask channel")
        return _PromptCompose.apply(clip_logit.contiguous(), mask_logit.contiguous())
