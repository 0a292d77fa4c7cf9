"""Import the UNMODIFIED reference modules from /root/reference (TEST INFRASTRUCTURE ONLY).

Used only in the authoring container (the GPU box has no /root/reference): by
``tests/golden/make_golden.py`` to generate the committed golden vectors and by the
``not gpu`` tests that compare the oracle with the live reference when it is present.

The reference imports two packages that are not installed here; they are replaced by empty
shims *before* import, nothing in the reference is edited:

* ``matplotlib`` / ``matplotlib.pyplot`` (utils/utils.py:3 imports it, never used on this path)
* ``tqdm.notebook.tqdm`` (utils/training.py:6; needs ipywidgets) -> pass-through iterator with a
  no-op ``set_postfix`` (utils/training.py:60 calls it)
"""
from __future__ import annotations

import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/build_ref.py (git-ignored copies)


def _default_root() -> str:
    for cand in (os.environ.get("UNET_REFERENCE_ROOT"), "/root/reference", _STAGED):
        if cand and os.path.isfile(os.path.join(cand, "unet", "unet.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _default_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "unet", "unet.py"))


class _PassThroughTqdm:
    def __init__(self, iterable=None, *a, **k):
        self._it = iterable

    def __iter__(self):
        return iter(self._it)

    def set_postfix(self, *a, **k):
        pass

    def update(self, *a, **k):
        pass

    def close(self):
        pass


def _install_shims():
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            m = types.ModuleType("matplotlib")
            p = types.ModuleType("matplotlib.pyplot")
            m.pyplot = p
            sys.modules["matplotlib"] = m
            sys.modules["matplotlib.pyplot"] = p
    import tqdm.notebook as tn
    tn.tqdm = _PassThroughTqdm


_REF_PACKAGES = ("utils", "unet", "autoencoder", "clip", "prompt_based")


def _is_reference_module(name: str) -> bool:
    return any(name == p or name.startswith(p + ".") for p in _REF_PACKAGES)


def random_init_clip_vit(seed: int = 0):
    """The reference builds its encoder with ``CLIPVisionModel.from_pretrained("openai/clip-vit-base-patch16")``
    (clip/clipunet.py:25-26), which needs the network.  BASELINE.json config 4 asks for RANDOM-INIT weights of that
    architecture: patch ``from_pretrained`` of both classes to build ViT-B/16 (hidden 768, 12 layers, 224 px, patch 16)
    from a default config.  Returns a context manager; nothing in the reference is edited."""
    import contextlib

    import torch
    from transformers import CLIPVisionConfig, CLIPVisionModel

    @contextlib.contextmanager
    def patched():
        cfg_fp, model_fp = CLIPVisionConfig.from_pretrained, CLIPVisionModel.from_pretrained

        def cfg_from(*a, **k):
            return CLIPVisionConfig(patch_size=16)

        def model_from(*a, **k):
            state = torch.random.get_rng_state()
            torch.manual_seed(seed)
            try:
                return CLIPVisionModel(CLIPVisionConfig(patch_size=16))
            finally:
                torch.random.set_rng_state(state)
        CLIPVisionConfig.from_pretrained = staticmethod(cfg_from)
        CLIPVisionModel.from_pretrained = staticmethod(model_from)
        try:
            yield
        finally:
            CLIPVisionConfig.from_pretrained, CLIPVisionModel.from_pretrained = cfg_fp, model_fp
    return patched()


def load():
    """Returns a namespace with the reference's hot-path symbols (raises if unavailable)."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_shims()
    # the reference's top-level packages are called `unet` and `utils`; import them under the
    # reference root without leaving it on sys.path for the rest of the process
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if _is_reference_module(k)}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import importlib
        ns = types.SimpleNamespace()
        ns.unet_mod = importlib.import_module("unet.unet")
        ns.loss_mod = importlib.import_module("utils.weighted_loss")
        ns.metrics_mod = importlib.import_module("utils.MetricsHistory")
        try:
            ns.training_mod = importlib.import_module("utils.training")
        except Exception as e:  # torchvision etc. missing
            ns.training_mod = None
            ns.training_error = e
        ns.unet = ns.unet_mod.unet
        ns.WeightedDiceCELoss = ns.loss_mod.WeightedDiceCELoss
        ns.WeightedMemoryEfficientDiceLoss = ns.loss_mod.WeightedMemoryEfficientDiceLoss
        ns.MetricsHistory = ns.metrics_mod.MetricsHistory
        # the other model families; each is optional
        for attr, modname in (("autoencoder_mod", "autoencoder.autoencoder"), ("clip_mod", "clip.clipunet"),
                              ("prompt_mod", "prompt_based.prompt")):
            try:
                setattr(ns, attr, importlib.import_module(modname))
            except Exception as e:  # e.g. transformers missing
                setattr(ns, attr, None)
                setattr(ns, attr + "_error", e)
        return ns
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if _is_reference_module(k)]:
            # keep reference modules reachable only through `ns`
            sys.modules.pop(k)
        sys.modules.update(saved)
