"""Stage the UNMODIFIED reference Python sources under oracle/_ref/ (TEST / BENCH INFRASTRUCTURE ONLY).

``/root/reference`` exists only in the authoring container.  ``oracle/_ref/`` is git-ignored (no reference source
enters this repository's history) but NOT gpurun-ignored, so -- exactly like the built ``libunetk.so`` -- the staged
files travel to the GPU box, where ``bench.py --impl reference``, ``cpu_baseline`` and ``tools/incumbent_bench.py``
time the reference's own modules instead of the oracle port.  Nothing under ``image_segmentation_b200/`` reads it.

    python -m oracle.build_ref        # also run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import os
import shutil

SRC = os.environ.get("UNET_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = (
    "unet/__init__.py", "unet/unet.py",
    "utils/weighted_loss.py", "utils/MetricsHistory.py", "utils/training.py", "utils/utils.py", "utils/dataset.py",
    "autoencoder/__init__.py", "autoencoder/autoencoder.py",
    "clip/clipunet.py", "clip/clipunet_noskips.py",
    "prompt_based/__init__.py", "prompt_based/prompt.py",
)


def stage(force: bool = False) -> str | None:
    """Byte-for-byte copies; returns the destination or None when the reference is not present."""
    if not os.path.isfile(os.path.join(SRC, "unet", "unet.py")):
        return DST if os.path.isfile(os.path.join(DST, "unet", "unet.py")) else None
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if force or not os.path.isfile(d) or open(s, "rb").read() != open(d, "rb").read():
            shutil.copyfile(s, d)
    return DST


if __name__ == "__main__":
    print(stage(force=True))
