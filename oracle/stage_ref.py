"""TEST / BASELINE INFRASTRUCTURE ONLY -- stage the reference's own UNet + sampler sources for the GPU box.

`/root/reference` exists only in the build container.  So that `bench.py --impl reference` (CPU) and the
`reference_cuda` leg of the native bench line (the reference's eager-PyTorch CUDA path, the thing the north star's
">= 1.8x" is measured against) run the GENUINE reference classes on the GPU box, this copies the nine files those
classes need, unmodified, into the git-ignored `oracle/_ref/reference/` (same relative paths).  That directory is
listed in .gitignore (never part of the history) but not in .gpurunignore (it travels with the snapshot, like the
built .so).  Nothing under videomv_b200/ imports it; oracle/ref_import.py falls back to it when /root/reference is
absent.  Run by __graft_entry__.build() when the reference tree is present.
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("VIDEOMV_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref", "reference")
FILES = [
    "tools/modules/unet/util.py", "tools/modules/unet/unet_t2v.py", "tools/modules/unet/unet_i2vgen.py",
    "tools/modules/diffusions/diffusion_ddim.py", "tools/modules/diffusions/schedules.py",
    "tools/modules/diffusions/losses.py", "tools/modules/autoencoder.py", "utils/registry.py", "utils/registry_class.py",
]


def stage() -> bool:
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        return False
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
    open(os.path.join(DST, "utils", "__init__.py"), "w").close()     # the only generated file: an empty package marker
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference tree not found", DST)
