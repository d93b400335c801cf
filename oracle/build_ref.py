"""Recipe for oracle/_ref: the reference's OWN implementation of the sampling path, staged where the GPU box can see it.

TEST INFRASTRUCTURE ONLY.  The reference is pure Python on PyTorch (nothing to compile): the four files that make
up its sampling path are copied, unmodified, from the read-only checkout into `oracle/_ref/` (git-ignored, so the
history stays free of reference sources; NOT gpurun-ignored, so the directory travels to the GPU box like a built
.so).  `bench.py --impl reference` and its `cpu_baseline` leg import them from there (`kind: "reference"`) and fall
back to the oracle restatement (`kind: "port"`) when the directory is absent.

    python oracle/build_ref.py            # run by __graft_entry__.build() when /root/reference exists
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("EO_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
# the import closure of EODiffusion / UNetModel / DDIMSampler; both packages are namespace
# packages in the reference (no __init__.py)
FILES = ["backbones/unet_openai.py", "diffusion/model.py", "diffusion/ddim.py", "diffusion/util.py"]


def build(verbose: bool = True) -> bool:
    if not os.path.isdir(REF):
        if verbose:
            print(f"[oracle/_ref] {REF} not present: keeping whatever oracle/_ref already holds")
        return os.path.isdir(DST)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "sha256": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"[oracle/_ref] staged {len(FILES)} reference files under {DST}")
    return True


def import_reference():
    """(EODiffusion, UNetModel, DDIMSampler, module of diffusion.model) of the staged reference, or None."""
    if not os.path.isfile(os.path.join(DST, "diffusion", "model.py")):
        return None
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import diffusion.model as ref_model_mod
    from backbones.unet_openai import UNetModel
    from diffusion.ddim import DDIMSampler
    return ref_model_mod.EODiffusion, UNetModel, DDIMSampler, ref_model_mod


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
