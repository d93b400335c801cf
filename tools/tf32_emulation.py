#!/usr/bin/env python
"""Would a TF32 tensor-core mode meet the fp32 bound (eps relative L2 <= 1e-4, BASELINE.json)?  CPU emulation: the
oracle's UNet forward with the operands of every contraction (conv, linear, attention einsum) rounded to TF32
(10-bit mantissa, round to nearest even), fp32 accumulation -- the arithmetic of tcgen05.mma kind::tf32.
TEST INFRASTRUCTURE (uses oracle/).  Result quoted in DESIGN.md section 2."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import build_unet, golden, golden_cfg, tt      # noqa: E402
from oracle import oracle as O                                # noqa: E402


def tf32(x):
    i = x.contiguous().view(torch.int32)
    r = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    return r.view(torch.float32)


def main():
    g = golden("base64_eps_t500")
    cfg = golden_cfg(g)
    m = build_unet(cfg, int(g["init_seed"]), int(g["dezero_seed"]))
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    x, t = tt(g["x"]), tt(g["t"])
    ref = O.unet_forward(sd, O.full_cfg(**cfg), x, t)
    o_conv2d, o_conv1d, o_linear, o_einsum = F.conv2d, F.conv1d, F.linear, torch.einsum
    F.conv2d = lambda a, w, b=None, **k: o_conv2d(tf32(a), tf32(w), b, **k)
    F.conv1d = lambda a, w, b=None, **k: o_conv1d(tf32(a), tf32(w), b, **k)
    F.linear = lambda a, w, b=None: o_linear(tf32(a), tf32(w), b)
    torch.einsum = lambda eq, a, b: o_einsum(eq, tf32(a), tf32(b))
    try:
        got = O.unet_forward(sd, O.full_cfg(**cfg), x, t)
    finally:
        F.conv2d, F.conv1d, F.linear, torch.einsum = o_conv2d, o_conv1d, o_linear, o_einsum
    print(f"oracle vs golden (fp32): {O.rel_l2(ref, tt(g['eps'])):.3e}")
    print(f"TF32-operand emulation vs fp32, base arch 64x64 t=500: eps relative L2 {O.rel_l2(got, ref):.3e} (bound 1e-4)")


if __name__ == "__main__":
    main()
