"""Development tool: the tensor-core convolution (C-ABI self-test entry) against torch conv2d on the
same bf16-rounded operands, over shapes of the BASELINE UNet, plus a per-CTA counter summary of the
persistent kernel (eo_debug_conv_trace).  (the trace / EO_TEST_* switches need a -DEO_DEVTOOLS build loaded through EO_B200_LIB)
usage: python tools/conv_check.py [trace]"""
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from eo_diffusion_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
L = _lib.lib()


def case(B, H, W, Cin, Cout, k, res, seed=0, trace=False):
    g = torch.Generator().manual_seed(seed + B + H + Cin + Cout)
    x = torch.randn((B, H, W, Cin), generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn((Cout, Cin, k, k), generator=g) / math.sqrt(Cin * k * k)).to(dev)
    b = torch.randn((Cout,), generator=g).to(dev)
    r = torch.randn((B, H, W, Cout), generator=g).to(dev).to(torch.bfloat16) if res else None
    y = torch.empty((B, H, W, Cout), dtype=torch.bfloat16, device=dev)
    tr = None
    if trace:
        tr = torch.zeros((296, 8), dtype=torch.int64, device=dev)
        L.eo_debug_conv_trace(_lib.ptr(tr), 296)
    _lib.check(L.eo_test_conv_tc(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(r), _lib.ptr(y),
                                 B, H, W, Cin, Cout, k, _lib.stream_ptr()), "conv")
    torch.cuda.synchronize()
    if trace:
        L.eo_debug_conv_trace(None, 0)
    nb = max(1, min(B, 4))          # reference on a few images only (fp32 conv of 64 x 256^2 is slow)
    gn_silu = os.environ.get("EO_TEST_GN", "") == "2" and k == 3 and H % 16 == 0 and W % 8 == 0 and Cin % 64 == 0
    act = (lambda v: F.silu(v).to(torch.bfloat16).float()) if gn_silu else (lambda v: v)
    want = F.conv2d(act(x[:nb].float()).permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), b, padding=k // 2)
    if res:
        want = want + r[:nb].float().permute(0, 3, 1, 2)
    got = y[:nb].float().permute(0, 3, 1, 2)
    err = float((got - want).norm() / want.norm())
    last = y[B - 1].float().permute(2, 0, 1)
    want_last = F.conv2d(act(x[B - 1:].float()).permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), b, padding=k // 2)[0]
    if res:
        want_last = want_last + r[B - 1].float().permute(2, 0, 1)
    err_last = float((last - want_last).norm() / want_last.norm())
    msg = f"B={B:3d} {H:3d}x{W:<3d} {Cin:4d}->{Cout:<4d} k={k} res={int(res)}  rel-L2 {err:.2e} (last image {err_last:.2e})"
    if trace:
        t = tr.cpu().numpy()
        t = t[t[:, 0] != 0]
        lead = t[t[:, 6] != 0]
        if len(lead):
            life = np.median(t[:, 0])
            flops = 2.0 * B * H * W * Cout * Cin * k * k
            msg += (f"\n      CTA life {life:.0f} clk, tiles/CTA {np.median(lead[:, 6]):.0f}; MMA waits: operands "
                    f"{100 * np.median(lead[:, 1]) / life:.0f}% accumulator {100 * np.median(lead[:, 2]) / life:.0f}%; "
                    f"epilogue: waits {100 * np.median(t[:, 3]) / life:.0f}% busy {100 * np.median(t[:, 4]) / life:.0f}%; "
                    f"producer waits {100 * np.median(t[:, 5]) / life:.0f}%; transform busy {100 * np.median(t[:, 7]) / life:.0f}%; "
                    f"{flops / life / 148 :.0f} flop/clk/SM")
    print(msg, flush=True)
    return max(err, err_last)


if __name__ == "__main__":
    trace = len(sys.argv) > 1 and sys.argv[1] == "trace"
    cases = [
        (1, 16, 16, 64, 128, 1, False), (2, 16, 16, 128, 128, 3, False), (2, 8, 8, 192, 192, 3, True),
        (3, 8, 8, 64, 64, 3, False), (1, 32, 32, 256, 256, 3, True), (1, 64, 64, 128, 384, 3, False),
        (2, 32, 32, 1024, 512, 3, False), (3, 16, 8, 64, 64, 3, True), (5, 32, 32, 64, 128, 3, True),
        (64, 32, 32, 512, 512, 3, False), (64, 64, 64, 384, 384, 3, True), (64, 64, 64, 384, 1536, 1, False),
        (64, 128, 128, 256, 256, 3, False), (64, 256, 256, 128, 128, 3, False), (64, 256, 256, 128, 128, 3, True),
        (64, 256, 256, 384, 128, 3, False),
    ]
    worst = 0.0
    for c in cases:
        worst = max(worst, case(*c, trace=trace))
    print("worst rel-L2", worst)
    sys.exit(0 if worst < 4e-3 else 1)
