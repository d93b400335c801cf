"""Development tool: per-CTA phase totals of the persistent tensor-core convolution (tc_conv3.cu) on a B200.
Slots (eo_debug_conv_trace): 0 lifetime, 1 MMA warp waits on operands, 2 MMA warp waits on a free accumulator
(= the epilogue is behind), 3 epilogue waits on the accumulator (= the main loop is behind), 4 epilogue busy,
5 producer waits on free stages, 6 tiles, 7 transform warps busy.  All in clk, printed per tile.
(the trace / EO_TEST_* switches need a -DEO_DEVTOOLS build loaded through EO_B200_LIB)
usage: [EO_TEST_GN=2] [EO_TEST_STATS=1] python tools/conv3_trace.py [B H W Cin Cout k res]"""
import math
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from eo_diffusion_b200 import _lib  # noqa: E402


def run(B, H, W, Cin, Cout, k, res):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    x = torch.randn((B, H, W, Cin), generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn((Cout, Cin, k, k), generator=g) / math.sqrt(Cin * k * k)).to(dev)
    b = torch.randn((Cout,), generator=g).to(dev)
    r = torch.randn((B, H, W, Cout), generator=g).to(dev).to(torch.bfloat16) if res else None
    y = torch.empty((B, H, W, Cout), dtype=torch.bfloat16, device=dev)
    L = _lib.lib()
    call = lambda: _lib.check(L.eo_test_conv_tc(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(r), _lib.ptr(y),
                                                B, H, W, Cin, Cout, k, _lib.stream_ptr()), "conv")
    call()
    torch.cuda.synchronize()
    tr = torch.zeros((4 * 148, 8), dtype=torch.int64, device=dev)
    L.eo_debug_conv_trace(_lib.ptr(tr), 148)
    call()
    torch.cuda.synchronize()
    L.eo_debug_conv_trace(None, 0)
    full = tr.cpu().numpy()
    t, ext, ext2 = full[:148], full[148:296], full[296:].reshape(2, 148, 8)
    lead = t[t[:, 6] > 0]                       # leader CTAs carry the MMA warp's tile count
    tiles = np.median(lead[:, 6])
    f = lambda a, c: np.median(a[:, c]) / tiles
    allc = t[t[:, 0] > 0]
    flops = 2.0 * B * H * W * Cout * Cin * k * k
    print(f"conv B={B} {H}x{W} {Cin}->{Cout} k={k} res={res}: {tiles:.0f} tiles per CTA pair, {f(allc, 0):.0f} clk per tile "
          f"(MMA floor {Cin * k * k / 64 * (256 if Cout % 256 else 512) / (1 if Cout % 256 else 1):.0f})\n"
          f"   per tile: MMA waits operands {f(lead, 1):.0f}, MMA waits accumulator {f(lead, 2):.0f}; epilogue waits accumulator "
          f"{f(allc, 3):.0f}, epilogue busy {f(allc, 4):.0f}; producer waits stages {f(lead, 5):.0f}; transform busy {f(allc, 7):.0f}")
    if ext.any():
        el = ext[t[:, 6] > 0]
        print(f"   MMA waits on operand A alone {np.median(el[:, 0]) / tiles:.0f}; transform warps wait for a free stage "
              f"{np.median(ext[t[:, 0] > 0][:, 1]) / tiles:.0f}, for the next patch's loads {np.median(ext[t[:, 0] > 0][:, 2]) / tiles:.0f}, before the stage wait {np.median(ext[t[:, 0] > 0][:, 3]) / tiles:.0f}")
        ea = ext[t[:, 0] > 0]
        print(f"   epilogue warp 0: tcgen05.ld + wait {np.median(ea[:, 4]) / tiles:.0f}, bias/residual/pack/stage {np.median(ea[:, 5]) / tiles:.0f}, "
              f"statistics read-back {np.median(ea[:, 6]) / tiles:.0f}, combine + barriers {np.median(ea[:, 7]) / tiles:.0f}")


    if ext2.any():
        live = t[:, 0] > 0
        f2 = np.concatenate([ext2[0][live], ext2[1][live]], axis=1)
        v = [np.median(f2[:, i]) / tiles for i in range(10)]
        print("   transform warp 0 per tile: advance + entry read %.0f, patch-load issue %.0f, pixel chunks %s, fence %.0f, arrive %.0f"
              % (v[0], v[1], " ".join("%.0f" % x for x in v[2:8]), v[8], v[9]))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        a = [int(v) for v in sys.argv[1:8]]
        run(*a[:6], bool(a[6]) if len(a) > 6 else False)
    else:
        run(64, 256, 256, 128, 128, 3, False)
        run(64, 256, 256, 128, 128, 3, True)
        run(64, 256, 256, 64, 128, 1, False)
        run(64, 64, 64, 384, 384, 1, True)
        run(64, 128, 128, 256, 256, 3, False)
