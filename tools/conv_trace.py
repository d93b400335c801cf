"""Development tool: per-CTA phase timeline of the tensor-core convolution kernel on a B200.

Runs one convolution through the C-ABI self-test entry with eo_debug_conv_trace armed and prints,
over the traced CTAs, the median / p90 clock counts of each phase:
  setup      entry -> barriers initialised, TMEM allocated, cluster synchronised
  first_ops  setup done -> first K block's operands landed (MMA warp passes full[0])
  mainloop   first operands -> accumulator complete (epilogue warps pass tmem_full)
  epilogue   accumulator complete -> epilogue warp 2 done
  exit       epilogue done -> CTA leaves (final cluster sync + TMEM dealloc)
usage: python tools/conv_trace.py [B H W Cin Cout k res]
"""
import math
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from eo_diffusion_b200 import _lib  # noqa: E402


def run(B, H, W, Cin, Cout, k, res, n_trace=40000):
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
    tr = torch.zeros((n_trace, 8), dtype=torch.int64, device=dev)
    L.eo_debug_conv_trace(_lib.ptr(tr), n_trace)
    call()
    torch.cuda.synchronize()
    L.eo_debug_conv_trace(None, 0)
    t = tr.cpu().numpy()
    t = t[t[:, 0] != 0]
    ph = {
        "setup": t[:, 2] - t[:, 1],
        "first_ops": t[:, 3] - t[:, 2],
        "mainloop": t[:, 4] - t[:, 3],
        "epilogue": t[:, 5] - t[:, 4],
        "exit": t[:, 6] - t[:, 5],
        "lifetime": t[:, 6] - t[:, 1],
    }
    lead = t[:, 3] != 0      # only the leader CTA of a pair has an MMA warp stamp
    print(f"conv B={B} {H}x{W} {Cin}->{Cout} k={k} res={res}: {len(t)} CTAs traced, "
          f"wall {(t[:, 0].max() - t[:, 0].min()) / 1e3:.1f} us between first and last traced entry")
    for name, v in ph.items():
        vv = v[lead] if name in ("first_ops", "mainloop") else v
        print(f"  {name:10s} median {np.median(vv):9.0f}  p10 {np.percentile(vv, 10):9.0f}  p90 {np.percentile(vv, 90):9.0f} clk")
    # CTAs per SM and the entry-to-entry interval on one SM
    sm = t[:, 7]
    one = np.sort(t[sm == sm[0], 0])
    if len(one) > 2:
        print(f"  SM {sm[0]}: {len(one)} traced CTAs, median entry-to-entry {np.median(np.diff(one)):.0f} ns")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        a = [int(v) for v in sys.argv[1:8]]
        run(*a[:6], bool(a[6]) if len(a) > 6 else False)
    else:
        run(64, 256, 256, 128, 128, 3, False)
        run(64, 256, 256, 128, 128, 3, True)
        run(64, 64, 64, 384, 1152, 1, False)
        run(64, 128, 128, 256, 256, 3, False)
