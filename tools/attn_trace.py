"""Development tool: per-CTA clock totals of the attention kernel on a B200 (eo_debug_conv_trace; needs a -DEO_DEVTOOLS
build: EO_LIB_VARIANT=dev EO_NVCC_EXTRA=-DEO_DEVTOOLS python -m eo_diffusion_b200.build, then EO_B200_LIB=.../libeo_b200_dev.so).
usage: python tools/attn_trace.py [B T heads ch]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from eo_diffusion_b200 import _lib  # noqa: E402


def run(B, T, heads, ch):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    qkv = (torch.randn((B, T, heads * 3 * ch), generator=g) * 1.5).to(dev).to(torch.bfloat16)
    out = torch.empty((B, T, heads * ch), dtype=torch.bfloat16, device=dev)
    L = _lib.lib()
    call = lambda: _lib.check(L.eo_test_attention_tc(_lib.ptr(qkv), _lib.ptr(out), B, T, heads, ch, _lib.stream_ptr()), "attn")
    call()
    torch.cuda.synchronize()
    n = 4096
    tr = torch.zeros((n, 8), dtype=torch.int64, device=dev)
    L.eo_debug_conv_trace(_lib.ptr(tr), n)
    call()
    torch.cuda.synchronize()
    L.eo_debug_conv_trace(None, 0)
    t = tr.cpu().numpy()
    t = t[t[:, 0] != 0]
    nb = np.median(t[:, 7])
    life = np.median(t[:, 0])
    f = lambda c: np.median(t[:, c]) / nb
    print(f"attention B={B} T={T} heads={heads} ch={ch}: {len(t)} CTAs traced, {nb:.0f} half-blocks, CTA life {life:.0f} clk "
          f"= {life / nb:.0f} clk per 64 keys\n   per 64 keys: softmax warp (tile 0, quadrant 0) spins on S {f(1):.0f}; issuer 0 waits for P {f(6):.0f}"
          f"   (phase by phase: tools/attn_timeline.py)")

if __name__ == "__main__":
    if len(sys.argv) > 4:
        run(*[int(v) for v in sys.argv[1:5]])
    else:
        run(8, 4096, 8, 48)
        run(16, 1024, 8, 64)
