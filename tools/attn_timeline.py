"""Development tool: clock64 timeline of the four softmax warps that share one warp scheduler (quadrant 0 of every query
tile) and of their MMA issuers, 16 consecutive 32-key rounds of one CTA (needs a -DEO_DEVTOOLS build, see attn_trace.py).
usage: python tools/attn_timeline.py [B T heads ch]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from eo_diffusion_b200 import _lib  # noqa: E402

B, T, heads, ch = [int(v) for v in sys.argv[1:5]] if len(sys.argv) > 4 else (8, 4096, 8, 48)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
qkv = (torch.randn((B, T, heads * 3 * ch), generator=g) * 1.5).to(dev).to(torch.bfloat16)
out = torch.empty((B, T, heads * ch), dtype=torch.bfloat16, device=dev)
L = _lib.lib()
call = lambda: _lib.check(L.eo_test_attention_tc(_lib.ptr(qkv), _lib.ptr(out), B, T, heads, ch, _lib.stream_ptr()), "attn")
call()
torch.cuda.synchronize()
n = 512
tr = torch.zeros((n + 288, 8), dtype=torch.int64, device=dev)
L.eo_debug_conv_trace(_lib.ptr(tr), n)
call()
torch.cuda.synchronize()
L.eo_debug_conv_trace(None, 0)
raw = tr[n:].cpu().numpy().reshape(-1)
# [tile][quadrant][round][top, first half done, next row requested, second half done, next row there, P V_q-1 seen,
#  P handed over, next maximum known]
sm = raw[:2048].reshape(4, 4, 16, 8)
iss = raw[2048:2304].reshape(4, 16, 4)    # [tile][round][round starts, next S issued, saw P, P V issued]
t0 = sm[:, :, :, 0].min()
print("softmax warp of quadrant 0: top, first half done, next row requested, second half done, next row there, PV seen, handed over, max known"
      " || issuer: round starts, next S issued, saw P, P V issued")
for r in range(16):
    for t in range(4):
        print(f"round {r:2d} tile {t}: " + " ".join(f"{int(v - t0):6d}" for v in sm[t, 0, r])
              + " || " + " ".join(f"{int(v - t0):6d}" for v in iss[t, r]))
print("round length per tile (clk):", [float(np.mean(np.diff(sm[t, 0, :, 0]))) for t in range(4)])
seg = np.diff(sm, axis=3).reshape(-1, 7).mean(axis=0)
print("softmax segments, mean: first half %.0f, wait for S + ld issue %.0f, second half %.0f, wait::ld + s_free %.0f, wait for PV %.0f, "
      "st + hand-over %.0f, maximum + vote %.0f; loop back %.0f" % (*seg, float(np.mean(sm[:, :, 1:, 0] - sm[:, :, :-1, 7]))))
last_arrive = sm[:, :, :, 6].max(axis=1)
print("spread of the four quadrants' hand-over (last - first), mean:", float((last_arrive - sm[:, :, :, 6].min(axis=1)).mean()))
print("last hand-over -> issuer saw P, mean:", float((iss[:, :, 2] - last_arrive).mean()))
print("issuer: S issue after round start %.0f, P V issue %.0f" % (float((iss[:, :, 1] - iss[:, :, 0]).mean()), float((iss[:, :, 3] - iss[:, :, 2]).mean())))
