"""Development tool: launch the tensor-core attention kernel a few times (for ncu).
usage: python tools/attn_run.py [B T heads ch]"""
import sys

import torch

sys.path.insert(0, ".")
from eo_diffusion_b200 import _lib  # noqa: E402

B, T, heads, ch = [int(v) for v in sys.argv[1:5]] if len(sys.argv) > 4 else (16, 4096, 8, 48)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
qkv = (torch.randn((B, T, heads * 3 * ch), generator=g) * 1.5).to(dev).to(torch.bfloat16)
out = torch.empty((B, T, heads * ch), dtype=torch.bfloat16, device=dev)
L = _lib.lib()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    e0.record()
    _lib.check(L.eo_test_attention_tc(_lib.ptr(qkv), _lib.ptr(out), B, T, heads, ch, _lib.stream_ptr()), "attn")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
print(f"attention B={B} T={T} heads={heads} ch={ch}: {ms:.3f} ms (incl. the test entry's qkv repack), "
      f"{4 * B * heads * T * T * ch / ms / 1e9:.1f} TFLOP/s algorithmic")
