"""Development tool: launch one tensor-core convolution a few times (for ncu).
(the trace / EO_TEST_* switches need a -DEO_DEVTOOLS build loaded through EO_B200_LIB)
usage: [EO_TEST_GN=2] [EO_TEST_STATS=1] python tools/conv_run.py [B H W Cin Cout k res]"""
import math
import sys

import torch

sys.path.insert(0, ".")
from eo_diffusion_b200 import _lib  # noqa: E402

a = [int(v) for v in sys.argv[1:8]] if len(sys.argv) > 6 else [64, 256, 256, 128, 128, 3, 0]
B, H, W, Cin, Cout, k, res = a + [0] * (7 - len(a))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
x = torch.randn((B, H, W, Cin), generator=g).to(dev).to(torch.bfloat16)
w = (torch.randn((Cout, Cin, k, k), generator=g) / math.sqrt(Cin * k * k)).to(dev)
b = torch.randn((Cout,), generator=g).to(dev)
r = torch.randn((B, H, W, Cout), generator=g).to(dev).to(torch.bfloat16) if res else None
y = torch.empty((B, H, W, Cout), dtype=torch.bfloat16, device=dev)
L = _lib.lib()
for _ in range(3):
    _lib.check(L.eo_test_conv_tc(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(r), _lib.ptr(y), B, H, W, Cin, Cout, k,
                                 _lib.stream_ptr()), "conv")
torch.cuda.synchronize()
print("ok")
