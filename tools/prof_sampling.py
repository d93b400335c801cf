"""Development aid: where does the host time of EODiffusion.sampling go at a launch-bound size?  (cProfile, B200)"""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from eo_diffusion_b200 import EODiffusion, UNetModel  # noqa: E402

size, B, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = bench.randomize_zero_init_(UNetModel(image_size=size, **bench.ARCH)).to(dev)
host, _ = bench.synth_inputs(B, size, 0, dev)
cond = torch.cat([host["gt"], host["mask"]], 1).pin_memory()


def call(T):
    diff = EODiffusion(model, size, 3, timesteps=T, cond_type="sum").to(dev)
    return diff.sampling(B, device=dev, cond=cond, write_pngs=False).cpu()


call(3)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    call(T)
    torch.cuda.synchronize()
    print(f"sampling(T={T}) {1e3 * (time.perf_counter() - t0):.1f} ms = {1e3 * (time.perf_counter() - t0) / T:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
call(T)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
