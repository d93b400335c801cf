import sys, time, torch, subprocess
sys.path.insert(0, ".")
import bench
from eo_diffusion_b200 import EODiffusion
dev = torch.device("cuda:0")
wl = bench.WORKLOADS["c1"]
ctx = bench.Ctx("c1", wl, 1, "bf16", dev, 0)
cond = torch.cat([ctx.host["gt"], ctx.host["mask"]], 1).pin_memory()
m = ctx.model
evs = []
orig = m.forward
def fwd(*a, **k):
    e = torch.cuda.Event(enable_timing=True); e.record(); evs.append(e)
    return orig(*a, **k)
m.forward = fwd
def smi():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,pstate", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
def call(T):
    global evs
    evs = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    diff = EODiffusion(m, 64, 3, timesteps=T, cond_type="sum").to(dev)
    out = diff.sampling(1, device=dev, cond=cond, write_pngs=False).cpu()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    g = [a.elapsed_time(b) for a, b in zip(evs[:-1], evs[1:])]
    print(f"T={T}: total {1e3*(t3-t0):.1f} ms; GPU time between forwards: first6 {[round(x,2) for x in g[:6]]} median {sorted(g)[len(g)//2]:.2f} max {max(g):.2f} sum {sum(g):.1f} | {smi()}")
for rep in range(4):
    call(3)
    call(50)
time.sleep(2); print("idle", smi())
call(50); call(50); call(50)
