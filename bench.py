#!/usr/bin/env python
"""Benchmark of the reverse-sampling hot path (BASELINE.json metric: 256x256 cloud-removal
images/sec at T=1000 DDPM; UNet ms/step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3]

One "step" = one DDPM reverse step of the whole per-GPU batch: UNet forward (eps) + fused
posterior update + next step's 'sum' conditioning mix.  images/sec = images in flight /
(T x seconds per step), T = 1000.  Prints ONE JSON line (rank 0).

  value        steps timed on the device (CUDA events) with every input resident in HBM
  e2e          the reference's own call, `EODiffusion.sampling(n, device=..., cond=<host tensor>)` followed by
               `.cpu()`: ONE full T = 1000 trajectory of the whole per-GPU batch when that takes <= 90 s (else, or with
               --short-e2e, a --steps-long instance of the same model); host RNG draw of x_T, host -> device copies
               of x_T and cond, every per-step launch of the public loop and the device -> host read of the result
               are inside the timed region
  e2e_streamed the per-step entry points (public methods) fed from pinned host buffers with the copies of step
               k+1 overlapping the compute of step k (what a streaming caller can reach)
  roofline     the dominant kernel family (tcgen05 implicit-GEMM conv), timed live with CUDA events per op
               (eo_unet_forward_timed) against MEASURED_PEAKS.json; `achieved` counts ALGORITHMIC FLOPs
               (SURVEY.md 8d), `achieved_executed` what the launches execute (padded rows, 4/9 sub-pixel form)
  secondary    the other BASELINE.json configurations (c1, c2, c4, c5), a few device-timed steps each
  cpu_baseline the reference's CPU path on the host cores, bounded sample (oracle/_ref = the reference's own
               classes when staged, else the oracle port)

`--impl reference` times the reference's CPU implementation of the path (same sources as cpu_baseline).
"""
from __future__ import annotations

import argparse
import contextlib
import gc
import json
import math
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_DDPM = 1000
ARCH = dict(in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2,
            attention_resolutions=[4, 8], channel_mult=[1, 2, 3, 4], num_heads=8)
# BASELINE.json configs; c3 (the configuration the metric is quoted on) is the default
WORKLOADS = {
    "c1": dict(size=64, batch=1, desc="64x64 batch 1 (reference notebook case)"),
    "c2": dict(size=128, batch=16, desc="128x128 batch 16"),
    "c3": dict(size=256, batch=64, desc="256x256 batch 64 per GPU"),
    "c4": dict(size=256, batch=256, kind="ddim", steps_per_image=50,
               desc="256x256 batch 256 per GPU, DDIM S=50 eta=0 (diffusion/ddim.py)"),
    "c5": dict(size=256, batch=64, kind="concat", cx=13, cc=15,
               desc="256x256 batch 64 per GPU, 13-band S2 + 15-channel conditioning concatenated (28 -> 13 ch)"),
}
METRIC = "ddpm_T1000_cloud_removal_images_per_sec"
UNIT = "images/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def randomize_zero_init_(model, seed=4321):
    """The reference zero-initialises 35 convs, so a fresh UNet outputs 0 and
    does no representative arithmetic.  Re-draw them and perturb the GroupNorm affines."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (torch.nn.Conv1d, torch.nn.Conv2d)) and not bool(m.weight.any()):
                fan_in = m.weight[0].numel()
                bound = 1.0 / math.sqrt(fan_in)
                m.weight.copy_((torch.rand(m.weight.shape, generator=g) * 2 - 1) * bound)
                m.bias.copy_((torch.rand(m.bias.shape, generator=g) * 2 - 1) * bound)
            elif isinstance(m, torch.nn.GroupNorm):
                m.weight.copy_(1 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
    return model


def synth_inputs(n, size, seed, device, pin=False, channels=3, cond_channels=0):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand((n, 3, size, size), generator=g)
    mask = (torch.rand((n, 1, size, size), generator=g) > 0.3).float()
    x = torch.randn((n, channels, size, size), generator=g)
    nz = [torch.randn((n, channels, size, size), generator=g) for _ in range(2)]
    host = dict(gt=gt, mask=mask, x=x, nz0=nz[0], nz1=nz[1])
    if cond_channels:
        host["cond"] = torch.rand((n, cond_channels, size, size), generator=g)
    if pin:
        host = {k: v.pin_memory() for k, v in host.items()}
    dev = {k: v.to(device) for k, v in host.items()}
    return host, dev


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv = None
            log(f"[bench] NVML unavailable: {e}")

    NAMES = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
             0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def run(self):
        if self.nv is None:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.1)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tflops=float(p["bf16_tflops_sustained"]), burst=float(p["bf16_tflops"]), gbs=float(p["hbm_gbs"]),
                    src="measured")
    except Exception:
        return dict(tflops=1590.0, burst=1590.0, gbs=6650.0, src="fallback")   # B200_PROFILING.md fallback (burst)


def ncu_traffic(kernel, workload, batch):
    """DRAM bytes (read + write) per launch of `kernel`, averaged over the launches of one UNet forward,
    from the committed `ncu --set full` capture of this workload (profiles/*_traffic.json, written by
    profiles/extract_traffic.py).  None when no capture matches."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
        try:
            with open(path) as f:
                t = json.load(f)
            if t.get("kernel") == kernel and t.get("workload") == workload and t.get("batch") == batch:
                return t["dram_bytes_per_launch"]
        except Exception:
            continue
    return None


# ----------------------------------------------------------------------------------------
# CPU legs: cpu_baseline and --impl reference
# ----------------------------------------------------------------------------------------
def cpu_reference_steps(size, batch, steps, warmup):
    """Seconds per sampler step ('sum' mix + UNet + clipped posterior) of the reference's CPU path on all host
    cores, and which implementation ran: the reference's own classes staged under oracle/_ref ("reference"),
    else the oracle restatement ("port").  The reference is driven through its public call,
    EODiffusion.sampling(n, cond=...), on a (warmup + steps)-step instance; step times are the intervals between
    consecutive UNet calls (forward pre-hook), the last one closed by the call's return."""
    torch.set_num_threads(os.cpu_count() or 1)
    from oracle import build_ref                 # test infrastructure: allowed in the CPU legs only
    ref = build_ref.import_reference()
    host, _ = synth_inputs(batch, size, 0, "cpu")
    cond = torch.cat([host["gt"], host["mask"]], 1)
    if ref is not None:
        RefDiffusion, RefUNet, _, ref_mod = ref
        torch.manual_seed(1234)
        unet = randomize_zero_init_(RefUNet(image_size=size, **ARCH)).eval()
        with contextlib.redirect_stdout(sys.stderr):     # the reference prints "loading model..." / "Loaded!!": stdout carries the JSON line only
            diff = RefDiffusion(unet, size, 3, timesteps=warmup + steps, cond_type="sum").eval()
        ref_mod.save_image = lambda *a, **k: None        # sampling() writes PNGs unconditionally
        stamps = []
        h = unet.register_forward_pre_hook(lambda m, a: stamps.append(time.perf_counter()))
        with torch.no_grad(), contextlib.redirect_stdout(sys.stderr):
            diff.sampling(batch, clipped_reverse_diffusion=True, device="cpu", cond=cond)
        stamps.append(time.perf_counter())
        h.remove()
        dt = [b - a for a, b in zip(stamps[:-1], stamps[1:])][warmup:]
        return sum(dt) / len(dt), "reference"
    from oracle import oracle as O
    from eo_diffusion_b200 import UNetModel
    torch.manual_seed(1234)
    m = randomize_zero_init_(UNetModel(image_size=size, **ARCH))
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    cfg = O.full_cfg(image_size=size, **ARCH)
    s = O.cosine_schedule(T_DDPM)
    x, gt, mask = host["x"], host["gt"], host["mask"]
    times = []
    with torch.no_grad():
        for k in range(warmup + steps):
            t0 = time.perf_counter()
            t = torch.full((batch,), T_DDPM - 1 - k, dtype=torch.long)
            nz = host["nz0"] if k % 2 == 0 else host["nz1"]
            x = O.sum_mix(s, x, gt, mask, t, nz)
            eps = O.unet_forward(sd, cfg, x, t)
            x = O.reverse_step_clip(s, x, t, nz, eps)
            if k >= warmup:
                times.append(time.perf_counter() - t0)
    return sum(times) / len(times), "port"


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    size = wl["size"]
    sample_b = 1
    sec, kind = cpu_reference_steps(size, sample_b, args.steps, args.warmup)
    val = sample_b / (T_DDPM * sec)
    cores = torch.get_num_threads()
    sample = (f"{sample_b} image(s) of {size}x{size}, {args.steps} sampler steps (of T={T_DDPM}) after {args.warmup} "
              f"warm-up, " + ("the reference's own EODiffusion.sampling + UNetModel (oracle/_ref)" if kind == "reference"
                              else "oracle restatement of the reference path"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}, DDPM T={T_DDPM}, cond 'sum', clipped; "
                               f"CPU arm runs a bounded sample: {sample}"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
class Ctx:
    """One workload on one rank: the model, its inputs and the per-step driver."""

    def __init__(self, name, wl, batch, mode, dev, rank):
        import ctypes as C
        from eo_diffusion_b200 import EODiffusion, UNetModel, _lib
        self.C, self.lib, self.L = C, _lib, _lib.lib()
        self.name, self.wl, self.dev, self.mode = name, wl, dev, mode
        self.size, self.B = wl["size"], batch
        self.kind = wl.get("kind", "sum")
        self.steps_per_image = wl.get("steps_per_image", T_DDPM)
        self.cx, self.cc = wl.get("cx", 3), wl.get("cc", 0)
        arch = dict(ARCH, in_channels=self.cx + self.cc, out_channels=self.cx)
        torch.manual_seed(1234)
        self.model = randomize_zero_init_(UNetModel(image_size=self.size, **arch)).to(dev).set_compute_mode(mode)
        self.diff = EODiffusion(self.model, self.size, self.cx, timesteps=T_DDPM,
                                cond_type="sum" if self.kind == "sum" else None).to(dev)
        self.host, self.d = synth_inputs(batch, self.size, 100 + rank, dev, pin=True, channels=self.cx,
                                         cond_channels=self.cc)
        self.tab = self.diff._coef_table(dev)
        self.ts_rows = self.diff._timestep_rows(batch, dev)
        self.hw = self.size * self.size
        self.nz = [self.d["nz0"], self.d["nz1"]]
        self.x = self.d["x"].clone()
        self.sampler = None
        if self.kind == "ddim":
            from eo_diffusion_b200 import DDIMSampler
            self.sampler = DDIMSampler(self.diff)
            self.sampler.make_schedule(ddim_num_steps=self.steps_per_image, ddim_eta=0.0, verbose=False)
            self.ddim_ts = [int(v) for v in self.sampler.ddim_timesteps]
            self.scal = []
            for idx in range(len(self.ddim_ts)):       # the scalars DDIMSampler.p_sample_ddim hands to eo_ddim_step
                a_t, a_prev = float(self.sampler.ddim_alphas[idx]), float(self.sampler.ddim_alphas_prev[idx])
                sg = float(self.sampler.ddim_sigmas[idx])
                self.scal.append((math.sqrt(a_t), float(self.sampler.ddim_sqrt_one_minus_alphas[idx]), math.sqrt(a_prev),
                                  math.sqrt(max(1. - a_prev - sg * sg, 0.)), sg))
            self.pred_x0 = torch.empty_like(self.x)

    def close(self):
        self.model = self.diff = self.sampler = None
        self.host = self.d = self.x = self.nz = None
        gc.collect()
        torch.cuda.empty_cache()

    def first_mix(self):
        if self.kind == "sum":
            p, d = self.lib.ptr, self.d
            self.lib.check(self.L.eo_ddpm_sum_mix(p(self.x), p(d["gt"]), p(d["mask"]), p(self.nz[0]),
                                                  p(self.ts_rows[T_DDPM - 1]), p(self.tab), p(self.x), self.B, 3, self.hw,
                                                  self.lib.stream_ptr()))

    def device_step(self, k):
        """One sampler step with every operand resident in HBM (the timestep tables of the schedule installed, as
        EODiffusion.sampling / DDIMSampler.sample do for their loops)."""
        p, L, d, x, B = self.lib.ptr, self.L, self.d, self.x, self.B
        st = self.lib.stream_ptr
        if self.kind == "ddim":
            idx = len(self.ddim_ts) - 1 - (k % len(self.ddim_ts))
            pred = self.model(x, self.ts_rows[self.ddim_ts[idx]])
            sa, s1, sp, dc, sg = self.scal[idx]
            self.lib.check(L.eo_ddim_step(p(x), p(pred), None, p(x), p(self.pred_x0), sa, s1, sp, dc, sg, 1.0, x.numel(),
                                          st()), "eo_ddim_step")
            return
        i = T_DDPM - 1 - (k % (T_DDPM - 1))           # i >= 1
        t = self.ts_rows[i]
        if self.kind == "concat":
            pred = self.model(x, t, cond=d["cond"])
            self.lib.check(L.eo_ddpm_step(p(x), p(pred), p(self.nz[k % 2]), p(t), p(self.tab), p(x), B, self.cx, self.hw,
                                          1, 1, st()), "eo_ddpm_step")
            return
        pred = self.model(x, t)
        self.lib.check(L.eo_ddpm_step_mix(p(x), p(pred), p(self.nz[k % 2]), p(t), p(d["gt"]), p(d["mask"]),
                                          p(self.nz[(k + 1) % 2]), p(self.ts_rows[i - 1]), p(self.tab), p(x), B, 3,
                                          self.hw, 1, 1, st()), "eo_ddpm_step_mix")

    def time_device_steps(self, steps, warmup, barrier, clocks=None):
        self.first_mix()
        with self.model.time_tables(T_DDPM + 1):
            for k in range(warmup):
                self.device_step(k)
            barrier()
            if clocks is not None:
                clocks.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(steps):
                self.device_step(warmup + k)
            e1.record()
            barrier()
            launches = self.model.launches_per_forward() + 1
        assert bool(torch.isfinite(self.x).all()), "sampler state diverged"
        return e0.elapsed_time(e1) / steps, launches

    def op_table(self):
        """Per-op CUDA-event times of one warm UNet forward + the engine's FLOP notes."""
        C, L, lib = self.C, self.L, self.lib
        h = C.c_void_p(self.model._handle)
        nops = L.eo_unet_num_ops(h)
        ms = (C.c_float * nops)()
        eps = torch.empty((self.B, self.cx, self.size, self.size), device=self.dev)
        cond_p = lib.ptr(self.d["cond"]) if self.kind == "concat" else None
        for _ in range(2):   # second pass is the warm one
            lib.check(L.eo_unet_forward_timed(h, lib.ptr(self.x), self.cx, cond_p, self.cc, lib.ptr(self.ts_rows[500]),
                                              None, lib.ptr(eps), self.B, lib.stream_ptr(), ms), "eo_unet_forward_timed")
        name, kern, fl, by = C.c_char_p(), C.c_char_p(), C.c_double(), C.c_double()
        ops = []
        for i in range(nops):
            L.eo_unet_op_info(h, i, C.byref(name), C.byref(kern), C.byref(fl), C.byref(by))
            ops.append(dict(name=name.value.decode(), kernel=kern.value.decode(), ms=ms[i], flops=fl.value * self.B,
                            exec_flops=L.eo_unet_op_executed_flops(h, i) * self.B, bytes=by.value * self.B))
        return ops

    def alg_flops_per_image(self):
        C, L = self.C, self.L
        h = C.c_void_p(self.model._handle)
        fl = C.c_double()
        tot = 0.0
        for i in range(L.eo_unet_num_ops(h)):
            L.eo_unet_op_info(h, i, None, None, C.byref(fl), None)
            tot += fl.value
        return tot


def e2e_public_sampling(ctx, steps, barrier):
    """The reference's call (inference.py:121-126): EODiffusion.sampling(n, device=..., cond=<host tensor>) then the
    result on the host.  Short-T instance (UNet and step kernels do not depend on T).  Returns (ms per step, h2d, d2h)."""
    from eo_diffusion_b200 import EODiffusion
    B, size, dev = ctx.B, ctx.size, ctx.dev
    if ctx.kind == "sum":
        cond = torch.cat([ctx.host["gt"], ctx.host["mask"]], 1).pin_memory()
    elif ctx.kind == "concat":
        cond = ctx.host["cond"]
    else:
        cond = None
    h2d = B * ctx.cx * size * size * 4 + (cond.numel() * 4 if cond is not None else 0)       # x_T + cond, once per call
    d2h = B * ctx.cx * size * size * 4

    def call(T):
        if ctx.kind == "ddim":
            with torch.no_grad():
                out, _ = ctx.sampler.sample(T, B, (ctx.cx, size, size), eta=0.0, verbose=False)
            return out.cpu()
        diff = EODiffusion(ctx.model, size, ctx.cx, timesteps=T, cond_type="sum" if ctx.kind == "sum" else None).to(dev)
        return diff.sampling(B, clipped_reverse_diffusion=True, device=dev, cond=cond, write_pngs=False).cpu()

    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):      # the DDIM sampler prints like the reference does
        call(3)                                          # warm: plans, graph capture, allocator
        barrier()
        t0 = time.perf_counter()
        out = call(steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    barrier()
    assert bool(torch.isfinite(out).all())
    return dt * 1e3 / steps, h2d / steps, d2h / steps


def e2e_streamed(ctx, n_steps, barrier):
    """Per-step public methods fed from pinned host buffers: every step copies ITS inputs host -> device and its result
    device -> host inside the timed region, on side streams into double-buffered device tensors, so step k+1's inputs
    travel while step k computes."""
    lib, L, d, host, B, kind = ctx.lib, ctx.L, ctx.d, ctx.host, ctx.B, ctx.kind
    hx = host["x"].clone().pin_memory()
    out_host = torch.empty_like(hx).pin_memory()

    def slot_buffers():
        b = dict(x=torch.empty_like(d["x"]), n=torch.empty_like(d["x"]))
        if kind == "sum":
            b["gt"], b["m"] = torch.empty_like(d["gt"]), torch.empty_like(d["mask"])
        if kind == "concat":
            b["cond"] = torch.empty_like(d["cond"])
        return b

    slots = [slot_buffers(), slot_buffers()]
    if kind == "sum":
        h2d = sum(t.numel() * 4 for t in (hx, host["nz0"], host["gt"], host["mask"]))
    elif kind == "concat":
        h2d = sum(t.numel() * 4 for t in (hx, host["nz0"], host["cond"]))
    else:
        h2d = hx.numel() * 4
    d2h = out_host.numel() * 4
    main_s = torch.cuda.current_stream()
    in_s, out_s = torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event(), torch.cuda.Event()]      # slot's inputs have landed
    ev_free = [torch.cuda.Event(), torch.cuda.Event()]    # the step that read the slot has finished
    ev_done, ev_out = torch.cuda.Event(), torch.cuda.Event()

    def stage_inputs(k):
        b = slots[k % 2]
        in_s.wait_event(ev_free[k % 2])
        with torch.cuda.stream(in_s):
            b["x"].copy_(hx, non_blocking=True)
            if kind != "ddim":
                b["n"].copy_(host["nz0"] if k % 2 == 0 else host["nz1"], non_blocking=True)
            if kind == "sum":
                b["gt"].copy_(host["gt"], non_blocking=True)
                b["m"].copy_(host["mask"], non_blocking=True)
            if kind == "concat":
                b["cond"].copy_(host["cond"], non_blocking=True)
            ev_in[k % 2].record(in_s)

    def step(k, last):
        b = slots[k % 2]
        if not last:
            stage_inputs(k + 1)
        main_s.wait_event(ev_in[k % 2])
        dx, dn = b["x"], b["n"]
        if kind == "ddim":
            idx = len(ctx.ddim_ts) - 1 - (k % len(ctx.ddim_ts))
            nxt, _ = ctx.sampler.p_sample_ddim(dx, None, ctx.ts_rows[ctx.ddim_ts[idx]], index=idx)
        else:
            i = T_DDPM - 1 - (k % (T_DDPM - 1))
            t = ctx.ts_rows[i]
            if kind == "concat":
                nxt = ctx.diff._reverse_diffusion_with_clip(dx, t, dn, cond=b["cond"])
            else:
                lib.check(L.eo_ddpm_sum_mix(lib.ptr(dx), lib.ptr(b["gt"]), lib.ptr(b["m"]), lib.ptr(dn), lib.ptr(t),
                                            lib.ptr(ctx.tab), lib.ptr(dx), B, 3, ctx.hw, lib.stream_ptr()))
                nxt = ctx.diff._reverse_diffusion_with_clip(dx, t, dn)
        ev_free[k % 2].record(main_s)
        ev_done.record(main_s)
        main_s.wait_event(ev_out)              # the previous result has left before its buffer can be reused
        out_s.wait_event(ev_done)
        with torch.cuda.stream(out_s):
            out_host.copy_(nxt, non_blocking=True)
            ev_out.record(out_s)
        nxt.record_stream(out_s)

    for ev in ev_free:
        ev.record(main_s)
    ev_out.record(out_s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ctx.model.time_tables(T_DDPM + 1):
        stage_inputs(0)
        step(0, True)
        barrier()
        e0.record()
        stage_inputs(1)
        for k in range(n_steps):
            step(k + 1, k == n_steps - 1)
        main_s.wait_event(ev_out)                  # the last result is on the host
        e1.record()
        barrier()
    return e0.elapsed_time(e1) / n_steps, h2d, d2h


def sharded_equals_single(ctx, rank, world):
    """Multi-GPU evidence the line can carry: three sampler steps of a small global batch through the
    whole-loop C entry point, each rank its contiguous slice of ONE shared noise tape; the NCCL all-gather of the
    slices must equal, bit for bit, rank 0's own run of the whole batch.  Returns (ok, collective ms of the bench-size
    gather)."""
    import torch.distributed as dist
    lib, L, C = ctx.lib, ctx.L, ctx.C
    dev, size, cx = ctx.dev, ctx.size, ctx.cx
    per, T = 2, 3
    n = per * world
    g = torch.Generator().manual_seed(4242)
    x_T = torch.randn((n, cx, size, size), generator=g)
    tape = torch.randn((T, n, cx, size, size), generator=g)
    gt = torch.rand((n, 3, size, size), generator=g)
    mask = (torch.rand((n, 1, size, size), generator=g) > 0.3).float()
    cond = torch.rand((n, max(ctx.cc, 1), size, size), generator=g)
    from eo_diffusion_b200 import EODiffusion
    diff = EODiffusion(ctx.model, size, cx, timesteps=T, cond_type="sum" if ctx.kind == "sum" else None).to(dev)
    tab = diff._coef_table(dev)

    def run(lo, hi):
        b = hi - lo
        x = x_T[lo:hi].contiguous().to(dev)
        tp = tape[:, lo:hi].contiguous().to(dev)
        rows = diff._timestep_rows(b, dev)
        eps = torch.empty_like(x)
        sum_mode = ctx.kind == "sum"
        g_, m_ = (gt[lo:hi].contiguous().to(dev), mask[lo:hi].contiguous().to(dev)) if sum_mode else (None, None)
        c_ = cond[lo:hi, :ctx.cc].contiguous().to(dev) if ctx.cc else None
        ctx.model(x, rows[0], cond=c_)      # plan for this batch size
        lib.check(L.eo_sample_ddpm(C.c_void_p(ctx.model._handle), lib.ptr(x), lib.ptr(tp), lib.ptr(g_), lib.ptr(m_),
                                   lib.ptr(c_), ctx.cc, None, lib.ptr(rows), lib.ptr(tab), lib.ptr(eps), T, b, cx, size,
                                   size, 1, lib.stream_ptr()), "eo_sample_ddpm")
        return x

    mine = run(rank * per, (rank + 1) * per)
    gathered = torch.empty((n, cx, size, size), device=dev)
    dist.all_gather_into_tensor(gathered, mine)
    ok = torch.ones((1,), device=dev)
    if rank == 0:
        whole = run(0, n)
        ok[0] = 1.0 if torch.equal(whole, gathered) else 0.0
    dist.broadcast(ok, 0)
    # the one collective of the path at bench size: gather the final images of all ranks
    big = torch.empty((world * ctx.B, cx, size, size), device=dev)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dist.all_gather_into_tensor(big, ctx.x)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(ok.item() == 1.0), float(t.item())


def run_ours(args, wl):
    from eo_diffusion_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.check(_lib.lib().eo_device_check(), "eo_device_check")
    mode = args.mode
    peaks = measured_peaks()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.batch or wl["batch"]
    ctx = Ctx(args.workload, wl, B, mode, dev, rank)
    size, kind, steps_per_image = ctx.size, ctx.kind, ctx.steps_per_image

    # ---- device-resident steps ------------------------------------------------------------
    clocks = ClockSampler(local)
    ms_local, launches_step = ctx.time_device_steps(args.steps, args.warmup, barrier, clocks)
    clk = clocks.finish()
    ms_step = max_over_ranks(ms_local)
    value = world * B / (steps_per_image * ms_step * 1e-3)

    # ---- end to end ---------------------------------------------------------------------------
    # the literal metric when it fits in ~1.5 min: one full T = 1000 call of the public sampler; else a shorter instance
    n_e2e = T_DDPM if ms_step * T_DDPM <= 90e3 and not args.short_e2e else max(args.steps, 10)
    if kind == "ddim":
        n_e2e = steps_per_image
    ms_pub, h2d_pub, d2h_pub = e2e_public_sampling(ctx, n_e2e, barrier)
    ms_pub = max_over_ranks(ms_pub)
    e2e_value = world * B / (steps_per_image * ms_pub * 1e-3)
    ms_str, h2d_str, d2h_str = e2e_streamed(ctx, max(2, args.steps), barrier)
    ms_str = max_over_ranks(ms_str)

    # ---- roofline of the dominant kernel family, timed per op with CUDA events --------------
    roofline = None
    flops_img = None
    attention = None
    if rank == 0:
        ops = ctx.op_table()
        breakdown = {}
        for o in ops:
            e = breakdown.setdefault(o["kernel"], dict(ms=0.0, flops=0.0, exec_flops=0.0, bytes=0.0, launches=0))
            e["ms"] += o["ms"]; e["flops"] += o["flops"]; e["exec_flops"] += o["exec_flops"]; e["bytes"] += o["bytes"]
            e["launches"] += 1
        fam = next(k for k in ("k_conv_tc3", "k_conv_simt") if k in breakdown)
        e = breakdown[fam]
        ach = e["flops"] / (e["ms"] * 1e-3) / 1e12
        ach_x = e["exec_flops"] / (e["ms"] * 1e-3) / 1e12
        total_ms = sum(v["ms"] for v in breakdown.values())
        total_fl = sum(v["flops"] for v in breakdown.values())
        flops_img = total_fl / B
        roofline = {"bound": "tensor", "kernel": fam, "achieved": ach, "peak": peaks["tflops"],
                    "unit": "TFLOP/s", "frac": ach / peaks["tflops"], "traffic": ncu_traffic(fam, args.workload, B),
                    "achieved_executed": ach_x, "frac_executed": ach_x / peaks["tflops"],
                    "peak_source": f"{peaks['src']} bf16_tflops_sustained (burst {peaks['burst']})",
                    "flops_per_launch_avg": e["flops"] / e["launches"], "launches_per_step": e["launches"],
                    "share_of_unet_time": e["ms"] / total_ms,
                    "whole_step": {"achieved": total_fl / (ms_step * 1e-3) / 1e12,
                                   "frac": total_fl / (ms_step * 1e-3) / 1e12 / peaks["tflops"]}}
        a = breakdown.get("k_attn_tc6")
        if a:
            attention = {"ms": a["ms"], "share_of_unet_time": a["ms"] / total_ms,
                         "achieved": a["flops"] / (a["ms"] * 1e-3) / 1e12,
                         "achieved_executed": a["exec_flops"] / (a["ms"] * 1e-3) / 1e12, "unit": "TFLOP/s",
                         "bound": "mufu+issue (one ex2 per logit; DESIGN.md section 4)"}
        log("[bench] per-kernel-family breakdown of one UNet forward (CUDA events per op):")
        for kname, v in sorted(breakdown.items(), key=lambda kv: -kv[1]["ms"]):
            extra = ""
            if v["flops"]:
                extra = (f"{v['flops'] / (v['ms'] * 1e-3) / 1e12:8.1f} TFLOP/s algorithmic "
                         f"{v['exec_flops'] / (v['ms'] * 1e-3) / 1e12:8.1f} executed")
            elif v["bytes"]:
                extra = f"{v['bytes'] / (v['ms'] * 1e-3) / 1e9:8.1f} GB/s"
            log(f"    {kname:18s} {v['ms']:9.3f} ms  {100 * v['ms'] / total_ms:5.1f}%  x{v['launches']:3d}  {extra}")
        if args.breakdown:
            with open(args.breakdown, "w") as f:
                json.dump(dict(workload=args.workload, batch=B, size=size, mode=mode, ops=ops), f, indent=1)

    # ---- multi-GPU: bit-equality of the sharded run + the one collective -----------------------
    multi = None
    if world > 1:
        ok, coll_ms = sharded_equals_single(ctx, rank, world)
        multi = {"sharded_equals_single": ok, "collective_ms": coll_ms,
                 "collective": "all_gather_into_tensor of the final images, once per trajectory (outside the timed steps)"}
    ctx.close()

    # ---- the other BASELINE.json configurations, a few device-timed steps each -----------------
    secondary = None
    if not args.no_secondary and args.workload == "c3" and not args.batch:
        secondary = {}
        names = ["c1", "c2", "c4", "c5"] if world == 1 else ["c5"]
        for name in names:
            w2 = WORKLOADS[name]
            try:
                c2 = Ctx(name, w2, w2["batch"], mode, dev, rank)
                k2 = 20 if name in ("c1", "c2") else 3
                ms2, l2 = c2.time_device_steps(k2, 3, barrier)
                ms2 = max_over_ranks(ms2)
                fl2 = c2.alg_flops_per_image()
                spi = c2.steps_per_image
                secondary[name] = {"workload": w2["desc"], "value": world * w2["batch"] / (spi * ms2 * 1e-3), "unit": UNIT,
                                   "steps_per_image": spi, "ms_per_step": ms2, "steps": k2, "batch_per_gpu": w2["batch"],
                                   "frac": fl2 * w2["batch"] / (ms2 * 1e-3) / 1e12 / peaks["tflops"],
                                   "gpu_launches_per_step": l2}
                if name == "c1":
                    # the reference's only published speed for this path (tqdm lines of the demo
                    # notebook, EO_Diffusion.ipynb:249-279 -- RTX 4000, fp32 eager, incl. its PNG writes): other
                    # hardware, so a note next to the line, never `vs_baseline`
                    secondary[name]["published"] = {"it_per_s": 26.3, "source": "EO_Diffusion.ipynb:249-279 (RTX 4000, torch 1.13 "
                                                    "fp32 eager, batch 1, 64x64, incl. save_image calls)",
                                                    "ours_it_per_s": 1e3 / ms2, "ratio": (1e3 / ms2) / 26.3}
                c2.close()
            except Exception as ex:      # a secondary line must not cost the headline
                secondary[name] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
                gc.collect()
                torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and kind == "sum":
        sample_b, sample_steps = 1, 10        # ~1 s per step on 16 host threads: 10-12 s of CPU work
        sec, ckind = cpu_reference_steps(size, sample_b, sample_steps, 1)
        cpu = {"value": sample_b / (T_DDPM * sec), "unit": UNIT, "cores": torch.get_num_threads(),
               "kind": ckind, "ms_per_step": sec * 1e3,
               "sample": f"{sample_b} image of {size}x{size}, {sample_steps} sampler steps after 1 warm-up ("
                         + ("the reference's own EODiffusion.sampling + UNetModel from oracle/_ref"
                            if ckind == "reference" else "oracle port of the reference CPU path") + ", fp32)"}

    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()

    if rank == 0:
        print(json.dumps({
            "metric": METRIC if kind != "ddim" else f"ddim_S{steps_per_image}_images_per_sec", "value": value,
            "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": mode, "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}, "
                                   + ("DDIM" if kind == "ddim" else f"DDPM T={T_DDPM}, "
                                      + ("cond 'sum' (3+3->3 ch), " if kind == "sum" else "cond concatenated, ")
                                      + "clipped posterior")
                                   + ", UNet base 128 mult [1,2,3,4] attn [4,8] 2 res blocks 8 heads",
                       "batch_per_gpu": B, "image": size, "parallelism": f"batch-sharded x{world}",
                       "l2": "activations per step (GBs) exceed the 126 MB L2; no flush needed"
                             if B * size * size * 128 * 2 > 4 * 126e6 else
                             "working set fits L2: numbers are L2-warm"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_pub, "d2h_bytes_per_step": d2h_pub,
                    "ms_per_step": ms_pub, "steps": n_e2e,
                    "call": ("EODiffusion.sampling(n, device, cond=<pinned host tensor>).cpu(), one call of "
                             f"{n_e2e} steps" + (" = the full T = 1000 trajectory" if n_e2e == T_DDPM else "")
                             + "; x_T drawn on the host like the reference; copies once per call, divided over its steps")
                            if kind != "ddim" else "DDIMSampler.sample(S, ...)[0].cpu()"},
            "e2e_streamed": {"value": world * B / (steps_per_image * ms_str * 1e-3), "unit": UNIT,
                             "h2d_bytes_per_step": h2d_str, "d2h_bytes_per_step": d2h_str, "ms_per_step": ms_str},
            "gpu_launches": launches_step * args.steps,
            "clocks": clk,
            "roofline": roofline,
            "attention": attention,
            "unet_algorithmic_gflop_per_image_step": flops_img / 1e9 if flops_img else None,
            "multi_gpu": multi,
            "secondary": secondary,
            "cpu_baseline": cpu,
        }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--breakdown", default="", help="write the per-op timing JSON here")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the c1/c2/c4/c5 block of the default line")
    ap.add_argument("--short-e2e", action="store_true", help="time a --steps-long sampling() call instead of the full T = 1000 one")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                    # timing rule: at least 3 warm-up steps
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
