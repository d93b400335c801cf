"""Round-2 parity evidence on the GPU (through the drop-in API -> C ABI -> CUDA): the geometries bench.py
measures, the full T = 1000 trajectory with the real UNet, wide attention heads, classifier-free guidance,
EODiffusion.forward, checkpoint files, the hoisted timestep tables, and the advisor's lifetime cases.

Fixtures come from the live reference (oracle/make_golden_r2.py); tolerances are BASELINE.json's
(eps relative L2 <= 1e-4 fp32 / <= 1e-2 bf16) and the trajectory bounds stated per test.  Needs a B200."""
import copy
import ctypes as C
import gc
import json

import pytest
import torch

from conftest import build_unet, golden, golden_cfg, rel_l2, replay, tt, weight_checksum
from eo_diffusion_b200 import DDIMSampler, EODiffusion, _lib
from eo_diffusion_b200.checkpoint import load_checkpoint, make_checkpoint
from oracle import oracle as O

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 1e-2}
TOL_BF16_NARROW = 1.25e-2          # 64-wide synthetic nets, see tests/test_gpu_unet.py
_models = {}


def model_for(g, dev, mode):
    cfg = golden_cfg(g)
    key = (json.dumps(cfg, sort_keys=True), mode)
    if key not in _models:
        m = build_unet(cfg, int(g["init_seed"]), int(g["dezero_seed"]))
        assert weight_checksum(m.state_dict()) == json.loads(str(g["wsum"]))["sha256"]
        _models[key] = m.to(dev).set_compute_mode(mode)
    return _models[key], cfg


def seeded_inputs(g, cfg):
    """x (and cond) exactly as oracle/make_golden_r2.py::eps_case drew them."""
    gen = torch.Generator().manual_seed(int(g["x_seed"]))
    B, cc, size = len(g["t"]), int(g["cond_ch"]), cfg["image_size"]
    x = torch.randn((B, cfg["in_channels"] - cc, size, size), generator=gen)
    cond = torch.rand((B, cc, size, size), generator=gen) if cc else None
    return x, cond


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["base128_eps_b2", "base256_eps_b2", "base64_ms_concat_eps", "heads1_eps"])
def test_eps_at_benchmarked_geometries(cuda_dev, name, mode):
    """BASELINE configs c2 (128 x 128), c3 (256 x 256), c5 (13 + 15 -> 13 channels at base width) and the
    reference scripts' own num_heads = 1 (head dimensions 128 / 256), against the live reference's eps."""
    g = golden(name)
    m, cfg = model_for(g, cuda_dev, mode)
    x, cond = seeded_inputs(g, cfg)
    eps = m(x.to(cuda_dev), tt(g["t"]).to(cuda_dev), cond=None if cond is None else cond.to(cuda_dev))
    assert eps.shape == g["eps"].shape and eps.dtype == torch.float32
    err = rel_l2(eps, tt(g["eps"]))
    print(f"[parity] {mode} {name}: eps rel L2 {err:.3e}")
    tol = TOL[mode] if (mode == "fp32" or cfg["model_channels"] >= 128) else TOL_BF16_NARROW
    assert err <= tol, f"{name} {mode}: rel L2 {err:.3e} > {tol}"
    # the second call is captured into a CUDA graph, the third replays it: same bits
    xd, td = x.to(cuda_dev), tt(g["t"]).to(cuda_dev)
    cd = None if cond is None else cond.to(cuda_dev)
    assert torch.equal(m(xd, td, cond=cd), eps) and torch.equal(m(xd, td, cond=cd), eps)
    _models.pop((json.dumps(cfg, sort_keys=True), mode), None)      # the 256 x 256 plan holds ~1 GB; free it
    gc.collect()


@pytest.mark.parametrize("mode,mae_tol", [("fp32", 1e-3), ("bf16", 2e-2)])
def test_c1_full_T1000_trajectory(cuda_dev, tmp_path, mode, mae_tol):
    """BASELINE config c1 exactly: EODiffusion.sampling, T = 1000, 64 x 64, batch 1, 'sum' conditioning, clipped,
    with the real UNet, replaying the noise tape of the live reference's run (reference == oracle bit-exact over the
    1000 steps).  Stated trajectory tolerance: mean-abs error of x_0 <= 1e-3 in fp32 mode,
    <= 2e-2 in bf16 mode; the timestep sequence is exact and the first mixed state bit-identical."""
    g = golden("base64_ddpm_sum_T1000")
    m, cfg = model_for(g, cuda_dev, mode)
    T, n, size = int(g["T"]), int(g["n"]), cfg["image_size"]
    d = EODiffusion(m, size, 3, timesteps=T, cond_type="sum").to(cuda_dev)
    x_T, tape = O.noise_tape((n, 3, size, size), T, seed=int(g["tape_seed"]))
    rec = {}
    keep = (0, 100, 500, 900, 999)
    count = [0]
    seq = []

    def hook(mod, args, kw, out):
        k = count[0]
        count[0] += 1
        seq.append(int(args[1][0]))
        if k in keep:
            rec[k] = (args[0].detach().clone(), out.detach().clone())

    h = m.register_forward_hook(hook, with_kwargs=True)
    try:
        with replay([x_T], tape, tmp_cwd=str(tmp_path)):
            out = d.sampling(n, clipped_reverse_diffusion=True, device=cuda_dev, cond=tt(g["cond"]))
    finally:
        h.remove()
    assert seq == list(g["t_seq"])
    assert torch.equal(rec[0][0].cpu(), tt(g["xt_step0"]))
    assert rel_l2(rec[0][1], tt(g["eps_step0"])) <= TOL[mode]
    drift = {k: float((rec[k][0].cpu() - tt(g[f"xt_step{k}"])).abs().mean()) for k in keep}
    mae = float((out.cpu() - tt(g["x0"])).abs().mean())
    print(f"[parity] {mode} c1 T=1000 trajectory: x_0 mean-abs {mae:.3e}; x_t mean-abs drift at loop iteration "
          + ", ".join(f"{k}: {v:.2e}" for k, v in drift.items()))
    assert bool(torch.isfinite(out).all())
    assert mae <= mae_tol, mae


def test_batch64_256_sample_equals_itself_alone(cuda_dev):
    """Config c3's geometry at full occupancy: 256 x 256, batch 64 (tile counts that fill the 148 SMs several times,
    odd tail waves, the persistent schedules of every launch).  Samples are independent, the kernels deterministic:
    sample 17 of the 64 is bit-identical to the same sample run alone, and two runs of the batch agree bit for bit."""
    g = golden("base256_eps_b2")
    m, cfg = model_for(g, cuda_dev, "bf16")
    gen = torch.Generator().manual_seed(64)
    x = torch.randn((64, 3, 256, 256), generator=gen).to(cuda_dev)
    t = torch.randint(0, 1000, (64,), generator=gen).to(cuda_dev)
    full = m(x, t)
    again = m(x, t)
    alone = m(x[17:18].contiguous(), t[17:18].contiguous())
    assert bool(torch.isfinite(full).all())
    assert torch.equal(full, again)
    assert torch.equal(full[17:18], alone)
    # and sample 0 / 1 of the fixture still match the reference inside the big plan
    xg, _ = seeded_inputs(g, cfg)
    eps = m(xg.to(cuda_dev), tt(g["t"]).to(cuda_dev))
    assert rel_l2(eps, tt(g["eps"])) <= TOL["bf16"]
    _models.clear()
    gc.collect()
    torch.cuda.empty_cache()


@pytest.mark.parametrize("B,T,heads,ch", [(1, 64, 1, 512), (2, 256, 1, 128), (1, 1024, 1, 1024), (2, 100, 2, 192),
                                          (1, 64, 2, 72), (1, 200, 3, 20)])
def test_attention_wide_heads(cuda_dev, B, T, heads, ch):
    """Head dimensions the tensor-core kernel does not take (> 64, or not a multiple of 8): k_attention_wide,
    against QKVAttentionLegacy in fp32 torch on the same bf16 operands."""
    from test_gpu_tc import _attn_ref
    gen = torch.Generator().manual_seed(T + heads + ch)
    qkv = torch.randn((B, T, heads * 3 * ch), generator=gen).to(cuda_dev).to(torch.bfloat16)
    out = torch.empty((B, T, heads * ch), dtype=torch.bfloat16, device=cuda_dev)
    _lib.check(_lib.lib().eo_test_attention_tc(_lib.ptr(qkv), _lib.ptr(out), B, T, heads, ch, _lib.stream_ptr()),
               "eo_test_attention_tc")
    torch.cuda.synchronize()
    assert rel_l2(out.float(), _attn_ref(qkv, heads, ch)) <= 4e-3


def test_cfg_ddim_trajectory_fp32(cuda_dev):
    """Classifier-free guidance through DDIMSampler.sample (ddim.py:176-181): doubled batch, cat(uncond, cond),
    e_u + s (e_c - e_u); against the live reference's run (tiny concat-conditioned UNet, eta 0.5, scale 3)."""
    g = golden("tiny_cfg_ddim_S4_T8")
    m, cfg = model_for(g, cuda_dev, "fp32")
    d = EODiffusion(m, 16, 3, timesteps=8).to(cuda_dev)
    smp = DDIMSampler(d)
    n, S = int(g["n"]), 4
    x_T, tape = O.noise_tape((n, 3, 16, 16), S, seed=int(g["tape_seed"]))
    cond = tt(g["cond"]).to(cuda_dev)
    with replay(tape, [None] * S):
        out, inter = smp.sample(S, n, (3, 16, 16), conditioning=cond, eta=float(g["eta"]), x_T=x_T.to(cuda_dev),
                                verbose=False, log_every_t=1, unconditional_guidance_scale=float(g["scale"]),
                                unconditional_conditioning=torch.zeros_like(cond))
    assert rel_l2(out, tt(g["x0"])) <= 5e-3
    assert rel_l2(inter["pred_x0"][-1], tt(g["pred_x0_last"])) <= 5e-3


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_cfg_ddim_against_oracle_64wide(cuda_dev, mode):
    """The same branch in both arithmetic modes on a 64-wide concat-conditioned net (the narrowest the tensor-core
    mode takes), against the oracle evaluated here."""
    cfg = dict(image_size=32, in_channels=5, model_channels=64, out_channels=3, num_res_blocks=1,
               attention_resolutions=[2], channel_mult=[1, 2], num_heads=2)
    m = build_unet(cfg, 81, 82)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    T, S, n = 20, 5, 2
    x_T, tape = O.noise_tape((n, 3, 32, 32), S, seed=83)
    cond = torch.rand((n, 2, 32, 32), generator=torch.Generator().manual_seed(84))
    want, _ = O.ddim_sample(sd, O.full_cfg(**cfg), O.cosine_schedule(T), S, x_T, tape, eta=0.3, cond=cond,
                            unconditional_guidance_scale=2.5, unconditional_conditioning=torch.zeros_like(cond))
    m = m.to(cuda_dev).set_compute_mode(mode)
    d = EODiffusion(m, 32, 3, timesteps=T).to(cuda_dev)
    smp = DDIMSampler(d)
    with replay(tape, [None] * S):
        out, _ = smp.sample(S, n, (3, 32, 32), conditioning=cond.to(cuda_dev), eta=0.3, x_T=x_T.to(cuda_dev),
                            verbose=False, unconditional_guidance_scale=2.5,
                            unconditional_conditioning=torch.zeros_like(cond).to(cuda_dev))
    err = rel_l2(out, want)
    print(f"[parity] {mode} CFG DDIM S=5: x_0 rel L2 {err:.3e}")
    assert err <= (5e-3 if mode == "fp32" else 8e-2)      # un-clipped DDIM amplifies eps error; relative, not absolute


def test_eodiffusion_forward_training_call(cuda_dev):
    """EODiffusion.forward (model.py:38-44): random t (replayed), q-sample, UNet -- against the live reference."""
    g = golden("tiny_forward_train")
    m, cfg = model_for(g, cuda_dev, "fp32")
    d = EODiffusion(m, 16, 3, timesteps=1000).to(cuda_dev)
    tdraw = tt(g["t"])
    o_randint = torch.randint
    torch.randint = lambda *a, **k: tdraw
    try:
        eps = d(tt(g["x"]).to(cuda_dev), tt(g["noise"]).to(cuda_dev))
    finally:
        torch.randint = o_randint
    assert rel_l2(eps, tt(g["eps"])) <= TOL["fp32"]


def test_checkpoint_file_to_forward(cuda_dev, tmp_path):
    """SURVEY.md 8f row 1 on the GPU: a checkpoint written the way reference train.py:137-138 writes it ->
    load_checkpoint(which="model_ema") into a freshly initialised drop-in -> forward == the golden eps."""
    g = golden("small_eps")
    cfg = golden_cfg(g)
    good = build_unet(cfg, int(g["init_seed"]), int(g["dezero_seed"]))
    src = EODiffusion(good, cfg["image_size"], 3, timesteps=1000)
    path = tmp_path / "clouds.pt"
    torch.save(make_checkpoint(src), path)
    fresh = build_unet(cfg, 5, 6).to(cuda_dev).set_compute_mode("fp32")
    dst = EODiffusion(fresh, cfg["image_size"], 3, timesteps=1000).to(cuda_dev)
    x, t = tt(g["x"]).to(cuda_dev), tt(g["t"]).to(cuda_dev)
    before = fresh(x, t)                                    # plans the engine with the wrong weights first
    res = load_checkpoint(dst, str(path), which="model_ema")
    assert not res.missing_keys and not res.unexpected_keys
    after = fresh(x, t)
    assert rel_l2(before, tt(g["eps"])) > 1e-2
    assert rel_l2(after, tt(g["eps"])) <= TOL["fp32"]
    fresh.set_compute_mode("bf16")
    assert rel_l2(fresh(x, t), tt(g["eps"])) <= TOL_BF16_NARROW


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_time_tables_are_bit_identical(cuda_dev, mode):
    """eo_unet_build_time_tables: the hoisted embedding rows equal the per-step path bit for bit;
    a timestep outside the table is loud (NaN), and closing the block restores the general path."""
    g = golden("small_eps")
    m, cfg = model_for(g, cuda_dev, mode)
    gen = torch.Generator().manual_seed(9)
    x = torch.randn((3, 3, 32, 32), generator=gen).to(cuda_dev)
    t = torch.tensor([999, 0, 437]).to(cuda_dev)
    plain = m(x, t)
    n_plain = m.launches_per_forward()
    with m.time_tables(1000):
        hoisted = m(x, t)
        hoisted2 = m(x, t)                 # the graph-replayed call
        assert m.launches_per_forward() == n_plain - 3
        outside = m(x, torch.tensor([999, 1000, 5]).to(cuda_dev))
    back = m(x, t)
    assert torch.equal(plain, hoisted) and torch.equal(plain, hoisted2) and torch.equal(plain, back)
    assert bool(torch.isnan(outside[1]).all()) and bool(torch.isfinite(outside[0]).all())
    assert m.launches_per_forward() == n_plain


def test_deepcopy_gets_its_own_engine(cuda_dev):
    """copy.deepcopy after a forward (torch.optim.swa_utils.AveragedModel, the reference's EMA wrapper, does this):
    the copy must not share the C handle -- each model keeps answering with ITS weights, and the copy survives the
    original's destruction."""
    g = golden("tiny_eps")
    cfg = golden_cfg(g)
    a = build_unet(cfg, int(g["init_seed"]), int(g["dezero_seed"])).to(cuda_dev).set_compute_mode("fp32")
    x, t = tt(g["x"]).to(cuda_dev), tt(g["t"]).to(cuda_dev)
    ya = a(x, t)
    b = copy.deepcopy(a)
    assert b._handle is None
    with torch.no_grad():
        for p in b.parameters():
            p.mul_(1.01)
    yb = b(x, t)
    assert b._handle is not None and b._handle != a._handle
    assert torch.equal(a(x, t), ya)                       # a was not re-finalized with b's weights
    assert rel_l2(ya, tt(g["eps"])) <= TOL["fp32"] and rel_l2(yb, ya) > 1e-4
    del a
    gc.collect()
    assert torch.equal(b(x, t), yb)


def test_concat_stem_after_dtype_round_trip(cuda_dev):
    """The (x, cond) split of the stem weight is packed at the first concat-conditioned forward, from the engine's
    private copy: parameters that are not fp32 (staged through a temporary) must still give the right stem."""
    g = golden("tiny_concat_eps")
    cfg = golden_cfg(g)
    m = build_unet(cfg, int(g["init_seed"]), int(g["dezero_seed"])).double().to(cuda_dev).set_compute_mode("fp32")
    x, t, cond = tt(g["x"]).to(cuda_dev), tt(g["t"]).to(cuda_dev), tt(g["cond"]).to(cuda_dev)
    junk = [torch.full((1 << 20,), float("nan"), device=cuda_dev) for _ in range(8)]     # recycle freed staging memory
    del junk
    eps = m(x, t, cond=cond)
    assert rel_l2(eps, tt(g["eps"])) <= TOL["fp32"]


def test_whole_loop_entry_points_validate_geometry(cuda_dev):
    g = golden("tiny_eps")
    m, cfg = model_for(g, cuda_dev, "fp32")
    x, t = tt(g["x"]).to(cuda_dev), tt(g["t"]).to(cuda_dev)
    m(x, t)
    d = EODiffusion(m, 16, 3, timesteps=4).to(cuda_dev)
    rows, tab = d._timestep_rows(2, cuda_dev), d._coef_table(cuda_dev)
    buf = torch.zeros((4, 2, 3, 32, 32), device=cuda_dev)
    rc = _lib.lib().eo_sample_ddpm(C.c_void_p(m._handle), _lib.ptr(buf[0]), _lib.ptr(buf), None, None, None, 0, None,
                                   _lib.ptr(rows), _lib.ptr(tab), _lib.ptr(buf[1]), 4, 2, 3, 32, 32, 1, _lib.stream_ptr())
    assert rc < 0 and "finalized geometry" in _lib.last_error()
