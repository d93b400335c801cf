"""Round-2 golden vectors from the LIVE reference (/root/reference): the geometries bench.py measures and the
branches the first set left out.  TEST INFRASTRUCTURE ONLY; run in the build container:

    python oracle/make_golden_r2.py [substring filter]

Same harness as oracle/make_golden.py (replayed noise tape, PNG side effect patched out, weights re-created from
(cfg, init seed, de-zero seed)).  Cases:

  base128_eps_b2        BASELINE architecture at 128 x 128 (config c2's geometry), batch 2, t = (999, 3)
  base256_eps_b2        ... at 256 x 256 (config c3's geometry), batch 2, t = (500, 0)
  base64_ms_concat_eps  ... with 13 + 15 -> 13 channels (config c5's stem / head), 64 x 64, batch 1
  heads1_eps            num_heads = 1 as in the reference's own scripts (inference.py:59, train.py:50):
                        head dimensions 128 and 256 (> 64)
  base64_ddpm_sum_T1000 config c1 exactly: EODiffusion.sampling over the full T = 1000 schedule with the real UNet,
                        'sum' conditioning, clipped (~60 TFLOP on the CPU, ~4 min; the oracle restatement is run over the
                        same 1000 steps and must agree)
  tiny_cfg_ddim_S4_T8   DDIMSampler.sample with classifier-free guidance (ddim.py:176-181) on a concat-conditioned
                        tiny UNet, eta 0.5, scale 3
  tiny_forward_train    EODiffusion.forward (model.py:38-44): random t, q-sample, UNet

Inputs that are pure functions of a seed (x, noise tapes) are stored as the seed; eps / x0 as fp32 arrays.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (imports the live reference and the oracle)

O = G.O
FILTER = sys.argv[1] if len(sys.argv) > 1 else ""


def want(name):
    return not FILTER or FILTER in name


def save(name, pinned, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    np.savez_compressed(os.path.join(G.GOLD, name + ".npz"), **out)
    path = os.path.join(G.GOLD, "MANIFEST.json")
    with open(path) as f:
        man = json.load(f)
    man["cases"][name] = {"oracle_vs_reference": pinned, "arrays": {k: list(np.shape(v)) for k, v in out.items()}}
    with open(path, "w") as f:
        json.dump(man, f, indent=1, sort_keys=True)
    print(f"[golden-r2] {name}: {pinned}", flush=True)


def seeded_x(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def eps_case(name, cfg, B, ts, x_seed, cond_ch=0, check_oracle=True):
    if not want(name):
        return
    t0 = time.time()
    m = G.build_ref_unet(cfg)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    size = cfg["image_size"]
    g = torch.Generator().manual_seed(x_seed)
    x = torch.randn((B, cfg["in_channels"] - cond_ch, size, size), generator=g)
    cond = torch.rand((B, cond_ch, size, size), generator=g) if cond_ch else None
    t = torch.tensor(ts, dtype=torch.long)
    with torch.no_grad():
        ref = m(x, t, cond=cond)
    pinned = "reference only"
    if check_oracle:
        got = O.unet_forward(sd, O.full_cfg(**cfg), x, t, cond=cond)
        pinned = "bit-exact" if G.eq(ref, got) else f"rel_l2={O.rel_l2(got, ref):.3e}"
        assert O.rel_l2(got, ref) < 1e-6, pinned
    save(name, pinned, x_seed=np.int64(x_seed), cond_ch=np.int64(cond_ch), t=t, eps=ref, cfg=json.dumps(cfg),
         init_seed=np.int64(G.INIT_SEED), dezero_seed=np.int64(G.DEZERO_SEED),
         wsum=json.dumps(G.weight_checksum(sd)))
    print(f"   ({time.time() - t0:.0f} s)", flush=True)


HEADS1 = dict(image_size=32, in_channels=3, model_channels=64, out_channels=3, num_res_blocks=1,
              attention_resolutions=[2], channel_mult=[1, 2, 4], num_heads=1)


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    eps_case("heads1_eps", HEADS1, 2, [999, 7], x_seed=301)
    eps_case("base64_ms_concat_eps", dict(G.BASE, in_channels=28, out_channels=13), 1, [400], x_seed=302, cond_ch=15)
    eps_case("base128_eps_b2", dict(G.BASE, image_size=128), 2, [999, 3], x_seed=303)
    eps_case("base256_eps_b2", dict(G.BASE, image_size=256), 2, [500, 0], x_seed=304)

    # ---- classifier-free guidance through the reference's DDIM sampler ------------------------------
    if want("tiny_cfg_ddim_S4_T8"):
        cfg = dict(G.TINY, in_channels=5)
        m = G.build_ref_unet(cfg)
        sd = {k: v.detach() for k, v in m.state_dict().items()}
        T, S, n, eta, scale = 8, 4, 2, 0.5, 3.0
        d = G.RefEODiffusion(m, 16, 3, timesteps=T).eval()
        smp = G.RefDDIMSampler(d)
        x_T, tape = O.noise_tape((n, 3, 16, 16), S, seed=51)
        cond = torch.rand((n, 2, 16, 16), generator=torch.Generator().manual_seed(52))
        uncond = torch.zeros_like(cond)
        with G.replay(tape, [None] * S):
            ref, inter = smp.sample(S, n, (3, 16, 16), conditioning=cond, eta=eta, x_T=x_T, verbose=False, log_every_t=1,
                                    unconditional_guidance_scale=scale, unconditional_conditioning=uncond)
        got, ginter = O.ddim_sample(sd, O.full_cfg(**cfg), O.cosine_schedule(T), S, x_T, tape, eta=eta, cond=cond,
                                    log_every_t=1, unconditional_guidance_scale=scale, unconditional_conditioning=uncond)
        assert O.rel_l2(got, ref) < 1e-5
        # the guidance must matter, or the fixture pins nothing
        plain, _ = O.ddim_sample(sd, O.full_cfg(**cfg), O.cosine_schedule(T), S, x_T, tape, eta=eta, cond=cond)
        assert O.rel_l2(plain, ref) > 1e-2
        save("tiny_cfg_ddim_S4_T8", "bit-exact" if G.eq(ref, got) else "rel_l2<1e-5", x0=ref, cond=cond,
             tape_seed=np.int64(51), n=np.int64(n), scale=np.float64(scale), eta=np.float64(eta), cfg=json.dumps(cfg),
             init_seed=np.int64(G.INIT_SEED), dezero_seed=np.int64(G.DEZERO_SEED),
             wsum=json.dumps(G.weight_checksum(sd)), pred_x0_last=inter["pred_x0"][-1])

    # ---- EODiffusion.forward (training-side call) --------------------------------------------------
    if want("tiny_forward_train"):
        m = G.build_ref_unet(G.TINY)
        sd = {k: v.detach() for k, v in m.state_dict().items()}
        d = G.RefEODiffusion(m, 16, 3, timesteps=1000).eval()
        g = torch.Generator().manual_seed(61)
        x = torch.rand((3, 3, 16, 16), generator=g)
        noise = torch.randn((3, 3, 16, 16), generator=g)
        tdraw = torch.tensor([17, 999, 0])
        o_randint = torch.randint
        torch.randint = lambda *a, **k: tdraw
        try:
            with torch.no_grad():
                ref = d(x, noise)
        finally:
            torch.randint = o_randint
        s = O.cosine_schedule(1000)
        got = O.unet_forward(sd, O.full_cfg(**G.TINY), O.forward_diffusion(s, x, tdraw, noise), tdraw)
        assert O.rel_l2(got, ref) < 1e-6
        save("tiny_forward_train", "bit-exact" if G.eq(ref, got) else "rel_l2<1e-6", x=x, noise=noise, t=tdraw, eps=ref,
             cfg=json.dumps(G.TINY), init_seed=np.int64(G.INIT_SEED), dezero_seed=np.int64(G.DEZERO_SEED),
             wsum=json.dumps(G.weight_checksum(sd)))

    # ---- config c1: the full T = 1000 trajectory with the real UNet (reference only) ------------------
    if want("base64_ddpm_sum_T1000"):
        t0 = time.time()
        m = G.build_ref_unet(G.BASE)
        sd = {k: v.detach() for k, v in m.state_dict().items()}
        T, n, size, seed = 1000, 1, 64, 71
        d = G.RefEODiffusion(m, size, 3, timesteps=T, cond_type="sum").eval()
        x_T, tape = O.noise_tape((n, 3, size, size), T, seed=seed)
        cond = O.synth_cond_sum(n, size, seed=seed + 1)
        rec = []
        h = G.hook_eps(d, rec)
        with G.replay([x_T], tape):
            ref = d.sampling(n, clipped_reverse_diffusion=True, cond=cond)
        h.remove()
        keep = (0, 100, 500, 900, 999)
        arrs = dict(x0=ref, cond=cond, tape_seed=np.int64(seed), n=np.int64(n), T=np.int64(T), cfg=json.dumps(G.BASE),
                    init_seed=np.int64(G.INIT_SEED), dezero_seed=np.int64(G.DEZERO_SEED),
                    wsum=json.dumps(G.weight_checksum(sd)), t_seq=np.asarray([r[0] for r in rec], dtype=np.int64))
        for k in keep:
            arrs[f"xt_step{k}"] = rec[k][1]
            arrs[f"eps_step{k}"] = rec[k][2]
        # the oracle restatement over the same 1000 steps (another ~4 min of CPU)
        got = O.ddpm_sample(sd, O.full_cfg(**G.BASE), O.cosine_schedule(T), x_T, tape, cond=cond, cond_type="sum", clipped=True)
        pinned = "bit-exact" if G.eq(ref, got) else f"rel_l2={O.rel_l2(got, ref):.3e}"
        assert O.rel_l2(got, ref) < 1e-5, pinned
        save("base64_ddpm_sum_T1000", pinned, **arrs)
        print(f"   ({time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    main()
