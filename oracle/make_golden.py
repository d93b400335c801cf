"""Generate tests/golden/*.npz from the LIVE reference (/root/reference), and pin the
oracle restatement (oracle/oracle.py) against it.  TEST INFRASTRUCTURE ONLY.

Run in the build container (the reference is not present on the GPU box):

    python oracle/make_golden.py

For every case the reference's own classes are driven through their public API
(`UNetModel.forward`, `EODiffusion.sampling`, `DDIMSampler.sample`) with
  * `torch.randn` / `torch.randn_like` patched to replay a pre-drawn noise tape in the
    reference's draw order,
  * `diffusion.model.save_image` patched out (F3: `sampling()` writes PNGs regardless of
    `save`),
  * `DDIMSampler.register_buffer` patched to a plain setattr (F7: it hard-requires CUDA),
and the oracle restatement is run on the same inputs.  The script asserts the two agree
BIT-FOR-BIT on CPU and records that in tests/golden/MANIFEST.json.  Weights are not
stored: they are re-created from (cfg, init seed, de-zero seed); each file carries a
checksum of the weights so a consumer can tell whether it reproduced them.
"""
from __future__ import annotations

import contextlib
import json
import os
import sys
import hashlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("EO_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import oracle as O  # noqa: E402

import diffusion.model as ref_model_mod  # noqa: E402
import diffusion.ddim as ref_ddim_mod  # noqa: E402
import diffusion.util as ref_util_mod  # noqa: E402
from backbones.unet_openai import UNetModel as RefUNet  # noqa: E402
from diffusion.model import EODiffusion as RefEODiffusion  # noqa: E402
from diffusion.ddim import DDIMSampler as RefDDIMSampler  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

TINY = dict(image_size=16, in_channels=3, model_channels=32, out_channels=3, num_res_blocks=1,
            attention_resolutions=[2], channel_mult=[1, 2], num_heads=2)
# mid-size: exercises every distinct layer kind of the BASELINE arch (non-pow2 channel
# counts, groups straddling the concat boundary, head dim 48, 3 levels) at small cost
SMALL = dict(image_size=32, in_channels=3, model_channels=64, out_channels=3, num_res_blocks=1,
             attention_resolutions=[2, 4], channel_mult=[1, 2, 3], num_heads=4)
BASE = dict(image_size=64, in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2,
            attention_resolutions=[4, 8], channel_mult=[1, 2, 3, 4], num_heads=8)
INIT_SEED = 1234
DEZERO_SEED = 4321


def weight_checksum(sd) -> dict:
    h = hashlib.sha256()
    tot = 0.0
    n = 0
    for k in sorted(sd.keys()):
        v = sd[k].detach().cpu().contiguous()
        h.update(k.encode())
        h.update(v.numpy().tobytes())
        tot += float(v.double().abs().sum())
        n += v.numel()
    return {"sha256": h.hexdigest(), "abs_sum": tot, "numel": n}


def build_ref_unet(cfg, init_seed=INIT_SEED, dezero_seed=DEZERO_SEED):
    torch.manual_seed(init_seed)
    m = RefUNet(**cfg)
    O.dezero_(m, dezero_seed)
    return m.eval()


@contextlib.contextmanager
def replay(randn_queue, randn_like_queue):
    """Patch torch.randn / torch.randn_like to pop pre-drawn tensors (None in the
    randn_like queue => return zeros: a draw the reference discards)."""
    o_randn, o_like = torch.randn, torch.randn_like
    rq, lq = list(randn_queue), list(randn_like_queue)

    def f_randn(*a, **k):
        t = rq.pop(0)
        dev = k.get("device", None)
        return t.to(dev) if dev is not None else t

    def f_like(x, **k):
        t = lq.pop(0)
        return torch.zeros_like(x) if t is None else t.to(x.device)

    torch.randn, torch.randn_like = f_randn, f_like
    o_save = ref_model_mod.save_image
    ref_model_mod.save_image = lambda *a, **k: None
    o_reg = RefDDIMSampler.register_buffer
    RefDDIMSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)
    try:
        yield
    finally:
        torch.randn, torch.randn_like = o_randn, o_like
        ref_model_mod.save_image = o_save
        RefDDIMSampler.register_buffer = o_reg


class Stub(torch.nn.Module):
    """Cheap stand-in for the UNet so that the sampler arithmetic can be pinned over the
    full T=1000 schedule: eps = 0.3 * x_t - 0.1 + 1e-3 * t."""

    def forward(self, x, t, cond=None, y=None):
        return 0.3 * x - 0.1 + 1e-3 * t.float().reshape(-1, 1, 1, 1)


def stub_eps(x, t, cond, y):
    return 0.3 * x - 0.1 + 1e-3 * t.float().reshape(-1, 1, 1, 1)


def hook_eps(diff, rec):
    def hook(mod, args, kwargs, out):
        rec.append((int(args[1][0]), args[0].detach().clone(), out.detach().clone()))
    return diff.model.register_forward_hook(hook, with_kwargs=True)


manifest = {"reference": "furio1999/EO_Diffusion", "torch": torch.__version__, "cases": {}}


ONLY = os.environ.get("EO_GOLDEN_ONLY", "")   # substring filter: regenerate matching cases only (manifest merged)


def save(name, pinned, **arrs):
    if ONLY and ONLY not in name:
        return
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    manifest["cases"][name] = {"oracle_vs_reference": pinned,
                               "arrays": {k: list(np.shape(v)) for k, v in out.items()}}
    print(f"[golden] {name}: {pinned}")


def eq(a, b):
    return bool(torch.equal(a, b))


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---- schedule tables -----------------------------------------------------------
    for T in (8, 20, 1000):
        d = RefEODiffusion(Stub(), 8, 3, timesteps=T)
        s = O.cosine_schedule(T)
        ok = all(eq(getattr(d, k), s[k]) for k in s)
        assert ok, f"schedule T={T} differs"
        save(f"schedule_T{T}", "bit-exact", **{k: getattr(d, k) for k in s})

    # ---- DDIM tables ---------------------------------------------------------------
    for (S, T, eta) in ((50, 1000, 0.0), (50, 1000, 0.5), (4, 8, 0.0), (4, 8, 0.5), (8, 8, 0.0)):
        d = RefEODiffusion(Stub(), 8, 3, timesteps=T)
        smp = RefDDIMSampler(d)
        with replay([], []):
            smp.make_schedule(S, ddim_eta=eta, verbose=False)
        tab = O.ddim_tables(O.cosine_schedule(T)["alphas_cumprod"], O.ddim_timesteps(S, T), eta)
        ok = np.array_equal(smp.ddim_timesteps, tab["ddim_timesteps"]) and \
            eq(smp.ddim_alphas, tab["ddim_alphas"]) and \
            np.array_equal(smp.ddim_alphas_prev, tab["ddim_alphas_prev"]) and \
            eq(torch.as_tensor(smp.ddim_sigmas), torch.as_tensor(tab["ddim_sigmas"])) and \
            eq(torch.as_tensor(smp.ddim_sqrt_one_minus_alphas),
               torch.as_tensor(tab["ddim_sqrt_one_minus_alphas"]))
        assert ok, f"ddim tables S={S} T={T} differ"
        save(f"ddim_tables_S{S}_T{T}_eta{eta}", "bit-exact",
             ddim_timesteps=smp.ddim_timesteps, ddim_alphas=smp.ddim_alphas,
             ddim_alphas_prev=smp.ddim_alphas_prev,
             ddim_sigmas=np.asarray(smp.ddim_sigmas, dtype=np.float64),
             ddim_sqrt_one_minus_alphas=np.asarray(smp.ddim_sqrt_one_minus_alphas))

    # ---- sampler arithmetic alone, full T=1000, stub eps ------------------------------
    for clipped in (True, False):
        T, n, size = 1000, 2, 8
        d = RefEODiffusion(Stub(), size, 3, timesteps=T, cond_type="sum")
        x_T, tape = O.noise_tape((n, 3, size, size), T, seed=77)
        cond = O.synth_cond_sum(n, size, seed=5)
        with replay([x_T], tape):
            ref = d.sampling(n, clipped_reverse_diffusion=clipped, cond=cond)
        got = O.ddpm_sample(None, None, O.cosine_schedule(T), x_T, tape, cond=cond,
                            cond_type="sum", clipped=clipped, eps_fn=stub_eps)
        assert eq(ref, got), "stub trajectory differs"
        save(f"stub_ddpm_sum_T1000_clip{int(clipped)}", "bit-exact", cond=cond, x0=ref,
             tape_seed=np.int64(77), n=np.int64(n), size=np.int64(size))
    # no conditioning, full T
    d = RefEODiffusion(Stub(), 8, 3, timesteps=1000, cond_type=None)
    x_T, tape = O.noise_tape((2, 3, 8, 8), 1000, seed=78)
    with replay([x_T], tape):
        ref = d.sampling(2)
    got = O.ddpm_sample(None, None, O.cosine_schedule(1000), x_T, tape, eps_fn=stub_eps)
    assert eq(ref, got)
    save("stub_ddpm_none_T1000_clip1", "bit-exact", x0=ref, tape_seed=np.int64(78),
         n=np.int64(2), size=np.int64(8))
    # DDIM with stub
    for eta in (0.0, 0.5):
        T, S, n, size = 1000, 50, 2, 8
        d = RefEODiffusion(Stub(), size, 3, timesteps=T)
        smp = RefDDIMSampler(d)
        x_T, tape = O.noise_tape((n, 3, size, size), S, seed=79)
        with replay(tape, [None] * S):
            ref, inter = smp.sample(S, n, (3, size, size), eta=eta, x_T=x_T, verbose=False)
        got, ginter = O.ddim_sample(None, None, O.cosine_schedule(T), S, x_T, tape, eta=eta,
                                    eps_fn=stub_eps)
        assert eq(ref, got) and len(inter["x_inter"]) == len(ginter["x_inter"])
        assert all(eq(a, b) for a, b in zip(inter["pred_x0"], ginter["pred_x0"]))
        save(f"stub_ddim_S50_T1000_eta{eta}", "bit-exact", x0=ref, tape_seed=np.int64(79),
             n=np.int64(n), size=np.int64(size),
             pred_x0_last=inter["pred_x0"][-1], n_inter=np.int64(len(inter["x_inter"])))

    # ---- UNet eps: tiny / small / BASELINE arch ------------------------------------------
    def eps_case(name, cfg, B, ts, cond_ch=0, seed=11):
        m = build_ref_unet(cfg)
        sd = {k: v.detach() for k, v in m.state_dict().items()}
        g = torch.Generator().manual_seed(seed)
        size = cfg["image_size"]
        xin = cfg["in_channels"] - cond_ch
        x = torch.randn((B, xin, size, size), generator=g)
        cond = torch.rand((B, cond_ch, size, size), generator=g) if cond_ch else None
        t = torch.tensor(ts, dtype=torch.long)
        with torch.no_grad():
            ref = m(x, t, cond=cond)
        got = O.unet_forward(sd, O.full_cfg(**cfg), x, t, cond=cond)
        pinned = "bit-exact" if eq(ref, got) else f"rel_l2={O.rel_l2(got, ref):.3e}"
        assert O.rel_l2(got, ref) < 1e-6, pinned
        arrs = dict(x=x, t=t, eps=ref, cfg=json.dumps(cfg), init_seed=np.int64(INIT_SEED),
                    dezero_seed=np.int64(DEZERO_SEED), wsum=json.dumps(weight_checksum(sd)))
        if cond is not None:
            arrs["cond"] = cond
        save(name, pinned, **arrs)
        return m, sd

    eps_case("tiny_eps", TINY, 2, [999, 0])
    eps_case("tiny_eps_b3", TINY, 3, [1, 500, 37])
    eps_case("tiny_concat_eps", dict(TINY, in_channels=5), 2, [250, 3], cond_ch=2)
    eps_case("small_eps", SMALL, 2, [999, 1])
    eps_case("small_ms_concat_eps", dict(SMALL, in_channels=28, out_channels=13), 1, [400], cond_ch=15)
    # the switches of the reference's UNet / UNetBig / UNetSmall factories (unet_openai.py:783-922)
    eps_case("tiny_film_eps", dict(TINY, use_scale_shift_norm=True), 2, [999, 3])
    eps_case("tiny_updown_eps", dict(TINY, resblock_updown=True), 2, [700, 0])
    eps_case("tiny_film_updown_eps", dict(TINY, use_scale_shift_norm=True, resblock_updown=True), 2, [250, 9])
    eps_case("small_film_updown_eps", dict(SMALL, use_scale_shift_norm=True, resblock_updown=True, num_head_channels=32,
                                           use_new_attention_order=True), 2, [999, 1])
    if ONLY:
        path = os.path.join(GOLD, "MANIFEST.json")
        with open(path) as f:
            old = json.load(f)
        old["cases"].update(manifest["cases"])
        with open(path, "w") as f:
            json.dump(old, f, indent=1, sort_keys=True)
        print("done (only:", ONLY + ")")
        return
    m_base, sd_base = eps_case("base64_eps", BASE, 1, [999])
    nparam = sum(p.numel() for p in m_base.parameters())
    assert nparam == 88220934, nparam  # EO_Diffusion.ipynb:151
    manifest["param_count_base"] = nparam
    for tval in (750, 500, 250, 1, 0):    # with 999 above: the timesteps SURVEY.md 8(d) lists
        g = torch.Generator().manual_seed(100 + tval)
        x = torch.randn((1, 3, 64, 64), generator=g)
        t = torch.tensor([tval])
        with torch.no_grad():
            ref = m_base(x, t)
        got = O.unet_forward(sd_base, O.full_cfg(**BASE), x, t)
        assert O.rel_l2(got, ref) < 1e-6
        save(f"base64_eps_t{tval}", "bit-exact" if eq(ref, got) else "rel_l2<1e-6", x=x, t=t,
             eps=ref, cfg=json.dumps(BASE), init_seed=np.int64(INIT_SEED),
             dezero_seed=np.int64(DEZERO_SEED), wsum=json.dumps(weight_checksum(sd_base)))

    # ---- trajectories through the reference's sampling() -------------------------------------
    def ddpm_case(name, cfg, m, sd, T, n, cond_type, clipped, tape_seed, keep=(0,)):
        size = cfg["image_size"]
        d = RefEODiffusion(m, size, 3, timesteps=T, cond_type=cond_type).eval()
        x_T, tape = O.noise_tape((n, 3, size, size), T, seed=tape_seed)
        cond = O.synth_cond_sum(n, size, seed=tape_seed + 1) if cond_type == "sum" else None
        rec = []
        h = hook_eps(d, rec)
        with replay([x_T], tape):
            ref = d.sampling(n, clipped_reverse_diffusion=clipped, cond=cond)
        h.remove()
        grec = []
        got = O.ddpm_sample(sd, O.full_cfg(**cfg), O.cosine_schedule(T), x_T, tape, cond=cond,
                            cond_type=cond_type, clipped=clipped, record=grec)
        exact = eq(ref, got) and all(eq(a[2], b[2]) and a[0] == b[0] for a, b in zip(rec, grec))
        assert O.rel_l2(got, ref) < 1e-5
        arrs = dict(x0=ref, tape_seed=np.int64(tape_seed), n=np.int64(n), T=np.int64(T),
                    cfg=json.dumps(cfg), init_seed=np.int64(INIT_SEED),
                    dezero_seed=np.int64(DEZERO_SEED), wsum=json.dumps(weight_checksum(sd)),
                    t_seq=np.asarray([r[0] for r in rec], dtype=np.int64))
        if cond is not None:
            arrs["cond"] = cond
        for k in keep:
            arrs[f"eps_step{k}"] = rec[k][2]
            arrs[f"xt_step{k}"] = rec[k][1]
        save(name, "bit-exact" if exact else "rel_l2<1e-5", **arrs)

    m_tiny = build_ref_unet(TINY)
    sd_tiny = {k: v.detach() for k, v in m_tiny.state_dict().items()}
    ddpm_case("tiny_ddpm_sum_T8", TINY, m_tiny, sd_tiny, 8, 2, "sum", True, 21, keep=(0, 3, 7))
    ddpm_case("tiny_ddpm_none_T8_noclip", TINY, m_tiny, sd_tiny, 8, 2, None, False, 23, keep=(0, 7))
    ddpm_case("base64_ddpm_sum_T20", BASE, m_base, sd_base, 20, 1, "sum", True, 31, keep=(0, 10, 19))

    # DDIM through the reference sampler with the tiny UNet
    for eta in (0.0, 0.5):
        T, S, n = 8, 4, 2
        d = RefEODiffusion(m_tiny, 16, 3, timesteps=T).eval()
        smp = RefDDIMSampler(d)
        x_T, tape = O.noise_tape((n, 3, 16, 16), S, seed=41)
        with replay(tape, [None] * S):
            ref, inter = smp.sample(S, n, (3, 16, 16), eta=eta, x_T=x_T, verbose=False,
                                    log_every_t=1)
        got, ginter = O.ddim_sample(sd_tiny, O.full_cfg(**TINY), O.cosine_schedule(T), S, x_T,
                                    tape, eta=eta, log_every_t=1)
        assert O.rel_l2(got, ref) < 1e-5
        save(f"tiny_ddim_S4_T8_eta{eta}", "bit-exact" if eq(ref, got) else "rel_l2<1e-5",
             x0=ref, tape_seed=np.int64(41), n=np.int64(n), cfg=json.dumps(TINY),
             init_seed=np.int64(INIT_SEED), dezero_seed=np.int64(DEZERO_SEED),
             wsum=json.dumps(weight_checksum(sd_tiny)), pred_x0_last=inter["pred_x0"][-1])

    with open(os.path.join(GOLD, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("done")


if __name__ == "__main__":
    main()
