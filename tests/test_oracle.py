"""The CPU oracle against the golden vectors generated from the live reference
(oracle/make_golden.py), and -- when /root/reference is present -- against the reference
itself.  No GPU."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLD, REFERENCE, golden, golden_cfg, tt, weight_checksum
from oracle import oracle as O


def _ref_unet(cfg, init_seed, dezero_seed):
    """The golden weights, rebuilt with the drop-in module tree (bit-identical state dict to
    the reference under the same seeds; tests/test_host.py checks that claim)."""
    from conftest import build_unet
    return build_unet(cfg, init_seed, dezero_seed)


@pytest.mark.parametrize("T", [8, 20, 1000])
def test_schedule_bit_exact(T):
    g = golden(f"schedule_T{T}")
    s = O.cosine_schedule(T)
    for k in s:
        assert np.array_equal(s[k].numpy(), g[k]), k


def test_schedule_published_endpoints():
    # the constants the survey probe recorded from the reference
    s = O.cosine_schedule(1000)
    assert abs(float(s["betas"][0]) - 4.1246e-05) < 1e-9
    assert float(s["betas"][999]) == pytest.approx(0.999)
    assert float(s["alphas_cumprod"][999]) == pytest.approx(2.428888645766847e-09, rel=1e-6)


@pytest.mark.parametrize("S,T,eta", [(50, 1000, 0.0), (50, 1000, 0.5), (4, 8, 0.0), (4, 8, 0.5), (8, 8, 0.0)])
def test_ddim_tables_bit_exact(S, T, eta):
    g = golden(f"ddim_tables_S{S}_T{T}_eta{eta}")
    ts = O.ddim_timesteps(S, T)
    tab = O.ddim_tables(O.cosine_schedule(T)["alphas_cumprod"], ts, eta)
    assert np.array_equal(ts, g["ddim_timesteps"])
    assert np.array_equal(tab["ddim_alphas"].numpy(), g["ddim_alphas"])
    assert np.array_equal(np.asarray(tab["ddim_alphas_prev"]), g["ddim_alphas_prev"])
    assert np.array_equal(np.asarray(tab["ddim_sigmas"], dtype=np.float64), g["ddim_sigmas"])
    assert np.array_equal(np.asarray(tab["ddim_sqrt_one_minus_alphas"]), g["ddim_sqrt_one_minus_alphas"])
    if S == 50:
        assert ts[0] == 1 and ts[-1] == 981     # F8: +1 offset


def _stub(x, t, cond, y):
    return 0.3 * x - 0.1 + 1e-3 * t.float().reshape(-1, 1, 1, 1)


@pytest.mark.parametrize("clip", [1, 0])
def test_stub_ddpm_sum_full_T(clip):
    g = golden(f"stub_ddpm_sum_T1000_clip{clip}")
    n, size = int(g["n"]), int(g["size"])
    x_T, tape = O.noise_tape((n, 3, size, size), 1000, seed=int(g["tape_seed"]))
    out = O.ddpm_sample(None, None, O.cosine_schedule(1000), x_T, tape, cond=tt(g["cond"]),
                        cond_type="sum", clipped=bool(clip), eps_fn=_stub)
    assert torch.equal(out, tt(g["x0"]))


def test_stub_ddpm_uncond_full_T():
    g = golden("stub_ddpm_none_T1000_clip1")
    x_T, tape = O.noise_tape((2, 3, 8, 8), 1000, seed=int(g["tape_seed"]))
    out = O.ddpm_sample(None, None, O.cosine_schedule(1000), x_T, tape, eps_fn=_stub)
    assert torch.equal(out, tt(g["x0"]))


@pytest.mark.parametrize("eta", [0.0, 0.5])
def test_stub_ddim(eta):
    g = golden(f"stub_ddim_S50_T1000_eta{eta}")
    n, size = int(g["n"]), int(g["size"])
    x_T, tape = O.noise_tape((n, 3, size, size), 50, seed=int(g["tape_seed"]))
    out, inter = O.ddim_sample(None, None, O.cosine_schedule(1000), 50, x_T, tape, eta=eta, eps_fn=_stub)
    assert torch.equal(out, tt(g["x0"]))
    assert torch.equal(inter["pred_x0"][-1], tt(g["pred_x0_last"]))
    assert len(inter["x_inter"]) == int(g["n_inter"])


@pytest.mark.parametrize("name", ["tiny_eps", "tiny_eps_b3", "tiny_concat_eps", "small_eps",
                                  "small_ms_concat_eps", "base64_eps", "base64_eps_t750", "base64_eps_t250",
                                  "tiny_film_eps", "tiny_updown_eps",
                                  "tiny_film_updown_eps", "small_film_updown_eps"])
def test_unet_eps_vs_golden(name):
    g = golden(name)
    cfg = golden_cfg(g)
    m = _ref_unet(cfg, int(g["init_seed"]), int(g["dezero_seed"]))
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    assert weight_checksum(sd) == json.loads(str(g["wsum"]))["sha256"], "fixture weights not reproduced"
    cond = tt(g["cond"]) if "cond" in g else None
    eps = O.unet_forward(sd, O.full_cfg(**cfg), tt(g["x"]), tt(g["t"]), cond=cond)
    # same ATen ops in the same order: bit-exact on the machine that made the fixtures,
    # <= 1e-6 relative if oneDNN picks another kernel on a different host CPU
    assert O.rel_l2(eps, tt(g["eps"])) < 1e-6


def test_tiny_trajectory_vs_golden():
    g = golden("tiny_ddpm_sum_T8")
    cfg = golden_cfg(g)
    m = _ref_unet(cfg, int(g["init_seed"]), int(g["dezero_seed"]))
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    n, T = int(g["n"]), int(g["T"])
    x_T, tape = O.noise_tape((n, 3, cfg["image_size"], cfg["image_size"]), T, seed=int(g["tape_seed"]))
    rec = []
    out = O.ddpm_sample(sd, O.full_cfg(**cfg), O.cosine_schedule(T), x_T, tape, cond=tt(g["cond"]),
                        cond_type="sum", clipped=True, record=rec)
    assert [r[0] for r in rec] == list(g["t_seq"])          # timestep sequence: exact
    assert O.rel_l2(rec[3][2], tt(g["eps_step3"])) < 1e-5
    assert O.rel_l2(out, tt(g["x0"])) < 1e-5


def test_param_count_known_answer():
    # EO_Diffusion.ipynb:151 prints 88.220934 M
    with open(os.path.join(GOLD, "MANIFEST.json")) as f:
        assert json.load(f)["param_count_base"] == 88220934


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")
def test_oracle_vs_live_reference_tiny():
    sys.path.insert(0, REFERENCE)
    try:
        from backbones.unet_openai import UNetModel as RefUNet
    finally:
        sys.path.remove(REFERENCE)
    cfg = dict(image_size=16, in_channels=3, model_channels=32, out_channels=3, num_res_blocks=1,
               attention_resolutions=[2], channel_mult=[1, 2], num_heads=2)
    torch.manual_seed(7)
    m = O.dezero_(RefUNet(**cfg), 8).eval()
    x = torch.randn(2, 3, 16, 16)
    t = torch.tensor([3, 700])
    with torch.no_grad():
        ref = m(x, t)
    got = O.unet_forward({k: v.detach() for k, v in m.state_dict().items()}, O.full_cfg(**cfg), x, t)
    assert torch.equal(ref, got)


# ---- round-2 fixtures (oracle/make_golden_r2.py) -----------------------------------------------------
@pytest.mark.parametrize("name", ["heads1_eps", "base64_ms_concat_eps"])
def test_round2_eps_goldens(name):
    """num_heads = 1 (head dimensions > 64, the reference scripts' own setting) and the 13 + 15 -> 13 channel stem / head
    at base width: the oracle reproduces the live reference's eps from the seeds the fixture stores."""
    g = golden(name)
    cfg = golden_cfg(g)
    m = _ref_unet(cfg, int(g["init_seed"]), int(g["dezero_seed"]))
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    assert weight_checksum(sd) == json.loads(str(g["wsum"]))["sha256"]
    gen = torch.Generator().manual_seed(int(g["x_seed"]))
    B, cc, size = len(g["t"]), int(g["cond_ch"]), cfg["image_size"]
    x = torch.randn((B, cfg["in_channels"] - cc, size, size), generator=gen)
    cond = torch.rand((B, cc, size, size), generator=gen) if cc else None
    got = O.unet_forward(sd, O.full_cfg(**cfg), x, tt(g["t"]), cond=cond)
    assert O.rel_l2(got, tt(g["eps"])) < 1e-6


def test_round2_cfg_ddim_golden():
    """Classifier-free guidance (ddim.py:176-181) in the oracle against the live reference's DDIMSampler.sample."""
    g = golden("tiny_cfg_ddim_S4_T8")
    cfg = golden_cfg(g)
    m = _ref_unet(cfg, int(g["init_seed"]), int(g["dezero_seed"]))
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    n, S = int(g["n"]), 4
    x_T, tape = O.noise_tape((n, 3, 16, 16), S, seed=int(g["tape_seed"]))
    cond = tt(g["cond"])
    got, inter = O.ddim_sample(sd, O.full_cfg(**cfg), O.cosine_schedule(8), S, x_T, tape, eta=float(g["eta"]), cond=cond,
                               log_every_t=1, unconditional_guidance_scale=float(g["scale"]),
                               unconditional_conditioning=torch.zeros_like(cond))
    assert O.rel_l2(got, tt(g["x0"])) < 1e-5
    assert O.rel_l2(inter["pred_x0"][-1], tt(g["pred_x0_last"])) < 1e-5


def test_round2_manifest_pins():
    with open(os.path.join(GOLD, "MANIFEST.json")) as f:
        cases = json.load(f)["cases"]
    for name in ("base128_eps_b2", "base256_eps_b2", "base64_ms_concat_eps", "heads1_eps", "base64_ddpm_sum_T1000",
                 "tiny_cfg_ddim_S4_T8", "tiny_forward_train"):
        assert cases[name]["oracle_vs_reference"] == "bit-exact", name
