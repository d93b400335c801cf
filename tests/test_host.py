"""Host-side logic and the C-ABI surface, on CPU (no compute calls into the library)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

from conftest import REFERENCE, ROOT, golden, tt

from eo_diffusion_b200 import DDIMSampler, EODiffusion, UNetModel, _lib
from eo_diffusion_b200.diffusion import ddpm_coef_table
from oracle import oracle as O

TINY = dict(image_size=16, in_channels=3, model_channels=32, out_channels=3, num_res_blocks=1,
            attention_resolutions=[2], channel_mult=[1, 2], num_heads=2)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "eo_b200.h")).read()
    declared = set(re.findall(r"^EO_API [\w\s\*]+?\b(eo_\w+)\(", hdr, flags=re.M))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert _lib.lib().eo_version() >= 100


def test_no_cpu_fallback_is_loud():
    m = UNetModel(**TINY)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 16, 16), torch.zeros(1, dtype=torch.long))
    d = EODiffusion(m, 16, 3, timesteps=8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d.sampling(1, device="cpu")
    with pytest.raises(AssertionError, match="class-conditional"):
        m(torch.zeros(1, 3, 16, 16), torch.zeros(1, dtype=torch.long), y=torch.zeros(1, dtype=torch.long))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "eo_diffusion_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_unsupported_ctor_options_raise():
    for kw in (dict(dims=3), dict(conv_resample=False)):
        with pytest.raises(NotImplementedError):
            UNetModel(**dict(TINY, **kw))


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")
@pytest.mark.parametrize("extra", [{}, dict(in_channels=5, num_classes=7, num_head_channels=16,
                                            use_new_attention_order=True),
                                   dict(use_scale_shift_norm=True, resblock_updown=True)])
def test_state_dict_identical_to_reference(extra):
    sys.path.insert(0, REFERENCE)
    try:
        from backbones.unet_openai import UNetModel as RefUNet
    finally:
        sys.path.remove(REFERENCE)
    cfg = dict(TINY, **extra)
    torch.manual_seed(1234)
    a = RefUNet(**cfg)
    torch.manual_seed(1234)
    b = UNetModel(**cfg)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    assert all(torch.equal(sa[k], sb[k]) for k in sa)
    b.load_state_dict(sa, strict=True)
    for attr in ("in_channels", "model_channels", "out_channels", "image_size", "num_classes", "dtype"):
        assert getattr(a, attr) == getattr(b, attr)


def test_schedule_buffers_match_golden():
    g = golden("schedule_T1000")
    d = EODiffusion(torch.nn.Identity(), 8, 3, timesteps=1000, cond_type="sum")
    for k in ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
        assert np.array_equal(getattr(d, k).numpy(), g[k]), k
    assert set(dict(d.named_buffers())) == {"betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod",
                                            "sqrt_one_minus_alphas_cumprod"}
    assert (d.timesteps, d.in_channels, d.image_size, d.cond_type, d.device) == (1000, 3, 8, "sum", "cpu")


def test_coef_table_reproduces_reference_scalars():
    """Every column equals what the reference computes per step from gathered values."""
    s = O.cosine_schedule(1000)
    tab = ddpm_coef_table(s["betas"], s["alphas"], s["alphas_cumprod"], s["sqrt_alphas_cumprod"],
                          s["sqrt_one_minus_alphas_cumprod"])
    for t in (999, 500, 1):
        tt_ = torch.tensor([t])
        a, acp, b = s["alphas"][tt_], s["alphas_cumprod"][tt_], s["betas"][tt_]
        prev = s["alphas_cumprod"][tt_ - 1]
        assert tab[t, 2] == torch.sqrt(1. / acp)
        assert tab[t, 3] == torch.sqrt(1. / acp - 1.)
        assert tab[t, 4] == (b * torch.sqrt(prev) / (1. - acp))
        assert tab[t, 5] == ((1. - prev) * torch.sqrt(a) / (1. - acp))
        assert tab[t, 6] == torch.sqrt(b * (1. - prev) / (1. - acp))
        assert tab[t, 8] == 1. / torch.sqrt(a)
        assert tab[t, 9] == (1.0 - a) / s["sqrt_one_minus_alphas_cumprod"][tt_]
    assert tab[0, 7] == s["betas"][0] / (1. - s["alphas_cumprod"][0])


@pytest.mark.parametrize("S,T,eta", [(50, 1000, 0.0), (50, 1000, 0.5), (4, 8, 0.5), (8, 8, 0.0)])
def test_ddim_make_schedule_matches_golden(S, T, eta):
    g = golden(f"ddim_tables_S{S}_T{T}_eta{eta}")
    d = EODiffusion(torch.nn.Identity(), 8, 3, timesteps=T)
    smp = DDIMSampler(d)
    smp.make_schedule(S, ddim_eta=eta, verbose=False)
    assert np.array_equal(smp.ddim_timesteps, g["ddim_timesteps"])
    assert np.array_equal(smp.ddim_alphas.cpu().numpy(), g["ddim_alphas"])
    assert isinstance(smp.ddim_alphas_prev, np.ndarray) and smp.ddim_alphas_prev.dtype == np.float64
    assert np.array_equal(smp.ddim_alphas_prev, g["ddim_alphas_prev"])
    assert np.array_equal(np.asarray(smp.ddim_sigmas, dtype=np.float64), g["ddim_sigmas"])
    assert np.array_equal(np.asarray(smp.ddim_sqrt_one_minus_alphas), g["ddim_sqrt_one_minus_alphas"])


def test_ddim_broken_reference_branches_refuse():
    d = EODiffusion(torch.nn.Identity(), 8, 3, timesteps=8)
    smp = DDIMSampler(d)
    with pytest.raises(NotImplementedError):
        smp.sample(4, 1, (3, 8, 8), mask=torch.ones(1, 1, 8, 8), x0=torch.zeros(1, 3, 8, 8), verbose=False)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")
@pytest.mark.parametrize("factory,size", [("UNet", 32), ("UNetSmall", 64), ("UNetBig", 28)])
def test_factories_identical_to_reference(factory, size):
    """UNet / UNetBig / UNetSmall (unet_openai.py:783-922): same tables, flags and initial weights."""
    sys.path.insert(0, REFERENCE)
    try:
        from backbones import unet_openai as R
    finally:
        sys.path.remove(REFERENCE)
    from eo_diffusion_b200 import unet as U
    torch.manual_seed(7)
    a = getattr(R, factory)(size, num_classes=5)
    torch.manual_seed(7)
    b = getattr(U, factory)(size, num_classes=5)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    assert all(torch.equal(sa[k], sb[k]) for k in sa)
    with pytest.raises(ValueError, match="unsupported image size"):
        getattr(U, factory)(48)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line with the
    contract's keys; here on the reference's own 64x64 batch-1 case, one sampler step."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["metric"] == "ddpm_T1000_cloud_removal_images_per_sec" and d["value"] > 0
    # oracle/_ref (the reference's own classes, staged by oracle/build_ref.py) when present, else the oracle port
    staged = os.path.isfile(os.path.join(root, "oracle", "_ref", "diffusion", "model.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


def test_grid_geometry_query_matches_torchvision():
    """eo_post_grid_u8 with null buffers only answers the grid geometry (host arithmetic, no launch): it must be
    torchvision.utils.make_grid's."""
    import ctypes as C
    from torchvision.utils import make_grid
    from eo_diffusion_b200 import _lib
    L = _lib.lib()
    geo = (C.c_int * 3)()
    for B, Cc, H, W, nrow, pad in [(16, 3, 64, 64, 4, 2), (5, 3, 33, 47, 8, 2), (7, 3, 20, 24, 3, 2), (1, 3, 40, 24, 1, 2),
                                   (6, 1, 16, 16, 2, 1), (4, 3, 32, 32, 2, 0), (1, 1, 8, 8, 8, 2)]:
        assert L.eo_post_grid_u8(None, None, B, Cc, H, W, nrow, pad, 0.0, 0, geo, None) == 0
        want = make_grid(torch.zeros(B, Cc, H, W), nrow=nrow, padding=pad).shape
        assert (geo[2], geo[0], geo[1]) == tuple(want)
    assert L.eo_post_grid_u8(None, None, 0, 3, 8, 8, 1, 2, 0.0, 0, geo, None) < 0
    assert L.eo_post_grid_u8(None, None, 2, 3, 8, 8, 1, 2, 0.0, 2, geo, None) < 0
