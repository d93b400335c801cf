"""Sampler-step kernels through the C ABI vs the CPU oracle: bit-exact (fp32, the
reference's op order).  Needs a B200."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import golden, replay, tt
from eo_diffusion_b200 import DDIMSampler, EODiffusion, _lib
from eo_diffusion_b200.diffusion import ddpm_coef_table
from oracle import oracle as O

pytestmark = pytest.mark.gpu


class Stub(torch.nn.Module):
    # the stand-in eps model of the stub_* fixtures (oracle/make_golden.py)
    def forward(self, x, t, cond=None, y=None):
        return 0.3 * x - 0.1 + 1e-3 * t.float().reshape(-1, 1, 1, 1)


def _table(dev, T=1000):
    s = O.cosine_schedule(T)
    tab = ddpm_coef_table(s["betas"], s["alphas"], s["alphas_cumprod"], s["sqrt_alphas_cumprod"],
                          s["sqrt_one_minus_alphas_cumprod"]).to(dev)
    return s, tab


@pytest.mark.parametrize("shape", [(2, 3, 8, 8), (3, 3, 5, 7), (1, 13, 16, 16), (4, 3, 64, 64)])
def test_sum_mix_and_steps_bit_exact(cuda_dev, shape):
    L = _lib.lib()
    s, tab = _table(cuda_dev)
    g = torch.Generator().manual_seed(sum(shape))
    n, c, h, w = shape
    x = torch.randn(shape, generator=g)
    eps = torch.randn(shape, generator=g)
    nz = torch.randn(shape, generator=g)
    nz2 = torch.randn(shape, generator=g)
    gt = torch.rand(shape, generator=g)
    mask = (torch.rand((n, 1, h, w), generator=g) > 0.5).float()
    for tval in (999, 500, 1, 0):
        t = torch.full((n,), tval, dtype=torch.long)
        d = lambda v: v.to(cuda_dev).contiguous()
        dx, de, dn, dn2, dg, dm, dt = d(x), d(eps), d(nz), d(nz2), d(gt), d(mask), d(t)
        out = torch.empty_like(dx)
        st = _lib.stream_ptr()
        # 'sum' mix (model.py:58-60)
        _lib.check(L.eo_ddpm_sum_mix(_lib.ptr(dx), _lib.ptr(dg), _lib.ptr(dm), _lib.ptr(dn), _lib.ptr(dt),
                                     _lib.ptr(tab), _lib.ptr(out), n, c, h * w, st))
        assert torch.equal(out.cpu(), O.sum_mix(s, x, gt, mask, t, nz)), f"mix t={tval}"
        # reverse steps (model.py:125-150 and :101-122)
        for clip, fn in ((1, O.reverse_step_clip), (0, O.reverse_step_noclip)):
            _lib.check(L.eo_ddpm_step(_lib.ptr(dx), _lib.ptr(de), _lib.ptr(dn), _lib.ptr(dt), _lib.ptr(tab),
                                      _lib.ptr(out), n, c, h * w, clip, int(tval > 0), st))
            assert torch.equal(out.cpu(), fn(s, x, t, nz, eps)), f"step clip={clip} t={tval}"
        # fused step + next iteration's mix
        if tval > 0:
            t2 = t - 1
            dt2 = d(t2)
            _lib.check(L.eo_ddpm_step_mix(_lib.ptr(dx), _lib.ptr(de), _lib.ptr(dn), _lib.ptr(dt), _lib.ptr(dg),
                                          _lib.ptr(dm), _lib.ptr(dn2), _lib.ptr(dt2), _lib.ptr(tab),
                                          _lib.ptr(out), n, c, h * w, 1, 1, st))
            want = O.sum_mix(s, O.reverse_step_clip(s, x, t, nz, eps), gt, mask, t2, nz2)
            assert torch.equal(out.cpu(), want), f"step_mix t={tval}"


def test_mixed_timesteps_per_sample(cuda_dev):
    """The kernels gather per-sample rows like the reference's .gather(-1, t)."""
    L = _lib.lib()
    s, tab = _table(cuda_dev)
    g = torch.Generator().manual_seed(3)
    shape = (4, 3, 8, 8)
    x, eps, nz = (torch.randn(shape, generator=g) for _ in range(3))
    t = torch.tensor([999, 3, 512, 1])
    out = torch.empty(shape, device=cuda_dev)
    dx, de, dn, dt = x.to(cuda_dev), eps.to(cuda_dev), nz.to(cuda_dev), t.to(cuda_dev)   # keep alive
    _lib.check(L.eo_ddpm_step(_lib.ptr(dx), _lib.ptr(de), _lib.ptr(dn), _lib.ptr(dt), _lib.ptr(tab),
                              _lib.ptr(out), 4, 3, 64, 1, 1, _lib.stream_ptr()))
    assert torch.equal(out.cpu(), O.reverse_step_clip(s, x, t, nz, eps))


@pytest.mark.parametrize("clip", [1, 0])
def test_ddpm_sum_full_trajectory_stub(cuda_dev, clip, tmp_path):
    """EODiffusion.sampling over the full T=1000 schedule with the reference's RNG order
    replayed: bit-exact against the trajectory the live reference produced."""
    g = golden(f"stub_ddpm_sum_T1000_clip{clip}")
    n, size = int(g["n"]), int(g["size"])
    d = EODiffusion(Stub(), size, 3, timesteps=1000, cond_type="sum").to(cuda_dev)
    x_T, tape = O.noise_tape((n, 3, size, size), 1000, seed=int(g["tape_seed"]))
    with replay([x_T], tape, tmp_cwd=str(tmp_path)):
        out = d.sampling(n, clipped_reverse_diffusion=bool(clip), device=cuda_dev, cond=tt(g["cond"]))
    assert torch.equal(out.cpu(), tt(g["x0"]))
    # the reference's unconditional PNG side effect (model.py:62-66) is reproduced
    assert (tmp_path / "results" / "prova" / "s0_200_pred.png").exists()
    assert (tmp_path / "results" / "prova" / "s0_gt.png").exists()
    assert not (tmp_path / "results" / "prova" / "s0_300_pred.png").exists()


def test_ddpm_uncond_full_trajectory_stub(cuda_dev, tmp_path):
    g = golden("stub_ddpm_none_T1000_clip1")
    d = EODiffusion(Stub(), 8, 3, timesteps=1000).to(cuda_dev)
    x_T, tape = O.noise_tape((2, 3, 8, 8), 1000, seed=int(g["tape_seed"]))
    with replay([x_T], tape, tmp_cwd=str(tmp_path)):
        out = d.sampling(2, device=cuda_dev, write_pngs=False)
    assert torch.equal(out.cpu(), tt(g["x0"]))
    assert not (tmp_path / "results" / "prova" / "s0_200_pred.png").exists()


def test_sampling_without_results_dir_raises_like_reference(cuda_dev, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    d = EODiffusion(Stub(), 8, 3, timesteps=250).to(cuda_dev)
    with pytest.raises((FileNotFoundError, OSError)):
        d.sampling(1, device=cuda_dev)


@pytest.mark.parametrize("eta", [0.0, 0.5])
def test_ddim_full_trajectory_stub(cuda_dev, eta):
    g = golden(f"stub_ddim_S50_T1000_eta{eta}")
    n, size = int(g["n"]), int(g["size"])
    d = EODiffusion(Stub(), size, 3, timesteps=1000).to(cuda_dev)
    smp = DDIMSampler(d)
    x_T, tape = O.noise_tape((n, 3, size, size), 50, seed=int(g["tape_seed"]))
    with replay(tape, [None] * 50):
        out, inter = smp.sample(50, n, (3, size, size), eta=eta, x_T=x_T.to(cuda_dev), verbose=False)
    assert np.array_equal(smp.ddim_timesteps[[0, -1]], [1, 981])
    assert torch.equal(out.cpu(), tt(g["x0"]))
    assert torch.equal(inter["pred_x0"][-1].cpu(), tt(g["pred_x0_last"]))
    assert len(inter["x_inter"]) == int(g["n_inter"])


def test_cfg_combine_and_forward_diffusion(cuda_dev):
    L = _lib.lib()
    g = torch.Generator().manual_seed(9)
    a, b = torch.randn(2, 3, 9, 9, generator=g), torch.randn(2, 3, 9, 9, generator=g)
    out = torch.empty_like(a, device=cuda_dev)
    da, db = a.to(cuda_dev), b.to(cuda_dev)
    _lib.check(L.eo_cfg_combine(_lib.ptr(da), _lib.ptr(db), 2.5, _lib.ptr(out), a.numel(), _lib.stream_ptr()))
    assert torch.equal(out.cpu(), a + 2.5 * (b - a))
    d = EODiffusion(Stub(), 9, 3, timesteps=1000).to(cuda_dev)
    t = torch.tensor([10, 900])
    got = d._forward_diffusion(a.to(cuda_dev), t.to(cuda_dev), b.to(cuda_dev))
    want = O.forward_diffusion(O.cosine_schedule(1000), a, t, b)
    # mask == 1 path: 1*q + 0*x == q exactly
    assert torch.equal(got.cpu(), want)


def test_full_size_mix_properties(cuda_dev):
    """BASELINE size (256x256, 64 images): where mask == 1 the mixed state is exactly
    q(gt | t); where mask == 0 it is exactly x_t (size-independent properties of the mix)."""
    L = _lib.lib()
    s, tab = _table(cuda_dev)
    n, c, h, w = 64, 3, 256, 256
    g = torch.Generator(device=cuda_dev).manual_seed(1)
    x = torch.randn((n, c, h, w), device=cuda_dev, generator=g)
    gt = torch.rand((n, c, h, w), device=cuda_dev, generator=g)
    nz = torch.randn((n, c, h, w), device=cuda_dev, generator=g)
    mask = (torch.rand((n, 1, h, w), device=cuda_dev, generator=g) > 0.5).float()
    t = torch.full((n,), 417, dtype=torch.long, device=cuda_dev)
    out = torch.empty_like(x)
    _lib.check(L.eo_ddpm_sum_mix(_lib.ptr(x), _lib.ptr(gt), _lib.ptr(mask), _lib.ptr(nz), _lib.ptr(t),
                                 _lib.ptr(tab), _lib.ptr(out), n, c, h * w, _lib.stream_ptr()))
    q = s["sqrt_alphas_cumprod"][417].item() * gt + s["sqrt_one_minus_alphas_cumprod"][417].item() * nz
    m = mask.expand(-1, c, -1, -1).bool()
    assert torch.equal(out[m], q[m])
    assert torch.equal(out[~m], x[~m])
