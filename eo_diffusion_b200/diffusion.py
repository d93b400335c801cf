"""Drop-in `EODiffusion`: cosine schedule + DDPM reverse sampler with the RePaint-style
'sum' conditioning mix, mirroring reference `diffusion/model.py::EODiffusion` (:12-150).

The loop is host code; all per-pixel arithmetic of a step runs in libeo_b200's fused
sampler kernels (`eo_ddpm_sum_mix`, `eo_ddpm_step`, `eo_ddpm_step_mix`), which evaluate the
reference's fp32 op sequence bit-exactly from per-timestep scalar tables that this module
computes with the reference's own torch ops.  The UNet call goes through `self.model`
(normally `eo_diffusion_b200.UNetModel`, i.e. libeo_b200 again)."""
from __future__ import annotations

import contextlib
import math
import os

import torch
import torch.nn as nn

from . import _lib

__all__ = ["EODiffusion", "ddpm_coef_table"]


def ddpm_coef_table(betas, alphas, alphas_cumprod, sqrt_acp, sqrt_1m_acp) -> torch.Tensor:
    """[T, EO_DDPM_NCOEF] fp32 table of the scalars that reference model.py:133-148 / :110-119
    derives from the gathered schedule values, computed with the same fp32 op order (each
    torch op rounds once, exactly like the reference's [n,1,1,1] tensors)."""
    T = betas.shape[0]
    b, a, acp = betas.float().cpu(), alphas.float().cpu(), alphas_cumprod.float().cpu()
    prev = torch.cat([acp[:1], acp[:-1]])      # acp[t-1]; row 0 is never used with t > 0
    tab = torch.zeros((T, _lib.EO_DDPM_NCOEF), dtype=torch.float32)
    tab[:, 0] = sqrt_acp.float().cpu()
    tab[:, 1] = sqrt_1m_acp.float().cpu()
    tab[:, 2] = torch.sqrt(1. / acp)
    tab[:, 3] = torch.sqrt(1. / acp - 1.)
    tab[:, 4] = b * torch.sqrt(prev) / (1. - acp)
    tab[:, 5] = (1. - prev) * torch.sqrt(a) / (1. - acp)
    tab[:, 6] = torch.sqrt(b * (1. - prev) / (1. - acp))
    tab[:, 7] = b / (1. - acp)
    tab[:, 8] = 1. / torch.sqrt(a)
    tab[:, 9] = (1.0 - a) / sqrt_1m_acp.float().cpu()
    return tab


class EODiffusion(nn.Module):
    """Same constructor, buffers and methods as the reference (model.py:13-36).

    Additive keyword arguments of `sampling` (defaults reproduce the reference):
      write_pngs : None -> the reference's PNG side effects (model.py:62-66: writes under
                   ./results/prova/ at i%25==0, i<=200 regardless of `save`); False -> none.
    """

    def __init__(self, model, image_size, in_channels, time_embedding_dim=256, timesteps=1000,
                 cond_type=None, device="cpu"):
        super().__init__()
        self.timesteps = timesteps
        self.in_channels = in_channels
        self.image_size = image_size
        self.cond_type = cond_type
        self.device = device

        betas = self._cosine_variance_schedule(timesteps)
        alphas = 1. - betas
        alphas_cumprod = torch.cumprod(alphas, dim=-1)
        self.register_buffer("betas", betas)
        self.register_buffer("alphas", alphas)
        self.register_buffer("alphas_cumprod", alphas_cumprod)
        self.register_buffer("sqrt_alphas_cumprod", torch.sqrt(alphas_cumprod))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", torch.sqrt(1. - alphas_cumprod))
        self.model = model
        self._tab_cache = None     # (key, device table)
        self._ts_cache = None      # (key, [T, n] int64 timesteps on device)

    # ------------------------------------------------------------------ schedule
    def _cosine_variance_schedule(self, timesteps, epsilon=0.008):
        # model.py:87-92 -- fp32 on the CPU, as in the reference constructor
        steps = torch.linspace(0, timesteps, steps=timesteps + 1, dtype=torch.float32)
        f_t = torch.cos(((steps / timesteps + epsilon) / (1.0 + epsilon)) * math.pi * 0.5) ** 2
        return torch.clip(1.0 - f_t[1:] / f_t[:timesteps], 0.0, 0.999)

    def _coef_table(self, device):
        key = (str(device), self.betas._version, self.alphas_cumprod._version, self.betas.data_ptr())
        if self._tab_cache is None or self._tab_cache[0] != key:
            tab = ddpm_coef_table(self.betas, self.alphas, self.alphas_cumprod,
                                  self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod)
            self._tab_cache = (key, tab.to(device).contiguous())
        return self._tab_cache[1]

    def _timestep_rows(self, n, device):
        key = (n, str(device), self.timesteps)
        if self._ts_cache is None or self._ts_cache[0] != key:
            rows = torch.arange(self.timesteps, dtype=torch.int64).reshape(-1, 1).expand(-1, n)
            self._ts_cache = (key, rows.contiguous().to(device))
        return self._ts_cache[1]

    # ------------------------------------------------------------------ reference methods
    def forward(self, x, noise, cond=None, y=None):
        # model.py:38-44 (training-side call; the UNet forward is the engine's)
        t = torch.randint(0, self.timesteps, (x.shape[0],)).to(x.device)
        x_t = self._forward_diffusion(x, t, noise)
        return self.model(x_t, t, cond=cond, y=y)

    def _forward_diffusion(self, x_0, t, noise):
        """q(x_t | x_0) (model.py:94-98) = the 'sum' mix kernel with mask == 1."""
        assert x_0.shape == noise.shape
        _lib.require_cuda_tensor(x_0, "x_0")
        dev = x_0.device
        n, c = x_0.shape[0], x_0.shape[1]
        hw = x_0[0, 0].numel()
        x0 = x_0.detach().float().contiguous()
        nz = noise.detach().to(dev).float().contiguous()
        ones = torch.ones((n, 1) + tuple(x_0.shape[2:]), dtype=torch.float32, device=dev)
        out = torch.empty_like(x0)
        ts = t.to(device=dev, dtype=torch.int64).contiguous()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().eo_ddpm_sum_mix(_lib.ptr(x0), _lib.ptr(x0), _lib.ptr(ones), _lib.ptr(nz),
                                                  _lib.ptr(ts), _lib.ptr(self._coef_table(dev)),
                                                  _lib.ptr(out), n, c, hw, _lib.stream_ptr()),
                       "eo_ddpm_sum_mix")
        return out

    def _step(self, x_t, pred, noise, t, clip, positive):
        dev = x_t.device
        n, c = x_t.shape[0], x_t.shape[1]
        hw = x_t[0, 0].numel()
        out = torch.empty_like(x_t)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().eo_ddpm_step(_lib.ptr(x_t), _lib.ptr(pred), _lib.ptr(noise), _lib.ptr(t),
                                               _lib.ptr(self._coef_table(dev)), _lib.ptr(out), n, c, hw,
                                               int(clip), int(positive), _lib.stream_ptr()),
                       "eo_ddpm_step")
        return out

    def _prep(self, x_t, t, noise):
        _lib.require_cuda_tensor(x_t, "x_t")
        dev = x_t.device
        return (x_t.detach().float().contiguous(), t.to(device=dev, dtype=torch.int64).contiguous(),
                noise.detach().to(dev).float().contiguous())

    @torch.no_grad()
    def _reverse_diffusion(self, x_t, t, noise, cond=None, y=None):
        """p(x_{t-1} | x_t) without clipping (model.py:101-122)."""
        x_t, t, noise = self._prep(x_t, t, noise)
        pred = self.model(x_t, t, cond=cond, y=y).float().contiguous()
        return self._step(x_t, pred, noise, t, clip=False, positive=bool(t.min() > 0))

    @torch.no_grad()
    def _reverse_diffusion_with_clip(self, x_t, t, noise, cond=None, y=None):
        """x0-prediction clipped to [-1, 1], then the posterior (model.py:125-150)."""
        x_t, t, noise = self._prep(x_t, t, noise)
        pred = self.model(x_t, t, cond=cond, y=y).float().contiguous()
        return self._step(x_t, pred, noise, t, clip=True, positive=bool(t.min() > 0))

    @torch.no_grad()
    def sampling(self, n_samples, clipped_reverse_diffusion=True, device="cpu", cond=None, y=None,
                 idx=0, save=False, write_pngs=None):
        """DDPM ancestral sampling (model.py:46-75).  RNG draw order is the reference's: one
        CPU `torch.randn` for x_T, then one `torch.randn_like` per step on `device`; in 'sum'
        mode the same noise forward-diffuses `gt` and drives the reverse step."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("eo_diffusion_b200.EODiffusion.sampling runs on CUDA (sm_100a) only; "
                               f"got device={device!r} (there is no CPU fallback)")
        L = _lib.lib()
        T, n = self.timesteps, n_samples
        x_t = torch.randn((n, self.in_channels, self.image_size, self.image_size)).to(dev)
        x_t = x_t.float().contiguous()
        sum_mode = self.cond_type == "sum"
        gt = mask = None
        if cond is not None and sum_mode:
            gt = cond[:n, :3].to(dev).float().contiguous()
            mask = cond[:n, 3][:, None].to(dev).float().contiguous()
            cond = None
        elif cond is not None:
            cond = cond.to(dev)
        if sum_mode and gt is None:
            raise RuntimeError("cond_type='sum' needs cond=[N,4,H,W] (gt in channels 0-2, mask in 3)")
        C_, hw = x_t.shape[1], x_t.shape[2] * x_t.shape[3]
        tab = self._coef_table(dev)
        ts_rows = self._timestep_rows(n, dev)
        clip = int(bool(clipped_reverse_diffusion))
        pngs = (write_pngs is None) or bool(write_pngs)

        # every UNet call of the loop uses a timestep value in [0, T): hoist the embedding path
        tables = getattr(self.model, "time_tables", None)
        hoist = tables(T) if (tables is not None and y is None) else contextlib.nullcontext()
        with torch.cuda.device(dev), hoist:
            st = _lib.stream_ptr
            noise = torch.randn_like(x_t).to(dev)
            if sum_mode:   # mix of the first iteration (model.py:58-60)
                _lib.check(L.eo_ddpm_sum_mix(_lib.ptr(x_t), _lib.ptr(gt), _lib.ptr(mask), _lib.ptr(noise),
                                             _lib.ptr(ts_rows[T - 1]), _lib.ptr(tab), _lib.ptr(x_t),
                                             n, C_, hw, st()), "eo_ddpm_sum_mix")
            for i in range(T - 1, -1, -1):
                t = ts_rows[i]
                if pngs and (i % 25 == 0 and i <= 200 or i % 100 == 0 and i <= T and save):
                    self._write_pngs(x_t, gt, mask, noise, t, i, idx, n)
                pred = self.model(x_t, t, cond=cond, y=y)
                if pred.dtype != torch.float32 or not pred.is_contiguous():
                    pred = pred.float().contiguous()
                # the next iteration's draw, made now so that its 'sum' mix fuses into this step
                noise_next = torch.randn_like(x_t).to(dev) if i > 0 else None
                if sum_mode and i > 0:
                    _lib.check(L.eo_ddpm_step_mix(_lib.ptr(x_t), _lib.ptr(pred), _lib.ptr(noise), _lib.ptr(t),
                                                  _lib.ptr(gt), _lib.ptr(mask), _lib.ptr(noise_next),
                                                  _lib.ptr(ts_rows[i - 1]), _lib.ptr(tab), _lib.ptr(x_t),
                                                  n, C_, hw, clip, 1, st()), "eo_ddpm_step_mix")
                else:
                    _lib.check(L.eo_ddpm_step(_lib.ptr(x_t), _lib.ptr(pred), _lib.ptr(noise), _lib.ptr(t),
                                              _lib.ptr(tab), _lib.ptr(x_t), n, C_, hw, clip, int(i > 0), st()),
                               "eo_ddpm_step")
                noise = noise_next
        return x_t

    def _write_pngs(self, x_t, gt, mask, noise, t, i, idx, n):
        # model.py:62-66 (the reference raises FileNotFoundError if ./results/prova is missing)
        from .postprocess import save_image     # torchvision's save_image, grid + quantisation on the device
        nrow = int(math.sqrt(n))
        save_image(x_t, f"results/prova/s{idx}_{i}_pred.png", nrow=nrow, signed=True)
        if self.cond_type == "sum":
            gt_noised = self._forward_diffusion(gt, t, noise)
            save_image(mask * gt_noised, f"results/prova/s{i}_masked.png", nrow=nrow)
            save_image(gt_noised, f"results/prova/s{i}_gt.png")
