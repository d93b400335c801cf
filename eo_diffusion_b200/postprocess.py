"""On-device post-processing and metrics of finished samples: what reference
`inference.py:128-150` does on the CPU/GPU with torch, torchvision and torchmetrics after `sampling()` returns
-- range mapping, the dimmed conditioning image, brightness adjustment, PSNR and SSIM, and the device half of
torchvision's `save_image` (grid + uint8 quantisation, so 3 bytes per pixel leave the GPU) -- as libeo_b200 kernels.
Function names follow the libraries the reference calls (`peak_signal_noise_ratio`,
`structural_similarity_index_measure`, `adjust_brightness`).  CUDA tensors only; there is no CPU fallback."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

__all__ = ["peak_signal_noise_ratio", "structural_similarity_index_measure", "adjust_brightness",
           "to_unit_range", "dim_masked", "tensor_stats", "postprocess_samples", "make_grid_u8", "save_image"]


def _f32(t, name):
    _lib.require_cuda_tensor(t, name)
    return t.detach().float().contiguous()


def _map(x, mode, factor=1.0):
    x = _f32(x, "x")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().eo_post_map(_lib.ptr(x), _lib.ptr(out), x.numel(), mode, float(factor), _lib.stream_ptr()),
                   "eo_post_map")
    return out


def tensor_stats(x) -> torch.Tensor:
    """[mean, min, max] of x as a 3-element fp32 device tensor (one pass over x)."""
    x = _f32(x, "x")
    ws = torch.empty(4, dtype=torch.float64, device=x.device)
    out = torch.empty(3, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().eo_post_stats(_lib.ptr(x), x.numel(), _lib.ptr(ws), _lib.ptr(out), _lib.stream_ptr()),
                   "eo_post_stats")
    return out


def to_unit_range(samples, image_min: float):
    """`samples.clip(0,1) if image.min()>=0 else (samples+1.)/2.` (inference.py:128)."""
    return _map(samples, 0 if image_min >= 0 else 1)


def adjust_brightness(img, brightness_factor: float):
    """torchvision.transforms.functional.adjust_brightness for float image tensors."""
    if brightness_factor < 0:
        raise ValueError(f"brightness_factor ({brightness_factor}) is not non-negative.")
    return _map(img, 2, brightness_factor)


def dim_masked(image, mask):
    """`image*((mask+0.7).clip(0,1))` with mask [B,1,H,W] (inference.py:135)."""
    image, mask = _f32(image, "image"), _f32(mask, "mask")
    B, Cc = image.shape[0], image.shape[1]
    hw = image[0, 0].numel()
    if mask.shape[0] != B or mask[0].numel() != hw:
        raise RuntimeError(f"mask {tuple(mask.shape)} does not broadcast over image {tuple(image.shape)}")
    out = torch.empty_like(image)
    with torch.cuda.device(image.device):
        _lib.check(_lib.lib().eo_post_dim_masked(_lib.ptr(image), _lib.ptr(mask), _lib.ptr(out), B, Cc, hw,
                                                 _lib.stream_ptr()), "eo_post_dim_masked")
    return out


def make_grid_u8(tensor, nrow: int = 8, padding: int = 2, pad_value: float = 0.0, signed: bool = False) -> torch.Tensor:
    """`torchvision.utils.make_grid(tensor, nrow, padding, pad_value=pad_value)` followed by `save_image`'s
    `mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8)`: a [GH, GW, C] uint8 DEVICE tensor,
    bit-identical to torchvision's.  `signed=True` maps (x + 1) / 2 first (model.py:63).  Accepts what
    `save_image` is given on this path: [B,C,H,W], [C,H,W] or [H,W] float tensors."""
    _lib.require_cuda_tensor(tensor, "tensor")
    if tensor.dim() == 2:
        tensor = tensor[None]
    if tensor.dim() == 3:
        tensor = tensor[None]
    if tensor.dim() != 4:
        raise ValueError(f"make_grid_u8 takes [B,C,H,W], [C,H,W] or [H,W]; got {tuple(tensor.shape)}")
    x = tensor.detach().float().contiguous()
    B, Cc, H, W = x.shape
    geo = (C.c_int * 3)()
    L = _lib.lib()
    args = (B, Cc, H, W, int(nrow), int(padding), float(pad_value), int(bool(signed)), geo)
    _lib.check(L.eo_post_grid_u8(None, None, *args, None), "eo_post_grid_u8")
    out = torch.empty((geo[0], geo[1], geo[2]), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(L.eo_post_grid_u8(_lib.ptr(x), _lib.ptr(out), *args, _lib.stream_ptr()), "eo_post_grid_u8")
    return out


def save_image(tensor, fp, format=None, nrow: int = 8, padding: int = 2, pad_value: float = 0.0, signed: bool = False) -> None:
    """`torchvision.utils.save_image(tensor, fp, nrow=...)` (inference.py:143-150, model.py:62-66) with the grid and
    the quantisation on the device: the uint8 grid is the only thing copied to the host; PIL writes the same file."""
    from PIL import Image
    grid = make_grid_u8(tensor, nrow=nrow, padding=padding, pad_value=pad_value, signed=signed)
    Image.fromarray(grid.cpu().numpy()).save(fp, format=format)


def peak_signal_noise_ratio(preds, target, data_range: float = 1.0) -> torch.Tensor:
    """torchmetrics.functional.peak_signal_noise_ratio(preds, target, data_range=...) (inference.py:138):
    0-dim fp32 device tensor."""
    preds, target = _f32(preds, "preds"), _f32(target, "target")
    if preds.shape != target.shape:
        raise RuntimeError("Predictions and targets are expected to have the same shape")
    ws = torch.empty(4, dtype=torch.float64, device=preds.device)
    out = torch.empty(1, dtype=torch.float32, device=preds.device)
    with torch.cuda.device(preds.device):
        _lib.check(_lib.lib().eo_psnr(_lib.ptr(preds), _lib.ptr(target), preds.numel(), float(data_range),
                                      _lib.ptr(ws), _lib.ptr(out), _lib.stream_ptr()), "eo_psnr")
    return out[0]


def _gaussian_window(kernel_size=11, sigma=1.5):
    # torchmetrics `_gaussian`, evaluated on the host in fp32 exactly as the library does
    dist = torch.arange(start=(1 - kernel_size) / 2, end=(1 + kernel_size) / 2, step=1, dtype=torch.float32)
    gauss = torch.exp(-torch.pow(dist / sigma, 2) / 2)
    return gauss / gauss.sum()


def structural_similarity_index_measure(preds, target, data_range: float = 1.0, reduction="elementwise_mean"):
    """torchmetrics.functional.structural_similarity_index_measure with its default window (inference.py:138).
    reduction 'elementwise_mean' -> 0-dim tensor, 'none' -> [B]."""
    preds, target = _f32(preds, "preds"), _f32(target, "target")
    if preds.shape != target.shape:
        raise RuntimeError("Predictions and targets are expected to have the same shape")
    if preds.dim() != 4:
        raise ValueError(f"Expected `preds` and `target` to have BxCxHxW shape. Got preds: {tuple(preds.shape)}")
    B, Cc, H, W = preds.shape
    g = (C.c_float * 11)(*[float(v) for v in _gaussian_window()])
    ws = torch.empty(B, dtype=torch.float64, device=preds.device)
    per = torch.empty(B, dtype=torch.float32, device=preds.device)
    mean = torch.empty(1, dtype=torch.float32, device=preds.device)
    with torch.cuda.device(preds.device):
        _lib.check(_lib.lib().eo_ssim(_lib.ptr(preds), _lib.ptr(target), B, Cc, H, W, float(data_range), g,
                                      _lib.ptr(ws), _lib.ptr(per), _lib.ptr(mean), _lib.stream_ptr()), "eo_ssim")
    return per if reduction in ("none", None) else mean[0]


def postprocess_samples(samples, image, mask=None, cond_type="sum", metrics=True) -> dict:
    """inference.py:128-150 for one batch, on the device (no PNG writes): returns `samples`, `gt`, `cond` as
    the reference would save them and, with `metrics`, `ssim` and `psnr` (0-dim device tensors).  The host
    branches of the reference (`image.min()>=0`, `x.mean()<0.2`) read one small statistics vector per tensor."""
    out = {}
    imin = float(tensor_stats(image)[1])
    samples = to_unit_range(samples, imin)
    if mask is not None or cond_type is not None:
        image = _f32(image, "image")
        cond = dim_masked(image, mask) if mask is not None else image
        gt, cond = (image, cond) if imin >= 0 else (_map(image, 1), _map(cond, 1))
        if metrics:
            out["ssim"] = structural_similarity_index_measure(samples, gt, data_range=1.0)
            out["psnr"] = peak_signal_noise_ratio(samples, gt, data_range=1.0)
        if float(tensor_stats(gt)[0]) < 0.2:
            gt = adjust_brightness(gt, 3)
        if cond_type != "sum" and float(tensor_stats(cond)[0]) < 0.2:
            cond = adjust_brightness(cond, 3)
        out["gt"], out["cond"] = gt, cond
    if samples.shape[0] == 1 and float(tensor_stats(samples)[0]) < 0.2:
        samples = adjust_brightness(samples, 3)
    out["samples"] = samples
    return out
