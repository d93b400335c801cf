"""CPU restatement of the post-processing and metrics that follow the sampling path -- TEST INFRASTRUCTURE
ONLY (imported by tests/; never by the product).

Follows reference inference.py:128-150.  `adjust_brightness` is torchvision's (pinned against the installed
torchvision in tests/test_postprocess.py).  The two metrics come from torchmetrics, a third-party dependency
that is NOT vendored in the reference checkout and not installed here (reference pin: eo_diffusion.yml
`torchmetrics==0.11.0`): their published algorithms (torchmetrics/functional/image/{psnr,ssim}.py) are restated
below with the same torch op sequence.  Pinning of `psnr` and `ssim`: no live torchmetrics output exists in this
image, so they are held to (1) the known-answer examples torchmetrics publishes in the docstrings of those two
functions (PSNR `tensor(2.5527)`; SSIM of `rand([3,3,256,256])` against `0.75 *` itself `tensor(0.9219)`, a value
that does not depend on the seed to six digits), (2) OpenCV's `cv2.PSNR`, (3) an independent scipy evaluation
of the SSIM definition, (4) analytic cases (tests/test_postprocess.py).  That is a four-digit pin on published
vectors, NOT a bit-level pin against the library: treat anything finer than 1e-4 as PARITY UNPINNED."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor


def to_unit_range(samples: Tensor, image_min: float) -> Tensor:
    """inference.py:128: `samples.clip(0,1) if image.min()>=0 else (samples+1.)/2.`"""
    return samples.clip(0, 1) if image_min >= 0 else (samples + 1.) / 2.


def dim_masked(image: Tensor, mask: Tensor) -> Tensor:
    """inference.py:135: `cond = image*((mask+0.7).clip(0,1))`"""
    return image * ((mask + 0.7).clip(0, 1))


def adjust_brightness(img: Tensor, factor: float) -> Tensor:
    """torchvision.transforms.functional.adjust_brightness for float tensors:
    `_blend(img, zeros_like(img), factor)` = `(factor * img + (1 - factor) * 0).clamp(0, 1)`."""
    return (factor * img + (1.0 - factor) * torch.zeros_like(img)).clamp(0, 1.0)


def psnr(preds: Tensor, target: Tensor, data_range: float = 1.0) -> Tensor:
    """torchmetrics.functional.peak_signal_noise_ratio(preds, target, data_range=...), base 10, dim None,
    reduction 'elementwise_mean': one value over the whole batch."""
    sum_squared_error = torch.sum(torch.pow(preds - target, 2))
    n_obs = torch.tensor(target.numel())
    dr = torch.tensor(float(data_range))
    psnr_base_e = 2 * torch.log(dr) - torch.log(sum_squared_error / n_obs)
    return psnr_base_e * (10 / torch.log(torch.tensor(10.0)))


def gaussian_window(kernel_size: int = 11, sigma: float = 1.5) -> Tensor:
    """torchmetrics `_gaussian`: fp32, normalised."""
    dist = torch.arange(start=(1 - kernel_size) / 2, end=(1 + kernel_size) / 2, step=1, dtype=torch.float32)
    gauss = torch.exp(-torch.pow(dist / sigma, 2) / 2)
    return gauss / gauss.sum()


def ssim(preds: Tensor, target: Tensor, data_range: float = 1.0, per_image: bool = False) -> Tensor:
    """torchmetrics.functional.structural_similarity_index_measure with its defaults (gaussian_kernel=True,
    sigma=1.5, kernel_size=11, k1=0.01, k2=0.03, reduction='elementwise_mean') -- `_ssim_update`:
    reflect-pad by 5, depthwise conv with the outer-product Gaussian, SSIM map, crop the padded border."""
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    channel = preds.size(1)
    g = gaussian_window()
    kernel = (g[:, None] @ g[None, :]).expand(channel, 1, 11, 11)
    pad = 5
    p = F.pad(preds, (pad, pad, pad, pad), mode="reflect")
    t = F.pad(target, (pad, pad, pad, pad), mode="reflect")
    inp = torch.cat((p, t, p * p, t * t, p * t))
    out = F.conv2d(inp, kernel, groups=channel).split(preds.shape[0])
    mu_p2, mu_t2, mu_pt = out[0].pow(2), out[1].pow(2), out[0] * out[1]
    s_p, s_t, s_pt = out[2] - mu_p2, out[3] - mu_t2, out[4] - mu_pt
    upper = 2 * s_pt + c2
    lower = s_p + s_t + c2
    full = ((2 * mu_pt + c1) * upper) / ((mu_p2 + mu_t2 + c1) * lower)
    idx = full[..., pad:-pad, pad:-pad]
    vals = idx.reshape(idx.shape[0], -1).mean(-1)
    return vals if per_image else vals.mean()


def postprocess(samples: Tensor, image: Tensor, mask=None, cond_type="sum", metrics=True) -> dict:
    """inference.py:128-150 for one batch (without the PNG writes); returns the tensors the reference saves and
    the two metric values."""
    out = {}
    imin = float(image.min())
    samples = to_unit_range(samples, imin)
    if mask is not None or cond_type is not None:
        cond = dim_masked(image, mask) if mask is not None else image
        gt, cond = (image, cond) if imin >= 0 else ((image + 1.) / 2., (cond + 1.) / 2.)
        if metrics:
            out["ssim"], out["psnr"] = ssim(samples, gt, 1.0), psnr(samples, gt, 1.0)
        gt = adjust_brightness(gt, 3) if gt.mean() < 0.2 else gt
        cond = adjust_brightness(cond, 3) if cond.mean() < 0.2 and cond_type != "sum" else cond
        out["gt"], out["cond"] = gt, cond
    samples = adjust_brightness(samples, 3) if samples.mean() < 0.2 and samples.shape[0] == 1 else samples
    out["samples"] = samples
    return out
