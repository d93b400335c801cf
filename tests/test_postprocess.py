"""The CPU restatement of the post-processing / metrics step (oracle/postprocess.py) against what can be
pinned here: torchvision (installed) bit-exactly, and -- torchmetrics being absent -- known answers plus an
independent float64 evaluation of the SSIM definition (Wang et al. 2004, the algorithm torchmetrics and
scikit-image both implement: Gaussian window sigma 1.5 truncated to 11 taps, population covariances, border
of 5 cropped) with scipy.ndimage."""
import math

import numpy as np
import pytest
import torch

from oracle import postprocess as P


def test_adjust_brightness_is_torchvision_bit_exact():
    TF = pytest.importorskip("torchvision.transforms.functional")
    g = torch.Generator().manual_seed(0)
    x = torch.rand((2, 3, 17, 19), generator=g)
    for f in (3.0, 0.5, 1.0, 0.0):
        assert torch.equal(P.adjust_brightness(x, f), TF.adjust_brightness(x, f))


def test_range_and_dim_follow_the_reference_expressions():
    g = torch.Generator().manual_seed(1)
    s = torch.randn((2, 3, 8, 8), generator=g)
    assert torch.equal(P.to_unit_range(s, 0.0), s.clip(0, 1))
    assert torch.equal(P.to_unit_range(s, -0.3), (s + 1.) / 2.)
    img = torch.rand((2, 3, 8, 8), generator=g)
    mask = (torch.rand((2, 1, 8, 8), generator=g) > 0.5).float()
    d = P.dim_masked(img, mask)
    assert torch.equal(d[mask.expand_as(d) == 1], img[mask.expand_as(d) == 1])
    assert torch.allclose(d[mask.expand_as(d) == 0], img[mask.expand_as(d) == 0] * 0.7)


def test_psnr_known_answers():
    a = torch.full((2, 3, 16, 16), 0.25)
    assert float(P.psnr(a + 0.1, a)) == pytest.approx(20.0, abs=1e-4)          # mse 0.01
    assert float(P.psnr(a + 0.5, a, data_range=2.0)) == pytest.approx(10 * math.log10(4 / 0.25), abs=1e-4)
    assert math.isinf(float(P.psnr(a, a)))


def test_psnr_and_ssim_reproduce_torchmetrics_published_examples():
    """The only torchmetrics outputs available without the package: the known-answer examples printed in the
    docstrings of torchmetrics.functional.peak_signal_noise_ratio (`tensor(2.5527)`) and
    structural_similarity_index_measure (`preds = torch.rand([3, 3, 256, 256]); target = preds * 0.75` ->
    `tensor(0.9219)`; data_range=None = the larger of the two value ranges).  The SSIM of that pair is a
    property of the uniform distribution, not of the seed (spread 1e-6 over seeds), so the printed four digits
    pin the restatement whatever generator state the docs were built with."""
    pred = torch.tensor([[0.0, 1.0], [2.0, 3.0]])
    target = torch.tensor([[3.0, 2.0], [1.0, 0.0]])
    assert round(float(P.psnr(pred, target, data_range=3.0)), 4) == 2.5527      # data_range=None: target.max - target.min
    for seed in (42, 0, 7):
        torch.manual_seed(seed)
        preds = torch.rand([3, 3, 256, 256])
        tgt = preds * 0.75
        dr = max(float(preds.max() - preds.min()), float(tgt.max() - tgt.min()))
        assert round(float(P.ssim(preds, tgt, data_range=dr)), 4) == 0.9219


def test_psnr_matches_opencv():
    """cv2.PSNR(a, b, R): an independent published implementation (OpenCV core) of the same definition."""
    cv2 = pytest.importorskip("cv2")
    g = torch.Generator().manual_seed(11)
    a = torch.rand((3, 48, 40), generator=g)
    b = (a + 0.05 * torch.randn((3, 48, 40), generator=g)).clamp(0, 1)
    for R in (1.0, 2.0):
        assert float(P.psnr(a, b, data_range=R)) == pytest.approx(cv2.PSNR(a.numpy(), b.numpy(), R), abs=1e-4)


def _ssim_scipy(x, y, data_range=1.0):
    ndi = pytest.importorskip("scipy.ndimage")
    w = np.exp(-0.5 * (np.arange(-5, 6) / 1.5) ** 2)
    w /= w.sum()

    def filt(a):
        a = ndi.correlate1d(a, w, axis=-1, mode="reflect")
        return ndi.correlate1d(a, w, axis=-2, mode="reflect")
    x, y = x.double().numpy(), y.double().numpy()
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    ux, uy = filt(x), filt(y)
    vx, vy, vxy = filt(x * x) - ux * ux, filt(y * y) - uy * uy, filt(x * y) - ux * uy
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    s = s[..., 5:-5, 5:-5]
    return s.reshape(s.shape[0], -1).mean(-1)


@pytest.mark.parametrize("shape", [(2, 3, 40, 40), (1, 1, 11, 23), (3, 13, 32, 27)])
def test_ssim_matches_an_independent_evaluation_of_the_definition(shape):
    g = torch.Generator().manual_seed(sum(shape))
    a = torch.rand(shape, generator=g)
    b = (a + 0.1 * torch.randn(shape, generator=g)).clamp(0, 1)
    got = P.ssim(a, b, per_image=True)
    want = _ssim_scipy(a, b)
    assert np.abs(got.numpy() - want).max() <= 2e-5
    assert float(P.ssim(a, b)) == pytest.approx(float(want.mean()), abs=2e-5)
    assert float(P.ssim(a, a)) == pytest.approx(1.0, abs=1e-6)
    assert float(P.ssim(a, b)) == pytest.approx(float(P.ssim(b, a)), abs=1e-6)


def test_postprocess_branches():
    g = torch.Generator().manual_seed(3)
    img = torch.rand((1, 3, 24, 24), generator=g) * 0.1               # dark image, range [0, 1]
    mask = (torch.rand((1, 1, 24, 24), generator=g) > 0.5).float()
    s = torch.randn((1, 3, 24, 24), generator=g) * 0.05
    o = P.postprocess(s, img, mask, "sum")
    assert torch.equal(o["gt"], P.adjust_brightness(img, 3))          # gt.mean() < 0.2
    assert torch.equal(o["cond"], P.dim_masked(img, mask))            # 'sum': never brightened
    assert torch.equal(o["samples"], P.adjust_brightness(s.clip(0, 1), 3))
    img2 = img * 2 - 1                                                # range [-1, 1] data
    o2 = P.postprocess(s, img2, mask, "concat")
    assert torch.equal(o2["samples"], P.adjust_brightness((s + 1.) / 2., 3)) or torch.equal(o2["samples"], (s + 1.) / 2.)
    assert torch.equal(o2["gt"], P.adjust_brightness((img2 + 1.) / 2., 3))
