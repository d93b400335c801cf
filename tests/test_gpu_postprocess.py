"""libeo_b200's post-processing and metric kernels (through eo_diffusion_b200.postprocess -> C ABI) against
the CPU oracle on the same seeded inputs.  Elementwise passes: bit-exact.  PSNR / SSIM: fp32 tolerance stated
per assert.  Needs a B200."""
import math

import pytest
import torch

from eo_diffusion_b200 import postprocess as G
from oracle import postprocess as P

pytestmark = pytest.mark.gpu


def test_elementwise_passes_bit_exact(cuda_dev):
    g = torch.Generator().manual_seed(0)
    x = torch.randn((3, 3, 33, 47), generator=g) * 1.5
    x[0, 0, 0, 0] = float("nan")
    xd = x.to(cuda_dev)
    assert torch.equal(G.to_unit_range(xd, 0.0).cpu().nan_to_num(7.0), P.to_unit_range(x, 0.0).nan_to_num(7.0))
    assert torch.equal(G.to_unit_range(xd, -1.0).cpu().nan_to_num(7.0), P.to_unit_range(x, -1.0).nan_to_num(7.0))
    for f in (3.0, 0.4):
        assert torch.equal(G.adjust_brightness(xd, f).cpu().nan_to_num(7.0), P.adjust_brightness(x, f).nan_to_num(7.0))
    img = torch.rand((3, 13, 16, 20), generator=g)
    mask = (torch.rand((3, 1, 16, 20), generator=g) > 0.4).float()
    assert torch.equal(G.dim_masked(img.to(cuda_dev), mask.to(cuda_dev)).cpu(), P.dim_masked(img, mask))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.adjust_brightness(x, 3.0)


def test_stats(cuda_dev):
    g = torch.Generator().manual_seed(1)
    x = torch.randn((5, 3, 64, 64), generator=g)
    s = G.tensor_stats(x.to(cuda_dev)).cpu()
    assert float(s[0]) == pytest.approx(float(x.double().mean()), abs=1e-7)
    assert float(s[1]) == float(x.min()) and float(s[2]) == float(x.max())


@pytest.mark.parametrize("shape", [(2, 3, 40, 40), (1, 1, 11, 23), (3, 13, 32, 27), (4, 3, 128, 128), (2, 3, 37, 53)])
def test_psnr_ssim_vs_oracle(cuda_dev, shape):
    g = torch.Generator().manual_seed(sum(shape))
    a = torch.rand(shape, generator=g)
    b = (a + 0.1 * torch.randn(shape, generator=g)).clamp(0, 1)
    ad, bd = a.to(cuda_dev), b.to(cuda_dev)
    assert float(G.peak_signal_noise_ratio(ad, bd, data_range=1.0)) == pytest.approx(float(P.psnr(a, b)), abs=2e-4)
    assert float(G.structural_similarity_index_measure(ad, bd, data_range=1.0)) == pytest.approx(float(P.ssim(a, b)), abs=2e-5)
    per = G.structural_similarity_index_measure(ad, bd, data_range=1.0, reduction="none").cpu()
    assert (per - P.ssim(a, b, per_image=True)).abs().max() <= 2e-5
    assert float(G.structural_similarity_index_measure(ad, ad)) == pytest.approx(1.0, abs=1e-6)


def test_full_size_properties(cuda_dev):
    # BASELINE size (batch 64 of 3 x 256 x 256): known answers instead of a CPU comparison
    g = torch.Generator(device=cuda_dev).manual_seed(2)
    a = torch.rand((64, 3, 256, 256), generator=g, device=cuda_dev) * 0.8
    assert float(G.structural_similarity_index_measure(a, a)) == pytest.approx(1.0, abs=1e-6)
    assert float(G.peak_signal_noise_ratio(a + 0.1, a)) == pytest.approx(20.0, abs=1e-3)
    assert math.isinf(float(G.peak_signal_noise_ratio(a, a)))
    s1 = float(G.structural_similarity_index_measure(a, (a + 0.05).clamp(0, 1)))
    s2 = float(G.structural_similarity_index_measure(a, (a + 0.2).clamp(0, 1)))
    assert 1.0 > s1 > s2 > 0.0
    with pytest.raises(Exception, match="smaller than the 11x11 window"):
        G.structural_similarity_index_measure(a[:, :, :8, :8].contiguous(), a[:, :, :8, :8].contiguous())


@pytest.mark.parametrize("signed,batch,cond_type", [(False, 1, "sum"), (True, 1, "concat"), (False, 4, "sum")])
def test_postprocess_samples_vs_oracle(cuda_dev, signed, batch, cond_type):
    g = torch.Generator().manual_seed(3 + batch)
    img = torch.rand((batch, 3, 48, 48), generator=g) * 0.15            # dark: the brightness branches fire
    if signed:
        img = img * 2 - 1
    mask = (torch.rand((batch, 1, 48, 48), generator=g) > 0.5).float()
    s = torch.randn((batch, 3, 48, 48), generator=g) * 0.1 + 0.05
    want = P.postprocess(s, img, mask, cond_type)
    got = G.postprocess_samples(s.to(cuda_dev), img.to(cuda_dev), mask.to(cuda_dev), cond_type)
    for k in ("samples", "gt", "cond"):
        assert torch.equal(got[k].cpu(), want[k]), k
    assert float(got["ssim"]) == pytest.approx(float(want["ssim"]), abs=2e-5)
    assert float(got["psnr"]) == pytest.approx(float(want["psnr"]), abs=2e-4)


def _tv_grid_u8(x, **kw):
    # torchvision.utils.save_image's own lines (installed torchvision, CPU)
    from torchvision.utils import make_grid
    return make_grid(x, **kw).mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8)


@pytest.mark.parametrize("shape, nrow, padding, pad_value", [
    ((16, 3, 64, 64), 4, 2, 0.0),          # sampling(): nrow = int(sqrt(n))
    ((5, 3, 33, 47), 8, 2, 0.0),           # fewer images than nrow: one ragged row
    ((7, 3, 20, 24), 3, 2, 0.5),           # last row partly empty, grey border
    ((1, 3, 40, 24), 1, 2, 0.0),           # one image: torchvision returns it without a border
    ((6, 1, 16, 16), 2, 1, 1.0),           # single-channel batch repeated to three channels
    ((4, 3, 32, 32), 2, 0, 0.0),           # no border at all
])
def test_grid_u8_is_torchvision_bit_exact(cuda_dev, shape, nrow, padding, pad_value):
    g = torch.Generator().manual_seed(sum(shape) + nrow)
    x = torch.rand(shape, generator=g) * 1.3 - 0.15          # values outside [0, 1] exercise the clamp
    x.view(-1)[:512] = torch.arange(512, dtype=torch.float32) / 510.0      # .5 ties of the quantiser
    got = G.make_grid_u8(x.to(cuda_dev), nrow=nrow, padding=padding, pad_value=pad_value)
    want = _tv_grid_u8(x, nrow=nrow, padding=padding, pad_value=pad_value)
    assert got.dtype == torch.uint8 and tuple(got.shape) == tuple(want.shape)
    assert torch.equal(got.cpu(), want)
    xs = x * 2 - 1
    got = G.make_grid_u8(xs.to(cuda_dev), nrow=nrow, padding=padding, pad_value=pad_value, signed=True)
    assert torch.equal(got.cpu(), _tv_grid_u8((xs + 1.) / 2., nrow=nrow, padding=padding, pad_value=pad_value))


def test_save_image_writes_the_file_torchvision_writes(cuda_dev, tmp_path):
    from PIL import Image
    from torchvision.utils import save_image
    import numpy as np
    g = torch.Generator().manual_seed(5)
    x = torch.rand((9, 3, 32, 32), generator=g)
    save_image(x, str(tmp_path / "tv.png"), nrow=3)
    G.save_image(x.to(cuda_dev), str(tmp_path / "eo.png"), nrow=3)
    assert (tmp_path / "tv.png").read_bytes() == (tmp_path / "eo.png").read_bytes()
    save_image(x[0, 0], str(tmp_path / "tv1.png"))                  # [H, W]
    G.save_image(x[0, 0].to(cuda_dev), str(tmp_path / "eo1.png"))
    assert np.array_equal(np.asarray(Image.open(tmp_path / "tv1.png")), np.asarray(Image.open(tmp_path / "eo1.png")))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.make_grid_u8(x)
