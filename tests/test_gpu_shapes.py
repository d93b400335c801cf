"""Shapes and constructor settings off the benchmark's beaten path, against the oracle evaluated here on the same
weights: non-square inputs, feature maps that no tile geometry divides evenly, five levels, three ResBlocks per level,
attention at every level, odd batch sizes, heads of 24 / 96 channels, 13-channel outputs.  Needs a B200."""
import pytest
import torch

from conftest import build_unet, rel_l2
from oracle import oracle as O

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-4
TOL_BF16_NARROW = 1.25e-2      # 64-wide synthetic nets (tests/test_gpu_unet.py); 128-wide ones are held to 1e-2

CASES = {
    # name: (cfg, H, W, batch)
    "nonsquare_64x32": (dict(image_size=64, in_channels=3, model_channels=64, out_channels=3, num_res_blocks=1,
                             attention_resolutions=[2], channel_mult=[1, 2], num_heads=4), 64, 32, 2),
    "80x80_three_levels": (dict(image_size=80, in_channels=3, model_channels=64, out_channels=3, num_res_blocks=1,
                                attention_resolutions=[4], channel_mult=[1, 2, 2], num_heads=2), 80, 80, 1),
    "40x24_odd_maps": (dict(image_size=40, in_channels=4, model_channels=64, out_channels=2, num_res_blocks=2,
                            attention_resolutions=[1, 2], channel_mult=[1, 2], num_heads=2), 40, 24, 3),
    "five_levels_128": (dict(image_size=128, in_channels=3, model_channels=64, out_channels=3, num_res_blocks=1,
                             attention_resolutions=[8, 16], channel_mult=[1, 1, 2, 3, 4], num_heads=4), 128, 128, 1),
    "three_resblocks_heads24": (dict(image_size=32, in_channels=3, model_channels=64, out_channels=13, num_res_blocks=3,
                                     attention_resolutions=[2, 4], channel_mult=[1, 3, 3], num_head_channels=24,
                                     use_new_attention_order=True), 32, 32, 5),
    "wide128_heads96": (dict(image_size=32, in_channels=3, model_channels=128, out_channels=3, num_res_blocks=1,
                             attention_resolutions=[2], channel_mult=[1, 3], num_heads=4), 32, 32, 2),
}


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_shape(cuda_dev, name, mode):
    cfg, H, W, B = CASES[name]
    m = build_unet(cfg, 301, 302)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    gen = torch.Generator().manual_seed(sum(map(ord, name)))
    x = torch.randn((B, cfg["in_channels"], H, W), generator=gen)
    t = torch.randint(0, 1000, (B,), generator=gen)
    want = O.unet_forward(sd, O.full_cfg(**cfg), x, t)
    got = m.to(cuda_dev).set_compute_mode(mode)(x.to(cuda_dev), t.to(cuda_dev))
    err = rel_l2(got, want)
    print(f"[parity] {mode} {name} ({H}x{W}, batch {B}): eps rel L2 {err:.3e}")
    tol = TOL_FP32 if mode == "fp32" else (1e-2 if cfg["model_channels"] >= 128 else TOL_BF16_NARROW)
    assert err <= tol, f"{name} {mode}: {err:.3e} > {tol}"
