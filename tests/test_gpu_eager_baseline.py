"""The practical "kernel to beat" (SURVEY.md 8d): the reference's own op sequence (the oracle restatement, i.e.
plain PyTorch eager: cuDNN convolutions, cuBLAS bmm + softmax with the [T, T] scores in HBM, ATen GroupNorm)
run ON THE SAME B200, against the engine, at the BASELINE architecture and 256 x 256.  A measurement with a
sanity bound, not a parity test: the numbers go to gpurun_out/eager_baseline.json (and the test log)."""
import json
import os

import pytest
import torch

from conftest import ROOT, rel_l2
from eo_diffusion_b200 import UNetModel
from oracle import oracle as O

pytestmark = pytest.mark.gpu

ARCH = dict(in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2,
            attention_resolutions=[4, 8], channel_mult=[1, 2, 3, 4], num_heads=8)


def _time(fn, warm=2, reps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def test_engine_vs_pytorch_eager_on_the_same_gpu(cuda_dev):
    size, B = 256, 4          # eager materialises [B*8, 4096, 4096] fp32 scores: 2.1 GB per attention block at B = 4
    torch.manual_seed(1234)
    m = O.dezero_(UNetModel(image_size=size, **ARCH)).eval()
    cfg = O.full_cfg(image_size=size, **ARCH)
    sd = {k: v.detach().to(cuda_dev) for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    x = torch.randn((B, 3, size, size), generator=g).to(cuda_dev)
    t = torch.full((B,), 500, dtype=torch.long, device=cuda_dev)
    res = {"config": f"UNet base 128 mult [1,2,3,4] attn [4,8], {size}x{size}, batch {B}, one forward", "ms": {}}
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        with torch.no_grad():
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            res["ms"]["eager_fp32"], ref = _time(lambda: O.unet_forward(sd, cfg, x, t))
            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cuda.matmul.allow_tf32 = True
            res["ms"]["eager_tf32"], _ = _time(lambda: O.unet_forward(sd, cfg, x, t))
            with torch.autocast("cuda", dtype=torch.bfloat16):
                res["ms"]["eager_bf16_autocast"], auto = _time(lambda: O.unet_forward(sd, cfg, x, t))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    m = m.to(cuda_dev).set_compute_mode("bf16")
    res["ms"]["engine_bf16"], ours = _time(lambda: m(x, t))
    res["rel_l2_vs_eager_fp32"] = {"engine_bf16": rel_l2(ours, ref), "eager_bf16_autocast": rel_l2(auto.float(), ref)}
    res["speedup_vs_eager_bf16_autocast"] = res["ms"]["eager_bf16_autocast"] / res["ms"]["engine_bf16"]
    res["speedup_vs_eager_fp32"] = res["ms"]["eager_fp32"] / res["ms"]["engine_bf16"]
    print("\n[eager baseline] " + json.dumps(res))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "eager_baseline.json"), "w") as f:
            json.dump(res, f, indent=1)
    assert res["rel_l2_vs_eager_fp32"]["engine_bf16"] <= 1e-2
    assert res["ms"]["engine_bf16"] < res["ms"]["eager_bf16_autocast"]
