"""The tcgen05 kernels in isolation (C-ABI self-test entry points) against plain PyTorch
fp32 references evaluated on the same bf16-rounded operands.  Needs a B200."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2
from eo_diffusion_b200 import _lib

pytestmark = pytest.mark.gpu


def _conv_case(dev, B, H, W, Cin, Cout, k, with_res, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((B, Cin, H, W), generator=g).to(dev)
    w = (torch.randn((Cout, Cin, k, k), generator=g) / math.sqrt(Cin * k * k)).to(dev)
    b = torch.randn((Cout,), generator=g).to(dev)
    res = torch.randn((B, Cout, H, W), generator=g).to(dev) if with_res else None
    x_bf = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)         # NHWC
    r_bf = res.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16) if with_res else None
    y = torch.empty((B, H, W, Cout), dtype=torch.bfloat16, device=dev)
    _lib.check(_lib.lib().eo_test_conv_tc(_lib.ptr(x_bf), _lib.ptr(w), _lib.ptr(b), _lib.ptr(r_bf), _lib.ptr(y),
                                          B, H, W, Cin, Cout, k, _lib.stream_ptr()), "eo_test_conv_tc")
    torch.cuda.synchronize()
    want = F.conv2d(x_bf.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), b, padding=k // 2)
    if with_res:
        want = want + r_bf.float().permute(0, 3, 1, 2)
    return y.float().permute(0, 3, 1, 2), want


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,res", [
    (1, 16, 16, 64, 128, 1, False),     # plain GEMM, one K block
    (2, 16, 16, 128, 128, 3, False),    # 3x3, zero padding through TMA OOB fill
    (2, 8, 8, 192, 192, 3, True),       # 8x8 map: two images per 128-pixel tile, Cout not /128
    (3, 8, 8, 64, 64, 3, False),        # odd batch: partially filled tile
    (1, 32, 32, 256, 256, 3, True),     # BN = 256 variant
    (1, 64, 64, 128, 384, 3, False),
    # tile grids that are not powers of two (tile decode by multiply-high): 3 x 3, 6 x 5 and 3 x 7 tiles of 16 x 8 pixels,
    # an odd number of tiles (all-masked partner), several images
    (2, 48, 24, 64, 128, 3, False),
    (1, 96, 40, 128, 128, 3, True),
    (3, 48, 56, 64, 192, 3, False),
    (5, 48, 32, 128, 256, 1, True),
    (3, 24, 24, 128, 256, 1, True),     # 8 x 8 pixels of two images per tile
    (9, 12, 12, 64, 128, 3, False),     # 4 x 4 pixels of eight images per tile, plain 3x3 taps
    (2, 32, 32, 1024, 512, 3, False),   # deepest K of the BASELINE UNet (144 K blocks)
])
def test_conv_tc(cuda_dev, B, H, W, Cin, Cout, k, res):
    got, want = _conv_case(cuda_dev, B, H, W, Cin, Cout, k, res, seed=B * 1000 + Cin + Cout + k)
    # bf16 output rounding (2^-9 relative) dominates
    assert rel_l2(got, want) <= 4e-3


def _attn_ref(qkv, heads, ch):
    # QKVAttentionLegacy (unet_openai.py:465-481) on [B, T, heads*3*ch] -> [B, T, heads*ch]
    B, T, _ = qkv.shape
    q, k, v = qkv.float().reshape(B, T, heads, 3, ch).permute(3, 0, 2, 1, 4)   # [B, heads, T, ch]
    if ch <= 48:
        # heads of <= 48 channels: the engine folds ch^-1/2 * log2(e) into the q projection, so q reaches the kernel as
        # bf16(q * that factor); the self-test entry rounds the same way, and so does this reference
        f = math.log2(math.e) / math.sqrt(ch)
        q = (q * f).to(torch.bfloat16).float() / f
    s = 1 / math.sqrt(math.sqrt(ch))
    w = torch.softmax(torch.einsum("bhtc,bhsc->bhts", q * s, k * s), dim=-1)
    a = torch.einsum("bhts,bhsc->bhtc", w, v)
    return a.permute(0, 2, 1, 3).reshape(B, T, heads * ch)


@pytest.mark.parametrize("B,T,heads,ch", [
    (1, 128, 1, 64), (2, 256, 4, 32), (1, 64, 8, 64), (2, 1024, 8, 48), (1, 4096, 2, 48), (1, 320, 2, 16),
    # ragged sequence lengths: a last half-block of < 32 and < 64 keys, a CTA with one, two and three query tiles
    (1, 100, 2, 48), (2, 576, 2, 64), (1, 40, 1, 64), (1, 1000, 1, 48), (1, 1336, 1, 64),
])
@pytest.mark.parametrize("gain", [1.5, 6.0])      # 6.0: logits of +-100, the reference maximum keeps moving
def test_attention_tc(cuda_dev, B, T, heads, ch, gain):
    g = torch.Generator().manual_seed(T + heads + ch)
    qkv = (torch.randn((B, T, heads * 3 * ch), generator=g) * gain).to(cuda_dev).to(torch.bfloat16)
    out = torch.empty((B, T, heads * ch), dtype=torch.bfloat16, device=cuda_dev)
    _lib.check(_lib.lib().eo_test_attention_tc(_lib.ptr(qkv), _lib.ptr(out), B, T, heads, ch,
                                               _lib.stream_ptr()), "eo_test_attention_tc")
    torch.cuda.synchronize()
    want = _attn_ref(qkv, heads, ch)
    assert rel_l2(out.float(), want) <= 1e-2
