"""Shared test helpers.  `-m "not gpu"` runs here on CPU; `-m gpu` needs a B200."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("EO_REFERENCE", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box)")


def golden(name: str) -> dict:
    with np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def golden_cfg(g: dict) -> dict:
    return json.loads(str(g["cfg"]))


def tt(a) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a))


def build_unet(cfg: dict, init_seed: int = 1234, dezero_seed: int = 4321):
    """The drop-in UNetModel with the golden fixtures' weights: constructed under
    `init_seed` (same parameter creation order as the reference => same values), then
    de-zeroed exactly like oracle/make_golden.py did for the reference model."""
    from eo_diffusion_b200 import UNetModel
    from oracle import oracle as O
    torch.manual_seed(init_seed)
    m = UNetModel(**cfg)
    O.dezero_(m, dezero_seed)
    return m.eval()


def weight_checksum(sd) -> dict:
    import hashlib
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        v = sd[k].detach().cpu().contiguous()
        h.update(k.encode())
        h.update(v.numpy().tobytes())
    return h.hexdigest()


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="session")
def cuda_dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from eo_diffusion_b200 import _lib
    assert _lib.lib().eo_device_check() == 0, _lib.last_error()
    return torch.device("cuda:0")


import contextlib


@contextlib.contextmanager
def replay(randn_queue, randn_like_queue, tmp_cwd=None):
    """Patch torch.randn / torch.randn_like to pop pre-drawn tensors in the reference's draw
    order -- the same harness oracle/make_golden.py drives the reference with.
    A None entry in the randn_like queue returns zeros (a draw the sampler discards)."""
    o_randn, o_like = torch.randn, torch.randn_like
    rq, lq = list(randn_queue), list(randn_like_queue)

    def f_randn(*a, **k):
        t = rq.pop(0)
        dev = k.get("device", None)
        return t.to(dev) if dev is not None else t

    def f_like(x, **k):
        t = lq.pop(0)
        return torch.zeros_like(x) if t is None else t.to(x.device)

    torch.randn, torch.randn_like = f_randn, f_like
    cwd = os.getcwd()
    if tmp_cwd is not None:
        os.makedirs(os.path.join(tmp_cwd, "results", "prova"), exist_ok=True)
        os.chdir(tmp_cwd)
    try:
        yield
    finally:
        torch.randn, torch.randn_like = o_randn, o_like
        os.chdir(cwd)
