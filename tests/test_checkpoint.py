"""Reference checkpoint files load into the drop-in modules.  CPU only."""
import os
import sys

import pytest
import torch
from torch.optim.swa_utils import AveragedModel

from conftest import REFERENCE

from eo_diffusion_b200 import EODiffusion, UNetModel
from eo_diffusion_b200.checkpoint import SCHEDULE_BUFFERS, extract_state_dict, load_checkpoint, make_checkpoint

TINY = dict(image_size=16, in_channels=3, model_channels=32, out_channels=3, num_res_blocks=1,
            attention_resolutions=[2], channel_mult=[1, 2], num_heads=2)


def _ema(model, decay=0.9):
    # reference script_utils/utils.py:56-67 (the file itself imports matplotlib, absent here): the same
    # AveragedModel subclass arguments
    def ema_avg(avg, p, n):
        return decay * avg + (1 - decay) * p
    return AveragedModel(model, "cpu", ema_avg, use_buffers=True)


def _perturb(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in module.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=g))


def _reference_checkpoint(path):
    """A checkpoint written the way reference train.py:137-138 writes it, from the reference's own
    classes when the checkout is present, else from the drop-in modules (identical state-dict layout,
    tests/test_host.py::test_state_dict_identical_to_reference)."""
    if os.path.isdir(REFERENCE):
        sys.path.insert(0, REFERENCE)
        try:
            from backbones.unet_openai import UNetModel as U
            from diffusion.model import EODiffusion as D
        finally:
            sys.path.remove(REFERENCE)
    else:
        U, D = UNetModel, EODiffusion
    torch.manual_seed(7)
    model = D(U(**TINY), 16, 3, timesteps=20, cond_type="sum")
    _perturb(model, 1)                       # the zero-initialised convs become non-trivial
    ema = _ema(model)
    _perturb(model, 2)
    ema.update_parameters(model)             # EMA weights now differ from the plain ones
    ckpt = {"model": model.state_dict(), "model_ema": ema.state_dict()}
    torch.save(ckpt, path)
    return model, ema


@pytest.mark.parametrize("which", ["model", "model_ema"])
def test_load_reference_checkpoint(tmp_path, which):
    path = tmp_path / "steps_00000100.pt"
    ref_model, ref_ema = _reference_checkpoint(path)
    want = ref_model.state_dict() if which == "model" else ref_ema.module.state_dict()

    ours = EODiffusion(UNetModel(**TINY), 16, 3, timesteps=20, cond_type="sum")
    res = load_checkpoint(ours, path, which=which)
    assert not res.missing_keys and not res.unexpected_keys
    got = ours.state_dict()
    assert list(got.keys()) == list(want.keys())
    for k in want:
        assert torch.equal(got[k], want[k]), k
    for b in SCHEDULE_BUFFERS:
        assert b in got
    assert "model.conv_out.weight" in got and "model.nout.weight" in got     # the dead duplicate head

    # a bare UNetModel takes the same file
    unet = UNetModel(**TINY)
    load_checkpoint(unet, path, which=which)
    for k, v in unet.state_dict().items():
        assert torch.equal(v, want["model." + k]), k


def test_strict_mismatch_raises_like_load_state_dict(tmp_path):
    path = tmp_path / "c.pt"
    _reference_checkpoint(path)
    other = EODiffusion(UNetModel(**dict(TINY, model_channels=64)), 16, 3, timesteps=20)
    with pytest.raises(RuntimeError, match="size mismatch"):
        load_checkpoint(other, path)
    ckpt = torch.load(path, weights_only=True)
    del ckpt["model"]["model.out.2.bias"]
    ours = EODiffusion(UNetModel(**TINY), 16, 3, timesteps=20)
    with pytest.raises(RuntimeError, match="Missing key"):
        load_checkpoint(ours, ckpt)
    assert load_checkpoint(ours, ckpt, strict=False).missing_keys == ["model.out.2.bias"]
    with pytest.raises(KeyError):
        load_checkpoint(ours, {"model_ema": ckpt["model_ema"]}, which="model")
    with pytest.raises(KeyError, match="prefix"):
        extract_state_dict({"model_ema": {"n_averaged": torch.tensor(1), "betas": torch.zeros(2)}}, "model_ema")


def test_make_checkpoint_round_trip():
    torch.manual_seed(3)
    a = EODiffusion(UNetModel(**TINY), 16, 3, timesteps=20)
    _perturb(a, 5)
    ema = _ema(a)
    ck = make_checkpoint(a, ema)
    assert set(ck) == {"model", "model_ema"}
    assert list(ck["model_ema"].keys())[0] == "n_averaged"
    b = EODiffusion(UNetModel(**TINY), 16, 3, timesteps=20)
    load_checkpoint(b, ck, which="model_ema")
    for k, v in a.state_dict().items():
        assert torch.equal(b.state_dict()[k], v), k
    ck2 = make_checkpoint(a)                 # no EMA object: plain weights under the EMA key layout
    assert [k for k in ck2["model_ema"] if k != "n_averaged"] == ["module." + k for k in ck2["model"]]
