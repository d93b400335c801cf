"""Reference checkpoint files -> the drop-in modules.

The reference writes `{"model": EODiffusion.state_dict(), "model_ema": ExponentialMovingAverage.state_dict()}`
(`train.py:137-138`) and reads it back with `model.load_state_dict(ckpt["model"])`
(`inference.py:79-87`).  The EMA wrapper is `torch.optim.swa_utils.AveragedModel`
(`script_utils/utils.py:56-67`, `use_buffers=True`), so its keys are the model's prefixed with
`module.` plus the counter `n_averaged`.  Key names and shapes are matched strictly, including the
schedule buffers and the dead duplicate head `model.nout.* / model.conv_out.*`
(`backbones/unet_openai.py:744`).

Nothing here computes: tensors are copied into the module's parameters, and the next
`UNetModel.forward` notices the change (parameter version counters) and repacks the engine's weights.
"""
from __future__ import annotations

import collections
import os
from typing import Mapping, Union

import torch

__all__ = ["load_checkpoint", "extract_state_dict", "make_checkpoint", "SCHEDULE_BUFFERS"]

SCHEDULE_BUFFERS = ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod")
_EMA_PREFIX = "module."
_EMA_COUNTER = "n_averaged"


def extract_state_dict(ckpt: Mapping, which: str = "model") -> "collections.OrderedDict[str, torch.Tensor]":
    """The `EODiffusion` state dict stored under `which` ("model" or "model_ema"), with the
    EMA wrapper's `module.` prefix and `n_averaged` counter removed.  A bare state dict (no
    "model" / "model_ema" entry) is accepted as is."""
    if which not in ("model", "model_ema"):
        raise ValueError(f"which must be 'model' or 'model_ema', got {which!r}")
    if "model" not in ckpt and "model_ema" not in ckpt:
        sd = ckpt                                   # already a state dict
    elif which not in ckpt:
        raise KeyError(f"checkpoint has no {which!r} entry (found {sorted(ckpt.keys())})")
    else:
        sd = ckpt[which]
    out = collections.OrderedDict()
    is_ema = any(k == _EMA_COUNTER or k.startswith(_EMA_PREFIX) for k in sd.keys())
    for k, v in sd.items():
        if is_ema:
            if k == _EMA_COUNTER:
                continue
            if not k.startswith(_EMA_PREFIX):
                raise KeyError(f"unexpected key {k!r} in an EMA state dict (expected the {_EMA_PREFIX!r} prefix)")
            k = k[len(_EMA_PREFIX):]
        out[k] = v
    return out


def load_checkpoint(diffusion: torch.nn.Module, ckpt: Union[str, os.PathLike, Mapping], which: str = "model",
                    strict: bool = True, map_location="cpu"):
    """`inference.py:79-87` for the drop-in `EODiffusion`: load the file (or an already loaded
    dict), pick the plain or the EMA weights, `load_state_dict` them strictly.  Also accepts a
    `UNetModel` as `diffusion` together with an `EODiffusion` checkpoint: the `model.` prefix is
    stripped and the schedule buffers are dropped.  Returns `load_state_dict`'s result."""
    if not isinstance(ckpt, Mapping):
        ckpt = torch.load(os.fspath(ckpt), map_location=map_location, weights_only=True)
    sd = extract_state_dict(ckpt, which)
    own = diffusion.state_dict()
    if not any(k.startswith("model.") for k in own) and any(k.startswith("model.") for k in sd):
        # a bare UNetModel receiving an EODiffusion state dict
        sd = collections.OrderedDict((k[len("model."):], v) for k, v in sd.items() if k.startswith("model."))
    return diffusion.load_state_dict(sd, strict=strict)


def make_checkpoint(diffusion: torch.nn.Module, ema: torch.nn.Module = None) -> dict:
    """The reference's checkpoint dict (`train.py:137-138`) from the drop-in modules; `ema` is a
    `torch.optim.swa_utils.AveragedModel` around an `EODiffusion`, or None to store the plain
    weights under both entries with the EMA key layout."""
    model_sd = collections.OrderedDict((k, v.detach().cpu().clone()) for k, v in diffusion.state_dict().items())
    if ema is not None:
        ema_sd = collections.OrderedDict((k, v.detach().cpu().clone()) for k, v in ema.state_dict().items())
    else:
        ema_sd = collections.OrderedDict([(_EMA_COUNTER, torch.tensor(0, dtype=torch.long))])
        for k, v in model_sd.items():
            ema_sd[_EMA_PREFIX + k] = v.clone()
    return {"model": model_sd, "model_ema": ema_sd}
