"""Drop-in `UNetModel` for EO_Diffusion's sampling path, executed by libeo_b200 (sm_100a).

Mirrors the constructor, attributes, state-dict layout and `forward(x, timesteps, cond, y)`
of the reference `backbones/unet_openai.py::UNetModel` (:522-780).  The module tree below
only OWNS parameters (same names, shapes, creation order and default initialisation as the
reference, including the zero-initialised convs, :62-68, and the dead duplicate head
`nout/act/conv_out`, :744, so that seeds, `state_dict()` and `load_state_dict()` are
interchangeable).  The arithmetic of `forward` happens in the CUDA library: parameters are
handed to the engine by their state-dict key, repacked there, and one call runs the whole
network.  Blocks have no PyTorch forward of their own and there is no CPU path.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import math
import operator
import os
import weakref

import torch
import torch.nn as nn

from . import _lib

__all__ = ["UNetModel", "UNet", "UNetBig", "UNetSmall", "ResBlock", "AttentionBlock", "Downsample", "Upsample",
           "TimestepEmbedSequential", "GroupNorm32", "timestep_embedding"]


class GroupNorm32(nn.GroupNorm):
    """Parameter holder for GroupNorm(32, C) evaluated in fp32 (reference :11-13); the
    normalisation itself runs inside the engine's GroupNorm kernels."""


def _zeroed(module: nn.Module) -> nn.Module:
    # reference zero_module (:62-68)
    with torch.no_grad():
        for p in module.parameters():
            p.zero_()
    return module


def timestep_embedding(timesteps, dim, max_period=10000):
    """Host-side restatement of the sinusoidal embedding (reference :81-99), exposed because
    callers of the reference import it; the engine evaluates the same formula on device."""
    half = dim // 2
    freqs = _embedding_freqs(dim, max_period).to(device=timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _embedding_freqs(dim, max_period=10000):
    half = dim // 2
    return torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half)


class _EngineBlock(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - deliberate
        raise NotImplementedError(
            f"{type(self).__name__} only holds parameters; it is executed inside "
            "UNetModel.forward by the libeo_b200 engine")


class TimestepEmbedSequential(nn.Sequential):
    """Container with the reference's child indexing (:195-208)."""

    def forward(self, *a, **k):  # pragma: no cover - deliberate
        return _EngineBlock.forward(self)


class _Resample(nn.Module):
    """Parameter-free Upsample(use_conv=False) / Downsample(use_conv=False) holder (reference :211-271):
    keeps the `h_upd` / `x_upd` module names of an up/down ResBlock."""

    def __init__(self, channels, up):
        super().__init__()
        self.channels = channels
        self.out_channels = channels
        self.use_conv = False
        self.up = up


class ResBlock(_EngineBlock):
    """Parameters of reference ResBlock (:274-385): additive or scale-shift (FiLM) embedding, optionally an
    up- or down-sampling block (`resblock_updown`)."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_scale_shift_norm=False,
                 up=False, down=False):
        super().__init__()
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.use_scale_shift_norm = use_scale_shift_norm
        self.updown = up or down
        oc = self.out_channels
        self.in_layers = nn.Sequential(GroupNorm32(32, channels), nn.SiLU(),
                                       nn.Conv2d(channels, oc, 3, padding=1))
        if up or down:
            self.h_upd = _Resample(channels, up)
            self.x_upd = _Resample(channels, up)
        else:
            self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, 2 * oc if use_scale_shift_norm else oc))
        self.out_layers = nn.Sequential(GroupNorm32(32, oc), nn.SiLU(), nn.Dropout(p=dropout),
                                        _zeroed(nn.Conv2d(oc, oc, 3, padding=1)))
        self.skip_connection = nn.Identity() if oc == channels else nn.Conv2d(channels, oc, 1)


class AttentionBlock(_EngineBlock):
    """Parameters of reference AttentionBlock (:388-433)."""

    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_new_attention_order=False):
        super().__init__()
        self.channels = channels
        if num_head_channels == -1:
            self.num_heads = num_heads
        else:
            assert channels % num_head_channels == 0, \
                f"q,k,v channels {channels} is not divisible by num_head_channels {num_head_channels}"
            self.num_heads = channels // num_head_channels
        self.use_new_attention_order = use_new_attention_order
        self.norm = GroupNorm32(32, channels)
        self.qkv = nn.Conv1d(channels, channels * 3, 1)
        self.proj_out = _zeroed(nn.Conv1d(channels, channels, 1))


class Downsample(_EngineBlock):
    """Parameters of reference Downsample with conv_resample (:245-271)."""

    def __init__(self, channels, out_channels=None):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.op = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=1)


class Upsample(_EngineBlock):
    """Parameters of reference Upsample with conv_resample (:211-242)."""

    def __init__(self, channels, out_channels=None):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=1)


_MODES = {"fp32": _lib.EO_MODE_FP32, "bf16": _lib.EO_MODE_BF16}
_GET_VERSION = operator.attrgetter("_version")


def _destroy_handle(handle):
    if handle:
        try:
            _lib.lib().eo_unet_destroy(C.c_void_p(handle))
        except Exception:  # interpreter shutdown
            pass


class UNetModel(nn.Module):
    """Same constructor as the reference (:553-575).  `dims != 2` and `conv_resample=False`, which no
    configuration of the reference sets, raise NotImplementedError rather than silently diverging.

    Additive API: `compute_mode` ("bf16": tcgen05 tensor-core kernels, the default, or
    "fp32": fp32 CUDA-core kernels for parity checks), selectable with `set_compute_mode`
    or the EO_B200_MODE environment variable."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks,
                 attention_resolutions, time_emb_factor=4, dropout=0, channel_mult=(1, 2, 4, 8),
                 conv_resample=True, dims=2, num_classes=None, use_checkpoint=False,
                 use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False):
        super().__init__()
        if dims != 2 or not conv_resample:
            raise NotImplementedError(
                "eo_diffusion_b200.UNetModel implements the reference's 2-D configurations "
                "(dims=2, conv_resample=True)")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.time_emb_factor = time_emb_factor
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float16 if use_fp16 else torch.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads_upsample
        self.use_new_attention_order = use_new_attention_order
        self.use_scale_shift_norm = use_scale_shift_norm
        self.resblock_updown = resblock_updown
        ssn = use_scale_shift_norm

        ted = model_channels * time_emb_factor
        self.time_embed = nn.Sequential(nn.Linear(model_channels, ted), nn.SiLU(), nn.Linear(ted, ted))
        if num_classes is not None:
            self.label_emb = nn.Embedding(num_classes, ted)

        def attn(ch, heads):
            return AttentionBlock(ch, num_heads=heads, num_head_channels=num_head_channels,
                                  use_new_attention_order=use_new_attention_order)

        ch = stem_ch = int(channel_mult[0] * model_channels)
        self.input_blocks = nn.ModuleList(
            [TimestepEmbedSequential(nn.Conv2d(in_channels, ch, 3, padding=1))])
        self._feature_size = ch
        skip_chans = [ch]
        ds = 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, ted, dropout, out_channels=int(mult * model_channels), use_scale_shift_norm=ssn)]
                ch = int(mult * model_channels)
                if ds in attention_resolutions:
                    layers.append(attn(ch, num_heads))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                self._feature_size += ch
                skip_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(
                    ResBlock(ch, ted, dropout, out_channels=ch, use_scale_shift_norm=ssn, down=True)
                    if resblock_updown else Downsample(ch, out_channels=ch)))
                skip_chans.append(ch)
                ds *= 2
                self._feature_size += ch

        self.middle_block = TimestepEmbedSequential(
            ResBlock(ch, ted, dropout, use_scale_shift_norm=ssn), attn(ch, num_heads),
            ResBlock(ch, ted, dropout, use_scale_shift_norm=ssn))
        self._feature_size += ch

        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = skip_chans.pop()
                layers = [ResBlock(ch + ich, ted, dropout, out_channels=int(model_channels * mult), use_scale_shift_norm=ssn)]
                ch = int(model_channels * mult)
                if ds in attention_resolutions:
                    layers.append(attn(ch, num_heads_upsample))
                if level and i == num_res_blocks:
                    layers.append(ResBlock(ch, ted, dropout, out_channels=ch, use_scale_shift_norm=ssn, up=True)
                                  if resblock_updown else Upsample(ch, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
                self._feature_size += ch

        self.out = nn.Sequential(GroupNorm32(32, ch), nn.SiLU(),
                                 _zeroed(nn.Conv2d(stem_ch, out_channels, 3, padding=1)))
        # dead duplicate head of the reference (:744): kept for state-dict and RNG parity
        self.nout = GroupNorm32(32, ch)
        self.act = nn.SiLU()
        self.conv_out = _zeroed(nn.Conv2d(stem_ch, out_channels, 3, padding=1))

        # ---- engine state (not part of the state dict)
        self._compute_mode = os.environ.get("EO_B200_MODE", "bf16")
        if self._compute_mode not in _MODES:
            raise ValueError(f"EO_B200_MODE must be one of {sorted(_MODES)}")
        self._handle = None
        self._finalizer = None
        self._plan_key = None
        self._staged = []          # fp32 copies handed to the engine (kept alive until finalize)
        self._params_flat = None   # cached list(self.parameters()) for the per-forward staleness check
        self._tt_n = 0             # timestep-table size requested by an open `time_tables` block
        self._tt_installed = 0     # size of the table the engine currently holds
        self._tt_on = False        # ... and whether forwards are using it

    # ------------------------------------------------------------------ engine plumbing
    def __getstate__(self):
        """copy.deepcopy (torch.optim.swa_utils.AveragedModel, the reference's EMA wrapper: train.py:66) and pickle
        must not share the C handle: a copy starts without engine state and plans its own on first use."""
        state = self.__dict__.copy()
        state["_handle"] = None
        state["_finalizer"] = None
        state["_plan_key"] = None
        state["_staged"] = []
        state["_params_flat"] = None
        state["_tt_n"] = 0
        state["_tt_installed"] = 0
        state["_tt_on"] = False
        return state

    @property
    def compute_mode(self) -> str:
        return self._compute_mode

    def set_compute_mode(self, mode: str) -> "UNetModel":
        if mode not in _MODES:
            raise ValueError(f"compute mode must be one of {sorted(_MODES)}, got {mode!r}")
        self._compute_mode = mode
        return self

    def _cfg_struct(self) -> _lib.EoUnetCfg:
        c = _lib.EoUnetCfg()
        c.in_channels = self.in_channels
        c.model_channels = self.model_channels
        c.out_channels = self.out_channels
        c.num_res_blocks = self.num_res_blocks
        ar = [int(a) for a in self.attention_resolutions]
        cm = [int(m) for m in self.channel_mult]
        if any(float(m) != int(m) for m in self.channel_mult):
            raise NotImplementedError("non-integer channel_mult")
        if len(ar) > 8 or len(cm) > 8:
            raise ValueError("at most 8 attention resolutions / channel multipliers")
        c.n_attention_resolutions = len(ar)
        for i, a in enumerate(ar):
            c.attention_resolutions[i] = a
        c.n_channel_mult = len(cm)
        for i, m in enumerate(cm):
            c.channel_mult[i] = m
        c.time_emb_factor = self.time_emb_factor
        c.num_classes = 0 if self.num_classes is None else int(self.num_classes)
        c.num_heads = self.num_heads
        c.num_head_channels = self.num_head_channels
        c.num_heads_upsample = self.num_heads_upsample
        c.use_new_attention_order = int(bool(self.use_new_attention_order))
        c.dims, c.conv_resample = 2, 1
        c.use_scale_shift_norm = int(bool(self.use_scale_shift_norm))
        c.resblock_updown = int(bool(self.resblock_updown))
        return c

    def _ensure_handle(self):
        if self._handle is None:
            L = _lib.lib()
            h = C.c_void_p()
            cfg = self._cfg_struct()
            _lib.check(L.eo_unet_create(C.byref(cfg), C.byref(h)), "eo_unet_create")
            self._handle = h.value
            self._finalizer = weakref.finalize(self, _destroy_handle, self._handle)
        return C.c_void_p(self._handle)

    def _apply(self, fn, *a, **k):
        self._params_flat = None        # .to() / .half() / .cuda(): parameters may be re-created
        return super()._apply(fn, *a, **k)

    def invalidate_engine(self) -> "UNetModel":
        """Force a re-plan at the next forward (only needed after replacing a Parameter OBJECT of a sub-module by
        hand; in-place updates, load_state_dict and .to() are detected)."""
        self._params_flat = None
        self._plan_key = None
        return self

    def _weights_version(self, device):
        # in-place updates bump _version; re-assignment of .data / .to() change data_ptr.  The walk over the module
        # tree (0.6 ms for the 343 tensors of the benchmark UNet) is done once; per forward only the two C properties
        # of each cached tensor are read (~60 us)
        flat = getattr(self, "_params_flat", None)
        if flat is None:
            flat = self._params_flat = list(self.parameters())
        return tuple(map(torch.Tensor.data_ptr, flat)), tuple(map(_GET_VERSION, flat)), str(device)

    def _plan(self, device, batch, H, W):
        """(Re)build the engine's packed weights and launch plan when parameters, device,
        compute mode or geometry changed since the last call."""
        key = (self._compute_mode, H, W, self._weights_version(device))
        if self._plan_key is not None and self._plan_key[0] == key and batch <= self._plan_key[1]:
            return
        L = _lib.lib()
        h = self._ensure_handle()
        sd = self.state_dict(keep_vars=True)
        n = L.eo_unet_num_weights(h)
        staged = []
        for i in range(n):
            name = L.eo_unet_weight_name(h, i).decode()
            if name == "time_embed.freqs":
                t = _embedding_freqs(self.model_channels)
            else:
                t = sd[name].detach()
            t = t.to(device=device, dtype=torch.float32).contiguous()
            staged.append(t)
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
            _lib.check(L.eo_unet_set_weight(h, name.encode(), _lib.ptr(t), shape, t.dim()),
                       f"eo_unet_set_weight({name})")
        max_batch = batch if self._plan_key is None or self._plan_key[0] != key else max(batch, self._plan_key[1])
        _lib.check(L.eo_unet_finalize(h, _MODES[self._compute_mode], max_batch, H, W, _lib.stream_ptr()),
                   "eo_unet_finalize")
        self._staged = []   # the engine has packed / copied everything it needs
        self._plan_key = (key, max_batch)
        self._tt_installed = 0   # timestep tables die with the plan
        self._tt_on = False

    @contextlib.contextmanager
    def time_tables(self, n_timesteps: int):
        """Additive API used by the samplers: inside the block, forwards without class labels take their
        timestep-embedding rows from a table of the timestep values 0 .. n_timesteps-1 computed once
        (`eo_unet_build_time_tables`; reference work hoisted: unet_openai.py:763, :374-376) -- bit-identical rows,
        one launch instead of four per step.  Timesteps outside the table give NaN, so only a caller that knows its
        schedule (EODiffusion.sampling, DDIMSampler.ddim_sampling) opens this block."""
        prev = self._tt_n
        self._tt_n = max(int(n_timesteps), prev)
        try:
            yield self
        finally:
            self._tt_n = prev
            if prev == 0 and self._tt_on and self._handle is not None:
                _lib.check(_lib.lib().eo_unet_clear_time_tables(C.c_void_p(self._handle)), "eo_unet_clear_time_tables")
                self._tt_on = False

    def launches_per_forward(self) -> int:
        return int(_lib.lib().eo_unet_launches_per_forward(self._ensure_handle()))

    def engine_device_bytes(self) -> int:
        return int(_lib.lib().eo_unet_device_bytes(self._ensure_handle()))

    def read_activation(self, name: str, batch: int) -> torch.Tensor:
        """Debug aid for the parity tests (needs EO_DEBUG_KEEP=1 in the environment when the
        plan is built): the output of reference module `name` in the last forward, NCHW fp32."""
        L = _lib.lib()
        dev = next(self.parameters()).device
        cap = batch * 2048 * self._plan_key[0][1] * self._plan_key[0][2]
        buf = torch.empty(cap, dtype=torch.float32, device=dev)
        n = _lib.check(L.eo_unet_read_activation(self._ensure_handle(), name.encode(), _lib.ptr(buf),
                                                 cap, batch, _lib.stream_ptr()), "read_activation")
        return buf[:n]

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def forward(self, x, timesteps, cond=None, y=None):
        """eps = UNet(cat(x, cond), timesteps[, y])  (reference :746-780).
        x [N,C,H,W]; timesteps [N] int; cond [N,Cc,H,W] or None; y [N] int labels iff
        class-conditional.  Returns [N, out_channels, H, W] in x's dtype."""
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        _lib.require_cuda_tensor(x, "x")
        dev = x.device
        if cond is not None:
            cond = cond.to(dev)
        B, Cx, H, W = x.shape
        Cc = 0 if cond is None else cond.shape[1]
        if Cx + Cc != self.in_channels:
            raise RuntimeError(f"expected {self.in_channels} input channels, got x with {Cx}"
                               + (f" + cond with {Cc}" if cond is not None else ""))
        if y is not None:
            assert y.shape == (B,), (y.shape, x.shape)
        with torch.cuda.device(dev):
            self._plan(dev, B, H, W)
            if self._tt_n and y is None and (not self._tt_on or self._tt_installed < self._tt_n):
                # (a table of at least this size kept by the handle from an earlier loop is only switched back on)
                _lib.check(_lib.lib().eo_unet_build_time_tables(C.c_void_p(self._handle), self._tt_n, _lib.stream_ptr()),
                           "eo_unet_build_time_tables")
                self._tt_installed = max(self._tt_installed, self._tt_n)
                self._tt_on = True
            xf = x.detach().to(torch.float32).contiguous()
            cf = None if cond is None else cond.detach().to(torch.float32).contiguous()
            ts = timesteps.to(device=dev, dtype=torch.int64).contiguous()
            yy = None if y is None else y.to(device=dev, dtype=torch.int64).contiguous()
            out = torch.empty((B, self.out_channels, H, W), dtype=torch.float32, device=dev)
            L = _lib.lib()
            _lib.check(L.eo_unet_forward(C.c_void_p(self._handle), _lib.ptr(xf), Cx, _lib.ptr(cf), Cc,
                                         _lib.ptr(ts), _lib.ptr(yy), _lib.ptr(out), B, _lib.stream_ptr()),
                       "eo_unet_forward")
        return out if x.dtype == torch.float32 else out.to(x.dtype)


def _factory(image_size, in_channels, out_channels, base_width, num_classes, num_res_blocks, mults, head_channels=64,
             time_emb_factor=4):
    # the reference's UNetBig / UNet / UNetSmall (unet_openai.py:783-922): same tables, same flags
    if image_size not in mults:
        raise ValueError(f"unsupported image size: {image_size}")
    attention_resolutions = "28,14,7" if image_size == 28 else "32,16,8"
    attention_ds = tuple(image_size // int(res) for res in attention_resolutions.split(","))
    return UNetModel(image_size=image_size, in_channels=in_channels, model_channels=base_width,
                     out_channels=out_channels, num_res_blocks=num_res_blocks, attention_resolutions=attention_ds,
                     time_emb_factor=time_emb_factor, dropout=0.1, channel_mult=mults[image_size], num_classes=num_classes, use_checkpoint=False,
                     use_fp16=False, num_heads=4, num_head_channels=head_channels, num_heads_upsample=-1,
                     use_scale_shift_norm=True, resblock_updown=True, use_new_attention_order=True)

_MULTS = {128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4), 32: (1, 2, 2, 2), 28: (1, 2, 2, 2)}


def UNetBig(image_size, in_channels=3, out_channels=3, base_width=192, num_classes=None):
    """Reference factory (unet_openai.py:783-828): 3 res blocks per level."""
    return _factory(image_size, in_channels, out_channels, base_width, num_classes, 3, _MULTS)


def UNet(image_size, in_channels=3, out_channels=3, base_width=64, num_classes=None):
    """Reference factory (unet_openai.py:830-875)."""
    return _factory(image_size, in_channels, out_channels, base_width, num_classes, 3, _MULTS)


def UNetSmall(image_size, in_channels=3, out_channels=3, base_width=32, num_classes=None):
    """Reference factory (unet_openai.py:877-922): 2 res blocks per level, 32-channel heads, time_emb_factor 2."""
    return _factory(image_size, in_channels, out_channels, base_width, num_classes, 2, _MULTS, head_channels=32,
                    time_emb_factor=2)
