#!/usr/bin/env python
"""Where does the bf16 mode's eps error come from, and what would a wider residual stream buy?  CPU emulation: the
oracle's UNet forward with the tensor classes the engine keeps in 16 bits rounded (weights, MMA operands, the first
convolution's output, the residual stream, qkv, the attention output) -- all of them = the engine's arithmetic
(8.1e-3 at the BASELINE architecture, 8.0e-3 .. 8.3e-3 measured on the GPU), then one class at a time exact, and the
residual stream as a bf16 (hi, lo) pair on the residual adds only.  `--fp16`: the same with IEEE half instead of bf16.
TEST INFRASTRUCTURE (uses oracle/).  Results quoted in DESIGN.md section 2.
usage: python tools/bf16_budget.py [golden name] [--fp16]"""
import os, sys, math
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import build_unet, golden, golden_cfg, tt
from oracle import oracle as O

NARROW = torch.float16 if "--fp16" in sys.argv else torch.bfloat16
def bf(x): return x.to(NARROW).to(torch.float32)

class Cfg: pass
E = Cfg()
E.weights = E.operands = E.conv1_out = E.stream = E.qkv = E.attn_out = True
E.stream_hilo = False   # residual adds read the exact stream; everything else reads its bf16 rounding

def W(sd, k): return bf(sd[k]) if E.weights else sd[k]
def opnd(x): return bf(x) if E.operands else x

def res_block(sd, p, x, emb, updown=None, scale_shift=False):
    # x: (exact, rounded) pair
    xe, xr = x
    h = O.group_norm32(xr, sd[p + "in_layers.0.weight"], sd[p + "in_layers.0.bias"])
    h = opnd(F.silu(h))
    h = F.conv2d(h, W(sd, p + "in_layers.2.weight"), sd[p + "in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])[..., None, None]
    h = h + e
    if E.conv1_out: h = bf(h)
    h = O.group_norm32(h, sd[p + "out_layers.0.weight"], sd[p + "out_layers.0.bias"])
    h = opnd(F.silu(h))
    h = F.conv2d(h, W(sd, p + "out_layers.3.weight"), sd[p + "out_layers.3.bias"], padding=1)
    if (p + "skip_connection.weight") in sd:
        r = F.conv2d(xr, W(sd, p + "skip_connection.weight"), sd[p + "skip_connection.bias"])
    else:
        r = xe if E.stream_hilo else xr
    return stream(r + h)

def stream(o):
    if not E.stream: return (o, o)
    hi = bf(o)
    if E.stream_hilo:
        return (hi + bf(o - hi), hi)
    return (hi, hi)

def attention_block(sd, p, x, n_heads, new_order):
    xe, xr = x
    b, c, *sp = xr.shape
    xe = xe.reshape(b, c, -1); xr = xr.reshape(b, c, -1)
    h = opnd(O.group_norm32(xr, sd[p + "norm.weight"], sd[p + "norm.bias"]))
    qkv = F.conv1d(h, W(sd, p + "qkv.weight"), sd[p + "qkv.bias"])
    if E.qkv: qkv = bf(qkv)
    h = O.qkv_attention(qkv, n_heads, new_order)
    if E.attn_out: h = bf(h)
    h = F.conv1d(h, W(sd, p + "proj_out.weight"), sd[p + "proj_out.bias"])
    r = xe if E.stream_hilo else xr
    oe, orr = stream(r + h)
    return (oe.reshape(b, c, *sp), orr.reshape(b, c, *sp))

def run_layers(sd, prefix, layers, h, emb, new_order):
    for j, layer in enumerate(layers):
        p = f"{prefix}{j}."; kind = layer[0]
        if kind == "conv_in":
            o = F.conv2d(h[1], W(sd, p + "weight"), sd[p + "bias"], padding=1); h = stream(o)
        elif kind == "res": h = res_block(sd, p, h, emb)
        elif kind == "attn": h = attention_block(sd, p, h, layer[2], new_order)
        elif kind == "down":
            o = F.conv2d(h[1], W(sd, p + "op.weight"), sd[p + "op.bias"], stride=2, padding=1); h = stream(o)
        elif kind == "up":
            o = F.interpolate(h[1], scale_factor=2, mode="nearest")
            o = F.conv2d(o, W(sd, p + "conv.weight"), sd[p + "conv.bias"], padding=1); h = stream(o)
    return h

@torch.no_grad()
def forward(sd, cfg, x, t):
    blocks = O.enumerate_blocks(cfg)
    emb = O.timestep_embedding(t, cfg["model_channels"])
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    hs = []
    h = (x.float(), x.float())
    for i, layers in enumerate(blocks["input"]):
        h = run_layers(sd, f"input_blocks.{i}.", layers, h, emb, False); hs.append(h)
    h = run_layers(sd, "middle_block.", blocks["middle"], h, emb, False)
    for i, layers in enumerate(blocks["output"]):
        s = hs.pop()
        h = (torch.cat([h[0], s[0]], 1), torch.cat([h[1], s[1]], 1))
        h = run_layers(sd, f"output_blocks.{i}.", layers, h, emb, False)
    o = opnd(F.silu(O.group_norm32(h[1], sd["out.0.weight"], sd["out.0.bias"])))
    return F.conv2d(o, W(sd, "out.2.weight"), sd["out.2.bias"], padding=1)

def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    name = args[0] if args else "base64_eps_t500"
    g = golden(name); cfg = golden_cfg(g)
    m = build_unet(cfg, int(g["init_seed"]), int(g["dezero_seed"]))
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    x, t = tt(g["x"]), tt(g["t"])
    fc = O.full_cfg(**cfg)
    ref = tt(g["eps"])
    def run(label, **kw):
        for k in ("weights","operands","conv1_out","stream","qkv","attn_out"): setattr(E, k, True)
        E.stream_hilo = False
        for k, v in kw.items(): setattr(E, k, v)
        got = forward(sd, fc, x, t)
        print(f"{label:50s} rel L2 {O.rel_l2(got, ref):.3e}", flush=True)
    for k in ("weights","operands","conv1_out","stream","qkv","attn_out"): setattr(E, k, False)
    print("sanity (nothing rounded):", O.rel_l2(forward(sd, fc, x, t), ref))
    run("all bf16 (engine emulation)")
    run("all bf16, stream hi+lo on the residual adds", stream_hilo=True)
    run("all bf16, stream exact everywhere", stream=False)
    run("all bf16, conv1_out exact", conv1_out=False)
    run("all bf16, weights exact", weights=False)
    run("only stream", weights=False, operands=False, conv1_out=False, qkv=False, attn_out=False)
    run("only stream, hi+lo", weights=False, operands=False, conv1_out=False, qkv=False, attn_out=False, stream_hilo=True)
main()
