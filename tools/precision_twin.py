#!/usr/bin/env python3
"""CPU model of the engine's precision choices (torch, no GPU): where do the LSBs against the reference come from, and what
would each candidate change buy?  The network is evaluated on single-tile goldens with activations rounded to fp16 exactly
where the engine stores fp16 (fp32 accumulation everywhere), then with one choice changed at a time:
    engine      : what csrc/ does (fp32 Conv_0, fp16 dense planes, the residual stream -- dense-block residual x0 AND the RRDB skip R --
                  carried as fp16 + int8 extension (kernels.h res_ext_encode: 1/256 of an fp16 ulp), fp16 tail activations)
    rr32        : the RRDB skip stream kept in fp32 instead (round 2's first carrier version)
    res16       : round-1 default: dense-block residual from the fp16 operand plane alone, fp32 RRDB skip
    res32       : round-1 option: dense-block residual from an fp32 copy, fp32 RRDB skip
    tail32      : + the four tail convs (after the trunk) read and write fp32 activations ("exact tail")
    subpix      : engine precision, but the two "nearest x2 + conv" pairs collapsed into four 2x2 sub-pixel convs on the low-res
                  grid with pre-summed fp16 weights (2.25x fewer MACs for those two convs, no upsampled tensors)
    wino        : engine precision, every 3x3 conv as Winograd F(2x2,3x3) with fp16-rounded transformed operands (2.25x fewer MACs)
    hr32 / v2hr32 / trunk32 : engine precision, but only the HR conv's output / the last two tail tensors / the trunk output exact
    body32      : fp16 only in the tail
    all32       : no rounding at all (the reference itself, up to summation order)
Prints max |diff|, share of bytes off by one and bytes off by more than one against the reference golden.
Usage: precision_twin.py [golden names...]   (default: the single-tile goldens)"""
import os, sys
import numpy as np
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import golden, manifest, read_convs, diff_stats

torch.set_num_threads(8)
H16 = lambda t: t.half().float()
ID = lambda t: t


def EXT8(t):
    """fp16 value + int8 extension counting 1/256 of its ulp, computed on the fp32 bit patterns exactly like kernels.h res_ext_encode /
    res_ext_decode"""
    hf = t.half().float()
    bv, bh = t.contiguous().view(torch.int32).to(torch.int64), hf.contiguous().view(torch.int32).to(torch.int64)
    q = torch.clamp((bv - bh) >> 5, -128, 127)
    return (bh + q * 32).to(torch.int32).view(torch.float32)


def upconv_collapsed(t, c):
    """nearest x2 followed by the 3x3 conv, evaluated on the LOW-RES grid: output phase (dy, dx) is a 2x2 conv whose weights
    are sums of the 3x3 taps that land on the same source pixel -- summed in fp32, then rounded to fp16 like any MMA operand"""
    w = torch.from_numpy(c["w"].astype(np.float32))  # [O, I, 3, 3]
    b = torch.from_numpy(c["b"].copy())
    O = w.shape[0]
    out = torch.zeros((1, O, t.shape[2] * 2, t.shape[3] * 2))
    rows = {0: [(0, [0]), (1, [1, 2])], 1: [(1, [0, 1]), (2, [2])]}  # phase -> [(position in a 3-tap window around i, taps summed)]
    for dy in (0, 1):
        for dx in (0, 1):
            k = torch.zeros_like(w)
            for (py, tys) in rows[dy]:
                for (px, txs) in rows[dx]:
                    k[:, :, py, px] = sum(w[:, :, ty, tx] for ty in tys for tx in txs)
            k = k.half().float()
            out[:, :, dy::2, dx::2] = F.conv2d(t, k, b, padding=1)
    return F.leaky_relu(out, 0.2) if c["act"] == 2 else out


def conv_winograd(t, c):
    """Winograd F(2x2, 3x3) with the operand rounding a tensor-core implementation would have: U = G g G^T and V = B^T d B are
    formed in fp32 from the fp16 weights / activations and rounded to fp16 (the MMA operands), the 16 channel contractions
    accumulate in fp32, the output transform is fp32."""
    w = torch.from_numpy(c["w"].astype(np.float32))
    b = torch.from_numpy(c["b"].copy())
    G = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float32)
    Bt = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float32)
    At = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float32)
    U = torch.einsum("ij,ocjk,lk->ocil", G, w, G).half().float()           # [O, C, 4, 4]
    _, C, H, W = t.shape
    Hp, Wp = (H + 1) // 2 * 2, (W + 1) // 2 * 2
    tp = F.pad(t, (1, 1 + Wp - W, 1, 1 + Hp - H))
    tiles = tp.unfold(2, 4, 2).unfold(3, 4, 2)                                # [1, C, Hp/2, Wp/2, 4, 4]
    V = torch.einsum("ij,nchwjk,lk->nchwil", Bt, tiles, Bt).half().float()
    M = torch.einsum("ocil,nchwil->nohwil", U, V)
    Y = torch.einsum("ij,nohwjk,lk->nohwil", At, M, At)                       # [1, O, Hp/2, Wp/2, 2, 2]
    y = Y.permute(0, 1, 2, 4, 3, 5).reshape(1, w.shape[0], Hp, Wp)[:, :, :H, :W] + b.view(1, -1, 1, 1)
    return F.leaky_relu(y, 0.2) if c["act"] == 2 else y


def net(x, convs, q_body, q_x0res, q_tail, nb=23, q_tails=None, collapse=False, wino=False, q_rrdb=None):
    """x: [1,3,H,W] fp32 in 0..1.  q_*: rounding applied where the engine stores fp16."""
    def conv(i, t):
        c = convs[i]
        if wino and i > 0:
            return conv_winograd(t, c)
        y = F.conv2d(t, torch.from_numpy(c["w"].astype(np.float32)), torch.from_numpy(c["b"].copy()), padding=1)
        return F.leaky_relu(y, 0.2) if c["act"] == 2 else y
    fea = conv(0, x)                      # fp32 (Conv_0 on CUDA cores)
    r32 = fea
    t32 = fea
    x0 = q_body(fea)
    idx = 1
    for rrdb in range(nb):
        for rdb in range(3):
            feats = [x0]
            for k in range(4):
                feats.append(q_body(conv(idx + k, torch.cat(feats, 1))))
            x5 = conv(idx + 4, torch.cat(feats, 1))
            t = 0.2 * x5 + (q_x0res(t32) if q_x0res is not ID else t32)
            if rdb == 2:
                t = 0.2 * t + (q_rrdb(r32) if q_rrdb else r32)
                r32 = t
            t32 = t
            x0 = q_body(t)
            idx += 5
    qt = q_tails or (q_tail,) * 4         # per tensor: trunk, v1 (2x), v2 (4x), hr (4x, the last conv's input)
    trunk = qt[0](conv(idx, x0) + fea)
    if collapse:
        v1 = qt[1](upconv_collapsed(trunk, convs[idx + 1]))
        v2 = qt[2](upconv_collapsed(v1, convs[idx + 2]))
    else:
        u1 = F.interpolate(trunk, scale_factor=2, mode="nearest")
        v1 = qt[1](conv(idx + 1, u1))
        u2 = F.interpolate(v1, scale_factor=2, mode="nearest")
        v2 = qt[2](conv(idx + 2, u2))
    hr = qt[3](conv(idx + 3, v2))
    return conv(idx + 4, hr)


def run(name, mode):
    case = [c for c in manifest()["cases"] if c["name"] == name][0]
    img, gold = golden(name)
    img = img[..., :3]
    h, w, _ = img.shape
    assert h <= case["tilesize"] and w <= case["tilesize"], "single-tile goldens only"
    convs = read_convs(case["model"])
    P = 10
    x = torch.from_numpy(img.astype(np.float32)).permute(2, 0, 1)[None] * (1 / 255.0)
    x = F.pad(x, (P, P, P, P), mode="reflect")
    q = dict(engine=(H16, EXT8, H16), rr32=(H16, EXT8, H16), res16=(H16, H16, H16), res32=(H16, ID, H16), tail32=(H16, ID, ID), body32=(ID, ID, H16), all32=(ID, ID, ID),
             hr32=(H16, H16, H16), v2hr32=(H16, H16, H16), trunk32=(H16, H16, H16), subpix=(H16, H16, H16), wino=(H16, H16, H16))[mode]
    # single tail tensors kept exact (engine precision elsewhere): which hand-off carries the error?
    q_tails = dict(hr32=(H16, H16, H16, ID), v2hr32=(H16, H16, ID, ID), trunk32=(ID, H16, H16, H16)).get(mode)
    with torch.no_grad():
        y = net(x, convs, *q, q_tails=q_tails, collapse=(mode == "subpix"), wino=(mode == "wino"), q_rrdb=(EXT8 if mode == "engine" else None))[0, :, 4 * P:4 * P + 4 * h, 4 * P:4 * P + 4 * w]
    out = torch.clamp((y * 255.0 + 0.5).to(torch.int32), 0, 255).permute(1, 2, 0).numpy().astype(np.uint8)  # (int) truncation like the reference
    return diff_stats(out, gold[..., :3])


if __name__ == "__main__":
    names = sys.argv[1:] or ["noise_40x36_jpeg_t200", "readme0_crop64_df2k_t200", "natural_28x20_df2k_tta"]
    for name in names:
        if "tta" in name:
            continue  # TTA goldens average 8 passes: not modelled here
        for mode in ("engine", "rr32", "res16", "res32", "tail32", "hr32", "v2hr32", "trunk32", "subpix", "wino", "body32", "all32"):
            s = run(name, mode)
            print("%-28s %-7s max %d  off-by-one %.3f %%  >1: %d of %d" % (name, mode, s["max"], 100 * s["frac_diff"], s["n_gt1"], s["n"]), flush=True)
