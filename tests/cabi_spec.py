"""Executable statement of what `tng_conv_gemm` computes from its descriptor (TEST INFRASTRUCTURE ONLY).

include/tango_b200.h describes the operator in prose; this is the same contract in a few lines of torch on the CPU, so
that the host-side packing logic (tango_b200/ops.py: weight layout, k-groups, parity views, hi/lo splitting, GEGLU row
interleave) can be checked against torch convolutions without a GPU. It is never imported by the package.

For output pixel (n, h, w) of the (NB, H, W) grid and output column j:
    acc[row, j] = sum over k-groups g, kk in [0, 64 * nkb):  A_g[row, kk] * B[j, b_k0 + kk]
    A_g[row, kk] = view[g.view] at (n, h + dh, w + dw, a_c0 + kk), ZERO outside the view's (NB, H, W, C) extent
    B[j, k] = weight[j, k], ZERO for k >= Ktot (what the TMA out-of-bounds fill provides)
    y = (acc + bias[j] + rowvec[n, j] + res[row, j]) * alpha (+ previous out_f32 if accumulate);  out_f32 = y
    out_bf16 = act(y) (SiLU / leaky-ReLU), or for GEGLU out[:, tn*BN/2 + i] = y[:, tn*BN + i] * gelu(y[:, tn*BN + BN/2 + i]);
    in split mode the bf16 rounding residual goes to column offset split_off.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

ACT_NONE, ACT_SILU, ACT_LRELU, ACT_GEGLU, ACT_GEGLU_TANH = 0, 1, 2, 3, 4
BK = 64


def _gather_view(v, n_idx, h_idx, w_idx, c0, nk):
    """[rows, nk] fp64 values of view v at (n, h, w, c0 + kk) with zero fill outside the view extent."""
    flat = v.t.reshape(-1).double()
    ok = (h_idx >= 0) & (h_idx < v.H) & (w_idx >= 0) & (w_idx < v.W) & (n_idx >= 0) & (n_idx < v.NB)
    base = v.off + n_idx.clamp(0, v.NB - 1) * v.s_n + h_idx.clamp(0, v.H - 1) * v.s_h + w_idx.clamp(0, v.W - 1) * v.s_w
    c = c0 + torch.arange(nk)
    c_ok = c < v.C
    idx = base[:, None] + c.clamp(max=max(v.C - 1, 0))[None, :]
    vals = flat[idx.clamp(0, flat.numel() - 1)]
    return vals * (ok[:, None] & c_ok[None, :])


def spec_conv_gemm(views, groups, weight, W, H, NB, *, bias=None, rowvec=None, res=None, alpha=1.0, accumulate=False,
                   out_f32=None, out_bf16=None, act=ACT_NONE, act_param=0.0, split_off=0, block_n=0, rowvec_ld=0, **_):
    rows = NB * H * W
    r = torch.arange(rows)
    n_idx, h_idx, w_idx = r // (H * W), (r // W) % H, r % W
    Ncols, Ktot = weight.shape
    wd = weight.double()
    acc = torch.zeros(rows, Ncols, dtype=torch.float64)
    for (vi, a_c0, dw, dh, b_k0, nkb) in groups:
        nk = nkb * BK
        a = _gather_view(views[vi], n_idx, h_idx + dh, w_idx + dw, a_c0, nk)
        b = torch.zeros(Ncols, nk, dtype=torch.float64)
        kmax = max(0, min(nk, Ktot - b_k0))
        b[:, :kmax] = wd[:, b_k0:b_k0 + kmax]
        acc += a @ b.t()
    y = acc
    if bias is not None:
        y = y + bias.double()[None, :]
    if rowvec is not None:
        ld = rowvec_ld or Ncols
        rv = rowvec.reshape(-1).double()
        y = y + rv[(n_idx * ld)[:, None] + torch.arange(Ncols)[None, :]]
    if res is not None:
        y = y + res.double()[:, :Ncols]
    y = y * alpha
    if out_f32 is not None:
        if accumulate:        # the accumulated value is also what the bf16 output (if any) is derived from
            y = y + out_f32[:, :Ncols].double()
        out_f32[:, :Ncols] = y.float()
    if out_bf16 is not None:
        z = y.float()
        if act == ACT_SILU:
            z = F.silu(z)
        elif act == ACT_LRELU:
            z = F.leaky_relu(z, act_param)
        elif act in (ACT_GEGLU, ACT_GEGLU_TANH):
            bn = block_n
            assert bn in (128, 256) and Ncols % bn == 0
            t = z.view(rows, Ncols // bn, bn)
            gate = F.gelu(t[..., bn // 2:], approximate="tanh" if act == ACT_GEGLU_TANH else "none")
            z = (t[..., :bn // 2] * gate).reshape(rows, Ncols // 2)
        hi = z.to(torch.bfloat16)
        out_bf16[:, :z.shape[1]] = hi
        if split_off > 0:
            out_bf16[:, split_off:split_off + z.shape[1]] = (z - hi.float()).to(torch.bfloat16)
