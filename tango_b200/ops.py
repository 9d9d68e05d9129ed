"""Host-side operator layer: weight packing and k-group construction for the tcgen05 conv/GEMM kernel.

Activations are channels-last row matrices:
  * fp32 [rows, C]              — the residual stream (hidden states),
  * bf16 [rows, C]              — tensor-core operands in perf mode ("bf16"),
  * bf16 [rows, 2C] = [hi | lo] — operands in parity mode ("split"): lo is the bf16 rounding residual, and every
                                   product is evaluated as hi*hi + lo*hi + hi*lo (~fp32 accuracy on bf16 MMAs).
Weights get the same treatment along K. See include/tango_b200.h for the kernel contract.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import lib as L

BK = 64


def ceil_div(a: int, b: int) -> int:
    return (a + b - 1) // b


def split_hi_lo(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    hi = w.to(torch.bfloat16)
    lo = (w.float() - hi.float()).to(torch.bfloat16)
    return hi, lo


class PackedConv:
    """A conv / linear layer packed for tng_conv_gemm.

    w: [Cout, Cin] (linear / 1x1), [Cout, Cin, k] (conv1d) or [Cout, Cin, kh, kw] (conv2d), fp32.
    K layout of the packed bf16 matrix: [tap0 cin.. | tap1 cin.. | ... | shortcut cin_sc..] (+ the same again for
    the lo halves in split mode).
    """

    def __init__(self, w: torch.Tensor, b: Optional[torch.Tensor], *, split: bool, device, dilation: int = 1,
                 stride: int = 1, sc_w: Optional[torch.Tensor] = None, sc_b: Optional[torch.Tensor] = None,
                 geglu_bn: int = 0, pad: Optional[int] = None, geglu_tanh: bool = False):
        w = w.detach().float()
        if w.dim() == 2:
            w = w[:, :, None, None]
            taps = [(0, 0)]
        elif w.dim() == 3:  # conv1d: positions run along W
            k = w.shape[2]
            p = (k * dilation - dilation) // 2 if pad is None else pad
            taps = [(j * dilation - p, 0) for j in range(k)]
            w = w[:, :, None, :]
        else:
            kh, kw = w.shape[2], w.shape[3]
            ph, pw = (kh // 2, kw // 2) if pad is None else (pad, pad)
            taps = [(kx - pw, ky - ph) for ky in range(kh) for kx in range(kw)]
        self.cout, self.cin = w.shape[0], w.shape[1]
        self.taps = taps
        self.stride = stride
        self.split = split
        self.geglu_bn = geglu_bn
        self.geglu_act = L.ACT_GEGLU_TANH if geglu_tanh else L.ACT_GEGLU
        wk = w.permute(0, 2, 3, 1).reshape(self.cout, len(taps) * self.cin)  # [Cout, tap*Cin]
        self.k_main = wk.shape[1]
        self.cin_sc = 0
        bias = b.detach().float().clone() if b is not None else None
        if sc_w is not None:
            scw = sc_w.detach().float().reshape(self.cout, -1)
            self.cin_sc = scw.shape[1]
            wk = torch.cat([wk, scw], dim=1)
            if sc_b is not None:
                bias = sc_b.detach().float() if bias is None else bias + sc_b.detach().float()
        if wk.shape[1] % 8:
            wk = torch.nn.functional.pad(wk, (0, 8 - wk.shape[1] % 8))
        if geglu_bn:
            # interleave hidden / gate rows so that both halves of a GEGLU pair land in the same N tile
            inner = self.cout // 2
            half = geglu_bn // 2
            assert inner % half == 0
            idx = []
            for i in range(inner // half):
                idx += list(range(i * half, (i + 1) * half))
                idx += list(range(inner + i * half, inner + (i + 1) * half))
            idx = torch.tensor(idx)
            wk = wk[idx]
            if bias is not None:
                bias = bias[idx]
        self.k_half = wk.shape[1]
        if split:
            hi, lo = split_hi_lo(wk)
            packed = torch.cat([hi, lo], dim=1)
        else:
            packed = wk.to(torch.bfloat16)
        self.weight = packed.contiguous().to(device)
        self.bias = bias.contiguous().to(device) if bias is not None else None

    # -- k-groups -------------------------------------------------------------------------------------------
    def groups(self, cin_views: Sequence[int] = (0,), sc_view: Optional[int] = None,
               parity_views: Optional[Sequence[int]] = None, lo_views: Optional[Sequence[int]] = None,
               sc_lo_view: Optional[int] = None) -> List[tuple]:
        """Build the (view, a_c0, dw, dh, b_k0, nkb) list.

        Normal convs read view `cin_views[0]`. Stride-2 convs read four parity views (index = 2*ph + pw).
        In split mode the lo half lives at channel offset cin of the same view when cin % 64 == 0, otherwise in
        the separate view `lo_views[i]`.
        """
        g: List[tuple] = []
        nkb = ceil_div(self.cin, BK)
        same_view_lo = (self.cin % BK == 0)

        def add(view, lo_view, dw, dh, bk, cin, nkb_):
            if not self.split:
                g.append((view, 0, dw, dh, bk, nkb_))
                return
            lv, lo_c0 = (view, cin) if (cin % BK == 0 and lo_view is None) else (lo_view, 0)
            g.append((view, 0, dw, dh, bk, nkb_))                       # hi * w_hi
            g.append((lv, lo_c0, dw, dh, bk, nkb_))                      # lo * w_hi
            g.append((view, 0, dw, dh, self.k_half + bk, nkb_))          # hi * w_lo

        for t, (dw, dh) in enumerate(self.taps):
            if self.stride == 2:
                # input coord = 2*out + d: parity plane ((d) & 1), plane coordinate out + floor(d / 2)
                ph, pw = dh & 1, dw & 1
                v = parity_views[2 * ph + pw]
                add(v, None, dw >> 1, dh >> 1, t * self.cin, self.cin, nkb)
            else:
                lo_v = None if (same_view_lo or lo_views is None) else lo_views[0]
                add(cin_views[0], lo_v, dw, dh, t * self.cin, self.cin, nkb)
        if self.cin_sc:
            lo_v = None if self.cin_sc % BK == 0 else sc_lo_view
            add(sc_view, lo_v, 0, 0, self.k_main, self.cin_sc, ceil_div(self.cin_sc, BK))
        return g


def act_views(x: torch.Tensor, NB: int, H: int, W: int, cin: int, split: bool) -> List[L.View]:
    """Views for a bf16 operand [rows, cin] (or [rows, 2*cin] = [hi|lo] in split mode).

    Returns [view] or, when split and cin is not a multiple of 64, [hi_view, lo_view]."""
    ld = x.stride(0)
    s_w, s_h, s_n = ld, W * ld, H * W * ld
    if not split:
        return [L.View(x, cin, W, H, NB, s_w, s_h, s_n)]
    if cin % BK == 0:
        return [L.View(x, 2 * cin, W, H, NB, s_w, s_h, s_n)]
    return [L.View(x, cin, W, H, NB, s_w, s_h, s_n), L.View(x, cin, W, H, NB, s_w, s_h, s_n, off=cin)]


def parity_views(x: torch.Tensor, NB: int, H: int, W: int, cin: int, split: bool) -> List[L.View]:
    """Four strided views (ph, pw) of a channels-last tensor for a stride-2 conv (cin % 64 == 0 in split mode)."""
    ld = x.stride(0)
    C_ = 2 * cin if split else cin
    out = []
    for ph in range(2):
        for pw in range(2):
            out.append(L.View(x, C_, (W - pw + 1) // 2, (H - ph + 1) // 2, NB, 2 * ld, 2 * W * ld, H * W * ld,
                              off=(ph * W + pw) * ld))
    return out


def run_conv(pc: PackedConv, x: torch.Tensor, NB: int, H: int, W: int, *, sc_x: Optional[torch.Tensor] = None,
             rowvec=None, res=None, alpha: float = 1.0, accumulate: bool = False, out_f32=None, out_bf16=None,
             act: int = L.ACT_NONE, act_param: float = 0.0, block_n: int = 0, rowvec_ld: int = 0,
             gn_stats: Optional[torch.Tensor] = None, stats_hw: int = 0) -> None:
    """Convolution / linear of the bf16 operand x ([rows, cin] or [hi|lo]) on the (NB, H, W) grid.

    For stride 2 (H, W) is the *input* grid; the output grid is (H/2, W/2).
    gn_stats / stats_hw: per-channel GroupNorm accumulators of the fp32 output (see lib.conv_gemm)."""
    split = pc.split
    views: List[L.View] = []
    if pc.stride == 2:
        assert (not split) or pc.cin % BK == 0
        views = parity_views(x, NB, H, W, pc.cin, split)
        Ho, Wo = (H + 1) // 2, (W + 1) // 2
        groups = pc.groups(parity_views=[0, 1, 2, 3])
    else:
        views = act_views(x, NB, H, W, pc.cin, split)
        lo_views = [1] if len(views) == 2 else None
        sc_view = sc_lo = None
        if pc.cin_sc:
            scv = act_views(sc_x, NB, H, W, pc.cin_sc, split)
            sc_view = len(views)
            views += scv
            sc_lo = sc_view + 1 if len(scv) == 2 else None
        Ho, Wo = H, W
        groups = pc.groups(cin_views=[0], lo_views=lo_views, sc_view=sc_view, sc_lo_view=sc_lo)
    so = 0
    if out_bf16 is not None and split:
        so = out_bf16.shape[-1] // 2
    if pc.geglu_bn:
        act, block_n = pc.geglu_act, pc.geglu_bn
    L.conv_gemm(views, groups, pc.weight, Wo, Ho, NB, bias=pc.bias, rowvec=rowvec, res=res, alpha=alpha,
                accumulate=accumulate, out_f32=out_f32, out_bf16=out_bf16, act=act, act_param=act_param,
                split_off=so, block_n=block_n, rowvec_ld=rowvec_ld,
                algo_k=len(pc.taps) * pc.cin + pc.cin_sc, gn_stats=gn_stats, stats_hw=stats_hw)


def run_linear(pc: PackedConv, x: torch.Tensor, **kw) -> None:
    """Linear layer over the rows of x (rows on the W axis of a 1 x 1 x rows grid)."""
    rows = x.shape[0]
    run_conv(pc, x, 1, 1, rows, **kw)
