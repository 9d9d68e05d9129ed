"""CPU: the host-side operator packing (tango_b200/ops.py) against torch convolutions, with `tng_conv_gemm` replaced by
its executable contract (tests/cabi_spec.py). Covers what only the GPU suite exercised before: tap / k-group layout,
TMA-style zero fill as conv padding, stride-2 parity views (diffusers' pad 1 and AudioLDM's right/bottom pad 0),
dilated conv1d, the fused 1x1 shortcut, hi/lo splitting (3-term products) and the GEGLU row interleave."""
import math

import pytest
import torch
import torch.nn.functional as F

from cabi_spec import spec_conv_gemm
from tango_b200 import lib as L
from tango_b200 import ops


@pytest.fixture(autouse=True)
def _spec_backend(monkeypatch):
    monkeypatch.setattr(L, "conv_gemm", spec_conv_gemm)


def bf(x):
    return x.to(torch.bfloat16)


def rows(x):   # [N,C,H,W] -> [N*H*W, C]
    return x.permute(0, 2, 3, 1).reshape(-1, x.shape[1]).contiguous()


def to_split(x):
    hi = bf(x)
    return torch.cat([hi, bf(x - hi.float())], dim=1).contiguous()


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("NB,H,W,Cin,Cout", [(2, 6, 5, 64, 32), (1, 4, 7, 72, 16), (2, 3, 3, 8, 24)])
@pytest.mark.parametrize("split", [False, True])
def test_conv3x3_padding_bias_rowvec_residual(NB, H, W, Cin, Cout, split):
    g = torch.Generator().manual_seed(NB + H + Cin)
    x = torch.randn(NB, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin)
    b, temb = torch.randn(Cout, generator=g), torch.randn(NB, Cout, generator=g)
    res = torch.randn(NB * H * W, Cout, generator=g)
    pc = ops.PackedConv(w, b, split=split, device="cpu")
    xin = to_split(rows(x)) if split else bf(rows(x))
    of = torch.full((NB * H * W, Cout), float("nan"))
    ops.run_conv(pc, xin, NB, H, W, rowvec=temb, res=res, alpha=0.5, out_f32=of)
    xr, wr = (x, w) if split else (bf(x).float(), bf(w).float())
    ref = (rows(F.conv2d(xr, wr, b, padding=1) + temb[:, :, None, None]) + res) * 0.5
    assert rel(of, ref) < (3e-5 if split else 1e-6)


@pytest.mark.parametrize("pad", [1, 0])
@pytest.mark.parametrize("split", [False, True])
def test_conv_stride2_parity_views(pad, split):
    """pad 1: diffusers Downsample2D (resnet.py:199-208); pad 0 on a right/bottom zero-padded input: AudioLDM's encoder
    Downsample (modules.py:76-94) — the out-of-range taps on the high side come from the zero fill."""
    NB, H, W, Cin, Cout = 2, 8, 6, 64, 48
    g = torch.Generator().manual_seed(pad)
    x = torch.randn(NB, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin)
    b = torch.randn(Cout, generator=g)
    pc = ops.PackedConv(w, b, split=split, device="cpu", stride=2, pad=pad)
    xin = to_split(rows(x)) if split else bf(rows(x))
    of = torch.empty(NB * (H // 2) * (W // 2), Cout)
    ops.run_conv(pc, xin, NB, H, W, out_f32=of)
    xr, wr = (x, w) if split else (bf(x).float(), bf(w).float())
    if pad:
        ref = F.conv2d(xr, wr, b, stride=2, padding=1)
    else:
        ref = F.conv2d(F.pad(xr, (0, 1, 0, 1)), wr, b, stride=2, padding=0)
    assert rel(of, rows(ref)) < (3e-5 if split else 1e-6)


@pytest.mark.parametrize("split", [False, True])
def test_conv_with_fused_1x1_shortcut(split):
    NB, H, W, Cin, Cout, Csc = 1, 5, 4, 64, 32, 24
    g = torch.Generator().manual_seed(3)
    x, xs = torch.randn(NB, Cin, H, W, generator=g), torch.randn(NB, Csc, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin)
    ws = torch.randn(Cout, Csc, 1, 1, generator=g) / math.sqrt(Csc)
    b, bs = torch.randn(Cout, generator=g), torch.randn(Cout, generator=g)
    pc = ops.PackedConv(w, b, split=split, device="cpu", sc_w=ws, sc_b=bs)
    xin, sin = (to_split(rows(x)), to_split(rows(xs))) if split else (bf(rows(x)), bf(rows(xs)))
    of = torch.empty(NB * H * W, Cout)
    ops.run_conv(pc, xin, NB, H, W, sc_x=sin, out_f32=of)
    c = (lambda t: t) if split else (lambda t: bf(t).float())
    ref = F.conv2d(c(x), c(w), b, padding=1) + F.conv2d(c(xs), c(ws), bs)
    assert rel(of, rows(ref)) < (3e-5 if split else 1e-6)


@pytest.mark.parametrize("k,dil", [(3, 1), (7, 3), (11, 5)])
def test_conv1d_dilated_hifigan(k, dil):
    B, Cc, Lx = 2, 64, 40
    g = torch.Generator().manual_seed(k)
    x = torch.randn(B, Cc, Lx, generator=g)
    w = torch.randn(Cc, Cc, k, generator=g) / math.sqrt(k * Cc)
    b = torch.randn(Cc, generator=g)
    pc = ops.PackedConv(w, b, split=False, device="cpu", dilation=dil)
    xin = bf(x.permute(0, 2, 1).reshape(B * Lx, Cc).contiguous())
    of = torch.empty(B * Lx, Cc)
    ops.run_conv(pc, xin, B, 1, Lx, out_f32=of)
    ref = F.conv1d(bf(x).float(), bf(w).float(), b, padding=(k * dil - dil) // 2, dilation=dil)
    assert rel(of, ref.permute(0, 2, 1).reshape(B * Lx, Cc)) < 1e-6


@pytest.mark.parametrize("tanh", [False, True])
@pytest.mark.parametrize("split", [False, True])
def test_gated_gelu_interleave(tanh, split):
    M, Cc, inner = 37, 64, 256
    g = torch.Generator().manual_seed(5)
    x = torch.randn(M, Cc, generator=g)
    w = torch.randn(2 * inner, Cc, generator=g) / math.sqrt(Cc)       # rows: [hidden | gate]
    b = torch.randn(2 * inner, generator=g)
    pc = ops.PackedConv(w, b, split=split, device="cpu", geglu_bn=128, geglu_tanh=tanh)
    s = 2 if split else 1
    ob = torch.zeros(M, s * inner, dtype=torch.bfloat16)
    ops.run_linear(pc, to_split(x) if split else bf(x), out_bf16=ob)
    c = (lambda t: t) if split else (lambda t: bf(t).float())
    proj = c(x) @ c(w).t() + b
    ref = proj[:, :inner] * F.gelu(proj[:, inner:], approximate="tanh" if tanh else "none")
    got = ob[:, :inner].float() + (ob[:, inner:].float() if split else 0)
    assert rel(got, ref) < (3e-5 if split else 3e-3)                  # single bf16 output rounding in perf mode


def test_silu_bf16_output_and_accumulate():
    M, K, N = 19, 128, 40
    g = torch.Generator().manual_seed(9)
    x, w = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / math.sqrt(K)
    pc = ops.PackedConv(w, None, split=False, device="cpu")
    of = torch.ones(M, N)
    ob = torch.zeros(M, N, dtype=torch.bfloat16)
    ops.run_linear(pc, bf(x), out_f32=of, out_bf16=ob, accumulate=True, act=L.ACT_SILU)
    y = bf(x).float() @ bf(w).float().t()
    assert rel(of, y + 1.0) < 1e-6
    assert rel(ob.float(), F.silu(y + 1.0)) < 3e-3
