"""Launch a few GEMM/conv/attention shapes once each (for `ncu --set full`)."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tango_b200 import lib as L, ops
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
def lin(M, N, K, res, f32, bf):
    w = torch.randn(N, K, device=dev) / math.sqrt(K)
    pc = ops.PackedConv(w, torch.zeros(N, device=dev), split=False, device=dev)
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    r = torch.randn(M, N, device=dev) if res else None
    of = torch.empty(M, N, device=dev) if f32 else None
    ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if bf else None
    for _ in range(3):
        ops.run_linear(pc, x, res=r, out_f32=of, out_bf16=ob)
    torch.cuda.synchronize()
def conv(NB, H, W, Cin, Cout):
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / math.sqrt(9 * Cin)
    pc = ops.PackedConv(w, torch.zeros(Cout, device=dev), split=False, device=dev)
    x = torch.randn(NB * H * W, Cin, device=dev).to(torch.bfloat16)
    r = torch.randn(NB * H * W, Cout, device=dev)
    of = torch.empty(NB * H * W, Cout, device=dev)
    for _ in range(3):
        ops.run_conv(pc, x, NB, H, W, res=r, out_f32=of)
    torch.cuda.synchronize()
def attn(B, heads, Lq):
    Cc = heads * 64
    qkv = torch.randn(B * Lq, 3 * Cc, device=dev).to(torch.bfloat16)
    out = torch.empty(B * Lq, Cc, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        L.attention(qkv, qkv, qkv, out, batch=B, heads=heads, Lq=Lq, Lk=Lq, scale=0.125, k_col0=Cc, v_col0=2 * Cc)
    torch.cuda.synchronize()
def norms():
    NB, HW, Cc = 16, 4096, 320
    x = torch.randn(NB * HW, Cc, device=dev)
    st = torch.zeros(NB, Cc, 2, device=dev, dtype=torch.float64)
    L.groupnorm_stats(x, NB, HW, st)
    g, b = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
    y = torch.empty(NB * HW, Cc, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        L.groupnorm(x, st, None, None, NB, HW, 32, g, b, 1e-5, L.ACT_SILU, y)
        L.layernorm(x, g, b, 1e-5, y)
    torch.cuda.synchronize()
def geglu(M, N, K):
    w = torch.randn(N, K, device=dev) / math.sqrt(K)
    pc = ops.PackedConv(w, torch.zeros(N, device=dev), split=False, device=dev, geglu_bn=256)
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    ob = torch.empty(M, N // 2, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.run_linear(pc, x, out_bf16=ob)
    torch.cuda.synchronize()
if which in ("all", "norms"):
    norms()
if which in ("all", "geglu"):
    geglu(65536, 2560, 320)
if which in ("all", "lin"):
    lin(65536, 320, 320, True, True, False)
    lin(65536, 960, 320, False, False, True)
if which in ("all", "conv"):
    conv(16, 256, 16, 320, 320)
    conv(16, 128, 8, 1280, 640)
if which in ("all", "attn"):
    attn(16, 5, 4096)
