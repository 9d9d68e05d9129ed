"""One launch of each hot kernel shape between cudaProfilerStart/Stop (after warm-ups outside the range), for
    ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/r2_full python tools/ncu_one.py
Shapes = the UNet-batch-16 / 10 s step of the benchmark: 3x3 conv 320->320 at 256x16 and 1280->640 at 128x8 (CTA-pair
tiles), the K = 320 linear with fp32 residual, the GEGLU GEMM, self-attention L = 4096, GroupNorm apply, LayerNorm, and
the VAE's one-head d = 512 attention at HW = 4096."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tango_b200 import lib as L, ops
dev = torch.device("cuda:0")
jobs = []

def lin(M, N, K, res, f32, bf, geglu=0):
    w = torch.randn(N, K, device=dev) / math.sqrt(K)
    pc = ops.PackedConv(w, torch.zeros(N, device=dev), split=False, device=dev, geglu_bn=geglu)
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    No = N // 2 if geglu else N
    r = torch.randn(M, No, device=dev) if res else None
    of = torch.empty(M, No, device=dev) if f32 else None
    ob = torch.empty(M, No, device=dev, dtype=torch.bfloat16) if bf else None
    jobs.append(lambda: ops.run_linear(pc, x, res=r, out_f32=of, out_bf16=ob))

def conv(NB, H, W, Cin, Cout):
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / math.sqrt(9 * Cin)
    pc = ops.PackedConv(w, torch.zeros(Cout, device=dev), split=False, device=dev)
    x = torch.randn(NB * H * W, Cin, device=dev).to(torch.bfloat16)
    r = torch.randn(NB * H * W, Cout, device=dev)
    of = torch.empty(NB * H * W, Cout, device=dev)
    st = torch.zeros(NB, Cout, 2, device=dev, dtype=torch.float64)
    jobs.append(lambda: ops.run_conv(pc, x, NB, H, W, res=r, out_f32=of, gn_stats=st, stats_hw=H * W))

def attn(B, heads, Lq):
    Cc = heads * 64
    qkv = torch.randn(B * Lq, 3 * Cc, device=dev).to(torch.bfloat16)
    out = torch.empty(B * Lq, Cc, device=dev, dtype=torch.bfloat16)
    jobs.append(lambda: L.attention(qkv, qkv, qkv, out, batch=B, heads=heads, Lq=Lq, Lk=Lq, scale=0.125, k_col0=Cc, v_col0=2 * Cc))

def attn_wide(B, Lq):
    qkv = torch.randn(B * Lq, 1536, device=dev).to(torch.bfloat16)
    out = torch.empty(B * Lq, 512, device=dev, dtype=torch.bfloat16)
    jobs.append(lambda: L.attention_wide(qkv, qkv, qkv, out, batch=B, L=Lq, dim=512, scale=512 ** -0.5, k_col0=512, v_col0=1024))

def norms():
    NB, HW, Cc = 16, 4096, 320
    x = torch.randn(NB * HW, Cc, device=dev)
    st = torch.zeros(NB, Cc, 2, device=dev, dtype=torch.float64)
    L.groupnorm_stats(x, NB, HW, st)
    g, b = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
    y = torch.empty(NB * HW, Cc, device=dev, dtype=torch.bfloat16)
    jobs.append(lambda: L.groupnorm(x, st, None, None, NB, HW, 32, g, b, 1e-5, L.ACT_SILU, y))
    jobs.append(lambda: L.layernorm(x, g, b, 1e-5, y))

conv(16, 256, 16, 320, 320)
conv(16, 128, 8, 1280, 640)
lin(65536, 320, 320, True, True, False)
lin(65536, 2560, 320, False, False, True, geglu=256)
attn(16, 5, 4096)
norms()
attn_wide(8, 4096)
for _ in range(2):
    for j in jobs:
        j()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for j in jobs:
    j()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
