import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tango_b200 import lib as L, ops
dev = torch.device("cuda:0")
def bench(name, fn, flops, reps=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"dbg={os.environ.get('TNG_GEMM_DBG','0')} {name:44s} {us:8.1f} us  {flops/us/1e6:8.1f} TF/s(nominal)")
def lin(M, N, K, bn=0):
    w = torch.randn(N, K, device=dev) / math.sqrt(K)
    pc = ops.PackedConv(w, None, split=False, device=dev)
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    bench(f"linear M={M} N={N} K={K} bn={bn}", lambda: ops.run_linear(pc, x, out_bf16=ob, block_n=bn), 2.0 * M * N * K)
for bn in (128, 160, 256):
    lin(8192, 8192 if bn != 160 else 8000, 8192, bn)
