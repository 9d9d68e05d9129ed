import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tango_b200 import lib as L, ops
dev = torch.device("cuda:0")
def bench(name, fn, flops, reps=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"dbg={os.environ.get('TNG_GEMM_DBG','0')} {name:44s} {us:8.1f} us  {flops/us/1e6:8.1f} TF/s")
def lin(M, N, K, res=True, bn=0):
    w = torch.randn(N, K, device=dev) / math.sqrt(K)
    pc = ops.PackedConv(w, torch.zeros(N, device=dev), split=False, device=dev)
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    r = torch.randn(M, N, device=dev) if res else None
    of = torch.empty(M, N, device=dev)
    bench(f"linear M={M} N={N} K={K} res={int(res)} bn={bn}", lambda: ops.run_linear(pc, x, res=r, out_f32=of, block_n=bn), 2.0 * M * N * K)
for K in (320, 640, 1280, 2560, 5120):
    lin(65536, 320, K, res=False)
lin(65536, 320, 320, res=True)
lin(8192, 8192, 8192, res=False, bn=160)
lin(8192, 8192, 8192, res=False, bn=256)
