"""Graph-timed 3x3 convs of the 32x2 UNet level (8 M tiles): A/B for TNG_GEMM_SPLITK."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tango_b200 import lib as L, ops
dev = torch.device("cuda:0")
def conv(NB, H, W, Cin, Cout, reps=20):
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / math.sqrt(9 * Cin)
    pc = ops.PackedConv(w, torch.zeros(Cout, device=dev), split=False, device=dev)
    x = torch.randn(NB * H * W, Cin, device=dev).to(torch.bfloat16)
    r = torch.randn(NB * H * W, Cout, device=dev)
    of = torch.empty(NB * H * W, Cout, device=dev)
    fn = lambda: ops.run_conv(pc, x, NB, H, W, res=r, out_f32=of)
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"SPLITK={os.environ.get('TNG_GEMM_SPLITK','1')} conv {NB}x{H}x{W} {Cin}->{Cout}: {us:7.1f} us  {2.0*NB*H*W*Cout*9*Cin/us/1e6:7.1f} TF/s")
conv(16, 32, 2, 1280, 1280); conv(16, 32, 2, 2560, 1280); conv(16, 64, 4, 1280, 1280); sys.stdout.flush()
