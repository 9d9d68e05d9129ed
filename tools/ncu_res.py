"""One launch of the level-0 K = C linear with an fp32 residual in and out (65536 x 320 x 320, the attention out-projection)
between cudaProfilerStart/Stop, for a source-level capture:
    ncu --set full --clock-control none --import-source on --profile-from-start off -o res python tools/ncu_res.py
    ncu -i res.ncu-rep --page source --csv --print-source cuda,sass"""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tango_b200 import ops
dev = torch.device("cuda:0")
w = torch.randn(320, 320, device=dev) / math.sqrt(320)
pc = ops.PackedConv(w, torch.zeros(320, device=dev), split=False, device=dev)
x = torch.randn(65536, 320, device=dev).to(torch.bfloat16)
res = torch.randn(65536, 320, device=dev)
out = torch.empty(65536, 320, device=dev)
def once():
    ops.run_linear(pc, x, res=res, out_f32=out)
for _ in range(3):
    once()
torch.cuda.synchronize()
torch.cuda.profiler.start()
once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
