"""One launch each of the pair-tile 3x3 conv (1280 -> 640 at 128x8, UNet batch 16) and the level-0 GEGLU GEMM between
cudaProfilerStart/Stop, for a source-level capture:
    ncu --set full --clock-control none --import-source on --profile-from-start off -o two python tools/ncu_two.py
    ncu -i two.ncu-rep --page source --csv   (per-instruction samples; needs the -lineinfo build, which is the default)"""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tango_b200 import ops
dev = torch.device("cuda:0")
w = torch.randn(640, 1280, 3, 3, device=dev) / math.sqrt(9 * 1280)
pc = ops.PackedConv(w, torch.zeros(640, device=dev), split=False, device=dev)
x = torch.randn(16 * 128 * 8, 1280, device=dev).to(torch.bfloat16)
of = torch.empty(16 * 128 * 8, 640, device=dev)
st = torch.zeros(16, 640, 2, device=dev, dtype=torch.float64)
wg = torch.randn(2560, 320, device=dev) / math.sqrt(320)
pg = ops.PackedConv(wg, torch.zeros(2560, device=dev), split=False, device=dev, geglu_bn=256)
xg = torch.randn(65536, 320, device=dev).to(torch.bfloat16)
og = torch.empty(65536, 1280, device=dev, dtype=torch.bfloat16)
def once():
    ops.run_conv(pc, x, 16, 128, 8, out_f32=of, gn_stats=st, stats_hw=1024)
    ops.run_linear(pg, xg, out_bf16=og)
for _ in range(3):
    once()
torch.cuda.synchronize()
torch.cuda.profiler.start()
once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
