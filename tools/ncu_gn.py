"""One launch of gn_apply (GroupNorm 32 + SiLU, 16 x 4096 pixels x 320 channels, fp32 in, bf16 out) between
cudaProfilerStart/Stop:
    ncu --set full --clock-control none --profile-from-start off -o gn python tools/ncu_gn.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tango_b200 import lib as L
dev = torch.device("cuda:0")
NB, HW, C = 16, 4096, 320
x = torch.randn(NB * HW, C, device=dev)
st = torch.zeros(NB, C, 2, device=dev, dtype=torch.float64)
L.groupnorm_stats(x, NB, HW, st)
gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
y = torch.empty(NB * HW, C, device=dev, dtype=torch.bfloat16)
def once():
    L.groupnorm(x, st, None, None, NB, HW, 32, gamma, beta, 1e-5, L.ACT_SILU, y)
for _ in range(3):
    once()
torch.cuda.synchronize()
torch.cuda.profiler.start()
once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
