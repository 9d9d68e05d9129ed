"""One eager UNet forward (UNet batch 16 = 8 prompts under CFG, shared prefix) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum,... --clock-control none` (the launch list in profiles/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tango_b200 import synth
from tango_b200.unet import UNet2DConditionModel
dev = torch.device("cuda:0")
cfg = synth.BASE_UNET_CONFIG
u = UNet2DConditionModel.from_config(cfg, precision="bf16").to(dev)
u.load_state_dict(synth.synth_state_dict(synth.unet_param_shapes(cfg), 0))
B, Bu = 8, 16
emb, mask = synth.synth_conditioning(B, 64, 1024)
u.set_conditioning(emb.to(dev), mask.to(dev))
temb = u.time_embedding_table(torch.full((Bu,), 500.0))
x = torch.randn(Bu * 4096, 8, device=dev).to(torch.bfloat16)
for _ in range(2):
    u.forward_rows(x, Bu, 256, 16, temb, temb.shape[1], cfg_shared=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
u.forward_rows(x, Bu, 256, 16, temb, temb.shape[1], cfg_shared=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
