"""Graph-replay time of one UNet forward (UNet batch 16, cfg_shared) — for A/B experiments via environment variables."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tango_b200 import synth
from tango_b200.unet import UNet2DConditionModel
dev = torch.device("cuda:0")
cfg = synth.BASE_UNET_CONFIG
u = UNet2DConditionModel.from_config(cfg, precision="bf16").to(dev)
u.load_state_dict(synth.synth_state_dict(synth.unet_param_shapes(cfg), 0))
B = 8; Bu = 16
emb, mask = synth.synth_conditioning(B, 64, 1024)
u.set_conditioning(emb.to(dev), mask.to(dev))
temb = u.time_embedding_table(torch.full((Bu,), 500.0))
x = torch.randn(Bu * 4096, 8, device=dev).to(torch.bfloat16)
u.forward_rows(x, Bu, 256, 16, temb, temb.shape[1], cfg_shared=True); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    u.forward_rows(x, Bu, 256, 16, temb, temb.shape[1], cfg_shared=True)
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): g.replay()
e1.record(); torch.cuda.synchronize()
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("TNG_")}, "forward_ms": e0.elapsed_time(e1) / 20}))
