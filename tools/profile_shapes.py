"""Graph-timed cost of every distinct GEMM / attention launch of one UNet forward (no host gaps): records each
launch's arguments during an eager forward, then replays each distinct one 10x inside a CUDA graph."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tango_b200 import lib as L, synth
import tango_b200.ops as ops
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
shared = "--shared" in sys.argv
u.forward_rows(x, Bu, 256, 16, temb, temb.shape[1], cfg_shared=shared); torch.cuda.synchronize()
calls = []
og, oa = L.conv_gemm, L.attention
def rec_g(views, groups, weight, W, H, NB, **k):
    kk = k.get("algo_k") or sum(g[5] for g in groups) * 64
    key = ("gemm", W * H * NB, weight.shape[0], kk, len(groups), k.get("res") is not None, k.get("out_f32") is not None, k.get("out_bf16") is not None, k.get("act", 0))
    calls.append((key, 2.0 * W * H * NB * weight.shape[0] * kk, lambda: og(views, groups, weight, W, H, NB, **k)))
    return og(views, groups, weight, W, H, NB, **k)
def rec_a(q, k_, v, out, **k):
    key = ("attn", k["batch"], k["heads"], k["Lq"], k["Lk"])
    calls.append((key, 4.0 * k["batch"] * k["heads"] * k["Lq"] * k["Lk"] * 64, lambda: oa(q, k_, v, out, **k)))
    return oa(q, k_, v, out, **k)
L.conv_gemm = rec_g; ops.L.conv_gemm = rec_g; L.attention = rec_a
u.forward_rows(x, Bu, 256, 16, temb, temb.shape[1], cfg_shared=shared); torch.cuda.synchronize()
L.conv_gemm = og; ops.L.conv_gemm = og; L.attention = oa
agg = collections.OrderedDict()
for key, fl, fn in calls:
    a = agg.setdefault(key, [0, fl, fn]); a[0] += 1
rows = []
for key, (n, fl, fn) in agg.items():
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    rows.append((n * us, n, us, fl / us / 1e6, key))
tot = sum(r[0] for r in rows)
lines = [f"distinct launches of one UNet forward (UNet batch {Bu}, shared_prefix={shared}); sum = {tot/1e3:.2f} ms",
         f"{'total us':>9s} {'n':>3s} {'us':>8s} {'TF/s':>7s}  key (kind, M, N, K, groups, res, f32, bf16, act) / (kind, B, heads, Lq, Lk)"]
for r in sorted(rows, key=lambda r: -r[0]):
    lines.append(f"{r[0]:9.1f} {r[1]:3d} {r[2]:8.1f} {r[3]:7.1f}  {r[4]}")
txt = "\n".join(lines)
print(txt)
out = [a for a in sys.argv[1:] if a.startswith("--out=")]
if out:
    open(out[0][6:], "w").write(txt + "\n")
