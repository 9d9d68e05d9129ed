"""Per-launch CUDA-event timing of one UNet forward (eager) — which layers are slow, at what TFLOP/s.

    python tools/profile_layers.py [--batch 8] [--out profiles/layers.txt]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tango_b200 import lib as L, synth  # noqa: E402
from tango_b200.unet import UNet2DConditionModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--out", default=None)
    ap.add_argument("--once", action="store_true", help="single forward, no event instrumentation (for ncu)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = synth.BASE_UNET_CONFIG
    u = UNet2DConditionModel.from_config(cfg, precision="bf16").to(dev)
    u.load_state_dict(synth.synth_state_dict(synth.unet_param_shapes(cfg), 0))
    Bu = 2 * args.batch
    emb, mask = synth.synth_conditioning(args.batch, 64, 1024)
    u.set_conditioning(emb.to(dev), mask.to(dev))
    temb = u.time_embedding_table(torch.full((Bu,), 500.0))
    x = torch.randn(Bu * 4096, 8, device=dev).to(torch.bfloat16)
    u.forward_rows(x, Bu, 256, 16, temb, temb.shape[1])
    torch.cuda.synchronize()
    if args.once:
        u.forward_rows(x, Bu, 256, 16, temb, temb.shape[1])
        torch.cuda.synchronize()
        return
    recs = []
    orig_gemm, orig_attn = L.conv_gemm, L.attention

    def timed(name, flops_fn, fn):
        def w(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            recs.append((name, flops_fn(*a, **k), e0, e1))
            return r
        return w

    def gemm_desc(views, groups, weight, W, H, NB, **k):
        kk = k.get("algo_k") or sum(g[5] for g in groups) * 64
        return (f"gemm M={W*H*NB:6d} N={weight.shape[0]:5d} K={kk:6d} grp={len(groups):2d} WxH={W}x{H}", 2.0 * W * H * NB * weight.shape[0] * kk)

    def attn_desc(q, k, v, out, **kw):
        return (f"attn B={kw['batch']} h={kw['heads']} Lq={kw['Lq']} Lk={kw['Lk']}", 4.0 * kw["batch"] * kw["heads"] * kw["Lq"] * kw["Lk"] * 64)

    import tango_b200.ops as ops
    L.conv_gemm = timed("gemm", gemm_desc, orig_gemm)
    ops.L.conv_gemm = L.conv_gemm
    L.attention = timed("attn", attn_desc, orig_attn)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    u.forward_rows(x, Bu, 256, 16, temb, temb.shape[1])
    t1.record()
    torch.cuda.synchronize()
    lines = []
    agg = {}
    tot = {"gemm": 0.0, "attn": 0.0}
    for name, (desc, fl), e0, e1 in recs:
        ms = e0.elapsed_time(e1)
        a = agg.setdefault(desc, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += ms
        a[2] += fl
        tot[name] += ms
    lines.append(f"one UNet forward, UNet batch {Bu} (eager, per-launch events): total {t0.elapsed_time(t1):.2f} ms; "
                 f"gemm {tot['gemm']:.2f} ms, attention {tot['attn']:.2f} ms, other {t0.elapsed_time(t1)-tot['gemm']-tot['attn']:.2f} ms")
    lines.append(f"{'launch':64s} {'n':>3s} {'ms':>8s} {'TF/s':>8s}")
    for desc, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{desc:64s} {n:3d} {ms:8.3f} {fl / ms / 1e9:8.1f}")
    txt = "\n".join(lines)
    print(txt)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        open(args.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
