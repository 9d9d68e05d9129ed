"""In-situ cost of each kernel family: time the CUDA-graph replay of one UNet forward with that family's launches
skipped (TNG_SKIP=...). Results are garbage with skips on; only the timing matters."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
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
    u.forward_rows(x, Bu, 256, 16, temb, temb.shape[1]); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        u.forward_rows(x, Bu, 256, 16, temb, temb.shape[1])
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"skip": os.environ.get("TNG_SKIP", ""), "ms": e0.elapsed_time(e1) / 10}))
else:
    base = None
    for skip in ["", "gn", "ln", "attn", "gemm", "gn,ln,attn", "gn,ln,attn,gemm"]:
        env = dict(os.environ, TNG_SKIP=skip)
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout.strip().splitlines()
        d = json.loads(out[-1])
        if base is None: base = d["ms"]
        print(f"skip={skip or '-':16s} forward {d['ms']:7.2f} ms   (family cost ~ {base - d['ms']:6.2f} ms)")
