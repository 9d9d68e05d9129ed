#!/usr/bin/env python
"""bench.py — audio-seconds generated per wall-second on the Tango hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic prompts: `denoise_steps` (200) CFG denoising steps
of the Tango base UNet on a batch of 8 prompts (UNet batch 16, 64 synthetic T5 tokens), then the VAE decoder and the
HiFi-GAN vocoder -> 8 x 163 872 int16 samples (10.24 s each). Workload = BASELINE.json configs[1].

Printed JSON (rank 0, one line):
  value      audio-s/s, inputs resident in HBM, device-timed (CUDA events), max over ranks, whole job
  e2e        same metric through the public API (Tango.generate_for_batch) with HOST buffers: pinned-host prompt
             embeddings copied H2D and the int16 waveforms copied D2H inside the timed region
  roofline   dominant kernel (tcgen05 implicit-GEMM conv/linear): algorithmic FLOPs / CUDA-event time per launch
             against the measured bf16 peak in MEASURED_PEAKS.json
  cpu_baseline  the oracle port timed on this box's host cores on a bounded sample (N=1 only)
`--impl reference` times the reference's CPU arithmetic (oracle port; the Python reference itself cannot travel to the
GPU box) on the same config / metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

AUDIO_S_PER_SAMPLE = 163872 / 16000.0  # 10.242 s (hifigan: 1024 mel frames -> 163 872 samples)
F_UNET, F_VAE, F_VOC = 803.181e9, 670.468e9, 1027.036e9  # SURVEY.md §8d, per sample, FLOP = 2 MAC


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d.get("bf16_tflops", 1590.0), "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_inputs(B: int, tokens: int, dim: int, rank: int):
    from tango_b200 import synth
    embeds, mask = synth.synth_conditioning(B, tokens, dim, seed=1 + rank)
    return embeds, mask


# ------------------------------------------------------------------------------------------------- reference arm
def pick_threads():
    """Host threads for the CPU arm: the fastest of a few counts on a representative 3x3 conv (oversubscribing a
    128-core box makes torch's CPU kernels several times slower, so 'all cores' is not 'all the threads it can use')."""
    import torch.nn.functional as F
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    x = torch.randn(2, 320, 256, 16)
    w = torch.randn(320, 320, 3, 3)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        F.conv2d(x, w, padding=1)
        t0 = time.perf_counter()
        for _ in range(3):
            F.conv2d(x, w, padding=1)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


class CpuReference:
    """The reference's CPU arithmetic for the path (oracle port, see oracle/__init__.py) on bounded samples."""

    def __init__(self, args):
        from oracle import hifigan as ohifi
        from oracle import unet as ounet
        from oracle import vae as ovae
        from tango_b200 import synth
        torch.set_grad_enabled(False)
        self.args, self.ounet, self.ovae, self.ohifi, self.synth = args, ounet, ovae, ohifi, synth
        self.cores = pick_threads()
        self.cfg = synth.BASE_UNET_CONFIG
        self.usd = synth.synth_state_dict(synth.unet_param_shapes(self.cfg), 0)
        self.vsd = synth.synth_state_dict(synth.vae_decoder_param_shapes(), 0)
        self.embeds, self.mask = synthetic_inputs(1, args.tokens, self.cfg["cross_attention_dim"], 0)
        g = torch.Generator().manual_seed(1234)
        self.x_full = torch.randn(2, 8, 256, 16, generator=g)
        self.x_small = self.x_full[:, :, :64].contiguous()
        # calibration on the quarter-length clip, FLOPs counted live
        from torch.utils.flop_counter import FlopCounterMode
        with FlopCounterMode(display=False) as fc:
            t0 = time.perf_counter()
            self.fwd(self.x_small, 0)
            self.t_small = time.perf_counter() - t0
        self.f_small = float(fc.get_total_flops())
        self.f_full = 2 * F_UNET          # CFG batch 2
        self.use_full = self.t_small * self.f_full / self.f_small <= 25.0

    def fwd(self, x, i):
        return self.ounet.unet_forward(self.usd, self.cfg, x, torch.tensor(995 - 5 * (i % 199)), self.embeds, self.mask)

    def step_seconds(self, i):
        """Seconds of one CFG UNet forward for one prompt at 256x16 (measured, or FLOP-scaled from the 64x16 sample)."""
        x = self.x_full if self.use_full else self.x_small
        t0 = time.perf_counter()
        self.fwd(x, i)
        dt = time.perf_counter() - t0
        return dt if self.use_full else dt * self.f_full / self.f_small

    def decode_seconds(self):
        """VAE decoder + HiFi-GAN for one sample; a quarter-length latent scaled by 4 when the box is slow."""
        z = self.x_full[:1] if self.use_full else self.x_small[:1]
        t0 = time.perf_counter()
        mel = self.ovae.decode_first_stage(self.vsd, z, self.synth.VAE_CONFIG["scale_factor"])
        self.ohifi.decode_to_waveform(self.vsd, mel)
        dt = time.perf_counter() - t0
        return dt if self.use_full else dt * 4.0

    def describe(self, t_fwd, t_dec):
        a = self.args
        what = ("1 UNet forward at CFG batch 2 (1 prompt, 64 tokens, 256x16 latent)" if self.use_full else
                f"1 UNet forward at CFG batch 2 on a quarter-length latent (64x16, {self.f_small / 1e9:.0f} GFLOP counted live) "
                f"scaled by the FLOP ratio to the 256x16 forward ({self.f_full / 1e9:.0f} GFLOP)")
        return (f"oracle port, torch CPU fp32, {self.cores} threads: per step {what} = {t_fwd:.2f} s; VAE+HiFi-GAN for 1 sample "
                f"= {t_dec:.2f} s (timed once); extrapolated linearly to {a.denoise_steps} denoising steps and batch {a.batch}")


def run_reference(args):
    """`--impl reference`: the reference's CPU arithmetic for this path on the box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference(args)
    t_dec = ref.decode_seconds()
    times = []
    for i in range(args.warmup + args.steps):
        dt = ref.step_seconds(i)
        if i >= args.warmup:
            times.append(dt)
    t_fwd = float(np.mean(times))
    per_sample = args.denoise_steps * t_fwd + t_dec       # one prompt = one CFG forward per denoising step
    value = AUDIO_S_PER_SAMPLE / per_sample
    line = {"impl": "reference", "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_sample * args.batch * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": ref.cores, "kind": "port",
                             "sample": ref.describe(t_fwd, t_dec)},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": "Tango " + args.unet + " UNet, batch %d prompts/GPU x %d GPU, %d %s steps, "
                        "CFG %.1f, %.2f s clips, 64 synthetic T5 tokens, + VAE decoder + HiFi-GAN -> int16 16 kHz"
                        % (args.batch, world, args.denoise_steps, args.scheduler.upper(), args.guidance,
                           (4 * args.latent_h * 160 + 32) / 16000.0),
            "global_batch": args.batch * world, "unet_batch_per_gpu": 2 * args.batch, "denoise_steps": args.denoise_steps,
            "scheduler": args.scheduler, "guidance": args.guidance, "precision": args.precision,
            "parallelism": f"prompt-shard x{world}", "l2": "working set (1.7 GB bf16 weights + activations) >> 126 MB L2"}


# ------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    from tango_b200 import lib as L
    from tango_b200 import parallel, synth
    from tango_b200.pipeline import Tango
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)

    cfg = synth.BASE_UNET_CONFIG if args.unet == "base" else synth.XL_UNET_CONFIG
    B = args.batch
    latent_shape = (args.latent_h, 16)
    audio_s = (4 * args.latent_h * 160 + 32) / 16000.0   # HiFi-GAN: 160 samples per mel frame (+32 tail)
    t = Tango.from_synthetic(unet_config=cfg, device=dev, precision=args.precision, scheduler=args.scheduler)
    if world > 1:
        # one-time NCCL broadcast of the (rank-0) weights over NVLink, as a sharded deployment would do at load
        usd = parallel.broadcast_state_dict({k: v.to(dev) for k, v in t.model.unet._sd.items()}, src=0)
        t.model.unet.load_state_dict(usd)
    embeds_h, mask_h = synthetic_inputs(B, args.tokens, cfg["cross_attention_dim"], rank)
    embeds_pin, mask_pin = embeds_h.pin_memory(), mask_h.pin_memory()
    embeds_d, mask_d = embeds_h.to(dev), mask_h.to(dev)
    prompts = [f"synthetic prompt {rank}-{i}" for i in range(B)]
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    def one_pass_device():
        lat = t.model.inference(prompts, t.scheduler, args.denoise_steps, args.guidance, prompt_embeds=embeds_d,
                                boolean_prompt_mask=mask_d, generator=gen, latent_shape=latent_shape)
        B_, Cl, H, W = lat.shape
        rows = lat.permute(0, 2, 3, 1).reshape(B_ * H * W, Cl).contiguous()
        mel = t.vae.decode_rows(rows, B_, H, W)
        return t.vae.vocoder_rows(mel.view(B_ * 4 * H, 4 * W), B_, 4 * H)

    def one_pass_e2e():
        return t.generate_for_batch(prompts, steps=args.denoise_steps, guidance=args.guidance, batch_size=B,
                                    prompt_embeds=embeds_pin, boolean_prompt_mask=mask_pin, generator=gen,
                                    latent_shape=latent_shape)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- warm-up
    for _ in range(max(args.warmup, 1)):
        one_pass_device()
    barrier()
    # ---------------- timed: device-resident
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = L.launch_count()
    graph_launches = 0
    unet_ms = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        one_pass_device()
        graph_launches += t.model.launches_per_forward * args.denoise_steps
        unet_ms.append(t.model.last_step_ms)
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    launches = (L.launch_count() - n0) + graph_launches
    dev_ms = parallel.max_over_ranks(dev_ms, dev)
    value = world * B * audio_s * args.steps / (dev_ms / 1e3)

    # ---------------- timed: end to end through the public API with host buffers
    one_pass_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        waves = one_pass_e2e()
    barrier()
    e2e_s = parallel.max_over_ranks(time.perf_counter() - t0, dev)
    e2e_value = world * B * audio_s * args.steps / e2e_s
    h2d = embeds_pin.numel() * 4 + mask_pin.numel()
    d2h = sum(int(w.nbytes) for w in waves)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel: one instrumented (eager, CUDA-event-per-launch) UNet forward
    pk = peaks()
    m = t.model
    m.use_cuda_graph = False
    m.inference(prompts, t.scheduler, 1, args.guidance, prompt_embeds=embeds_d, boolean_prompt_mask=mask_d, generator=gen,
                latent_shape=latent_shape)
    torch.cuda.synchronize()
    # Park the GPU behind a ~0.25 s spin kernel so that the (slow, Python-driven) eager launches queue up ahead of the
    # GPU: the per-launch CUDA events then bracket pure kernel execution, not host launch latency.
    torch.cuda._sleep(int(0.25 * 1.9e9))
    L.PROF.start()
    m.inference(prompts, t.scheduler, 2, args.guidance, prompt_embeds=embeds_d, boolean_prompt_mask=mask_d, generator=gen,
                latent_shape=latent_shape)
    prof = L.PROF.stop()
    m.use_cuda_graph = True
    gm = prof.get("gemm_tc", {"launches": 1, "ms": 1.0, "flops": 0.0})
    at = prof.get("attention_tc", {"launches": 1, "ms": 1.0, "flops": 0.0})
    achieved = gm["flops"] / (gm["ms"] / 1e3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 implicit-GEMM conv/linear)",
            "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
            "frac": achieved / pk["bf16_tflops_sustained"], "traffic": traffic,
            "peak_source": pk["source"] + ", sustained bf16 figure (kernel timed inside a long step)",
            "launches_profiled": gm["launches"], "avg_launch_ms": gm["ms"] / max(1, gm["launches"]),
            "algorithmic_gflop_per_launch": gm["flops"] / max(1, gm["launches"]) / 1e9,
            "attention_tc": {"achieved": at["flops"] / (at["ms"] / 1e3) / 1e12, "launches": at["launches"],
                             "avg_launch_ms": at["ms"] / max(1, at["launches"])}}
    if args.latent_h == 256:
        f_unet, f_vae, f_voc = (F_UNET if args.unet == "base" else 806.453e9), F_VAE, F_VOC
    else:  # SURVEY.md §8d figures for the 30 s extension (768 x 16); other lengths are not tabulated
        f_unet, f_vae, f_voc = (3137.856e9 if args.unet == "base" else 3141.13e9), 2217.564e9, 3080.714e9
    f_total = B * (2 * args.denoise_steps * f_unet + f_vae + f_voc)
    whole = f_total * args.steps / (dev_ms / 1e3) / 1e12
    roof["whole_path_tflops"] = whole
    roof["whole_path_frac"] = whole / pk["bf16_tflops_sustained"]

    # ---------------- text-conditioning front-end: reported beside the metric, not inside it
    # (BASELINE.json's metric excludes text encoding). FLAN-T5 encoder of the UNet's width, seeded random weights,
    # `batch` prompts x `tokens` tokens (the "" prompt is cached by the pipeline and not re-encoded).
    text = None
    if world == 1 and not args.no_text_encoder:          # like the CPU baseline: reported at N = 1 only
        text = text_encoder_leg(args, dev)

    # ---------------- CPU baseline (oracle port) on a bounded sample, N = 1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args)

    line = {"metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (split)",
            "data": "synthetic (seeded random weights of the Tango base architecture, random 64-token conditioning)",
            "config": workload_config(args, world), "unet_step_ms": float(np.mean(unet_ms)),
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof}
    if text is not None:
        line["text_encoder"] = text
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def text_encoder_leg(args, dev):
    from tango_b200 import synth
    from tango_b200.t5 import T5EncoderModel
    cfg = synth.FLAN_T5_LARGE_CONFIG if args.unet == "base" else synth.FLAN_T5_XL_CONFIG
    m = T5EncoderModel.from_config(cfg, precision=args.precision).to(dev)
    m.load_state_dict(synth.synth_state_dict(synth.t5_encoder_param_shapes(cfg), 0))
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(2, cfg["vocab_size"], (args.batch, args.tokens), generator=g).to(dev)
    mask = torch.ones(args.batch, args.tokens, dtype=torch.long, device=dev)
    for _ in range(2):
        m(ids, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        m(ids, mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rows, d, ff, inner = args.batch * args.tokens, cfg["d_model"], cfg["d_ff"], cfg["num_heads"] * cfg["d_kv"]
    flops = cfg["num_layers"] * (2 * rows * (4 * d * inner + 3 * d * ff) + 4 * args.batch * cfg["num_heads"] * args.tokens ** 2 * 64)
    del m
    torch.cuda.empty_cache()
    return {"model": ("flan-t5-large" if args.unet == "base" else "flan-t5-xl") + " encoder, seeded random weights",
            "prompts": args.batch, "tokens": args.tokens, "ms": ms, "tflops": flops / (ms / 1e3) / 1e12,
            "note": "eager launches (not graph-captured); outside the timed metric"}


def cpu_baseline(args):
    ref = CpuReference(args)
    t_fwd = ref.step_seconds(0)
    t_dec = ref.decode_seconds()
    v = AUDIO_S_PER_SAMPLE / (args.denoise_steps * t_fwd + t_dec)
    return {"value": v, "unit": "audio-s/s", "cores": ref.cores, "kind": "port", "sample": ref.describe(t_fwd, t_dec)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="prompts per GPU")
    ap.add_argument("--denoise-steps", type=int, default=200)
    ap.add_argument("--guidance", type=float, default=3.0)
    ap.add_argument("--tokens", type=int, default=64)
    ap.add_argument("--scheduler", default="ddim", choices=["ddim", "ddpm"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "split"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-text-encoder", action="store_true", help="skip the FLAN-T5 front-end timing leg")
    ap.add_argument("--latent-h", type=int, default=256, help="latent time frames: 256 = 10.24 s (reference), 768 = 30.7 s")
    ap.add_argument("--unet", default="base", choices=["base", "xl"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
