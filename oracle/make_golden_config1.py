"""The canonical config-1 correctness artefact (BASELINE.json configs[0]) — TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_config1        (needs /root/reference; run in the build container, ~10 min on 8 cores)

Full-size Tango base UNet (configs/diffusion_model_config.json, 866 M parameters, seeded synthetic weights), ONE prompt
(64 synthetic T5 tokens, the unconditional half masked like T5("")), classifier-free guidance 3.0, TEN denoising steps
at the real latent size 256 x 16, fp32 on the CPU, run through the UNMODIFIED reference:
`AudioDiffusion.inference` (/root/reference/models.py:210-257) with the fork's DDPMScheduler and DDIMScheduler, then
`AutoencoderKL.decode_first_stage` / `decode_to_waveform` (tango.py:46-48) on the DDIM latents. The oracle restatement
runs on the same inputs and is asserted equal to round-off (this pins the oracle at full size), and the reference outputs
are stored in tests/golden/config1.npz (~0.8 MB): final latents of both loops, per-step latent norms, mel, int16
waveform. Inputs are not stored: tango_b200.synth regenerates them from the seeds below.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import hifigan as ohifi  # noqa: E402
from oracle import pipeline as opipe  # noqa: E402
from oracle import refshim  # noqa: E402
from oracle import schedulers as osched  # noqa: E402
from oracle import vae as ovae  # noqa: E402
from tango_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SEEDS = {"weights": 0, "conditioning": 1, "noise": 1234}
STEPS, GUIDANCE, TOKENS, MASKED_TAIL = 10, 3.0, 64, 9


def inputs():
    """The config-1 inputs, regenerated from SEEDS (shared with tests/test_config1_gpu.py)."""
    cfg = dict(synth.BASE_UNET_CONFIG)
    embeds, mask = synth.synth_conditioning(1, TOKENS, cfg["cross_attention_dim"], seed=SEEDS["conditioning"],
                                            masked_tail=MASKED_TAIL)
    lat0, noises = synth.synth_noise(1, STEPS, shape=(8, 256, 16), seed=SEEDS["noise"])
    return cfg, embeds, mask, lat0, noises


def maxdiff(a, b):
    return float((torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs().max())


def main():
    torch.set_grad_enabled(False)
    t00 = time.time()
    cfg, embeds, mask, lat0, noises = inputs()
    sd = synth.synth_state_dict(synth.unet_param_shapes(cfg), seed=SEEDS["weights"])
    U = refshim.unet_class()
    ref_unet = U.from_config(dict(cfg)).eval()
    ref_unet.load_state_dict(sd, strict=True)
    DDPM, DDIM = refshim.schedulers()
    refmod = refshim.audio_diffusion_module()
    import diffusers.schedulers.scheduling_ddpm as ref_ddpm_mod
    sc = dict(osched.SD21_CONFIG)

    class _Stub:
        pass

    out = {}
    checks = {}
    for name in ("ddpm", "ddim"):
        stub = _Stub()
        stub.unet = ref_unet
        stub.set_from = "random"
        stub.text_encoder = _Stub()
        stub.text_encoder.device = torch.device("cpu")
        stub.encode_text_classifier_free = lambda prompt, n: (embeds, mask)
        queue = [lat0] + list(noises)
        stub.prepare_latents = lambda bs, sch, ch, dt, dev: queue.pop(0) * sch.init_noise_sigma
        norms = []
        if name == "ddpm":
            r = DDPM(num_train_timesteps=1000, beta_start=sc["beta_start"], beta_end=sc["beta_end"],
                     beta_schedule=sc["beta_schedule"], prediction_type=sc["prediction_type"], clip_sample=False)
            o = osched.OracleDDPM(**sc)
        else:
            r = DDIM(num_train_timesteps=1000, beta_start=sc["beta_start"], beta_end=sc["beta_end"],
                     beta_schedule=sc["beta_schedule"], prediction_type=sc["prediction_type"], clip_sample=False,
                     set_alpha_to_one=False, steps_offset=1)
            o = osched.OracleDDIM(**sc)
        orig = ref_ddpm_mod.randn_tensor
        ref_ddpm_mod.randn_tensor = lambda *a, **k: queue.pop(0)
        t0 = time.time()
        try:
            lat_ref = refmod.AudioDiffusion.inference(stub, ["synthetic prompt"], r, STEPS, GUIDANCE, 1, True)
        finally:
            ref_ddpm_mod.randn_tensor = orig
        t_ref = time.time() - t0
        trace = []
        t0 = time.time()
        lat_orc = opipe.inference(sd, cfg, o, embeds, mask, STEPS, GUIDANCE, lat0, noises if name == "ddpm" else None,
                                  trace=trace)
        t_orc = time.time() - t0
        d = maxdiff(lat_ref, lat_orc)
        norms = [float(x.norm()) for x in trace]
        print(f"config-1 {name}: {STEPS} steps, |lat| max {lat_ref.abs().max():.3f}, oracle-vs-reference max diff {d:.3e} "
              f"(reference {t_ref:.0f} s, oracle {t_orc:.0f} s)", flush=True)
        assert d < 5e-4, d
        out[f"latents_{name}"] = lat_ref.numpy()
        out[f"timesteps_{name}"] = r.timesteps.numpy()
        out[f"step_norms_{name}"] = np.asarray(norms, dtype=np.float64)
        checks[name] = {"latents_max_abs": d, "reference_s": round(t_ref, 1), "oracle_s": round(t_orc, 1)}
    del ref_unet

    # ---- decode the DDIM latents (the benchmark's sampler) with the reference VAE + HiFi-GAN
    A = refshim.autoencoder_class()
    vae = A(**synth.VAE_CONFIG).eval()
    vsd = synth.synth_state_dict(synth.vae_decoder_param_shapes(), seed=SEEDS["weights"])
    full = vae.state_dict()
    full.update(vsd)
    vae.load_state_dict(full, strict=True)
    lat = torch.from_numpy(out["latents_ddim"])
    mel_ref = vae.decode_first_stage(lat)
    wav_ref_i16 = vae.decode_to_waveform(mel_ref)
    wav_ref_f = vae.vocoder(mel_ref.squeeze(1).permute(0, 2, 1)).squeeze(1)
    mel_orc = ovae.decode_first_stage(vsd, lat, synth.VAE_CONFIG["scale_factor"])
    wav_orc_f, wav_orc_i16 = ohifi.decode_to_waveform(vsd, mel_orc)
    dm, dw = maxdiff(mel_ref, mel_orc), maxdiff(wav_ref_f, wav_orc_f)
    di = int(np.abs(wav_ref_i16.astype(np.int32) - wav_orc_i16.astype(np.int32)).max())
    print(f"config-1 decode: mel {tuple(mel_ref.shape)} diff {dm:.3e}; wave {tuple(wav_ref_f.shape)} diff {dw:.3e}; "
          f"int16 diff {di}", flush=True)
    assert dm < 1e-3 and dw < 1e-3 and di <= 8
    checks["decode"] = {"mel_max_abs": dm, "wave_max_abs": dw, "int16_max": di}
    out["mel"] = mel_ref.numpy().astype(np.float32)
    out["wave_i16"] = wav_ref_i16
    np.savez_compressed(os.path.join(GOLD, "config1.npz"), **out)
    mp = os.path.join(GOLD, "MANIFEST.json")
    manifest = json.load(open(mp)) if os.path.exists(mp) else {"checks": {}}
    manifest["checks"]["config1"] = dict(checks, seeds=SEEDS, steps=STEPS, guidance=GUIDANCE, tokens=TOKENS,
                                         generated=time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
                                         what="BASELINE.json configs[0]: full base UNet, 1 prompt, 10 steps, fp32 CPU, "
                                              "through the unmodified reference loop, VAE and HiFi-GAN")
    with open(mp, "w") as f:
        json.dump(manifest, f, indent=1)
    print(f"tests/golden/config1.npz written ({os.path.getsize(os.path.join(GOLD, 'config1.npz')) / 1e6:.2f} MB) in "
          f"{time.time() - t00:.0f} s")


if __name__ == "__main__":
    main()
