"""Generate the golden vectors under tests/golden/ from the REAL reference and pin the oracle against it.

    python -m oracle.make_golden          (needs /root/reference; run in the build container, not on the GPU box)

For every piece of the hot path the unmodified reference module (imported through oracle/refshim.py) is run on
seeded synthetic weights / inputs (tango_b200/synth.py), the oracle restatement is run on the same data, the two are
asserted equal to fp32 round-off, and the reference output is stored as a small fixture. TEST INFRASTRUCTURE ONLY.
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
from oracle import unet as ounet  # noqa: E402
from oracle import vae as ovae  # noqa: E402
from tango_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def maxdiff(a, b):
    return float((a.double() - b.double()).abs().max())


def main():
    torch.set_grad_enabled(False)
    os.makedirs(GOLD, exist_ok=True)
    manifest = {"torch": torch.__version__, "generated": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
                "reference": "declare-lab/tango @ /root/reference (diffusers fork 0.15.0.dev0)", "checks": {}}

    # ------------------------------------------------------------------ 1. tiny UNet forward
    U = refshim.unet_class()
    cfg = dict(synth.TINY_UNET_CONFIG)
    sd = synth.synth_state_dict(synth.unet_param_shapes(cfg), seed=0)
    ref_unet = U.from_config(dict(cfg)).eval()
    ref_unet.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(7)
    sample = torch.randn(2, 8, 32, 16, generator=g)
    ehs = torch.randn(2, 12, cfg["cross_attention_dim"], generator=g)
    mask = torch.ones(2, 12, dtype=torch.bool)
    mask[0, 1:] = False
    mask[1, 9:] = False
    t = torch.tensor(481)
    out_ref = ref_unet(sample, t, encoder_hidden_states=ehs, encoder_attention_mask=mask).sample
    out_orc = ounet.unet_forward(sd, cfg, sample, t, ehs, mask)
    d = maxdiff(out_ref, out_orc)
    print(f"tiny UNet forward: |ref| max {out_ref.abs().max():.3f}  oracle-vs-reference max diff {d:.3e}")
    assert d < 5e-5
    # no-mask / python-int timestep variant
    out_ref2 = ref_unet(sample, 7, encoder_hidden_states=ehs).sample
    out_orc2 = ounet.unet_forward(sd, cfg, sample, 7, ehs, None)
    d2 = maxdiff(out_ref2, out_orc2)
    assert d2 < 5e-5
    manifest["checks"]["tiny_unet"] = {"oracle_vs_reference_max_abs": d, "nomask": d2}
    np.savez_compressed(os.path.join(GOLD, "tiny_unet.npz"), sample=sample.numpy(), ehs=ehs.numpy(),
                        mask=mask.numpy(), t=np.int64(481), out=out_ref.numpy(), out_nomask_t7=out_ref2.numpy())

    # ------------------------------------------------------------------ 2. schedulers
    DDPM, DDIM = refshim.schedulers()
    sc = dict(osched.SD21_CONFIG)
    sched_gold = {}
    for n in (10, 200):
        r = DDPM(num_train_timesteps=1000, beta_start=sc["beta_start"], beta_end=sc["beta_end"],
                 beta_schedule=sc["beta_schedule"], prediction_type=sc["prediction_type"], clip_sample=False)
        r.set_timesteps(n)
        o = osched.OracleDDPM(**sc)
        o.set_timesteps(n)
        assert torch.equal(r.timesteps, o.timesteps)
        sched_gold[f"ddpm_timesteps_{n}"] = r.timesteps.numpy()
        ri = DDIM(num_train_timesteps=1000, beta_start=sc["beta_start"], beta_end=sc["beta_end"],
                  beta_schedule=sc["beta_schedule"], prediction_type=sc["prediction_type"], clip_sample=False,
                  set_alpha_to_one=False, steps_offset=1)
        ri.set_timesteps(n)
        oi = osched.OracleDDIM(**sc)
        oi.set_timesteps(n)
        assert torch.equal(ri.timesteps, oi.timesteps)
        sched_gold[f"ddim_timesteps_{n}"] = ri.timesteps.numpy()
    # full loops with a deterministic "model" and injected noise; bit-exact oracle == reference
    import diffusers.schedulers.scheduling_ddpm as ref_ddpm_mod
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(2, 8, 16, 16, generator=g)
    noises = [torch.randn(2, 8, 16, 16, generator=g) for _ in range(10)]
    for pred in ("v_prediction", "epsilon"):
        r = DDPM(num_train_timesteps=1000, beta_start=sc["beta_start"], beta_end=sc["beta_end"],
                 beta_schedule=sc["beta_schedule"], prediction_type=pred, clip_sample=False)
        r.set_timesteps(10)
        o = osched.OracleDDPM(**dict(sc, prediction_type=pred))
        o.set_timesteps(10)
        queue = list(noises)
        orig = ref_ddpm_mod.randn_tensor
        ref_ddpm_mod.randn_tensor = lambda *a, **k: queue.pop(0)
        xr, xo = x0.clone(), x0.clone()
        for i, tt in enumerate(r.timesteps):
            mo_r = torch.sin(xr * 3.0 + float(tt) / 1000)
            xr = r.step(mo_r, tt, xr).prev_sample
            mo_o = torch.sin(xo * 3.0 + float(tt) / 1000)
            xo = o.step(mo_o, tt, xo, noises[i])
        ref_ddpm_mod.randn_tensor = orig
        assert torch.equal(xr, xo), f"DDPM {pred} oracle not bit-exact"
        sched_gold[f"ddpm_loop_{pred}"] = xr.numpy()
        ri = DDIM(num_train_timesteps=1000, beta_start=sc["beta_start"], beta_end=sc["beta_end"],
                  beta_schedule=sc["beta_schedule"], prediction_type=pred, clip_sample=False, set_alpha_to_one=False,
                  steps_offset=1)
        ri.set_timesteps(10)
        oi = osched.OracleDDIM(**dict(sc, prediction_type=pred))
        oi.set_timesteps(10)
        xr, xo = x0.clone(), x0.clone()
        for tt in ri.timesteps:
            xr = ri.step(torch.sin(xr * 3.0 + float(tt) / 1000), tt, xr).prev_sample
            xo = oi.step(torch.sin(xo * 3.0 + float(tt) / 1000), tt, xo)
        assert torch.equal(xr, xo), f"DDIM {pred} oracle not bit-exact"
        sched_gold[f"ddim_loop_{pred}"] = xr.numpy()
    sched_gold["x0"] = x0.numpy()
    sched_gold["noises"] = torch.stack(noises).numpy()
    np.savez_compressed(os.path.join(GOLD, "schedulers.npz"), **sched_gold)
    manifest["checks"]["schedulers"] = "oracle == reference bit-exact (timesteps, DDPM/DDIM 10-step loops, v/eps)"
    print("schedulers: oracle bit-exact vs reference")

    # ------------------------------------------------------------------ 3. VAE decoder + HiFi-GAN
    A = refshim.autoencoder_class()
    vae = A(**synth.VAE_CONFIG).eval()
    vsd = synth.synth_state_dict(synth.vae_decoder_param_shapes(), seed=0)
    full = vae.state_dict()
    full.update(vsd)
    vae.load_state_dict(full, strict=True)
    g = torch.Generator().manual_seed(11)
    z = torch.randn(1, 8, 8, 16, generator=g)
    mel_ref = vae.decode_first_stage(z)
    wav_ref_i16 = vae.decode_to_waveform(mel_ref)
    wav_ref_f = vae.vocoder(mel_ref.squeeze(1).permute(0, 2, 1)).squeeze(1)
    mel_orc = ovae.decode_first_stage(vsd, z, synth.VAE_CONFIG["scale_factor"])
    wav_orc_f, wav_orc_i16 = ohifi.decode_to_waveform(vsd, mel_orc)
    dm, dw = maxdiff(mel_ref, mel_orc), maxdiff(wav_ref_f, wav_orc_f)
    di = int(np.abs(wav_ref_i16.astype(np.int32) - wav_orc_i16.astype(np.int32)).max())
    print(f"VAE mel {tuple(mel_ref.shape)} |max| {mel_ref.abs().max():.3f} diff {dm:.3e}; wave {wav_ref_f.shape} "
          f"|max| {wav_ref_f.abs().max():.3f} diff {dw:.3e}; int16 diff {di}")
    assert dm < 1e-4 and dw < 1e-4 and di <= 2
    manifest["checks"]["vae_vocoder"] = {"mel_max_abs": dm, "wave_max_abs": dw, "int16_max": di}
    np.savez_compressed(os.path.join(GOLD, "tiny_vae_vocoder.npz"), z=z.numpy(), mel=mel_ref.numpy(),
                        wave=wav_ref_f.numpy(), wave_i16=wav_ref_i16)

    # ------------------------------------------------------------------ 4. AudioDiffusion.inference (tiny UNet)
    refmod = refshim.audio_diffusion_module()
    B, steps, guidance = 1, 4, 3.0
    embeds, bmask = synth.synth_conditioning(B, 10, cfg["cross_attention_dim"], seed=5, masked_tail=3)
    lat0, noises = synth.synth_noise(B, steps, shape=(8, 32, 16), seed=99)

    class _Stub:
        pass

    stub = _Stub()
    stub.unet = ref_unet
    stub.set_from = "random"
    stub.text_encoder = _Stub()
    stub.text_encoder.device = torch.device("cpu")
    stub.encode_text_classifier_free = lambda prompt, n: (embeds, bmask)
    queue = [lat0] + list(noises)
    stub.prepare_latents = lambda bs, sch, ch, dt, dev: queue.pop(0) * sch.init_noise_sigma
    orig = ref_ddpm_mod.randn_tensor
    ref_ddpm_mod.randn_tensor = lambda *a, **k: queue.pop(0)
    r = DDPM(num_train_timesteps=1000, beta_start=sc["beta_start"], beta_end=sc["beta_end"],
             beta_schedule=sc["beta_schedule"], prediction_type=sc["prediction_type"], clip_sample=False)
    lat_ref = refmod.AudioDiffusion.inference(stub, ["synthetic prompt"], r, steps, guidance, 1, True)
    ref_ddpm_mod.randn_tensor = orig
    o = osched.OracleDDPM(**sc)
    lat_orc = opipe.inference(sd, cfg, o, embeds, bmask, steps, guidance, lat0, noises)
    dl = maxdiff(lat_ref, lat_orc)
    print(f"AudioDiffusion.inference (tiny, {steps} DDPM steps, CFG {guidance}): |lat| max {lat_ref.abs().max():.3f} diff {dl:.3e}")
    assert dl < 2e-4
    manifest["checks"]["tiny_inference"] = {"latents_max_abs": dl, "steps": steps, "guidance": guidance}
    np.savez_compressed(os.path.join(GOLD, "tiny_inference.npz"), latents=lat_ref.numpy(), embeds=embeds.numpy(),
                        mask=bmask.numpy(), lat0=lat0.numpy(), noises=torch.stack(noises).numpy())

    with open(os.path.join(GOLD, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
