"""`Tango(name)` on a snapshot directory with the reference's file layout (tango.py:12-36): vae_config.json,
stft_config.json, main_config.json, pytorch_model_{vae,stft,main}.bin — written here with tiny synthetic weights (the real
checkpoint is not reachable offline). CPU: every file is read, validated and routed to the right module; GPU: the loaded
instance generates audio and agrees with an instance built directly from the same tensors."""
import json
import os

import numpy as np
import pytest
import torch

from tango_b200 import synth
from tango_b200.pipeline import Tango
from tango_b200.stft import fourier_basis


def write_snapshot(root, with_text_encoder=False):
    ucfg = dict(synth.TINY_UNET_CONFIG)
    os.makedirs(root, exist_ok=True)
    json.dump(ucfg, open(os.path.join(root, "unet_config.json"), "w"))
    json.dump(dict(synth.VAE_CONFIG), open(os.path.join(root, "vae_config.json"), "w"))
    stft_cfg = dict(synth.STFT_CONFIG, filter_length=256, hop_length=40, win_length=256, n_mel_channels=16)
    json.dump(stft_cfg, open(os.path.join(root, "stft_config.json"), "w"))
    sched_dir = os.path.join(root, "sched")
    os.makedirs(os.path.join(sched_dir, "scheduler"), exist_ok=True)
    json.dump({"_class_name": "DDPMScheduler", "num_train_timesteps": 1000, "beta_start": 0.00085, "beta_end": 0.012,
               "beta_schedule": "scaled_linear", "prediction_type": "v_prediction", "clip_sample": False,
               "steps_offset": 1, "set_alpha_to_one": False}, open(os.path.join(sched_dir, "scheduler", "scheduler_config.json"), "w"))
    json.dump({"text_encoder_name": None, "scheduler_name": sched_dir, "unet_model_name": None,
               "unet_model_config_path": os.path.join(root, "unet_config.json"), "snr_gamma": 5.0},
              open(os.path.join(root, "main_config.json"), "w"))
    usd = synth.synth_state_dict(synth.unet_param_shapes(ucfg), seed=4)
    vsd = synth.synth_state_dict(synth.vae_decoder_param_shapes(), seed=4)
    vsd.update(synth.synth_state_dict(synth.vae_encoder_param_shapes(), seed=4))
    g = torch.Generator().manual_seed(2)
    ssd = {"mel_basis": torch.rand(16, 129, generator=g), "stft_fn.forward_basis": fourier_basis(256, 256),
           "stft_fn.inverse_basis": torch.randn(258, 1, 256, generator=g)}
    torch.save({"unet." + k: v for k, v in usd.items()}, os.path.join(root, "pytorch_model_main.bin"))
    torch.save(vsd, os.path.join(root, "pytorch_model_vae.bin"))
    torch.save(ssd, os.path.join(root, "pytorch_model_stft.bin"))
    return usd, vsd, ssd


def test_snapshot_directory_loads_on_cpu(tmp_path, capsys):
    usd, vsd, ssd = write_snapshot(str(tmp_path / "snap"))
    t = Tango(str(tmp_path / "snap"), device="cpu")
    assert "Successfully loaded checkpoint from:" in capsys.readouterr().out      # tango.py:30
    assert set(t.model.unet._sd) == set(usd) and torch.equal(t.model.unet._sd["conv_in.weight"], usd["conv_in.weight"])
    assert torch.equal(t.stft.mel_basis, ssd["mel_basis"]) and t.stft.mel_basis_source == "checkpoint"
    assert t.stft.filter_length == 256 and t.stft.hop_length == 40
    assert t.vae._esd is not None                                                 # encoder.* / quant_conv.* picked up
    assert t.scheduler.config["prediction_type"] == "v_prediction" and t.scheduler.config["beta_end"] == 0.012
    assert t.model.unet.config["cross_attention_dim"] == synth.TINY_UNET_CONFIG["cross_attention_dim"]
    # the reference's failure modes: a missing file and a wrong tensor shape are errors, not silent fallbacks
    os.remove(tmp_path / "snap" / "pytorch_model_stft.bin")
    with pytest.raises(FileNotFoundError):
        Tango(str(tmp_path / "snap"), device="cpu")
    with pytest.raises(FileNotFoundError):
        Tango("declare-lab/tango", device="cpu")                                   # hub names need a local snapshot


def test_snapshot_rejects_mismatched_weights(tmp_path):
    write_snapshot(str(tmp_path / "snap"))
    bad = torch.load(tmp_path / "snap" / "pytorch_model_main.bin")
    bad["unet.conv_in.weight"] = torch.zeros(3, 3)
    torch.save(bad, tmp_path / "snap" / "pytorch_model_main.bin")
    with pytest.raises(RuntimeError, match="size mismatch"):
        Tango(str(tmp_path / "snap"), device="cpu")


@pytest.mark.gpu
def test_snapshot_directory_generates_audio(cuda, tmp_path):
    usd, vsd, _ = write_snapshot(str(tmp_path / "snap"))
    t = Tango(str(tmp_path / "snap"), device=cuda, precision="split")
    g = torch.Generator(device=cuda).manual_seed(3)
    wave = t.generate("rain on a tin roof", steps=3, guidance=3, latent_shape=(32, 16), generator=g)
    assert wave.dtype == np.int16 and wave.shape == (20512,)
    ref = Tango.from_synthetic(unet_config=synth.TINY_UNET_CONFIG, device=cuda, precision="split")
    ref.model.unet.load_state_dict(usd)
    ref.vae.load_state_dict(vsd)
    g = torch.Generator(device=cuda).manual_seed(3)
    want = ref.generate("rain on a tin roof", steps=3, guidance=3, latent_shape=(32, 16), generator=g)
    assert np.abs(wave.astype(np.int32) - want.astype(np.int32)).max() <= 2
    # tango.stft is usable as inference.py:81 uses it: waveform -> fbank for the VAE encoder
    from tango_b200.stft import wav_to_fbank
    fbank, _, _ = wav_to_fbank([torch.from_numpy(wave.astype(np.float32) / 32768.0).to(cuda)], target_length=512, fn_STFT=t.stft)
    assert fbank.shape == (1, 512, 16) and torch.isfinite(fbank).all()
