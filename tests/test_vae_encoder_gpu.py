"""GPU: AutoencoderKL.encode_first_stage against the reference posterior
(tests/golden/tiny_vae_encoder.npz); validated on hardware at the end of round 1."""
import os

import numpy as np
import pytest
import torch

from tango_b200 import synth
from tango_b200.vae import AutoencoderKL

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("precision,tol", [("split", 1e-3), ("bf16", 3e-2)])
def test_vae_encoder_vs_reference_golden(cuda, precision, tol):
    gd = np.load(os.path.join(GOLD, "tiny_vae_encoder.npz"))
    vae = AutoencoderKL(**synth.VAE_CONFIG, precision=precision).to(cuda)
    sd = synth.synth_state_dict(synth.vae_decoder_param_shapes(), seed=0)
    sd.update(synth.synth_state_dict(synth.vae_encoder_param_shapes(), seed=0))
    vae.load_state_dict(sd)
    post = vae.encode_first_stage(torch.from_numpy(gd["mel"]).to(cuda))
    assert post.mean.shape == (2, 8, 16, 16) and post.mean.is_cuda
    e_mean, e_std = rel(post.mean, gd["mean"]), rel(post.std, gd["std"])
    print(f"VAE encoder {precision}: mean rel {e_mean:.3e}, std rel {e_std:.3e}")
    assert e_mean < tol and e_std < tol
    # encode -> decode round trip runs end to end (shapes only: random weights are not an autoencoder)
    mel = vae.decode_first_stage(post.mode() * vae.scale_factor)
    assert mel.shape == (2, 1, 64, 64) and torch.isfinite(mel).all()
