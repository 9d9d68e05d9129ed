"""Pin oracle/vae.py:encode_first_stage against the UNMODIFIED reference encoder and write tests/golden/tiny_vae_encoder.npz
(TEST INFRASTRUCTURE ONLY; build container only — imports /root/reference through oracle/refshim.py).

    python -m oracle.make_golden_vae_encoder
"""
import json
import os

import numpy as np
import torch

from oracle import refshim
from oracle import vae as ovae
from tango_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    torch.set_grad_enabled(False)
    refshim.install()
    A = refshim.autoencoder_class()
    vae = A(**synth.VAE_CONFIG).eval()
    esd = synth.synth_state_dict(synth.vae_encoder_param_shapes(), seed=0)
    full = vae.state_dict()
    missing = [k for k in esd if k not in full]
    assert not missing, missing
    for k, v in esd.items():
        assert tuple(full[k].shape) == tuple(v.shape), (k, full[k].shape, v.shape)
        full[k] = v
    vae.load_state_dict(full, strict=True)
    g = torch.Generator().manual_seed(4)
    mel = torch.randn(2, 1, 64, 64, generator=g) * 2.0 - 4.0          # log-mel-like range, 64 frames x 64 bins
    post = vae.encode_first_stage(mel)
    mean, std = ovae.encode_first_stage(esd, mel)
    dm, ds = float((post.mean - mean).abs().max()), float((post.std - std).abs().max())
    print(f"VAE encoder: |mean| max {float(post.mean.abs().max()):.3f}; oracle vs reference mean {dm:.3e} std {ds:.3e}")
    assert dm < 1e-5 and ds < 1e-5
    np.savez_compressed(os.path.join(GOLD, "tiny_vae_encoder.npz"), mel=mel.numpy(), mean=post.mean.numpy(),
                        std=post.std.numpy())
    mp = os.path.join(GOLD, "MANIFEST.json")
    man = json.load(open(mp))
    man["checks"]["vae_encoder"] = {"mean_max_abs": dm, "std_max_abs": ds}
    json.dump(man, open(mp, "w"), indent=1)
    print("wrote", os.path.join(GOLD, "tiny_vae_encoder.npz"))


if __name__ == "__main__":
    main()
