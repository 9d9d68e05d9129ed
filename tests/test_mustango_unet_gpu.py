"""GPU: the Mustango UNet variant against the fork's UNet2DConditionModelMusic output
(tests/golden/tiny_unet_music.npz); validated on hardware at the end of round 1."""
import os

import numpy as np
import pytest
import torch

from tango_b200 import synth
from tango_b200.unet import UNet2DConditionModel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("precision,tol", [("split", 1e-3), ("bf16", 3e-2)])
def test_mustango_unet_forward_vs_reference_golden(cuda, precision, tol):
    gd = np.load(os.path.join(GOLD, "tiny_unet_music.npz"))
    cfg = synth.TINY_MUSIC_UNET_CONFIG
    u = UNet2DConditionModel.from_config(cfg, precision=precision).to(cuda)
    u.load_state_dict(synth.synth_state_dict(synth.unet_param_shapes(cfg), seed=0))
    tt = lambda n: torch.from_numpy(gd[n]).to(cuda)
    out = u(tt("sample"), torch.tensor(int(gd["t"])), tt("ehs"), encoder_attention_mask=tt("mask"),
            beat_features=tt("beat"), chord_features=tt("chord"), beat_attention_mask=tt("bmask"),
            chord_attention_mask=tt("cmask")).sample
    assert out.shape == (2, 8, 32, 16)
    e = rel(out, gd["out"])
    print(f"Mustango tiny UNet {precision}: rel err vs reference golden {e:.3e}")
    assert e < tol
