"""Pin oracle/unet.py:unet_forward(extra_streams=...) against the UNMODIFIED Mustango UNet of the diffusers fork
(mustango/diffusers/src/diffusers/models/unet_2d_condition_music.py: UNet2DConditionModelMusic) and write
tests/golden/tiny_unet_music.npz (TEST INFRASTRUCTURE ONLY; build container only). SURVEY.md section 8(f).4.

    python -m oracle.make_golden_music
"""
import json
import os

import numpy as np
import torch

from oracle import refshim
from oracle import unet as ounet
from tango_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TINY_MUSIC_CONFIG = dict(
    synth.TINY_UNET_CONFIG,
    down_block_types=["CrossAttnDownBlock2DMusic", "CrossAttnDownBlock2DMusic", "CrossAttnDownBlock2DMusic", "DownBlock2D"],
    mid_block_type="UNetMidBlock2DCrossAttnMusic",
    up_block_types=["UpBlock2D", "CrossAttnUpBlock2DMusic", "CrossAttnUpBlock2DMusic", "CrossAttnUpBlock2DMusic"])


def main():
    torch.set_grad_enabled(False)
    refshim.install()
    from diffusers.models.unet_2d_condition_music import UNet2DConditionModelMusic as M
    cfg = TINY_MUSIC_CONFIG
    torch.manual_seed(0)
    ref = M(**{k: v for k, v in cfg.items()}).eval()
    # seeded, well-scaled weights keyed by parameter name (same generator as every other fixture)
    sd = {k: synth.synth_tensor(k, tuple(v.shape), 0) for k, v in ref.state_dict().items()}
    ref.load_state_dict(sd, strict=True)
    n_extra = sum(1 for k in sd if ".attentions2." in k or ".attentions3." in k)
    g = torch.Generator().manual_seed(21)
    B, L, Lb, Lc, D = 2, 10, 6, 5, cfg["cross_attention_dim"]
    x = torch.randn(B, 8, 32, 16, generator=g)
    ehs, beat, chord = (torch.randn(B, n, D, generator=g) for n in (L, Lb, Lc))
    mask = torch.ones(B, L, dtype=torch.bool); mask[1, 7:] = False
    bmask = torch.ones(B, Lb, dtype=torch.bool); bmask[0, 4:] = False
    cmask = torch.ones(B, Lc, dtype=torch.bool); cmask[1, 2:] = False
    t = torch.tensor(417)
    want = ref(x, t, ehs, beat, chord, encoder_attention_mask=mask, beat_attention_mask=bmask,
               chord_attention_mask=cmask).sample
    got = ounet.unet_forward(sd, cfg, x, t, ehs, mask, extra_streams=((beat, bmask), (chord, cmask)))
    d = float((want - got).abs().max())
    print(f"Mustango tiny UNet ({len(sd)} tensors, {n_extra} in the beat/chord attentions): |out| max "
          f"{float(want.abs().max()):.3f}, oracle vs reference {d:.3e}")
    assert d < 5e-5
    np.savez_compressed(os.path.join(GOLD, "tiny_unet_music.npz"), sample=x.numpy(), t=np.int64(417), ehs=ehs.numpy(),
                        beat=beat.numpy(), chord=chord.numpy(), mask=mask.numpy(), bmask=bmask.numpy(),
                        cmask=cmask.numpy(), out=want.numpy(),
                        keys=np.array(sorted(sd.keys())), shapes=np.array([str(tuple(sd[k].shape)) for k in sorted(sd)]))
    mp = os.path.join(GOLD, "MANIFEST.json")
    man = json.load(open(mp))
    man["checks"]["tiny_unet_music"] = {"oracle_vs_reference_max_abs": d, "tensors": len(sd)}
    json.dump(man, open(mp, "w"), indent=1)
    print("wrote", os.path.join(GOLD, "tiny_unet_music.npz"))


if __name__ == "__main__":
    main()
