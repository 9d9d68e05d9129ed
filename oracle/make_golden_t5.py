"""Generate tests/golden/tiny_t5.npz from the installed `transformers.T5EncoderModel` (TEST INFRASTRUCTURE ONLY).

The FLAN-T5 encoder is a pip dependency of the reference (models.py:98-100), not code under /root/reference, so this
script needs no reference checkout: it instantiates T5EncoderModel from the tiny config, loads the seeded synthetic
state_dict, runs it in eval mode on CPU fp32 and stores inputs + outputs; it also checks oracle/t5.py against it.
Run from the repo root:  python -m oracle.make_golden_t5
"""
import json
import os

import numpy as np
import torch

from oracle import t5 as ot5
from tango_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def hf_model(cfg, sd):
    import transformers
    from transformers import T5Config, T5EncoderModel
    hc = T5Config(vocab_size=cfg["vocab_size"], d_model=cfg["d_model"], d_kv=cfg["d_kv"], d_ff=cfg["d_ff"],
                  num_layers=cfg["num_layers"], num_heads=cfg["num_heads"],
                  relative_attention_num_buckets=cfg["relative_attention_num_buckets"],
                  relative_attention_max_distance=cfg["relative_attention_max_distance"],
                  layer_norm_epsilon=cfg["layer_norm_epsilon"], feed_forward_proj=cfg["feed_forward_proj"],
                  dropout_rate=0.0)
    m = T5EncoderModel(hc).eval()
    full = dict(sd)
    full["encoder.embed_tokens.weight"] = sd["shared.weight"]
    missing, unexpected = m.load_state_dict(full, strict=False)
    assert not unexpected and all("embed_tokens" in k or "shared" in k for k in missing), (missing, unexpected)
    return m, transformers.__version__


def inputs(cfg, B=3, Lt=10, seed=11):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, cfg["vocab_size"], (B, Lt), generator=g)
    mask = torch.ones(B, Lt, dtype=torch.long)
    mask[1, 6:] = 0           # padded prompt
    if B > 2:
        mask[2, 1:] = 0       # the "" prompt of classifier-free guidance: one EOS token, the rest padding
    ids = ids * mask          # pad id 0
    return ids, mask


def main():
    torch.set_grad_enabled(False)
    cfg = synth.TINY_T5_CONFIG
    sd = synth.synth_state_dict(synth.t5_encoder_param_shapes(cfg), seed=0)
    m, ver = hf_model(cfg, sd)
    ids, mask = inputs(cfg)
    ref = m(input_ids=ids, attention_mask=mask)[0]
    orc = ot5.t5_encoder(sd, cfg, ids, mask)
    d = float((ref - orc).abs().max())
    print(f"tiny T5 encoder: |out| max {float(ref.abs().max()):.3f}, oracle vs transformers {ver}: {d:.3e}")
    assert d < 2e-5
    # a longer sequence exercises the logarithmic buckets (|key - query| >= 8) and the 64-key tiling of the kernel
    ids2, mask2 = inputs(cfg, B=2, Lt=150, seed=12)
    mask2[:] = 1
    mask2[1, 140:] = 0
    ref2 = m(input_ids=ids2, attention_mask=mask2)[0]
    d2 = float((ref2 - ot5.t5_encoder(sd, cfg, ids2, mask2)).abs().max())
    print(f"  L = 150: oracle vs transformers {d2:.3e}")
    assert d2 < 2e-5
    np.savez_compressed(os.path.join(GOLD, "tiny_t5.npz"), ids=ids.numpy(), mask=mask.numpy(), out=ref.numpy(),
                        ids_long=ids2.numpy(), mask_long=mask2.numpy(), out_long=ref2.numpy())
    mp = os.path.join(GOLD, "MANIFEST.json")
    man = json.load(open(mp))
    man["checks"]["tiny_t5"] = {"source": f"transformers {ver} T5EncoderModel (pip dependency of the reference)",
                                "oracle_vs_transformers_max_abs": d, "L150": d2}
    json.dump(man, open(mp, "w"), indent=1)
    print("wrote", os.path.join(GOLD, "tiny_t5.npz"))


if __name__ == "__main__":
    main()
