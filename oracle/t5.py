"""CPU restatement of the FLAN-T5 encoder Tango conditions on (TEST INFRASTRUCTURE ONLY).

The reference builds it with `T5EncoderModel.from_pretrained(text_encoder_name)` (/root/reference/models.py:98-100) and
calls it as `self.text_encoder(input_ids=..., attention_mask=...)[0]` (models.py:129-147, 266-305). The arithmetic is not
under /root/reference: it lives in the pip dependency `transformers` (requirements.txt pins transformers==4.27.0; this
container has 5.5.0, same encoder arithmetic) in transformers/models/t5/modeling_t5.py — T5LayerNorm, T5Attention
(relative position buckets, no 1/sqrt(d) scaling), T5DenseGatedActDense ("gated-gelu": tanh-form GELU), T5Stack.
This file restates that published algorithm functionally over a plain state_dict with the T5EncoderModel key names;
tests/test_oracle_pins.py pins it against the installed transformers.T5EncoderModel on seeded random weights, and
tests/golden/tiny_t5.npz holds outputs generated from that model by oracle/make_golden.py.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

TINY_T5_CONFIG = {"vocab_size": 96, "d_model": 128, "d_kv": 64, "num_heads": 2, "d_ff": 256, "num_layers": 2,
                  "relative_attention_num_buckets": 32, "relative_attention_max_distance": 128,
                  "layer_norm_epsilon": 1e-6, "feed_forward_proj": "gated-gelu"}


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """T5LayerNorm: scale by the root mean square only (no mean subtraction, no bias), statistics in fp32."""
    var = x.float().pow(2).mean(-1, keepdim=True)
    return weight * (x * torch.rsqrt(var + eps))


def gelu_new(x: torch.Tensor) -> torch.Tensor:
    """tanh-form GELU ("gelu_new" activation of the gated-gelu feed-forward)."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def relative_position_bucket(rel: torch.Tensor, num_buckets: int, max_distance: int) -> torch.Tensor:
    """Bidirectional bucket of rel = key position - query position: half of the buckets for each sign; within a sign
    the first half are exact offsets, the rest logarithmic bins up to max_distance (clamped to the last bucket)."""
    half = num_buckets // 2
    bucket = (rel > 0).to(torch.long) * half
    n = rel.abs()
    exact = half // 2
    log_bin = exact + (torch.log(n.float() / exact) / math.log(max_distance / exact) * (half - exact)).to(torch.long)
    log_bin = torch.clamp(log_bin, max=half - 1)
    return bucket + torch.where(n < exact, n, log_bin)


def position_bias(sd: SD, cfg: dict, L: int) -> torch.Tensor:
    """[heads, L, L] learned bias of block 0 (shared by every block): table[bucket(key - query), head]."""
    pos = torch.arange(L, dtype=torch.long)
    rel = pos[None, :] - pos[:, None]
    b = relative_position_bucket(rel, cfg["relative_attention_num_buckets"], cfg["relative_attention_max_distance"])
    table = sd["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]  # [buckets, heads]
    return table[b].permute(2, 0, 1)


def t5_encoder(sd: SD, cfg: dict, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """last_hidden_state [B, L, d_model] of T5EncoderModel(input_ids, attention_mask) in eval mode (dropout off)."""
    if cfg.get("feed_forward_proj", "gated-gelu") != "gated-gelu":
        raise ValueError("only the FLAN-T5 (gated-gelu) feed-forward is restated")
    B, L = input_ids.shape
    H, dk, eps = cfg["num_heads"], cfg["d_kv"], cfg["layer_norm_epsilon"]
    emb = sd["shared.weight"] if "shared.weight" in sd else sd["encoder.embed_tokens.weight"]
    x = F.embedding(input_ids, emb).float()
    # additive mask: 0 for kept keys, the most negative fp32 for padded ones, added to the position bias
    ext = (1.0 - attention_mask[:, None, None, :].float()) * torch.finfo(torch.float32).min
    bias = position_bias(sd, cfg, L)[None] + ext                                   # [B, H, L, L]
    for i in range(cfg["num_layers"]):
        p = f"encoder.block.{i}.layer."
        n = rms_norm(x, sd[p + "0.layer_norm.weight"], eps)
        q = F.linear(n, sd[p + "0.SelfAttention.q.weight"]).view(B, L, H, dk).transpose(1, 2)
        k = F.linear(n, sd[p + "0.SelfAttention.k.weight"]).view(B, L, H, dk).transpose(1, 2)
        v = F.linear(n, sd[p + "0.SelfAttention.v.weight"]).view(B, L, H, dk).transpose(1, 2)
        scores = torch.matmul(q, k.transpose(3, 2)) + bias                          # no 1/sqrt(d) in T5
        attn = F.softmax(scores.float(), dim=-1)
        ctx = torch.matmul(attn, v).transpose(1, 2).reshape(B, L, H * dk)
        x = x + F.linear(ctx, sd[p + "0.SelfAttention.o.weight"])
        n = rms_norm(x, sd[p + "1.layer_norm.weight"], eps)
        h = gelu_new(F.linear(n, sd[p + "1.DenseReluDense.wi_0.weight"])) * F.linear(n, sd[p + "1.DenseReluDense.wi_1.weight"])
        x = x + F.linear(h, sd[p + "1.DenseReluDense.wo.weight"])
    return rms_norm(x, sd["encoder.final_layer_norm.weight"], eps)
