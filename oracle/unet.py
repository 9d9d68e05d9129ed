"""CPU restatement of diffusers' UNet2DConditionModel.forward as used by Tango (TEST INFRASTRUCTURE ONLY).

Functional PyTorch fp32 over a plain state_dict with the diffusers key names. D/ below abbreviates
/root/reference/mustango/diffusers/src/diffusers/.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def timestep_embedding(timesteps: torch.Tensor, dim: int, flip_sin_to_cos: bool, freq_shift: float) -> torch.Tensor:
    """D/models/embeddings.py:22-62 (scale = 1, max_period = 10000)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32)
    exponent = exponent / (half - freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


def _lin(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _conv(sd: SD, p: str, x: torch.Tensor, stride: int = 1, padding: int = 1) -> torch.Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def _gn(sd: SD, p: str, x: torch.Tensor, groups: int, eps: float) -> torch.Tensor:
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def _ln(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def resnet_block(sd: SD, p: str, x: torch.Tensor, temb: torch.Tensor, groups: int, eps: float) -> torch.Tensor:
    """D/models/resnet.py:549-597 (time_embedding_norm='default', output_scale_factor=1, silu)."""
    h = F.silu(_gn(sd, p + ".norm1", x, groups, eps))
    h = _conv(sd, p + ".conv1", h)
    t = _lin(sd, p + ".time_emb_proj", F.silu(temb))[:, :, None, None]
    h = h + t
    h = F.silu(_gn(sd, p + ".norm2", h, groups, eps))
    h = _conv(sd, p + ".conv2", h)
    if (p + ".conv_shortcut.weight") in sd:
        x = _conv(sd, p + ".conv_shortcut", x, padding=0)
    return (x + h) / 1.0


def downsample2d(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """D/models/resnet.py:199-208 (Downsample2D, use_conv=True, padding=1): 3x3 stride-2 conv."""
    return _conv(sd, p + ".conv", x, stride=2, padding=1)


def upsample2d(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """D/models/resnet.py:126-161 (Upsample2D, use_conv=True): nearest x2 then 3x3 conv."""
    return _conv(sd, p + ".conv", F.interpolate(x, scale_factor=2.0, mode="nearest"))


def attention(sd: SD, p: str, x: torch.Tensor, ctx: Optional[torch.Tensor], heads: int,
              bias: Optional[torch.Tensor]) -> torch.Tensor:
    """D/models/attention_processor.py:500-540 — AttnProcessor2_0, the processor the reference selects on any torch that
    has F.scaled_dot_product_attention (attention_processor.py:104-111): q / k / v projections, heads split, SDPA with
    the additive mask bias expanded over heads and queries, out projection. (AttnProcessor's baddbmm + softmax + bmm,
    :263-299, gives the same fp32 result to round-off; SDPA is what actually runs — and what the CPU baseline of bench.py
    should time: it does not materialise the [B, heads, Lq, Lk] score tensor.) bias: additive [B, 1, Lk] or None."""
    B, Lq, Cc = x.shape
    src = x if ctx is None else ctx
    q = F.linear(x, sd[p + ".to_q.weight"])
    k = F.linear(src, sd[p + ".to_k.weight"])
    v = F.linear(src, sd[p + ".to_v.weight"])
    d = Cc // heads
    Lk = src.shape[1]
    q = q.view(B, Lq, heads, d).transpose(1, 2)
    k = k.view(B, Lk, heads, d).transpose(1, 2)
    v = v.view(B, Lk, heads, d).transpose(1, 2)
    mask = None if bias is None else bias[:, None, :, :].expand(B, heads, Lq, Lk)
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(B, Lq, Cc)
    return _lin(sd, p + ".to_out.0", o)


def transformer_2d(sd: SD, p: str, x: torch.Tensor, ctx: torch.Tensor, heads: int, groups: int,
                   bias: Optional[torch.Tensor]) -> torch.Tensor:
    """D/models/transformer_2d.py:214-321 (use_linear_projection=True, one BasicTransformerBlock,
    D/models/attention.py:276-335 with GEGLU feed-forward :412-433)."""
    B, Cc, H, W = x.shape
    res = x
    h = _gn(sd, p + ".norm", x, groups, 1e-6)
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, Cc)
    h = _lin(sd, p + ".proj_in", h)
    b = p + ".transformer_blocks.0"
    h = attention(sd, b + ".attn1", _ln(sd, b + ".norm1", h), None, heads, None) + h
    h = attention(sd, b + ".attn2", _ln(sd, b + ".norm2", h), ctx, heads, bias) + h
    n = _ln(sd, b + ".norm3", h)
    proj = _lin(sd, b + ".ff.net.0.proj", n)
    hid, gate = proj.chunk(2, dim=-1)
    h = _lin(sd, b + ".ff.net.2", hid * F.gelu(gate)) + h
    h = _lin(sd, p + ".proj_out", h)
    h = h.reshape(B, H, W, Cc).permute(0, 3, 1, 2)
    return h + res


def _mask_bias(m: Optional[torch.Tensor], dtype) -> Optional[torch.Tensor]:
    """bool mask (True = keep) -> additive bias [B, 1, L] (unet_2d_condition.py:575-579); other dtypes pass as bias."""
    if m is None:
        return None
    if m.dtype is torch.bool:
        m = (1 - m.to(dtype)) * -10000.0
    return m.unsqueeze(1)


def unet_forward(sd: SD, cfg: dict, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor,
                 encoder_attention_mask: Optional[torch.Tensor] = None, taps: Optional[dict] = None,
                 extra_streams=()) -> torch.Tensor:
    """D/models/unet_2d_condition.py:520-707 for the block types Tango's configs use
    (CrossAttnDownBlock2D / DownBlock2D / UNetMidBlock2DCrossAttn / UpBlock2D / CrossAttnUpBlock2D,
    D/models/unet_2d_blocks.py:1022-1074,1325-1349,577-598,2490-2513,2195-2248).

    `extra_streams` restates the Mustango variant (D/models/unet_2d_condition_music.py:536-757 with the *Music blocks of
    unet_2d_blocks.py:603-760,1079-1270,2251-2440): a sequence of (features [B, L, D], mask or None) pairs — beats then
    chords — each consumed by one more Transformer2DModel (`attentions2`, `attentions3`) right after the text one at
    every attention position. Empty for Tango."""
    boc = cfg["block_out_channels"]
    groups = cfg.get("norm_num_groups", 32)
    eps = cfg.get("norm_eps", 1e-5)
    ahd = cfg["attention_head_dim"]
    heads = ahd if isinstance(ahd, (list, tuple)) else [ahd] * len(boc)
    lpb = cfg.get("layers_per_block", 2)

    bias = _mask_bias(encoder_attention_mask, sample.dtype)
    extras = [(f, _mask_bias(m, sample.dtype)) for f, m in extra_streams]

    def attend(prefix: str, j: int, h: torch.Tensor, nheads: int) -> torch.Tensor:
        h = transformer_2d(sd, f"{prefix}.attentions.{j}", h, encoder_hidden_states, nheads, groups, bias)
        for n, (feat, fbias) in enumerate(extras):
            h = transformer_2d(sd, f"{prefix}.attentions{n + 2}.{j}", h, feat, nheads, groups, fbias)
        return h

    t = timestep
    if not torch.is_tensor(t):
        t = torch.tensor([t], dtype=torch.int64)
    elif t.dim() == 0:
        t = t[None]
    t = t.expand(sample.shape[0])
    temb = timestep_embedding(t, boc[0], cfg.get("flip_sin_to_cos", True), cfg.get("freq_shift", 0))
    temb = _lin(sd, "time_embedding.linear_2", F.silu(_lin(sd, "time_embedding.linear_1", temb)))

    h = _conv(sd, "conv_in", sample)
    skips = [h]
    for i, bt in enumerate(cfg["down_block_types"]):
        for j in range(lpb):
            h = resnet_block(sd, f"down_blocks.{i}.resnets.{j}", h, temb, groups, eps)
            if bt in ("CrossAttnDownBlock2D", "CrossAttnDownBlock2DMusic"):
                h = attend(f"down_blocks.{i}", j, h, heads[i])
            skips.append(h)
        if i != len(boc) - 1:
            h = downsample2d(sd, f"down_blocks.{i}.downsamplers.0", h)
            skips.append(h)
    if taps is not None:
        taps["down"] = h

    h = resnet_block(sd, "mid_block.resnets.0", h, temb, groups, eps)
    h = attend("mid_block", 0, h, heads[-1])
    h = resnet_block(sd, "mid_block.resnets.1", h, temb, groups, eps)
    if taps is not None:
        taps["mid"] = h

    rheads = list(reversed(heads))
    for i, bt in enumerate(cfg["up_block_types"]):
        for j in range(lpb + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet_block(sd, f"up_blocks.{i}.resnets.{j}", h, temb, groups, eps)
            if bt in ("CrossAttnUpBlock2D", "CrossAttnUpBlock2DMusic"):
                h = attend(f"up_blocks.{i}", j, h, rheads[i])
        if i != len(boc) - 1:
            h = upsample2d(sd, f"up_blocks.{i}.upsamplers.0", h)

    h = F.silu(_gn(sd, "conv_norm_out", h, groups, eps))
    return _conv(sd, "conv_out", h)
