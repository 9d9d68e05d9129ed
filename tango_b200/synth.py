"""Synthetic (seeded, random) weights and inputs in the reference's state_dict layouts.

No pretrained checkpoint is reachable offline, so parity and the benchmark run on random weights that exercise
identical arithmetic. Every tensor is drawn from its own CPU generator seeded by
crc32(key) ^ seed, so any subset of keys reproduces the same values on any machine.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import torch

Shapes = "OrderedDict[str, Tuple[int, ...]]"

TINY_UNET_CONFIG = {
    "act_fn": "silu", "attention_head_dim": [1, 2, 4, 4], "block_out_channels": [64, 128, 256, 256],
    "center_input_sample": False, "cross_attention_dim": 128,
    "down_block_types": ["CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"],
    "downsample_padding": 1, "flip_sin_to_cos": True, "freq_shift": 0, "in_channels": 8, "layers_per_block": 2,
    "mid_block_scale_factor": 1, "norm_eps": 1e-05, "norm_num_groups": 32, "out_channels": 8,
    "up_block_types": ["UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"],
    "use_linear_projection": True, "upcast_attention": True,
}

_MUSIC_BLOCKS = dict(
    down_block_types=["CrossAttnDownBlock2DMusic", "CrossAttnDownBlock2DMusic", "CrossAttnDownBlock2DMusic", "DownBlock2D"],
    mid_block_type="UNetMidBlock2DCrossAttnMusic",
    up_block_types=["UpBlock2D", "CrossAttnUpBlock2DMusic", "CrossAttnUpBlock2DMusic", "CrossAttnUpBlock2DMusic"])
# Mustango variant (mustango/configs/music_diffusion_model_config.json): beat + chord cross-attentions at every position
TINY_MUSIC_UNET_CONFIG = dict(TINY_UNET_CONFIG, **_MUSIC_BLOCKS)

BASE_UNET_CONFIG = dict(TINY_UNET_CONFIG, attention_head_dim=[5, 10, 20, 20], block_out_channels=[320, 640, 1280, 1280],
                        cross_attention_dim=1024)
XL_UNET_CONFIG = dict(BASE_UNET_CONFIG, cross_attention_dim=2048)
MUSIC_UNET_CONFIG = dict(BASE_UNET_CONFIG, **_MUSIC_BLOCKS)

VAE_CONFIG = {"image_key": "fbank", "subband": 1, "embed_dim": 8, "time_shuffle": 1,
              "ddconfig": {"double_z": True, "z_channels": 8, "resolution": 256, "downsample_time": False,
                           "in_channels": 1, "out_ch": 1, "ch": 128, "ch_mult": [1, 2, 4], "num_res_blocks": 2,
                           "attn_resolutions": [], "dropout": 0.0},
              "scale_factor": 0.9227914214134216}

# stft_config.json of the declare-lab/tango checkpoints (tango.py:15)
STFT_CONFIG = {"filter_length": 1024, "hop_length": 160, "win_length": 1024, "n_mel_channels": 64,
               "sampling_rate": 16000, "mel_fmin": 0, "mel_fmax": 8000}

HIFIGAN_CONFIG = {"upsample_rates": [5, 4, 2, 2, 2], "upsample_kernel_sizes": [16, 16, 8, 4, 4],
                  "upsample_initial_channel": 1024, "resblock_kernel_sizes": [3, 7, 11],
                  "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]], "num_mels": 64}


def _heads(cfg):
    ahd = cfg["attention_head_dim"]
    return list(ahd) if isinstance(ahd, (list, tuple)) else [ahd] * len(cfg["block_out_channels"])


def unet_param_shapes(cfg: dict) -> Shapes:
    """Parameter names and shapes of diffusers' UNet2DConditionModel for Tango's block types (686 tensors for the
    base config), in construction order (unet_2d_condition.py ctor; unet_2d_blocks.py:get_down/up_block)."""
    s: Shapes = OrderedDict()
    boc = cfg["block_out_channels"]
    cin, cout = cfg["in_channels"], cfg["out_channels"]
    xd = cfg["cross_attention_dim"]
    lpb = cfg["layers_per_block"]
    ted = boc[0] * 4

    def lin(p, o, i, bias=True):
        s[p + ".weight"] = (o, i)
        if bias:
            s[p + ".bias"] = (o,)

    def conv(p, o, i, k):
        s[p + ".weight"] = (o, i, k, k)
        s[p + ".bias"] = (o,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    def resnet(p, i, o):
        norm(p + ".norm1", i)
        conv(p + ".conv1", o, i, 3)
        lin(p + ".time_emb_proj", o, ted)
        norm(p + ".norm2", o)
        conv(p + ".conv2", o, o, 3)
        if i != o:
            conv(p + ".conv_shortcut", o, i, 1)

    def attn(p, c, kv):
        lin(p + ".to_q", c, c, False)
        lin(p + ".to_k", c, kv, False)
        lin(p + ".to_v", c, kv, False)
        lin(p + ".to_out.0", c, c)

    def transformer(p, c):
        norm(p + ".norm", c)
        lin(p + ".proj_in", c, c)
        b = p + ".transformer_blocks.0"
        norm(b + ".norm1", c)
        attn(b + ".attn1", c, c)
        norm(b + ".norm2", c)
        attn(b + ".attn2", c, xd)
        norm(b + ".norm3", c)
        lin(b + ".ff.net.0.proj", 8 * c, c)
        lin(b + ".ff.net.2", c, 4 * c)
        lin(p + ".proj_out", c, c)

    conv("conv_in", boc[0], cin, 3)
    lin("time_embedding.linear_1", ted, boc[0])
    lin("time_embedding.linear_2", ted, ted)
    out_ch = boc[0]
    for i, bt in enumerate(cfg["down_block_types"]):
        in_ch, out_ch = out_ch, boc[i]
        # diffusers registers attentions before resnets inside each block
        if bt in ("CrossAttnDownBlock2D", "CrossAttnDownBlock2DMusic"):
            # the Mustango blocks (unet_2d_blocks.py:1079-1270) add attentions2 (beats) and attentions3 (chords)
            for name in ("attentions", "attentions2", "attentions3")[:3 if bt.endswith("Music") else 1]:
                for j in range(lpb):
                    transformer(f"down_blocks.{i}.{name}.{j}", out_ch)
        for j in range(lpb):
            resnet(f"down_blocks.{i}.resnets.{j}", in_ch if j == 0 else out_ch, out_ch)
        if i != len(boc) - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", out_ch, out_ch, 3)
    rboc = list(reversed(boc))
    out_ch = rboc[0]
    up: Shapes = OrderedDict()
    main = s
    s = up
    for i, bt in enumerate(cfg["up_block_types"]):
        prev_out, out_ch = out_ch, rboc[i]
        in_ch = rboc[min(i + 1, len(boc) - 1)]
        if bt in ("CrossAttnUpBlock2D", "CrossAttnUpBlock2DMusic"):
            for name in ("attentions", "attentions2", "attentions3")[:3 if bt.endswith("Music") else 1]:
                for j in range(lpb + 1):
                    transformer(f"up_blocks.{i}.{name}.{j}", out_ch)
        for j in range(lpb + 1):
            skip = in_ch if j == lpb else out_ch
            rin = prev_out if j == 0 else out_ch
            resnet(f"up_blocks.{i}.resnets.{j}", rin + skip, out_ch)
        if i != len(boc) - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", out_ch, out_ch, 3)
    s = main
    s.update(up)
    transformer("mid_block.attentions.0", boc[-1])
    if cfg.get("mid_block_type", "UNetMidBlock2DCrossAttn") == "UNetMidBlock2DCrossAttnMusic":
        transformer("mid_block.attentions2.0", boc[-1])
        transformer("mid_block.attentions3.0", boc[-1])
    resnet("mid_block.resnets.0", boc[-1], boc[-1])
    resnet("mid_block.resnets.1", boc[-1], boc[-1])
    norm("conv_norm_out", boc[0])
    conv("conv_out", cout, boc[0], 3)
    return s


def vae_decoder_param_shapes(vae_cfg: dict = VAE_CONFIG) -> Shapes:
    """decoder.*, post_quant_conv.* and vocoder.* of the AudioLDM AutoencoderKL state_dict
    (audioldm/variational_autoencoder/modules.py:545-648; audioldm/hifigan/models.py:106-147)."""
    s: Shapes = OrderedDict()
    dd = vae_cfg["ddconfig"]
    ch, mult, nrb, zc = dd["ch"], dd["ch_mult"], dd["num_res_blocks"], dd["z_channels"]

    def conv(p, o, i, k):
        s[p + ".weight"] = (o, i, k, k)
        s[p + ".bias"] = (o,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    def res(p, i, o):
        norm(p + ".norm1", i)
        conv(p + ".conv1", o, i, 3)
        norm(p + ".norm2", o)
        conv(p + ".conv2", o, o, 3)
        if i != o:
            conv(p + ".nin_shortcut", o, i, 1)

    conv("post_quant_conv", zc, vae_cfg["embed_dim"], 1)
    bi = ch * mult[-1]
    conv("decoder.conv_in", bi, zc, 3)
    res("decoder.mid.block_1", bi, bi)
    norm("decoder.mid.attn_1.norm", bi)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"decoder.mid.attn_1.{n}", bi, bi, 1)
    res("decoder.mid.block_2", bi, bi)
    for lvl in reversed(range(len(mult))):
        bo = ch * mult[lvl]
        for b in range(nrb + 1):
            res(f"decoder.up.{lvl}.block.{b}", bi, bo)
            bi = bo
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", bi, bi, 3)
    norm("decoder.norm_out", bi)
    conv("decoder.conv_out", dd["out_ch"], bi, 3)

    h = HIFIGAN_CONFIG
    c0 = h["upsample_initial_channel"]
    s["vocoder.conv_pre.weight"] = (c0, h["num_mels"], 7)
    s["vocoder.conv_pre.bias"] = (c0,)
    nk = len(h["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        ci, co = c0 // (2 ** i), c0 // (2 ** (i + 1))
        s[f"vocoder.ups.{i}.weight"] = (ci, co, k)
        s[f"vocoder.ups.{i}.bias"] = (co,)
        for j, rk in enumerate(h["resblock_kernel_sizes"]):
            for grp in ("convs1", "convs2"):
                for d in range(3):
                    s[f"vocoder.resblocks.{i * nk + j}.{grp}.{d}.weight"] = (co, co, rk)
                    s[f"vocoder.resblocks.{i * nk + j}.{grp}.{d}.bias"] = (co,)
    s["vocoder.conv_post.weight"] = (1, c0 // (2 ** len(h["upsample_rates"])), 7)
    s["vocoder.conv_post.bias"] = (1,)
    return s


def vae_encoder_param_shapes(vae_cfg: dict = None) -> Shapes:
    """encoder.* and quant_conv.* of the AudioLDM AutoencoderKL state_dict
    (audioldm/variational_autoencoder/modules.py:419-517; autoencoder.py:28-29) — the "next" row 2 of the scope table."""
    vae_cfg = vae_cfg or VAE_CONFIG
    s: Shapes = OrderedDict()
    dd = vae_cfg["ddconfig"]
    ch, mult, nrb, zc = dd["ch"], dd["ch_mult"], dd["num_res_blocks"], dd["z_channels"]

    def conv(p, o, i, k):
        s[p + ".weight"] = (o, i, k, k)
        s[p + ".bias"] = (o,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    def res(p, i, o):
        norm(p + ".norm1", i)
        conv(p + ".conv1", o, i, 3)
        norm(p + ".norm2", o)
        conv(p + ".conv2", o, o, 3)
        if i != o:
            conv(p + ".nin_shortcut", o, i, 1)

    conv("encoder.conv_in", ch, dd["in_channels"], 3)
    bi = ch
    for lvl in range(len(mult)):
        bo = ch * mult[lvl]
        for b in range(nrb):
            res(f"encoder.down.{lvl}.block.{b}", bi, bo)
            bi = bo
        if lvl != len(mult) - 1:
            conv(f"encoder.down.{lvl}.downsample.conv", bi, bi, 3)
    res("encoder.mid.block_1", bi, bi)
    norm("encoder.mid.attn_1.norm", bi)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"encoder.mid.attn_1.{n}", bi, bi, 1)
    res("encoder.mid.block_2", bi, bi)
    norm("encoder.norm_out", bi)
    conv("encoder.conv_out", 2 * zc if dd.get("double_z", True) else zc, bi, 3)
    conv("quant_conv", 2 * vae_cfg["embed_dim"], 2 * zc, 1)
    return s


TINY_T5_CONFIG = {"vocab_size": 96, "d_model": 128, "d_kv": 64, "num_heads": 2, "d_ff": 256, "num_layers": 2,
                  "relative_attention_num_buckets": 32, "relative_attention_max_distance": 128,
                  "layer_norm_epsilon": 1e-6, "feed_forward_proj": "gated-gelu"}
# google/flan-t5-large (Tango, Tango 2) and google/flan-t5-xl (the XL UNet config) encoder stacks
FLAN_T5_LARGE_CONFIG = dict(TINY_T5_CONFIG, vocab_size=32128, d_model=1024, num_heads=16, d_ff=2816, num_layers=24)
FLAN_T5_XL_CONFIG = dict(TINY_T5_CONFIG, vocab_size=32128, d_model=2048, num_heads=32, d_ff=5120, num_layers=24)


def t5_encoder_param_shapes(cfg: dict) -> Shapes:
    """T5EncoderModel state_dict layout (transformers/models/t5/modeling_t5.py; built at models.py:98-100)."""
    d, inner, ff = cfg["d_model"], cfg["num_heads"] * cfg["d_kv"], cfg["d_ff"]
    s: Shapes = OrderedDict()
    s["shared.weight"] = (cfg["vocab_size"], d)
    for i in range(cfg["num_layers"]):
        p = f"encoder.block.{i}.layer."
        for n in ("q", "k", "v"):
            s[p + f"0.SelfAttention.{n}.weight"] = (inner, d)
        s[p + "0.SelfAttention.o.weight"] = (d, inner)
        if i == 0:
            s[p + "0.SelfAttention.relative_attention_bias.weight"] = (cfg["relative_attention_num_buckets"],
                                                                      cfg["num_heads"])
        s[p + "0.layer_norm.weight"] = (d,)
        s[p + "1.DenseReluDense.wi_0.weight"] = (ff, d)
        s[p + "1.DenseReluDense.wi_1.weight"] = (ff, d)
        s[p + "1.DenseReluDense.wo.weight"] = (d, ff)
        s[p + "1.layer_norm.weight"] = (d,)
    s["encoder.final_layer_norm.weight"] = (d,)
    return s


def synth_tensor(key: str, shape, seed: int = 0) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    leaf = key.rsplit(".", 1)[-1]
    is_norm = any(t in key for t in (".norm", "norm_out", "conv_norm_out", "layer_norm")) and len(shape) == 1
    if key == "shared.weight":                       # token embeddings: unit scale, like the T5 initialiser
        return torch.randn(shape, generator=g)
    if key.endswith("SelfAttention.q.weight"):       # T5 has no 1/sqrt(d_kv) in the scores: fold it into q as T5's init does
        return torch.randn(shape, generator=g) / math.sqrt(shape[1] * 64)
    if is_norm:
        if leaf == "weight":
            return 1.0 + 0.1 * torch.randn(shape, generator=g)
        return 0.05 * torch.randn(shape, generator=g)
    if leaf == "bias":
        return 0.02 * torch.randn(shape, generator=g)
    if "vocoder.ups" in key:  # ConvTranspose1d weight (Cin, Cout, k): fan_in per output sample ~ Cin * k / stride
        fan_in = shape[0] * max(1, shape[2] // 2)
    else:
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
    return torch.randn(shape, generator=g) / math.sqrt(fan_in)


def synth_state_dict(shapes: Shapes, seed: int = 0, device="cpu") -> Dict[str, torch.Tensor]:
    return OrderedDict((k, synth_tensor(k, shp, seed).to(device)) for k, shp in shapes.items())


def synth_conditioning(batch: int, seq: int, dim: int, seed: int = 1, masked_tail: int = 0):
    """Synthetic stand-in for encode_text_classifier_free (models.py:266-305): returns
    (prompt_embeds [2B, L, D] = [uncond; cond], bool mask [2B, L])."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    cond = torch.randn(batch, seq, dim, generator=g)
    uncond = torch.randn(1, seq, dim, generator=g).expand(batch, seq, dim).clone()
    mask = torch.ones(2 * batch, seq, dtype=torch.bool)
    if masked_tail:
        mask[batch:, seq - masked_tail:] = False  # padded prompt tokens
        mask[:batch, 1:] = False                  # T5("") is a single EOS token followed by padding
    return torch.cat([uncond, cond]), mask


def synth_noise(batch: int, steps: int, shape=(8, 256, 16), seed: int = 1234):
    """Initial latents + one noise tensor per step, drawn in the reference's order (models.py:261;
    scheduling_ddpm.py:331-335)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    lat = torch.randn((batch, *shape), generator=g)
    noises = [torch.randn((batch, *shape), generator=g) for _ in range(steps)]
    return lat, noises
