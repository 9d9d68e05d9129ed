"""CPU restatement of the AudioLDM VAE decoder (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/audioldm/variational_autoencoder/autoencoder.py:60-64,116-124 and modules.py
(Decoder.forward :650-683, ResnetBlock :155-175, AttnBlock :204-230, Upsample :53-57) with the
mustango/configs/vae_config.json hyper-parameters (ch 128, ch_mult [1,2,4], 2 res blocks, z_channels 8).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _swish(x):
    return x * torch.sigmoid(x)


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)


def _conv(sd, p, x, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=padding)


def resnet_block(sd: SD, p: str, x):
    h = _conv(sd, p + ".conv1", _swish(_gn(sd, p + ".norm1", x)))
    h = _conv(sd, p + ".conv2", _swish(_gn(sd, p + ".norm2", h)))
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv(sd, p + ".nin_shortcut", x, padding=0)
    return x + h


def attn_block(sd: SD, p: str, x):
    h = _gn(sd, p + ".norm", x)
    q = _conv(sd, p + ".q", h, 0)
    k = _conv(sd, p + ".k", h, 0)
    v = _conv(sd, p + ".v", h, 0)
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(sd, p + ".proj_out", h, 0)


def decode_first_stage(sd: SD, z: torch.Tensor, scale_factor: float, ch_mult=(1, 2, 4), num_res_blocks=2) -> torch.Tensor:
    """z (B, 8, T/4, 16) -> mel (B, 1, T, 64). Keys as in AutoencoderKL.state_dict()."""
    z = 1.0 / scale_factor * z
    z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    d = "decoder"
    h = _conv(sd, d + ".conv_in", z)
    h = resnet_block(sd, d + ".mid.block_1", h)
    h = attn_block(sd, d + ".mid.attn_1", h)
    h = resnet_block(sd, d + ".mid.block_2", h)
    for lvl in reversed(range(len(ch_mult))):
        for blk in range(num_res_blocks + 1):
            h = resnet_block(sd, f"{d}.up.{lvl}.block.{blk}", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"{d}.up.{lvl}.upsample.conv", h)
    h = _swish(_gn(sd, d + ".norm_out", h))
    return _conv(sd, d + ".conv_out", h)


# ---------------------------------------------------------------------------------------------------------------------
# Encoder side (SURVEY.md section 8(f).2 — the step *before* the hot path, used for training / audio-to-audio; restated
# and pinned now so that the kernels of a later round have their checker ready).
def downsample(sd: SD, p: str, x):
    """modules.py:76-94 (Downsample, with_conv): zero-pad one column / row at the END of each axis, 3x3 stride-2 conv."""
    x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
    return F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], stride=2, padding=0)


def encoder_forward(sd: SD, x: torch.Tensor, ch_mult=(1, 2, 4), num_res_blocks=2) -> torch.Tensor:
    """modules.py:519-543 (Encoder.forward) for the Tango VAE config (no attention inside the levels:
    attn_resolutions = [], no time-stride-4 levels): mel (B, 1, T, 64) -> (B, 2 * z_channels, T/4, 16)."""
    e = "encoder"
    h = _conv(sd, e + ".conv_in", x)
    for lvl in range(len(ch_mult)):
        for blk in range(num_res_blocks):
            h = resnet_block(sd, f"{e}.down.{lvl}.block.{blk}", h)
        if lvl != len(ch_mult) - 1:
            h = downsample(sd, f"{e}.down.{lvl}.downsample", h)
    h = resnet_block(sd, e + ".mid.block_1", h)
    h = attn_block(sd, e + ".mid.attn_1", h)
    h = resnet_block(sd, e + ".mid.block_2", h)
    h = _swish(_gn(sd, e + ".norm_out", h))
    return _conv(sd, e + ".conv_out", h)


def encode_first_stage(sd: SD, mel: torch.Tensor, ch_mult=(1, 2, 4), num_res_blocks=2):
    """autoencoder.py:52-58,110-112 (encode / encode_first_stage) + distributions.py:24-41: returns (mean, std) of the
    diagonal Gaussian posterior; `posterior.sample()` = mean + std * randn, `posterior.mode()` = mean;
    the latent handed to the diffusion model is `scale_factor * sample` (models.py: get_first_stage_encoding)."""
    moments = F.conv2d(encoder_forward(sd, mel, ch_mult, num_res_blocks), sd["quant_conv.weight"], sd["quant_conv.bias"])
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    return mean, torch.exp(0.5 * logvar)
