"""CPU restatement of the HiFi-GAN generator + vocoder_infer (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/audioldm/hifigan/models.py:96-103 (ResBlock), :149-165 (Generator.forward) and
utilities.py:76-86 (vocoder_infer) with the HIFIGAN_16K_64 config (utilities.py:9-40). Keys: `vocoder.*`
(weight norm already removed, tango.py load format).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
UPSAMPLE_RATES = (5, 4, 2, 2, 2)
UPSAMPLE_KERNELS = (16, 16, 8, 4, 4)
RES_KERNELS = (3, 7, 11)
RES_DILATIONS = (1, 3, 5)
LRELU_SLOPE = 0.1


def _pad(k, d=1):
    return int((k * d - d) / 2)


def resblock(sd: SD, p: str, x, k):
    for i, d in enumerate(RES_DILATIONS):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, sd[f"{p}.convs1.{i}.weight"], sd[f"{p}.convs1.{i}.bias"], padding=_pad(k, d), dilation=d)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, sd[f"{p}.convs2.{i}.weight"], sd[f"{p}.convs2.{i}.bias"], padding=_pad(k, 1))
        x = xt + x
    return x


def generator(sd: SD, mel: torch.Tensor, prefix: str = "vocoder") -> torch.Tensor:
    """mel (B, 64, T) -> (B, 1, 160*T + ...) in [-1, 1]."""
    p = prefix
    x = F.conv1d(mel, sd[p + ".conv_pre.weight"], sd[p + ".conv_pre.bias"], padding=3)
    nk = len(RES_KERNELS)
    for i, (u, k) in enumerate(zip(UPSAMPLE_RATES, UPSAMPLE_KERNELS)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, sd[f"{p}.ups.{i}.weight"], sd[f"{p}.ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, rk in enumerate(RES_KERNELS):
            r = resblock(sd, f"{p}.resblocks.{i * nk + j}", x, rk)
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x)  # default slope 0.01 (models.py:161)
    x = F.conv1d(x, sd[p + ".conv_post.weight"], sd[p + ".conv_post.bias"], padding=3)
    return torch.tanh(x)


def decode_to_waveform(sd: SD, mel_img: torch.Tensor, prefix: str = "vocoder"):
    """autoencoder.py:66-69 + utilities.py:76-86: (B,1,T,64) -> (wave float32 (B, L), int16 (B, L))."""
    dec = mel_img.squeeze(1).permute(0, 2, 1)
    wav = generator(sd, dec, prefix).squeeze(1)
    with np.errstate(invalid="ignore"):
        wi = (wav.cpu().numpy() * 32768).astype("int16")
    return wav, wi
