"""CPU restatement of AudioDiffusion.inference and the Tango.generate tail (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/models.py:210-264 (CFG denoising loop, prepare_latents) and tango.py:43-49. Text encoding
(models.py:266-305) is outside the accelerated path: `prompt_embeds` / `mask` are injected, already in the
[uncond; cond] order that encode_text_classifier_free returns. Randomness is injected too: `latents0` replaces the
randn_tensor of models.py:261 and `noises[i]` the per-step draw of scheduling_ddpm.py:331-335.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import unet as ounet


def inference(unet_sd, unet_cfg, scheduler, prompt_embeds, mask, num_steps, guidance_scale, latents0,
              noises: Optional[List[torch.Tensor]] = None, trace: Optional[list] = None, extra_streams=()):
    """`extra_streams`: the Mustango loop (mustango/models.py:540-600, MusicAudioDiffusion.inference) is this same loop
    with the (already CFG-duplicated) encoded beats and chords handed to the UNet at every step."""
    cfg = guidance_scale > 1.0
    scheduler.set_timesteps(num_steps)
    latents = latents0 * scheduler.init_noise_sigma
    for i, t in enumerate(scheduler.timesteps):
        x = torch.cat([latents] * 2) if cfg else latents
        pred = ounet.unet_forward(unet_sd, unet_cfg, x, t, prompt_embeds, mask, extra_streams=extra_streams)
        if cfg:
            u, c = pred.chunk(2)
            pred = u + guidance_scale * (c - u)
        if noises is not None:
            latents = scheduler.step(pred, t, latents, noises[i] if int(t) > 0 else None)
        else:
            latents = scheduler.step(pred, t, latents)
        if trace is not None:
            trace.append(latents.clone())
    return latents
