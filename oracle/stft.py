"""CPU restatement of the mel front-end in front of the VAE encoder (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/audioldm/audio/stft.py (STFT.__init__ :18-50, STFT.transform :52-83,
TacotronSTFT.mel_spectrogram :161-186), audioldm/audio/audio_processing.py:85-91 (dynamic_range_compression) and
tools/torch_tools.py:9-54,57-77 (normalize_wav, pad_wav, _pad_spec, get_mel_from_wav / wav_to_fbank after file I/O).
The mel filter bank is a *buffer* of the checkpoint (`pytorch_model_stft.bin: mel_basis`; the reference builds it with
librosa, which this image does not have), so it is an input here; the windowed Fourier basis is rebuilt from its
definition and must equal the checkpoint's `stft_fn.forward_basis`.
Tango's stft_config.json: filter_length 1024, hop_length 160, win_length 1024, n_mel_channels 64, 16 kHz, fmin 0, fmax 8000.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def forward_basis(filter_length: int, win_length: int) -> torch.Tensor:
    """[2 * (filter_length/2 + 1), 1, filter_length]: real then imaginary rows of the DFT matrix, each multiplied by the
    periodic Hann window of `win_length` zero-padded (centred) to `filter_length` (stft.py:26-47)."""
    fb = np.fft.fft(np.eye(filter_length))
    cutoff = filter_length // 2 + 1
    fb = np.vstack([np.real(fb[:cutoff, :]), np.imag(fb[:cutoff, :])])
    n = np.arange(win_length)
    window = 0.5 - 0.5 * np.cos(2.0 * math.pi * n / win_length)          # scipy get_window("hann", fftbins=True)
    lpad = (filter_length - win_length) // 2
    window = np.pad(window, (lpad, filter_length - win_length - lpad))    # librosa.util.pad_center
    return torch.FloatTensor(fb[:, None, :]) * torch.from_numpy(window).float()


def stft_magnitude(y: torch.Tensor, basis: torch.Tensor, filter_length: int, hop_length: int) -> torch.Tensor:
    """STFT.transform: reflect-pad filter_length/2 on both sides, strided conv1d with the windowed basis, magnitude.
    y: (B, T) -> (B, filter_length/2 + 1, 1 + T // hop_length)."""
    B, T = y.shape
    x = F.pad(y.view(B, 1, 1, T), (filter_length // 2, filter_length // 2, 0, 0), mode="reflect").squeeze(1)
    ft = F.conv1d(x, basis, stride=hop_length, padding=0)
    cutoff = filter_length // 2 + 1
    return torch.sqrt(ft[:, :cutoff, :] ** 2 + ft[:, cutoff:, :] ** 2)


def mel_spectrogram(y: torch.Tensor, basis: torch.Tensor, mel_basis: torch.Tensor, filter_length: int = 1024,
                    hop_length: int = 160):
    """TacotronSTFT.mel_spectrogram: (log-mel (B, n_mel, frames), log-magnitudes (B, bins, frames), energy (B, frames)),
    log = natural log of the value clamped at 1e-5."""
    assert float(y.min()) >= -1 and float(y.max()) <= 1
    mag = stft_magnitude(y, basis, filter_length, hop_length)
    mel = torch.log(torch.clamp(torch.matmul(mel_basis, mag), min=1e-5))
    return mel, torch.log(torch.clamp(mag, min=1e-5)), torch.norm(mag, dim=1)


def normalize_wav(w: torch.Tensor) -> torch.Tensor:
    w = w - torch.mean(w)
    return w / (torch.max(torch.abs(w)) + 1e-8) * 0.5


def prepare_waveform(w: torch.Tensor, segment_length: int) -> torch.Tensor:
    """read_wav_file after decoding/resampling (torch_tools.py:44-54): remove DC, peak-normalise to 0.5, pad / cut to
    the segment, peak-normalise again to 0.5. w: (T,) -> (1, segment_length)."""
    w = normalize_wav(w)
    if w.numel() > segment_length:
        w = w[:segment_length]
    elif w.numel() < segment_length:
        w = torch.cat([w, torch.zeros(segment_length - w.numel())])
    w = w.unsqueeze(0)
    return 0.5 * (w / torch.max(torch.abs(w)))


def pad_spec(fbank: torch.Tensor, target_length: int) -> torch.Tensor:
    """_pad_spec: (B, frames, ch) zero-padded / cut to target_length frames; an odd channel count drops the last one."""
    B, n, ch = fbank.shape
    if n < target_length:
        fbank = torch.cat([fbank, torch.zeros(B, target_length - n, ch)], 1)
    elif n > target_length:
        fbank = fbank[:, :target_length, :]
    return fbank[:, :, :-1] if ch % 2 else fbank


def wav_to_fbank(waves, basis, mel_basis, target_length: int = 1024, filter_length: int = 1024, hop_length: int = 160):
    """wav_to_fbank without the file reads: list of 16 kHz mono waveforms -> (fbank (B, target_length, n_mel),
    log-magnitudes (B, target_length, bins - 1), waveform (B, target_length * hop))."""
    wav = torch.cat([prepare_waveform(w, target_length * hop_length) for w in waves], 0)
    audio = torch.nan_to_num(torch.clip(wav, -1, 1))
    mel, logmag, _ = mel_spectrogram(audio, basis, mel_basis, filter_length, hop_length)
    return pad_spec(mel.transpose(1, 2), target_length), pad_spec(logmag.transpose(1, 2), target_length), wav
