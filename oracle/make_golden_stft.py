"""Pin oracle/stft.py against the UNMODIFIED reference STFT / TacotronSTFT / torch_tools arithmetic and write
tests/golden/tiny_stft.npz (TEST INFRASTRUCTURE ONLY; build container only).

librosa is not installed, so the reference modules are instantiated without running their librosa-dependent
constructors: the buffers they would compute (`forward_basis`, `mel_basis`) are supplied — exactly what loading
`pytorch_model_stft.bin` does in tango.py:19,26 — and every forward-path function is the reference's own.

    python -m oracle.make_golden_stft
"""
import json
import os
import sys
import types

import numpy as np
import torch

from oracle import refshim
from oracle import stft as ostft

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
FL, HOP, WIN, NMEL = 256, 40, 256, 16          # a quarter-size stft_config (1024 / 160 / 1024 / 64)


def main():
    torch.set_grad_enabled(False)
    refshim.install()
    from audioldm.audio import stft as rstft
    for nm in ("torchaudio", "tools.mix"):
        if nm not in sys.modules:
            sys.modules[nm] = types.ModuleType(nm)
    sys.modules["tools.mix"].mix = None
    from tools import torch_tools as rtt

    basis = ostft.forward_basis(FL, WIN)
    # the reference's own basis construction, with the two librosa helpers written out (pad_center is a centred np.pad)
    from scipy.signal import get_window
    fb = np.fft.fft(np.eye(FL))
    cutoff = FL // 2 + 1
    fb = np.vstack([np.real(fb[:cutoff, :]), np.imag(fb[:cutoff, :])])
    ref_basis = torch.FloatTensor(fb[:, None, :]) * torch.from_numpy(get_window("hann", WIN, fftbins=True)).float()
    assert torch.equal(basis, ref_basis.float())

    g = torch.Generator().manual_seed(8)
    mel_basis = torch.rand(NMEL, cutoff, generator=g) * (torch.rand(NMEL, cutoff, generator=g) > 0.8).float()

    fn = rstft.TacotronSTFT.__new__(rstft.TacotronSTFT)
    torch.nn.Module.__init__(fn)
    fn.n_mel_channels, fn.sampling_rate = NMEL, 16000
    st = rstft.STFT.__new__(rstft.STFT)
    torch.nn.Module.__init__(st)
    st.filter_length, st.hop_length, st.win_length, st.window = FL, HOP, WIN, "hann"
    st.register_buffer("forward_basis", basis.clone())
    fn.stft_fn = st
    fn.register_buffer("mel_basis", mel_basis.clone())

    waves = [torch.randn(n, generator=g) * 0.3 + 0.05 for n in (5000, 7300)]
    target = 160                                         # frames; segment = target * HOP samples
    # the reference pipeline after file decoding (torch_tools.py:44-77 with read_wav_file's torchaudio part removed)
    def ref_prepare(w):
        w = rtt.normalize_wav(w)
        w = rtt.pad_wav(w, target * HOP).unsqueeze(0)
        w = w / torch.max(torch.abs(w))
        return 0.5 * w
    wav_ref = torch.cat([ref_prepare(w) for w in waves], 0)
    fb_ref, lm_ref, _ = rtt.get_mel_from_wav(wav_ref, fn)
    fb_ref, lm_ref = rtt._pad_spec(fb_ref.transpose(1, 2), target), rtt._pad_spec(lm_ref.transpose(1, 2), target)

    fb_o, lm_o, wav_o = ostft.wav_to_fbank(waves, basis, mel_basis, target_length=target, filter_length=FL, hop_length=HOP)
    d = [float((a - b).abs().max()) for a, b in ((wav_ref, wav_o), (fb_ref, fb_o), (lm_ref, lm_o))]
    print(f"STFT front-end: frames {fb_ref.shape[1]}, oracle vs reference: waveform {d[0]:.3e} fbank {d[1]:.3e} log-mag {d[2]:.3e}")
    assert max(d) < 1e-5
    np.savez_compressed(os.path.join(GOLD, "tiny_stft.npz"), wave0=waves[0].numpy(), wave1=waves[1].numpy(),
                        mel_basis=mel_basis.numpy(), fbank=fb_ref.numpy(), log_mag=lm_ref.numpy(), wav=wav_ref.numpy(),
                        cfg=np.array([FL, HOP, WIN, NMEL, target]))
    mp = os.path.join(GOLD, "MANIFEST.json")
    man = json.load(open(mp))
    man["checks"]["stft_frontend"] = {"waveform": d[0], "fbank": d[1], "log_magnitudes": d[2]}
    json.dump(man, open(mp, "w"), indent=1)
    print("wrote", os.path.join(GOLD, "tiny_stft.npz"))


if __name__ == "__main__":
    main()
