"""One eager pass of everything on the path OUTSIDE the UNet forward, between cudaProfilerStart/Stop, for the ncu launch
list of the decoder-side kernels (profiles/): fused CFG + scheduler step, VAE decoder (GroupNorm, 3x3 convs, the
single-head d = 512 attention), HiFi-GAN (conv1d stacks, ConvTranspose overlap-add, tanh -> int16), the FLAN-T5
encoder front-end (rmsnorm, gather, relative-position attention) and the TacotronSTFT mel front-end.
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv python tools/ncu_decode.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tango_b200 import lib as L, synth
from tango_b200.schedulers import DDIMScheduler
from tango_b200.stft import TacotronSTFT
from tango_b200.t5 import T5EncoderModel
from tango_b200.vae import AutoencoderKL

dev = torch.device("cuda:0")
B = 2
vae = AutoencoderKL(**synth.VAE_CONFIG, precision="bf16").to(dev)
vae.load_state_dict(synth.synth_state_dict(synth.vae_decoder_param_shapes(), 0))
z = torch.randn(B * 256 * 16, 8, device=dev)
t5cfg = dict(synth.FLAN_T5_LARGE_CONFIG, num_layers=2)
t5 = T5EncoderModel.from_config(t5cfg, precision="bf16").to(dev)
t5.load_state_dict(synth.synth_state_dict(synth.t5_encoder_param_shapes(t5cfg), 0))
ids = torch.randint(2, 1000, (8, 64), device=dev)
am = torch.ones(8, 64, dtype=torch.long, device=dev)
stft = TacotronSTFT(**synth.STFT_CONFIG).to(dev)
wav = (torch.rand(2, 163840, device=dev) - 0.5)
sch = DDIMScheduler.from_pretrained()
sch.set_timesteps(200, device=dev)
coef = sch.coefficient_table(dev)
mo = torch.randn(16 * 4096, 8, device=dev)
sample = torch.randn(8, 8, 256, 16, device=dev)
x_in = torch.zeros(16 * 4096, 8, device=dev, dtype=torch.bfloat16)


def once():
    L.sched_step(mo, True, 3.0, sample, None, coef[5], sample, x_in, B=8, Cc=8, HW=4096)
    vae.decode_rows_to_waveform(z, B, 256, 16, use_cuda_graph=False)
    t5(ids, am)
    stft.mel_spectrogram(wav)


once()
torch.cuda.synchronize()
torch.cuda.profiler.start()
once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
