"""GPU: state handling of AudioDiffusion.inference / Tango.generate_for_batch around the captured CUDA graphs, and the
seed / sharding contract (diffusers torch_utils.py:29-70 `randn_tensor`)."""
import os
import socket

import numpy as np
import pytest
import torch

from tango_b200 import parallel, synth
from tango_b200.pipeline import AudioDiffusion, Tango
from tango_b200.schedulers import DDPMScheduler

pytestmark = pytest.mark.gpu
SHAPE = (32, 16)


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _model(cuda, seed=0, graph=True, precision="split"):
    cfg = synth.TINY_UNET_CONFIG
    m = AudioDiffusion(unet_config=cfg, precision=precision, use_cuda_graph=graph).to(cuda)
    m.unet.load_state_dict(synth.synth_state_dict(synth.unet_param_shapes(cfg), seed=seed))
    return m


def test_graph_cache_distinguishes_cfg_from_no_cfg_at_equal_unet_batch(cuda):
    """1 prompt at guidance 3 and 2 prompts at guidance 1 both give UNet batch 2, but only the first may share the CFG
    prefix: alternating them must replay the right graph each time (compared with eager, graph-free runs)."""
    cfg = synth.TINY_UNET_CONFIG
    e1, m1 = synth.synth_conditioning(1, 10, cfg["cross_attention_dim"], seed=5, masked_tail=3)       # [uncond; cond]
    e2, m2 = synth.synth_conditioning(2, 10, cfg["cross_attention_dim"], seed=6, masked_tail=2)
    e2, m2 = e2[2:], m2[2:]                                                                            # cond only
    l1, n1 = synth.synth_noise(1, 3, shape=(8, *SHAPE), seed=1)
    l2, n2 = synth.synth_noise(2, 3, shape=(8, *SHAPE), seed=2)
    mg, me = _model(cuda, graph=True), _model(cuda, graph=False)

    def run(m, which):
        if which == "cfg":
            return m.inference(["a"], DDPMScheduler.from_pretrained(), 3, 3.0, prompt_embeds=e1, boolean_prompt_mask=m1,
                               latents=l1, noises=n1, latent_shape=SHAPE).clone()
        return m.inference(["a", "b"], DDPMScheduler.from_pretrained(), 3, 1.0, prompt_embeds=e2, boolean_prompt_mask=m2,
                           latents=l2, noises=n2, latent_shape=SHAPE).clone()

    want = {k: run(me, k) for k in ("cfg", "plain")}
    for k in ("cfg", "plain", "cfg", "plain"):
        got = run(mg, k)
        assert rel(got, want[k]) < 2e-4, k     # graph replay vs eager: GroupNorm atomics round-off only
    assert len(mg._state) == 2
    # sample 1 of the no-CFG batch differs from sample 0 (a stale CFG graph would have duplicated the first half)
    assert rel(want["plain"][1], want["plain"][0]) > 1e-2


def test_reloading_weights_or_moving_drops_the_captured_graphs(cuda):
    cfg = synth.TINY_UNET_CONFIG
    e, mk = synth.synth_conditioning(1, 10, cfg["cross_attention_dim"], seed=5, masked_tail=3)
    l0, ns = synth.synth_noise(1, 2, shape=(8, *SHAPE), seed=4)

    def run(m):
        return m.inference(["a"], DDPMScheduler.from_pretrained(), 2, 3.0, prompt_embeds=e, boolean_prompt_mask=mk,
                           latents=l0, noises=ns, latent_shape=SHAPE).clone()

    m = _model(cuda, seed=0)
    a0 = run(m)
    gen0 = m.unet.pack_generation
    m.unet.load_state_dict(synth.synth_state_dict(synth.unet_param_shapes(cfg), seed=1))     # straight on the UNet
    a1 = run(m)
    assert m.unet.pack_generation == gen0 + 1
    fresh = run(_model(cuda, seed=1))
    assert rel(a1, fresh) < 2e-4 and rel(a1, a0) > 1e-2
    m.load_state_dict({"unet." + k: v for k, v in synth.synth_state_dict(synth.unet_param_shapes(cfg), seed=0).items()})
    assert len(m._state) == 0                                                               # dropped eagerly too
    assert rel(run(m), a0) < 2e-4


def test_masked_text_length_is_bucketed(cuda):
    """Prompts of 10 and 12 tokens share one padded length (and one captured graph); padding changes nothing."""
    cfg = synth.TINY_UNET_CONFIG
    m = _model(cuda)
    me = _model(cuda, graph=False)
    me.LK_BUCKET = 1                       # eager run at the raw length
    for Lk in (10, 12):
        e, mk = synth.synth_conditioning(1, Lk, cfg["cross_attention_dim"], seed=Lk, masked_tail=3)
        l0, ns = synth.synth_noise(1, 2, shape=(8, *SHAPE), seed=9)
        kw = dict(prompt_embeds=e, boolean_prompt_mask=mk, latents=l0, noises=ns, latent_shape=SHAPE)
        a = m.inference(["a"], DDPMScheduler.from_pretrained(), 2, 3.0, **kw).clone()
        b = me.inference(["a"], DDPMScheduler.from_pretrained(), 2, 3.0, **kw).clone()
        assert rel(a, b) < 2e-4
    assert len(m._state) == 1


def _tango(cuda):
    return Tango.from_synthetic(unet_config=synth.TINY_UNET_CONFIG, device=cuda, precision="split")


def test_per_sample_generators_make_noise_independent_of_batching(cuda):
    """torch_utils.py:60-66: a list of generators draws every sample on its own -> chunking by batch_size 4 or 2 (and
    therefore any sharding) yields the same waveforms."""
    t = _tango(cuda)
    prompts = [f"prompt number {i}" for i in range(4)]

    def gens():
        return [torch.Generator(device=cuda).manual_seed(1000 + i) for i in range(4)]

    a = t.generate_for_batch(prompts, steps=3, guidance=3, batch_size=4, latent_shape=SHAPE, generator=gens())
    b = t.generate_for_batch(prompts, steps=3, guidance=3, batch_size=2, latent_shape=SHAPE, generator=gens())
    assert len(a) == len(b) == 4
    for x, y in zip(a, b):
        assert np.abs(x.astype(np.int32) - y.astype(np.int32)).max() <= 2       # same noise; GN atomics round-off only
    with pytest.raises(ValueError):
        t.generate_for_batch(prompts, steps=1, guidance=3, batch_size=4, latent_shape=SHAPE, generator=gens()[:3])


def test_sharded_chunks_reproduce_the_single_gpu_run(cuda, monkeypatch):
    """The prompt-shard path on ONE device: ranks 0 and 1 of a world of 2 are run one after the other (same seed on
    both, as the CLI sets it), their rows concatenated, and compared with the unsharded run on that seed."""
    t = _tango(cuda)
    prompts = [f"prompt number {i}" for i in range(5)]           # chunks of 4 + 1: the second chunk leaves rank 1 empty

    def run(world, r):
        monkeypatch.setattr(parallel, "world_size", lambda: world)
        monkeypatch.setattr(parallel, "rank", lambda: r)
        monkeypatch.setattr(parallel, "allgather_waves", lambda w, dev=None: w)   # keep the local block
        g = torch.Generator(device=cuda).manual_seed(77)
        return t.generate_for_batch(prompts, steps=3, guidance=3, batch_size=4, latent_shape=SHAPE, generator=g,
                                    shard=world > 1)

    full = run(1, 0)
    r0, r1 = run(2, 0), run(2, 1)
    assert len(full) == 5 and len(r0) == 3 and len(r1) == 2      # rank 0: prompts 0,1,4; rank 1: prompts 2,3
    for got, want in zip([r0[0], r0[1], r1[0], r1[1], r0[2]], full):
        assert np.abs(got.astype(np.int32) - want.astype(np.int32)).max() <= 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        t = Tango.from_synthetic(unet_config=synth.TINY_UNET_CONFIG, device=dev, precision="split")
        g = torch.Generator(device=dev).manual_seed(77)
        prompts = [f"prompt number {i}" for i in range(5)]
        out = t.generate_for_batch(prompts, steps=3, guidance=3, batch_size=4, latent_shape=SHAPE, generator=g, shard=True)
        q.put((rank, [o.tolist() for o in out]))
    finally:
        dist.destroy_process_group()


def test_two_rank_nccl_run_equals_one_rank_run(cuda):
    """Two processes, two GPUs, NCCL all-gather of the int16 waveforms: every rank ends with all five waveforms and they
    equal the one-GPU run on the same seed (needs 2 GPUs: `gpurun --gpus 2`)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    t = _tango(cuda)
    g = torch.Generator(device=cuda).manual_seed(77)
    prompts = [f"prompt number {i}" for i in range(5)]
    want = t.generate_for_batch(prompts, steps=3, guidance=3, batch_size=4, latent_shape=SHAPE, generator=g)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in (0, 1):
        assert len(res[r]) == 5
        for got, w in zip(res[r], want):
            assert np.abs(np.asarray(got, dtype=np.int32) - w.astype(np.int32)).max() <= 2


def test_cli_two_ranks_reproduces_single_rank_wavs(cuda, tmp_path):
    """`torchrun --nproc-per-node 2 -m tango_b200.cli --seed S` writes the wav files of the one-GPU run with the same
    seed (every chunk of prompts is split over the ranks, each rank keeps its rows of the shared noise stream) - needs
    2 GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import json
    import subprocess
    import sys
    import wave
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    man = tmp_path / "prompts.json"
    man.write_text("\n".join(json.dumps({"captions": c}) for c in ["a dog barking", "rain", "church bells", "a train", "wind"]))
    common = ["--checkpoint", "synthetic:tiny", "--test_file", str(man), "--num_steps", "3", "--batch_size", "4",
              "--latent_h", "32", "--seed", "11", "--precision", "split"]
    env = dict(os.environ, PYTHONPATH=root)
    r1 = subprocess.run([sys.executable, "-m", "tango_b200.cli", *common, "--output_root", str(tmp_path / "one"), "--exp_id", "a"],
                        cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0, r1.stderr[-2000:]
    r2 = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                         "127.0.0.1", "--master-port", str(_free_port()), "-m", "tango_b200.cli", *common, "--output_root",
                         str(tmp_path / "two"), "--exp_id", "a"], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r2.returncode == 0, r2.stderr[-2000:]
    for j in range(5):
        frames = []
        for d in ("one", "two"):
            with wave.open(str(tmp_path / d / "a_steps_3_guidance_3" / f"output_{j}.wav")) as w:
                frames.append(np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.int32))
        assert frames[0].shape == frames[1].shape and np.abs(frames[0] - frames[1]).max() <= 2, j
    line = (tmp_path / "two" / "tango_checkpoint_summary.jsonl").read_text().strip()
    assert json.loads(line)["n_gpus"] == 2
