"""CPU, world_size 2 over gloo: the N>1 plumbing (prompt sharding, weight broadcast, waveform gather, max-over-ranks)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tango_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prompts = [f"p{i}" for i in range(5)]
        lo, hi = parallel.shard_range(len(prompts), rank, world)
        mine = prompts[lo:hi]
        # weights: rank 0 has the real values, rank 1 garbage; after the broadcast both agree
        sd = {"a.weight": torch.full((3, 4), float(rank + 1)), "b.bias": torch.arange(5, dtype=torch.float32) * (rank + 1)}
        sd = parallel.broadcast_state_dict(sd, src=0)
        ok_bcast = bool((sd["a.weight"] == 1).all() and torch.equal(sd["b.bias"], torch.arange(5, dtype=torch.float32)))
        # RNG contract: same seed on every rank, full-batch draw, keep my rows == the rows a single
        # process would use; per-sample generator lists only consume the local generators
        from tango_b200.pipeline import AudioDiffusion
        g = torch.Generator().manual_seed(7)
        part = AudioDiffusion.randn_rows((hi - lo, 8), g, "cpu", rows=(lo, hi, 5))
        g2 = torch.Generator().manual_seed(7)
        nxt = AudioDiffusion.randn_rows((hi - lo, 8), g, "cpu", rows=(lo, hi, 5))      # second draw of the stream
        full = torch.randn(5, 8, generator=g2)
        full2 = torch.randn(5, 8, generator=g2)
        ok_rng = torch.equal(part, full[lo:hi]) and torch.equal(nxt, full2[lo:hi])
        gl = [torch.Generator().manual_seed(100 + i) for i in range(5)]
        ps = AudioDiffusion.randn_rows((hi - lo, 8), gl, "cpu", rows=(lo, hi, 5))
        want = torch.cat([torch.randn(1, 8, generator=torch.Generator().manual_seed(100 + i)) for i in range(lo, hi)])
        ok_rng = ok_rng and torch.equal(ps, want)
        waves = np.stack([np.full(4, int(p[1:]), dtype=np.int16) for p in mine])
        out = parallel.allgather_waves(waves)
        mx = parallel.max_over_ranks(float(rank + 1))
        sm = parallel.sum_over_ranks(float(rank + 1))
        q.put((rank, mine, ok_bcast, ok_rng, [int(w[0]) for w in out], mx, sm))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    for n in (0, 1, 5, 8, 64):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, mine0, b0, n0, out0, mx0, sm0), (r1, mine1, b1, n1, out1, mx1, sm1) = res
    assert mine0 == ["p0", "p1", "p2"] and mine1 == ["p3", "p4"]
    assert b0 and b1
    assert n0 and n1
    assert out0 == out1 == [0, 1, 2, 3, 4]  # every rank holds every waveform, in prompt order
    assert mx0 == mx1 == 2.0 and sm0 == sm1 == 3.0
