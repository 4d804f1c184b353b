"""World-size-2 gloo test of the only multi-rank logic on the path: utterance sharding and the
all-reduce of the BER / SNR statistics vector (bench.py uses the same vector over NCCL)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


from image_in_speech_watermarking_b200 import sharding as SH


def _shard(n_utt, rank, world):
    return list(SH.shard_range(n_utt, rank, world))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # data-parallel training step: ONE all-reduce of the flat gradient buffer (train_modelA.sync_gradients)
    from image_in_speech_watermarking_b200 import train_modelA as TM
    flat = torch.full((17655,), float(rank + 1))
    assert TM.sync_gradients(flat) == world and torch.all(flat == sum(range(1, world + 1)))
    g = torch.Generator().manual_seed(0)
    per_utt = torch.rand(10, 7, generator=g, dtype=torch.float64)          # same table on every rank
    mine = _shard(10, rank, world)
    vec = SH.allreduce_stats(SH.stats_vector(per_utt[mine]))
    if rank == 0:
        out.put((vec, SH.stats_vector(per_utt)))
    dist.destroy_process_group()


def test_stats_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got, want = q.get()
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert torch.allclose(got, want, rtol=0, atol=1e-12)
    assert sorted(_shard(10, 0, 2) + _shard(10, 1, 2)) == list(range(10))
    assert _shard(3, 3, 4) == []
