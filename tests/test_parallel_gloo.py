"""world_size-2 gloo test of the multi-GPU host logic: shard ranges cover the job and
the counter all-reduce reproduces the single-process BER."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import aware_oracle as O
from aware_b200.parallel import allreduce_counters, ber_percent, shard_range


def test_shard_range_partitions():
    for n in (1, 7, 256, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_clips, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    ref = O.synth_bits(n_clips)
    dec = ref ^ (rng.random(ref.shape) < 0.1)          # same "decoded" bits on every rank
    lo, hi = shard_range(n_clips, rank, world)
    errs = int((ref[lo:hi] != dec[lo:hi]).sum())
    counters = torch.tensor([errs, (hi - lo) * 20, hi - lo], dtype=torch.int64)
    sums = torch.tensor([float(hi - lo)], dtype=torch.float64)
    allreduce_counters(counters, sums)
    if rank == 0:
        ret["counters"] = counters.tolist()
        ret["sums"] = sums.tolist()
        ret["expected_errs"] = int((ref != dec).sum())
    dist.destroy_process_group()


def test_counter_allreduce_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n_clips = 37
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, n_clips, ret), nprocs=2, join=True)
        assert ret["counters"] == [ret["expected_errs"], n_clips * 20, n_clips]
        assert ret["sums"] == [float(n_clips)]
        ber = ber_percent(torch.tensor(ret["counters"]))
        assert abs(ber - 100.0 * ret["expected_errs"] / (n_clips * 20)) < 1e-12
