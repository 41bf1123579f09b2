"""Frame-sharded long-form mode (aware_b200/longform.py, aw_*_sharded).

CPU (gloo, world_size 2): the shard plan and the two collectives the C library calls back for,
driven exactly as the library drives them (offsets into the arena, float64 sum / int64 max, the
halo all-gather layout).  GPU: the sharded path against the whole-clip path -- with one rank, and
with two and three ranks that SHARE the one GPU and meet only in host-side gloo collectives."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import aware_oracle as O
from aware_b200 import _lib
from aware_b200.longform import HALO, Comm, plan_shards


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_plan_covers_the_clip_with_even_boundaries_and_halos():
    for n in (44100 * 60, 44100 * 3600, 16000 * 95 + 77, 256 * 400, 256 * 400 + 255):
        T = 1 + n // 256
        for world in (1, 2, 3, 8):
            plan = plan_shards(n, world)
            assert plan[0]["f0"] == 0 and plan[-1]["f1"] == T
            assert all(a["f1"] == b["f0"] for a, b in zip(plan, plan[1:]))
            assert all(p["f0"] % 2 == 0 for p in plan)                       # AvgPool pairs stay on one rank
            assert sum((p["f1"] - p["f0"]) // 2 for p in plan) == T // 2       # pooled frames add up to T'
            assert plan[0]["out_lo"] == 0 and plan[-1]["out_hi"] == 256 * (T - 1)
            assert all(a["out_hi"] == b["out_lo"] for a, b in zip(plan, plan[1:]))
            for p in plan:
                assert p["own_lo"] == (HALO if p["rank"] > 0 else 0)
                assert p["e1"] - p["f1"] == (HALO if p["rank"] < world - 1 else 0)
                assert p["f1"] - p["f0"] >= 2 * HALO
                assert 0 <= p["s0"] and p["s0"] + p["n_seg"] <= n
                assert 1 + p["n_seg"] // 256 == p["e1"] - p["e0"]              # the segment frames like a clip
    with pytest.raises(ValueError):
        plan_shards(256 * 20, 2)


def _collective_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = Comm(torch.device("cpu"))
    a = comm.arena
    # (1) float64 sum at offset 0, as sh_reduce_sum does
    a[:8 * 256].view(torch.float64).copy_(torch.arange(256, dtype=torch.float64) * (rank + 1))
    assert comm.struct.allreduce(None, 0, 256, _lib.COMM_F64, _lib.COMM_SUM, None) == 0
    s = a[:8 * 256].view(torch.float64).clone()
    # (2) packed peak word: int64 max == the larger (|y| bits, index) word
    word = (int(np.float32(0.25 + 0.5 * rank).view(np.uint32)) << 32) | (1000 + rank)
    a[:8].view(torch.int64)[0] = word
    assert comm.struct.allreduce(None, 0, 1, _lib.COMM_I64, _lib.COMM_MAX, None) == 0
    w = int(a[:8].view(torch.int64)[0])
    # (3) halo all-gather: [2][H][nb] floats per rank at the send offset, rank-major at the recv offset
    nb = 81
    per = HALO * nb
    send = torch.arange(2 * per, dtype=torch.float32) + 10000.0 * rank
    a[32768:32768 + 8 * per].view(torch.float32).copy_(send)
    assert comm.struct.allgather(None, 32768, 65536, 8 * per, None) == 0
    recv = a[65536:65536 + world * 8 * per].view(torch.float32).view(world, 2, HALO, nb).clone()
    if rank == 0:
        ret["sum"] = s.tolist()
        ret["word"] = w
        ret["recv"] = recv.numpy()
        ret["counts"] = (comm.n_allreduce, comm.n_allgather)
    dist.destroy_process_group()


def test_callbacks_reduce_and_gather_world2():
    """The halo / statistic exchange as the C library issues it, on 2 gloo ranks (CPU arena)."""
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_collective_worker, args=(2, _port(), ret), nprocs=2, join=True)
        assert ret["sum"] == [3.0 * i for i in range(256)]
        assert ret["word"] == (int(np.float32(0.75).view(np.uint32)) << 32) | 1001
        recv = ret["recv"]
        # rank 0's right halo comes from rank 1's FIRST H own frames = recv[1][0]
        np.testing.assert_array_equal(recv[1, 0].ravel(), np.arange(HALO * 81, dtype=np.float32) + 10000.0)
        np.testing.assert_array_equal(recv[0, 1].ravel(), np.arange(HALO * 81, 2 * HALO * 81, dtype=np.float32))
        assert ret["counts"] == (2, 1)


# ------------------------------------------------------------------------------------ GPU
def _clip(secs, sr=44100, idx=5):
    return O.synth_clip(idx, secs, sr)


@pytest.mark.gpu
def test_sharded_path_with_one_rank_equals_the_batch_path():
    """world = 1 exercises the global-T plumbing and the own-range logic without any halo: detector
    values and a 3-iteration embed equal the ordinary batch path."""
    from aware_b200.longform import Longform
    from aware_b200.utils.models import load
    emb, det = load()
    emb.verbose = False
    eng = emb.engine
    sr = 44100
    x = _clip(4.0)
    lf = Longform(emb)
    eng.set_precision("fp32")
    try:
        v_ref = eng.detect(torch.from_numpy(x[None]).cuda(), sr).cpu().numpy()[0]
        v = lf.detect(x, sr)
    finally:
        eng.set_precision("tf32")
    assert np.abs(v - v_ref).max() <= 1e-6
    np.testing.assert_allclose(v, O.detect(x, sr), atol=1e-5)
    pat = O.encode_bits(O.synth_bits(8)[5])
    eng.set_tc_spectral(False)          # the sharded mode runs the fp32 FFT spectral kernels: compare like with like
    for prec in ("fp32", "fp16"):
        y_ref = eng.embed(torch.from_numpy(x[None]).cuda(), sr, torch.from_numpy(pat[None]), iters=3,
                          precision=prec).cpu().numpy()[0]
        y = lf.embed(x, sr, pat, iters=3, precision=prec)
        assert y.shape == y_ref.shape
        assert np.abs(y - y_ref).max() <= 1e-5, prec
    eng.set_tc_spectral(True)
    assert lf.last_stats["allgathers"] == 0 and lf.last_stats["allreduces"] == 2 + 3 * 12


def _gpu_worker(rank, world, port, secs, iters, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)                      # every rank on the ONE GPU: they meet only in gloo collectives
    from aware_b200.longform import Longform
    from aware_b200.utils.models import load
    emb, det = load()
    emb.verbose = False
    eng = emb.engine
    sr = 44100
    x = _clip(secs)
    pat = O.encode_bits(O.synth_bits(8)[5])
    lf = Longform(emb)
    out = {}
    with eng._with_precision("fp32"):
        out["v_fp32"] = lf.detect(x, sr)
    out["v_tf32"] = lf.detect(x, sr, precision="tf32")
    out["y1_fp32"] = lf.embed(x, sr, pat, iters=1, precision="fp32")
    for prec in ("fp32", "fp16"):
        y, losses = lf.embed(x, sr, pat, iters=iters, precision=prec, return_losses=True)
        out["y_" + prec], out["loss_" + prec] = y, losses
    out["stats"] = dict(lf.last_stats)
    yl = lf.embed(x, sr, pat, iters=60, precision="fp16")
    out["v_after"] = lf.detect(yl, sr)
    out["y_long"] = yl
    if rank == 0:
        # the whole-clip path on the same GPU, same process (fp32 FFT spectral kernels, as the sharded mode)
        eng.set_tc_spectral(False)
        xd = torch.from_numpy(x[None]).cuda()
        with eng._with_precision("fp32"):
            out["ref_v_fp32"] = eng.detect(xd, sr).cpu().numpy()[0]
        out["ref_y1_fp32"] = eng.embed(xd, sr, torch.from_numpy(pat[None]), iters=1, precision="fp32").cpu().numpy()[0]
        for prec in ("fp32", "fp16"):
            yr, _, lr = eng.embed(xd, sr, torch.from_numpy(pat[None]), iters=iters, precision=prec, return_losses=True)
            out["ref_y_" + prec], out["ref_loss_" + prec] = yr.cpu().numpy()[0], lr.cpu().numpy()[:, 0]
        ret.update(out)
    else:
        ret["v_fp32_r%d" % rank] = out["v_fp32"]
        ret["y_fp32_r%d" % rank] = out["y_fp32"]
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world,secs", [(2, 6.0), (3, 9.5)])
def test_sharded_ranks_reproduce_the_whole_clip_run(world, secs):
    """2 and 3 ranks (sharing the one GPU, gloo collectives): halo + statistic exchange.  Detector values
    equal the whole-clip run to 1e-6 and decode identically; every rank holds the same result.  Embed: the
    only difference to the whole-clip run is the summation order of the float64 statistics, so the first
    loss is identical and ONE step gives the same waveform (1e-4 on >= 99.5 % of the samples: a NAdam first
    step is sign-like, a gradient within rounding of 0 can flip); after that the iteration is chaotic
    (reference vs reference 6e-7 after 1 step, 1.8e-3 after 3), so three steps are gated
    loosely; a 60-iteration sharded embed is decoded correctly by the sharded detector and the CPU oracle."""
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_gpu_worker, args=(world, _port(), secs, 3, ret), nprocs=world, join=True)
        r = dict(ret)
    assert np.abs(r["v_fp32"] - r["ref_v_fp32"]).max() <= 1e-6
    assert np.array_equal(r["v_fp32"] > 0, r["ref_v_fp32"] > 0)
    assert np.abs(r["v_tf32"] - r["ref_v_fp32"]).max() <= 1e-3
    for k in range(1, world):
        np.testing.assert_array_equal(r["v_fp32_r%d" % k], r["v_fp32"])          # identical on every rank
        np.testing.assert_array_equal(r["y_fp32_r%d" % k], r["y_fp32"])
    d1 = np.abs(r["y1_fp32"] - r["ref_y1_fp32"])
    print("world %d: 1-iteration waveform max diff %.2e, %.4f %% within 1e-4" % (world, d1.max(), 100 * (d1 <= 1e-4).mean()))
    assert r["y1_fp32"].shape == r["ref_y1_fp32"].shape
    assert (d1 <= 1e-4).mean() >= 0.995 and d1.max() <= 3e-3
    for prec, tol0, tol_l in (("fp32", 1e-6, 1e-3), ("fp16", 1e-3, 1e-2)):
        dl = np.abs(r["loss_" + prec][:3] - r["ref_loss_" + prec][:3])
        assert dl[0] <= tol0 and dl.max() <= tol_l, (prec, dl)
        d = np.abs(r["y_" + prec] - r["ref_y_" + prec])
        print("world %d %s: 3-iteration waveform max diff %.2e, %.4f %% within 1e-4" % (
            world, prec, d.max(), 100 * (d <= 1e-4).mean()))
        if prec == "fp32":
            assert (d <= 1e-3).mean() >= 0.97 and d.max() <= 2e-2
        else:       # 16-bit activations turn a last-bit difference of a statistic into 1e-3 steps: faster divergence
            assert d.mean() <= 3e-3 and d.max() <= 3e-2
    st = r["stats"]
    assert st["allreduces"] == 2 + 3 * 12 and st["allgathers"] == 2 * 3 + 1     # of the 3-iteration fp16 run
    bits = O.synth_bits(8)[5]
    np.testing.assert_array_equal((r["v_after"] > 0).astype(np.int32), bits)
    np.testing.assert_array_equal(O.detect_watermark(r["y_long"], 44100), bits)   # the CPU oracle agrees
