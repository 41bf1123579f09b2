"""BASELINE configs[4]: one long clip (default 1 h @ 44.1 kHz) embedded and detected in the exact
frame-sharded mode on the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        tests/longform_run.py --seconds 3600 --out gpurun_out/r2_longform_8gpu.json

Rank 0 prints / writes ONE JSON record: device-timed embed (400 iterations) and detect, max over ranks;
audio-seconds per second; collectives per iteration and their share of the iteration (from one
instrumented run); and the exactness checks against the whole-clip path on ONE GPU (rank 0): detector
values of the watermarked clip, decoded bits, BER."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def synth_long(seconds, sr):
    """Deterministic long clip: the 30 s synthetic clips 0.. concatenated (each peak <= 0.9)."""
    from aware_b200.synth import synth_clip
    n = int(round(seconds * sr))
    parts, got, i = [], 0, 0
    while got < n:
        c = synth_clip(100 + i % 8, min(30.0, seconds), sr)
        c = np.roll(c, 997 * (i // 8)) * np.float32(1.0 - 0.02 * (i % 5))
        parts.append(c)
        got += len(c)
        i += 1
    return np.concatenate(parts)[:n].astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3600.0)
    ap.add_argument("--sr", type=int, default=44100)
    ap.add_argument("--iters", type=int, default=400)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--out", default="")
    ap.add_argument("--skip-single", action="store_true", help="skip the whole-clip run on rank 0")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from aware_b200.longform import Longform, plan_shards
    from aware_b200.synth import synth_bits
    from aware_b200.utils.models import load
    from aware_b200.utils.watermark import PatternEncoder
    emb, det = load()
    emb.verbose = False
    eng = emb.engine
    sr = args.sr
    x = synth_long(args.seconds, sr)
    bits = synth_bits(1)[0]
    pat = PatternEncoder()(bits)
    lf = Longform(emb)
    dev = eng.device

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn):
        sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        sync()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return r, float(t.item())

    # warm-up (allocations, NCCL channels), then the timed runs
    lf.embed(x, sr, pat, iters=3, precision=args.precision, gather=False)
    lf.detect(x, sr)
    (own, off), ms_embed = timed(lambda: lf.embed(x, sr, pat, iters=args.iters, precision=args.precision, gather=False))
    stats = dict(lf.last_stats)
    plan = plan_shards(len(x), world)
    y = lf.gather(own, plan)
    v_sh, ms_detect = timed(lambda: lf.detect(y, sr))
    with eng._with_precision("fp32"):
        v_sh32 = lf.detect(y, sr)
    # one instrumented run (events around every launch and around every collective): shares of an iteration
    eng.profile(True)
    lf.embed(x, sr, pat, iters=20, precision=args.precision, gather=False)
    torch.cuda.synchronize()
    eng.profile(False)
    eng.profile_read()
    tl = eng.profile_read_named()
    tot = sum(v[1] for v in tl.values())
    comm_ms = sum(v[1] for k, v in tl.items() if k in ("allreduce", "allgather"))
    if world > 1:
        dist.barrier()

    rec = None
    if rank == 0:
        rec = {"workload": "BASELINE configs[4]: 1 clip x %g s @ %d Hz, embed (%d NAdam it, %s loop) + detect, "
                           "frames sharded over %d GPU(s), exact mode" % (args.seconds, sr, args.iters, args.precision, world),
               "n_gpus": world, "frames": plan[0]["T"], "frames_per_rank": [p["f1"] - p["f0"] for p in plan],
               "embed_ms": ms_embed, "embed_audio_s_per_s": args.seconds / (ms_embed / 1e3),
               "embed_ms_per_iteration": ms_embed / max(args.iters, 1),
               "detect_ms": ms_detect, "detect_audio_s_per_s": args.seconds / (ms_detect / 1e3),
               "collectives": {"allreduces_total": stats["allreduces"], "allgathers_total": stats["allgathers"],
                               "allreduces_per_iteration": (stats["allreduces"] - 2) / max(args.iters, 1),
                               "allgathers_per_iteration": (stats["allgathers"] - 1) / max(args.iters, 1),
                               "share_of_iteration_instrumented": comm_ms / tot if tot else None,
                               "ms_per_iteration_instrumented": comm_ms / 20.0,
                               "note": "small latency-bound NCCL collectives on NVLink (<= 16 KB all-reduce, 5 KB/rank "
                                       "all-gather); measured with CUDA events around each callback on the compute stream"},
               "ber_percent_sharded_detect": 100.0 * float(np.mean((v_sh > 0).astype(np.int32) != bits)),
               "min_margin": float(np.abs(v_sh).min())}
        if not args.skip_single:
            yd = torch.from_numpy(y[None]).to(dev)
            t0 = time.perf_counter()
            with eng._with_precision("fp32"):
                v1 = eng.detect(yd, sr).cpu().numpy()[0]          # whole clip on ONE GPU, exact GEMMs
            rec["single_gpu_detect_fp32_s"] = time.perf_counter() - t0
            rec["values_max_abs_diff_vs_single_gpu_fp32"] = float(np.abs(v_sh32 - v1).max())
            rec["bits_identical_vs_single_gpu"] = bool(np.array_equal(v_sh32 > 0, v1 > 0) and np.array_equal(v_sh > 0, v1 > 0))
            xd = torch.from_numpy(x[None]).to(dev)
            pt = torch.from_numpy(pat[None])
            eng.embed(xd, sr, pt, iters=3, precision=args.precision)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            y1 = eng.embed(xd, sr, pt, iters=args.iters, precision=args.precision)
            b.record()
            torch.cuda.synchronize()
            rec["single_gpu_embed_ms"] = a.elapsed_time(b)
            rec["speedup_vs_single_gpu"] = a.elapsed_time(b) / ms_embed
            v_cross = eng.detect(y1, sr).cpu().numpy()[0]
            rec["single_gpu_embed_ber_percent"] = 100.0 * float(np.mean((v_cross > 0).astype(np.int32) != bits))
            from aware_b200.metrics.audio import SNR
            L = len(y)
            rec["snr_db_sharded"] = SNR()(y, x[:L])
            rec["snr_db_single_gpu"] = SNR()(y1.cpu().numpy()[0], x[:L])
        print(json.dumps(rec))
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            open(args.out, "w").write(json.dumps(rec) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
