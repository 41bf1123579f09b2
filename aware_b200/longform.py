"""Exact frame-sharded mode for ONE long clip across the GPUs of a box (BASELINE
configs[4]: "streaming long-form audio, 1 h clip, embed+detect on 8 B200").

The reference processes a clip whole and every statistic it takes is whole-clip: the peak
normalisers (utils/audio/waveform.py:19), InstanceNorm over all frames and the global
standardisation (detection/multibit_detector_net.py:50,126, modules/globalStandardize.py:17-19), the
global average pool (modules/BRH.py:18).  Cutting the clip into independent segments would change
the result.  Here the FRAMES are split over the ranks instead:

  * rank r owns global frames [f0, f1) (even boundaries, so AvgPool1d(2,2) pairs never straddle two
    ranks) and holds a segment with 8 halo frames per inner side;
  * every per-clip sum / maximum (peak of x and of y, mel channel sums, the four InstanceNorm
    statistics forward and backward, the head's sums, the front-end adjoint sums, the Euler sum of the
    normaliser sub-gradient) is all-reduced between the kernel that produces the partials and the
    kernel that consumes them -- 13 small all-reduces per optimisation iteration;
  * the halo frames of the coefficients c and of the spectral gradient dA are refreshed from their
    owners by two small all-gathers per iteration (the fused STFT/iSTFT kernels couple frames t-3..t+3).

The C library drives the loop (aw_embed_sharded / aw_detect_sharded) and calls back into `Comm` for
the collectives, which run through torch.distributed: NCCL over NVLink on the box, gloo in the tests
(several ranks may then share one GPU: the ranks only meet in host-side collectives).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .engine import N_BITS, _ptr, _stream

HALO = 8          # AW_HALO in csrc/api.cu
HOP = 256


def plan_shards(n_samples: int, world: int, halo: int = HALO):
    """Split the frames of an n_samples clip over `world` ranks.  Returns a list of dicts with the
    global frame ranges (own [f0, f1), segment [e0, e1)), the segment's sample range [s0, s0 + n_seg)
    and the local own-frame indices the C ABI takes."""
    T = 1 + n_samples // HOP
    if T < 2 * halo * world + 2:
        raise ValueError(f"clip too short to shard over {world} ranks (needs >= {2 * halo} frames per rank)")
    bounds = [2 * int(round(T * r / world / 2.0)) for r in range(world)] + [T]
    plan = []
    for r in range(world):
        f0, f1 = bounds[r], bounds[r + 1]
        e0 = f0 - halo if r > 0 else 0
        e1 = f1 + halo if r < world - 1 else T
        s0 = HOP * e0
        n_seg = n_samples - s0 if r == world - 1 else HOP * (e1 - e0 - 1) + 1
        out_lo, out_hi = HOP * f0, min(HOP * f1, HOP * (T - 1))
        plan.append(dict(rank=r, f0=f0, f1=f1, e0=e0, e1=e1, s0=s0, n_seg=n_seg, own_lo=f0 - e0, own_hi=f1 - e0,
                         out_lo=out_lo, out_hi=out_hi, T=T))
    return plan


class Comm:
    """The two collectives the C library asks for, on a caller-owned arena tensor (device memory on
    the box; a CPU tensor in the host-only tests)."""

    def __init__(self, device, group=None):
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if on else 0
        self.world = dist.get_world_size(group) if on else 1
        self.backend = dist.get_backend(group) if on else "none"
        self.arena = torch.zeros((64 << 10) + self.world * (32 << 10), dtype=torch.uint8, device=device)
        self.error = None
        self.n_allreduce = self.n_allgather = 0
        self._ar = _lib.ALLREDUCE_FN(self._allreduce)       # keep the thunks alive
        self._ag = _lib.ALLGATHER_FN(self._allgather)
        self.struct = _lib.AwComm(None, self._ar, self._ag, C.c_void_p(self.arena.data_ptr()),
                                  self.arena.numel(), self.rank, self.world)

    # ---- callbacks (called from inside aw_*_sharded, on the calling thread) ----------------------
    def _allreduce(self, user, offset, count, dtype, op, stream):
        try:
            t = self.arena[offset:offset + 8 * count].view(torch.float64 if dtype == _lib.COMM_F64 else torch.int64)
            if self.world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.SUM if op == _lib.COMM_SUM else dist.ReduceOp.MAX,
                                group=self.group)
            self.n_allreduce += 1
            return 0
        except Exception as e:  # noqa: BLE001  (must not propagate through the C frame)
            self.error = e
            return 1

    def _allgather(self, user, send_offset, recv_offset, nbytes, stream):
        try:
            send = self.arena[send_offset:send_offset + nbytes]
            recv = self.arena[recv_offset:recv_offset + self.world * nbytes]
            if self.world == 1:
                recv.copy_(send)
            elif self.backend == "nccl" or not send.is_cuda:
                dist.all_gather_into_tensor(recv, send, group=self.group)
            else:       # gloo has no device all-gather: stage through the host (test path)
                parts = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(self.world)]
                dist.all_gather(parts, send.cpu(), group=self.group)
                recv.copy_(torch.cat(parts).to(recv.device))
            self.n_allgather += 1
            return 0
        except Exception as e:  # noqa: BLE001
            self.error = e
            return 1

    def check(self, status):
        if status != 0 and self.error is not None:
            e, self.error = self.error, None
            raise RuntimeError("collective failed inside the sharded long-form call") from e
        _lib.check(status)


class Longform:
    """Frame-sharded embed / detect of one long clip.  Every rank calls the same method with the SAME
    full clip (numpy, host memory); each uploads only its own segment."""

    def __init__(self, embedder, group=None):
        self.embedder = embedder
        self.engine = embedder.engine
        self.comm = Comm(self.engine.device, group)
        self.last_stats = {}

    def _segment(self, audio, plan):
        p = plan[self.comm.rank]
        seg = np.ascontiguousarray(np.asarray(audio, dtype=np.float32)[p["s0"]:p["s0"] + p["n_seg"]])
        return p, torch.from_numpy(seg).to(self.engine.device)

    def detect(self, audio: np.ndarray, sample_rate: int, precision: str | None = None) -> np.ndarray:
        """AWAREDetector.detect for a clip sharded over the ranks: float32 [20], identical everywhere.
        Clips whose margin is below the engine's exact margin are evaluated again with fp32 GEMMs (the
        decision is taken on identical values on every rank, so all ranks repeat together)."""
        eng = self.engine
        plan = plan_shards(len(audio), self.comm.world)
        p, seg = self._segment(audio, plan)
        out = torch.empty(N_BITS, dtype=torch.float32, device=eng.device)

        def run():
            self.comm.check(eng.lib.aw_detect_sharded(
                eng._ctx, _ptr(seg), p["n_seg"], p["e0"], p["own_lo"], p["own_hi"], p["T"], int(sample_rate),
                C.byref(self.comm.struct), _ptr(out), _stream()))
            return out.cpu().numpy()
        with eng._with_precision(precision):
            v = run()
        if eng.precision != "fp32" and precision in (None, "tf32", "fp16", "bf16") and eng.exact_margin > 0 \
                and np.abs(v - eng.threshold).min() < eng.exact_margin:
            with eng._with_precision("fp32"):
                v = run()
        return v

    def embed(self, audio: np.ndarray, sample_rate: int, watermark, iters: int | None = None,
              precision: str | None = None, gather: bool = True, return_losses: bool = False):
        """AWAREEmbedder.embed for a clip sharded over the ranks.  gather=True: the whole watermarked,
        peak-normalised clip (float32 [256*(T-1)]) on every rank; gather=False: (own samples, offset)."""
        eng = self.engine
        iters = self.embedder.num_iterations if iters is None else iters
        plan = plan_shards(len(audio), self.comm.world)
        p, seg = self._segment(audio, plan)
        pat = torch.as_tensor(np.asarray(watermark)).to(device=eng.device, dtype=torch.int32).contiguous()
        if pat.shape != (N_BITS,):
            raise ValueError("Invalid watermark length.")
        n_own = p["out_hi"] - p["out_lo"]
        own = torch.empty(n_own, dtype=torch.float32, device=eng.device)
        best = torch.empty(1, dtype=torch.float32, device=eng.device)
        losses = torch.zeros(max(iters, 1), dtype=torch.float32, device=eng.device) if return_losses else None
        counts = (C.c_int64 * 2)()
        with eng._with_precision(precision or self.embedder.embed_precision):
            self.comm.check(eng.lib.aw_embed_sharded(
                eng._ctx, _ptr(seg), p["n_seg"], p["e0"], p["own_lo"], p["own_hi"], p["T"], int(sample_rate),
                _ptr(pat), int(iters), C.byref(self.comm.struct), _ptr(own), n_own, _ptr(best), _ptr(losses),
                counts, _stream()))
        self.last_stats = {"allreduces": int(counts[0]), "allgathers": int(counts[1]), "iterations": iters,
                           "allreduces_per_iteration": (int(counts[0]) - 3) / max(iters, 1),
                           "own_frames": p["f1"] - p["f0"], "segment_frames": p["e1"] - p["e0"]}
        if not gather:
            res = (own, p["out_lo"])
        else:
            res = self.gather(own, plan)
        return (res, losses.cpu().numpy()) if return_losses else res

    def gather(self, own: torch.Tensor, plan) -> np.ndarray:
        """Own slices -> the whole clip on every rank (padded all-gather; host staging under gloo)."""
        if self.comm.world == 1:
            return own.cpu().numpy()
        width = max(q["out_hi"] - q["out_lo"] for q in plan)
        pad = torch.zeros(width, dtype=torch.float32, device=own.device)
        pad[:own.numel()] = own
        if self.comm.backend == "nccl":
            full = torch.empty(self.comm.world * width, dtype=torch.float32, device=own.device)
            dist.all_gather_into_tensor(full, pad, group=self.comm.group)
            parts = full.view(self.comm.world, width).cpu()
        else:
            lst = [torch.empty(width, dtype=torch.float32) for _ in range(self.comm.world)]
            dist.all_gather(lst, pad.cpu(), group=self.comm.group)
            parts = torch.stack(lst)
        return np.concatenate([parts[q["rank"], :q["out_hi"] - q["out_lo"]].numpy() for q in plan])


# ---- service-level wrappers (service/embed.py:7-80, service/detect.py:7-55 semantics, mono) ---------------
def embed_watermark_longform(audio: np.ndarray, sample_rate: int, watermark_bits, model, group=None,
                             longform: Longform | None = None) -> np.ndarray:
    from .service.embed import _check_rate, _encode
    _check_rate(sample_rate, model)
    wm = _encode(watermark_bits, model)
    audio = np.asarray(audio)
    if audio.ndim != 1:
        raise ValueError("Invalid audio shape. Expected 1D or 2D numpy array.")
    lf = longform or Longform(model, group)
    return np.max(audio) * lf.embed(audio, sample_rate, wm)         # signed max (service/embed.py:69,73)


def detect_watermark_longform(audio: np.ndarray, sample_rate: int, detector, group=None,
                              longform: Longform | None = None) -> np.ndarray:
    from .service.detect import _check_rate
    from .utils.watermark import PatternDecoder
    _check_rate(sample_rate, detector)
    lf = longform or Longform(detector._engine_owner or detector, group)
    lf.engine.set_threshold(detector.threshold)
    v = lf.detect(np.asarray(audio), sample_rate)
    return PatternDecoder(encoder_mode=detector.pattern_mode, threshold=detector.threshold)(v)
