"""Metrics with the reference's call signatures (metrics/audio.py there).

BER and SNR accept numpy arrays like the reference (torch tensors, host or device, are copied
to the host first): they are the reference's per-clip host metrics.  The batched device
versions are `Engine.decide` (BER counters, aw_decide_and_count) and `Engine.snr` (aw_snr_batch)."""
import numpy as np
import torch


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class BER:
    """mean(output != target) * 100 -- percent (metrics/audio.py:9-17)."""

    def __call__(self, output, target) -> float:
        return float(np.mean(_np(output) != _np(target)) * 100)


class SNR:
    """10 log10(mean(out^2) / mean((out - target)^2)), inf when identical, inputs
    truncated to the shorter one, stereo averaged to mono (metrics/audio.py:69-89)."""

    def __call__(self, output, target) -> float:
        o, t = _np(output), _np(target)
        if o.ndim == 2 and o.shape[1] == 2:
            o, t = o.mean(axis=1), t.mean(axis=1)
        n = min(len(o), len(t))
        o, t = o[:n], t[:n]
        if np.all(o == t):
            return float("inf")
        return float(10 * np.log10(np.mean(o ** 2) / np.mean((o - t) ** 2)))


def _perceptual(name):
    class _Metric:
        """Optional host metric: needs the third-party `pesq` / `pystoi` packages and a
        16 kHz resampler, none of which is part of the device hot path."""

        def __call__(self, output, target, sampling_rate: int) -> float:
            o, t = _np(output), _np(target)
            if o.ndim == 2 and o.shape[1] == 2:
                o, t = o.mean(axis=1), t.mean(axis=1)
            n = min(len(o), len(t))
            o, t = o[:n], t[:n]
            if sampling_rate != 16000:
                from scipy.signal import resample_poly
                o, t = resample_poly(o, 16000, sampling_rate), resample_poly(t, 16000, sampling_rate)
            if name == "PESQ":
                from pesq import pesq
                return pesq(16000, t, o, "wb")
            from pystoi import stoi
            return float(stoi(t, o, 16000))
    _Metric.__name__ = name
    return _Metric


PESQ = _perceptual("PESQ")
STOI = _perceptual("STOI")
