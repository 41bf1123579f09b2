"""Metrics with the reference's call signatures (metrics/audio.py there).

BER and SNR accept numpy arrays like the reference (torch tensors, host or device, are copied
to the host first): they are the reference's per-clip host metrics.  STOI runs on the GPU.  The batched device
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


STOI_FS = 10000          # pystoi's internal rate
_OCT_PLANS = {}


def resample_window_oct(p: int, q: int) -> np.ndarray:
    """pystoi.utils._resample_window_oct (port of Octave's `resample`): Kaiser-windowed sinc with 60 dB
    rejection, the FIR pystoi hands to scipy.signal.resample_poly when the input is not at 10 kHz."""
    import math
    g = math.gcd(p, q)
    p, q = p // g, q // g
    stop = 1.0 / (2 * max(p, q))
    roll = stop / 10
    rej = 60.0
    L = math.ceil((rej - 8) / (28.714 * roll))
    t = np.arange(-L, L + 1)
    beta = 0.1102 * (rej - 8.7)
    return np.kaiser(2 * L + 1, beta) * (2 * p * stop * np.sinc(2 * stop * t))


def resample_oct_batch(x: torch.Tensor, p: int, q: int, engine) -> torch.Tensor:
    """pystoi.utils.resample_oct for a CUDA float32 batch [n, N]: resample_poly(x, p, q, window=h / sum(h))
    on the polyphase kernel of the `Resample` attack."""
    from .. import attacks as A
    key = (x.shape[1], p, q, engine.device.index)
    if key not in _OCT_PLANS:
        h = resample_window_oct(p, q)
        h_tf, tpp, first, n_out = A.polyphase_plan(x.shape[1], p, q, taps=h / np.sum(h))
        import math
        g = math.gcd(p, q)
        _OCT_PLANS[key] = (torch.from_numpy(h_tf).to(engine.device), tpp, p // g, q // g, first, n_out)
    h, tpp, up, down, first, n_out = _OCT_PLANS[key]
    return engine.attack_upfirdn(x, h, tpp, up, down, first, n_out)


class STOI:
    """Short-time objective intelligibility with the reference's call signature (metrics/audio.py:43-64:
    stereo averaged to mono, both signals truncated to the shorter one, then pystoi.stoi(target, output, sr),
    extended=False).  Computed on the GPU (`Engine.stoi` -> aw_stoi_batch, csrc/stoi.cuh); `STOI.batch`
    scores a whole [n, N] batch in one call.  pystoi is a third-party package that is absent here: the kernels
    follow its algorithm (oracle/stoi_oracle.py restates it; parity unpinned to the package itself)."""

    def __init__(self, engine=None):
        self.engine = engine

    def _eng(self):
        if self.engine is None:
            from .. import attacks as A
            self.engine = A._eng()
        return self.engine

    def batch(self, output: torch.Tensor, target: torch.Tensor, sampling_rate: int, stoi_sum=None):
        return self._eng().stoi(target, output, sampling_rate, stoi_sum)

    def __call__(self, output, target, sampling_rate: int) -> float:
        o, t = _np(output), _np(target)
        if o.ndim == 2 and o.shape[1] == 2:
            o, t = o.mean(axis=1), t.mean(axis=1)
        n = min(len(o), len(t))
        eng = self._eng()
        od = torch.from_numpy(np.ascontiguousarray(o[:n], dtype=np.float32)).reshape(1, -1).to(eng.device)
        td = torch.from_numpy(np.ascontiguousarray(t[:n], dtype=np.float32)).reshape(1, -1).to(eng.device)
        return float(eng.stoi(td, od, sampling_rate)[0].item())


class PESQ:
    """ITU-T P.862 through the third-party `pesq` package (metrics/audio.py:19-40), a host metric: there is no
    arithmetic of it in the reference to restate, so this stays a host wrapper (ImportError without the
    package).  `PESQ.batch` scores clips on a host thread pool; its aggregate joins the counter all-reduce
    in `evaluate_clips` when the package is present."""

    def __call__(self, output, target, sampling_rate: int) -> float:
        from pesq import pesq
        o, t = _np(output), _np(target)
        if o.ndim == 2 and o.shape[1] == 2:
            o, t = o.mean(axis=1), t.mean(axis=1)
        n = min(len(o), len(t))
        o, t = o[:n], t[:n]
        if sampling_rate != 16000:
            from scipy.signal import resample_poly
            o, t = resample_poly(o, 16000, sampling_rate), resample_poly(t, 16000, sampling_rate)
        return pesq(16000, t, o, "wb")

    def batch(self, outputs, targets, sampling_rate: int, workers: int = 8):
        """Per-clip scores (NaN where pesq raises, as scripts/test.py:79-84 skips silent clips)."""
        from concurrent.futures import ThreadPoolExecutor

        def one(pair):
            try:
                return float(self(pair[0], pair[1], sampling_rate))
            except ImportError:
                raise
            except Exception:  # noqa: BLE001
                return float("nan")
        import pesq  # noqa: F401  (fail before spawning workers)
        with ThreadPoolExecutor(max_workers=workers) as ex:
            return list(ex.map(one, zip(outputs, targets)))
