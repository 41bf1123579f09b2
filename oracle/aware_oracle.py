"""CPU oracle for the AWARE hot path (embed -> attack -> detect -> BER).

TEST INFRASTRUCTURE ONLY.  This file is a CPU restatement (torch-CPU / numpy /
scipy) of the reference algorithm.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import it.  Nothing under
aware_b200/ imports it; the product path has no CPU fallback.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so this oracle is pinned against outputs of the
*unmodified reference itself*, run in the build container through
oracle/make_ref_shims.py: tests/golden/*.npz (written by
oracle/gen_golden.py) hold reference outputs for detect, 1..3 embed iterations,
a full embed, and every in-scope attack; tests/test_oracle.py checks this file
against those vectors on every CPU run, and -- when /root/reference is present
-- against the live reference as well.

All file:line citations are relative to /root/reference/.  Third-party
arithmetic the reference relies on (torch.stft/istft, Conv1d, InstanceNorm1d,
NAdam, scipy.signal) is called from the same libraries here; the restatement
covers the reference's own code.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# constants: src/AWARE/cards/config.yaml:3-46
# ----------------------------------------------------------------------------
N_FFT = 1024
HOP = 256
BANDS = (500.0, 4000.0)
TOLERANCE_DB = 6.0
NUM_ITERS = 400
LR = 0.1
N_MELS = 128
MEL_SR = 16000           # config.yaml:34 -- mel basis is built for 16 kHz always
CHANNELS = (128, 512, 1024, 1024, 40)
N_BITS = 20
THRESHOLD = 0.0
SEED = 328656719         # detection/multibit_detector_net.py:78


# ----------------------------------------------------------------------------
# weights: detection/multibit_detector_net.py:58-80,98-107
# ----------------------------------------------------------------------------
def make_weights():
    """Seeded xavier_uniform_ conv weights, in module order (block 0..3).

    `self.apply(_init_weights)` visits the four Conv1d modules in order and
    draws each weight with nn.init.xavier_uniform_ from the global CPU
    generator seeded with SEED; biases are zero (and cancelled by the
    InstanceNorm that follows every conv).  Returns [W0..W3], W_l of shape
    (C_out, C_in) float32.  Restores the caller's RNG state (the reference
    does not: load_model.py side effect Q21).
    """
    state = torch.random.get_rng_state()
    torch.manual_seed(SEED)
    ws = []
    for cin, cout in zip(CHANNELS[:-1], CHANNELS[1:]):
        w = torch.empty(cout, cin, 1)
        torch.nn.init.xavier_uniform_(w)
        ws.append(w[:, :, 0].contiguous())
    torch.random.set_rng_state(state)
    return ws


# ----------------------------------------------------------------------------
# mel basis: detection/modules/mel.py:6-149 (Slaney, librosa-compatible)
# ----------------------------------------------------------------------------
def _hz_to_mel(f):
    f = np.atleast_1d(np.asarray(f, dtype=np.float64))
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    log_t = f >= min_log_hz
    mels[log_t] = min_log_mel + np.log(f[log_t] / min_log_hz) / logstep
    return mels


def _mel_to_hz(m):
    m = np.atleast_1d(np.asarray(m, dtype=np.float64))
    f_sp = 200.0 / 3
    hz = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    log_t = m >= min_log_mel
    hz[log_t] = min_log_hz * np.exp(logstep * (m[log_t] - min_log_mel))
    return hz


def mel_basis(sr=MEL_SR, n_fft=N_FFT, n_mels=N_MELS):
    """mel.py:105-149 get_mel_filter_bank(norm='slaney', dtype=float32)."""
    fmax = float(sr) / 2
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.linspace(0, sr / 2, 1 + n_fft // 2, endpoint=True)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0)[0], _hz_to_mel(fmax)[0], n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


# ----------------------------------------------------------------------------
# band bins: embedding/multibit_embedder.py:43-47, detection/multibit_detector.py:34-37
# ----------------------------------------------------------------------------
def band_indices(sample_rate, n_fft=N_FFT, bands=BANDS):
    freqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sample_rate)   # == librosa.fft_frequencies
    mask = (freqs >= bands[0]) & (freqs <= bands[1])
    return np.where(mask)[0], np.where(~mask)[0]


# ----------------------------------------------------------------------------
# audio ops: utils/audio/waveform.py:19, utils/audio/stft.py:28,48,55,62
# ----------------------------------------------------------------------------
def normalize_waveform(x: torch.Tensor) -> torch.Tensor:
    return x / torch.max(torch.abs(x) + 1e-8)


_WINDOW = None


def hann():
    global _WINDOW
    if _WINDOW is None:
        _WINDOW = torch.hann_window(N_FFT)
    return _WINDOW


def stft(x: torch.Tensor) -> torch.Tensor:
    return torch.stft(x, n_fft=N_FFT, hop_length=HOP, center=True, window=hann(),
                      return_complex=True)


def istft(spec: torch.Tensor) -> torch.Tensor:
    return torch.istft(spec, n_fft=N_FFT, hop_length=HOP, center=True, window=hann())


def stft_manual(x: np.ndarray) -> np.ndarray:
    """reflect-pad 512, frame 1024/256, periodic Hann, rFFT (float64 math).

    Independent restatement used by tests to cross-check torch.stft and the
    CUDA kernels' framing conventions."""
    w = hann().numpy().astype(np.float64)
    xp = np.pad(np.asarray(x, dtype=np.float64), (N_FFT // 2, N_FFT // 2), mode="reflect")
    T = 1 + len(x) // HOP
    frames = np.stack([xp[t * HOP:t * HOP + N_FFT] * w for t in range(T)], axis=1)
    return np.fft.rfft(frames, axis=0)


def istft_manual(spec: np.ndarray) -> np.ndarray:
    """irfft * w, overlap-add, trim 512 each side, THEN / sum w^2."""
    w = hann().numpy().astype(np.float64)
    T = spec.shape[1]
    L = HOP * (T - 1)
    out = np.zeros(L + N_FFT)
    env = np.zeros(L + N_FFT)
    fr = np.fft.irfft(spec, n=N_FFT, axis=0)
    for t in range(T):
        out[t * HOP:t * HOP + N_FFT] += fr[:, t] * w
        env[t * HOP:t * HOP + N_FFT] += w * w
    h = N_FFT // 2
    return out[h:h + L] / env[h:h + L]


# ----------------------------------------------------------------------------
# detector net: detection/multibit_detector_net.py:109-140 and modules/*
# ----------------------------------------------------------------------------
class Net:
    def __init__(self):
        self.W = make_weights()
        self.mel = torch.from_numpy(mel_basis())

    def forward(self, mag: torch.Tensor, keep=None) -> torch.Tensor:
        """mag: (513, T) magnitude with out-of-band rows already zeroed -> (20,).

        `keep`, if a dict, receives intermediates (for per-stage kernel tests).
        The first GlobalStandardize of the reference is computed and discarded
        (multibit_detector_net.py:121 vs :124) so it is not applied here.
        """
        x = mag.unsqueeze(0)                                   # (1, 513, T)
        m = torch.matmul(x.transpose(1, 2), self.mel.T).transpose(1, 2)   # mel.py:195
        mh = F.instance_norm(m, eps=1e-5)                       # nn.InstanceNorm1d(128)
        g = (mh - mh.mean()) / (mh.std() + 1e-8)               # globalStandardize.py:17-19
        p = F.avg_pool1d(g, kernel_size=2, stride=2)           # :131
        if keep is not None:
            keep.update(mel=m, mel_in=mh, gs=g, p0=p)
        for l, w in enumerate(self.W):                         # conv1d.py:38-42
            h = F.conv1d(p, w.unsqueeze(-1))
            p = F.leaky_relu(F.instance_norm(h, eps=1e-5), 0.2)
            if keep is not None:
                keep["h%d" % (l + 1)] = h
                keep["p%d" % (l + 1)] = p
        z = p.mean(dim=2)                                      # BRH.py:18
        v = torch.tanh(z[:, 0::2] - z[:, 1::2])                # BRH.py:21-25
        if keep is not None:
            keep.update(z=z, v=v)
        return v.reshape(-1)


_NET = None


def net() -> Net:
    global _NET
    if _NET is None:
        _NET = Net()
    return _NET


# ----------------------------------------------------------------------------
# detect: detection/multibit_detector.py:28-42
# ----------------------------------------------------------------------------
def detect(audio: np.ndarray, sample_rate: int, keep=None) -> np.ndarray:
    x = torch.from_numpy(np.asarray(audio)).float()            # utils/utils.py:21
    spec = stft(normalize_waveform(x))
    mag = spec.abs()
    _, oob = band_indices(sample_rate)
    mag[oob] = 0.0
    if keep is not None:
        keep["mag"] = mag
    with torch.no_grad():
        return net().forward(mag, keep).numpy()


# ----------------------------------------------------------------------------
# pattern codec: utils/watermark/encoder.py:35-45, decoder.py:40-51,59-63
# ----------------------------------------------------------------------------
def encode_bits(bits) -> np.ndarray:
    return np.array([2 * int(b) - 1 for b in bits], dtype=np.int32)


def decode_values(values: np.ndarray, threshold: float = THRESHOLD) -> np.ndarray:
    bipolar = 2 * (np.asarray(values) > threshold).astype(np.int32) - 1
    return (bipolar > 0).astype(np.int32)


# ----------------------------------------------------------------------------
# metrics: metrics/audio.py:8-17, 68-89
# ----------------------------------------------------------------------------
def ber_percent(output, target) -> float:
    return float(np.mean(np.asarray(output) != np.asarray(target)) * 100)


def snr_db(output, target) -> float:
    o = np.asarray(output)
    t = np.asarray(target)
    n = min(len(o), len(t))
    o, t = o[:n], t[:n]
    if np.all(o == t):
        return float("inf")
    return float(10 * np.log10(np.mean(o ** 2) / np.mean((o - t) ** 2)))


# ----------------------------------------------------------------------------
# loss: embedding/losses.py:38-42
# ----------------------------------------------------------------------------
def push_extremes_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return F.mse_loss(pred, target) - 0.1 * torch.mean(torch.abs(pred))


# ----------------------------------------------------------------------------
# NAdam: torch/optim/nadam.py _single_tensor_nadam (non-capturable CPU path),
# selected by embedding/optimizers.py:5,20 with config.yaml:17-21
# ----------------------------------------------------------------------------
def nadam_scalars(num_iters, lr=LR, beta1=0.9, beta2=0.999, momentum_decay=4e-3):
    """Per-step scalar factors, exactly as the Python-float code computes them.

    Returns float64 array (num_iters, 3): [a_g, a_m, bias_correction2] with
      c += a_g * g / denom + a_m * m / denom ; denom = sqrt(v / bc2) + eps.
    mu_product lives in a float32 0-d tensor in torch (`mu_product *= mu`), so
    it is accumulated in float32 here too.
    """
    out = np.zeros((num_iters, 3), dtype=np.float64)
    mu_product = np.float32(1.0)
    for step in range(1, num_iters + 1):
        bc2 = 1 - beta2 ** step
        mu = beta1 * (1.0 - 0.5 * (0.96 ** (step * momentum_decay)))
        mu_next = beta1 * (1.0 - 0.5 * (0.96 ** ((step + 1) * momentum_decay)))
        mu_product = np.float32(mu_product * np.float32(mu))
        mp = float(mu_product)
        out[step - 1] = (-lr * (1.0 - mu) / (1.0 - mp),
                         (-lr * mu_next) / (1.0 - mp * mu_next),
                         bc2)
    return out


def nadam_step(c, g, m, v, scal, beta1=0.9, beta2=0.999, eps=1e-8):
    """In-place single NAdam step on float32 tensors with the scalar row `scal`."""
    a_g, a_m, bc2 = (float(s) for s in scal)
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    denom = v.div(bc2).sqrt().add_(eps)
    c.addcdiv_(g, denom, value=a_g)
    c.addcdiv_(m, denom, value=a_m)


# ----------------------------------------------------------------------------
# embed: embedding/multibit_embedder.py:141-197 with _optimize :70-138
# ----------------------------------------------------------------------------
def embed(audio: np.ndarray, sample_rate: int, pattern: np.ndarray,
          num_iters: int = NUM_ITERS, keep=None) -> np.ndarray:
    """Returns the watermarked, peak-normalised waveform, length 256*(T-1).

    The reference builds `bounds` with a Python loop over ~140k 0-d tensors
    (multibit_embedder.py:159-160, ~3.7 s/clip); this restatement computes the
    same float32 values vectorised -- (c - d) and (c + d) with
    d = c * 10**(-6/20) -- so a CPU-baseline timed on it is *faster* than the
    real reference, never slower.

    keep (dict, optional) receives: mag0, phase, c0, lo, hi, losses (list),
    grads (list of first 3 gradient tensors), coeffs_after (dict it->coeffs),
    best_loss.
    """
    x = torch.from_numpy(np.asarray(audio)).float()
    spec = stft(normalize_waveform(x))                         # :143-147
    magnitude, phase = spec.abs(), torch.angle(spec)
    target = torch.from_numpy(np.asarray(pattern)).float()
    fi, nfi = band_indices(sample_rate)                        # :152
    fi_t = torch.from_numpy(fi)
    nfi_t = torch.from_numpy(nfi)
    c0 = magnitude[fi_t].flatten()                             # :154 (bin-major)
    delta = c0 * 10 ** (-TOLERANCE_DB / 20)                    # :157
    lo = torch.clamp(c0 - delta, min=0.0)                      # :159 max(0, c - d)
    hi = c0 + delta
    scal = nadam_scalars(num_iters)
    coeffs = c0.clone().requires_grad_(True)                   # :79
    m = torch.zeros_like(c0)
    v = torch.zeros_like(c0)
    best_loss = float("inf")
    best = c0.clone()
    rot = torch.exp(1j * phase)                                # stft.py:62 (constant)
    if keep is not None:
        keep.update(mag0=magnitude, phase=phase, c0=c0.clone(), lo=lo, hi=hi,
                    losses=[], grads=[], coeffs_after={}, values=[])
    nn_ = net()
    for it in range(num_iters):                                # :95
        if coeffs.grad is not None:
            coeffs.grad = None
        wm = magnitude.clone()
        wm[fi_t] = coeffs.reshape(len(fi), -1)                 # :99-101
        y = istft(wm * rot)                                    # :103 -> :49-67
        y = normalize_waveform(normalize_waveform(y))          # post + pre pipelines
        mag2 = stft(y).abs()
        mag2[nfi_t] = 0.0                                      # :104
        pred = nn_.forward(mag2)                               # :107
        loss = push_extremes_loss(pred, target)                # :109
        loss.backward()                                        # :111
        g = coeffs.grad
        with torch.no_grad():
            nadam_step(coeffs, g, m, v, scal[it])              # :112
            coeffs.copy_(torch.clamp(coeffs, lo, hi))          # :116-117
        lv = loss.item()
        if lv < best_loss:                                     # :120-122
            best_loss = lv
            best = coeffs.detach().clone()
        if keep is not None:
            keep["losses"].append(lv)
            keep["values"].append(pred.detach().numpy().copy())
            if it < 3:
                keep["grads"].append(g.detach().clone())
            if it < 3 or it == num_iters - 1:
                keep["coeffs_after"][it + 1] = coeffs.detach().clone()
    wm = magnitude.clone()
    wm[fi_t] = best.reshape(len(fi), -1)                       # :173-174
    with torch.no_grad():
        y = normalize_waveform(istft(wm * rot))                # :185-192
    if keep is not None:
        keep["best_loss"] = best_loss
        keep["best"] = best
    return y.numpy()


# ----------------------------------------------------------------------------
# service layer: service/embed.py:7-80, service/detect.py:7-55 (mono path)
# ----------------------------------------------------------------------------
def embed_watermark(audio: np.ndarray, sample_rate: int, bits, num_iters=NUM_ITERS,
                    enforce_16k=False) -> np.ndarray:
    if enforce_16k and sample_rate != 16000:                   # embed.py:24-26
        raise ValueError("Invalid sample rate. Expected 16000Hz.")
    wm = encode_bits(bits)
    if len(wm) != N_BITS:                                      # embed.py:32-34
        raise ValueError("Invalid watermark length.")
    mx = np.max(audio)                                         # embed.py:69 (signed max)
    return mx * embed(audio, sample_rate, wm, num_iters)       # embed.py:73


def detect_watermark(audio: np.ndarray, sample_rate: int) -> np.ndarray:
    return decode_values(detect(audio, sample_rate))           # detect.py:44-51


# ----------------------------------------------------------------------------
# attacks: scripts/attacks.py (A1..A8 of SURVEY section 8a).  Randomness is an
# explicit argument (the reference draws it unseeded: attacks.py:170,340,378).
# ----------------------------------------------------------------------------
def attack_pcm(audio: np.ndarray, bits: int) -> np.ndarray:
    """attacks.py:44-70 PCMBitDepthConversion (peak-normalise, scale, clip, trunc)."""
    a = audio / np.max(np.abs(audio) + 1e-8)
    s, lo, hi, dt = {8: (127.0, -128, 127, np.int8), 12: (4095.0, -4096, 4095, np.int16),
                     16: (32767.0, -32768, 32767, np.int16),
                     24: (8388607.0, -8388608, 8388607, np.int32)}[bits]
    ai = np.clip(a * s, lo, hi).astype(dt)
    return ai.astype(np.float32) / s


def attack_delete(audio: np.ndarray, percentage: float, start: int) -> np.ndarray:
    """attacks.py:162-178 DeleteSamples; start = np.random.randint(0, N - n_del)."""
    n_del = int(percentage * len(audio))
    return np.concatenate([audio[:start], audio[start + n_del:]])


def attack_cropout(audio: np.ndarray, percentage: float, sr: int) -> np.ndarray:
    """attacks.py:192-205 Cropout: drop the first int(p*sr) samples."""
    return audio[int(percentage * sr):]


def attack_suppress(audio: np.ndarray, percentage: float, sr: int, start: int) -> np.ndarray:
    """attacks.py:370-385 SampleSupression: zero int(p*sr) samples from start."""
    n = int(percentage * sr)
    out = audio.copy()
    out[start:start + n] = 0
    return out


def attack_resample(audio: np.ndarray, sr: int, target_sr: int = 16000) -> np.ndarray:
    """attacks.py:267-294 Resample (decimate+interp if sr//target>1 else polyphase)."""
    f = sr // target_sr
    if f > 1:
        down = audio[::f]
        return np.interp(np.arange(len(audio)), np.arange(0, len(audio), f), down)
    from scipy.signal import resample_poly
    return resample_poly(resample_poly(audio, 441, 160), 160, 441)


def attack_compression_approx(x, sr, step_db=1.5, floor_db=-60.0, bands=(500.0, 4000.0)):
    """numpy / torch-CPU restatement of aware_b200.attacks.CompressionApprox for ONE clip.  PARITY UNPINNED
    to the reference: upstream's MP3Compression shells out to ffmpeg (scripts/attacks.py:73-148) and has
    no arithmetic to follow; this pins the CUDA kernels to the attack's own definition."""
    x = np.asarray(x, dtype=np.float32)
    win = torch.hann_window(1024)
    S = torch.stft(torch.from_numpy(x), n_fft=1024, hop_length=256, win_length=1024, window=win, center=True,
                   pad_mode="reflect", return_complex=True)
    f = np.fft.rfftfreq(1024, 1.0 / sr)
    band = torch.from_numpy((f >= bands[0]) & (f <= bands[1]))
    mag = S.abs()[band].numpy()                                    # [nb, T]
    k_log = np.float32(20.0 * np.log10(2.0) / step_db)
    k_exp = np.float32(step_db / (20.0 * np.log10(2.0)))
    floor = mag.max(axis=0, keepdims=True) * np.float32(10.0 ** (floor_db / 20.0))
    with np.errstate(divide="ignore"):
        q = np.exp2(np.rint(np.log2(mag) * k_log) * k_exp).astype(np.float32)
    q = np.where((mag >= floor) & (mag > 0), q, np.float32(0.0))
    D = torch.zeros_like(S)
    ph = S[band] / torch.clamp(S[band].abs(), min=1e-30)
    D[band] = torch.from_numpy(q - mag) * ph
    delta = torch.istft(D, n_fft=1024, hop_length=256, win_length=1024, window=win, center=True).numpy()
    return (x[:len(delta)] + delta).astype(np.float32), mag.T, q.T


def butter_coeffs(kind: str, sr: int, f_low: float | None = None):
    """Filter designs used by attacks.py:342-349, 413-416, 451-453 (host-side, scipy)."""
    from scipy.signal import butter
    nyq = 0.5 * sr
    if kind == "lowpass":
        return butter(6, 4000.0 / nyq, btype="low", analog=False)
    if kind == "highpass":
        return butter(4, 500.0 / nyq, btype="highpass", analog=False)
    if kind == "bandstop":
        return butter(4, [f_low / nyq, (f_low + 200.0) / nyq], btype="bandstop")
    raise ValueError(kind)


def attack_lowpass(audio: np.ndarray, sr: int) -> np.ndarray:
    """attacks.py:400-423 LowPassFilter: butter(6, 4 kHz) + causal lfilter (float64 out)."""
    from scipy.signal import lfilter
    b, a = butter_coeffs("lowpass", sr)
    return lfilter(b, a, audio)


def attack_highpass(audio: np.ndarray, sr: int) -> np.ndarray:
    """attacks.py:438-455 HighPassFilter: butter(4, 500 Hz) + lfilter."""
    from scipy.signal import lfilter
    b, a = butter_coeffs("highpass", sr)
    return lfilter(b, a, audio)


def attack_bandstop(audio: np.ndarray, sr: int, f_low: float) -> np.ndarray:
    """attacks.py:324-356 RandomBandstop with f_low = random.uniform(300, 3800) given."""
    from scipy.signal import filtfilt
    b, a = butter_coeffs("bandstop", sr, f_low)
    return filtfilt(b, a, np.asarray(audio).astype(np.float64)).astype(audio.dtype)


def filtfilt_manual(b, a, x):
    """filtfilt(method='pad', padtype='odd', padlen=3*max(len(a),len(b)))
    spelled out as two lfilter passes -- the form the CUDA kernel follows."""
    from scipy.signal import lfilter, lfilter_zi
    x = np.asarray(x, dtype=np.float64)
    n = 3 * max(len(a), len(b))
    ext = np.concatenate([2 * x[0] - x[n:0:-1], x, 2 * x[-1] - x[-2:-n - 2:-1]])
    zi = lfilter_zi(b, a)
    y, _ = lfilter(b, a, ext, zi=zi * ext[0])
    y, _ = lfilter(b, a, y[::-1], zi=zi * y[-1])
    return y[::-1][n:-n]


# ----------------------------------------------------------------------------
# synthetic workload: SURVEY section 8(d)
# ----------------------------------------------------------------------------
def synth_clip(i: int, seconds: float, sr: int) -> np.ndarray:
    """Deterministic synthetic clip i (8 log-uniform tones + noise, peak <= 0.9)."""
    rng = np.random.default_rng(1000 + i)
    n = int(round(seconds * sr))
    t = np.arange(n, dtype=np.float64) / sr
    f = np.exp(rng.uniform(np.log(100.0), np.log(8000.0), 8))
    f = np.minimum(f, 0.45 * sr)
    a = rng.uniform(0.05, 0.3, 8)
    ph = rng.uniform(0, 2 * np.pi, 8)
    x = (a[:, None] * np.sin(2 * np.pi * f[:, None] * t[None, :] + ph[:, None])).sum(0)
    x = x + 0.02 * rng.standard_normal(n)
    peak = np.max(np.abs(x))
    if peak > 0.9:
        x = x * (0.9 / peak)
    return x.astype(np.float32)


def synth_bits(n_clips: int) -> np.ndarray:
    return np.random.default_rng(7).integers(0, 2, (n_clips, N_BITS), dtype=np.int32)
