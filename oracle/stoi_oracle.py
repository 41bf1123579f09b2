"""TEST INFRASTRUCTURE -- numpy restatement of STOI as the reference computes it.

The reference calls the third-party package `pystoi` (pinned `pystoi==0.4.1` in /root/reference/setup.py:24;
call site /root/reference/src/AWARE/metrics/audio.py:43-64: `stoi(resampled_target, resampled_output, 16000)`
after truncating both signals to the shorter one; /root/reference/scripts/test.py:86-88 keeps scores > 0.1).
pystoi is NOT vendored under /root/reference and is not installable here (no network), so this file restates
its published algorithm (C. H. Taal et al., "An Algorithm for Intelligibility Prediction of Time-Frequency
Weighted Noisy Speech", IEEE TASL 2011; pystoi/stoi.py and pystoi/utils.py of the 0.4 series):

    PARITY UNPINNED: no pystoi output is available to check this restatement against.  The conventions that
    a different pystoi version could change are spelled out below (frame ranges); the CUDA path
    (aw_stoi_batch) is tested against THIS file.

Constants (pystoi/stoi.py): FS 10 kHz, frame 256, FFT 512, 15 one-third octave bands from 150 Hz, N = 30
frames per segment, BETA = -15 dB, DYN_RANGE = 40 dB, EPS = float64 machine epsilon.
"""
import math

import numpy as np
from scipy.signal import resample_poly

FS = 10000
N_FRAME = 256
NFFT = 512
NUMBAND = 15
MINFREQ = 150
N = 30
BETA = -15.0
DYN_RANGE = 40
EPS = np.finfo("float").eps


def thirdoct(fs=FS, nfft=NFFT, num_bands=NUMBAND, min_freq=MINFREQ):
    """pystoi/utils.py thirdoct: 0/1 band matrix [num_bands][nfft/2+1]; band i covers the bins from the one
    nearest to its lower edge (inclusive) to the one nearest to its upper edge (exclusive)."""
    f = np.linspace(0, fs, nfft + 1)[:nfft // 2 + 1]
    k = np.arange(num_bands, dtype=float)
    freq_low = min_freq * np.power(2.0, (2 * k - 1) / 6)
    freq_high = min_freq * np.power(2.0, (2 * k + 1) / 6)
    obm = np.zeros((num_bands, len(f)))
    edges = []
    for i in range(num_bands):
        lo = int(np.argmin(np.square(f - freq_low[i])))
        hi = int(np.argmin(np.square(f - freq_high[i])))
        obm[i, lo:hi] = 1
        edges.append((lo, hi))
    return obm, edges


def resample_window_oct(p, q):
    """pystoi/utils.py _resample_window_oct (port of Octave's resample): Kaiser-windowed sinc, 60 dB."""
    g = math.gcd(p, q)
    p, q = p // g, q // g
    log10_rejection = -3.0
    stopband_cutoff_f = 1.0 / (2 * max(p, q))
    roll_off_width = stopband_cutoff_f / 10
    rejection_db = -20 * log10_rejection
    L = math.ceil((rejection_db - 8) / (28.714 * roll_off_width))
    t = np.arange(-L, L + 1)
    ideal = 2 * p * stopband_cutoff_f * np.sinc(2 * stopband_cutoff_f * t)
    if 21 <= rejection_db <= 50:
        beta = 0.5842 * (rejection_db - 21) ** 0.4 + 0.07886 * (rejection_db - 21)
    elif rejection_db > 50:
        beta = 0.1102 * (rejection_db - 8.7)
    else:
        beta = 0.0
    return np.kaiser(2 * L + 1, beta) * ideal


def resample_oct(x, p, q):
    """pystoi/utils.py resample_oct: resample_poly with the Octave window normalised to unit sum."""
    h = resample_window_oct(p, q)
    return resample_poly(x, p, q, window=h / np.sum(h))


def hann_matlab(n):
    return np.hanning(n + 2)[1:-1]


def remove_silent_frames(x, y, dyn_range=DYN_RANGE, framelen=N_FRAME, hop=N_FRAME // 2):
    """pystoi/utils.py remove_silent_frames (0.4 series: frame starts range(0, len(x) - framelen + 1, hop)),
    overlap-add of the kept windowed frames.  Returns (x_sil, y_sil, mask)."""
    w = hann_matlab(framelen)
    starts = range(0, len(x) - framelen + 1, hop)
    xf = np.array([w * x[i:i + framelen] for i in starts])
    yf = np.array([w * y[i:i + framelen] for i in starts])
    e = 20 * np.log10(np.linalg.norm(xf, axis=1) + EPS)
    mask = (np.max(e) - dyn_range - e) < 0
    xf, yf = xf[mask], yf[mask]

    def ola(fr):
        out = np.zeros((len(fr) - 1) * hop + framelen if len(fr) else 0)
        for k, f_ in enumerate(fr):
            out[k * hop:k * hop + framelen] += f_
        return out
    return ola(xf), ola(yf), mask


def stft(x, win_size=N_FRAME, fft_size=NFFT, overlap=2):
    """pystoi/utils.py stft: frame starts range(0, len(x) - win_size, hop) (no + 1 here)."""
    hop = win_size // overlap
    w = hann_matlab(win_size)
    fr = [np.fft.rfft(w * x[i:i + win_size], n=fft_size) for i in range(0, len(x) - win_size, hop)]
    return np.array(fr).reshape(-1, fft_size // 2 + 1)


def stoi_10k(x, y):
    """STOI of two equal-length float64 signals already at 10 kHz (pystoi/stoi.py, extended=False)."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    x, y, _ = remove_silent_frames(x, y)
    xs, ys = stft(x).T, stft(y).T                      # [257][G]
    if xs.shape[-1] < N:
        return 1e-5                                    # pystoi warns and returns 1e-5
    obm, _ = thirdoct()
    xt = np.sqrt(obm @ np.square(np.abs(xs)))          # [15][G]
    yt = np.sqrt(obm @ np.square(np.abs(ys)))
    G = xt.shape[1]
    xseg = np.array([xt[:, m - N:m] for m in range(N, G + 1)])      # [J][15][30]
    yseg = np.array([yt[:, m - N:m] for m in range(N, G + 1)])
    norm = np.linalg.norm(xseg, axis=2, keepdims=True) / (np.linalg.norm(yseg, axis=2, keepdims=True) + EPS)
    yn = yseg * norm
    clip = 10 ** (-BETA / 20)
    yp = np.minimum(yn, xseg * (1 + clip))
    yp = yp - np.mean(yp, axis=2, keepdims=True)
    xseg = xseg - np.mean(xseg, axis=2, keepdims=True)
    yp = yp / (np.linalg.norm(yp, axis=2, keepdims=True) + EPS)
    xseg = xseg / (np.linalg.norm(xseg, axis=2, keepdims=True) + EPS)
    J, M = xseg.shape[0], xseg.shape[1]
    return float(np.sum(yp * xseg) / (J * M))


def stoi(clean, processed, fs_sig):
    """pystoi.stoi(clean, processed, fs_sig, extended=False)."""
    clean, processed = np.asarray(clean), np.asarray(processed)
    if clean.shape != processed.shape:
        raise Exception("x and y should have the same length, found {} and {}".format(clean.shape, processed.shape))
    if fs_sig != FS:
        clean = resample_oct(clean, FS, fs_sig)
        processed = resample_oct(processed, FS, fs_sig)
    return stoi_10k(clean, processed)
