"""Float64 numpy model of the *kernel decomposition* used by aware_b200/csrc.

This is not the oracle: it spells out, stage by stage, the algebra the CUDA
kernels implement (band-only STFT / iSTFT by linearity, analytic
GlobalStandardize statistics, hand-derived adjoints, sub-gradient of the two
peak normalisers, fused NAdam) so that tests/test_kernel_model.py can check it
against torch autograd on the oracle's forward.  Layouts follow the kernels:
spectra are [T][B] (frame-major), activations are [T'][C] (channels-last).
"""
import numpy as np

N_FFT, HOP, H = 1024, 256, 512


def hann64():
    n = np.arange(N_FFT)
    return 0.5 - 0.5 * np.cos(2 * np.pi * n / N_FFT)


def envelope(T):
    """sum_t w^2[m - 256 t] on the padded axis m in [0, 256(T-1)+1024)."""
    w2 = hann64() ** 2
    env = np.zeros(HOP * (T - 1) + N_FFT)
    for t in range(T):
        env[t * HOP:t * HOP + N_FFT] += w2
    return env


def reflect_index(i, L):
    i = np.where(i < 0, -i, i)
    return np.where(i >= L, 2 * (L - 1) - i, i)


def analysis(sig, T, bins, mode):
    """Frames of `sig` -> band spectrum [T][B] (e^{-i} DFT, Hann-windowed).

    mode 'reflect': sig has length L, padded axis value = sig[reflect(m - 512)].
    mode 'zero'   : sig already lives on the padded axis (length 256(T-1)+1024).
    """
    w = hann64()
    if mode == "reflect":
        L = len(sig)
        m = np.arange(HOP * (T - 1) + N_FFT)
        pad = sig[reflect_index(m - H, L)]
    else:
        pad = sig
    n = np.arange(N_FFT)
    out = np.zeros((T, len(bins)), dtype=np.complex128)
    tw = np.exp(-2j * np.pi * np.outer(n, bins) / N_FFT)      # [n][b]
    for t in range(T):
        out[t] = (pad[t * HOP:t * HOP + N_FFT] * w) @ tw
    return out


def synthesis(spec, bins, scale):
    """Band spectrum [T][B] -> windowed overlap-add on the padded axis.

    Per frame: x[n] = scale * 2 * Re sum_b X[b] e^{+i theta k_b n}  (the CUDA kernel
    builds the hermitian pair Z[k]=s*X, Z[N-k]=s*conj(X) and runs a complex
    inverse FFT; 2*Re(.) is the same thing).  scale = 1/N gives irfft.
    """
    w = hann64()
    T = spec.shape[0]
    n = np.arange(N_FFT)
    tw = np.exp(2j * np.pi * np.outer(bins, n) / N_FFT)       # [b][n]
    out = np.zeros(HOP * (T - 1) + N_FFT)
    for t in range(T):
        out[t * HOP:t * HOP + N_FFT] += w * (2 * scale) * np.real(spec[t] @ tw)
    return out


class Model:
    def __init__(self, W, mel, bins):
        self.W = [np.asarray(w, dtype=np.float64) for w in W]
        self.bins = np.asarray(bins)
        self.melB = np.asarray(mel, dtype=np.float64)[:, self.bins]   # [128][B]

    # ------------------------------------------------------------------ init
    def init(self, x):
        x = np.asarray(x, dtype=np.float64)
        self.T = T = 1 + len(x) // HOP
        self.L = L = HOP * (T - 1)
        self.Tp = T // 2
        self.env = envelope(T)
        xn = x / (np.max(np.abs(x)) + 1e-8)
        S = analysis(xn, T, self.bins, "reflect")
        self.A0 = np.abs(S)
        self.u = np.where(self.A0 > 0, S / np.where(self.A0 > 0, self.A0, 1), 0)
        yb = synthesis(self.A0 * self.u, self.bins, 1.0 / N_FFT)[H:H + L] / self.env[H:H + L]
        self.y_oob = xn[:L] - yb
        r = 10 ** (-6.0 / 20)
        self.lo = np.maximum(0.0, self.A0 - self.A0 * r)
        self.hi = self.A0 + self.A0 * r
        return self.A0.copy()

    # --------------------------------------------------------------- forward
    def forward(self, c, pattern):
        T, L, Tp, env = self.T, self.L, self.Tp, self.env
        s = {}
        y = synthesis(c * self.u, self.bins, 1.0 / N_FFT)[H:H + L] / env[H:H + L] + self.y_oob
        s["y"] = y
        nstar = int(np.argmax(np.abs(y)))
        p1 = abs(y[nstar])
        d1 = p1 + 1e-8
        d2 = p1 / d1 + 1e-8
        y2 = y / d1 / d2
        s.update(nstar=nstar, d1=d1, d2=d2, y2=y2)
        St = analysis(y2, T, self.bins, "reflect")
        At = np.abs(St)
        s["q"] = np.where(At > 0, St / np.where(At > 0, At, 1), 0)
        M = At @ self.melB.T                                     # [T][128]
        mu = M.mean(0)
        var = M.var(0)
        rstd = 1.0 / np.sqrt(var + 1e-5)
        Mh = (M - mu) * rstd
        n = 128 * T
        mean_g = 0.0                                             # analytic: sum_t Mh = 0
        sigma = np.sqrt(T * np.sum(var / (var + 1e-5)) / (n - 1))
        G = (Mh - mean_g) / (sigma + 1e-8)
        P = 0.5 * (G[0:2 * Tp:2] + G[1:2 * Tp:2])                # [T'][128]
        s.update(Mh=Mh, rstd0=rstd, sigma=sigma, P=[P], rstd=[], n=n)
        for Wl in self.W:
            Hl = P @ Wl.T
            mu_l = Hl.mean(0)
            r_l = 1.0 / np.sqrt(Hl.var(0) + 1e-5)
            Hh = (Hl - mu_l) * r_l
            P = np.where(Hh > 0, Hh, 0.2 * Hh)
            s["P"].append(P)
            s["rstd"].append(r_l)
        z = P.mean(0)
        v = np.tanh(z[0::2] - z[1::2])
        loss = np.mean((v - pattern) ** 2) - 0.1 * np.mean(np.abs(v))
        s.update(v=v, loss=loss)
        self.s = s
        return loss, v

    # -------------------------------------------------------------- backward
    def backward(self, pattern):
        s, T, L, Tp, env = self.s, self.T, self.L, self.Tp, self.env
        v = s["v"]
        nb = len(v)
        dv = 2 * (v - pattern) / nb - 0.1 * np.sign(v) / nb
        dd = dv * (1 - v * v)
        dz = np.zeros(2 * nb)
        dz[0::2], dz[1::2] = dd, -dd
        dP = np.tile(dz / Tp, (Tp, 1))                           # [T'][40]
        for l in range(len(self.W) - 1, -1, -1):
            P = s["P"][l + 1]
            Hh = np.where(P > 0, P, P / 0.2)                     # recover IN output
            dHh = dP * np.where(P > 0, 1.0, 0.2)
            s1 = dHh.mean(0)
            s2 = (dHh * Hh).mean(0)
            dH = s["rstd"][l] * (dHh - s1 - Hh * s2)
            dP = dH @ self.W[l]
        # pool adjoint
        dG = np.zeros((T, 128))
        dG[0:2 * Tp:2] = 0.5 * dP
        dG[1:2 * Tp:2] = 0.5 * dP
        # GlobalStandardize adjoint (unbiased std over n elements, eps outside)
        Mh, sigma, n = s["Mh"], s["sigma"], s["n"]
        alpha = 1.0 / (sigma + 1e-8)
        beta = np.sum(dG * Mh) / ((n - 1) * sigma * (sigma + 1e-8) ** 2)
        dMh = alpha * (dG - dG.mean()) - beta * Mh
        # InstanceNorm adjoint per mel channel
        dM = s["rstd0"] * (dMh - dMh.mean(0) - Mh * (dMh * Mh).mean(0))
        dA = dM @ self.melB                                      # [T][B]
        dS = dA * s["q"]
        # STFT adjoint: un-normalised one-sided synthesis (scale 1/2) + reflect fold
        dpad = synthesis(dS, self.bins, 0.5)
        dy2 = dpad[H:H + L].copy()
        dy2[1:H + 1] += dpad[H - 1::-1][:H]                      # i = 512 - m, m in [0,512)
        i = np.arange(L - 513, L - 1)                            # right pad
        dy2[i] += dpad[H + 2 * (L - 1) - i]
        # two peak normalisers (sub-gradient to the arg-max sample)
        d1, d2, nstar = s["d1"], s["d2"], s["nstar"]
        s2n = np.sum(dy2 * s["y2"])
        s1n = s2n * 1e-8 / d2
        dy = dy2 / (d1 * d2)
        dy[nstar] -= np.sign(s["y"][nstar]) * (s2n / d2 + s1n) / d1
        # iSTFT adjoint: /env, window, forward DFT, * 2/N ; dc = Re(dX conj(u))
        dola = np.zeros(HOP * (T - 1) + N_FFT)
        dola[H:H + L] = dy / env[H:H + L]
        dX = analysis(dola, T, self.bins, "zero") * (2.0 / N_FFT)
        return np.real(dX * np.conj(self.u))


# ---------------------------------------------------------------------------------------------
# k_fir_tiled (aware_b200/csrc/attacks.cuh): the index arithmetic of the register-tiled FIR --
# staged tile with one pad word per eight, a sliding window of R inputs per thread rotated in
# place, non-fused float32 product and sum in scipy's tap order.  Vectorised over the threads of
# one block; `threads` is a parameter so that the CPU test can use small tiles.
# ---------------------------------------------------------------------------------------------
def fir_tiled_model(x, h_tf, k_off, n_out, threads=256, R=8):
    x = np.asarray(x, dtype=np.float32)
    h_tf = np.asarray(h_tf, dtype=np.float32)
    hpp, n_in, tile = len(h_tf), len(x), threads * R
    addr = lambda i: i + (i >> 3)                                   # noqa: E731  (fir_addr)
    out = np.zeros(n_out, dtype=np.float32)
    base = np.arange(threads) * R
    for tile0 in range(0, n_out, tile):
        g0 = tile0 + k_off - hpp + 1
        n_stage = tile + hpp + 7
        xs = np.full(addr(n_stage) + 1, np.nan, dtype=np.float32)  # NaN = a word the kernel never staged
        g = g0 + np.arange(n_stage)
        ok = (g >= 0) & (g < n_in)
        xs[addr(np.arange(n_stage))] = np.where(ok, x[np.clip(g, 0, n_in - 1)], np.float32(0))
        acc = np.zeros((threads, R), dtype=np.float32)
        w = np.stack([xs[addr(base + r)] for r in range(R)], axis=1)
        j = 0
        while j + 8 <= hpp:
            for u in range(8):
                for r in range(R):
                    acc[:, r] = acc[:, r] + w[:, (u + r) & 7] * h_tf[j + u]     # float32 product, then float32 sum
                w[:, u] = xs[addr(base + j + u + 8)]
            j += 8
        while j < hpp:
            for r in range(R):
                acc[:, r] = acc[:, r] + xs[addr(base + j + r)] * h_tf[j]
            j += 1
        n_here = min(tile, n_out - tile0)
        out[tile0:tile0 + n_here] = acc.reshape(-1)[:n_here]
    return out
