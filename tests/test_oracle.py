"""The oracle (oracle/aware_oracle.py) against the committed reference outputs
(tests/golden/*.npz, written by oracle/gen_golden.py from the unmodified
reference) and, where /root/reference exists, against the live reference."""
import os
import random

import numpy as np
import pytest
import torch

import aware_oracle as O

HAVE_REF = os.path.isdir("/root/reference/src/AWARE")


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _snr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return 10 * np.log10((b ** 2).sum() / max(((a - b) ** 2).sum(), 1e-300))


def test_weights_are_seed_deterministic():
    w1 = O.make_weights()
    w2 = O.make_weights()
    assert [tuple(w.shape) for w in w1] == [(512, 128), (1024, 512), (1024, 1024), (40, 1024)]
    for a, b in zip(w1, w2):
        assert torch.equal(a, b)
    assert sum(w.numel() for w in w1) + sum(w.shape[0] for w in w1) == 1681960


def test_band_bins():
    fi, nfi = O.band_indices(44100)
    assert fi[0] == 12 and fi[-1] == 92 and len(fi) == 81 and len(nfi) == 432
    fi, _ = O.band_indices(16000)
    assert fi[0] == 32 and fi[-1] == 256 and len(fi) == 225


def test_mel_basis_sparsity():
    m = O.mel_basis()
    assert m.shape == (128, 513) and m.dtype == np.float32
    assert int((m != 0).sum()) == 1009


def test_detect_matches_golden(golden_dir):
    g = _g(golden_dir, "detect.npz")
    for key in g.files:
        _, sr, clip, secs = key.split("_")
        sr, clip, secs = int(sr[2:]), int(clip[4:]), float(secs[1:])
        v = O.detect(O.synth_clip(clip, secs, sr), sr)
        np.testing.assert_allclose(v, g[key], atol=2e-6, rtol=0, err_msg=key)


def test_embed_short_matches_golden(golden_dir):
    g = _g(golden_dir, "embed_short.npz")
    bits = O.synth_bits(8)[0]
    for sr in (16000, 44100):
        x = O.synth_clip(0, 1.0, sr)
        y = O.embed(x, sr, O.encode_bits(bits), num_iters=1)
        ref = g["wave_sr%d_it1" % sr]
        assert y.shape == ref.shape == (256 * (len(x) // 256),)
        # A NAdam first step is lr * g / (|g| + 1e-8): every coefficient moves by +-0.1056 and the
        # SIGN of a gradient within fp32 reduction noise of 0 (|g| ~ 1e-6 against a median of 2e-4)
        # depends on the host's BLAS / FFT code path.  The golden vector was written on another host
        # than the one this test may run on -- the live reference on THIS host equals the oracle
        # (test_oracle_equals_live_reference) while both differ from the golden in a few dozen
        # coefficients -- so the gate is the one tests/test_gpu_parity.py::_gate_waveform_1e4 uses:
        # the stated 1e-4 / 80 dB on the samples no flipped coefficient reaches (>= 90 %; measured
        # 100 % at 16 kHz and 95.3 % at 44.1 kHz on an AVX-512 EPYC), a loose bound on the rest.
        d = np.abs(y.astype(np.float64) - ref)
        ok = d <= 1e-4
        assert ok.mean() >= 0.90, (sr, ok.mean())
        assert _snr(y[ok], ref[ok]) >= 80.0, sr
        assert d.max() <= 3e-3 and _snr(y, ref) >= 70.0, (sr, d.max())


@pytest.mark.slow
def test_embed_full_functional_parity(golden_dir):
    """400 iterations are chaotic even reference-vs-reference; the gate
    is functional: bits recovered, SNR within 1 dB of the reference's."""
    g = _g(golden_dir, "embed_full.npz")
    x = O.synth_clip(1, 2.0, 16000)
    y = O.embed_watermark(x, 16000, g["bits"])
    assert y.shape == g["wave"].shape
    dec = O.detect_watermark(y, 16000)
    assert O.ber_percent(g["bits"], dec) == 0.0 == float(g["ber"])
    assert abs(O.snr_db(y, x) - float(g["snr"])) < 1.0
    # cross-detection: the oracle detector on the reference's watermarked audio
    np.testing.assert_array_equal(O.detect_watermark(g["wave"], 16000), g["decoded"])
    np.testing.assert_allclose(O.detect(g["wave"], 16000), g["values"], atol=2e-6)
    np.testing.assert_array_equal(O.detect_watermark(g["wave44"], 44100), g["decoded44"])
    np.testing.assert_array_equal(g["decoded44"], g["bits44"])


def test_attacks_match_golden(golden_dir):
    g = _g(golden_dir, "attacks.npz")
    for sr in (16000, 44100):
        x = O.synth_clip(3, 0.4, sr)
        n = len(x)
        for pcm in (8, 12, 16, 24):
            np.testing.assert_array_equal(O.attack_pcm(x, pcm), g["pcm%d_sr%d" % (pcm, sr)])
        for p in (0.1, 0.15, 0.2):
            np.random.seed(11)
            start = np.random.randint(0, n - int(p * n))
            np.testing.assert_array_equal(O.attack_delete(x, p, start), g["delete%g_sr%d" % (p, sr)])
        for p in (0.1, 0.25):
            np.random.seed(12)
            start = np.random.randint(0, n - int(p * sr))
            np.testing.assert_array_equal(O.attack_suppress(x, p, sr, start), g["suppress%g_sr%d" % (p, sr)])
        np.testing.assert_array_equal(O.attack_cropout(x, 0.1, sr), g["cropout0.1_sr%d" % sr])
        np.testing.assert_array_equal(O.attack_resample(x, sr), g["resample_sr%d" % sr])
        random.seed(13)
        f_low = random.uniform(300.0, 4000.0 - 200.0)
        np.testing.assert_array_equal(O.attack_bandstop(x, sr, f_low), g["bandstop_sr%d" % sr])
        b, a = O.butter_coeffs("bandstop", sr, f_low)
        np.testing.assert_allclose(O.filtfilt_manual(b, a, x), g["bandstop_sr%d" % sr], atol=1e-6)
        np.testing.assert_array_equal(O.attack_lowpass(x, sr), g["lowpass_sr%d" % sr])
        np.testing.assert_array_equal(O.attack_highpass(x, sr), g["highpass_sr%d" % sr])


def test_manual_stft_istft_match_torch():
    x = O.synth_clip(5, 0.5, 44100)
    xn = O.normalize_waveform(torch.from_numpy(x))
    s = O.stft(xn).numpy()
    sm = O.stft_manual(xn.numpy())
    assert np.abs(s - sm).max() < 2e-4          # float32 FFT vs float64
    y = O.istft(torch.from_numpy(s)).numpy()
    ym = O.istft_manual(s.astype(np.complex128))
    assert y.shape == ym.shape and np.abs(y - ym).max() < 1e-5
    assert np.abs(y - xn.numpy()[:len(y)]).max() < 1e-5     # perfect reconstruction


def test_nadam_restatement_matches_torch_optim():
    torch.manual_seed(1)
    c = torch.rand(1000) + 0.5
    c_ref = c.clone().requires_grad_(True)
    opt = torch.optim.NAdam([c_ref], lr=O.LR)
    m, v = torch.zeros_like(c), torch.zeros_like(c)
    scal = O.nadam_scalars(5)
    for it in range(5):
        g = torch.randn(1000) * 1e-4
        c_ref.grad = g.clone()
        opt.step()
        O.nadam_step(c, g, m, v, scal[it])
        assert torch.equal(c, c_ref.detach()), it


def test_codec_and_metrics():
    bits = np.array([0, 1] * 10, dtype=np.int32)
    assert O.encode_bits(bits).tolist() == [-1, 1] * 10
    v = np.array([0.0, 1e-9, -0.3, 0.7])
    assert O.decode_values(v).tolist() == [0, 1, 0, 1]          # strict '>'
    assert O.ber_percent(bits, 1 - bits) == 100.0
    assert O.ber_percent(bits, bits) == 0.0
    assert O.snr_db(np.ones(4), np.ones(4)) == float("inf")
    assert abs(O.snr_db(np.ones(10), np.ones(8) * 0.9) - 20.0) < 1e-6


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present (GPU box)")
def test_oracle_equals_live_reference():
    import logging
    import make_ref_shims
    make_ref_shims.activate()
    from aware.utils.logger import logger
    from aware.utils.models import load
    logger.setLevel(logging.ERROR)
    emb, det = load()
    for sr in (16000, 44100):
        x = O.synth_clip(6, 1.0, sr)
        np.testing.assert_array_equal(det.detect(x, sr), O.detect(x, sr))
        emb.num_iterations = 2
        wm = O.encode_bits(O.synth_bits(8)[3])
        np.testing.assert_allclose(emb.embed(x, sr, wm), O.embed(x, sr, wm, num_iters=2), atol=1e-6)


def test_stoi_restatement_invariants():
    """oracle/stoi_oracle.py restates pystoi 0.4.1 (absent here: parity unpinned).  What can be pinned without the
    package: the published constants (15 one-third-octave bands from 150 Hz at 10 kHz / 512: bins 7..218), the
    Octave-style resampling window pystoi builds for 16 kHz -> 10 kHz, STOI(x, x) = 1, monotone decrease with
    additive noise, invariance to a common gain, the 1e-5 answer for clips with fewer than 30 frames, and the
    vectorised overlap-add equal to a direct loop."""
    import stoi_oracle as S
    obm, edges = S.thirdoct()
    assert obm.shape == (15, 257) and edges[0] == (7, 9) and edges[-1] == (174, 219)
    assert all(edges[i][1] == edges[i + 1][0] for i in range(14))          # contiguous bands
    h = S.resample_window_oct(10000, 16000)
    assert len(h) == 581 and abs(h.sum() - 5.0) < 1e-3 and np.allclose(h, h[::-1])
    x = O.synth_clip(0, 3.0, 16000).astype(np.float64)
    rng = np.random.default_rng(0)
    noise = rng.standard_normal(len(x))
    assert abs(S.stoi(x, x, 16000) - 1.0) < 1e-12
    d = [S.stoi(x, x + s * noise, 16000) for s in (1e-3, 1e-2, 5e-2, 2e-1)]
    assert all(a > b for a, b in zip(d, d[1:])) and d[0] > 0.99 and d[-1] < 0.4
    assert abs(S.stoi(3.0 * x, 3.0 * (x + 0.01 * noise), 16000) - d[1]) < 1e-9
    assert S.stoi(x[:4000], x[:4000], 16000) == 1e-5                         # 0.25 s: fewer than 30 frames
    with pytest.raises(Exception):
        S.stoi(x, x[:-1], 16000)
    # silence gate: frames 40 dB below the loudest frame are removed before the analysis
    y = x.copy()
    y[8000:20000] *= 1e-4
    y10 = S.resample_oct(y, 10000, 16000)
    xs, ys, mask = S.remove_silent_frames(y10, y10)
    assert 0 < mask.sum() < len(mask) and len(xs) == (mask.sum() - 1) * 128 + 256
