"""Tensor-core spectral path of the fp16 embed loop (csrc/spectc.cuh) against the fp32 FFT kernels it
replaces (csrc/spec.cuh, themselves gated against the oracle in test_gpu_parity.py): same inputs, same
state, one switch (Engine.set_tc_spectral).  The Toeplitz GEMMs carry fp16 operands, so the comparison
is at fp16 resolution for the interior frames and at fp32 resolution for the six edge frames per clip,
which the path evaluates with the FFT kernel in edge mode."""
import numpy as np
import pytest
import torch

import aware_oracle as O

pytestmark = pytest.mark.gpu
SR = 44100


@pytest.fixture(scope="module")
def eng():
    from aware_b200.utils.models import load
    emb, _ = load()
    emb.verbose = False
    return emb.engine


def _run(eng, x, pat, iters, tc):
    eng.set_tc_spectral(tc, min_frames=2)          # the default only switches the path on for large batches
    try:
        out, best, losses = eng.embed(x, SR, pat, iters=iters, precision="fp16", return_losses=True)
    finally:
        eng.set_tc_spectral(True, min_frames=24 * 1024)
    n, T = x.shape[0], 1 + x.shape[1] // 256
    _, nb = eng.band_bins(SR)
    st = {k: eng.embed_state(k, n, T, SR).cpu().numpy() for k in ("c", "m", "c0")}
    st["mag"] = eng.debug_buffer(14, n * T * nb).view(n, T, nb).cpu().numpy()
    st["q"] = eng.debug_buffer(16, n * T * nb * 2).view(n, T, nb, 2).cpu().numpy()
    st["dA"] = eng.debug_buffer(13, n * T * nb).view(n, T, nb).cpu().numpy()
    st["peak"] = eng.debug_buffer(12, 2 * n).cpu().numpy()[1::2].copy()        # high word of the packed peak = |y| max
    return out.cpu().numpy(), losses.cpu().numpy(), st


def _rms(a):
    return float(np.sqrt(np.mean(np.asarray(a, dtype=np.float64) ** 2)))


@pytest.mark.parametrize("secs,idx", [(3.0, [0, 1, 2]), (1.6, [4, 7])])
def test_one_iteration_matches_the_fft_path(eng, secs, idx):
    x = torch.from_numpy(np.stack([O.synth_clip(i, secs, SR) for i in idx])).cuda()
    pat = torch.from_numpy(np.stack([O.encode_bits(O.synth_bits(8)[i]) for i in idx]))
    y0, l0, a = _run(eng, x, pat, 1, False)
    y1, l1, b = _run(eng, x, pat, 1, True)
    T = a["mag"].shape[1]
    np.testing.assert_array_equal(a["c0"], b["c0"])
    # forward: peak of y, |S| and the phasor of every frame
    assert np.abs(b["peak"] / a["peak"] - 1).max() <= 2e-3, (a["peak"], b["peak"])
    scale = a["mag"].max(axis=(1, 2), keepdims=True)
    d = np.abs(b["mag"] - a["mag"]) / scale
    print("|S|: max diff %.2e of the clip maximum, rel. RMS %.2e; peak rel. diff %.2e"
          % (d.max(), _rms(b["mag"] - a["mag"]) / _rms(a["mag"]), np.abs(b["peak"] / a["peak"] - 1).max()))
    assert d.max() <= 4e-3 and _rms(b["mag"] - a["mag"]) <= 2e-3 * _rms(a["mag"])
    edge = np.r_[0:3, T - 3:T]
    assert (np.abs(b["mag"][:, edge] - a["mag"][:, edge]) / scale).max() <= 1e-5      # edge frames: the fp32 kernel
    strong = a["mag"] > 1e-2 * scale
    assert np.abs(b["q"] - a["q"])[strong].max() <= 2e-2
    assert np.abs(b["q"][:, edge] - a["q"][:, edge])[strong[:, edge]].max() <= 1e-4
    assert np.abs(l1[0] - l0[0]).max() <= 3e-3, (l0[0], l1[0])
    # backward: the spectral gradient entering (dA) and the coefficient gradient leaving (m = 0.1 g)
    g0, g1 = a["m"] / 0.1, b["m"] / 0.1
    rel = _rms(g1 - g0) / _rms(g0)
    agree = float(np.mean(np.sign(g1) == np.sign(g0)))
    rel_e = _rms(g1[:, edge] - g0[:, edge]) / _rms(g0[:, edge])
    print("gradient: rel. RMS diff %.3f (edge frames %.3f), sign agreement %.4f; dA rel. RMS diff %.3f"
          % (rel, rel_e, agree, _rms(b["dA"] - a["dA"]) / _rms(a["dA"])))
    assert rel <= 0.12 and agree >= 0.96 and rel_e <= 0.15
    assert np.isfinite(y1).all() and y1.shape == y0.shape


def test_loss_trajectory_and_functional_parity(eng):
    """12 iterations: the loss follows the FFT path's; 400 iterations: BER 0 by the CUDA detector and by the
    CPU oracle, SNR against the host within 1 dB (mean) of the FFT path's result."""
    from aware_b200.metrics.audio import SNR
    idx = [0, 1, 2, 3]
    xn = np.stack([O.synth_clip(i, 3.0, SR) for i in idx])
    x = torch.from_numpy(xn).cuda()
    bits = O.synth_bits(8)[idx]
    pat = torch.from_numpy(np.stack([O.encode_bits(b) for b in bits]))
    _, l0, _ = _run(eng, x, pat, 12, False)
    _, l1, _ = _run(eng, x, pat, 12, True)
    print("loss after 12 iterations: fft %s tc %s" % (l0[11], l1[11]))
    assert np.abs(l1[:12] - l0[:12]).max() <= 3e-2
    assert (l1[11] < l1[0] - 0.05).all()
    y0, _, _ = _run(eng, x, pat, 400, False)
    y1, _, _ = _run(eng, x, pat, 400, True)
    L = y1.shape[1]
    v = eng.detect(torch.from_numpy(y1).cuda(), SR).cpu().numpy()
    np.testing.assert_array_equal((v > 0).astype(np.int32), bits)
    assert np.abs(v).min() > 0.05
    for i in range(len(idx)):
        np.testing.assert_array_equal(O.detect_watermark(y1[i], SR), bits[i])
    s0 = np.array([SNR()(y0[i], xn[i][:L]) for i in range(len(idx))])
    s1 = np.array([SNR()(y1[i], xn[i][:L]) for i in range(len(idx))])
    print("SNR fft path %s, tc path %s" % (s0, s1))
    assert abs(s1.mean() - s0.mean()) <= 1.0
    assert int(eng.embed_status().sum()) == 0
