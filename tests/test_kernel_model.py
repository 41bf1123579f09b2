"""The kernel decomposition (tests/kernel_model.py) against torch autograd on the
oracle's forward, in float64 -- proves the hand-derived adjoints the CUDA
kernels implement are the reference's gradient."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import aware_oracle as O
from kernel_model import Model


def _autograd_reference(x, sr, pattern):
    """Oracle loop body (multibit_embedder.py:95-111) in float64."""
    dt = torch.float64
    w = torch.hann_window(1024, dtype=dt)
    xn = torch.from_numpy(x).to(dt)
    xn = xn / torch.max(torch.abs(xn) + 1e-8)
    spec = torch.stft(xn, 1024, 256, window=w, center=True, return_complex=True)
    mag, phase = spec.abs(), torch.angle(spec)
    fi, nfi = O.band_indices(sr)
    c = mag[fi].clone().requires_grad_(True)
    wm = mag.clone()
    wm[fi] = c
    y = torch.istft(wm * torch.exp(1j * phase), 1024, 256, window=w, center=True)
    y = y / torch.max(torch.abs(y) + 1e-8)
    y = y / torch.max(torch.abs(y) + 1e-8)
    m2 = torch.stft(y, 1024, 256, window=w, center=True, return_complex=True).abs()
    m2[nfi] = 0.0
    mel = torch.from_numpy(O.mel_basis()).to(dt)
    m = (mel @ m2).unsqueeze(0)
    mh = F.instance_norm(m, eps=1e-5)
    g = (mh - mh.mean()) / (mh.std() + 1e-8)
    p = F.avg_pool1d(g, 2, 2)
    for wl in O.make_weights():
        p = F.leaky_relu(F.instance_norm(F.conv1d(p, wl.to(dt).unsqueeze(-1)), eps=1e-5), 0.2)
    z = p.mean(2)
    v = torch.tanh(z[:, 0::2] - z[:, 1::2]).reshape(-1)
    t = torch.from_numpy(pattern).to(dt)
    loss = F.mse_loss(v, t) - 0.1 * torch.mean(torch.abs(v))
    loss.backward()
    return loss.item(), v.detach().numpy(), c.grad.numpy(), mag[fi].numpy()


@pytest.mark.parametrize("sr,secs", [(16000, 0.6), (44100, 0.35)])
def test_decomposition_matches_autograd(sr, secs):
    x = O.synth_clip(2, secs, sr)
    pattern = O.encode_bits(O.synth_bits(4)[2]).astype(np.float64)
    loss_ref, v_ref, g_ref, c0_ref = _autograd_reference(x, sr, pattern)
    fi, _ = O.band_indices(sr)
    mdl = Model(O.make_weights(), O.mel_basis(), fi)
    c0 = mdl.init(x)
    np.testing.assert_allclose(c0, c0_ref.T, rtol=1e-9, atol=1e-11)
    loss, v = mdl.forward(c0, pattern)
    np.testing.assert_allclose(v, v_ref, rtol=0, atol=1e-9)
    assert abs(loss - loss_ref) < 1e-10
    g = mdl.backward(pattern)
    scale = np.abs(g_ref).max()
    assert np.abs(g - g_ref.T).max() <= 1e-7 * scale


@pytest.mark.parametrize("numtaps", [1, 5, 8, 21, 64, 101])
def test_fir_tile_model_equals_scipy_upfirdn(numtaps):
    """Index arithmetic of k_fir_tiled (padded staging, in-place window rotation, zero-staged samples
    outside the clip, remainder taps) on the CPU: bit-exact against scipy's float32 upfirdn for output
    windows at the start, inside and past the end of the clip.  The CUDA kernel itself is compared with
    scipy in tests/test_gpu_parity.py::test_register_tiled_fir_equals_scipy_for_every_tap_count_and_window."""
    from scipy.signal import upfirdn
    from kernel_model import fir_tiled_model
    rng = np.random.default_rng(numtaps)
    h = (rng.standard_normal(numtaps) / np.sqrt(numtaps)).astype(np.float32)
    for n in (1, 7, 255, 256, 700):
        x = rng.standard_normal(n).astype(np.float32)
        full = upfirdn(h, x, 1, 1)
        assert full.dtype == np.float32 and len(full) == n + numtaps - 1
        for first, n_out in ((0, n), (0, n + numtaps - 1), (n // 2, n - n // 2)):
            got = fir_tiled_model(x, h[::-1].copy(), first, n_out, threads=32)
            assert not np.isnan(got).any()
            np.testing.assert_array_equal(got, full[first:first + n_out], err_msg=str((n, first, n_out)))


@pytest.mark.parametrize("f", [2, 3, 4])
def test_decimate_interp_expression_equals_numpy_interp(f):
    """k_attack_decim_interp's per-sample expression (attacks.cuh): knots are returned as they are, other
    samples are fl64(fl64(slope * r) + y0) with slope = (y1 - y0) / f -- for f = 2 the kernel's 16-byte path
    drops the exact `* 1.0`, for a power-of-two f the division is a multiplication by 1/f -- and the tail
    holds the last knot.  Equal to the oracle's np.interp (scripts/attacks.py:276-287) bit for bit."""
    rng = np.random.default_rng(f)
    for n in (5, 64, 1001, 4410):
        x = rng.standard_normal(n).astype(np.float32)
        want = O.attack_resample(x, 16000 * f + 100).astype(np.float32)
        last = ((n - 1) // f) * f
        i = np.arange(n)
        k0 = np.minimum((i // f) * f, last)
        r = (i - k0).astype(np.float64)
        y0 = x[k0].astype(np.float64)
        y1 = x[np.minimum(k0 + f, last)].astype(np.float64)
        d = y1 - y0
        slope = d * (1.0 / f) if f & (f - 1) == 0 else d / f
        got = np.where(f == 2, slope + y0, slope * r + y0)
        got = np.where((i >= last) | (r == 0), x[np.minimum(k0, last)].astype(np.float64), got).astype(np.float32)
        np.testing.assert_array_equal(got, want, err_msg=str((f, n)))
