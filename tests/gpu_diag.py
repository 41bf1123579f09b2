"""Stage-by-stage GPU diagnostics against the CPU oracle (development tool).

    python tests/gpu_diag.py [stage ...]      # default: all stages, each in its own process

Each stage runs in a subprocess under a timeout so that a faulting or hanging
kernel cannot take the other stages (or the GPU box) with it.  Results are
appended to gpurun_out/diag.jsonl and printed.
"""
import json
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
OUT = os.path.join(ROOT, "gpurun_out")

STAGES = ["stft", "istft", "attacks", "detect_fp32", "gemm_tc", "detect_tf32", "embed1_fp32",
          "embed1_tf32", "embed3", "embed_full", "timing", "dual", "timeline", "gradprec", "longform"]


def _engine(precision="fp32"):
    import aware_oracle as O
    from aware_b200.engine import Engine
    return Engine([w.numpy() for w in O.make_weights()], O.mel_basis(), O.hann().numpy(), precision=precision)


def _clips(idx, secs, sr):
    import numpy as np
    import aware_oracle as O
    return np.stack([O.synth_clip(i, secs, sr) for i in idx])


def stage_stft(res):
    import numpy as np
    import torch
    import aware_oracle as O
    eng = _engine()
    for sr, secs in ((16000, 1.3), (44100, 1.0)):
        x = _clips([0, 1, 2], secs, sr)
        mag, ph = eng.stft_band(torch.from_numpy(x).cuda(), sr, phasor=True)
        mag, ph = mag.cpu().numpy(), ph.cpu().numpy()
        fi, _ = O.band_indices(sr)
        for i in range(len(x)):
            s = O.stft(O.normalize_waveform(torch.from_numpy(x[i]))).numpy()[fi].T    # [T][B]
            em = np.abs(mag[i] - np.abs(s)).max() / np.abs(s).max()
            ephase = np.abs((ph[i][..., 0] + 1j * ph[i][..., 1]) * np.abs(s) - s).max() / np.abs(s).max()
            res["sr%d_clip%d" % (sr, i)] = dict(mag_rel=float(em), spec_rel=float(ephase))
        mag2 = eng.stft_band(torch.from_numpy(x).cuda(), sr).cpu().numpy()
        res["sr%d_magonly_equal" % sr] = bool(np.array_equal(mag2, mag))


def stage_istft(res):
    import numpy as np
    import torch
    import aware_oracle as O
    eng = _engine()
    for sr, secs in ((16000, 1.3), (44100, 1.0)):
        x = _clips([3, 4], secs, sr)
        mag, ph = eng.stft_band(torch.from_numpy(x).cuda(), sr, phasor=True)
        y = eng.istft_band(mag, ph, sr).cpu().numpy()
        fi, nfi = O.band_indices(sr)
        for i in range(len(x)):
            s = O.stft(O.normalize_waveform(torch.from_numpy(x[i])))
            s[torch.from_numpy(nfi)] = 0
            yr = O.istft(s).numpy()
            res["sr%d_clip%d" % (sr, i)] = dict(shape_ok=bool(y[i].shape == yr.shape),
                                                max_abs=float(np.abs(y[i] - yr).max()),
                                                ref_peak=float(np.abs(yr).max()))


def stage_attacks(res):
    import random
    import numpy as np
    import torch
    import aware_oracle as O
    from aware_b200 import attacks as A
    eng = _engine()
    A.set_engine(eng)
    for sr in (16000, 44100):
        x = _clips([3, 5], 0.8, sr)
        xd = torch.from_numpy(x).cuda()
        n = x.shape[1]

        def cmp(name, got, want_fn):
            got = got.cpu().numpy()
            errs = []
            for i in range(len(x)):
                want = np.asarray(want_fn(x[i]), dtype=np.float32)
                if got[i].shape != want.shape:
                    errs.append(float("nan"))
                else:
                    errs.append(float(np.abs(got[i] - want).max()))
            res["%s_sr%d" % (name, sr)] = errs

        for pcm in (8, 12, 16, 24):
            cmp("pcm%d" % pcm, A.PCMBitDepthConversion(pcm).apply_batch(xd, sr), lambda a: O.attack_pcm(a, pcm))
        st = np.array([100, n // 2])
        got = A.DeleteSamples(0.15, start=st).apply_batch(xd, sr).cpu().numpy()
        res["delete_sr%d" % sr] = [float(np.abs(got[i] - O.attack_delete(x[i], 0.15, int(st[i]))).max()) for i in range(2)]
        got = A.SampleSupression(0.25, start=st).apply_batch(xd, sr).cpu().numpy()
        res["suppress_sr%d" % sr] = [float(np.abs(got[i] - O.attack_suppress(x[i], 0.25, sr, int(st[i]))).max()) for i in range(2)]
        cmp("cropout", A.Cropout(0.1).apply_batch(xd, sr), lambda a: O.attack_cropout(a, 0.1, sr))
        cmp("resample", A.Resample().apply_batch(xd, sr), lambda a: O.attack_resample(a, sr))
        cmp("lowpass", A.LowPassFilter().apply_batch(xd, sr), lambda a: O.attack_lowpass(a, sr))
        cmp("highpass", A.HighPassFilter().apply_batch(xd, sr), lambda a: O.attack_highpass(a, sr))
        random.seed(5)
        f_low = random.uniform(300.0, 3800.0)
        cmp("bandstop", A.RandomBandstop(f_low=f_low).apply_batch(xd, sr), lambda a: O.attack_bandstop(a, sr, f_low))
    # long clip: chunked IIR scan across many chunks
    sr = 44100
    x = _clips([6], 5.0, sr)
    xd = torch.from_numpy(x).cuda()
    res["lowpass_long"] = float(np.abs(A.LowPassFilter().apply_batch(xd, sr).cpu().numpy()[0] - O.attack_lowpass(x[0], sr).astype(np.float32)).max())
    res["highpass_long"] = float(np.abs(A.HighPassFilter().apply_batch(xd, sr).cpu().numpy()[0] - O.attack_highpass(x[0], sr).astype(np.float32)).max())
    res["bandstop_long"] = float(np.abs(A.RandomBandstop(f_low=3700.0).apply_batch(xd, sr).cpu().numpy()[0] - O.attack_bandstop(x[0], sr, 3700.0)).max())


def _detect_stage(res, precision):
    import numpy as np
    import torch
    import aware_oracle as O
    eng = _engine(precision)
    for sr, secs in ((16000, 2.0), (44100, 1.5), (44100, 3.1)):
        x = _clips([0, 1, 2, 3, 4], secs, sr)
        v = eng.detect(torch.from_numpy(x).cuda(), sr).cpu().numpy()
        ref = np.stack([O.detect(x[i], sr) for i in range(len(x))])
        res["sr%d_s%g" % (sr, secs)] = dict(max_abs=float(np.abs(v - ref).max()),
                                            bits_equal=bool(np.array_equal(v > 0, ref > 0)),
                                            min_margin=float(np.abs(ref).min()))
    # batch of one and a second call with another shape (workspace regrow)
    x = _clips([7], 1.2, 16000)
    v = eng.detect(torch.from_numpy(x).cuda(), 16000).cpu().numpy()
    res["single"] = float(np.abs(v[0] - O.detect(x[0], 16000)).max())


def stage_detect_fp32(res):
    _detect_stage(res, "fp32")


def stage_detect_tf32(res):
    _detect_stage(res, "tf32")


def stage_gemm_tc(res):
    import torch
    eng = _engine("tf32")
    torch.manual_seed(0)
    for rows, n, k in ((128, 64, 32), (128, 64, 64), (256, 128, 128), (256, 256, 512), (384, 512, 128),
                       (256, 1024, 1024), (1280, 1024, 512), (256, 64, 1024), (256, 128, 512), (256, 1024, 64),
                       (128 * 300, 1024, 1024), (128 * 149, 512, 128)):
        a = torch.randn(rows, k, device="cuda")
        b = torch.randn(n, k, device="cuda") / k ** 0.5
        want = (a.double() @ b.double().T)
        got_tc = eng.gemm(a, b, "tf32").double()
        got_ex = eng.gemm(a, b, "fp32").double()
        r = dict(tf32_rel=float((got_tc - want).abs().max().item() / (want.abs().max().item())),
                 fp32_rel=float((got_ex - want).abs().max().item() / (want.abs().max().item())))
        if k % 64 == 0:
            got_bf = eng.gemm(a, b, "bf16").double()
            want_bf = a.bfloat16().double() @ b.bfloat16().double().T
            r["bf16_rel_vs_bf16_inputs"] = float((got_bf - want_bf).abs().max().item() / want_bf.abs().max().item())
        torch.cuda.synchronize()
        res["%dx%dx%d" % (rows, n, k)] = r


def _embed_iters(res, precision, iters, sr=16000, secs=1.0, idx=(0, 1)):
    import numpy as np
    import torch
    import aware_oracle as O
    eng = _engine(precision)
    x = _clips(list(idx), secs, sr)
    bits = O.synth_bits(8)[:len(idx)]
    pat = np.stack([O.encode_bits(b) for b in bits])
    out, best, losses = eng.embed(torch.from_numpy(x).cuda(), sr, torch.from_numpy(pat), iters=iters,
                                  return_losses=True)
    T = 1 + x.shape[1] // 256
    c = eng.embed_state("c", len(idx), T, sr).cpu().numpy()
    c0 = eng.embed_state("c0", len(idx), T, sr).cpu().numpy()
    out, losses = out.cpu().numpy(), losses.cpu().numpy()
    for i in range(len(idx)):
        keep = {}
        y = O.embed(x[i], sr, pat[i], num_iters=iters, keep=keep)
        B = c.shape[2]
        c_ref = keep["coeffs_after"][iters].numpy().reshape(B, T).T
        c0_ref = keep["c0"].numpy().reshape(B, T).T
        d = np.abs(c[i] - c_ref)
        res["clip%d" % i] = dict(
            c0_rel=float(np.abs(c0[i] - c0_ref).max() / np.abs(c0_ref).max()),
            loss_gpu=[float(v) for v in losses[:iters, i]], loss_ref=[float(v) for v in keep["losses"]],
            coeff_max_abs=float(d.max()), coeff_frac_within_1e4=float((d <= 1e-4 * np.maximum(1, np.abs(c_ref))).mean()),
            coeff_frac_within_1e2=float((d <= 1e-2).mean()),
            wave_max_abs=float(np.abs(out[i] - y).max()),
            wave_snr_db=float(O.snr_db(out[i], y)))


def stage_embed1_fp32(res):
    _embed_iters(res, "fp32", 1)
    _embed_iters(res.setdefault("sr44100", {}), "fp32", 1, sr=44100, secs=0.8, idx=(2,))


def stage_embed1_tf32(res):
    _embed_iters(res, "tf32", 1)


def stage_embed3(res):
    _embed_iters(res.setdefault("fp32", {}), "fp32", 3)
    _embed_iters(res.setdefault("tf32", {}), "tf32", 3)


def stage_embed_full(res):
    import numpy as np
    import torch
    import aware_oracle as O
    for precision in ("tf32", "bf16", "fp32"):
        eng = _engine("tf32" if precision == "bf16" else precision)
        eng.embed_precision = precision
        sr = 16000
        x = _clips([1, 2, 3], 2.0, sr)
        bits = O.synth_bits(8)[1:4]
        pat = np.stack([O.encode_bits(b) for b in bits])
        t0 = time.time()
        out, best, losses = eng.embed(torch.from_numpy(x).cuda(), sr, torch.from_numpy(pat), iters=400,
                                      return_losses=True)
        torch.cuda.synchronize()
        dt = time.time() - t0
        v = eng.detect(out, sr).cpu().numpy()
        out = out.cpu().numpy()
        r = dict(seconds=dt, best=[float(b) for b in best.cpu().numpy()],
                 loss_first=[float(v_) for v_ in losses[0].cpu().numpy()],
                 loss_last=[float(v_) for v_ in losses[-1].cpu().numpy()])
        r["ber_gpu"] = [float(np.mean((v[i] > 0).astype(np.int32) != bits[i]) * 100) for i in range(3)]
        r["ber_oracle_cross"] = [O.ber_percent(bits[i], O.detect_watermark(out[i], sr)) for i in range(3)]
        r["snr_db"] = [O.snr_db(out[i] * np.max(x[i]), x[i]) for i in range(3)]
        r["min_margin"] = float(np.abs(v).min())
        res[precision] = r


def stage_timing(res):
    import numpy as np
    import torch
    from aware_b200.synth import synth_batch, synth_bits
    from aware_b200.utils.watermark import PatternEncoder
    sr = 44100
    for n, iters, prec in ((32, 20, "tf32"), (128, 10, "tf32"), (128, 10, "bf16")):
        eng = _engine("tf32")
        eng.embed_precision = prec
        x = torch.from_numpy(synth_batch(n, 10.0, sr)).cuda()
        pat = torch.from_numpy(np.stack([PatternEncoder()(b) for b in synth_bits(n)]))
        eng.embed(x, sr, pat, iters=2)
        torch.cuda.synchronize()
        t0 = time.time()
        out = eng.embed(x, sr, pat, iters=iters)
        torch.cuda.synchronize()
        dt = time.time() - t0
        t1 = time.time()
        v = eng.detect(out, sr)
        torch.cuda.synchronize()
        dd = time.time() - t1
        res["n%d_%s" % (n, prec)] = dict(ms_per_iter=1e3 * dt / iters, est_400it_audio_s_per_s=n * 10.0 / (dt / iters * 400),
                              detect_ms=1e3 * dd, detect_audio_s_per_s=n * 10.0 / dd)


def stage_timeline(res):
    """Device time per kernel class inside the embed loop (events around every launch)."""
    import numpy as np
    import torch
    from aware_b200.synth import synth_batch, synth_bits
    from aware_b200.utils.watermark import PatternEncoder
    for n, iters, prec, sr in ((128, 10, "tf32", 44100), (128, 10, "fp16", 44100), (256, 10, "fp16", 16000)):
        eng = _engine("tf32")
        eng.embed_precision = prec
        x = torch.from_numpy(synth_batch(n, 10.0, sr)).cuda()
        pat = torch.from_numpy(np.stack([PatternEncoder()(b) for b in synth_bits(n)]))
        eng.embed(x, sr, pat, iters=2)
        torch.cuda.synchronize()
        eng.profile(True)
        eng.embed(x, sr, pat, iters=iters)
        torch.cuda.synchronize()
        tl = eng.profile_read_named()
        eng.profile(False)
        tot = sum(v[1] for v in tl.values())
        res["n%d_%s_%d" % (n, prec, sr)] = dict(total_ms_per_iter=tot / iters,
                                         classes={k: [v[0], round(v[1] / iters, 4)] for k, v in
                                                  sorted(tl.items(), key=lambda kv: -kv[1][1])})


def stage_gradprec(res):
    """First-iteration gradient (m_1 / 0.1) of each GEMM precision against the fp32 path."""
    import numpy as np
    import torch
    import aware_oracle as O
    sr = 44100
    x = torch.from_numpy(_clips([0, 1, 2, 3], 2.0, sr)).cuda()
    pat = torch.from_numpy(np.stack([O.encode_bits(b) for b in O.synth_bits(4)]))
    T = 1 + x.shape[1] // 256
    g = {}
    for prec in ("fp32", "tf32", "fp16", "bf16"):
        eng = _engine(prec)
        eng.embed(x, sr, pat, iters=1)
        g[prec] = eng.embed_state("m", 4, T, sr).cpu().numpy() / 0.1
    ref = g["fp32"]
    for prec in ("tf32", "fp16", "bf16"):
        d = g[prec] - ref
        res[prec] = dict(rel_rms=float(np.sqrt((d ** 2).mean() / (ref ** 2).mean())),
                         finite=bool(np.isfinite(g[prec]).all()),
                         sign_agree=float((np.sign(g[prec]) == np.sign(ref)).mean()))


def stage_longform(res):
    """BASELINE configs[0] (one 10 s clip) and configs[4] (one 1 h clip) on a single GPU."""
    import numpy as np
    import torch
    from aware_b200.synth import synth_batch, synth_bits
    from aware_b200.utils.watermark import PatternEncoder
    sr = 44100
    pat = torch.from_numpy(np.stack([PatternEncoder()(b) for b in synth_bits(1)]))
    for name, secs, iters in (("1x10s", 10.0, 100), ("1x3600s", 3600.0, 5)):
        eng = _engine("tf32")
        eng.embed_precision = "fp16"
        base = synth_batch(1, min(secs, 60.0), sr)
        x = torch.from_numpy(np.tile(base, (1, int(round(secs / min(secs, 60.0)))))).cuda()
        eng.embed(x, sr, pat, iters=2)
        torch.cuda.synchronize()
        eng.profile(True)
        t0 = time.time()
        out = eng.embed(x, sr, pat, iters=iters)
        torch.cuda.synchronize()
        dt = time.time() - t0
        tl = eng.profile_read_named()
        eng.profile(False)
        t1 = time.time()
        v = eng.detect(out, sr)
        torch.cuda.synchronize()
        dd = time.time() - t1
        res[name] = dict(ms_per_iter=1e3 * dt / iters, est_400it_audio_s_per_s=secs / (dt / iters * 400),
                         detect_ms=1e3 * dd, finite=bool(torch.isfinite(out).all()),
                         classes={k: round(v_[1] / iters, 4) for k, v_ in
                                  sorted(tl.items(), key=lambda kv: -kv[1][1])[:10]})
        del eng, x, out
        torch.cuda.empty_cache()


def stage_dual(res):
    """Two half-batches on two streams / two contexts from two host threads: do the HBM-bound
    elementwise kernels of one half overlap the tensor-bound GEMMs of the other?"""
    import threading
    import numpy as np
    import torch
    from aware_b200.synth import synth_batch, synth_bits
    from aware_b200.utils.watermark import PatternEncoder
    sr, n, iters = 44100, 128, 10
    x = torch.from_numpy(synth_batch(n, 10.0, sr)).cuda()
    pat = torch.from_numpy(np.stack([PatternEncoder()(b) for b in synth_bits(n)]))
    for prec in ("tf32", "bf16"):
        for parts in (1, 2, 4):
            engs = [_engine("tf32") for _ in range(parts)]
            streams = [torch.cuda.Stream() for _ in range(parts)]
            h = n // parts

            def work(i, it):
                with torch.cuda.stream(streams[i]):
                    engs[i].embed_precision = prec
                    engs[i].embed(x[i * h:(i + 1) * h], sr, pat[i * h:(i + 1) * h], iters=it)

            def run(it):
                th = [threading.Thread(target=work, args=(i, it)) for i in range(parts)]
                [t.start() for t in th]
                [t.join() for t in th]
                torch.cuda.synchronize()

            run(2)
            t0 = time.time()
            run(iters)
            dt = time.time() - t0
            res["%s_parts%d" % (prec, parts)] = dict(ms_per_iter=1e3 * dt / iters)
            del engs


def run_stage(name):
    res = {}
    t0 = time.time()
    try:
        globals()["stage_" + name](res)
        status = "ok"
    except Exception as e:  # noqa: BLE001
        status = "error: %s" % (str(e).splitlines()[0] if str(e) else type(e).__name__)
        res["traceback"] = traceback.format_exc()[-1500:]
    rec = dict(stage=name, status=status, seconds=round(time.time() - t0, 2), result=res)
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "diag.jsonl"), "a") as f:
        f.write(json.dumps(rec) + "\n")
    print(json.dumps(rec, indent=1)[:6000], flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_stage(sys.argv[2])
        sys.exit(0)
    stages = sys.argv[1:] or STAGES
    for s in stages:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", s], timeout=600)
            if p.returncode != 0:
                print("stage %s exited %d" % (s, p.returncode), flush=True)
        except subprocess.TimeoutExpired:
            print("stage %s TIMED OUT" % s, flush=True)
            os.makedirs(OUT, exist_ok=True)
            with open(os.path.join(OUT, "diag.jsonl"), "a") as f:
                f.write(json.dumps(dict(stage=s, status="timeout")) + "\n")
