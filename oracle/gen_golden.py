"""Generate tests/golden/*.npz by running the UNMODIFIED reference (through
oracle/make_ref_shims.py).  Only runs where /root/reference exists (the build
container); the fixtures it writes are committed and travel to the GPU box.

TEST INFRASTRUCTURE ONLY.

Inputs are never stored: every fixture records the generator call
(`aware_oracle.synth_clip(i, seconds, sr)`, `synth_bits`) that reproduces them
and the reference outputs for it.

    python oracle/gen_golden.py            # writes tests/golden/
"""
import logging
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ref_shims  # noqa: E402

make_ref_shims.activate()
import aware_oracle as O  # noqa: E402
import torch  # noqa: E402,F401
from aware.utils.logger import logger  # noqa: E402
from aware.utils.models import load  # noqa: E402
from aware.utils.watermark import PatternDecoder, PatternEncoder  # noqa: E402
from aware.service import embed_watermark, detect_watermark  # noqa: E402
from aware.metrics.audio import BER, SNR  # noqa: E402
import attacks as A  # noqa: E402  (/root/reference/scripts/attacks.py)

logger.setLevel(logging.ERROR)
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def main():
    emb, det = load()
    enc = PatternEncoder("bits2bipolar")
    bits_all = O.synth_bits(8)

    # --- detect: raw 20-vectors on un-watermarked clips, both rates --------
    rec = {}
    for sr in (16000, 44100):
        for i in range(4):
            secs = [1.0, 2.0, 3.0, 2.5][i]
            x = O.synth_clip(i, secs, sr)
            rec["values_sr%d_clip%d_s%g" % (sr, i, secs)] = det.detect(x, sr)
    np.savez_compressed(os.path.join(OUT, "detect.npz"), **rec)

    # --- embed: 1 and 3 iterations (waveforms), both rates ------------------
    rec = {}
    for sr in (16000, 44100):
        x = O.synth_clip(0, 1.0, sr)
        wm = enc(bits_all[0])
        for iters in (1, 3):
            emb.num_iterations = iters
            y = emb.embed(x, sr, wm)
            rec["wave_sr%d_it%d" % (sr, iters)] = y.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "embed_short.npz"), **rec)

    # --- full 400-iteration embed through the service API (16 kHz) ----------
    emb.num_iterations = 400
    x = O.synth_clip(1, 2.0, 16000)
    t0 = time.time()
    y = embed_watermark(x, 16000, bits_all[1], emb)
    t_embed = time.time() - t0
    dec = detect_watermark(y, 16000, det)
    vals = det.detect(y, 16000)
    rec = dict(wave=y.astype(np.float32), bits=bits_all[1], decoded=dec, values=vals,
               ber=np.float64(BER()(bits_all[1], dec)),
               snr=np.float64(SNR()(y, x)), seconds=np.float64(t_embed))
    # and at 44.1 kHz through the model interface (service rejects != 16 kHz)
    x44 = O.synth_clip(2, 2.0, 44100)
    y44 = np.max(x44) * emb.embed(x44, 44100, enc(bits_all[2]))
    v44 = det.detect(y44, 44100)
    rec.update(wave44=y44.astype(np.float32), bits44=bits_all[2], values44=v44,
               decoded44=PatternDecoder(0.0, "bits2bipolar")(v44),
               snr44=np.float64(SNR()(y44, x44)))
    np.savez_compressed(os.path.join(OUT, "embed_full.npz"), **rec)
    print("full embed: %.1fs  ber=%.1f snr=%.2f snr44=%.2f" % (t_embed, rec["ber"], rec["snr"], rec["snr44"]))

    # --- attacks -------------------------------------------------------------
    rec = {}
    for sr in (16000, 44100):
        x = O.synth_clip(3, 0.4, sr)
        for pcm in (8, 12, 16, 24):
            rec["pcm%d_sr%d" % (pcm, sr)] = A.PCMBitDepthConversion(pcm).apply(x, sr)
        for p in (0.1, 0.15, 0.2):
            np.random.seed(11)
            rec["delete%g_sr%d" % (p, sr)] = A.DeleteSamples(p).apply(x, sr)
        for p in (0.1, 0.25):
            np.random.seed(12)
            rec["suppress%g_sr%d" % (p, sr)] = A.SampleSupression(p).apply(x, sr)
        rec["cropout0.1_sr%d" % sr] = A.Cropout(0.1).apply(x, sr)
        rec["resample_sr%d" % sr] = A.Resample().apply(x, sr)
        random.seed(13)
        rec["bandstop_sr%d" % sr] = A.RandomBandstop().apply(x, sr)
        rec["lowpass_sr%d" % sr] = A.LowPassFilter().apply(x, sr)
        rec["highpass_sr%d" % sr] = A.HighPassFilter().apply(x, sr)
    np.savez_compressed(os.path.join(OUT, "attacks.npz"), **rec)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
