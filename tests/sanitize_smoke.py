"""Reduced pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tests/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tests/sanitize_smoke.py
    compute-sanitizer --tool synccheck python tests/sanitize_smoke.py

Small shapes (2 clips x ~0.7 s, 3 optimisation iterations) so that the 10-50x slow-down of the tools
stays within a GPU call: the tcgen05 / TMA / mbarrier GEMM pipeline in all three operand types, the
fused spectral kernels with their aliasing shared-memory tiles (tile seams, both edges), the front
end / head, the exact re-evaluation path and every attack kernel.  Logs go under profiles/."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    from aware_b200 import attacks as A
    from aware_b200.synth import synth_batch, synth_bits
    from aware_b200.utils.models import load
    emb, det = load()
    emb.verbose = False
    eng = emb.engine
    A.set_engine(eng)
    sr = 44100
    x = torch.from_numpy(synth_batch(2, 0.7, sr)).cuda()           # T = 121 frames: 3 spectral tiles
    pat = torch.from_numpy(2 * synth_bits(2) - 1)
    for prec in ("fp16", "tf32", "bf16", "fp32"):
        y = eng.embed(x, sr, pat, iters=3, scale="signed_max", precision=prec)
        assert torch.isfinite(y).all(), prec
    eng.embed(x, sr, pat, iters=5, precision="fp16")               # >= 4 iterations: CUDA-graph replay path
    v = eng.detect(x, sr)                                          # un-watermarked: exact re-evaluation runs
    assert torch.isfinite(v).all()
    eng.decide(v, torch.from_numpy(synth_bits(2)), torch.zeros(3, dtype=torch.int64, device="cuda"))
    eng.snr(y, x[:, :y.shape[1]])
    x16 = torch.from_numpy(synth_batch(2, 0.5, 16000)).cuda()
    eng.embed(x16, 16000, pat, iters=2, precision="fp16")          # the <1,8> band-group instantiations
    eng.detect(x16, 16000)
    n = y.shape[1]
    st = np.array([10, n // 3])
    suite = [A.PCMBitDepthConversion(8), A.PCMBitDepthConversion(24), A.DeleteSamples(0.1, start=st), A.Cropout(0.1),
             A.SampleSupression(0.1, start=st), A.Resample(), A.RandomBandstop(f_low=1000.0),
             A.RandomBandstop(f_low=1000.0, fast=True), A.LowPassFilter(), A.LowPassFilter(fast=True),
             A.HighPassFilter(fast=True), A.AdditiveNoise(0.01), A.Gain(0.5), A.FIRFilter("bandpass", [500.0, 4000.0], 65)]
    for att in suite:
        z = att.apply_batch(y, sr)
        assert torch.isfinite(z).all(), att.name
    A.Resample().apply_batch(x16, 16000)                           # polyphase branch
    torch.cuda.synchronize()
    print("sanitize_smoke: ok, %d launches" % eng.launch_count())


if __name__ == "__main__":
    main()
