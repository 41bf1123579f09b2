"""Round-2 GPU diagnostics (development aid, not a test): python tests/gpu_diag2.py <what>"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch


def big_detect():
    from aware_b200.synth import synth_batch, synth_bits
    from aware_b200.utils.models import load
    from aware_b200 import attacks as A
    emb, det = load(); emb.verbose = False
    eng = emb.engine; A.set_engine(eng)
    sr = 44100
    x = torch.from_numpy(synth_batch(256, 10.0, sr)).cuda()
    pat = torch.from_numpy(2 * synth_bits(256) - 1)
    for tc in (False, True):
        eng.set_tc_spectral(tc)
        y = eng.embed(x, sr, pat, iters=6, scale="signed_max", precision="fp16")
        torch.cuda.synchronize()
        print("embed tc=%s ok" % tc, flush=True)
        for n in (256, 512, 1024):
            yy = y.repeat(n // 256, 1)
            try:
                v = eng.detect(yy, sr)
                torch.cuda.synchronize()
                print("  detect n=%d ok" % n, float(v.abs().min()), flush=True)
            except Exception as e:  # noqa: BLE001
                print("  detect n=%d FAILED: %s" % (n, e), flush=True)


if __name__ == "__main__":
    {"big_detect": big_detect}[sys.argv[1]]()
