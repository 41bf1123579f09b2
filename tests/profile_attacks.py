"""Development aid: the streaming attack passes on 256 x 10 s clips, for an ncu capture
    python tests/profile_attacks.py                       # must exit 0 first
    ncu --set full --clock-control none --import-source on -k regex:'k_fir_tiled|k_attack_affine|k_attack_pcm|k_attack_decim' \
        -c 5 -o gpurun_out/r2_attacks python tests/profile_attacks.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from aware_b200 import attacks as A                       # noqa: E402
from aware_b200.utils.models import load                  # noqa: E402

emb, _ = load()
A.set_engine(emb.engine)
sr, n = 44100, 256
L = 256 * (10 * sr // 256)
g = torch.Generator(device="cuda").manual_seed(1)
y = torch.rand((n, L), generator=g, device="cuda") - 0.5
buf = torch.randn((n, L), generator=g, device="cuda")
for att in (A.FIRFilter("lowpass", 4000.0), A.AdditiveNoise(0.01, buffer=buf), A.Gain(0.5),
            A.PCMBitDepthConversion(16), A.Resample()):
    z = att.apply_batch(y, sr)
torch.cuda.synchronize()
print("ok", tuple(z.shape))
