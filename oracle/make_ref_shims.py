"""Materialise the import shims that let the UNMODIFIED reference
(/root/reference, read-only, Python-only) run in this container.

TEST INFRASTRUCTURE ONLY.  Output goes to oracle/_ref/ (git-ignored).  Nothing
under aware_b200/ may import from here.  The reference cannot travel to the GPU
box, so this script is only ever useful where /root/reference exists; it is used
to (a) validate oracle/aware_oracle.py and (b) generate tests/golden/*.npz.

Shims (none touches hot-path arithmetic):
  aware            -> symlink to /root/reference/src/AWARE (package dir is
                      upper-case, imports are lower-case)
  librosa          -> fft_frequencies only (== np.fft.rfftfreq)
  webrtcvad        -> Vad.is_speech -> True (C extension not installed)
  resampy, pesq, pystoi, soundfile, pyrubberband, matplotlib -> empty stubs
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("AWARE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref", "shims")

FILES = {
    "librosa/__init__.py": (
        "import numpy as np\n"
        "def fft_frequencies(sr=22050, n_fft=2048):\n"
        "    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)\n"
        "def resample(*a, **k):\n    raise NotImplementedError('librosa stub')\n"
        "def load(*a, **k):\n    raise NotImplementedError('librosa stub')\n"
    ),
    "librosa/display.py": "",
    "matplotlib/__init__.py": "",
    "matplotlib/pyplot.py": "",
    "soundfile.py": "",
    "pyrubberband.py": "",
    "resampy.py": "",
    "webrtcvad.py": (
        "class Vad:\n"
        "    def __init__(self, mode=0):\n        self.mode = mode\n"
        "    def is_speech(self, frame, sample_rate):\n        return True\n"
    ),
    "pesq.py": "def pesq(*a, **k):\n    raise NotImplementedError('pesq stub')\n",
    "pystoi.py": "def stoi(*a, **k):\n    raise NotImplementedError('pystoi stub')\n",
}


def build():
    if not os.path.isdir(os.path.join(REF, "src", "AWARE")):
        return None
    os.makedirs(OUT, exist_ok=True)
    link = os.path.join(OUT, "aware")
    if not os.path.islink(link):
        os.symlink(os.path.join(REF, "src", "AWARE"), link)
    for rel, body in FILES.items():
        p = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write(body)
    return OUT


def activate():
    """Put the shims (and the reference's scripts/) first on sys.path."""
    out = build()
    if out is None:
        raise RuntimeError("reference tree not present at %s" % REF)
    for p in (os.path.join(REF, "scripts"), out):
        if p not in sys.path:
            sys.path.insert(0, p)
    return out


if __name__ == "__main__":
    print(build())
