"""Host-side audio helpers kept for interface compatibility.  The arithmetic of
WaveformNormalizer / STFT / ISTFT lives in the CUDA kernels; only the VAD gate
(reference utils/audio/waveform.py:22-46) is a host object (webrtcvad is a C extension)."""
import numpy as np


class SilenceChecker:
    """True when the clip holds less than `min_speech_seconds` of voiced frames
    (utils/audio/waveform.py:22-46 upstream: webrtcvad, 8/16/32/48 kHz only).

    Upstream imports webrtcvad unconditionally, so a host without it cannot embed at all.  Here
    the gate is switched by the model card (`vad_gate`, see cards/config.yaml): when it is ON and
    webrtcvad is missing this raises ImportError -- it never silently reports "not silent"."""

    def __init__(self, sample_rate=16000, aggr=3, frame_ms=30.0, min_speech_seconds=0.01):
        self.sample_rate, self.aggr = sample_rate, aggr
        self.frame_ms, self.min_speech_seconds = frame_ms, min_speech_seconds

    def __call__(self, data: np.ndarray) -> bool:
        try:
            import webrtcvad
        except ImportError as e:
            raise ImportError("the speech / silence gate needs the `webrtcvad` package; install it or "
                              "switch the gate off explicitly (vad_gate: false in cards/config.yaml, "
                              "or model.vad_gate = False)") from e
        pcm = (np.asarray(data) * 32767).astype(np.int16).tobytes()
        vad = webrtcvad.Vad(self.aggr)
        step = int(self.sample_rate * self.frame_ms / 1000) * 2
        voiced = sum(vad.is_speech(pcm[i:i + step], self.sample_rate)
                     for i in range(0, len(pcm) - step + 1, step))
        return voiced * (self.frame_ms / 1000.0) < self.min_speech_seconds


def silent_mask(clips, sample_rate: int, model) -> np.ndarray:
    """bool [n]: True where the VAD gate rejects the clip (all False when `model.vad_gate` is off)."""
    n = len(clips)
    if not getattr(model, "vad_gate", False):
        return np.zeros(n, dtype=bool)
    chk = SilenceChecker(sample_rate=sample_rate)
    return np.array([bool(chk(np.asarray(c))) for c in clips], dtype=bool)


__all__ = ["SilenceChecker", "silent_mask"]
