"""Host-side audio helpers kept for interface compatibility.  The arithmetic of
WaveformNormalizer / STFT / ISTFT lives in the CUDA kernels; only the VAD gate
(reference utils/audio/waveform.py:22-46) is a host object, and it is pluggable:
webrtcvad is a C extension that may be absent."""
import numpy as np

from ..logger import logger


class SilenceChecker:
    """True when the clip holds less than `min_speech_seconds` of voiced frames.

    Uses webrtcvad when importable (8/16/32/48 kHz only, as upstream); otherwise the
    gate is skipped (returns False) with a one-time warning."""
    _warned = False

    def __init__(self, sample_rate=16000, aggr=3, frame_ms=30.0, min_speech_seconds=0.01):
        self.sample_rate, self.aggr = sample_rate, aggr
        self.frame_ms, self.min_speech_seconds = frame_ms, min_speech_seconds

    def __call__(self, data: np.ndarray) -> bool:
        try:
            import webrtcvad
        except ImportError:
            if not SilenceChecker._warned:
                logger.warning("webrtcvad not installed: speech/silence gate skipped")
                SilenceChecker._warned = True
            return False
        pcm = (np.asarray(data) * 32767).astype(np.int16).tobytes()
        vad = webrtcvad.Vad(self.aggr)
        step = int(self.sample_rate * self.frame_ms / 1000) * 2
        voiced = sum(vad.is_speech(pcm[i:i + step], self.sample_rate)
                     for i in range(0, len(pcm) - step + 1, step))
        return voiced * (self.frame_ms / 1000.0) < self.min_speech_seconds


__all__ = ["SilenceChecker"]
