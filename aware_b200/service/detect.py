"""detect_watermark with the reference's signature and error behaviour
(service/detect.py:7-55 there), plus detect_watermark_batch."""
import numpy as np
import torch

from ..utils.logger import logger
from ..utils.watermark import PatternDecoder


def _check_rate(sample_rate, detector):
    if sample_rate != 16000 and getattr(detector, "enforce_16k", True):
        logger.error(f"Invalid sample rate. Expected 16000Hz, got {sample_rate}Hz.")
        raise ValueError("Invalid sample rate. Expected 16000Hz.")


def detect_watermark(audio: np.ndarray, sample_rate: int, detector):
    decode = PatternDecoder(encoder_mode=detector.pattern_mode, threshold=detector.threshold)
    _check_rate(sample_rate, detector)
    audio = np.asarray(audio)
    if audio.ndim == 2 and audio.shape[1] == 2:                       # stereo: per-bit larger |v|
        v = detector.detect_batch(np.stack([audio[:, 0], audio[:, 1]]), sample_rate).cpu().numpy()
        left, right = v[0], v[1]
        return decode(np.where(np.abs(left) > np.abs(right), left, right))
    if audio.ndim == 1:
        return decode(detector.detect(audio, sample_rate))
    logger.error("Invalid audio shape. Expected 1D or 2D numpy array.")
    raise ValueError("Invalid audio shape. Expected 1D or 2D numpy array.")


def detect_watermark_batch(audio, sample_rate: int, detector, ref_bits=None, counters=None):
    """audio [n, N] -> int32 bits [n, 20] on the device.  With `ref_bits` also returns the
    per-clip error counts and adds {errors, bits, clips} to `counters` (int64[3], device)."""
    _check_rate(sample_rate, detector)
    values = detector.detect_batch(audio, sample_rate)
    if ref_bits is not None and not isinstance(ref_bits, torch.Tensor):
        ref_bits = torch.as_tensor(np.asarray(ref_bits))
    return detector.engine.decide(values, ref_bits, counters, threshold=detector.threshold)
