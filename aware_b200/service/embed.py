"""embed_watermark with the reference's signature and error behaviour
(service/embed.py:7-80 there), plus embed_watermark_batch for [n, N] batches."""
import numpy as np
import torch

from ..utils.audio import silent_mask
from ..utils.logger import logger
from ..utils.watermark import PatternEncoder


def _check_rate(sample_rate, model):
    if sample_rate != 16000 and getattr(model, "enforce_16k", True):
        logger.error(f"Invalid sample rate. Expected 16000Hz, got {sample_rate}Hz.")
        raise ValueError("Invalid sample rate. Expected 16000Hz.")


def _encode(watermark_bits, model):
    watermark = PatternEncoder(mode=model.pattern_mode)(watermark_bits)
    if len(watermark) != model.detection_net.output_length:
        logger.error(f"Invalid watermark length. Expected {model.detection_net.output_length}, got {len(watermark)}.")
        raise ValueError("Invalid watermark length.")
    return watermark


_SILENT = ("Signal you provided doesn't contain any speach. Please provide signal that contains speach.")


def embed_watermark(audio: np.ndarray, sample_rate: int, watermark_bits, model) -> np.ndarray:
    _check_rate(sample_rate, model)
    watermark = _encode(watermark_bits, model)
    audio = np.asarray(audio)
    if audio.ndim == 2 and audio.shape[1] == 2:                       # stereo: per channel
        chans = [audio[:, 0], audio[:, 1]]
        if silent_mask(chans, sample_rate, model).all():
            logger.error(_SILENT)
            raise ValueError(_SILENT)
        mx = np.array([np.max(c) for c in chans], dtype=np.float32)     # signed max (embed.py:41-42)
        out = model.embed_batch(np.stack(chans), sample_rate, watermark).cpu().numpy()
        return np.column_stack((mx[0] * out[0], mx[1] * out[1]))
    if audio.ndim == 1 or (audio.ndim == 2 and audio.shape[1] == 1):  # mono
        if silent_mask([audio.reshape(-1)], sample_rate, model)[0]:
            logger.error(_SILENT)
            raise ValueError(_SILENT)
        audio_mx = np.max(audio)
        return audio_mx * model.embed(audio, sample_rate, watermark)
    logger.error("Invalid audio shape. Expected 1D or 2D numpy array.")
    raise ValueError("Invalid audio shape. Expected 1D or 2D numpy array.")


def embed_watermark_batch(audio, sample_rate: int, watermark_bits, model) -> torch.Tensor:
    """audio [n, N] (numpy / tensor, host or device), watermark_bits [n, 20] or [20] of 0/1.
    Returns a CUDA tensor [n, 256*(N//256)] already rescaled by each clip's signed max (taken on
    the device in the same pass as the peak).  With the VAD gate on (`model.vad_gate`) a silent clip
    raises ValueError as upstream does per clip; pre-filter with `utils.audio.silent_mask`."""
    _check_rate(sample_rate, model)
    if getattr(model, "vad_gate", False):
        host = audio.detach().cpu().numpy() if isinstance(audio, torch.Tensor) else np.asarray(audio)
        if silent_mask(list(host), sample_rate, model).any():
            logger.error(_SILENT)
            raise ValueError(_SILENT)
    bits = np.asarray(watermark_bits.cpu() if isinstance(watermark_bits, torch.Tensor) else watermark_bits)
    wm = np.stack([_encode(b, model) for b in np.atleast_2d(bits)])
    x = audio if isinstance(audio, torch.Tensor) else torch.from_numpy(np.asarray(audio))
    x = x.float().to(model.engine.device, non_blocking=True)
    if wm.shape[0] == 1 and x.shape[0] > 1:
        wm = np.repeat(wm, x.shape[0], axis=0)
    return model.embed_batch(x, sample_rate, wm, scale="signed_max")
