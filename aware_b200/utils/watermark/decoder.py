"""PatternDecoder: detector outputs -> payload (reference utils/watermark/decoder.py).
Strict '>' against the threshold (decoder.py:51,63 there)."""
import numpy as np


class PatternDecoder:
    def __init__(self, threshold: float = 0.5, encoder_mode: str = "bits2bipolar"):
        self.threshold = threshold
        self.encoder_mode = encoder_mode

    def __call__(self, detected_values: np.ndarray):
        v = np.asarray(detected_values)
        if self.encoder_mode == "bits2bipolar":
            return self._bipolar_to_bits(self._detect_bipolar(v))
        if self.encoder_mode == "bytes2bipolar":
            return self._bits_to_bytes(self._bipolar_to_bits(self._detect_bipolar(v)))
        if self.encoder_mode == "bytes2bits":
            return self._bits_to_bytes(self._detect_binary(v))
        if self.encoder_mode == "bits":
            return self._detect_binary(v)
        raise ValueError(f"Invalid mode: {self.encoder_mode}")

    def _detect_binary(self, v):
        return (v > self.threshold).astype(np.int32)

    def _detect_bipolar(self, v):
        return 2 * (v > self.threshold).astype(np.int32) - 1

    @staticmethod
    def _bipolar_to_bits(v):
        return (v > 0).astype(np.int32)

    @staticmethod
    def _bits_to_bytes(bits) -> bytes:
        return bytes(int(b) for b in bits)
