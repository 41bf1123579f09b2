"""PatternEncoder: watermark payload -> detector target (reference
utils/watermark/encoder.py).  The device path only consumes 'bits2bipolar'
(config.yaml:9); the byte modes stay host-side."""
import numpy as np


class PatternEncoder:
    MODES = ("bits2bipolar", "bytes2bipolar", "bytes2bits", "bits")

    def __init__(self, mode: str = "bits2bipolar"):
        self.mode = mode

    def __call__(self, inputs):
        if self.mode == "bits2bipolar":
            return self._bits_to_bipolar(inputs)
        if self.mode == "bytes2bipolar":
            return self._bits_to_bipolar(self._bytes_to_bits(inputs))
        if self.mode == "bytes2bits":
            return self._bytes_to_bits(inputs)
        if self.mode == "bits":
            return inputs
        raise ValueError(f"Invalid mode: {self.mode}")

    @staticmethod
    def _bytes_to_bits(data: bytes) -> np.ndarray:
        return np.unpackbits(np.frombuffer(bytes(data), dtype=np.uint8)).astype(np.int32)

    @staticmethod
    def _bits_to_bipolar(bits) -> np.ndarray:
        return (2 * np.asarray(bits).astype(np.int64) - 1).astype(np.int32)
