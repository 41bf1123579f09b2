"""AWAREDetector with the reference's interface (detection/multibit_detector.py there):
`detect(audio, sample_rate) -> float32[20]`, plus a batched `detect_batch`."""
from __future__ import annotations

import numpy as np
import torch

from ..utils.utils import to_tensor


class AWAREDetector:
    def __init__(self, model, threshold: float = 0.0, frame_length: int = 1024, hop_length: int = 256,
                 window: str = "hann", win_length: int = 1024, pattern_mode: str = "bits2bipolar",
                 embedding_bands=(500, 4000), engine_owner=None, precision: str = "tf32"):
        if (frame_length, hop_length, win_length, window) != (1024, 256, 1024, "hann"):
            raise ValueError("aware_b200 kernels are specialised for n_fft=1024, hop=256, hann")
        self.threshold = threshold
        self.pattern_mode = pattern_mode
        self.embedding_bands = tuple(embedding_bands)
        self.win_length = self.frame_length = frame_length
        self.hop_length = hop_length
        self.detection_net = model
        self._engine = None
        self._engine_owner = engine_owner     # object whose .engine is shared (the embedder)
        self._precision = precision
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")

    @property
    def precision(self):
        """GEMM arithmetic of the detector stack.  With a shared engine (load()) this IS the engine's
        setting: there is one context, so embedder and detector cannot disagree silently."""
        if self._engine_owner is not None and self._engine_owner._engine is not None:
            return self._engine_owner.engine.precision
        return self._precision

    @precision.setter
    def precision(self, value):
        self._precision = value
        if self._engine_owner is not None:
            self._engine_owner.engine.set_precision(value)
        elif self._engine is not None:
            self._engine.set_precision(value)

    @property
    def engine(self):
        if self._engine_owner is not None:
            return self._engine_owner.engine
        if self._engine is None:
            from ..engine import Engine
            self._engine = Engine(self.detection_net.weights, self.detection_net.mel_filter_bank,
                                  torch.hann_window(1024).numpy(), bands=self.embedding_bands,
                                  threshold=self.threshold, precision=self._precision)
        return self._engine

    def detect_batch(self, audio, sample_rate: int) -> torch.Tensor:
        """[n, N] (numpy or tensor, any float dtype) -> CUDA float32 [n, 20]."""
        x = to_tensor(audio)
        if x.dim() != 2:
            raise ValueError("detect_batch expects [n_clips, n_samples]")
        eng = self.engine
        x = x.to(eng.device, non_blocking=True).contiguous()
        eng.set_threshold(self.threshold)       # a shared engine follows THIS detector's threshold
        return eng.detect(x, sample_rate)

    def detect(self, audio: np.ndarray, sample_rate: int) -> np.ndarray:
        x = to_tensor(audio).reshape(1, -1)
        return self.detect_batch(x, sample_rate)[0].cpu().numpy()
