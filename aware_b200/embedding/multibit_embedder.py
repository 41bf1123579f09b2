"""AWAREEmbedder with the reference's interface (embedding/multibit_embedder.py there):
`embed(audio, sample_rate, watermark) -> float32[256*(T-1)]`, plus `embed_batch`.

Only the configured optimisation (NAdam lr 0.1, push_extremes loss; the plateau
scheduler with patience 500 can never fire in <= 500 iterations) has kernels."""
from __future__ import annotations

import numpy as np
import torch

from ..detection.multibit_detector_net import AWAREDetectorNet
from ..utils.logger import logger
from ..utils.utils import to_tensor


class AWAREEmbedder:
    def __init__(self, frame_length: int = 1024, hop_length: int = 256, window: str = "hann",
                 win_length: int = 1024, pattern_mode: str = "bits2bipolar", embedding_bands=(500, 4000),
                 tolerance_db: float = 6.0, num_iterations: int = 400, detection_net_cfg: dict = None,
                 optimizer_cfg: dict = None, scheduler_cfg: dict = None, loss: str = "push_extremes",
                 verbose: bool = True, precision: str = "tf32", wave_clips: int = 0,
                 embed_precision: str = "fp16"):
        if (frame_length, hop_length, win_length, window) != (1024, 256, 1024, "hann"):
            raise ValueError("aware_b200 kernels are specialised for n_fft=1024, hop=256, hann")
        optimizer_cfg = optimizer_cfg or {"name": "nadam", "params": {"lr": 0.1}}
        scheduler_cfg = scheduler_cfg or {"name": "reduce_lr_on_plateau", "params": {"factor": 0.9, "patience": 500}}
        if optimizer_cfg["name"] != "nadam" or float(optimizer_cfg["params"].get("lr", 0.1)) != 0.1:
            raise ValueError("aware_b200 implements the configured optimiser only: nadam, lr=0.1")
        if loss not in ("push_extremes", "push"):
            raise ValueError("aware_b200 implements the configured loss only: push_extremes")
        if scheduler_cfg["name"] == "reduce_lr_on_plateau" and \
                int(scheduler_cfg["params"].get("patience", 500)) < num_iterations:
            raise ValueError("plateau scheduler that can fire is not implemented on device")
        self.frame_length, self.hop_length = frame_length, hop_length
        self.embedding_bands = tuple(embedding_bands)
        self.tolerance_db = tolerance_db
        self.num_iterations = num_iterations
        self.pattern_mode = pattern_mode
        self.detection_net = AWAREDetectorNet(**(detection_net_cfg or {}))
        self.optimizer_name, self.optimizer_params = optimizer_cfg["name"], optimizer_cfg["params"]
        self.scheduler_name, self.scheduler_params = scheduler_cfg["name"], scheduler_cfg["params"]
        self.loss = loss
        self.verbose = verbose
        self.precision = precision              # detector GEMMs
        # GEMMs inside the 400-step optimisation loop: "fp16" = tcgen05 kind::f16 with fp32
        # accumulation -- TF32's 10-bit mantissa at half the bytes (measured: not less accurate)
        self.embed_precision = embed_precision
        self.wave_clips = wave_clips
        self.vad_gate = False                   # speech / silence gate (needs webrtcvad), see load()
        self.exact_margin = 1e-3                # detect: exact re-evaluation margin (engine option)
        self.check_finite = True                # 16-bit loops: re-run flagged clips in TF32
        self.threshold = 0.0
        self._engine = None
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")

    @property
    def engine(self):
        if self._engine is None:
            from ..engine import Engine
            self._engine = Engine(self.detection_net.weights, self.detection_net.mel_filter_bank,
                                  torch.hann_window(1024).numpy(), bands=self.embedding_bands,
                                  tolerance_db=self.tolerance_db, threshold=self.threshold,
                                  precision=self.precision)
            self._engine.set_exact_margin(self.exact_margin)
        return self._engine

    def embed_batch(self, audio, sample_rate: int, watermark, scale=None) -> torch.Tensor:
        """[n, N] audio + [n, 20] (or [20]) bipolar watermark -> CUDA float32 [n, 256*(N//256)]."""
        x = to_tensor(audio)
        if x.dim() != 2:
            raise ValueError("embed_batch expects [n_clips, n_samples]")
        x = x.to(self.engine.device, non_blocking=True).contiguous()
        wm = torch.as_tensor(np.asarray(watermark)).to(torch.int32)
        if wm.dim() == 1:
            wm = wm.unsqueeze(0).expand(x.shape[0], -1)
        if self.verbose:
            _, nb = self.engine.band_bins(sample_rate)
            logger.info(f"Starting optimization with {nb * (1 + x.shape[1] // 256)} variables per clip, "
                        f"{x.shape[0]} clip(s), {self.num_iterations} iterations")
        eng = self.engine
        out = eng.embed(x, sample_rate, wm.contiguous(), iters=self.num_iterations, scale=scale,
                        wave_clips=self.wave_clips, precision=self.embed_precision)
        if self.check_finite and (self.embed_precision or eng.precision) in ("fp16", "bf16"):
            # a 16-bit loop skips (and flags) updates whose gradient overflowed; such a clip is
            # embedded again with TF32 loop GEMMs instead of being returned half-optimised
            bad = torch.nonzero(eng.embed_status()).flatten()
            if bad.numel():
                logger.warning(f"{bad.numel()} clip(s) met a non-finite gradient in the "
                               f"{self.embed_precision} loop; re-embedding them with TF32 GEMMs")
                sc = scale[bad] if isinstance(scale, torch.Tensor) else scale
                out[bad] = eng.embed(x[bad].contiguous(), sample_rate, wm[bad].contiguous(),
                                     iters=self.num_iterations, scale=sc, wave_clips=self.wave_clips,
                                     precision="tf32")
        return out

    def embed(self, audio: np.ndarray, sample_rate: int, watermark: np.ndarray) -> np.ndarray:
        x = to_tensor(audio).reshape(1, -1)
        return self.embed_batch(x, sample_rate, watermark)[0].cpu().numpy()
