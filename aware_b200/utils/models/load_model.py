"""load() -> (embedder, detector), as the reference's utils/models/load_model.py:6-76:
reads cards/config.yaml, builds the embedder, and a detector that SHARES the
embedder's network (and here also its device engine)."""
from pathlib import Path

from ...detection import AWAREDetector
from ...embedding import AWAREEmbedder
from ..logger import logger
from ..utils import load_config

CARDS_DIR = Path(__file__).resolve().parent.parent.parent / "cards"


def load(config_path=None):
    try:
        config = load_config(config_path or CARDS_DIR / "config.yaml")
    except Exception as e:                                   # noqa: BLE001  (reference behaviour)
        logger.error(f"Error loading configs: {e}")
        return
    try:
        embedder = AWAREEmbedder(
            frame_length=config.get("frame_length", 1024), hop_length=config.get("hop_length", 256),
            window=config.get("window", "hann"), win_length=config.get("win_length", 1024),
            pattern_mode=config.get("pattern_mode", "bits2bipolar"),
            embedding_bands=tuple(config.get("embedding_bands", [500, 4000])),
            tolerance_db=config.get("tolerance_db", 6.0), num_iterations=config.get("num_iterations", 400),
            detection_net_cfg=config.get("detection_net_cfg", {}),
            optimizer_cfg=config.get("optimizer_cfg", {"name": "nadam", "params": {"lr": 0.1}}),
            scheduler_cfg=config.get("scheduler_cfg", {"name": "reduce_lr_on_plateau",
                                                       "params": {"factor": 0.9, "patience": 500}}),
            loss=config.get("loss", "push_extremes"), verbose=config.get("verbose", True),
            precision=config.get("precision", "tf32"), wave_clips=config.get("wave_clips", 0),
            embed_precision=config.get("embed_precision", "fp16"))
        embedder.enforce_16k = bool(config.get("enforce_16k", True))
        embedder.threshold = config.get("threshold", 0.0)
        embedder.vad_gate = bool(config.get("vad_gate", True))    # absent key = upstream behaviour (gate on)
        embedder.exact_margin = float(config.get("exact_margin", 1e-3))
    except Exception as e:                                   # noqa: BLE001
        logger.error(f"Error creating embedder: {e}")
        return
    try:
        detector = AWAREDetector(
            model=embedder.detection_net, threshold=config.get("threshold", 0.0),
            frame_length=config.get("frame_length", 1024), hop_length=config.get("hop_length", 256),
            window=config.get("window", "hann"), win_length=config.get("win_length", 1024),
            pattern_mode=config.get("pattern_mode", "bipolar"),
            embedding_bands=tuple(config.get("embedding_bands", [500, 4000])),
            precision=config.get("precision", "tf32"),
            engine_owner=embedder)     # one context / weights / workspace for both (shared net upstream)
        detector.enforce_16k = bool(config.get("enforce_16k", True))
    except Exception as e:                                   # noqa: BLE001
        logger.error(f"Error creating detector: {e}")
        return
    return embedder, detector
