"""aware_b200 -- B200-native (sm_100a) implementation of the AWARE audio-watermarking
hot path: batched embed -> attack -> detect -> BER behind the reference's Python API.

    from aware_b200.utils.models import load
    from aware_b200.service import embed_watermark, detect_watermark
    from aware_b200.metrics.audio import BER, SNR

`install_as_aware()` registers the package under the reference's import name so
unmodified callers (`from aware.service import embed_watermark`) use this path.
"""
import importlib
import sys

__version__ = "0.1.0"

_SUBMODULES = ["service", "service.embed", "service.detect", "utils", "utils.models", "utils.watermark",
               "utils.audio", "utils.logger", "utils.utils", "embedding", "embedding.multibit_embedder",
               "detection", "detection.multibit_detector", "detection.multibit_detector_net", "metrics",
               "metrics.audio"]


def install_as_aware():
    """Alias aware_b200[.x] as aware[.x] in sys.modules (drop-in for the reference's imports)."""
    me = importlib.import_module(__name__)
    sys.modules["aware"] = me
    for sub in _SUBMODULES:
        sys.modules["aware." + sub] = importlib.import_module(__name__ + "." + sub)
    return me
