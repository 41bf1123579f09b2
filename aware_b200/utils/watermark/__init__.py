from .decoder import PatternDecoder
from .encoder import PatternEncoder

__all__ = ["PatternEncoder", "PatternDecoder"]
