from .audio import BER, PESQ, SNR, STOI

__all__ = ["BER", "SNR", "PESQ", "STOI"]
