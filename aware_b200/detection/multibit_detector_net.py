"""AWAREDetectorNet -- host-side description of the detector (weights, mel basis).

Mirrors the constructor signature and attributes of the reference class
(detection/multibit_detector_net.py:15-80 there).  It is not an nn.Module that
executes: the forward (and input-gradient) pass runs in the CUDA kernels; this
object only materialises the parameters -- the mel basis and the
xavier-uniform conv weights drawn under the reference's fixed seed -- and hands
them to the device engine.
"""
from __future__ import annotations

import numpy as np
import torch

INIT_SEED = 328656719      # torch.manual_seed in the reference constructor (:78)


def _slaney_hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f / (200.0 / 3)
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore"):
        log = 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) / logstep
    return np.where(f >= 1000.0, log, lin)


def _slaney_mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    logstep = np.log(6.4) / 27.0
    return np.where(m >= 15.0, 1000.0 * np.exp(logstep * (m - 15.0)), (200.0 / 3) * m)


def slaney_mel_basis(sample_rate: int, n_fft: int, n_mels: int) -> np.ndarray:
    """librosa-compatible Slaney mel filter bank, float32 (n_mels, 1 + n_fft//2);
    same construction as the reference's detection/modules/mel.py:105-149."""
    n_freq = 1 + n_fft // 2
    fft_f = np.linspace(0.0, sample_rate / 2.0, n_freq)
    edges = _slaney_mel_to_hz(np.linspace(_slaney_hz_to_mel(0.0), _slaney_hz_to_mel(sample_rate / 2.0),
                                          n_mels + 2))
    width = np.diff(edges)
    ramps = edges[:, None] - fft_f[None, :]
    basis = np.zeros((n_mels, n_freq), dtype=np.float32)
    for i in range(n_mels):
        basis[i] = np.maximum(0.0, np.minimum(-ramps[i] / width[i], ramps[i + 2] / width[i + 1]))
    basis *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return basis


class AWAREDetectorNet:
    def __init__(self, sample_rate: int = 16000, n_fft: int = 1024, n_mels: int = 128,
                 initial_pool_size: int = 2, initial_pool_stride: int = 2, num_blocks: int = 3,
                 n_filters=(512, 1024, 1024), kernel_size: int = 1, stride: int = 1, padding: int = 0,
                 norm_layer: str = "instance", activation: str = "leaky_relu", output_length: int = 20,
                 final_activation: str = "tanh"):
        n_filters = list(n_filters)
        assert len(n_filters) == num_blocks, "Number of filters must match number of blocks"
        fixed = dict(n_fft=1024, n_mels=128, initial_pool_size=2, initial_pool_stride=2, num_blocks=3,
                     n_filters=[512, 1024, 1024], kernel_size=1, stride=1, padding=0,
                     norm_layer="instance", activation="leaky_relu", output_length=20,
                     final_activation="tanh")
        given = dict(n_fft=n_fft, n_mels=n_mels, initial_pool_size=initial_pool_size,
                     initial_pool_stride=initial_pool_stride, num_blocks=num_blocks, n_filters=n_filters,
                     kernel_size=kernel_size, stride=stride, padding=padding, norm_layer=norm_layer,
                     activation=activation, output_length=output_length, final_activation=final_activation)
        bad = {k: v for k, v in given.items() if fixed[k] != v}
        if bad:
            raise ValueError("aware_b200 kernels are specialised for the released architecture; "
                             f"unsupported detection_net_cfg entries: {bad}")
        self.sample_rate, self.n_fft, self.n_mels = sample_rate, n_fft, n_mels
        self.num_blocks, self.initial_pool_size = num_blocks, initial_pool_size
        self.output_length, self.final_activation = output_length, final_activation
        self.channels = [n_mels] + n_filters + [2 * output_length]
        self.mel_filter_bank = slaney_mel_basis(sample_rate, n_fft, n_mels)
        # Conv1d(k=1) weights, xavier_uniform_ in module order under the fixed seed; biases are 0
        # and are cancelled by the InstanceNorm that follows each conv.  Like the reference this
        # resets the global torch RNG (load_model side effect).
        torch.manual_seed(INIT_SEED)
        self.weights = []
        for cin, cout in zip(self.channels[:-1], self.channels[1:]):
            w = torch.empty(cout, cin, 1)
            torch.nn.init.xavier_uniform_(w)
            self.weights.append(w[:, :, 0].contiguous().numpy())

    def eval(self):
        return self

    def to(self, device):
        return self

    def parameters(self):
        return []

    def get_model_info(self):
        total = sum(w.size + w.shape[0] for w in self.weights)
        return {"sample_rate": self.sample_rate, "n_fft": self.n_fft, "n_mels": self.n_mels,
                "num_blocks": self.num_blocks, "output_length": self.output_length,
                "final_activation": self.final_activation, "total_parameters": total,
                "trainable_parameters": total}
