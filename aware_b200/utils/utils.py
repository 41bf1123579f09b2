import os

import numpy as np
import torch
import yaml


def load_config(path):
    """YAML -> dict (reference utils/utils.py:5-14)."""
    with open(os.fspath(path), "r") as f:
        return yaml.safe_load(f)


def to_tensor(x):
    """Anything array-like -> float32 torch tensor (reference utils/utils.py:16-23)."""
    if isinstance(x, torch.Tensor):
        return x.float()
    return torch.from_numpy(np.asarray(x)).float()
