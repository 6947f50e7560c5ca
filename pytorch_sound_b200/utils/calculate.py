"""dB <-> log helpers, mel normalisation and RMS volume normalisation — mirrors
pytorch_sound/utils/calculate.py:10-63.

These are scalar/elementwise host helpers; `norm_mel` is ALSO available fused into the kernel
epilogue (LogMelSpectrogram(..., ).forward(..., norm=True) / the `norm_mel` flag of b200mel_epilogue).
"""
from typing import Union

import numpy as np
import torch

from .. import settings

TensorOrArr = Union[torch.Tensor, np.ndarray]


def db2log(db: TensorOrArr) -> TensorOrArr:
    """ln(10 ** (db / 10)) — utils/calculate.py:10-19."""
    if isinstance(db, torch.Tensor):
        return torch.log(torch.pow(10, db / 10.))
    return np.log(np.power(10, np.asarray(db, dtype=np.float64) / 10))


def _bounds():
    return float(db2log(settings.MIN_DB)), float(db2log(settings.MAX_DB))


def unnorm_mel(x: TensorOrArr) -> TensorOrArr:
    """utils/calculate.py:22-29."""
    mel_min, mel_max = _bounds()
    return ((x + 1) / 2) * (mel_max - mel_min) + mel_min


def norm_mel(x: TensorOrArr) -> TensorOrArr:
    """utils/calculate.py:32-43."""
    mel_min, mel_max = _bounds()
    if type(x) == np.ndarray:
        x = x.clip(mel_min, mel_max)
    else:
        x = x.clamp(mel_min, mel_max)
    return (x - mel_min) / (mel_max - mel_min) * 2 - 1


def volume_norm_log(x: np.ndarray, target_db: float = -11.5) -> np.ndarray:
    """RMS volume normalisation of a numpy waveform (loader-worker side) — utils/calculate.py:46-53.
    Like the reference it applies the 10^(dB/10) power ratio to an amplitude."""
    return x / (np.std(x) / 10 ** (target_db / 10))


def volume_norm_log_torch(x: torch.Tensor, target_db: float = -11.5) -> torch.Tensor:
    """utils/calculate.py:56-63 on a CUDA tensor: x / (torch.std(x) / 10^(target_db/10)), std over the whole
    tensor (unbiased).  Two streaming kernels (double-precision moments, scale); no CPU path."""
    from .. import functional

    return functional.volume_norm(x, target_db)
