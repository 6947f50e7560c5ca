"""Waveform-side operators of pytorch_sound/models/sound.py that sit directly in front of the spectral path.

    reference symbol (file:line)                  here
    PreEmphasis.forward  models/sound.py:66-81    PreEmphasis.forward (one HBM-bound sm_100a kernel)
    build_stft_functions models/sound.py:106-117  build_stft_functions (plans are cached per geometry, no per-call rebuild)
    multi_stft_loss      models/sound.py:120-147  multi_stft_loss (forward value: magnitude kernels + one reduction pass)

CUDA float32 only; there is no CPU path.  InversePreEmphasis (an RNN used at synthesis time) is outside the
feature-extraction path.  multi_stft_loss is the forward VALUE (validation metric / monitoring); training through it
needs autograd, which stays with the reference implementation (see patch.py: inputs that require grad run it).
"""
from typing import List, Tuple

import torch

from .. import functional
from .transforms import STFTTorchAudio as STFT  # the reference imports it under this name (models/sound.py:3)


class PreEmphasis(torch.nn.Module):
    """Drop-in for pytorch_sound.models.sound.PreEmphasis: y[n] = x[n] - coef x[n-1], input and output (B, 1, L).

    The reference pads one reflected sample on the left (`F.pad(input, (1, 0), 'reflect')`, models/sound.py:80) and
    runs a 2-tap conv1d with the registered `flipped_filter` buffer; the buffer is kept (same name, shape and
    values) so state_dicts stay loadable, the work is one streaming kernel."""

    def __init__(self, coef: float = 0.97):
        super().__init__()
        self.coef = coef
        self.register_buffer('flipped_filter', torch.FloatTensor([-self.coef, 1.]).unsqueeze(0).unsqueeze(0))

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        assert len(input.size()) == 3, 'The number of dimensions of input tensor must be 3!'
        if input.shape[1] != 1:
            raise ValueError("PreEmphasis expects (B, 1, L): the reference's conv1d filter has one input channel")
        if input.shape[2] < 2:
            raise ValueError("PreEmphasis: reflect padding needs L >= 2")
        return functional.preemphasis(input[:, 0], self.coef).unsqueeze(1)


_stft_cache = {}


def build_stft_functions(*params: Tuple[int, int, int]):
    """STFT modules for (n_fft, window size, hop size) triplets (models/sound.py:106-117).  The reference rebuilds
    them and calls .cuda() on every loss evaluation; here they are built once per geometry
    (the modules are stateless: a plan per device, cached in the library binding)."""
    out = []
    for fft, win, hop in params:
        key = (int(fft), int(win), int(hop))
        if key not in _stft_cache:
            _stft_cache[key] = STFT(win, hop, win, fft)
        out.append(_stft_cache[key])
    return out


def multi_stft_loss(pred: torch.Tensor, target: torch.Tensor, stft_params: List[Tuple[int, int, int]],
                    eps: float = 1e-5) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Multi-resolution STFT loss value (models/sound.py:120-147): per resolution the spectral-convergence term
    mean_b ||t - p||_F / ||t||_F and the log-magnitude term mean_b ||ln(t + eps) - ln(p + eps)||_1 / (F T);
    returns (sum of both, spectral convergence, magnitude), each averaged over the resolutions, as 0-dim CUDA tensors.

    pred / target: (N, T) CUDA float32.  Every resolution is two launches of the magnitude kernel (no phase pass,
    unlike the reference's transform()[0]) and one streaming reduction over both magnitude tensors; n_fft must be a
    power of two <= 2048 (the usual (1024, 600, 120), (2048, 1200, 240), (512, 240, 50) set qualifies)."""
    if pred.shape != target.shape:
        raise ValueError(f"pred {tuple(pred.shape)} and target {tuple(target.shape)} differ")
    funcs = build_stft_functions(*stft_params)
    acc = torch.zeros(2, device=pred.device, dtype=torch.float32)
    for f in funcs:
        functional.stft_loss_terms(f.magnitude(pred), f.magnitude(target), eps, acc)
    acc = acc / float(len(funcs))
    return acc[0] + acc[1], acc[0], acc[1]
