"""Waveform-side operators of pytorch_sound/models/sound.py that sit directly in front of the spectral path.

    reference symbol (file:line)                  here
    PreEmphasis.forward  models/sound.py:66-81    PreEmphasis.forward (one HBM-bound sm_100a kernel)

CUDA float32 only; there is no CPU path.  InversePreEmphasis (an RNN used at synthesis time) and the loss
helpers of that file are outside the feature-extraction path.
"""
import torch

from .. import functional


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
