"""HiFi-GAN mel front-end — mirrors pytorch_sound/interface/hifi_gan.py:10-63.

Only the ENCODER side (wave -> mel) is the hot path; `InterfaceHifiGAN.decode` (the vocoder) stays in
the reference.  `MelSpectrogram` here has the reference's constructor, buffers (`mel_filter`,
`window`) and output: (B, L) -> (B, num_mels, L // hop_size) = ln(max(mel_filter @ sqrt(re^2+im^2+1e-9), 1e-5)).
"""
import torch

from .. import _lib, functional
from ..models.transforms import _PlanUser


class AudioParameters:
    """interface/hifi_gan.py:10-18."""
    sampling_rate: int = 22050
    n_fft: int = 1024
    window_size: int = 1024
    hop_size: int = 256
    num_mels: int = 80
    fmin: float = 0.
    fmax: float = 8000.


def audio_parameters() -> dict:
    """vars(AudioParameters()) is empty for class attributes, so the reference's
    `MelSpectrogram(**vars(AudioParameters()))` (interface/hifi_gan.py:98) always builds the defaults."""
    return dict(vars(AudioParameters()))


class MelSpectrogram(_PlanUser):
    """Drop-in for pytorch_sound.interface.hifi_gan.MelSpectrogram (interface/hifi_gan.py:29-63)."""

    _fb_buffer_name = 'mel_filter'

    def __init__(self, sampling_rate: int = 22050, n_fft: int = 1024, window_size: int = 1024, hop_size: int = 256,
                 num_mels: int = 80, fmin: float = 0., fmax: float = 8000.):
        super().__init__()
        self.n_fft = n_fft
        self.hop_size = hop_size
        self.window_size = window_size
        self.pad_size = (self.n_fft - self.hop_size) // 2
        mel_filter = _lib.mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax)
        self.register_buffer('mel_filter', torch.from_numpy(mel_filter))
        self._fb_default = torch.from_numpy(mel_filter).clone()
        self.register_buffer('window', torch.from_numpy(_lib.hann_window(window_size)))
        self._plan_kwargs = dict(sample_rate=sampling_rate, n_fft=n_fft, win_length=window_size, hop_length=hop_size,
                                 n_mels=num_mels, fmin=fmin, fmax=fmax, pad_mode=_lib.PAD_HIFI, mag_eps=1e-9)
        self._center_kwargs = dict(self._plan_kwargs, pad_mode=_lib.PAD_CENTER)
        self._epi = _lib.make_epilogue(_lib.LOG_LN_FLOOR, 1e-5)

    def forward(self, wav: torch.Tensor, is_center: bool = False) -> torch.Tensor:
        if is_center:
            if not wav.is_cuda:
                raise RuntimeError(functional.NO_CPU_MSG)
            # the reference pads pad_size by reflection and THEN lets torch.stft centre-pad n_fft//2 more
            # (interface/hifi_gan.py:48-54): materialise the first pad, then run the centred plan
            wav = torch.nn.functional.pad(wav.unsqueeze(1), [self.pad_size, self.pad_size], mode='reflect').squeeze(1)
            idx = wav.device.index if wav.device.index is not None else torch.cuda.current_device()
            plan = _lib.cached_plan(idx, **self._center_kwargs)
            self._fb_sync()
            if self._fb_dirty:
                raise NotImplementedError("is_center=True with a state_dict-supplied mel_filter")
        else:
            plan = self._plan(wav.device)
        mel, _, _ = functional.run(plan, wav, self._epi)
        return mel
