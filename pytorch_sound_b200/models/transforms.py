"""Spectral operators with the names, constructor arguments, buffers and outputs of
pytorch_sound/models/transforms.py — computed by the fused sm_100a kernel (libb200mel.so).

    reference symbol (file:line)                         here
    STFT.transform            transforms.py:53-69        STFT.transform           -> (mag, phase)
    LogMelSpectrogram.forward transforms.py:231-244      LogMelSpectrogram.forward(wav, log_offset=1e-6)
    STFTTorchAudio.forward    transforms.py:297-303      STFTTorchAudio.forward   -> (real, imag)
    STFTTorchAudio.transform  transforms.py:305-311      STFTTorchAudio.transform -> (mag, phase)
    Audio2Mel.forward         transforms.py:351-366      Audio2Mel.forward(audio (B,1,L))
    LogMelSpectrogramTorchAudio.forward transforms.py:387-394  same kernel, power / HTK / un-normalised filterbank
    MelToMFCC.forward         transforms.py:428-430      MelToMFCC.forward (DCT kernel on the mel tensor)
    MFCC.forward              transforms.py:451-455      LogMelSpectrogram + the DCT kernel

Inputs must be CUDA float32; there is no CPU path.  The modules are forward-only (feature
extraction); the reference's trainable / synthesis-direction classes (LearnableSTFT, PQMF,
STFT.inverse's use in vocoding) are outside the hot path and are not provided here.
"""
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, functional


def _db_to_log(db: Optional[float]) -> Optional[float]:
    # `if min_db:` truthiness as in the reference (transforms.py:222-229): 0 / None mean "no clamp"
    if db:
        return float(np.log(np.power(10.0, db / 10.0)))
    return None


class _PlanUser(nn.Module):
    """Shared plumbing: plan lookup per device, custom-filterbank handling.

    The kernel applies the filterbank of its PLAN (device tables), not the registered torch buffer, so the buffer
    is watched: whenever it is replaced or written in place (load_state_dict, `.copy_`, `.to`) — detected through
    the tensor's identity and `_version` counter, two integer compares per call — it is compared with the default
    filterbank of the geometry; a differing buffer gets a private plan carrying exactly those weights, and going
    back to the default goes back to the shared plan."""

    _fb_buffer_name: Optional[str] = None

    def __init__(self):
        super().__init__()
        self._plan_kwargs = {}
        self._private_plans = {}
        self._fb_dirty = False
        self._fb_default: Optional[torch.Tensor] = None
        self._fb_seen = None      # (id, _version) of the buffer the current _fb_dirty / private plans belong to
        self._fb_owner = None     # patch.py hybrids: the reference module whose buffer is the one to watch

    def _fb_tensor(self) -> Optional[torch.Tensor]:
        """The registered filterbank as (n_mels, n_freq), or None for operators without one."""
        if self._fb_buffer_name is None:
            return None
        owner = self._fb_owner() if self._fb_owner is not None else self
        return getattr(owner, self._fb_buffer_name, None) if owner is not None else None

    def _fb_sync(self) -> None:
        fb = self._fb_tensor()
        if fb is None or self._fb_default is None:
            return
        key = (id(fb), fb._version)
        if key == self._fb_seen:
            return
        self._fb_seen = key
        loaded = fb.detach().to('cpu', torch.float32)
        self._fb_dirty = loaded.shape != self._fb_default.shape or not torch.equal(loaded, self._fb_default)
        self._private_plans = {}

    def _plan(self, device: torch.device) -> "_lib.Plan":
        if device.type != 'cuda':
            raise RuntimeError(functional.NO_CPU_MSG)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        self._fb_sync()
        if not self._fb_dirty:
            return _lib.cached_plan(idx, **self._plan_kwargs)
        # the registered filterbank differs from the geometry's default: private plan carrying those weights
        pl = self._private_plans.get(idx)
        if pl is None:
            pl = _lib.Plan(_lib.make_config(**self._plan_kwargs), idx)
            pl.set_filterbank(self._fb_tensor().detach().to('cpu', torch.float32).contiguous().numpy())
            self._private_plans[idx] = pl
        return pl


class STFT(_PlanUser):
    """Drop-in for pytorch_sound.models.transforms.STFT (transforms.py:13-101), analysis direction.

    Same constructor; `transform(wav (B, L)) -> (magnitude, phase)`, each (B, filter_length//2+1,
    1 + L//hop).  Registers `square_window`, like the reference; the 4 MB `forward_basis` /
    `inverse_basis` conv kernels of the reference are not materialised (no conv is run) and are
    ignored if present in a loaded state_dict.
    """

    def __init__(self, filter_length: int = 1024, hop_length: int = 512, win_length: int = None,
                 window: str = 'hann'):
        super().__init__()
        self.filter_length = filter_length
        self.hop_length = hop_length
        self.win_length = win_length if win_length else filter_length
        self.window = window
        self.pad_amount = self.filter_length // 2
        assert (filter_length >= self.win_length)
        if window != 'hann':
            raise NotImplementedError(f'{window} is not implemented ! Use hann')
        fft_window = torch.from_numpy(_lib.hann_window(self.win_length, filter_length))
        self.register_buffer('square_window', fft_window ** 2)
        self._plan_kwargs = dict(sample_rate=1, n_fft=filter_length, win_length=self.win_length,
                                 hop_length=hop_length, n_mels=0, pad_mode=_lib.PAD_CENTER)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        for k in ('forward_basis', 'inverse_basis'):
            state_dict.pop(prefix + k, None)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def transform(self, wav: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        _, mag, phase = functional.run(self._plan(wav.device), wav, None, want_mel=False,
                                       spec_kind=_lib.SPEC_MAG_PHASE)
        return mag, phase

    def magnitude(self, wav: torch.Tensor) -> torch.Tensor:
        """transform() without the phase pass (half the HBM writes)."""
        _, mag, _ = functional.run(self._plan(wav.device), wav, None, want_mel=False, spec_kind=_lib.SPEC_MAG)
        return mag

    def forward(self, wav: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.transform(wav)


class LogMelSpectrogram(_PlanUser):
    """Drop-in for pytorch_sound.models.transforms.LogMelSpectrogram (transforms.py:206-244).

    forward(wav (B, L), log_offset=1e-6) -> (B, mel_size, 1 + L//hop) = clamp(ln(mel_filter @ |STFT| + log_offset)).
    As in the reference the STFT uses filter_length = win_length (transforms.py:217), so n_fft must
    equal win_length (otherwise the reference's own matmul shapes mismatch).
    """

    _fb_buffer_name = 'mel_filter'

    def __init__(self, sample_rate: int, mel_size: int, n_fft: int, win_length: int,
                 hop_length: int, min_db: float = None, max_db: float = None,
                 mel_min: float = 0., mel_max: float = None):
        super().__init__()
        if n_fft != win_length:
            raise ValueError("LogMelSpectrogram needs n_fft == win_length: the reference builds "
                             "STFT(filter_length=win_length) and a (mel_size, n_fft//2+1) filter (transforms.py:217-221)")
        self.mel_size = mel_size
        self.stft = STFT(filter_length=win_length, hop_length=hop_length)
        mel_filter = _lib.mel_filterbank(sample_rate, n_fft, mel_size, mel_min, mel_max)
        self.register_buffer('mel_filter', torch.from_numpy(mel_filter))
        self._fb_default = torch.from_numpy(mel_filter).clone()
        self.min_db = _db_to_log(min_db)
        self.max_db = _db_to_log(max_db)
        self._plan_kwargs = dict(sample_rate=sample_rate, n_fft=n_fft, win_length=win_length, hop_length=hop_length,
                                 n_mels=mel_size, fmin=mel_min, fmax=mel_max, pad_mode=_lib.PAD_CENTER)

    def forward(self, wav: torch.Tensor, log_offset: float = 1e-6, lengths: Optional[torch.Tensor] = None,
                norm: bool = False, frame_mask: bool = False, out: Optional[torch.Tensor] = None, reserve_sms: int = 0,
                preemphasis: float = 0.0):
        """`lengths` (int32 (B,), optional, extension): true clip lengths of a zero-padded batch — reflect at
        each clip's own end and zero the frames past it.  `norm=True` (extension) fuses utils.calculate.norm_mel.
        `frame_mask=True` (extension) also returns the (B, T) SpectrogramMasker frame mask (transforms.py:397-416),
        written by the same launch.  `out` (extension): preallocated (B, mel_size, T) tensor to write into.
        `reserve_sms` (extension): SMs left free for a concurrent kernel on another stream (distributed.py).
        `preemphasis` (extension): coefficient of models.sound.PreEmphasis (models/sound.py:66-81) fused into the launch —
        equal to `LogMelSpectrogram(PreEmphasis(coef)(wav[:, None])[:, 0])` without the extra pass over the waveform."""
        epi = _lib.make_epilogue(_lib.LOG_LN_OFFSET, log_offset, self.min_db, self.max_db, norm)
        fm = None
        if frame_mask:
            fm = torch.empty((wav.shape[0], 1 + wav.shape[1] // self.stft.hop_length), device=wav.device, dtype=torch.float32)
        mel, _, _ = functional.run(self._plan(wav.device), wav, epi, lengths=lengths, out=out, frame_mask=fm,
                                   reserve_sms=reserve_sms, preemphasis=preemphasis)
        return (mel, fm) if frame_mask else mel


class LogMelScale(_PlanUser):
    """pytorch_sound.models.transforms.LogMelScale (transforms.py:247-268): mel filterbank + log + clamp applied to a
    MAGNITUDE spectrogram (B, n_fft//2+1, T) -> (B, mel_size, T).

    The reference class cannot be constructed (`torch.Tensor(mel_filter, dtype=torch.float)` raises TypeError,
    :258-259); its intended forward — `torch.matmul(self.mel_filter, magnitude)`,
    `log(mel + log_offset)`, `clamp(min_db, max_db)` — is what runs here, as ONE tcgen05 tensor-core GEMM with the
    accumulator in TMEM (csrc/mel_tc.cuh): the path's dense contraction, for callers that keep magnitudes around."""

    _fb_buffer_name = 'mel_filter'

    def __init__(self, sample_rate: int, mel_size: int, n_fft: int, min_db: float, max_db: float,
                 mel_min: float = 0., mel_max: float = None):
        super().__init__()
        self.mel_size = mel_size
        self.min_db = float(np.log(np.power(10, min_db / 10)))  # unconditional, unlike LogMelSpectrogram (:253-254)
        self.max_db = float(np.log(np.power(10, max_db / 10)))
        mel_filter = _lib.mel_filterbank(sample_rate, n_fft, mel_size, mel_min, mel_max)
        self.register_buffer('mel_filter', torch.from_numpy(mel_filter))
        self._fb_default = torch.from_numpy(mel_filter).clone()
        self._plan_kwargs = dict(sample_rate=sample_rate, n_fft=n_fft, win_length=n_fft, hop_length=n_fft // 4,
                                 n_mels=mel_size, fmin=mel_min, fmax=mel_max, pad_mode=_lib.PAD_CENTER)

    def forward(self, magnitude: torch.Tensor, log_offset: float = 1e-6) -> torch.Tensor:
        epi = _lib.make_epilogue(_lib.LOG_LN_OFFSET, log_offset, self.min_db, self.max_db)
        return functional.logmel_from_magnitude(self._plan(magnitude.device), magnitude, epi)


class STFTTorchAudio(_PlanUser):
    """Drop-in for pytorch_sound.models.transforms.STFTTorchAudio (transforms.py:271-319), analysis direction."""

    def __init__(self, filter_length: int = 1024, hop_length: int = 512, win_length: int = None, n_fft: int = None,
                 window: str = 'hann'):
        super().__init__()
        self.filter_length = filter_length
        self.hop_length = hop_length
        self.win_length = win_length if win_length else self.filter_length
        if window == 'hann':
            self.register_buffer('window', torch.from_numpy(_lib.hann_window(self.win_length)))
        else:
            raise NotImplementedError(f'{window} is not implemented ! Use hann')
        self.n_fft = n_fft if n_fft else self.win_length
        self._plan_kwargs = dict(sample_rate=1, n_fft=self.n_fft, win_length=self.win_length, hop_length=hop_length,
                                 n_mels=0, pad_mode=_lib.PAD_CENTER)

    def forward(self, wav: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        _, re, im = functional.run(self._plan(wav.device), wav, None, want_mel=False, spec_kind=_lib.SPEC_RE_IM)
        return re, im

    def transform(self, wav: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        _, mag, phase = functional.run(self._plan(wav.device), wav, None, want_mel=False,
                                       spec_kind=_lib.SPEC_MAG_PHASE)
        return mag, phase

    def magnitude(self, wav: torch.Tensor) -> torch.Tensor:
        """transform()[0] without the phase pass (what multi_stft_loss consumes, models/sound.py:136-137)."""
        _, mag, _ = functional.run(self._plan(wav.device), wav, None, want_mel=False, spec_kind=_lib.SPEC_MAG)
        return mag


class Audio2Mel(_PlanUser):
    """Drop-in for pytorch_sound.models.transforms.Audio2Mel (MelGAN front-end, transforms.py:322-366).

    forward(audio (B, 1, L)) -> (B, n_mel_channels, L // hop) = log10(max(mel_basis @ |STFT|, 1e-5)).
    """

    _fb_buffer_name = 'mel_basis'

    def __init__(self, n_fft: int = 1024, hop_length: int = 256, win_length: int = 1024, sampling_rate: int = 22050,
                 n_mel_channels: int = 80, mel_fmin: float = 0.0, mel_fmax: float = None):
        super().__init__()
        mel_basis = _lib.mel_filterbank(sampling_rate, n_fft, n_mel_channels, mel_fmin, mel_fmax)
        self.register_buffer("mel_basis", torch.from_numpy(mel_basis))
        self._fb_default = torch.from_numpy(mel_basis).clone()
        self.register_buffer("window", torch.from_numpy(_lib.hann_window(win_length)))
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.win_length = win_length
        self.sampling_rate = sampling_rate
        self.n_mel_channels = n_mel_channels
        self._plan_kwargs = dict(sample_rate=sampling_rate, n_fft=n_fft, win_length=win_length,
                                 hop_length=hop_length, n_mels=n_mel_channels, fmin=mel_fmin, fmax=mel_fmax,
                                 pad_mode=_lib.PAD_HIFI)
        self._epi = _lib.make_epilogue(_lib.LOG_LOG10_FLOOR, 1e-5)

    def forward(self, audio: torch.Tensor) -> torch.Tensor:
        if audio.dim() == 3:
            if audio.shape[1] != 1:
                raise ValueError("Audio2Mel expects (B, 1, L)")
            audio = audio[:, 0]
        mel, _, _ = functional.run(self._plan(audio.device), audio, self._epi)
        return mel


class _Buffers(nn.Module):
    """Name-only container so state_dict keys match the torchaudio sub-modules of the reference."""


class LogMelSpectrogramTorchAudio(_PlanUser):
    """Drop-in for pytorch_sound.models.transforms.LogMelSpectrogramTorchAudio (transforms.py:369-394).

    The reference wraps torchaudio 0.7.0 `MelSpectrogram(..., window_fn=torch.hann_window)` with its defaults:
    POWER spectrogram (power=2), centre reflect padding, HTK mel scale, no area normalisation — different numerics
    from LogMelSpectrogram (magnitude, Slaney) — then `log(mel + log_offset)` and an unconditional clamp.
    Same fused kernel here with `power=2, mel_scale=HTK, mel_norm=none`.  The buffers torchaudio registers keep
    their state_dict names (`melfunc.spectrogram.window`, `melfunc.mel_scale.fb` with torchaudio's (n_freq, n_mels)
    layout)."""

    def __init__(self, sample_rate: int, mel_size: int, n_fft: int, win_length: int,
                 hop_length: int, min_db: float, max_db: float,
                 mel_min: float = 0., mel_max: float = None):
        super().__init__()
        self.mel_size = mel_size
        # db to log (unconditional, unlike LogMelSpectrogram: transforms.py:380-381)
        self.min_db = float(np.log(np.power(10, min_db / 10)))
        self.max_db = float(np.log(np.power(10, max_db / 10)))
        f_max = float(mel_max) if mel_max is not None else float(sample_rate // 2)  # torchaudio's default
        sr_even = 2 * (sample_rate // 2)  # torchaudio spaces the FFT bins over [0, sample_rate // 2]
        fb = _lib.mel_filterbank(sr_even, n_fft, mel_size, mel_min, f_max, _lib.MEL_HTK, _lib.NORM_NONE)
        self.melfunc = _Buffers()
        self.melfunc.spectrogram = _Buffers()
        self.melfunc.mel_scale = _Buffers()
        self.melfunc.spectrogram.register_buffer('window', torch.from_numpy(_lib.hann_window(win_length)))
        self.melfunc.mel_scale.register_buffer('fb', torch.from_numpy(fb.T.copy()))
        self._fb_default = torch.from_numpy(fb).clone()
        self._plan_kwargs = dict(sample_rate=sr_even, n_fft=n_fft, win_length=win_length, hop_length=hop_length,
                                 n_mels=mel_size, fmin=mel_min, fmax=f_max, pad_mode=_lib.PAD_CENTER,
                                 mel_scale=_lib.MEL_HTK, mel_norm=_lib.NORM_NONE, power=2)

    _fb_buffer_name = 'melfunc.mel_scale.fb'

    def _fb_tensor(self) -> Optional[torch.Tensor]:
        owner = self._fb_owner() if self._fb_owner is not None else self
        try:
            return owner.melfunc.mel_scale.fb.t()  # torchaudio keeps (n_freq, n_mels)
        except AttributeError:
            return None

    def _fb_sync(self) -> None:
        # `.t()` makes a new view object per call: key the watch on the underlying buffer instead
        owner = self._fb_owner() if self._fb_owner is not None else self
        try:
            base = owner.melfunc.mel_scale.fb
        except AttributeError:
            return
        key = (id(base), base._version)
        if key == self._fb_seen:
            return
        self._fb_seen = key
        loaded = base.detach().to('cpu', torch.float32).t()
        self._fb_dirty = loaded.shape != self._fb_default.shape or not torch.equal(loaded, self._fb_default)
        self._private_plans = {}

    def forward(self, wav: torch.Tensor, log_offset: float = 1e-6) -> torch.Tensor:
        epi = _lib.make_epilogue(_lib.LOG_LN_OFFSET, log_offset, self.min_db, self.max_db)
        mel, _, _ = functional.run(self._plan(wav.device), wav, epi)
        return mel


class MelToMFCC(nn.Module):
    """pytorch_sound.models.transforms.MelToMFCC (transforms.py:419-430): ortho DCT-II of a mel spectrogram.
    One streaming kernel over the mel tensor (b200mel_mel_to_mfcc); torchaudio's functional.create_dct is
    restated so torchaudio is not needed."""

    def __init__(self, n_mfcc: int, mel_size: int, norm: str = 'ortho'):
        super().__init__()
        self.n_mfcc = n_mfcc
        n = torch.arange(float(mel_size))
        k = torch.arange(float(n_mfcc)).unsqueeze(1)
        dct = torch.cos(np.pi / float(mel_size) * (n + 0.5) * k)  # (n_mfcc, mel_size)
        if norm is None:
            dct *= 2.0
        else:
            assert norm == 'ortho'
            dct[0] *= 1.0 / np.sqrt(2.0)
            dct *= np.sqrt(2.0 / float(mel_size))
        self.register_buffer('dct_mat', dct)

    def forward(self, mel_spec: torch.Tensor) -> torch.Tensor:
        assert len(mel_spec.size()) == 3
        return functional.mel_to_mfcc(mel_spec, self.dct_mat)


class MFCC(nn.Module):
    """pytorch_sound.models.transforms.MFCC (transforms.py:433-455): LogMelSpectrogram followed by the ortho DCT.

    The reference asserts a 3-D input but then feeds its 2-D-only STFT; this class takes the
    (B, L) waveform its LogMelSpectrogram needs (a (B, 1, L) tensor is squeezed)."""

    def __init__(self, sample_rate: int, mel_size: int, n_fft: int, win_length: int, n_mfcc: int,
                 hop_length: int, min_db: float, max_db: float,
                 mel_min: float = 0., mel_max: float = None, norm: str = 'ortho'):
        super().__init__()
        self.n_mfcc = n_mfcc
        self.mel_func = LogMelSpectrogram(
            sample_rate, mel_size, n_fft, win_length, hop_length, min_db, max_db,
            mel_min, mel_max
        )
        self.register_buffer('dct_mat', MelToMFCC(n_mfcc, mel_size, norm).dct_mat.clone())

    def forward(self, wav: torch.Tensor) -> torch.Tensor:
        if wav.dim() == 3 and wav.shape[1] == 1:
            wav = wav[:, 0]
        # one launch: the DCT as an epilogue of the mel kernel, while the log-mel column is still on chip
        # (io.out_mfcc, include/b200mel.h); geometries the fused kernel does not serve take the two launches
        mf = self.mel_func
        epi = _lib.make_epilogue(_lib.LOG_LN_OFFSET, 1e-6, mf.min_db, mf.max_db, False)
        fused = functional.mfcc_fused(mf._plan(wav.device), wav, epi, self.dct_mat)
        if fused is not None:
            return fused[0]
        mel_spectrogram = self.mel_func(wav)
        return functional.mel_to_mfcc(mel_spectrogram, self.dct_mat)


class SpectrogramMasker(nn.Module):
    """pytorch_sound.models.transforms.SpectrogramMasker (transforms.py:397-416): wave-level validity mask ->
    frame-level mask with the kernel's frame geometry (T = 1 + L // hop).

    The reference runs a strided mean-filter conv over the mask padded with win//2 ones on the left and win//2
    zeros on the right and takes ceil(): a frame is 1 iff any sample of its window is valid.  Same result here from
    one cumulative sum (no conv weights, runs on the mask's device; the reference hard-codes .cuda()).
    `from_lengths` gives the same frame mask straight from per-clip lengths (the `lengths=` argument of the
    extractor), without materialising a wave-level mask."""

    def __init__(self, win_length: int, hop_length: int):
        super().__init__()
        self.win_length = win_length
        self.hop_length = hop_length

    def forward(self, wav_mask: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            half = self.win_length // 2
            m = wav_mask.float()
            m = torch.nn.functional.pad(m, [0, half], value=0.)
            m = torch.nn.functional.pad(m, [half, 0], value=1.)
            csum = torch.nn.functional.pad(torch.cumsum(m.double(), dim=-1), [1, 0])
            n_frames = (m.shape[-1] - self.win_length) // self.hop_length + 1
            start = torch.arange(n_frames, device=m.device) * self.hop_length
            frame_sum = csum[..., start + self.win_length] - csum[..., start]
            return (frame_sum > 0).float()

    def from_lengths(self, lengths: torch.Tensor, n_samples: int) -> torch.Tensor:
        half = self.win_length // 2
        n_frames = (n_samples + 2 * half - self.win_length) // self.hop_length + 1
        t = torch.arange(n_frames, device=lengths.device)
        # frame t covers padded samples [t*hop, t*hop + win) = wave samples [t*hop - half, t*hop - half + win)
        return ((t * self.hop_length - half) < lengths.view(-1, 1)).float()
