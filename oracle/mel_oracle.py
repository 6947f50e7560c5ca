"""CPU oracle for the STFT -> magnitude -> mel -> log path of AppleHolic/pytorch_sound.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  The product path
(pytorch_sound_b200) never imports anything from oracle/ and has no CPU fallback.

It restates, in numpy (float64 = ground truth) and in torch-CPU float32 ("the
reference as it would run"), the arithmetic of

  pytorch_sound/models/transforms.py:19-69    STFT.__init__/transform (conv-DFT)
  pytorch_sound/models/transforms.py:211-244  LogMelSpectrogram
  pytorch_sound/models/transforms.py:276-311  STFTTorchAudio
  pytorch_sound/models/transforms.py:327-366  Audio2Mel
  pytorch_sound/interface/hifi_gan.py:34-63   MelSpectrogram (HiFi-GAN front-end)
  pytorch_sound/utils/calculate.py:10-43      db2log / norm_mel / unnorm_mel
  pytorch_sound/settings.py:9-22              constants

and of the third-party functions the reference calls but does not vendor
(requirements.txt: librosa==0.8.0, torch==1.7.0, scipy unpinned):

  librosa.filters.mel (0.8.0: htk=False, norm='slaney', dtype=float32)
  librosa.util.pad_center, scipy.signal.get_window('hann', fftbins=True), torch.hann_window
  torch.stft (center/reflect, onesided, unnormalised), F.pad(mode='reflect')

Parity pin status
-----------------
The reference ships NO golden vectors / known-answer tests for this path
(tests/ are assertion-free loader smoke scripts), so the oracle is pinned against
outputs of the REFERENCE'S OWN CODE executed in the build container:
tests/golden/make_golden.py imports /root/reference/pytorch_sound/models/transforms.py
and interface/hifi_gan.py unmodified and runs STFT.transform, LogMelSpectrogram,
STFTTorchAudio, Audio2Mel and MelSpectrogram on seeded inputs; the committed
tests/golden/*.npz are what tests/test_oracle_golden.py checks this file against.
Two third-party pieces had to be shimmed to import the reference on torch 2.11 /
no-librosa: torch.stft's removed legacy (non-complex) return, and
librosa.filters.mel, which is absent from the image.  The filterbank is therefore
"parity pinned by restatement + cross-check" only: `mel_filterbank` below follows the
librosa 0.8.0 source and is cross-checked against torchaudio's independent
melscale_fbanks(norm='slaney', mel_scale='slaney') (<= 2e-7 abs) in the tests.
"""
from __future__ import annotations

import math

import numpy as np

# ---------------------------------------------------------------------------------------------
# settings.py:9-22
# ---------------------------------------------------------------------------------------------
SAMPLE_RATE = 22050
N_FFT = 1024
WIN_LENGTH = 1024
HOP_LENGTH = 256
SPEC_SIZE = WIN_LENGTH // 2 + 1
MEL_SIZE = 80
MFCC_SIZE = 40
MEL_MIN = 0
MEL_MAX = 8000
MIN_DB = -50
MAX_DB = 30


# ---------------------------------------------------------------------------------------------
# utils/calculate.py:10-43
# ---------------------------------------------------------------------------------------------
def db2log(db):
    """utils/calculate.py:10-19 — ln(10 ** (db / 10))."""
    return np.log(np.power(10.0, np.asarray(db, dtype=np.float64) / 10.0))


def unnorm_mel(x):
    """utils/calculate.py:22-29."""
    lo, hi = db2log(MIN_DB), db2log(MAX_DB)
    return ((x + 1) / 2) * (hi - lo) + lo


def norm_mel(x):
    """utils/calculate.py:32-43."""
    lo, hi = db2log(MIN_DB), db2log(MAX_DB)
    x = np.clip(x, lo, hi)
    return (x - lo) / (hi - lo) * 2 - 1


# ---------------------------------------------------------------------------------------------
# third-party restatements
# ---------------------------------------------------------------------------------------------
def hann_periodic(win_length: int, n_fft: int | None = None) -> np.ndarray:
    """scipy.signal.get_window('hann', M, fftbins=True) (= torch.hann_window(M), periodic) in float64,
    centre-padded to n_fft as librosa.util.pad_center does (models/transforms.py:30-31)."""
    n = np.arange(win_length, dtype=np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)
    if n_fft is None or n_fft == win_length:
        return w
    lpad = (n_fft - win_length) // 2
    out = np.zeros(n_fft, dtype=np.float64)
    out[lpad:lpad + win_length] = w
    return out


def _hz_to_mel(f, htk=False):
    f = np.asanyarray(f, dtype=np.float64)
    if htk:
        return 2595.0 * np.log10(1.0 + f / 700.0)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore"):
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def _mel_to_hz(m, htk=False):
    m = np.asanyarray(m, dtype=np.float64)
    if htk:
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney") -> np.ndarray:
    """librosa.filters.mel, librosa 0.8.0 (called positionally as (sr, n_fft, n_mels, fmin, fmax) at
    models/transforms.py:220,339-341 and interface/hifi_gan.py:42).  float64 math, float32 storage,
    including librosa's order of roundings (triangles stored as float32, then scaled in place)."""
    if fmax is None:
        fmax = float(sr) / 2
    n_freq = 1 + n_fft // 2
    weights = np.zeros((n_mels, n_freq), dtype=np.float32)
    fftfreqs = np.linspace(0, float(sr) / 2, n_freq, endpoint=True)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin, htk), _hz_to_mel(fmax, htk), n_mels + 2), htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if norm == "slaney":
        enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
        weights *= enorm[:, np.newaxis]
    return weights


def reflect_pad(x: np.ndarray, p: int) -> np.ndarray:
    """F.pad(x, (p, p), mode='reflect') on the last axis (no edge repeat)."""
    if p == 0:
        return x
    if x.shape[-1] <= p:
        raise ValueError("reflect padding needs L > pad")
    return np.pad(x, [(0, 0)] * (x.ndim - 1) + [(p, p)], mode="reflect")


def num_frames(L: int, n_fft: int, hop: int, pad: int) -> int:
    span = L + 2 * pad - n_fft
    return 0 if span < 0 else span // hop + 1


# ---------------------------------------------------------------------------------------------
# float64 ground truth
# ---------------------------------------------------------------------------------------------
def stft_complex(wav, n_fft, hop, win_length=None, pad=None, dtype=np.float64) -> np.ndarray:
    """Windowed one-sided DFT of every frame: (B, L) -> complex (B, F, T).

    Same numbers as the strided conv with the windowed DFT basis (models/transforms.py:35-45,63-66)
    and as torch.stft(normalized=False, onesided=True) (models/transforms.py:298-301)."""
    wav = np.asarray(wav, dtype=dtype)
    if wav.ndim == 1:
        wav = wav[None]
    win_length = win_length or n_fft
    pad = n_fft // 2 if pad is None else pad
    x = reflect_pad(wav, pad)
    T = num_frames(wav.shape[-1], n_fft, hop, pad)
    w = hann_periodic(win_length, n_fft).astype(np.float32).astype(dtype)  # the reference stores fp32 windows
    idx = np.arange(n_fft)[None, :] + hop * np.arange(T)[:, None]
    frames = x[:, idx] * w  # (B, T, n_fft)
    spec = np.fft.rfft(frames.astype(np.float64), axis=-1)
    return np.transpose(spec, (0, 2, 1))


def stft_transform(wav, filter_length=1024, hop_length=512, win_length=None):
    """STFT.transform (models/transforms.py:53-69) / STFTTorchAudio.transform (:305-311): (mag, phase)."""
    s = stft_complex(wav, filter_length, hop_length, win_length, pad=filter_length // 2)
    return np.abs(s), np.arctan2(s.imag, s.real)


def apply_log(mel, kind, arg):
    if kind == "ln_offset":
        return np.log(mel + arg)
    if kind == "ln_floor":
        return np.log(np.maximum(mel, arg))
    if kind == "log10_floor":
        return np.log10(np.maximum(mel, arg))
    if kind == "none":
        return mel
    raise ValueError(kind)


def log_mel_spectrogram(wav, sample_rate=SAMPLE_RATE, mel_size=MEL_SIZE, n_fft=N_FFT, win_length=WIN_LENGTH,
                        hop_length=HOP_LENGTH, min_db=None, max_db=None, mel_min=0.0, mel_max=None,
                        log_offset=1e-6, clamp=True):
    """LogMelSpectrogram.forward (models/transforms.py:231-244), float64.  `clamp=False` returns the
    pre-clamp log-mel (the quantity the parity metric is defined on)."""
    mag, _ = stft_transform(wav, win_length, hop_length)  # STFT(filter_length=win_length) — :217
    fb = mel_filterbank(sample_rate, n_fft, mel_size, mel_min, mel_max).astype(np.float64)
    mel = np.einsum("mf,bft->bmt", fb, mag)
    y = np.log(mel + log_offset)
    if clamp:
        if min_db:  # truthiness as in the reference (:222-229,240-243)
            y = np.maximum(y, db2log(min_db))
        if max_db:
            y = np.minimum(y, db2log(max_db))
    return y


def audio2mel(audio, n_fft=1024, hop_length=256, win_length=1024, sampling_rate=22050, n_mel_channels=80,
              mel_fmin=0.0, mel_fmax=None):
    """Audio2Mel.forward (models/transforms.py:351-366), float64. audio: (B, 1, L) or (B, L)."""
    audio = np.asarray(audio, dtype=np.float64)
    if audio.ndim == 3:
        audio = audio[:, 0]
    p = (n_fft - hop_length) // 2
    s = stft_complex(audio, n_fft, hop_length, win_length, pad=p)
    fb = mel_filterbank(sampling_rate, n_fft, n_mel_channels, mel_fmin, mel_fmax).astype(np.float64)
    mel = np.einsum("mf,bft->bmt", fb, np.abs(s))
    return np.log10(np.maximum(mel, 1e-5))


def hifi_mel_spectrogram(wav, sampling_rate=22050, n_fft=1024, window_size=1024, hop_size=256, num_mels=80,
                         fmin=0.0, fmax=8000.0, floor=True):
    """interface.hifi_gan.MelSpectrogram.forward(is_center=False) (interface/hifi_gan.py:46-63), float64."""
    p = (n_fft - hop_size) // 2
    s = stft_complex(wav, n_fft, hop_size, window_size, pad=p)
    mag = np.sqrt(s.real ** 2 + s.imag ** 2 + 1e-9)
    fb = mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax).astype(np.float64)
    mel = np.einsum("mf,bft->bmt", fb, mag)
    return np.log(np.maximum(mel, 1e-5)) if floor else mel


# ---------------------------------------------------------------------------------------------
# float32 "reference as it would run" (torch CPU) — also the timed CPU baseline

# ---------------------------------------------------------------------------------------------
# torchaudio-wrapping variant and the operators either side of the spectral path
# ---------------------------------------------------------------------------------------------
def log_mel_spectrogram_torchaudio(wav, sample_rate, mel_size, n_fft, win_length, hop_length, min_db, max_db,
                                   mel_min=0.0, mel_max=None, log_offset=1e-6, dtype=np.float64):
    """LogMelSpectrogramTorchAudio (models/transforms.py:369-394).  The arithmetic lives in torchaudio 0.7.0
    (requirements.txt:3, not vendored): transforms.MelSpectrogram defaults = Spectrogram(power=2.0, centre reflect
    pad n_fft//2, window zero-padded to n_fft) and MelScale = create_fb_matrix with the HTK scale, no area norm,
    all_freqs = linspace(0, sample_rate // 2, n_fft // 2 + 1), f_max default float(sample_rate // 2).  Then
    log(mel + log_offset) and an unconditional clamp to [ln 10^(min_db/10), ln 10^(max_db/10)]."""
    spec = stft_complex(wav, n_fft, hop_length, win_length, pad=n_fft // 2, dtype=dtype)
    power = spec.real ** 2 + spec.imag ** 2
    f_max = float(mel_max) if mel_max is not None else float(sample_rate // 2)
    fb = mel_filterbank(2 * (sample_rate // 2), n_fft, mel_size, mel_min, f_max, htk=True, norm=None).astype(dtype)
    mel = np.einsum("mf,bft->bmt", fb, power)
    return np.clip(np.log(mel + log_offset), db2log(min_db), db2log(max_db))


def pre_emphasis(x, coef=0.97):
    """PreEmphasis.forward (models/sound.py:66-81) on (B, L): one reflected sample on the left, then the 2-tap
    cross-correlation with [-coef, 1]: y[n] = x[n] - coef x[n-1], y[0] = x[0] - coef x[1]."""
    x = np.asarray(x, dtype=np.float64)
    prev = np.concatenate([x[..., 1:2], x[..., :-1]], axis=-1)
    return x - coef * prev


def volume_norm_log(x, target_db=-11.5):
    """volume_norm_log / volume_norm_log_torch (utils/calculate.py:46-63): np.std is the population std (ddof 0),
    torch.std the unbiased one (ddof 1) — `ddof` picks which twin is restated."""
    x = np.asarray(x, dtype=np.float64)
    return x / (np.std(x) / 10 ** (target_db / 10))


def volume_norm_log_torch(x, target_db=-11.5):
    x = np.asarray(x, dtype=np.float64)
    return x / (np.std(x, ddof=1) / 10 ** (target_db / 10))


def create_dct(n_mfcc, n_mels, norm="ortho"):
    """torchaudio.functional.create_dct (0.7.0), the (n_mels, n_mfcc) DCT-II matrix MelToMFCC transposes
    (models/transforms.py:427)."""
    n = np.arange(float(n_mels))
    k = np.arange(float(n_mfcc))[:, None]
    dct = np.cos(np.pi / float(n_mels) * (n + 0.5) * k)  # (n_mfcc, n_mels)
    if norm is None:
        dct *= 2.0
    else:
        dct[0] *= 1.0 / np.sqrt(2.0)
        dct *= np.sqrt(2.0 / float(n_mels))
    return dct.T


def mel_to_mfcc(mel, n_mfcc, norm="ortho"):
    """MelToMFCC.forward (models/transforms.py:428-430): dct_mat (n_mfcc, n_mels) @ mel (B, n_mels, T)."""
    mel = np.asarray(mel, dtype=np.float64)
    return np.einsum("cm,bmt->bct", create_dct(n_mfcc, mel.shape[1], norm).T, mel)

def multi_stft_loss(pred, target, stft_params, eps=1e-5):
    """models/sound.py:120-147, float64.  stft_params: (n_fft, window size, hop size) triplets; every resolution runs
    STFTTorchAudio(filter_length=win, hop_length=hop, win_length=win, n_fft=fft).transform (models/sound.py:113-117),
    i.e. a Hann window of `win` samples centre-padded to n_fft (torch.stft), centre reflect padding of n_fft // 2.
    Returns (loss, spectral-convergence loss, log-magnitude loss), each averaged over the resolutions."""
    loss = sc = mg = 0.0
    for fft, win, hop in stft_params:
        p = np.abs(stft_complex(pred, fft, hop, win))
        t = np.abs(stft_complex(target, fft, hop, win))
        n = t.shape[1] * t.shape[2]
        sc_ = np.mean(np.sqrt(np.sum((t - p) ** 2, axis=(1, 2))) / np.sqrt(np.sum(t ** 2, axis=(1, 2))))
        mg_ = np.mean(np.sum(np.abs(np.log(t + eps) - np.log(p + eps)), axis=(1, 2))) / n
        loss += sc_ + mg_
        sc += sc_
        mg += mg_
    k = len(stft_params)
    return loss / k, sc / k, mg / k


def spectrogram_mask(wav_mask, win_length, hop_length):
    """SpectrogramMasker.forward (models/transforms.py:408-416): pad win//2 zeros on the right and win//2 ones on the
    left, mean-filter conv of width win / stride hop, ceil."""
    m = np.asarray(wav_mask, dtype=np.float64)
    half = win_length // 2
    m = np.concatenate([np.ones(m.shape[:-1] + (half,)), m, np.zeros(m.shape[:-1] + (half,))], axis=-1)
    T = (m.shape[-1] - win_length) // hop_length + 1
    idx = np.arange(win_length)[None, :] + hop_length * np.arange(T)[:, None]
    return np.ceil(m[..., idx].mean(axis=-1))


# ---------------------------------------------------------------------------------------------
class TorchReference:
    """The reference modules' op sequences on torch-CPU float32, with today's torch API
    (torch.stft(return_complex=True) instead of the removed legacy return)."""

    def __init__(self, sample_rate=SAMPLE_RATE, mel_size=MEL_SIZE, n_fft=N_FFT, win_length=WIN_LENGTH,
                 hop_length=HOP_LENGTH, min_db=None, max_db=None, mel_min=0.0, mel_max=None):
        import torch

        self.torch = torch
        self.n_fft, self.win, self.hop = n_fft, win_length, hop_length
        self.min_db = float(db2log(min_db)) if min_db else None
        self.max_db = float(db2log(max_db)) if max_db else None
        self.mel_filter = torch.from_numpy(mel_filterbank(sample_rate, n_fft, mel_size, mel_min, mel_max))
        window = hann_periodic(win_length, win_length)
        self.window = torch.from_numpy(window).float()
        # conv-DFT basis, models/transforms.py:35-45 (filter_length = win_length, :217)
        N = win_length
        cut = N // 2 + 1
        basis = np.fft.fft(np.eye(N))
        basis = np.vstack([np.real(basis[:cut]), np.imag(basis[:cut])])
        self.forward_basis = torch.FloatTensor(basis[:, None, :]) * self.window

    def to(self, device):
        """Move the registered buffers (the reference modules are plain nn.Modules and run on either device)."""
        self.mel_filter = self.mel_filter.to(device)
        self.window = self.window.to(device)
        self.forward_basis = self.forward_basis.to(device)
        return self

    def logmel_conv(self, wav, log_offset=1e-6):
        """STFT.transform + LogMelSpectrogram.forward (models/transforms.py:53-69,231-244)."""
        torch, F = self.torch, self.torch.nn.functional
        x = wav.unsqueeze(1).unsqueeze(1)
        x = F.pad(x, (self.win // 2, self.win // 2, 0, 0), mode="reflect").squeeze(1)
        ft = F.conv1d(x, self.forward_basis, stride=self.hop, padding=0)
        re, im = ft.chunk(2, 1)
        mag = torch.sqrt(re ** 2 + im ** 2)
        # STFT.transform also computes the phase (models/transforms.py:69) although LogMelSpectrogram discards
        # it (:232): part of the stock op sequence, so it is part of what gets timed
        self.last_phase = torch.atan2(im.data, re.data)
        mel = torch.matmul(self.mel_filter, mag)
        mel = torch.log(mel + log_offset)
        if self.min_db:
            mel = mel.clamp_min(self.min_db)
        if self.max_db:
            mel = mel.clamp_max(self.max_db)
        return mel

    def logmel_stft(self, wav, log_offset=1e-6):
        """STFTTorchAudio.transform (models/transforms.py:297-311) + the same mel/log tail."""
        torch = self.torch
        s = torch.stft(wav, self.n_fft, self.hop, self.win, self.window, center=True, pad_mode="reflect",
                       normalized=False, onesided=True, return_complex=True)
        mag = torch.sqrt(s.real ** 2 + s.imag ** 2)
        mel = torch.matmul(self.mel_filter, mag)
        mel = torch.log(mel + log_offset)
        if self.min_db:
            mel = mel.clamp_min(self.min_db)
        if self.max_db:
            mel = mel.clamp_max(self.max_db)
        return mel

    def hifi(self, wav):
        """interface/hifi_gan.py:46-63."""
        torch = self.torch
        p = (self.n_fft - self.hop) // 2
        x = torch.nn.functional.pad(wav.unsqueeze(1), [p, p], mode="reflect").squeeze(1)
        s = torch.stft(x, self.n_fft, hop_length=self.hop, win_length=self.win, window=self.window, center=False,
                       pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
        spec = torch.sqrt(s.real ** 2 + s.imag ** 2 + 1e-9)
        spec = torch.matmul(self.mel_filter, spec)
        return torch.log(torch.clamp(spec, min=1e-5))


# ---------------------------------------------------------------------------------------------
# synthetic inputs — shared by tests and bench so both see the same waveforms
# ---------------------------------------------------------------------------------------------
def synth_clips(B: int, L: int, sr: int, seed: int, first_clip: int = 0, noise: float = 0.01) -> np.ndarray:
    """Clip i: 0.5*sin(2*pi*f_i*n/sr) + noise*N(0,1), f_i = 110*2^((i mod 72)/12) Hz capped at 0.45*sr."""
    rng = np.random.default_rng(seed)
    n = np.arange(L, dtype=np.float64)
    out = np.empty((B, L), dtype=np.float32)
    for i in range(B):
        f = min(110.0 * 2.0 ** (((first_clip + i) % 72) / 12.0), 0.45 * sr)
        x = 0.5 * np.sin(2.0 * math.pi * f * n / sr)
        if noise:
            x = x + noise * rng.standard_normal(L)
        out[i] = x.astype(np.float32)
    return out


def parity_error(y, y_ref64) -> float:
    """max |y - ref| / max(1, |ref|)  — the mixed abs/rel metric of SURVEY 8d (tolerance 1e-4)."""
    y = np.asarray(y, dtype=np.float64)
    r = np.asarray(y_ref64, dtype=np.float64)
    return float(np.max(np.abs(y - r) / np.maximum(1.0, np.abs(r)))) if y.size else 0.0
