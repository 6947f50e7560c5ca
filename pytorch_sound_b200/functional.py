"""Host-side launcher: torch tensors in, one b200mel_forward call, torch tensors out."""
import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib


NO_CPU_MSG = ("pytorch_sound_b200 runs on CUDA tensors only (no CPU fallback): move the batch to the GPU first, "
              "as Trainer does with to_device (utils/tensor.py:6-15)")


def _check_wav(wav: torch.Tensor) -> torch.Tensor:
    if not isinstance(wav, torch.Tensor):
        raise TypeError("wav must be a torch.Tensor")
    if wav.dim() != 2:
        raise ValueError(f"wav must be (B, L), got {tuple(wav.shape)}")
    if not wav.is_cuda:
        raise RuntimeError(NO_CPU_MSG)
    if wav.dtype != torch.float32:
        raise TypeError(f"wav must be float32, got {wav.dtype}")
    if wav.requires_grad and torch.is_grad_enabled():
        raise RuntimeError("pytorch_sound_b200 is forward-only feature extraction; detach() the waveform or run "
                           "under torch.no_grad()")
    if wav.shape[1] > 0 and wav.stride(1) != 1:
        wav = wav.contiguous()
    return wav


def run(plan: "_lib.Plan", wav: torch.Tensor, epi: Optional["_lib.Epilogue"], want_mel: bool = True,
        spec_kind: int = _lib.SPEC_NONE, lengths: Optional[torch.Tensor] = None,
        out: Optional[torch.Tensor] = None, frame_mask: Optional[torch.Tensor] = None, reserve_sms: int = 0,
        preemphasis: float = 0.0) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Launch the fused kernel on wav's device / current stream. Returns (mel, out_a, out_b).

    `out`: optional preallocated contiguous (B, n_mels, T) float32 CUDA tensor for the mel frames (e.g. a slice of a
    symmetric-memory gather buffer).  `frame_mask`: optional (B, T) float32 tensor the same launch fills with the
    SpectrogramMasker frame mask.  `reserve_sms`: SMs the persistent mel kernel leaves free for concurrent kernels.
    `preemphasis`: coefficient of the fused pre-emphasis prologue (0 = off)."""
    wav = _check_wav(wav)
    dev = wav.device
    if dev.index != plan.device_index:
        raise RuntimeError(f"plan lives on cuda:{plan.device_index}, wav on {dev}")
    B, L = wav.shape
    T = plan.out_frames(L)
    n_freq = plan.cfg.n_fft // 2 + 1
    mel = out_a = out_b = None
    if want_mel:
        if out is not None:
            if out.shape != (B, plan.cfg.n_mels, T) or out.dtype != torch.float32 or out.device != dev or not out.is_contiguous():
                raise ValueError(f"out must be a contiguous float32 {(B, plan.cfg.n_mels, T)} tensor on {dev}")
            mel = out
        else:
            mel = torch.empty((B, plan.cfg.n_mels, T), device=dev, dtype=torch.float32)
    if frame_mask is not None and (frame_mask.shape != (B, T) or frame_mask.dtype != torch.float32 or
                                   frame_mask.device != dev or not frame_mask.is_contiguous()):
        raise ValueError(f"frame_mask must be a contiguous float32 {(B, T)} tensor on {dev}")
    if spec_kind != _lib.SPEC_NONE:
        out_a = torch.empty((B, n_freq, T), device=dev, dtype=torch.float32)
        if spec_kind in (_lib.SPEC_MAG_PHASE, _lib.SPEC_RE_IM):
            out_b = torch.empty((B, n_freq, T), device=dev, dtype=torch.float32)
    len_ptr = None
    if lengths is not None:
        if lengths.device != dev or lengths.dtype != torch.int32 or lengths.numel() != B:
            lengths = lengths.to(device=dev, dtype=torch.int32).reshape(B)
        lengths = lengths.contiguous()
        len_ptr = lengths.data_ptr()
    if B == 0:
        return mel, out_a, out_b
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        io = _lib.IO(C.sizeof(_lib.IO), spec_kind, wav.data_ptr(), B, L, wav.stride(0) if B > 1 else max(L, 1), len_ptr,
                     mel.data_ptr() if mel is not None else None, out_a.data_ptr() if out_a is not None else None,
                     out_b.data_ptr() if out_b is not None else None,
                     frame_mask.data_ptr() if frame_mask is not None else None, int(reserve_sms), float(preemphasis),
                     None, None, 0, 0)
        rc = _lib.lib().b200mel_forward_io(plan.handle, C.byref(io), C.byref(epi) if epi is not None else None,
                                           C.c_void_p(stream))
    _lib.check(rc)
    return mel, out_a, out_b


def mfcc_fused(plan: "_lib.Plan", wav: torch.Tensor, epi: "_lib.Epilogue", dct: torch.Tensor,
               want_mel: bool = False) -> Optional[Tuple[torch.Tensor, Optional[torch.Tensor]]]:
    """MFCC.forward (models/transforms.py:433-455) in ONE launch: the DCT applied as an epilogue of the mel kernel
    (io.out_mfcc).  Returns (mfcc (B, n_mfcc, T), mel or None), or None when the library answers B200MEL_EUNSUP for
    this plan / call — the caller then runs the mel launch and b200mel_mel_to_mfcc."""
    wav = _check_wav(wav)
    dev = wav.device
    if dev.index != plan.device_index:
        raise RuntimeError(f"plan lives on cuda:{plan.device_index}, wav on {dev}")
    if dct.dim() != 2 or dct.shape[1] != plan.cfg.n_mels:
        raise ValueError(f"dct must be (n_mfcc, {plan.cfg.n_mels}), got {tuple(dct.shape)}")
    dct = dct.to(device=dev, dtype=torch.float32).contiguous()
    B, L = wav.shape
    T = plan.out_frames(L)
    out = torch.empty((B, dct.shape[0], T), device=dev, dtype=torch.float32)
    mel = torch.empty((B, plan.cfg.n_mels, T), device=dev, dtype=torch.float32) if want_mel else None
    if B == 0:
        return out, mel
    with torch.cuda.device(dev):
        io = _lib.IO(C.sizeof(_lib.IO), _lib.SPEC_NONE, wav.data_ptr(), B, L, wav.stride(0) if B > 1 else max(L, 1), None,
                     mel.data_ptr() if mel is not None else None, None, None, None, 0, 0.0,
                     dct.data_ptr(), out.data_ptr(), int(dct.shape[0]), 0)
        rc = _lib.lib().b200mel_forward_io(plan.handle, C.byref(io), C.byref(epi),
                                           C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc == _lib.EUNSUP:
        return None
    _lib.check(rc)
    return out, mel


def _check_cuda_f32(x: torch.Tensor, what: str) -> None:
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{what} must be a torch.Tensor")
    if not x.is_cuda:
        raise RuntimeError(NO_CPU_MSG)
    if x.dtype != torch.float32:
        raise TypeError(f"{what} must be float32, got {x.dtype}")
    if x.requires_grad and torch.is_grad_enabled():
        raise RuntimeError("pytorch_sound_b200 is forward-only; detach() the input or run under torch.no_grad()")


def preemphasis(x: torch.Tensor, coef: float) -> torch.Tensor:
    """(B, L) CUDA float32 -> y[n] = x[n] - coef x[n-1] with the reference's 1-sample reflect pad (models/sound.py:66-81)."""
    _check_cuda_f32(x, "input")
    if x.dim() != 2:
        raise ValueError(f"expected (B, L), got {tuple(x.shape)}")
    if x.shape[1] > 0 and x.stride(1) != 1:
        x = x.contiguous()
    B, L = x.shape
    y = torch.empty((B, L), device=x.device, dtype=torch.float32)
    if B == 0 or L == 0:
        return y
    with torch.cuda.device(x.device):
        rc = _lib.lib().b200mel_preemphasis(x.data_ptr(), B, L, x.stride(0) if B > 1 else L, float(coef), y.data_ptr(), L,
                                            C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    _lib.check(rc)
    return y


def volume_norm(x: torch.Tensor, target_db: float) -> torch.Tensor:
    """x / (std(x) / 10^(target_db / 10)), std over the whole tensor (utils/calculate.py:56-63)."""
    _check_cuda_f32(x, "x")
    xc = x.contiguous()
    y = torch.empty_like(xc)
    if xc.numel() == 0:
        return y
    scratch = torch.empty(2, device=x.device, dtype=torch.float64)
    with torch.cuda.device(x.device):
        rc = _lib.lib().b200mel_volume_norm(xc.data_ptr(), xc.numel(), float(target_db), y.data_ptr(), scratch.data_ptr(),
                                            C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    _lib.check(rc)
    return y


def mel_to_mfcc(mel: torch.Tensor, dct: torch.Tensor) -> torch.Tensor:
    """dct (n_mfcc, n_mels) @ mel (B, n_mels, T) -> (B, n_mfcc, T) (models/transforms.py:428-430)."""
    _check_cuda_f32(mel, "mel_spec")
    if mel.dim() != 3 or dct.dim() != 2 or dct.shape[1] != mel.shape[1]:
        raise ValueError(f"expected mel (B, M, T) and dct (C, M), got {tuple(mel.shape)} and {tuple(dct.shape)}")
    mel = mel.contiguous()
    dct = dct.to(device=mel.device, dtype=torch.float32).contiguous()
    B, M, T = mel.shape
    out = torch.empty((B, dct.shape[0], T), device=mel.device, dtype=torch.float32)
    if B == 0 or T == 0:
        return out
    with torch.cuda.device(mel.device):
        rc = _lib.lib().b200mel_mel_to_mfcc(mel.data_ptr(), dct.data_ptr(), B, M, dct.shape[0], T, out.data_ptr(),
                                            C.c_void_p(torch.cuda.current_stream(mel.device).cuda_stream))
    _lib.check(rc)
    return out


def stft_loss_terms(pred_mag: torch.Tensor, target_mag: torch.Tensor, eps: float, out2: torch.Tensor) -> None:
    """Adds one resolution's (spectral-convergence, log-magnitude) terms of multi_stft_loss (models/sound.py:139-141)
    into the 2-float CUDA tensor `out2`."""
    _check_cuda_f32(pred_mag, "pred magnitudes")
    _check_cuda_f32(target_mag, "target magnitudes")
    if pred_mag.shape != target_mag.shape or pred_mag.dim() != 3:
        raise ValueError(f"expected two (B, F, T) tensors, got {tuple(pred_mag.shape)} and {tuple(target_mag.shape)}")
    pred_mag, target_mag = pred_mag.contiguous(), target_mag.contiguous()
    B, F, T = pred_mag.shape
    scratch = torch.empty(3 * B, device=pred_mag.device, dtype=torch.float64)
    with torch.cuda.device(pred_mag.device):
        rc = _lib.lib().b200mel_stft_loss_terms(pred_mag.data_ptr(), target_mag.data_ptr(), B, F * T, float(eps),
                                                scratch.data_ptr(), out2.data_ptr(),
                                                C.c_void_p(torch.cuda.current_stream(pred_mag.device).cuda_stream))
    _lib.check(rc)


def logmel_from_magnitude(plan: "_lib.Plan", mag: torch.Tensor, epi: "_lib.Epilogue") -> torch.Tensor:
    """(B, n_fft//2+1, T) magnitudes -> (B, n_mels, T) through the tcgen05 / TMEM mel GEMM (b200mel_logmel_from_magnitude)."""
    _check_cuda_f32(mag, "magnitude")
    if mag.dim() != 3 or mag.shape[1] != plan.cfg.n_fft // 2 + 1:
        raise ValueError(f"expected magnitudes (B, {plan.cfg.n_fft // 2 + 1}, T), got {tuple(mag.shape)}")
    if mag.device.index != plan.device_index:
        raise RuntimeError(f"plan lives on cuda:{plan.device_index}, magnitudes on {mag.device}")
    mag = mag.contiguous()
    B, _, T = mag.shape
    out = torch.empty((B, plan.cfg.n_mels, T), device=mag.device, dtype=torch.float32)
    if B == 0 or T == 0:
        return out
    with torch.cuda.device(mag.device):
        rc = _lib.lib().b200mel_logmel_from_magnitude(plan.handle, mag.data_ptr(), B, T, C.byref(epi), out.data_ptr(),
                                                      C.c_void_p(torch.cuda.current_stream(mag.device).cuda_stream))
    _lib.check(rc)
    return out
