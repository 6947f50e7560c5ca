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
        spec_kind: int = _lib.SPEC_NONE, lengths: Optional[torch.Tensor] = None
        ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Launch the fused kernel on wav's device / current stream. Returns (mel, out_a, out_b)."""
    wav = _check_wav(wav)
    dev = wav.device
    if dev.index != plan.device_index:
        raise RuntimeError(f"plan lives on cuda:{plan.device_index}, wav on {dev}")
    B, L = wav.shape
    T = plan.out_frames(L)
    n_freq = plan.cfg.n_fft // 2 + 1
    mel = out_a = out_b = None
    if want_mel:
        mel = torch.empty((B, plan.cfg.n_mels, T), device=dev, dtype=torch.float32)
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
        rc = _lib.lib().b200mel_forward(
            plan.handle, wav.data_ptr(), B, L, wav.stride(0) if B > 1 else max(L, 1), len_ptr,
            C.byref(epi) if epi is not None else None, mel.data_ptr() if mel is not None else None, spec_kind,
            out_a.data_ptr() if out_a is not None else None, out_b.data_ptr() if out_b is not None else None,
            C.c_void_p(stream))
    _lib.check(rc)
    return mel, out_a, out_b
