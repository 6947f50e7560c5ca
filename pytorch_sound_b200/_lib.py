"""ctypes binding of libb200mel.so (the C ABI declared in include/b200mel.h).

There is no fallback: if the shared library is missing this module raises at import
of the first symbol user, and every compute entry point raises when CUDA is absent.
"""
import ctypes as C
import os
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libb200mel.so")

OK, EINVAL, ECUDA, ENODEV, ENOMEM, EUNSUP = 0, -1, -2, -3, -4, -5
PAD_CENTER, PAD_HIFI = 0, 1
MEL_SLANEY, MEL_HTK = 0, 1
NORM_NONE, NORM_SLANEY = 0, 1
LOG_NONE, LOG_LN_OFFSET, LOG_LN_FLOOR, LOG_LOG10_FLOOR = 0, 1, 2, 3
SPEC_NONE, SPEC_MAG_PHASE, SPEC_RE_IM, SPEC_MAG = 0, 1, 2, 3


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("sample_rate", C.c_int32), ("n_fft", C.c_int32), ("win_length", C.c_int32),
        ("hop_length", C.c_int32), ("n_mels", C.c_int32), ("fmin", C.c_float), ("fmax", C.c_float),
        ("pad_mode", C.c_int32), ("mel_scale", C.c_int32), ("mel_norm", C.c_int32), ("power", C.c_int32),
        ("mag_eps", C.c_float),
    ]


class Epilogue(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("log_kind", C.c_int32), ("log_arg", C.c_float), ("has_clamp_lo", C.c_int32),
        ("clamp_lo", C.c_float), ("has_clamp_hi", C.c_int32), ("clamp_hi", C.c_float), ("norm_mel", C.c_int32),
    ]


class IO(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("spec_kind", C.c_int32), ("wav", C.c_void_p), ("B", C.c_int64), ("L", C.c_int64),
        ("row_stride", C.c_int64), ("lengths", C.c_void_p), ("out_mel", C.c_void_p), ("out_a", C.c_void_p),
        ("out_b", C.c_void_p), ("out_frame_mask", C.c_void_p), ("reserve_sms", C.c_int32), ("preemphasis", C.c_float),
        ("dct_mat", C.c_void_p), ("out_mfcc", C.c_void_p), ("n_mfcc", C.c_int32), ("reserved0", C.c_int32),
    ]


# every symbol include/b200mel.h declares: (restype, argtypes)
SYMBOLS = {
    "b200mel_version": (C.c_int, []),
    "b200mel_last_error": (C.c_char_p, []),
    "b200mel_mel_filterbank": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32,
                                         C.c_int32, C.c_void_p]),
    "b200mel_hann_window": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p]),
    "b200mel_out_frames": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "b200mel_plan_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "b200mel_plan_set_filterbank": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "b200mel_plan_destroy": (C.c_int, [C.c_void_p]),
    "b200mel_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                  C.POINTER(Epilogue), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200mel_forward_io": (C.c_int, [C.c_void_p, C.POINTER(IO), C.POINTER(Epilogue), C.c_void_p]),
    "b200mel_logmel_from_magnitude": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(Epilogue),
                                                C.c_void_p, C.c_void_p]),
    "b200mel_forward_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.POINTER(Epilogue),
                                       C.c_void_p, C.c_void_p]),
    "b200mel_launch_count": (C.c_int64, []),
    "b200mel_preemphasis": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_int64,
                                      C.c_void_p]),
    "b200mel_volume_norm": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200mel_stft_loss_terms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "b200mel_gather_pull": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, C.c_int32,
                                      C.POINTER(C.c_int64), C.c_int32, C.c_void_p]),
    "b200mel_gather_tma": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, C.c_int32,
                                     C.POINTER(C.c_int64), C.c_int32, C.c_void_p]),
    "b200mel_gather_copy": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, C.c_int32,
                                      C.POINTER(C.c_int64), C.c_void_p]),
    "b200mel_mel_to_mfcc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int64, C.c_void_p,
                                      C.c_void_p]),
}

_lib = None
_lock = threading.Lock()


def lib():
    """Load libb200mel.so once.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        f"{LIB_PATH} is missing: build it with `python -m pytorch_sound_b200.build` "
                        "(nvcc, sm_100a). pytorch_sound_b200 has no CPU / PyTorch fallback.")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in SYMBOLS.items():
                    fn = getattr(handle, name)  # AttributeError if the .so is stale
                    fn.restype, fn.argtypes = res, args
                _lib = handle
    return _lib


class B200MelError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc == OK:
        return
    msg = lib().b200mel_last_error().decode("utf-8", "replace")
    if rc in (EINVAL, EUNSUP):
        raise ValueError(f"b200mel: {msg}")
    raise B200MelError(f"b200mel (code {rc}): {msg}")


def make_config(sample_rate, n_fft, win_length, hop_length, n_mels, fmin=0.0, fmax=None, pad_mode=PAD_CENTER,
                mel_scale=MEL_SLANEY, mel_norm=NORM_SLANEY, power=1, mag_eps=0.0) -> Config:
    return Config(C.sizeof(Config), int(sample_rate), int(n_fft), int(win_length), int(hop_length), int(n_mels),
                  float(fmin or 0.0), float(fmax) if fmax else 0.0, int(pad_mode), int(mel_scale), int(mel_norm),
                  int(power), float(mag_eps))


def make_epilogue(log_kind=LOG_NONE, log_arg=0.0, clamp_lo=None, clamp_hi=None, norm_mel=False) -> Epilogue:
    return Epilogue(C.sizeof(Epilogue), int(log_kind), float(log_arg), int(clamp_lo is not None),
                    float(clamp_lo or 0.0), int(clamp_hi is not None), float(clamp_hi or 0.0), int(bool(norm_mel)))


def mel_filterbank(sample_rate, n_fft, n_mels, fmin=0.0, fmax=None, mel_scale=MEL_SLANEY, mel_norm=NORM_SLANEY):
    """Host helper (no GPU): librosa.filters.mel restated in C++ -> numpy float32 (n_mels, n_fft//2+1)."""
    import numpy as np

    out = np.empty((n_mels, n_fft // 2 + 1), dtype=np.float32)
    check(lib().b200mel_mel_filterbank(int(sample_rate), int(n_fft), int(n_mels), float(fmin or 0.0),
                                       float(fmax) if fmax else 0.0, mel_scale, mel_norm, out.ctypes.data))
    return out


def hann_window(win_length, n_fft=None):
    import numpy as np

    n_fft = n_fft or win_length
    out = np.empty(n_fft, dtype=np.float32)
    check(lib().b200mel_hann_window(int(win_length), int(n_fft), out.ctypes.data))
    return out


class Plan:
    """Owns one b200mel_plan (device tables) on one CUDA device."""

    def __init__(self, cfg: Config, device_index: int):
        import torch

        if not torch.cuda.is_available():
            raise B200MelError("b200mel: CUDA is not available and there is no CPU fallback")
        self.cfg = cfg
        self.device_index = device_index
        self._h = C.c_void_p()
        with torch.cuda.device(device_index):
            check(lib().b200mel_plan_create(C.byref(cfg), C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def out_frames(self, L: int) -> int:
        T = C.c_int64()
        check(lib().b200mel_out_frames(self._h, int(L), C.byref(T)))
        return T.value

    def set_filterbank(self, weights) -> None:
        import numpy as np

        w = np.ascontiguousarray(weights, dtype=np.float32)
        check(lib().b200mel_plan_set_filterbank(self._h, w.ctypes.data, w.shape[0], w.shape[1]))

    def close(self):
        if self._h:
            lib().b200mel_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_plan_cache = {}
_plan_lock = threading.Lock()


def cached_plan(device_index: int, **kw) -> Plan:
    """Plans are immutable, so modules with equal geometry on one device share one."""
    key = (device_index,) + tuple(sorted(kw.items()))
    with _plan_lock:
        pl = _plan_cache.get(key)
        if pl is None:
            pl = Plan(make_config(**kw), device_index)
            _plan_cache[key] = pl
        return pl


def launch_count() -> int:
    return int(lib().b200mel_launch_count())
