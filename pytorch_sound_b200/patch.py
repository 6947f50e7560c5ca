"""Monkey-patch the B200 operators into an importable `pytorch_sound` package so that existing user
code (`from pytorch_sound.models.transforms import LogMelSpectrogram`, `InterfaceHifiGAN().encode`)
picks them up unchanged.  See INTEGRATION.md.

What gets installed is a HYBRID of each reference class, not a bare replacement: a subclass of the reference
class, constructed by the reference's own `__init__` (so every buffer, `STFT.inverse`, `STFTTorchAudio.inverse`
and the autograd path are still there), carrying a B200 twin.  The analysis-direction methods run the twin —
one fused kernel launch — when the input is a CUDA tensor that needs no gradient, and the reference's own
implementation otherwise (CPU tensors such as `InterfaceHifiGAN(device='cpu')`, or `multi_stft_loss(pred, ...)`
during training, where `pred` requires grad).  `patch()` also reads the live `pytorch_sound.settings` values into
`pytorch_sound_b200.settings` (an edited settings.py is honoured by `norm_mel` and `logmel_kwargs`).
"""
import sys
import weakref

import torch


def _use_kernel(x) -> bool:
    return isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and \
        not (x.requires_grad and torch.is_grad_enabled())


def make_hybrid(ref_cls, b200_cls, methods):
    """Subclass of `ref_cls` whose `methods` dispatch to a `b200_cls` twin for CUDA no-grad inputs.
    Falls back to `b200_cls` itself when the reference symbol is not an nn.Module class (stand-ins)."""
    if not (isinstance(ref_cls, type) and issubclass(ref_cls, torch.nn.Module)):
        return b200_cls

    def init(self, *args, **kwargs):
        ref_cls.__init__(self, *args, **kwargs)
        twin = b200_cls(*args, **kwargs)
        twin._fb_owner = weakref.ref(self)       # the kernel follows THIS module's (reference-named) filterbank buffer
        self.__dict__['_b200'] = twin            # not a registered sub-module: state_dict stays the reference's

    ns = {'__init__': init, '__doc__': (ref_cls.__doc__ or '') + '\n[pytorch_sound_b200 hybrid: CUDA no-grad inputs run '
          'the fused sm_100a kernel, everything else the reference implementation]', '_reference_class': ref_cls,
          '_b200_class': b200_cls}
    for name in methods:
        if not hasattr(ref_cls, name) or not hasattr(b200_cls, name):
            continue

        def dispatch(self, x, *args, __name=name, **kwargs):
            if _use_kernel(x):
                return getattr(self.__dict__['_b200'], __name)(x, *args, **kwargs)
            return getattr(ref_cls, __name)(self, x, *args, **kwargs)

        dispatch.__name__ = name
        ns[name] = dispatch
    return type(ref_cls.__name__, (ref_cls,), ns)


_TRANSFORMS = {  # class -> analysis-direction methods served by the kernel
    "STFT": ("transform",),
    "LogMelSpectrogram": ("forward",),
    "STFTTorchAudio": ("forward", "transform"),
    "Audio2Mel": ("forward",),
    "LogMelSpectrogramTorchAudio": ("forward",),
    "MelToMFCC": ("forward",),
    "MFCC": ("forward",),
}


def patch() -> bool:
    try:
        import pytorch_sound.models.transforms as ref_t  # type: ignore
    except Exception:
        return False
    from . import settings
    from .models import transforms as t

    settings.from_reference()
    for name, methods in _TRANSFORMS.items():
        ref_cls = getattr(ref_t, "_reference_" + name, None) or getattr(ref_t, name, None)
        setattr(ref_t, "_reference_" + name, ref_cls)
        setattr(ref_t, name, make_hybrid(ref_cls, getattr(t, name), methods))
    # SpectrogramMasker: same result on any device, and the reference hard-codes .cuda() in __init__ — plain swap
    ref_t._reference_SpectrogramMasker = getattr(ref_t, "_reference_SpectrogramMasker", None) or \
        getattr(ref_t, "SpectrogramMasker", None)
    ref_t.SpectrogramMasker = t.SpectrogramMasker

    # models.sound / interface.hifi_gan are only touched when they import (or were imported) cleanly; sound.py binds
    # `STFT = STFTTorchAudio` at import time, so its alias is re-pointed at the hybrid (which still differentiates).
    try:
        import pytorch_sound.models.sound as ref_s  # type: ignore
        from .models import sound as snd

        ref_s._reference_PreEmphasis = getattr(ref_s, "_reference_PreEmphasis", None) or ref_s.PreEmphasis
        ref_s.PreEmphasis = make_hybrid(ref_s._reference_PreEmphasis, snd.PreEmphasis, ("forward",))
        if getattr(ref_s, "STFT", None) is ref_t._reference_STFTTorchAudio:
            ref_s.STFT = ref_t.STFTTorchAudio
    except Exception:
        pass
    try:
        import pytorch_sound.interface.hifi_gan as ref_h  # type: ignore
        from .interface import hifi_gan as h

        ref_h._reference_MelSpectrogram = getattr(ref_h, "_reference_MelSpectrogram", None) or ref_h.MelSpectrogram
        ref_h.MelSpectrogram = make_hybrid(ref_h._reference_MelSpectrogram, h.MelSpectrogram, ("forward",))
    except Exception:
        pass
    return True


def unpatch() -> None:
    """Put the reference classes back (tests; A/B comparisons)."""
    for mod_name in ("pytorch_sound.models.transforms", "pytorch_sound.models.sound", "pytorch_sound.interface.hifi_gan"):
        mod = sys.modules.get(mod_name)
        if mod is None:
            continue
        for key in [k for k in vars(mod) if k.startswith("_reference_")]:
            orig = getattr(mod, key)
            if orig is not None:
                setattr(mod, key[len("_reference_"):], orig)
            delattr(mod, key)
