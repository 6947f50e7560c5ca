"""Monkey-patch the B200 operators into an importable `pytorch_sound` package so that existing user
code (`from pytorch_sound.models.transforms import LogMelSpectrogram`, `InterfaceHifiGAN().encode`)
picks them up unchanged.  See INTEGRATION.md."""


def patch() -> bool:
    try:
        import pytorch_sound.models.transforms as ref_t  # type: ignore
    except Exception:
        return False
    from .models import transforms as t

    for name in ("STFT", "LogMelSpectrogram", "STFTTorchAudio", "Audio2Mel", "LogMelSpectrogramTorchAudio", "MelToMFCC",
                 "MFCC", "SpectrogramMasker"):
        setattr(ref_t, "_reference_" + name, getattr(ref_t, name, None))
        setattr(ref_t, name, getattr(t, name))
    try:
        import pytorch_sound.models.sound as ref_s  # type: ignore
        from .models import sound as snd

        ref_s._reference_PreEmphasis = ref_s.PreEmphasis
        ref_s.PreEmphasis = snd.PreEmphasis
    except Exception:
        pass
    try:
        import pytorch_sound.interface.hifi_gan as ref_h  # type: ignore
        from .interface import hifi_gan as h

        ref_h._reference_MelSpectrogram = ref_h.MelSpectrogram
        ref_h.MelSpectrogram = h.MelSpectrogram
    except Exception:
        pass
    return True
