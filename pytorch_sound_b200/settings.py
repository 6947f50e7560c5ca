"""Audio / STFT constants of pytorch_sound (pytorch_sound/settings.py:9-22), mirrored value for value.

When the reference package is importable, `from_reference()` reads the live values instead, so a
user who edited settings.py ("If you want to change sound settings, Change settings.py",
README.md:111) gets the same geometry here.  The reference's settings module pulls in text
cleaners (unidecode) at import; failure to import simply keeps the mirrored defaults.
"""
SAMPLE_RATE: int = 22050  # settings.py:9
N_FFT: int = 1024  # :10
WIN_LENGTH: int = 1024  # :11
HOP_LENGTH: int = 256  # :12
HOP_STRIDE: int = WIN_LENGTH // HOP_LENGTH  # :13
SPEC_SIZE: int = WIN_LENGTH // 2 + 1  # :14
MEL_SIZE: int = 80  # :15
MFCC_SIZE: int = 40  # :16
MEL_MIN: int = 0  # :17
MEL_MAX: int = 8000  # :18
MIN_DB: int = -50  # :19
MAX_DB: int = 30  # :20
VN_DB: float = -11.5  # :21

_NAMES = ("SAMPLE_RATE", "N_FFT", "WIN_LENGTH", "HOP_LENGTH", "HOP_STRIDE", "SPEC_SIZE", "MEL_SIZE", "MFCC_SIZE",
          "MEL_MIN", "MEL_MAX", "MIN_DB", "MAX_DB", "VN_DB")


def from_reference() -> bool:
    """Overwrite the mirrored constants with pytorch_sound.settings' if that module imports."""
    try:
        from pytorch_sound import settings as ref  # type: ignore
    except Exception:
        return False
    g = globals()
    for name in _NAMES:
        if hasattr(ref, name):
            g[name] = getattr(ref, name)
    return True


def logmel_kwargs() -> dict:
    """Constructor kwargs of LogMelSpectrogram at the settings geometry."""
    return dict(sample_rate=SAMPLE_RATE, mel_size=MEL_SIZE, n_fft=N_FFT, win_length=WIN_LENGTH,
                hop_length=HOP_LENGTH, min_db=MIN_DB, max_db=MAX_DB, mel_min=float(MEL_MIN), mel_max=float(MEL_MAX))
