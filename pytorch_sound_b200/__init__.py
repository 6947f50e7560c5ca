"""pytorch_sound_b200 — B200-native STFT -> magnitude -> mel -> log feature extraction behind
pytorch_sound's operator surface.

    from pytorch_sound_b200.models.transforms import LogMelSpectrogram, STFT, STFTTorchAudio, Audio2Mel
    from pytorch_sound_b200.interface.hifi_gan import MelSpectrogram
    from pytorch_sound_b200.utils.calculate import db2log, norm_mel, unnorm_mel

Same class names, constructor arguments, buffers and output tensors as
pytorch_sound.models.transforms / pytorch_sound.interface.hifi_gan; the work is done by one
hand-written sm_100a kernel launch per clip batch (libb200mel.so, C ABI in include/b200mel.h).
CUDA float32 tensors only; there is no CPU or PyTorch-op fallback.
"""
from . import settings  # noqa: F401

__version__ = "0.1.0"
__all__ = ["settings", "patch_pytorch_sound"]


def patch_pytorch_sound() -> bool:
    """Swap the B200 operators into an importable `pytorch_sound` (monkey patch). Returns False if the
    reference package cannot be imported."""
    from .patch import patch

    return patch()
