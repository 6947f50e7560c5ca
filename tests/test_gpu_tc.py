"""GPU parity of the tcgen05 STFT kernel (csrc/stft_tc.cuh: both radix-32 DFT stages as tensor-core GEMMs on fp16 hi / lo
limbs, accumulators in TMEM).  The kernel is an opt-in path of b200mel_forward (B200MEL_TC=1 / b200mel_debug_set_tc_mode):
every test here switches it on, checks that it really launched, and compares with the float64 oracle, the reference's
golden outputs and the CUDA-core kernel.  Needs a B200: `-m gpu`."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from oracle import mel_oracle as mo

sys.path.insert(0, os.path.dirname(__file__))
import tc_model as tm  # noqa: E402

pytestmark = pytest.mark.gpu

TOL = 1e-4
GEO = dict(sample_rate=22050, mel_size=80, n_fft=1024, win_length=1024, hop_length=256, mel_min=0.0, mel_max=8000.0)


@pytest.fixture()
def tc(built_lib):
    """The library with the tensor-core kernel forced on; yields (torch, lib, launch counter) and restores the mode."""
    import torch

    assert torch.cuda.is_available(), "GPU tests need CUDA"
    lib = built_lib.lib()
    lib.b200mel_debug_set_tc_taps.argtypes = [C.c_void_p] * 4
    lib.b200mel_debug_set_tc_taps.restype = None
    lib.b200mel_debug_set_tc_mode.argtypes = [C.c_int]
    lib.b200mel_debug_set_tc_mode.restype = C.c_int
    lib.b200mel_debug_tc_launch_count.restype = C.c_int64
    prev = lib.b200mel_debug_set_tc_mode(1)
    try:
        yield torch, lib
    finally:
        lib.b200mel_debug_set_tc_taps(None, None, None, None)
        lib.b200mel_debug_set_tc_mode(prev)


def cuda(torch, x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_operands_stage_by_stage(tc):
    """Debug taps of batch 0 (clip 0, frames 0..7) against the numpy model of the kernel, and the magnitudes of every
    frame against numpy.fft in float64: the tensor-core arithmetic (fp16 limbs, three products, fp32 TMEM accumulation)
    is as accurate as an fp32 FFT."""
    torch, lib = tc
    from pytorch_sound_b200.models import transforms as T

    wav = mo.synth_clips(4, 22050, 22050, seed=3)
    x = cuda(torch, wav)
    mod = T.LogMelSpectrogram(**GEO).cuda()
    n_frames = 22050 // 256 + 1
    mag = torch.zeros(4, 384, n_frames, device="cuda")
    d1 = torch.zeros(256, 32, device="cuda")
    d2 = torch.zeros(128, 48, device="cuda")
    lib.b200mel_debug_set_tc_taps(mag.data_ptr(), d1.data_ptr(), d2.data_ptr(), None)
    n0 = lib.b200mel_debug_tc_launch_count()
    mod(x)
    torch.cuda.synchronize()
    assert lib.b200mel_debug_tc_launch_count() == n0 + 1, "the tensor-core kernel did not run"
    taps = {}
    tm.group_magnitudes(mo.reflect_pad(wav[0].astype(np.float64), 512).astype(np.float32)[:tm.SPAN], emulate=True, taps=taps)
    D1, D2 = taps["D1"].reshape(256, 32), taps["D2"].reshape(128, 48)
    assert np.abs(d1.cpu().numpy() - D1).max() < 2e-6 * np.abs(D1).max()
    assert np.abs(d2.cpu().numpy() - D2).max() < 2e-6 * np.abs(D2).max()
    ref = np.abs(mo.stft_complex(wav.astype(np.float64), 1024, 256))[:, :384, :]
    assert np.abs(mag.cpu().numpy() - ref).max() < 1e-6 * ref.max()


@pytest.mark.parametrize("B,L", [(1, 513), (2, 600), (3, 2815), (5, 22050), (37, 8000), (256, 22050)])
def test_logmel_vs_oracle_and_cuda_core_kernel(tc, B, L):
    """Clips of every length class: shorter than one batch of 8 frames, one frame past a batch, the C2 shape.  1e-4 on
    the pre-clamp log-mel against float64 (the north-star tolerance), and within 1e-4 of the CUDA-core kernel."""
    torch, lib = tc
    from pytorch_sound_b200.models import transforms as T

    wav = mo.synth_clips(B, L, 22050, seed=11 + B)
    x = cuda(torch, wav)
    mod = T.LogMelSpectrogram(22050, 80, 1024, 1024, 256, None, None, 0.0, 8000.0).cuda()
    n0 = lib.b200mel_debug_tc_launch_count()
    y = mod(x)
    assert lib.b200mel_debug_tc_launch_count() == n0 + 1
    lib.b200mel_debug_set_tc_mode(0)
    y_cc = mod(x)
    lib.b200mel_debug_set_tc_mode(1)
    assert lib.b200mel_debug_tc_launch_count() == n0 + 1
    sub = slice(0, min(B, 48))
    ref = mo.log_mel_spectrogram(wav[sub].astype(np.float64), **GEO, clamp=False)
    assert y.shape == y_cc.shape and y.shape[1:] == ref.shape[1:]
    assert mo.parity_error(y[sub].cpu().numpy(), ref) < TOL
    assert mo.parity_error(y.cpu().numpy(), y_cc.cpu().numpy()) < TOL      # same mixed abs / rel metric
    assert torch.equal(mod(x), y), "not deterministic"


def test_golden_reference_outputs(tc, golden):
    """The reference's own LogMelSpectrogram / hifi MelSpectrogram outputs (tests/golden/make_golden.py)."""
    torch, lib = tc
    from pytorch_sound_b200.interface.hifi_gan import MelSpectrogram
    from pytorch_sound_b200.models import transforms as T

    x = cuda(torch, golden["clips.wav"])
    n0 = lib.b200mel_debug_tc_launch_count()
    y = T.LogMelSpectrogram(22050, 80, 1024, 1024, 256, -50, 30, 0.0, 8000.0).cuda()(x)
    assert mo.parity_error(y.cpu().numpy(), golden["clips.logmel_clamped"]) < TOL
    h = MelSpectrogram().cuda()(x)        # hifi padding, sqrt(re^2 + im^2 + 1e-9), ln(max(., 1e-5))
    assert mo.parity_error(h.cpu().numpy(), golden["clips.hifi"]) < TOL
    assert lib.b200mel_debug_tc_launch_count() == n0 + 2


def test_scales_silence_and_non_finite(tc):
    """The per-batch power-of-two scale: outputs follow the input over 50 orders of magnitude, silence gives the floor,
    and a NaN sample spoils exactly the frames that contain it, as in the reference (its batch of 8 frames merely
    runs unscaled)."""
    torch, lib = tc
    from pytorch_sound_b200.models import transforms as T

    mod = T.LogMelSpectrogram(22050, 80, 1024, 1024, 256, None, None, 0.0, 8000.0).cuda()
    wav = mo.synth_clips(3, 6000, 22050, seed=5)
    for scale in (1e-30, 1e-12, 1e-4, 1.0, 3e4, 1e20):
        w = (wav.astype(np.float64) * scale).astype(np.float32)
        ref = mo.log_mel_spectrogram(w.astype(np.float64), **GEO, clamp=False)
        assert mo.parity_error(mod(cuda(torch, w)).cpu().numpy(), ref) < TOL, scale
    z = mod(torch.zeros(2, 4000, device="cuda"))
    assert torch.all(z == z[0, 0, 0]) and abs(float(z[0, 0, 0]) - np.log(1e-6)) < 1e-5
    w = wav.copy()
    w[1, 3000] = np.nan
    y = mod(cuda(torch, w)).cpu().numpy()
    ok = mod(cuda(torch, wav)).cpu().numpy()
    assert not np.isfinite(y[1, :, 10:14]).any()               # padded sample 3512 lies in frames 10..13 (the shared
                                                               # epilogue's min / max turn the NaN into -inf)
    rest = np.r_[0:10, 14:y.shape[2]]
    assert mo.parity_error(y[1][:, rest], ok[1][:, rest]) < TOL
    assert np.array_equal(y[1, :, 16:], ok[1, :, 16:]) and np.array_equal(y[[0, 2]], ok[[0, 2]])


def test_power_spectrogram_variant_and_norm_epilogue(tc):
    """LogMelSpectrogramTorchAudio (power 2, HTK scale, no area norm) with a filterbank below bin 384, and the fused
    norm_mel epilogue, through the tensor-core kernel."""
    torch, lib = tc
    from pytorch_sound_b200.models import transforms as T

    wav = mo.synth_clips(6, 12000, 22050, seed=21)
    x = cuda(torch, wav)
    n0 = lib.b200mel_debug_tc_launch_count()
    y = T.LogMelSpectrogramTorchAudio(22050, 64, 1024, 1024, 256, -50, 30, 0.0, 8000.0).cuda()(x)
    ref = mo.log_mel_spectrogram_torchaudio(wav.astype(np.float64), 22050, 64, 1024, 1024, 256, -50, 30, 0.0, 8000.0)
    assert mo.parity_error(y.cpu().numpy(), ref) < TOL
    lm = T.LogMelSpectrogram(22050, 80, 1024, 1024, 256, -50, 30, 0.0, 8000.0).cuda()
    yn = lm(x, norm=True)
    refn = mo.norm_mel(mo.log_mel_spectrogram(wav.astype(np.float64), min_db=-50, max_db=30, **GEO))
    assert mo.parity_error(yn.cpu().numpy(), refn) < TOL
    assert lib.b200mel_debug_tc_launch_count() == n0 + 2


def test_misaligned_and_strided_tensors(tc):
    """Waveform views at every 4-byte offset inside NaN guard floats, and rows with a stride: the staged copy is
    widened to 16-byte boundaries but never leaves the tensor, and the floats the clamp drops come from global memory."""
    torch, lib = tc
    from pytorch_sound_b200.models import transforms as T

    B, L = 5, 22050
    x = cuda(torch, mo.synth_clips(B, L, 22050, seed=77))
    mod = T.LogMelSpectrogram(**GEO).cuda()
    y0 = mod(x)
    for off in (1, 2, 3, 5):
        buf = torch.full((B * L + 8,), float("nan"), device="cuda")
        v = buf[off:off + B * L].view(B, L)
        v.copy_(x)
        assert torch.equal(mod(v), y0), off
    wide = torch.zeros(B, L + 3, device="cuda")
    wide[:, :L] = x
    assert torch.equal(mod(wide[:, :L]), y0)


def test_ineligible_plans_and_arguments_use_the_cuda_core_kernels(tc):
    """hop 128, a filterbank up to Nyquist, per-clip lengths: the launch silently stays on the CUDA-core kernels."""
    torch, lib = tc
    from pytorch_sound_b200.models import transforms as T

    x = cuda(torch, mo.synth_clips(4, 9000, 22050, seed=2))
    n0 = lib.b200mel_debug_tc_launch_count()
    T.LogMelSpectrogram(22050, 80, 1024, 1024, 128, -50, 30, 0.0, 8000.0).cuda()(x)
    T.LogMelSpectrogram(22050, 80, 1024, 1024, 256, -50, 30, 0.0, None).cuda()(x)
    T.LogMelSpectrogram(**GEO).cuda()(x, lengths=torch.full((4,), 9000, device="cuda", dtype=torch.int32))
    assert lib.b200mel_debug_tc_launch_count() == n0
