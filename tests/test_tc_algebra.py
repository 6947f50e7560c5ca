"""CPU checks of the tensor-core STFT formulation (pytorch_sound_b200/csrc/stft_tc.cuh): the numpy model of its algebra
against numpy.fft, the fp16 hi / lo limb arithmetic against an fp32 FFT, and the operand blob the library builds
(b200mel_debug_tc_tables, host only) against the model's matrices."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import tc_model as tm  # noqa: E402

from pytorch_sound_b200 import _lib  # noqa: E402


def _ref_mags(span):
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(1024) / 1024)
    fr = np.stack([span[256 * t:256 * t + 1024].astype(np.float64) * w for t in range(8)])
    return np.abs(np.fft.rfft(fr, axis=1))[:, :384].T


def _spans():
    rng = np.random.default_rng(0)
    n = np.arange(tm.SPAN)
    return {
        "noise": 0.1 * rng.standard_normal(tm.SPAN),
        "sine+noise": 0.5 * np.sin(2 * np.pi * 440 / 22050 * n) + 0.01 * rng.standard_normal(tm.SPAN),
        "dc+noise": 0.8 + 1e-3 * rng.standard_normal(tm.SPAN),
        "tiny": 1e-6 * rng.standard_normal(tm.SPAN),
        "loud": 3e4 * rng.standard_normal(tm.SPAN),
        "silence": np.zeros(tm.SPAN),
    }


@pytest.mark.parametrize("name", list(_spans()))
def test_algebra_is_exact(name):
    span = _spans()[name].astype(np.float32)
    r = _ref_mags(span)
    got = tm.group_magnitudes(span, emulate=False)
    assert np.abs(got - r).max() <= 1e-12 * max(np.abs(r).max(), 1e-30)


@pytest.mark.parametrize("name", list(_spans()))
def test_fp16_limbs_match_an_fp32_fft(name):
    span = _spans()[name].astype(np.float32)
    r = _ref_mags(span)
    got = tm.group_magnitudes(span, emulate=True)
    # an fp32 FFT is 1-2e-7 of the largest bin away from float64; three fp16-limb products are within 4e-7
    assert np.abs(got - r).max() <= 6e-7 * max(np.abs(r).max(), 1e-30)


def _blob(W):
    fn = _lib.lib().b200mel_debug_tc_tables
    fn.restype = C.c_int64
    fn.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
    info = np.zeros(4, dtype=np.int32)
    W = np.ascontiguousarray(W, dtype=np.float32)
    n = fn(W.ctypes.data, W.shape[0], W.shape[1], None, 0, info.ctypes.data)
    if n < 0:
        return None, info
    out = np.zeros(n, dtype=np.uint8)
    assert fn(W.ctypes.data, W.shape[0], W.shape[1], out.ctypes.data, n, info.ctypes.data) == n
    return out, info


def _limbs(raw, n_chunks, n_rows):
    """[k chunk][n][8] fp16 -> float64 [n, K]"""
    a = raw.view(np.float16).astype(np.float64).reshape(n_chunks, n_rows, 8)
    return np.transpose(a, (1, 0, 2)).reshape(n_rows, n_chunks * 8)


def test_library_operands_equal_the_model():
    W = _lib.mel_filterbank(22050, 1024, 80, 0.0, 8000.0)
    blob, info = _blob(W)
    assert blob is not None
    b1 = _limbs(blob[:4096], 4, 64)                       # rows: 32 hi columns, 32 lo columns
    F = tm.stage1_matrix()                                # [n1, col]
    assert np.abs((b1[:32] + b1[32:]).T - F).max() < 2.0 ** -21
    assert np.array_equal(b1[:32].T, F.astype(np.float16).astype(np.float64))
    b2 = _limbs(blob[4096:4096 + 24576], 8, 192)
    B, Bp = tm.stage2_matrices()
    assert np.abs((b2[0:48] + b2[96:144]).T - B).max() < 2.0 ** -21
    assert np.abs((b2[48:96] + b2[144:192]).T - Bp).max() < 2.0 ** -21
    tw = blob[4096 + 24576:4096 + 24576 + 17 * 32 * 8].view(np.float32).reshape(17, 32, 2)
    ref = tm.twiddles() * 2.0 ** -tm.S2_SHIFT
    assert np.abs(tw[..., 0] - ref.real).max() < 1e-9 and np.abs(tw[..., 1] - ref.imag).max() < 1e-9


@pytest.mark.parametrize("cfg", [(22050, 80, 8000.0), (22050, 40, 7600.0), (24000, 128, 8000.0), (22050, 7, 3000.0)])
def test_mel_schedule_replays_to_the_dense_filterbank(cfg):
    sr, n_mels, fmax = cfg
    W = _lib.mel_filterbank(sr, 1024, n_mels, 0.0, fmax)
    blob, info = _blob(W)
    assert blob is not None, "filterbank below bin 384 must be eligible"
    mel = blob[4096 + 24576 + 17 * 32 * 8:]
    assert len(mel) == info[0] and info[0] % 16 == 0
    hdr = mel[:1056].view(np.int32)                      # trip[4] | woff[4] | ent[128] {m, lo}
    trip, woff, ent = hdr[:4], hdr[4:8], hdr[8:].reshape(128, 2)
    w = mel[1056:].view(np.float32)
    dense = np.zeros_like(W)
    seen = set()
    for v in range(4):
        for lane in range(32):
            m, lo = ent[32 * v + lane]
            assert lo % 4 == lane % 4                                   # neighbouring lanes start in different banks
            assert lo >= -4 and lo + trip[v] <= 384                     # stays inside the (padded) magnitude tile
            if m >= 0:
                assert m not in seen
                seen.add(int(m))
            for i in range(trip[v]):
                val = w[woff[v] + 32 * i + lane]
                if val != 0:
                    assert m >= 0 and lo + i >= 0
                    dense[m, lo + i] += val
    assert len(seen) == n_mels
    assert np.array_equal(dense, W)
    assert trip[0] >= trip[1] >= trip[2] >= trip[3]                      # rows sorted by length: idle warps come last


def test_filterbank_above_bin_384_is_not_eligible():
    W = _lib.mel_filterbank(22050, 1024, 80, 0.0, None)   # reaches Nyquist = bin 512
    blob, _ = _blob(W)
    assert blob is None
