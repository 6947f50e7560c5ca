"""Round-2 GPU parity: transform sizes below 1024, multi_stft_loss, the fast-path / generic kernel pair, misaligned
tensors (the bulk-copy clamp), the fused frame mask.  Golden values come from the reference itself
(tests/golden/reference_round2.npz, made by tests/golden/make_golden.py --round2-only).  Needs a B200: `-m gpu`."""
import os

import numpy as np
import pytest

from oracle import mel_oracle as mo

pytestmark = pytest.mark.gpu

TOL = 1e-4
GEO = dict(sample_rate=22050, mel_size=80, n_fft=1024, win_length=1024, hop_length=256, mel_min=0.0, mel_max=8000.0)


@pytest.fixture(scope="module")
def golden2():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_round2.npz"))


@pytest.fixture(scope="module")
def torch_cuda(built_lib):
    import torch

    assert torch.cuda.is_available(), "GPU tests need CUDA"
    return torch


def cuda(torch, x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_small_transform_sizes_vs_reference(torch_cuda, golden2):
    """STFT(filter_length=512 / 256 / 128), STFTTorchAudio(win 400 in n_fft 512), LogMelSpectrogram / Audio2Mel at
    n_fft 512 against the reference's outputs (the kernel runs them as every r-th bin of the 1024-point transform
    of the zero-extended frame)."""
    torch = torch_cuda
    from pytorch_sound_b200.models import transforms as T

    x = cuda(torch, golden2["clips.wav"])
    for n, hop in ((512, 128), (256, 64), (128, 100)):
        st = T.STFT(filter_length=n, hop_length=hop).cuda()
        mag, phase = st.transform(x)
        ref = golden2[f"clips.stft{n}_mag"]
        assert mag.shape == ref.shape
        assert np.abs(mag.cpu().numpy() - ref).max() < 3e-6 * ref.max()
        assert np.abs(st.magnitude(x).cpu().numpy() - ref).max() < 3e-6 * ref.max()
        strong = ref > 1e-2 * ref.max()
        d = np.angle(np.exp(1j * (phase.cpu().numpy().astype(np.float64) - golden2[f"clips.stft{n}_phase"])))
        assert np.abs(d[strong]).max() < 1e-3
    re, im = T.STFTTorchAudio(filter_length=400, hop_length=160, win_length=400, n_fft=512).cuda()(x)
    scale = np.abs(golden2["clips.stfta_win400_fft512_re"]).max()
    assert re.shape == golden2["clips.stfta_win400_fft512_re"].shape
    assert np.abs(re.cpu().numpy() - golden2["clips.stfta_win400_fft512_re"]).max() < 3e-6 * scale
    assert np.abs(im.cpu().numpy() - golden2["clips.stfta_win400_fft512_im"]).max() < 3e-6 * scale
    x16 = cuda(torch, golden2["c16.wav"])
    y = T.LogMelSpectrogram(16000, 40, 512, 512, 128, -50, 30, 0.0, 8000.0).cuda()(x16)
    assert y.shape == golden2["c16.logmel512"].shape
    assert mo.parity_error(y.cpu().numpy(), golden2["c16.logmel512"]) < TOL
    a = T.Audio2Mel(n_fft=512, hop_length=128, win_length=512, sampling_rate=16000, n_mel_channels=40).cuda()(x16.unsqueeze(1))
    assert mo.parity_error(a.cpu().numpy(), golden2["c16.audio2mel512"]) < TOL
    # and against the float64 oracle at sizes the fixtures do not hold
    xs = mo.synth_clips(3, 3000, 22050, seed=5)
    for n, hop, win in ((64, 16, 64), (32, 8, 32), (512, 100, 300), (2048, 300, 1200)):
        mag = T.STFT(filter_length=n, hop_length=hop, win_length=win).cuda().magnitude(cuda(torch, xs)).cpu().numpy()
        rm, _ = mo.stft_transform(xs, n, hop, win)
        assert mag.shape == rm.shape and np.abs(mag - rm).max() < 3e-6 * rm.max(), (n, hop, win)
    with pytest.raises(ValueError):
        T.STFT(filter_length=4096, hop_length=1024).cuda().magnitude(cuda(torch, xs))  # not built: loud, no fallback
    with pytest.raises(ValueError):
        T.STFT(filter_length=400, hop_length=100).cuda().magnitude(cuda(torch, xs))


def test_multi_stft_loss_vs_reference(torch_cuda, golden2):
    """models/sound.py:120-147: the three scalars the reference returned on seeded (pred, target), at 1e-4."""
    torch = torch_cuda
    from pytorch_sound_b200.models.sound import multi_stft_loss

    pred, target = cuda(torch, golden2["loss.pred"]), cuda(torch, golden2["loss.target"])
    params = [tuple(int(v) for v in row) for row in golden2["loss.params"]]
    got = [float(v) for v in multi_stft_loss(pred, target, params)]
    np.testing.assert_allclose(got, golden2["loss.values"], rtol=1e-4)
    np.testing.assert_allclose(got, mo.multi_stft_loss(golden2["loss.pred"], golden2["loss.target"], params), rtol=1e-4)
    got = [float(v) for v in multi_stft_loss(pred, target, [(512, 512, 128)], eps=1e-3)]
    np.testing.assert_allclose(got, golden2["loss.values_512"], rtol=1e-4)
    assert got[0] == pytest.approx(got[1] + got[2], rel=1e-6)
    zero = [float(v) for v in multi_stft_loss(target, target, params)]
    assert zero == [0.0, 0.0, 0.0]


def test_fast_and_generic_kernels_agree(torch_cuda):
    """b200mel_forward picks the compile-time-specialised kernel (logmel_fast.cuh) for the common geometry — its kLen
    instance when `lengths` / the frame mask are asked for — and the generic body otherwise: separately compiled
    instances of the same arithmetic.  The generic body is reached here through the fused pre-emphasis prologue with
    a coefficient of 1e-30, which leaves every fp32 sample unchanged.  Complete frame pairs agree to a few ulp of the
    log-mel value (measured 2.4e-6); the odd last frame of a clip is paired with the reflected continuation of the
    clip in the fast kernel and with zeros in the generic one, so the fp32 rounding of its pair partner differs
    (measured 1.4e-5 on noise-floor bands).  Bar: 3e-5, a third of the parity tolerance — and each kernel is within
    1e-4 of the float64 oracle on its own.  The kLen instance with full lengths is bit-equal to the plain one."""
    torch = torch_cuda
    from pytorch_sound_b200.models import transforms as T

    for geo, L in ((GEO, 22050), (dict(GEO, sample_rate=16000), 8000), (dict(GEO, mel_max=None), 5000)):
        x = cuda(torch, mo.synth_clips(9, L, geo["sample_rate"], seed=31))
        lm = T.LogMelSpectrogram(**geo).cuda()
        full = torch.full((9,), L, device="cuda", dtype=torch.int32)
        y_fast, y_len, y_gen = lm(x), lm(x, lengths=full), lm(x, preemphasis=1e-30)
        assert torch.equal(y_len, y_fast)
        assert float((y_fast - y_gen).abs().max()) < 3e-5
        ref = mo.log_mel_spectrogram(x.cpu().numpy(), **geo, clamp=False)
        assert mo.parity_error(y_fast.cpu().numpy(), ref) < TOL and mo.parity_error(y_gen.cpu().numpy(), ref) < TOL


def test_misaligned_and_tight_tensors(torch_cuda):
    """The bulk copies are widened to 16-byte boundaries but clamped to the tensor (include/b200mel.h): a waveform
    tensor whose first / last byte is not 16-byte aligned (a view at an odd float offset into a larger buffer, rows
    of 22050 floats) gives bit-identical results to an aligned copy, for the mel and the spectrum kernels."""
    torch = torch_cuda
    from pytorch_sound_b200.models import transforms as T

    B, L = 5, 22050
    x = cuda(torch, mo.synth_clips(B, L, 22050, seed=77))
    lm = T.LogMelSpectrogram(**GEO).cuda()
    st = T.STFT(filter_length=1024, hop_length=256).cuda()
    y0, m0 = lm(x), st.magnitude(x)
    full = torch.full((B,), L, device="cuda", dtype=torch.int32)
    g0 = lm(x, lengths=full)
    for off in (1, 2, 3, 5):
        buf = torch.full((B * L + 8,), float("nan"), device="cuda")  # NaN guard floats around the view
        v = buf[off:off + B * L].view(B, L)
        v.copy_(x)
        assert v.data_ptr() % 16 == (4 * off) % 16
        assert torch.equal(lm(v), y0), off
        assert torch.equal(lm(v, lengths=full), g0), off
        assert torch.equal(st.magnitude(v), m0), off
    # strided rows: every row start has its own alignment
    wide = torch.zeros(B, L + 3, device="cuda")
    wide[:, :L] = x
    assert torch.equal(lm(wide[:, :L]), y0)


def test_frame_mask_from_the_same_launch(torch_cuda, golden2):
    """LogMelSpectrogram(..., frame_mask=True): the (B, T) SpectrogramMasker mask (models/transforms.py:397-416)
    written by the mel launch itself, against the reference's own masker output."""
    torch = torch_cuda
    from pytorch_sound_b200.models import transforms as T

    lens = golden2["masker.1024_256.lengths"]
    L = int(golden2["masker.1024_256.L"])
    x = cuda(torch, mo.synth_clips(len(lens), L, 22050, seed=3))
    lengths = torch.from_numpy(lens.astype(np.int32)).cuda()
    for i, n in enumerate(lens):
        x[i, int(n):] = 0
    lm = T.LogMelSpectrogram(**GEO).cuda()
    y, fmask = lm(x, lengths=lengths, frame_mask=True)
    np.testing.assert_array_equal(fmask.cpu().numpy(), golden2["masker.1024_256.out"])
    assert torch.equal(y, lm(x, lengths=lengths))
    _, ones = lm(x, frame_mask=True)  # without lengths every frame is valid
    assert bool((ones == 1).all())


def test_logmelscale_tcgen05_vs_oracle(torch_cuda):
    """LogMelScale (models/transforms.py:247-268; the reference class itself raises TypeError at construction): the mel
    filterbank as a tcgen05 tensor-core GEMM (bf16 hi/lo split, fp32 accumulator in TMEM) on magnitudes in HBM, against
    the float64 product of the same filterbank and the same fp32 magnitudes — so only the GEMM's own error is seen."""
    torch = torch_cuda
    from pytorch_sound_b200.models import transforms as T

    for sr, n_mels, fmax, B, L in ((22050, 80, 8000.0, 7, 22050), (16000, 80, 8000.0, 33, 8000), (22050, 40, None, 3, 3000),
                                   (22050, 96, 8000.0, 2, 5000)):
        x = mo.synth_clips(B, L, sr, seed=11 + B)
        x[0] *= 1e-4  # a quiet clip: bands far below the log offset
        mag = T.STFT(filter_length=1024, hop_length=256).cuda().magnitude(cuda(torch, x))
        lms = T.LogMelScale(sr, n_mels, 1024, -50, 30, 0.0, fmax).cuda()
        y = lms(mag)
        assert y.shape == (B, n_mels, mag.shape[2]) and y.dtype == torch.float32
        fb = mo.mel_filterbank(sr, 1024, n_mels, 0.0, fmax).astype(np.float64)
        ref = np.log(np.einsum("mf,bft->bmt", fb, mag.cpu().numpy().astype(np.float64)) + 1e-6)
        ref = np.clip(ref, mo.db2log(-50), mo.db2log(30))
        err = mo.parity_error(y.cpu().numpy(), ref)
        print(f"LogMelScale tcgen05 sr={sr} mels={n_mels}: max err {err:.2e}")
        assert err < TOL
        # and it agrees with the fused kernel's own filterbank path on the same waveform
        lm = T.LogMelSpectrogram(sr, n_mels, 1024, 1024, 256, -50, 30, 0.0, fmax).cuda()
        assert float((lm(cuda(torch, x)) - y).abs().max()) < 2e-4
    # tile boundaries: B * T not a multiple of 128, one clip, one frame
    for B, Tn in ((1, 1), (1, 129), (5, 127), (3, 300)):
        mag = torch.rand(B, 513, Tn, device="cuda") * 3
        y = T.LogMelScale(22050, 80, 1024, -50, 30, 0.0, 8000.0).cuda()(mag)
        fb = mo.mel_filterbank(22050, 1024, 80, 0.0, 8000.0).astype(np.float64)
        ref = np.clip(np.log(np.einsum("mf,bft->bmt", fb, mag.cpu().numpy().astype(np.float64)) + 1e-6), mo.db2log(-50), mo.db2log(30))
        assert mo.parity_error(y.cpu().numpy(), ref) < TOL, (B, Tn)
    # a filterbank whose bf16 limbs do not fit in shared memory next to the pipeline buffers: loud, no fallback
    with pytest.raises(ValueError, match="too large"):
        T.LogMelScale(22050, 128, 1024, -50, 30, 0.0, None).cuda()(torch.rand(1, 513, 4, device="cuda"))


def test_fused_preemphasis_prologue(torch_cuda):
    """LogMelSpectrogram(..., preemphasis=c): models.sound.PreEmphasis (models/sound.py:66-81) applied while the samples
    are staged.  Same fmaf as the standalone kernel and the same generic extraction kernel afterwards, so where the
    two-pass path also runs the generic kernel (n_fft 2048, hop 300) it is BIT-equal to PreEmphasis ->
    LogMelSpectrogram, including reflected edge frames, per-clip lengths and misaligned rows; at the common geometry the
    two-pass path takes the compile-time specialised kernel, a separately compiled instance of the same arithmetic
    (bar 3e-5 as in test_fast_and_generic_kernels_agree).  And it is within tolerance of the float64 oracle chain."""
    torch = torch_cuda
    from pytorch_sound_b200.models import transforms as T
    from pytorch_sound_b200.models.sound import PreEmphasis

    for geo, L in ((GEO, 22050), (GEO, 1500), (dict(sample_rate=44100, mel_size=128, n_fft=2048, win_length=2048, hop_length=512), 9000),
                   (dict(GEO, hop_length=300), 5000)):
        B = 5
        x = cuda(torch, mo.synth_clips(B, L, geo["sample_rate"], seed=5 + L))
        lm = T.LogMelSpectrogram(**geo).cuda()
        full = torch.full((B,), L, device="cuda", dtype=torch.int32)
        pre = PreEmphasis(0.97).cuda()(x.unsqueeze(1))[:, 0]
        two_pass = lm(pre, lengths=full)             # extraction of the pre-emphasised waveform
        fused = lm(x, lengths=full, preemphasis=0.97)
        if geo["n_fft"] == 1024 and geo["hop_length"] == 256:
            assert float((fused - two_pass).abs().max()) < 3e-5, (geo, L)
        else:
            assert torch.equal(fused, two_pass), (geo, L)
        assert float((lm(x, preemphasis=0.97) - fused).abs().max()) == 0.0  # without lengths: same generic kernel
        ref = mo.log_mel_spectrogram(mo.pre_emphasis(x.cpu().numpy()[:, None, :], 0.97)[:, 0], **geo, clamp=False)
        assert mo.parity_error(fused.cpu().numpy(), ref) < TOL
    # ragged lengths: each clip is pre-emphasised and reflected within its own length
    lens = [6000, 4500, 5999, 700]
    x = torch.zeros(4, 6000, device="cuda")
    clips = [mo.synth_clips(1, n, 22050, seed=40 + i)[0] for i, n in enumerate(lens)]
    for i, c in enumerate(clips):
        x[i, :len(c)] = torch.from_numpy(c).cuda()
    lm = T.LogMelSpectrogram(**GEO).cuda()
    y = lm(x, lengths=torch.tensor(lens, device="cuda", dtype=torch.int32), preemphasis=0.97)
    for i, c in enumerate(clips):
        Ti = 1 + len(c) // 256
        ref = mo.log_mel_spectrogram(mo.pre_emphasis(c[None, None, :], 0.97)[:, 0], **GEO, clamp=False)[0]
        assert mo.parity_error(y[i, :, :Ti].cpu().numpy(), ref) < TOL
    buf = torch.zeros(5 * 22050 + 8, device="cuda")
    v = buf[1:1 + 5 * 22050].view(5, 22050)
    xs = cuda(torch, mo.synth_clips(5, 22050, 22050, seed=9))
    v.copy_(xs)
    assert torch.equal(lm(v, preemphasis=0.97), lm(xs, preemphasis=0.97))


def test_fused_mfcc_epilogue(torch_cuda):
    """MFCC.forward (models/transforms.py:433-455) as ONE launch (io.out_mfcc: the DCT applied while the log-mel column
    is on chip) against the reference's MFCC output, the float64 oracle and the two-launch path (mel kernel + DCT
    kernel); n_mfcc 40 (two coefficient slots per lane), 13 and 64, an odd mel count; geometries the fused kernel does
    not serve answer B200MEL_EUNSUP and the module falls back to two launches."""
    torch = torch_cuda
    from pytorch_sound_b200 import _lib, functional
    from pytorch_sound_b200.models import transforms as T

    extra = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_extra.npz"))
    x = cuda(torch, extra["clips.wav"])
    mf = T.MFCC(22050, 80, 1024, 1024, 40, 256, -50, 30, 0.0, 8000.0).cuda()
    n0 = _lib.launch_count()
    y = mf(x)
    assert _lib.launch_count() == n0 + 1, "the fused epilogue is one launch"
    assert np.abs(y.cpu().numpy() - extra["clips.mfcc"]).max() < 1e-3     # 80-term sums of log values, each within 1e-4
    two = functional.mel_to_mfcc(mf.mel_func(x), mf.dct_mat)
    assert float((y - two).abs().max()) < 2e-5                            # same log-mel values, different summation order
    # mel frames and coefficients from the same launch
    epi = _lib.make_epilogue(_lib.LOG_LN_OFFSET, 1e-6, mf.mel_func.min_db, mf.mel_func.max_db, False)
    y2, mel = functional.mfcc_fused(mf.mel_func._plan(x.device), x, epi, mf.dct_mat, want_mel=True)
    assert torch.equal(y2, y) and torch.equal(mel, mf.mel_func(x))
    wav = mo.synth_clips(7, 9001, 22050, seed=4)
    for n_mels, n_mfcc in ((80, 13), (80, 64), (79, 40), (40, 40)):
        m = T.MFCC(22050, n_mels, 1024, 1024, n_mfcc, 256, -50, 30, 0.0, 8000.0).cuda()
        n0 = _lib.launch_count()
        got = m(cuda(torch, wav)).cpu().numpy()
        ref_mel = mo.log_mel_spectrogram(wav, 22050, n_mels, 1024, 1024, 256, -50, 30, 0.0, 8000.0)
        ref = np.einsum("km,bmt->bkt", mo.create_dct(n_mfcc, n_mels).T, ref_mel)
        assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-3, (n_mels, n_mfcc)
    # hop 128 / a filterbank the specialised kernel has no instance for: two launches, same numbers
    m = T.MFCC(22050, 80, 1024, 1024, 40, 128, -50, 30, 0.0, 8000.0).cuda()
    n0 = _lib.launch_count()
    got = m(cuda(torch, wav)).cpu().numpy()
    assert _lib.launch_count() == n0 + 2
    ref = np.einsum("km,bmt->bkt", mo.create_dct(40, 80).T, mo.log_mel_spectrogram(wav, 22050, 80, 1024, 1024, 128, -50, 30, 0.0, 8000.0))
    assert np.abs(got - ref).max() < 1e-3


def test_feature_loader_prefetch(torch_cuda):
    """GpuFeatureLoader(prefetch=True): batch i + 1 is copied and extracted on a side stream while the consumer works on
    batch i — same batches, same order, bit-equal features, also when the consumer keeps its stream busy and when
    it runs on a non-default stream."""
    torch = torch_cuda
    from pytorch_sound_b200.data.feature_loader import GpuFeatureLoader
    from pytorch_sound_b200.models import transforms as T

    lm = T.LogMelSpectrogram(**GEO)
    rng = np.random.default_rng(3)
    batches = []
    for i in range(7):
        n = 3000 + 500 * i
        wav = torch.from_numpy((0.1 * rng.standard_normal((4 + i, n))).astype(np.float32)).pin_memory()
        mask = torch.ones(4 + i, n).pin_memory()
        mask[0, n // 2:] = 0
        batches.append([wav, mask])
    plain = [[t.clone() for t in b] for b in GpuFeatureLoader(batches, [(0, lm)], mask_index=-1)]
    burn = torch.randn(2048, 2048, device="cuda")
    for stream in (torch.cuda.current_stream(), torch.cuda.Stream()):
        with torch.cuda.stream(stream):
            got = []
            for b in GpuFeatureLoader(batches, [(0, lm)], mask_index=-1, prefetch=True):
                burn = burn @ burn * 1e-3          # the consumer's own work on its stream
                got.append([t.clone() for t in b])
        torch.cuda.synchronize()
        assert len(got) == len(plain)
        for a, b in zip(got, plain):
            assert len(a) == len(b) == 3
            assert all(torch.equal(x, y) for x, y in zip(a, b))
