"""CPU model of the index algebra of csrc/fft32.cuh + logmel_kernel.cuh (no GPU, numpy only).

It replays, lane by lane and slot by slot, what a warp does to one task — pair packing of two real frames,
radix-32 pass over n1 with the digit-reversed output slots of fft32_pos(), transpose, inter-pass twiddle table in
its [n2/2][lane][2] layout, radix-32 pass over n2, conjugate-symmetry separation with the partner-lane / slot
formulas of the kernel (including the special cases of lane 0 and bin 512), and the split pass of n_fft = 2048 —
and checks the result against numpy's rfft.  A refactor of the kernel's indexing can be tried here first."""
import numpy as np


def fft32_pos(k):  # fft32.cuh: slot of output bin k after the in-place 4 x 8 decomposition (an involution)
    return 8 * (k & 3) + (k & 4) + (k >> 3)


def fft32_slots(a):
    """32-point forward DFT of the 32 slots of every lane; output bin k lands in slot fft32_pos(k)."""
    A = np.fft.fft(a, axis=-1)
    out = np.empty_like(A)
    for k in range(32):
        out[..., fft32_pos(k)] = A[..., k]
    return out


def twiddle_table():
    """b200mel.cu: tw[((j >> 1) * 32 + lane) * 2 + (j & 1)] = exp(-2 pi i j lane / 1024)."""
    tw = np.zeros(32 * 32, dtype=np.complex128)
    for j in range(32):
        for lane in range(32):
            tw[((j >> 1) * 32 + lane) * 2 + (j & 1)] = np.exp(-2j * np.pi * j * lane / 1024)
    return tw


def warp_fft1024(z):
    """1024-point complex FFT as the warp computes it.  z[n], n = 32 n1 + n2.  Returns Z as [lane = k1][slot]
    with Z[k1 + 32 k2] at slot fft32_pos(k2)."""
    a = np.empty((32, 32), dtype=np.complex128)  # pass 1: lane = n2, slot j = n1
    for lane in range(32):
        for j in range(32):
            a[lane, j] = z[32 * j + lane]
    a = fft32_slots(a)                            # Y[k1] of lane n2 at slot fft32_pos(k1)
    buf = np.empty((32, 32), dtype=np.complex128)
    for lane in range(32):                        # transpose buffer: row k1, column n2
        for k1 in range(32):
            buf[k1, lane] = a[lane, fft32_pos(k1)]
    tw = twiddle_table()
    b = np.empty((32, 32), dtype=np.complex128)   # read back: lane = k1, slot j = n2, twiddle W^(n2 k1)
    for lane in range(32):
        for j in range(32):
            b[lane, j] = buf[lane, j] * tw[((j >> 1) * 32 + lane) * 2 + (j & 1)]
    return fft32_slots(b)


def test_fft32_pos_is_an_involution_and_a_permutation():
    assert sorted(fft32_pos(k) for k in range(32)) == list(range(32))
    assert all(fft32_pos(fft32_pos(k)) == k for k in range(32))


def test_two_pass_fft_matches_numpy():
    rng = np.random.default_rng(0)
    z = rng.standard_normal(1024) + 1j * rng.standard_normal(1024)
    Z = warp_fft1024(z)
    ref = np.fft.fft(z)
    for k1 in range(32):
        for k2 in range(32):
            assert abs(Z[k1, fft32_pos(k2)] - ref[k1 + 32 * k2]) < 1e-9


def test_pair_mode_separation_matches_rfft_of_both_frames():
    """Frames t, t+1 packed as re / im (window pre-scaled by 1/2); lane k1 pairs its Z[k] with Z[1024 - k] held by
    lane (32 - k1) & 31: that lane's slot 31 - k2, or for lane 0 its own slot (32 - k2) & 31."""
    rng = np.random.default_rng(1)
    x0, x1 = rng.standard_normal(1024), rng.standard_normal(1024)
    Z = warp_fft1024(0.5 * (x0 + 1j * x1))
    X0, X1 = np.fft.rfft(x0), np.fft.rfft(x1)
    for lane in range(32):
        partner = (32 - lane) & 31
        for k2 in range(16):
            A = Z[lane, fft32_pos(k2)]
            # what the partner lane sends: lane 0 reads itself and needs slot (32 - k2) & 31, others slot 31 - k2
            src_slot = fft32_pos((32 - k2) & 31) if partner == 0 else fft32_pos(31 - k2)
            B = Z[partner, src_slot]
            E, D = A + np.conj(B), A - np.conj(B)
            k = lane + 32 * k2
            assert abs(E - X0[k]) < 1e-9                 # frame t
            assert abs(-1j * D - X1[k]) < 1e-9           # frame t+1 = -i (A - conj(B)); the kernel only needs |D|
    A = Z[0, fft32_pos(16)]                              # bin 512 is its own partner
    assert abs(2 * A.real - X0[512]) < 1e-9 and abs(2 * A.imag - X1[512]) < 1e-9


def test_split_mode_matches_rfft_2048():
    """n_fft = 2048: even / odd samples packed as re / im of a 1024-point FFT, then
    X[k] = E + W_2048^k O, X[1024 - k] = conj(E - W_2048^k O), W_2048^k = W_2048^lane * W_64^k2."""
    rng = np.random.default_rng(2)
    x = rng.standard_normal(2048)
    Z = warp_fft1024(0.5 * (x[0::2] + 1j * x[1::2]))
    X = np.fft.rfft(x)
    for lane in range(32):
        partner = (32 - lane) & 31
        wl = np.exp(-2j * np.pi * lane / 2048)
        for k2 in range(16):
            A = Z[lane, fft32_pos(k2)]
            B = Z[partner, fft32_pos((32 - k2) & 31) if partner == 0 else fft32_pos(31 - k2)]
            E, D = A + np.conj(B), A - np.conj(B)
            Wk = wl * np.exp(-2j * np.pi * k2 / 64)
            P = D * (-1j * Wk)
            k = lane + 32 * k2
            assert abs((E + P) - X[k]) < 1e-9
            assert abs(np.conj(E - P) - X[1024 - k]) < 1e-9
    A = Z[0, fft32_pos(16)]
    assert abs((2 * A.real - 2j * A.imag) - X[512]) < 1e-9


def test_generated_hann_window_matches_the_table():
    """load_windowed_pair: 0.5 w[32 j + lane] = 0.25 - t, 0.5 w[32 (j + 16) + lane] = 0.25 + t with
    t = 0.25 cos(2 pi j / 32 + 2 pi lane / 1024), evaluated in float32 like the kernel."""
    n = np.arange(1024)
    table = (0.5 * (0.5 - 0.5 * np.cos(2 * np.pi * n / 1024))).astype(np.float32)
    c32 = np.cos(2 * np.pi * np.arange(16) / 32).astype(np.float32)
    s32 = (-np.sin(2 * np.pi * np.arange(16) / 32)).astype(np.float32)
    worst = 0.0
    for lane in range(32):
        cs = np.float32(0.25) * np.float32(np.cos(np.pi * lane / 512))
        sn = np.float32(0.25) * np.float32(np.sin(np.pi * lane / 512))
        for j in range(16):
            t = np.float32(s32[j] * sn + np.float32(c32[j] * cs))
            worst = max(worst, abs(float(np.float32(0.25) - t) - float(table[32 * j + lane])),
                        abs(float(np.float32(0.25) + t) - float(table[32 * (j + 16) + lane])))
    assert worst < 6e-8  # a few float32 ulps of 0.25; relative to the window's peak of 0.5 that is 1.2e-7
