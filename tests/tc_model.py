"""numpy model of the tcgen05 STFT kernel (pytorch_sound_b200/csrc/logmel_tc.cuh), operand by operand.

Test infrastructure: tests/test_tc_algebra.py checks this model against numpy.fft (the algebra) and the host tables the
library exports (b200mel_debug_tc_tables) against the tables built here; the GPU tests compare the kernel's debug taps
with the intermediate arrays this model returns.

The transform: a 1024-point real DFT with a periodic Hann window, n = 32 n1 + n2, k = k1 + 32 k2,

  stage 1 (GEMM, un-windowed)   A[k1, n2]   = sum_n1 W32^(k1 n1) x[32 n1 + n2]               k1 = 0..16 (real input)
  twiddle                       A'[k1, n2]  = W1024^(k1 n2) A[k1, n2]
  Hann as a 3-tap in k1         Aw'[k1, n2] = 0.5 A'[k1] - 0.25 A'[k1-1] - 0.25 A'[k1+1]      A'[-1] = conj A'[1],
                                                                                             A'[17] = W32^n2 conj A'[15]
  stage 2 (GEMM)                X[k1 + 32 k2] = sum_n2 W32^(n2 k2) Aw'[k1, n2]

Rows k1 = 1..15 of stage 2 produce k2 in {0..11} u {20..31}; the upper twelve are the conjugates of bins
(32 - k1) + 32 (31 - k2), so magnitudes of all bins < 384 with k1 not in {0, 16} come from 15 rows.  Rows 0 and 16 have
real stage-1 outputs; they share ONE packed row (re slot: Aw'[0, n2], im slot: r16[n2] with Aw'[16, n2] = W64^n2 r16[n2])
that is multiplied by a second matrix B' whose columns produce X[32 k2] from the re slot and X[16 + 32 k2] from the im slot.
"""
from __future__ import annotations

import numpy as np

N = 1024
HOP = 256
GROUP = 8                 # frames per group (one UMMA core-matrix of rows)
SPAN = (GROUP - 1) * HOP + N   # 2816 samples staged per group
QLEN = SPAN // 32         # 88 fp16 per Hankel row
J = 24                    # complex outputs per stage-2 row
S2_SHIFT = 6              # stage-2 operands are scaled by 2^-6 (|Aw'| <= 32 max|x|)


def k2_of_j(j):
    return j if j < 12 else j + 8


def stage1_matrix():
    """[32 n1, 32 cols]: col 0 = A[0].re, col 1 = A[16].re, col 2 k1 / 2 k1 + 1 = A[k1].re / .im (k1 = 1..15)."""
    n1 = np.arange(32, dtype=np.float64)
    F = np.zeros((32, 32))
    F[:, 0] = 1.0
    F[:, 1] = np.cos(np.pi * n1)
    for k1 in range(1, 16):
        th = 2 * np.pi * ((k1 * np.arange(32)) % 32) / 32
        F[:, 2 * k1] = np.cos(th)
        F[:, 2 * k1 + 1] = -np.sin(th)
    return F


def stage2_matrices():
    """B, B' as [64 = (n2, c), 48 = (j, re/im)] float64."""
    B = np.zeros((64, 2 * J))
    Bp = np.zeros((64, 2 * J))
    n2 = np.arange(32)
    for j in range(J):
        th = 2 * np.pi * ((n2 * k2_of_j(j)) % 32) / 32
        B[0::2, 2 * j] = np.cos(th)
        B[1::2, 2 * j] = np.sin(th)
        B[0::2, 2 * j + 1] = -np.sin(th)
        B[1::2, 2 * j + 1] = np.cos(th)
        if j < 12:   # U[k2 = j] = X[32 j] from the re slot
            Bp[0::2, 2 * j] = np.cos(th)
            Bp[0::2, 2 * j + 1] = -np.sin(th)
        else:        # V = X[16 + 32 (23 - j)] from the im slot
            ph = 2 * np.pi * ((n2 * (16 + 32 * (23 - j))) % 1024) / 1024
            Bp[1::2, 2 * j] = np.cos(ph)
            Bp[1::2, 2 * j + 1] = -np.sin(ph)
    return B, Bp


def twiddles():
    """tw[k1, n2] = W1024^(k1 n2), k1 = 0..16, complex128 (the kernel's table carries the 2^-S2_SHIFT scale too)."""
    k1 = np.arange(17)[:, None]
    n2 = np.arange(32)[None, :]
    return np.exp(-2j * np.pi * ((k1 * n2) % 1024) / 1024)


def bin_of(slot, j):
    """spectrum bin whose magnitude row `slot` (0 = packed, else k1), column j of stage 2 carries."""
    if j < 12:
        return slot + 32 * j
    return (16 if slot == 0 else 32 - slot) + 32 * (23 - j)


def split_f16(v):
    hi = v.astype(np.float16)
    lo = (v.astype(np.float32) - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def mm3(a, b, emulate):
    """a @ b the way the kernel forms it: fp16 hi/lo limbs, hi*hi + lo*hi + hi*lo, fp32 accumulation."""
    if not emulate:
        return a.astype(np.float64) @ b.astype(np.float64)
    ah, al = split_f16(a.astype(np.float32))
    bh, bl = split_f16(b.astype(np.float32))
    f = lambda u: u.astype(np.float32)
    return (f(ah) @ f(bh) + f(al) @ f(bh) + f(ah) @ f(bl)).astype(np.float64)


def input_scale(xmax):
    """power of two s with xmax * s in [2^13, 2^14); 1 for silence (the kernel: exponent arithmetic on the max's bits)."""
    if xmax == 0 or not np.isfinite(xmax):
        return 1.0
    e = int(np.floor(np.log2(xmax)))
    return float(2.0 ** (13 - e))


def group_magnitudes(span, emulate=True, taps=None):
    """|STFT| of the 8 frames of one group.  span: SPAN padded-coordinate samples (frame t = span[256 t : 256 t + 1024]).
    Returns mags [384 bins, 8 frames].  taps (dict) receives the intermediate operands."""
    span = np.asarray(span, dtype=np.float32)
    assert span.shape == (SPAN,)
    s_in = input_scale(float(np.max(np.abs(span))))
    xs = span * np.float32(s_in)
    # Hankel buffer S[n2][q] = xs[32 q + n2]; row (n2, t) of the stage-1 A operand = S[n2][8 t : 8 t + 32]
    S = xs.reshape(QLEN, 32).T.copy()
    A_op = np.stack([S[n2, 8 * t:8 * t + 32] for n2 in range(32) for t in range(GROUP)])  # [(n2, t), n1]
    D1 = mm3(A_op, stage1_matrix(), emulate).reshape(32, GROUP, 32)                         # [n2, t, col]
    if not emulate:
        D1 = D1.astype(np.float64)
    else:
        D1 = D1.astype(np.float32).astype(np.float64)
    A = np.zeros((17, 32, GROUP), dtype=np.complex128)       # [k1, n2, t]
    A[0] = D1[:, :, 0]
    A[16] = D1[:, :, 1]
    for k1 in range(1, 16):
        A[k1] = D1[:, :, 2 * k1] + 1j * D1[:, :, 2 * k1 + 1]
    tw = twiddles()
    sc = 2.0 ** -S2_SHIFT
    Ap = A * (tw * sc)[:, :, None]
    Aw = np.zeros((16, 32, GROUP), dtype=np.complex128)      # slots: 0 packed, 1..15 = k1
    for k1 in range(1, 16):
        lo_n = Ap[k1 - 1]
        hi_n = Ap[k1 + 1] if k1 < 15 else Ap[16]
        Aw[k1] = 0.5 * Ap[k1] - 0.25 * lo_n - 0.25 * hi_n
    a0 = 0.5 * Ap[0].real - 0.5 * Ap[1].real
    tw1 = tw[1][:, None]
    r16 = 0.5 * sc * A[16].real - 0.5 * sc * (np.conj(tw1) * A[15]).real
    Aw[0] = a0 + 1j * r16
    if emulate:
        Aw = Aw.astype(np.complex64).astype(np.complex128)
    # stage-2 A operand rows (slot, t), K = (n2, c)
    A2 = np.zeros((16, GROUP, 64))
    A2[:, :, 0::2] = np.transpose(Aw.real, (0, 2, 1))
    A2[:, :, 1::2] = np.transpose(Aw.imag, (0, 2, 1))
    A2 = A2.reshape(16 * GROUP, 64)
    B, Bp = stage2_matrices()
    D2 = mm3(A2, B, emulate).reshape(16, GROUP, 2 * J)
    D2p = mm3(A2, Bp, emulate).reshape(16, GROUP, 2 * J)
    D2[0] = D2p[0]
    mags = np.zeros((384, GROUP))
    descale = 2.0 ** S2_SHIFT / s_in
    for slot in range(16):
        for j in range(J):
            re, im = D2[slot, :, 2 * j], D2[slot, :, 2 * j + 1]
            mags[bin_of(slot, j)] = np.sqrt(re * re + im * im) * descale
    if taps is not None:
        taps.update(s_in=s_in, S=S, D1=D1, Aw=Aw, A2=A2, D2=D2)
    return mags
