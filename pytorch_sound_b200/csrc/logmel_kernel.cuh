// logmel_kernel.cuh — the fused STFT -> |.| -> mel -> log kernel (sm_100a).
//
// Persistent grid: one CTA per SM, 16 warps per CTA, every warp an independent pipeline that walks the
// task list with stride (#warps in the grid; the warps of a CTA take tasks gridDim apart so a partial
// last round spreads over all SMs).  A task is one 1024-point complex FFT held in
// registers (32 complex values per lane, two radix-32 passes, one shared-memory transpose):
//   kPair  (n_fft = 1024): frames (2q, 2q+1) of a clip packed as re/im, separated by conjugate symmetry;
//   !kPair (n_fft = 2048): frame q packed even/odd, finished by the real-input split pass.
//
// Per-warp shared-memory region (bytes):
//   [0, 4160)              magnitude tile   (pair: float2[520] = {|X_t[k]|, |X_t+1[k]|}; split: float[1032])
//   [4224, 4224 + stage)   sample stage     (filled by ONE cp.async.bulk = TMA 1-D copy per task, mbarrier-signalled)
//   [0, 8704)              transpose buffer (float2[32][34]) — overlaps both, live only between the two
//                          radix-32 passes, i.e. after the stage was consumed and before the next TMA is issued.
// The TMA for task i+1 is issued right after the transpose of task i, so the copy lands while the warp
// does its second FFT pass, the separation, the mel contraction and the epilogue of task i.
//
// CTA-shared tables (fetched once per CTA by four bulk copies on one mbarrier): inter-pass twiddles, window,
// banded mel filterbank in a lane-balanced schedule (rows sorted by length, 32 rows per round, weights padded
// to float4 groups).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200mel.h"
#include "fft32.cuh"

// Optional phase timing (tools/phase_timing.py): compile with -DB200MEL_PHASE_TIMING; every warp accumulates the
// clock64() deltas of its phases into dbg[phase] (one atomicAdd per phase per task).
#ifdef B200MEL_PHASE_TIMING
// light-weight: one 32-bit clock read per mark, per-thread register accumulators, flushed once at kernel end
#define PHASE_MARK(i)                                  \
    do {                                               \
        const unsigned now_ = (unsigned)clock();       \
        phase_acc_[i] += now_ - tmark_;                \
        tmark_ = now_;                                 \
    } while (0)
#else
#define PHASE_MARK(i) \
    do {              \
    } while (0)
#endif

namespace b200mel {

#ifndef B200MEL_WARPS_PER_CTA
#define B200MEL_WARPS_PER_CTA 16
#endif
// warps per CTA of the mel kernel: 16 -> 128 registers per thread, 20 -> 96 (register budget 65536 / (32 * warps))
constexpr int kMaxWarps = B200MEL_WARPS_PER_CTA;
#ifndef B200MEL_MEL_CHUNK
#define B200MEL_MEL_CHUNK (B200MEL_WARPS_PER_CTA > 16 ? 4 : 8)
#endif
#ifndef B200MEL_XPOSE_BATCH
#define B200MEL_XPOSE_BATCH (B200MEL_WARPS_PER_CTA > 16 ? 8 : 16)
#endif
// float2 per transposed row: 272 B = 17 x 16 B, so a lane reads ITS row with 128-bit loads (two values per LDS)
// and the 8 lanes of a quarter warp hit 8 different 16-byte bank groups (17 l mod 8 = l); the write side stores
// 32 consecutive float2 per row, conflict-free for any stride.
constexpr int kBufStride = 34;
constexpr int kXposeBytes = 32 * kBufStride * 8;          // 8704
constexpr int kTileBytes = 4160;                          // magnitude tile
constexpr int kStageOff = 4224;                           // stage offset inside the warp region (128B aligned)
constexpr int kPairTileLen = 520;                         // float2 entries (513 used, tail zeroed)
constexpr int kSplitTileLen = 1032;                       // float entries (1025 used, tail zeroed)

constexpr int kMaxMelRounds = 8;  // 32 rows per round -> up to 256 mel rows

struct MelEntry {  // one filterbank row as seen by one lane in one round; 8 bytes so a warp reads its 32 entries with
                   // one conflict-free 64-bit load (16-byte records read field by field were 4-way conflicts)
    int lo;        // first spectrum bin of the row's read window (16-byte aligned in the tile, slid for bank spread)
    int m;         // mel row index, -1 = idle lane
};

struct KParams {
    const float *wav;
    // first / last 16-byte boundary INSIDE the caller's tensor [wav, wav + (B-1) row_stride + L): bulk copies are
    // widened to 16-byte boundaries but never leave this range (the <= 3 floats cut off at a misaligned tensor
    // edge are fetched by the halo patch instead)
    unsigned long long wav_lo16, wav_hi16;
    long long row_stride;
    long long B;
    int L;
    const int *lengths;
    int T, hop, pad;
    int n_fft;        // PHYSICAL transform size the kernel runs: 1024 (pair) or 2048 (split)
    int n_fft_log;    // logical n_fft of the plan: a power of two <= n_fft.  A 1024 / r-point DFT is bin r k of the
                      // 1024-point DFT of the zero-extended frame, so smaller transforms run the same kernel with a
                      // window table that is zero beyond n_fft_log samples, and use every r-th bin
    int bin_step;     // r = n_fft / n_fft_log
    int n_freq_out;   // n_fft_log / 2 + 1: rows of the spectrum outputs
    int pair_frames;  // frames per task: 2 (pair mode) or 1 (split mode, or pair mode with hop > n_fft)
    int hann_full;    // 1: win_length == n_fft, the kernel may generate the periodic Hann instead of reading the table
    // global copies of the CTA tables
    const float *window;    // [n_fft], periodic Hann centre-padded, pre-scaled by 0.5
    const float2 *tw;       // [16][32][2]  tw[((j >> 1) * 32 + lane) * 2 + (j & 1)] = exp(-2 pi i j lane / 1024)
    const float2 *tw_post;  // [32]      exp(-2 pi i lane / 2048)            (split mode)
    const MelEntry *mel_entries;  // [rounds][32]
    const float *mel_w;           // [mel_w_len]
    int n_mels, n_freq, mel_rounds, mel_w_len;
    int round_groups[kMaxMelRounds];  // float4 groups every lane runs in round r (rows padded with zero weights)
    int round_wbase[kMaxMelRounds];   // float4 index of the round's weights, stored [group][lane]
    // shared memory layout (bytes from the dynamic smem base)
    int off_window, off_entries, off_melw, off_bar, off_regions, region_bytes, stage_bytes;
    // outputs
    float *out_mel, *out_a, *out_b;
    float preemph;     // != 0: pre-emphasis y[n] = x[n] - preemph x[n-1] (y[0] = x[0] - preemph x[1]) applied to the staged
                       // samples before framing (PreEmphasis.forward, models/sound.py:66-81), generic mel kernel only
    // fused MFCC epilogue (logmel_fast_kernel<..., kDct = true> only): dct (n_mfcc, n_mels) row-major in global memory,
    // its shared-memory copy [n_mels rounded up to even][64] at off_dct, the per-warp log-mel column float2[...] at off_col
    const float *dct;
    float *out_mfcc;
    int n_mfcc, off_dct, off_col, col_bytes;
    float *out_fmask;  // nullable (B, T): SpectrogramMasker frame mask, 1 iff t * hop - win_half < clip length
    int win_half;
    long long *dbg;  // phase-timing accumulators (debug builds only)
    float mag_eps;
    // branch-free epilogue: y = min(max(lg2(max(x, floor) + offset) * log_scale, lo), hi) * norm_scale + norm_bias
    int use_log;
    float ep_floor, ep_offset, log_scale, lo, hi, norm_scale, norm_bias;
    // task list: task = b * tasks_per_clip + q; the grid-wide warp stride is pre-split as stride_b * tasks_per_clip + stride_q
    long long n_tasks;
    int tasks_per_clip, stride_b, stride_q;
};

// ------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Waits for the phase with the given parity.  The spin is bounded (seconds): a protocol bug or a lost bulk copy ends in
// a trap, i.e. a CUDA error on the host, instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 24)) __trap();
    } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // one MUFU.SQRT, max rel. error 2^-23
    return y;
}

// Frames of a clip of Li valid samples (`lengths` mode).  A clip that is too short to be reflect-padded (Li <= pad; the
// reference's F.pad would raise) has no frames: its rows of the output are written as zeros.
__device__ __forceinline__ int frames_of(int Li, int n_fft, int hop, int pad) {
    int span = Li + 2 * pad - n_fft;
    return (span < 0 || Li <= pad) ? 0 : span / hop + 1;
}
__device__ __forceinline__ int reflect_index(int i, int Li) {
    if (i < 0) i = -i;
    if (i >= Li) i = 2 * (Li - 1) - i;
    return min(max(i, 0), Li - 1);
}

// ln / log10 through one MUFU.LG2 (abs. error ~1e-6 on the log value, two orders below the 1e-4 tolerance);
// clamps are +-inf and the affine is the identity when the caller did not ask for them.
__device__ __forceinline__ float epilogue(float x, const KParams &p) {
    float y = x;
    if (p.use_log) {  // warp-uniform
        float t = fmaxf(x, p.ep_floor) + p.ep_offset;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(t));
        y *= p.log_scale;
    }
    y = fminf(fmaxf(y, p.lo), p.hi);
    return fmaf(y, p.norm_scale, p.norm_bias);
}

// magnitudes of two complex values at once: {|u|, |v|} (power 1) or {|u|^2, |v|^2} (power 2)
template <int kPower>
__device__ __forceinline__ float2 magnitude2(float2 u, float2 v, float eps) {
    float2 sq = make_float2(fmaf(u.x, u.x, u.y * u.y), fmaf(v.x, v.x, v.y * v.y));
    if constexpr (kPower == 2) return sq;
    sq = __fadd2_rn(sq, make_float2(eps, eps));
    return make_float2(sqrt_approx(sq.x), sqrt_approx(sq.y));
}

// Pair-mode separation and magnitudes of one bin in four packed instructions (+ the two square roots):
//   E = A + conj(B) is frame t, D = A - conj(B) has |D| = |frame t+1| (the -i rotation is never formed);
//   {E.x, D.x} = {A.x, A.x} + {B.x, -B.x},  {E.y, D.y} = {A.y, A.y} + {-B.y, B.y},
//   {|E|^2, |D|^2} = {E.x, D.x}^2 + {E.y, D.y}^2  — the same roundings as the scalar form x*x + (y*y).
template <int kPower>
__device__ __forceinline__ float2 pair_magnitudes(float2 A, float2 Bv, float eps) {
    const float2 px = __fadd2_rn(make_float2(A.x, A.x), make_float2(Bv.x, -Bv.x));
    const float2 py = __fadd2_rn(make_float2(A.y, A.y), make_float2(-Bv.y, Bv.y));
    float2 sq = __ffma2_rn(px, px, __fmul2_rn(py, py));
    if constexpr (kPower == 2) return sq;
    sq = __fadd2_rn(sq, make_float2(eps, eps));
    return make_float2(sqrt_approx(sq.x), sqrt_approx(sq.y));
}

template <int kPower>
__device__ __forceinline__ float magnitude(float re, float im, float eps) {
    const float sq = fmaf(re, re, im * im);
    if constexpr (kPower == 2) return sq;
    return sqrt_approx(sq + eps);
}

// Geometry of the ONE bulk copy that stages a task's samples, computed identically by the request side and the
// halo patch.  Padded-coordinate sample s of the span [s_first, s_first + span) sits at stage[s - s_first + delta];
// delta in 0..3 makes stage and global address 16-byte congruent, as cp.async.bulk requires.  The copy covers the
// in-range samples [max(s_first, 0), min(s_first + span, Li)) widened to 16-byte boundaries on both sides — the
// extra <= 3 floats per side are neighbouring samples of the same tensor, never read as data — and is clamped to
// the tensor: when a widened end would leave [wav_lo16, wav_hi16) (first / last row of a tensor whose ends are
// not 16-byte aligned) that end moves inwards by 16 bytes and the samples it drops are marked uncovered.
struct CopyGeom {
    int c_lo, c_hi;        // padded-coordinate samples [c_lo, c_hi) are valid in the stage once the copy has landed
    int delta;
    int dst;               // stage index (floats) the copy starts at, a multiple of 4
    uint32_t bytes;        // multiple of 16; 0 = nothing to copy
    const void *src;       // 16-byte aligned global address
    bool patch;            // part of the span is not covered by the copy: reflect halo and / or clamped floats
};
__device__ __forceinline__ CopyGeom copy_geom(const KParams &p, long long b, int s_first, int span, int Li) {
    CopyGeom g;
    const int p_lo = max(s_first, 0), p_hi = min(s_first + span, Li);
    const uintptr_t a = reinterpret_cast<uintptr_t>(p.wav + b * p.row_stride + p_lo);
    const int mis = (int)(a >> 2) & 3;     // floats by which the first in-range sample misses a 16-byte boundary
    const int off = p_lo - s_first;        // its position in the span
    g.delta = (mis - off) & 3;
    uintptr_t a16 = a - 4 * (uintptr_t)mis;
    uintptr_t e16 = (a + 4 * (uintptr_t)(p_hi - p_lo) + 15) & ~(uintptr_t)15;
    g.c_lo = p_lo, g.c_hi = p_hi;
    g.dst = off + g.delta - mis;
    if (a16 < p.wav_lo16) a16 += 16, g.dst += 4, g.c_lo = p_lo + 4 - mis;
    if (e16 > p.wav_hi16) e16 -= 16, g.c_hi = p_lo + (int)((e16 - a) >> 2);
    g.src = reinterpret_cast<const void *>(a16);
    g.bytes = e16 > a16 ? (uint32_t)(e16 - a16) : 0u;
    if (g.bytes == 0u) g.c_hi = g.c_lo;
    g.patch = s_first < g.c_lo || s_first + span > g.c_hi;
    return g;
}
// Fill what the bulk copy did not cover, after it has landed: the reflected halo of an edge task (<= 4 of 44
// tasks per 1-s clip) and the floats dropped by the tensor-edge clamp.  The source sample is almost always
// inside the staged part, so it is copied within shared memory; only a reflection that leaves the staged span
// (or a clamped float) falls back to global memory.
// sample r of a row as the FFT sees it: the raw sample, or the pre-emphasised one (same fmaf as preemph_kernel)
__device__ __forceinline__ float row_sample(const float *row, int r, float preemph) {
    const float v = __ldg(row + r);
    return preemph != 0.f ? fmaf(-preemph, __ldg(row + (r > 0 ? r - 1 : 1)), v) : v;
}
__device__ __forceinline__ void patch_stage(const KParams &p, long long b, int s_first, int span, int Li,
                                            const CopyGeom &g, float *stage, int lane, float preemph = 0.f) {
    const float *row = p.wav + b * p.row_stride;
    float *st = stage + g.delta - s_first;  // st[s] = sample at padded-coordinate position s
    for (int s = s_first + lane; s < g.c_lo; s += 32) {
        const int r = reflect_index(s, Li);
        st[s] = (r >= g.c_lo && r < g.c_hi) ? st[r] : row_sample(row, r, preemph);
    }
    for (int s = g.c_hi + lane; s < s_first + span; s += 32) {
        const int r = reflect_index(s, Li);
        st[s] = (r >= g.c_lo && r < g.c_hi) ? st[r] : row_sample(row, r, preemph);
    }
    __syncwarp();
}
// Fused pre-emphasis (PreEmphasis.forward, models/sound.py:66-81, as a prologue of the extraction): the staged samples
// [c_lo, c_hi) are replaced in place by y[s] = x[s] - c x[s-1] (y[0] = x[0] - c x[1], the reference's 1-sample reflect
// pad) BEFORE the halo patch, so the reflected halo mirrors the pre-emphasised signal as F.pad of the reference's
// output would.  Chunks of 256 samples are processed from the top down: a chunk reads its samples and the one just
// below it (still raw — lower chunks come later) into registers, then writes.  Out of line: its registers must not
// weigh on the allocation of the FFT loop it is called from.
__device__ __noinline__ void preemphasize_stage(const float *row, float *st, int c_lo, int c_hi, float coef, int lane) {
    const int n = c_hi - c_lo;
    for (int base = ((n - 1) >> 8) << 8; base >= 0; base -= 256) {
        float cur[8], prev[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int s = c_lo + base + 32 * i + lane;
            if (s < c_hi) {
                cur[i] = st[s];
                prev[i] = s > c_lo ? st[s - 1] : __ldg(row + (s > 0 ? s - 1 : 1));
            }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int s = c_lo + base + 32 * i + lane;
            if (s < c_hi) st[s] = fmaf(-coef, prev[i], cur[i]);
        }
        __syncwarp();
    }
}
__device__ __forceinline__ void issue_copy(const CopyGeom &g, uint32_t stage_s, uint32_t bar) {
    fence_proxy_async();
    mbar_arrive_expect_tx(bar, g.bytes);  // bytes may be 0: the arrive alone completes the phase
    if (g.bytes) tma_load_1d(stage_s + (uint32_t)(g.dst * 4), g.src, g.bytes, bar);
}

// Everything that locates a task of the spectrum kernel (spec_kernel.cuh); recomputed identically by the prefetch
// and the consume side.
struct Task {
    long long b;
    int t0, Li, s_first;  // first frame, clip length, first padded-coordinate sample of the span
    bool valid0, valid1;
    int span;             // samples staged: n_fft (+ hop when the second frame is valid)
};

template <bool kPair>
__device__ __forceinline__ Task decode_task(const KParams &p, long long b, int q) {
    Task t;
    t.b = b;
    t.Li = p.lengths ? min(__ldg(p.lengths + b), p.L) : p.L;
    const int Ti = p.lengths ? min(frames_of(t.Li, p.n_fft_log, p.hop, p.pad), p.T) : p.T;
    t.t0 = q * p.pair_frames;
    t.valid0 = t.t0 < Ti;
    t.valid1 = kPair && p.pair_frames == 2 && (t.t0 + 1 < Ti);
    t.s_first = t.t0 * p.hop - p.pad;
    t.span = p.n_fft + (t.valid1 ? p.hop : 0);
    return t;
}

// One mel round for one lane, G float4 weight groups, straight-line: all 3G 128-bit loads are issued before the
// 8G FFMAs (four independent accumulation chains), so the shared-memory latency is paid once per round instead of
// once per group.  `w4` points at the lane's first weight group ([group][lane] layout, stride 32 float4).
template <bool kPair, int G>
__device__ __forceinline__ void mel_round(const float4 *w4, const void *tile_at_lo, float &acc0, float &acc1) {
    if constexpr (kPair) {
        // {|X_t[k]|, |X_t+1[k]|, |X_t[k+1]|, |X_t+1[k+1]|}: both frames of a bin advance in one FFMA2
        const float4 *mg = reinterpret_cast<const float4 *>(tile_at_lo);
        float4 w[G], u[G], v[G];
#pragma unroll
        for (int g = 0; g < G; ++g) w[g] = w4[g * 32], u[g] = mg[2 * g], v[g] = mg[2 * g + 1];
        float2 a = make_float2(acc0, acc1), b = make_float2(0.f, 0.f);  // two independent accumulation chains
#pragma unroll
        for (int g = 0; g < G; ++g) {
            a = __ffma2_rn(make_float2(u[g].x, u[g].y), make_float2(w[g].x, w[g].x), a);
            b = __ffma2_rn(make_float2(u[g].z, u[g].w), make_float2(w[g].y, w[g].y), b);
            a = __ffma2_rn(make_float2(v[g].x, v[g].y), make_float2(w[g].z, w[g].z), a);
            b = __ffma2_rn(make_float2(v[g].z, v[g].w), make_float2(w[g].w, w[g].w), b);
        }
        a = __fadd2_rn(a, b);
        acc0 = a.x, acc1 = a.y;
    } else {
        const float4 *mg = reinterpret_cast<const float4 *>(tile_at_lo);
        float4 w[G], u[G];
#pragma unroll
        for (int g = 0; g < G; ++g) w[g] = w4[g * 32], u[g] = mg[g];
        float2 a = make_float2(acc0, 0.f);  // even / odd terms of the single frame in the two halves
#pragma unroll
        for (int g = 0; g < G; ++g) {
            a = __ffma2_rn(make_float2(u[g].x, u[g].y), make_float2(w[g].x, w[g].y), a);
            a = __ffma2_rn(make_float2(u[g].z, u[g].w), make_float2(w[g].z, w[g].w), a);
        }
        acc0 = a.x + a.y;
    }
}

// Runs `groups` (warp-uniform) weight groups as straight-line chunks of 8 / 4 / 2 / 1.
template <bool kPair, int kMaxChunk>
__device__ __forceinline__ void mel_groups(int groups, const float4 *w4, const unsigned char *tile_at_lo, float &acc0,
                                           float &acc1) {
    constexpr int kBytesPerGroup = kPair ? 32 : 16;  // tile bytes one weight group covers
    while (groups >= kMaxChunk) {
        mel_round<kPair, kMaxChunk>(w4, tile_at_lo, acc0, acc1);
        groups -= kMaxChunk, w4 += kMaxChunk * 32, tile_at_lo += kMaxChunk * kBytesPerGroup;
    }
    if constexpr (kMaxChunk > 4) {
        if (groups & 4) {
            mel_round<kPair, 4>(w4, tile_at_lo, acc0, acc1);
            w4 += 4 * 32, tile_at_lo += 4 * kBytesPerGroup;
        }
    }
    if constexpr (kMaxChunk > 2) {
        if (groups & 2) {
            mel_round<kPair, 2>(w4, tile_at_lo, acc0, acc1);
            w4 += 2 * 32, tile_at_lo += 2 * kBytesPerGroup;
        }
    }
    if constexpr (kMaxChunk > 1) {
        if (groups & 1) mel_round<kPair, 1>(w4, tile_at_lo, acc0, acc1);
    }
}

// Transposed read-back between the two radix-32 passes (lane = k1, slot = n2) with the inter-pass twiddle
// W_1024^{n2 k1} applied on the read side.  Both the lane's row of the transpose buffer and the twiddle table
// ([n2 / 2][lane][2]) are read with 128-bit loads — two complex values per LDS — and each batch issues all its
// loads back to back before the multiplies (deep memory-level parallelism; the a[] registers are free here).
template <int kB>
__device__ __forceinline__ void xpose_read_twiddle(float2 *a, const float2 *buf, const float2 *s_tw, int lane) {
    const float4 *row = reinterpret_cast<const float4 *>(buf + lane * kBufStride);
    const float4 *tw4 = reinterpret_cast<const float4 *>(s_tw) + lane;
    static_for<0, 32 / kB>([&](auto h_) {
        constexpr int h = decltype(h_)::value;
        float4 v[kB / 2], t[kB / 2];
#pragma unroll
        for (int i = 0; i < kB / 2; ++i) v[i] = row[h * (kB / 2) + i];
#pragma unroll
        for (int i = 0; i < kB / 2; ++i)
            t[i] = tw4[(h * (kB / 2) + i) * 32];
#pragma unroll
        for (int i = 0; i < kB / 2; ++i) {
            const int j = h * kB + 2 * i;
            a[j] = j > 0 ? cmul(make_float2(v[i].x, v[i].y), make_float2(t[i].x, t[i].y)) : make_float2(v[i].x, v[i].y);
            a[j + 1] = cmul(make_float2(v[i].z, v[i].w), make_float2(t[i].z, t[i].w));
        }
    });
}

// What the consume side of a task needs to know, computed ONCE when the task's samples are requested and carried
// in registers until the task is processed (the prefetch runs one task ahead).
struct Desc {
    int b, t0;        // clip, first frame
    int delta;        // stage shift: sample s of the span sits at stage[s - s_first + delta]
    int Li;           // clip length (p.L unless `lengths`)
    unsigned flags;   // 1: frame t0 exists, 2: frame t0+1 exists (pair mode), 4: part of the span needs the patch
};

// Locate a task and (one elected lane) request its samples: arm the warp's mbarrier and issue ONE bulk copy
// (copy_geom).  Everything is computed uniformly by all lanes in straight-line code, so the scheduler can sink it
// into the shadow of the surrounding FFT arithmetic; only the PTX instructions at the end are predicated.
template <bool kPair>
__device__ __forceinline__ Desc request_task(const KParams &p, int b, int q, int lane, uint32_t stage_s, uint32_t bar) {
    Desc d;
    d.b = b;
    int Li = p.L, Ti = p.T;
    if (p.lengths) {
        Li = min(__ldg(p.lengths + b), p.L);
        Ti = min(frames_of(Li, p.n_fft_log, p.hop, p.pad), p.T);
    }
    d.Li = Li;
    d.t0 = q * p.pair_frames;
    const bool v0 = d.t0 < Ti, v1 = kPair && p.pair_frames == 2 && d.t0 + 1 < Ti;
    const int s_first = d.t0 * p.hop - p.pad;
    const int span = p.n_fft + (v1 ? p.hop : 0);
    const CopyGeom g = copy_geom(p, b, s_first, span, Li);
    d.delta = g.delta;
    d.flags = (v0 ? 1u : 0u) | (v1 ? 2u : 0u) | (g.patch ? 4u : 0u);
    if (v0 && lane == 0) issue_copy(g, stage_s, bar);
    return d;
}

template <bool kPre>
__device__ __forceinline__ void patch_halo_smem(const KParams &p, const Desc &d, float *stage, int lane) {
    const int s_first = d.t0 * p.hop - p.pad;
    const int span = p.n_fft + ((d.flags & 2u) ? p.hop : 0);
    const CopyGeom g = copy_geom(p, d.b, s_first, span, d.Li);
    if constexpr (kPre) {
        if (g.c_hi > g.c_lo)  // warp-uniform
            preemphasize_stage(p.wav + (long long)d.b * p.row_stride, stage + g.delta - s_first, g.c_lo, g.c_hi, p.preemph, lane);
    }
    if (g.patch) patch_stage(p, d.b, s_first, span, d.Li, g, stage, lane, kPre ? p.preemph : 0.f);
}

// Pair mode, stage -> registers with the window applied: a[j] = {x_t[32 j + lane], x_t+1[32 j + lane]} * w[32 j + lane].
// hann_cs = (0.25 cos phi, 0.25 sin phi), phi = 2 pi lane / 1024 (only read when p.hann_full).
__device__ __forceinline__ void load_windowed_pair(float2 *a, const float *x0, const float *s_win, float2 hann_cs,
                                                   const KParams &p, int lane, bool valid1) {
    // keep the generated window a per-task computation: hoisted out of the task loop it would be 32 live registers
    // (in practice 32 local-memory reloads per task, i.e. the table loads this path exists to avoid)
    asm volatile("" : "+f"(hann_cs.x), "+f"(hann_cs.y));
    if (valid1 && p.hop == 256) {
        // hop = 8 * 32: element j of frame t+1 IS element j+8 of frame t in the stage -> 40 distinct
        // shared-memory loads per lane for the two frames instead of 64 (the kernel is LSU-bound)
        float raw[40];
#pragma unroll
        for (int j = 0; j < 40; ++j) raw[j] = x0[32 * j];
        if (p.hann_full) {
            // Full-length periodic Hann generated in registers: 0.5 w[32 j + lane] = 0.25 - 0.25 cos(theta_j + phi),
            // theta_j = 2 pi j / 32 (compile-time cos / sin), so no window-table loads on the common path.
            static_for<0, 16>([&](auto j_) {
                constexpr int j = decltype(j_)::value;
                // t = 0.25 cos(theta_j + phi); TwConst::s32 holds -sin
                const float t = fmaf(TwConst::s32[j], hann_cs.y, TwConst::c32[j] * hann_cs.x);
                const float w0 = 0.25f - t, w1 = 0.25f + t;  // slots j and j + 16 (theta + pi)
                a[j] = __fmul2_rn(make_float2(raw[j], raw[j + 8]), make_float2(w0, w0));
                a[j + 16] = __fmul2_rn(make_float2(raw[j + 16], raw[j + 24]), make_float2(w1, w1));
            });
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float w = s_win[32 * j + lane];
                a[j] = __fmul2_rn(make_float2(raw[j], raw[j + 8]), make_float2(w, w));
            }
        }
    } else if (valid1) {
        const float *x1 = x0 + p.hop;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float w = s_win[32 * j + lane];
            a[j] = __fmul2_rn(make_float2(x0[32 * j], x1[32 * j]), make_float2(w, w));  // one FMUL2
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            a[j].x = x0[32 * j] * s_win[32 * j + lane];
            a[j].y = 0.f;
        }
    }
}

// kPower: 1 magnitude, 2 power.  kTop: 32-bin groups of the spectrum that are separated in pair mode — 16 = all
// 513 bins, 12 = bins 0..383 only (plans whose filterbank ends below bin 384, e.g. fmax 8000 Hz at 22050 Hz; the
// unused FFT outputs are dead code for the compiler).  16 warps per CTA (128 registers per thread).
// kPre: the fused pre-emphasis prologue (io.preemphasis != 0) is a separate instantiation — as a run-time branch of the
// one body it cost the split-mode kernel 72 bytes of spills and 10 % of its speed at C4 (55 us instead of 49 us).
template <bool kPair, int kPower, int kTop, bool kPre = false>
__global__ void __launch_bounds__(kMaxWarps * 32, 1) logmel_kernel(const KParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_warps = blockDim.x >> 5;
#if defined(B200MEL_PHASE_TIMING) || defined(B200MEL_SPAN_TIMING)
    if (p.dbg && threadIdx.x == 0) {
        unsigned long long t_;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        atomicMin(reinterpret_cast<unsigned long long *>(p.dbg) + 14, t_);
    }
#endif

    float2 *s_tw = reinterpret_cast<float2 *>(smem_raw);
    float *s_win = reinterpret_cast<float *>(smem_raw + p.off_window);
    const MelEntry *s_ent = reinterpret_cast<const MelEntry *>(smem_raw + p.off_entries);
    const float *s_melw = reinterpret_cast<const float *>(smem_raw + p.off_melw);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + p.off_bar);
    unsigned char *region = smem_raw + p.off_regions + warp * p.region_bytes;
    float2 *buf = reinterpret_cast<float2 *>(region);          // transpose buffer
    float2 *tile2 = reinterpret_cast<float2 *>(region);        // pair-mode magnitude tile
    float *tile1 = reinterpret_cast<float *>(region);          // split-mode magnitude tile
    float *stage = reinterpret_cast<float *>(region + kStageOff);
    const uint32_t stage_s = smem_u32(stage);
    const uint32_t bar = smem_u32(s_bar + warp);
    uint32_t parity = 0;

    // task = cb * tasks_per_clip + cq, advanced by the grid-wide warp stride without any division.  The warps of
    // one CTA take tasks gridDim.x apart (task = warp * gridDim.x + blockIdx.x), so the partial last round of a
    // launch spreads over ALL SMs instead of filling the first CTAs' 16 warps and leaving the others idle.
    const long long stride = (long long)gridDim.x * n_warps;
    long long task = (long long)warp * gridDim.x + blockIdx.x;
    int cb = (int)(task / p.tasks_per_clip);
    int cq = (int)(task - (long long)cb * p.tasks_per_clip);

    auto request = [&](int b, int q) -> Desc { return request_task<kPair>(p, b, q, lane, stage_s, bar); };

    // Prologue, ordered for programmatic dependent launch (PDL): everything that does not touch caller memory —
    // mbarrier init and the fetch of the plan-owned tables — is started BEFORE griddepcontrol.wait, i.e. it overlaps
    // the tail of whatever kernel precedes this one in the stream.  The tables come in as four bulk copies (TMA)
    // signalled on one mbarrier, so no thread carries them through registers and their L2 latency runs concurrently
    // with the first sample fetch.  Caller memory (wav, outputs) is only touched after the wait;
    // griddepcontrol.launch_dependents then lets the next launch start its own prologue the same way.
    const uint32_t tbar = smem_u32(s_bar + kMaxWarps);
    if (lane == 0) mbar_init(bar, 1);
    if (threadIdx.x == 0) {
        mbar_init(tbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t b_tw = 32 * 32 * 8, b_win = (uint32_t)p.n_fft * 4u;
        const uint32_t b_ent = (uint32_t)p.mel_rounds * 32u * (uint32_t)sizeof(MelEntry), b_w = (uint32_t)p.mel_w_len * 4u;
        mbar_arrive_expect_tx(tbar, b_tw + b_win + b_ent + b_w);
        tma_load_1d(smem_u32(s_tw), p.tw, b_tw, tbar);
        tma_load_1d(smem_u32(s_win), p.window, b_win, tbar);
        tma_load_1d(smem_u32(smem_raw + p.off_entries), p.mel_entries, b_ent, tbar);
        tma_load_1d(smem_u32(smem_raw + p.off_melw), p.mel_w, b_w, tbar);
    } else if (lane == 0) {
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // Hint only: pull the first task's span towards L2 while the previous kernel may still be running.  L2 is the
    // point of coherence, so a producer that writes wav afterwards simply updates the line; the real (TMA) read
    // happens after griddepcontrol.wait.
    if (lane == 0 && task < p.n_tasks) {
        const int s0 = max(cq * p.pair_frames * p.hop - p.pad, 0);
        const int s1 = min(s0 + p.n_fft + p.hop, p.L);
        const float *row = p.wav + (long long)cb * p.row_stride;
        const uintptr_t a16 = reinterpret_cast<uintptr_t>(row + s0) & ~(uintptr_t)15;
        const uintptr_t e16 = reinterpret_cast<uintptr_t>(row + s1) & ~(uintptr_t)15;
        if (e16 > a16)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(reinterpret_cast<const void *>(a16))),
                         "r"((uint32_t)(e16 - a16))
                         : "memory");
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // every warp gets its first task's samples moving (its own mbarrier, initialised by its own lane 0)
    Desc cur;
    cur.b = cur.t0 = cur.delta = cur.Li = 0;
    cur.flags = 0;
    __syncwarp();
    if (task < p.n_tasks) cur = request(cb, cq);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __syncthreads();      // the only block-wide barrier: the table mbarrier is initialised for everyone
    mbar_wait(tbar, 0);   // tables have landed (async-proxy writes are visible to the waiting threads)

    float2 wl = make_float2(1.f, 0.f);
    if (!kPair) wl = __ldg(p.tw_post + lane);
    // Full-length periodic Hann (win_length == n_fft) is generated in registers instead of read from the table:
    // 0.5 w[32 j + lane] = 0.25 - 0.25 cos(2 pi j / 32 + phi), phi = 2 pi lane / 1024 — the angle-addition form
    // with the compile-time cos/sin of 2 pi j / 32 and the lane's own (0.25 cos phi, 0.25 sin phi).
    float2 hann_cs = make_float2(0.f, 0.f);
    if (kPair && p.hann_full) {
        float sn, cs;
        sincospif((float)lane * (1.0f / 512.0f), &sn, &cs);
        hann_cs = make_float2(0.25f * cs, 0.25f * sn);
    }

#ifdef B200MEL_PHASE_TIMING
    unsigned phase_acc_[14];
#pragma unroll
    for (int i_ = 0; i_ < 14; ++i_) phase_acc_[i_] = 0;
    unsigned tmark_ = (unsigned)clock();
#endif
    PHASE_MARK(0);  // prologue
#if defined(B200MEL_PHASE_TIMING) || defined(B200MEL_SPAN_TIMING)
    if (p.dbg && lane == 0) {
        unsigned long long t_;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        atomicMin(reinterpret_cast<unsigned long long *>(p.dbg) + 16, t_);   // first warp past the prologue
        atomicMax(reinterpret_cast<unsigned long long *>(p.dbg) + 17, t_);   // last warp past the prologue
    }
#endif

    auto patch_halo = [&](const Desc &d) { patch_halo_smem<kPre>(p, d, stage, lane); };

    for (; task < p.n_tasks; task += stride) {
        const Desc d = cur;
        // next task of this warp
        cb += p.stride_b;
        cq += p.stride_q;
        if (cq >= p.tasks_per_clip) cq -= p.tasks_per_clip, ++cb;
        const bool valid0 = d.flags & 1u, valid1 = d.flags & 2u;
        float2 a[32];
        if (p.out_fmask && lane < p.pair_frames && d.t0 + lane < p.T)  // frame mask of this task's frames (models/transforms.py:397-416)
            p.out_fmask[(long long)d.b * p.T + d.t0 + lane] = ((d.t0 + lane) * p.hop - p.win_half < d.Li) ? 1.f : 0.f;

        if (valid0) {
            PHASE_MARK(1);  // decode
            // -------------------------------------------------------------- stage -> registers, windowed
            mbar_wait(bar, parity);
            parity ^= 1;
            if ((d.flags & 4u) || kPre) patch_halo(d);
            const float *x0 = stage + d.delta + lane;
            PHASE_MARK(2);  // wait for the TMA stage
            if constexpr (kPair) {
                load_windowed_pair(a, x0, s_win, hann_cs, p, lane, valid1);
            } else {
                const float2 *w2 = reinterpret_cast<const float2 *>(s_win);
                // The even / odd sample pair of a lane is ONE aligned 64-bit load whenever the stage shift is even (always
                // for 8-byte aligned rows of even length): 32 conflict-free LDS.64 instead of 64 two-way conflicting
                // LDS.32 per task (measured at C4: 49.9 -> 48.5 us).
                if ((d.delta & 1) == 0) {
                    const float2 *x2 = reinterpret_cast<const float2 *>(stage + d.delta) + lane;
#pragma unroll
                    for (int j = 0; j < 32; ++j) a[j] = __fmul2_rn(x2[32 * j], w2[32 * j + lane]);
                } else
#pragma unroll
                for (int j = 0; j < 32; ++j)  // samples 64 j + 2 lane, 64 j + 2 lane + 1
                    a[j] = __fmul2_rn(make_float2(x0[64 * j + lane], x0[64 * j + lane + 1]), w2[32 * j + lane]);
            }

            // -------------------------------------------------------------- 1024-point complex FFT
            PHASE_MARK(3);  // stage -> registers (windowed)
            fft32(a);  // pass 1: lane = n2, FFT over n1 -> Y[k1] at a[pos(k1)]
            PHASE_MARK(4);  // pass 1
            __syncwarp();  // every lane has consumed the stage before the transpose buffer overwrites it
            static_for<0, 32>([&](auto k1_) {
                constexpr int k1 = decltype(k1_)::value;
                buf[k1 * kBufStride + lane] = a[fft32_pos(k1)];
            });
            __syncwarp();
            xpose_read_twiddle<B200MEL_XPOSE_BATCH>(a, buf, s_tw, lane);
            __syncwarp();  // transpose buffer is dead: magnitude tile and next stage may reuse it
            PHASE_MARK(5);  // transpose + twiddle
        }

        // ------------------------------------------------------------------ request the next task's samples
        cur.flags = 0;
        if (task + stride < p.n_tasks) cur = request(cb, cq);

        int2 e_first = make_int2(0, -1);  // {lo, m} of the first mel round, fetched early so its latency hides under pass 2
        if (valid0) {
            e_first = reinterpret_cast<const int2 *>(s_ent)[lane];
            PHASE_MARK(6);  // prefetch issue
            fft32(a);  // pass 2: lane = k1, FFT over n2 -> Z[k1 + 32 k2] at a[pos(k2)]
            PHASE_MARK(7);  // pass 2

            // -------------------------------------------------------------- real-input separation + magnitudes
            const int partner = (32 - lane) & 31;
            constexpr int kGroups = kPair ? kTop : 16;
            static_for<0, kGroups>([&](auto k2_) {
                constexpr int k2 = decltype(k2_)::value;
                const float2 A = a[fft32_pos(k2)];
                // value my reader needs: lane 0 is read by itself and wants Z[32*((32-k2)&31)];
                // lane s != 0 is read by lane 32-s, which wants my slot 31-k2.
                const float2 g0 = a[fft32_pos((32 - k2) & 31)];
                const float2 g1 = a[fft32_pos(31 - k2)];
                float2 Bv;
                Bv.x = __shfl_sync(0xffffffffu, lane == 0 ? g0.x : g1.x, partner);
                Bv.y = __shfl_sync(0xffffffffu, lane == 0 ? g0.y : g1.y, partner);
                const int k = lane + 32 * k2;
                // E = A + conj(B), D = A - conj(B) (the 1/2 is folded into the window): frame t is E, frame t+1 is
                // O = -i D, and |O| = |D|, so the rotation is never formed.  conj() is an operand sign pattern.
                if constexpr (kPair) {
                    tile2[k] = pair_magnitudes<kPower>(A, Bv, p.mag_eps);
                } else {
                    const float2 Bc = make_float2(Bv.x, -Bv.y);
                    const float2 E = __fadd2_rn(A, Bc);
                    const float2 D = __fadd2_rn(A, make_float2(-Bc.x, -Bc.y));
                    // X[k] = E + W_2048^k O,  X[1024-k] = conj(E - W_2048^k O),  O = -i D,  W_2048^k = wl * W_64^{k2}
                    // -> P = D * (-i wl W_64^{k2})
                    constexpr float w64c = TwConst::c64[k2], w64s = TwConst::s64[k2];
                    const float2 Wk = cmul(wl, make_float2(w64c, w64s));
                    const float2 P = cmul(D, make_float2(Wk.y, -Wk.x));
                    const float2 m = magnitude2<kPower>(__fadd2_rn(E, P), __fadd2_rn(E, make_float2(-P.x, -P.y)), p.mag_eps);
                    tile1[k] = m.x;
                    tile1[1024 - k] = m.y;
                }
            });
            if constexpr (kGroups == 16) {
                if (lane == 0) {  // bin 512 (k1 = 0, k2 = 16) is its own partner
                    const float2 A = a[fft32_pos(16)];
                    if constexpr (kPair) {
                        tile2[512] = magnitude2<kPower>(make_float2(2.f * A.x, 0.f), make_float2(2.f * A.y, 0.f), p.mag_eps);
                    } else {  // E = 2 Re A, O = 2 Im A, W_2048^512 = -i  ->  X[512] = 2 (Re A - i Im A)
                        tile1[512] = magnitude2<kPower>(make_float2(2.f * A.x, -2.f * A.y), make_float2(0.f, 0.f), p.mag_eps).x;
                    }
                }
                if (lane < 7) {  // zero the padded tail the float4 weight groups may touch
                    if constexpr (kPair) tile2[513 + lane] = make_float2(0.f, 0.f);
                    else tile1[1025 + lane] = 0.f;
                }
            }  // kTop < 16: the plan keeps every read window below bin 32 kTop, all of which were just written
            __syncwarp();
        }

        // ---------------------------------------------------------------------- frames past the clip's end
        if (!valid0 || (kPair && p.pair_frames == 2 && !valid1)) {
            // only with `lengths` (or the odd last frame of a pair): zero-fill, as pad_collate_fn zero-pads
            // per-item features (data/dataset.py:230-250).
            const int tz0 = valid0 ? d.t0 + 1 : d.t0;
            const int tz1 = d.t0 + p.pair_frames - 1;
            for (int tt = tz0; tt <= tz1 && tt < p.T; ++tt)
                for (int m = lane; m < p.n_mels; m += 32) p.out_mel[((long long)d.b * p.n_mels + m) * (long long)p.T + tt] = 0.f;
        }
        PHASE_MARK(8);  // separation + magnitudes

        // ---------------------------------------------------------------------- banded mel + log epilogue
        if (valid0) {
            float *orow = p.out_mel + (long long)d.b * p.n_mels * (long long)p.T + d.t0;
            const int2 *ent2 = reinterpret_cast<const int2 *>(s_ent) + lane;
            const float4 *wbase = reinterpret_cast<const float4 *>(s_melw) + lane;
            const unsigned char *tile_bytes = region;
            int2 e = e_first;  // {lo, m}; the next round's entry is fetched while this one computes
#pragma unroll 1
            for (int r = 0; r < p.mel_rounds; ++r) {
                const int2 ce = e;
                if (r + 1 < p.mel_rounds) e = ent2[(r + 1) * 32];
                float acc0 = 0.f, acc1 = 0.f;
                PHASE_MARK(9);  // round setup
                mel_groups<kPair, B200MEL_MEL_CHUNK>(p.round_groups[r], wbase + p.round_wbase[r], tile_bytes + ce.x * (kPair ? 8 : 4), acc0, acc1);
                PHASE_MARK(10);  // mel FMAs
                const float y0 = epilogue(acc0, p), y1 = epilogue(acc1, p);
                PHASE_MARK(11);  // log epilogue
                if (ce.y >= 0) {
                    float *o = orow + ce.y * p.T;  // n_mels * T < 2^31 is checked by the host (a 64-bit product here spills in split mode)
                    o[0] = y0;
                    if (kPair && valid1) o[1] = y1;
                }
                PHASE_MARK(12);  // stores
            }
            __syncwarp();  // tile reads done before the next task's transpose overwrites the region
            PHASE_MARK(13);  // final syncwarp
        }
    }
#ifdef B200MEL_PHASE_TIMING
    if (p.dbg && lane == 0) {
#pragma unroll
        for (int i_ = 0; i_ < 14; ++i_)
            atomicAdd(reinterpret_cast<unsigned long long *>(p.dbg) + i_, (unsigned long long)phase_acc_[i_]);
    }
#endif
#if defined(B200MEL_PHASE_TIMING) || defined(B200MEL_SPAN_TIMING)
    if (p.dbg && lane == 0) {
        unsigned long long t_;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        atomicMax(reinterpret_cast<unsigned long long *>(p.dbg) + 15, t_);
    }
#endif
}

}  // namespace b200mel
