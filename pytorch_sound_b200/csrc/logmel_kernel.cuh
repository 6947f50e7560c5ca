// logmel_kernel.cuh — the fused STFT -> |.| -> mel -> log kernel (sm_100a), v2.
//
// Persistent grid: one CTA per SM, up to 16 warps per CTA, every warp an independent pipeline that
// walks the task list with stride (#warps in the grid).  A task is one 1024-point complex FFT held in
// registers (32 complex values per lane, two radix-32 passes, one shared-memory transpose):
//   kPair  (n_fft = 1024): frames (2q, 2q+1) of a clip packed as re/im, separated by conjugate symmetry;
//   !kPair (n_fft = 2048): frame q packed even/odd, finished by the real-input split pass.
//
// Per-warp shared-memory region (bytes):
//   [0, 4160)              magnitude tile   (pair: float2[520] = {|X_t[k]|, |X_t+1[k]|}; split: float[1032])
//   [4224, 4224 + stage)   sample stage     (filled by ONE cp.async.bulk = TMA 1-D copy per task, mbarrier-signalled)
//   [0, 8448)              transpose buffer (float2[32][33]) — overlaps both, live only between the two
//                          radix-32 passes, i.e. after the stage was consumed and before the next TMA is issued.
// The TMA for task i+1 is issued right after the transpose of task i, so the copy lands while the warp
// does its second FFT pass, the separation, the mel contraction and the epilogue of task i.
//
// CTA-shared tables (loaded once per CTA): inter-pass twiddles, window, banded mel filterbank in a
// lane-balanced schedule (rows sorted by length, 32 rows per round, weights padded to float4 groups).
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

constexpr int kMaxWarps = 24;  // upper bound over all kernel variants (mbarrier array size)
constexpr int kBufStride = 33;                           // float2 per transposed row (+1 pad)
constexpr int kXposeBytes = 32 * kBufStride * 8;          // 8448
constexpr int kTileBytes = 4160;                          // magnitude tile
constexpr int kStageOff = 4224;                           // stage offset inside the warp region (128B aligned)
constexpr int kPairTileLen = 520;                         // float2 entries (513 used, tail zeroed)
constexpr int kSplitTileLen = 1032;                       // float entries (1025 used, tail zeroed)

constexpr int kMaxMelRounds = 8;  // 32 rows per round -> up to 256 mel rows

struct MelEntry {  // one filterbank row as seen by one lane in one round
    int lo;        // first spectrum bin of the row's read window (16-byte aligned in the tile, slid for bank spread)
    int groups;    // float4 weight groups the row itself needs (informational; the loop runs the round's count)
    int woff;      // unused (weights are addressed [round base + group][lane])
    int m;         // mel row index, -1 = idle lane
};

struct KParams {
    const float *wav;
    long long row_stride;
    long long B;
    int L;
    const int *lengths;
    int T, hop, pad, n_fft;
    int pair_frames;  // frames per task: 2 (pair mode) or 1 (split mode, or pair mode with hop > n_fft)
    // global copies of the CTA tables
    const float *window;    // [n_fft], periodic Hann centre-padded, pre-scaled by 0.5
    const float2 *tw;       // [32][32]  tw[k1*32 + lane] = exp(-2 pi i k1 lane / 1024)
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
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

__device__ __forceinline__ int frames_of(int Li, int n_fft, int hop, int pad) {
    int span = Li + 2 * pad - n_fft;
    return span < 0 ? 0 : span / hop + 1;
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

template <int kPower>
__device__ __forceinline__ float magnitude(float re, float im, float eps) {
    const float sq = fmaf(re, re, im * im);
    if constexpr (kPower == 2) return sq;
    return sqrt_approx(sq + eps);
}

// Everything that locates a task; recomputed identically by the prefetch and the consume side.
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
    const int Ti = p.lengths ? min(frames_of(t.Li, p.n_fft, p.hop, p.pad), p.T) : p.T;
    t.t0 = q * p.pair_frames;
    t.valid0 = t.t0 < Ti;
    t.valid1 = kPair && p.pair_frames == 2 && (t.t0 + 1 < Ti);
    t.s_first = t.t0 * p.hop - p.pad;
    t.span = p.n_fft + (t.valid1 ? p.hop : 0);
    return t;
}

// Stage index of padded-coordinate sample position `s` is (s - s_first + delta): delta in 0..3 makes the first
// in-range sample land 16-byte-congruent with its global address, as the bulk copy requires.
__device__ __forceinline__ int stage_delta(const float *row, const Task &t) {
    const int p_lo = max(t.s_first, 0);
    return (int)(((reinterpret_cast<uintptr_t>(row + p_lo) >> 2) - (uintptr_t)(p_lo - t.s_first)) & 3);
}

// One elected lane: arm the warp's mbarrier and issue the bulk copy of the in-range part of the span.
// The copy is widened to 16-byte boundaries on both sides (the extra <= 3 floats on each side are never read
// as samples: interior tasks ignore them, edge tasks overwrite the halo after the copy has landed).
template <bool kPair>
__device__ __forceinline__ void issue_stage(const KParams &p, const Task &t, float *stage, uint32_t bar) {
    const float *row = p.wav + t.b * p.row_stride;
    const int p_lo = max(t.s_first, 0), p_hi = min(t.s_first + t.span, t.Li);
    const uintptr_t a = reinterpret_cast<uintptr_t>(row + p_lo), e = reinterpret_cast<uintptr_t>(row + p_hi);
    const uintptr_t a16 = a & ~(uintptr_t)15, e16 = (e + 15) & ~(uintptr_t)15;
    const int delta = stage_delta(row, t);
    const int idx_lo = p_lo - t.s_first + delta;              // stage index of sample p_lo
    float *dst = stage + (idx_lo - (int)((a >> 2) & 3));      // multiple of 4 floats by construction
    const uint32_t bytes = (uint32_t)(e16 - a16);
    fence_proxy_async();
    mbar_arrive_expect_tx(bar, bytes);
    tma_load_1d(smem_u32(dst), reinterpret_cast<const void *>(a16), bytes, bar);
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

// kPower: 1 magnitude, 2 power.  kWarps: warps per CTA the variant is compiled for (register budget =
// 65536 / (32 kWarps)).
template <bool kPair, int kPower, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) logmel_kernel(const KParams p) {
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
    const uint32_t bar = smem_u32(s_bar + warp);
    uint32_t parity = 0;

    // task = cb * tasks_per_clip + cq, advanced by the grid-wide warp stride without any division
    const long long stride = (long long)gridDim.x * n_warps;
    long long task = (long long)blockIdx.x * n_warps + warp;
    long long cb = task / p.tasks_per_clip;
    int cq = (int)(task - cb * p.tasks_per_clip);

    // Prologue, ordered for programmatic dependent launch (PDL): everything that does not touch caller memory —
    // mbarrier init and the loads of the plan-owned tables — runs BEFORE griddepcontrol.wait, i.e. it overlaps the
    // tail of whatever kernel precedes this one in the stream.  Caller memory (wav, outputs) is only touched after
    // the wait; griddepcontrol.launch_dependents then lets the next launch start its own prologue the same way.
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    int4 t_tw = make_int4(0, 0, 0, 0), t_win = t_tw, t_ent = t_tw;
    const int tid = threadIdx.x;
    if (blockDim.x >= 512) {
        t_tw = __ldg(reinterpret_cast<const int4 *>(p.tw) + tid);
        if (tid < p.n_fft / 4) t_win = __ldg(reinterpret_cast<const int4 *>(p.window) + tid);
        if (tid < p.mel_rounds * 32) t_ent = __ldg(reinterpret_cast<const int4 *>(p.mel_entries) + tid);
    }
    // Hint only: pull the first task's span towards L2 while the previous kernel may still be running.  L2 is the
    // point of coherence, so a producer that writes wav afterwards simply updates the line; the real (TMA) read
    // happens after griddepcontrol.wait.
    if (lane == 0 && task < p.n_tasks) {
        const Task t = decode_task<kPair>(p, cb, cq);
        if (t.valid0) {
            const float *row = p.wav + t.b * p.row_stride;
            const int p_lo = max(t.s_first, 0), p_hi = min(t.s_first + t.span, t.Li);
            const uintptr_t a16 = reinterpret_cast<uintptr_t>(row + p_lo) & ~(uintptr_t)15;
            const uintptr_t e16 = (reinterpret_cast<uintptr_t>(row + p_hi) + 15) & ~(uintptr_t)15;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(reinterpret_cast<const void *>(a16))),
                         "r"((uint32_t)(e16 - a16))
                         : "memory");
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // every warp gets its first task's samples moving before the tables are stored
    if (lane == 0 && task < p.n_tasks) {
        const Task t = decode_task<kPair>(p, cb, cq);
        if (t.valid0) issue_stage<kPair>(p, t, stage, bar);
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __syncwarp();
    {
        const int4 *g;
        int4 *s;
        if (blockDim.x >= 512) {
            reinterpret_cast<int4 *>(s_tw)[tid] = t_tw;
            if (tid < p.n_fft / 4) reinterpret_cast<int4 *>(s_win)[tid] = t_win;
            if (tid < p.mel_rounds * 32) reinterpret_cast<int4 *>(smem_raw + p.off_entries)[tid] = t_ent;
            g = reinterpret_cast<const int4 *>(p.window);
            s = reinterpret_cast<int4 *>(s_win);
            for (int i = tid + 512; i < p.n_fft / 4; i += blockDim.x) s[i] = __ldg(g + i);  // n_fft > 2048 only
        } else {
            g = reinterpret_cast<const int4 *>(p.tw);
            s = reinterpret_cast<int4 *>(s_tw);
            for (int i = tid; i < 32 * 32 * 8 / 16; i += blockDim.x) s[i] = __ldg(g + i);
            g = reinterpret_cast<const int4 *>(p.window);
            s = reinterpret_cast<int4 *>(s_win);
            for (int i = tid; i < p.n_fft / 4; i += blockDim.x) s[i] = __ldg(g + i);
            g = reinterpret_cast<const int4 *>(p.mel_entries);
            s = reinterpret_cast<int4 *>(smem_raw + p.off_entries);
            for (int i = tid; i < p.mel_rounds * 32; i += blockDim.x) s[i] = __ldg(g + i);
        }
        g = reinterpret_cast<const int4 *>(p.mel_w);
        s = reinterpret_cast<int4 *>(smem_raw + p.off_melw);
        for (int i = tid; i < p.mel_w_len / 4; i += blockDim.x) s[i] = __ldg(g + i);
    }
    __syncthreads();  // the only block-wide barrier; warps are independent from here on

    float2 wl = make_float2(1.f, 0.f);
    if (!kPair) wl = __ldg(p.tw_post + lane);

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

    // Wait for a task's staged samples and patch its reflected halo (edge tasks only: the out-of-range part of the
    // span is overwritten after the bulk copy has landed).  Returns the stage shift delta.
    auto acquire_stage = [&](const Task &tk) -> int {
        const float *row = p.wav + tk.b * p.row_stride;
        const int delta = stage_delta(row, tk);
        mbar_wait(bar, parity);
        parity ^= 1;
        if (tk.s_first < 0 || tk.s_first + tk.span > tk.Li) {
            for (int i = lane; i < tk.span; i += 32) {
                const int s = tk.s_first + i;
                if (s < 0 || s >= tk.Li) stage[i + delta] = __ldg(row + reflect_index(s, tk.Li));
            }
            __syncwarp();
        }
        return delta;
    };

    for (; task < p.n_tasks; task += stride) {
        const Task t = decode_task<kPair>(p, cb, cq);
        // next task of this warp
        cb += p.stride_b;
        cq += p.stride_q;
        if (cq >= p.tasks_per_clip) cq -= p.tasks_per_clip, ++cb;
        const long long b = t.b;
        const int t0 = t.t0;
        float2 a[32];

        if (t.valid0) {
            PHASE_MARK(1);  // decode
            // -------------------------------------------------------------- stage -> registers, windowed
            const float *x0 = stage + acquire_stage(t) + lane;
            PHASE_MARK(2);  // wait for the TMA stage
            if constexpr (kPair) {
                if (t.valid1 && p.hop == 256) {
                    // hop = 8 * 32: element j of frame t+1 IS element j+8 of frame t in the stage -> 40 distinct
                    // shared-memory loads per lane for the two frames instead of 64 (the kernel is LSU-bound)
                    float raw[40];
#pragma unroll
                    for (int j = 0; j < 40; ++j) raw[j] = x0[32 * j];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float w = s_win[32 * j + lane];
                        a[j] = __fmul2_rn(make_float2(raw[j], raw[j + 8]), make_float2(w, w));
                    }
                } else if (t.valid1) {
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
            } else {
                const float2 *w2 = reinterpret_cast<const float2 *>(s_win);
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
            // read back transposed (lane = k1, slot = n2) and apply the inter-pass twiddle W_1024^{n2 k1} on the
            // read side: the a[] registers are free here, so each batch issues its loads back to back (deep
            // memory-level parallelism) instead of serialising load -> multiply -> store per element.
            {
                constexpr int kB = kWarps > 16 ? 8 : 16;  // loads in flight per batch (register budget of the variant)
                static_for<0, 32 / kB>([&](auto h_) {
                    constexpr int h = decltype(h_)::value;
                    float2 tw[kB];
#pragma unroll
                    for (int i = 0; i < kB; ++i) a[h * kB + i] = buf[lane * kBufStride + h * kB + i];
#pragma unroll
                    for (int i = 0; i < kB; ++i) tw[i] = s_tw[(h * kB + i) * 32 + lane];
#pragma unroll
                    for (int i = 0; i < kB; ++i)
                        if (h * kB + i > 0) a[h * kB + i] = cmul(a[h * kB + i], tw[i]);
                });
            }
            __syncwarp();  // transpose buffer is dead: magnitude tile and next stage may reuse it
            PHASE_MARK(5);  // transpose + twiddle
        }

        // ------------------------------------------------------------------ prefetch the next task's samples
        Task nxt;
        nxt.valid0 = false;
        if (task + stride < p.n_tasks) {
            nxt = decode_task<kPair>(p, cb, cq);
            if (nxt.valid0 && lane == 0) issue_stage<kPair>(p, nxt, stage, bar);
        }

        if (t.valid0) {
            PHASE_MARK(6);  // prefetch issue
            fft32(a);  // pass 2: lane = k1, FFT over n2 -> Z[k1 + 32 k2] at a[pos(k2)]
            PHASE_MARK(7);  // pass 2

            // -------------------------------------------------------------- real-input separation + magnitudes
            const int partner = (32 - lane) & 31;
            static_for<0, 16>([&](auto k2_) {
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
                const float2 Bc = make_float2(Bv.x, -Bv.y);
                const float2 E = __fadd2_rn(A, Bc);
                const float2 D = __fadd2_rn(A, make_float2(-Bc.x, -Bc.y));
                if constexpr (kPair) {
                    tile2[k] = magnitude2<kPower>(E, D, p.mag_eps);
                } else {
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
            __syncwarp();
        }

        // ---------------------------------------------------------------------- frames past the clip's end
        if (!t.valid0 || (kPair && p.pair_frames == 2 && !t.valid1)) {
            // only with `lengths` (or the odd last frame of a pair): zero-fill, as pad_collate_fn zero-pads
            // per-item features (data/dataset.py:230-250).
            const int tz0 = t.valid0 ? t0 + 1 : t0;
            const int tz1 = t0 + p.pair_frames - 1;
            for (int tt = tz0; tt <= tz1 && tt < p.T; ++tt)
                for (int m = lane; m < p.n_mels; m += 32) p.out_mel[(b * p.n_mels + m) * (long long)p.T + tt] = 0.f;
        }
        PHASE_MARK(8);  // separation + magnitudes

        // ---------------------------------------------------------------------- banded mel + log epilogue
        if (t.valid0) {
            float *orow = p.out_mel + b * p.n_mels * (long long)p.T + t0;
            const int4 *ent4 = reinterpret_cast<const int4 *>(s_ent) + lane;
            const float4 *wbase = reinterpret_cast<const float4 *>(s_melw) + lane;
            const unsigned char *tile_bytes = region;
            int4 e = ent4[0];  // {lo, groups, woff, m}; the next round's entry is fetched while this one computes
#pragma unroll 1
            for (int r = 0; r < p.mel_rounds; ++r) {
                const int4 cur = e;
                if (r + 1 < p.mel_rounds) e = ent4[(r + 1) * 32];
                float acc0 = 0.f, acc1 = 0.f;
                PHASE_MARK(9);  // round setup
                mel_groups<kPair, (kWarps > 16 ? 4 : 8)>(p.round_groups[r], wbase + p.round_wbase[r],
                                                                     tile_bytes + cur.x * (kPair ? 8 : 4), acc0, acc1);
                PHASE_MARK(10);  // mel FMAs
                const float y0 = epilogue(acc0, p), y1 = epilogue(acc1, p);
                PHASE_MARK(11);  // log epilogue
                if (cur.w >= 0) {
                    float *o = orow + cur.w * p.T;
                    o[0] = y0;
                    if (kPair && t.valid1) o[1] = y1;
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
