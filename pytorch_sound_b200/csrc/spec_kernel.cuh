// spec_kernel.cuh — spectrum-output variant of the fused STFT kernel (sm_100a): STFT.transform (mag, phase),
// STFTTorchAudio.forward (re, im) and magnitude-only.  Replaces models/transforms.py:53-69 and :297-311.
//
// These operators write (B, n_fft/2+1, T) tensors — 6.4x (one array) to 12.8x (two arrays) the bytes of the
// mel output — so they are bound by how well the stores coalesce, not by the FFT.  A warp only ever holds two
// frames of a bin (8 contiguous bytes of an output row); the 8 warps of a CTA therefore work in lock step on 8
// CONSECUTIVE tasks (16 consecutive frames of a clip in pair mode) and pool their spectra in a CTA-wide shared
// tile [bin][frame column], which the whole CTA then writes out row-wise: 16 consecutive threads store 64
// contiguous bytes of one output row.
//
// FFT pipeline per warp identical to logmel_kernel.cuh (TMA stage -> window -> radix-32 pass -> transpose +
// twiddle -> radix-32 pass -> real-input separation); see there for the index algebra.
#pragma once
#include "logmel_kernel.cuh"

namespace b200mel {

struct SpecSlot {  // where the columns of one task go (written by lane 0 of the task's warp every round)
    long long row0;  // element offset of (clip b, bin 0, frame t0) in the output arrays
    int n_frames;    // valid frames of the task (0 = idle slot, 1, or 2 in pair mode)
    int zero;        // 1: the task lies past the clip's own end (lengths) -> its columns are written as zeros
};

// atan2 with ~3e-7 rad absolute error (degree-7 minimax polynomial in t^2 on [0,1], octant reduction, one
// MUFU.RCP); the library atan2f costs about twice the instructions and dominated the mag+phase kernel.
__device__ __forceinline__ float fast_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float t = mx > 0.f ? __fdividef(mn, mx) : 0.f;
    const float s = t * t;
    float r = -0.004054528195410967f;
    r = fmaf(r, s, 0.021862806752324104f);
    r = fmaf(r, s, -0.055912092328071594f);
    r = fmaf(r, s, 0.09642178565263748f);
    r = fmaf(r, s, -0.13908621668815613f);
    r = fmaf(r, s, 0.19946563243865967f);
    r = fmaf(r, s, -0.33329859375953674f);
    r = fmaf(r, s, 0.9999993443489075f);
    r *= t;
    if (ay > ax) r = 1.57079632679489662f - r;
    if (x < 0.f) r = 3.14159265358979324f - r;
    return copysignf(r, y);
}

template <int kSpec>
__device__ __forceinline__ void spec_deposit(float *ta, float *tb, int idx, float re, float im, float eps) {
    if constexpr (kSpec == B200MEL_SPEC_RE_IM) {
        ta[idx] = re;
        tb[idx] = im;
    } else {
        ta[idx] = sqrt_approx(fmaf(re, re, im * im) + eps);
        if constexpr (kSpec == B200MEL_SPEC_MAG_PHASE) tb[idx] = fast_atan2(im, re);
    }
}

// kWarps warps per CTA, each running kRT tasks per round; the CTA tile holds kWarps * kRT * (2 | 1) frame columns:
//   |X| only (one tile):        16 warps x 1 task  -> 32 columns in pair mode: a warp stores 128 contiguous bytes of a row
//   two outputs (two tiles):     8 warps x 2 tasks -> 32 columns too (the tiles leave room for 8 warp regions only)
// Round = {every warp: kRT x (TMA stage -> window -> radix-32 -> transpose + twiddle -> radix-32 -> separation ->
// deposit its columns)}.  The row-wise write-out of round r is NOT a phase of its own: every thread's share of it
// (its column of 33 rows) is cut into four chunks that are issued between the FFT steps of round r + 1, so the
// stores drain in the background of the arithmetic instead of in a burst during which the FMA pipe idles and HBM
// sees all 148 SMs at once (measured before: 23 us of FFT + 17 us of write-out at C2).
// Two CTA barriers per round remain: "every chunk of round r has been read out of the tile" before the first
// deposit of round r + 1, and "every deposit is in" after the last.
// Shared layout (bytes): tw 8192 | window 4 n_fft | mbarriers | slots | warp regions | tile A | tile B.  A warp region
// is the transpose buffer with the sample stage overlaid at offset 0 (the stage is consumed before the transpose is
// written, and the next TMA is issued only after the transpose has been read back).
template <bool kPair, int kSpec, int kWarps, int kRT>
__global__ void __launch_bounds__(kWarps * 32, 1) spec_kernel(const KParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    constexpr int kFr = kPair ? 2 : 1;
    constexpr int kCols = kWarps * kRT * kFr;  // frame columns of the CTA tile
    constexpr int kRowStride = kCols + 1;      // odd stride: conflict-free deposits and row reads
    constexpr bool kTwo = kSpec != B200MEL_SPEC_MAG;

    float2 *s_tw = reinterpret_cast<float2 *>(smem_raw);
    float *s_win = reinterpret_cast<float *>(smem_raw + p.off_window);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + p.off_bar);
    SpecSlot *s_slot = reinterpret_cast<SpecSlot *>(smem_raw + p.off_entries);
    unsigned char *region = smem_raw + p.off_regions + warp * p.region_bytes;
    float2 *buf = reinterpret_cast<float2 *>(region);
    float *stage = reinterpret_cast<float *>(region);
    float *tile_a = reinterpret_cast<float *>(smem_raw + p.off_melw);
    float *tile_b = tile_a + p.n_freq * kRowStride;
    const uint32_t bar = smem_u32(s_bar + warp);
    const uint32_t stage_s = smem_u32(stage);
    uint32_t parity = 0;

    // task sequence of this warp: round r, sub-task s -> task (blockIdx + r gridDim) * kWarps * kRT + s * kWarps + warp
    constexpr int kRoundTasks = kWarps * kRT;
    const long long round_stride = (long long)gridDim.x * kRoundTasks;
    long long task = (long long)blockIdx.x * kRoundTasks + warp;
    long long cb = task / p.tasks_per_clip;
    int cq = (int)(task - cb * p.tasks_per_clip);
    const int sub_db = kWarps / p.tasks_per_clip, sub_dq = kWarps % p.tasks_per_clip;  // + kWarps tasks
    // + (round_stride - (kRT - 1) kWarps) tasks: pre-split on the host as stride_b / stride_q
    auto advance = [&](long long &b, int &q, int db, int dq) {
        b += db;
        q += dq;
        if (q >= p.tasks_per_clip) q -= p.tasks_per_clip, ++b;
    };

    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 32 * 32 * 8 / 16; i += blockDim.x)
        reinterpret_cast<int4 *>(s_tw)[i] = __ldg(reinterpret_cast<const int4 *>(p.tw) + i);
    for (int i = tid; i < p.n_fft / 4; i += blockDim.x)
        reinterpret_cast<int4 *>(s_win)[i] = __ldg(reinterpret_cast<const int4 *>(p.window) + i);
    asm volatile("griddepcontrol.wait;" ::: "memory");  // PDL: caller memory is only touched below this line
    if (lane == 0 && task < p.n_tasks) {
        const Task t = decode_task<kPair>(p, cb, cq);
        if (t.valid0) issue_copy(copy_geom(p, t.b, t.s_first, t.span, t.Li), stage_s, bar);
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __syncthreads();

    float2 wl = make_float2(1.f, 0.f);
    if (!kPair) wl = __ldg(p.tw_post + lane);
    const int partner = (32 - lane) & 31;

    // this thread's share of a round's write-out: column `wcol` of the tile, rows wr0, wr0 + kRowsPerPass, ...
    constexpr int kRowsPerPass = kWarps * 32 / kCols;
    const int wcol = tid % kCols, wr0 = tid / kCols;
    const int w_iters = (p.n_freq_out - wr0 + kRowsPerPass - 1) / kRowsPerPass;
    const int w_chunk = (w_iters + 3) / 4;
    auto writeout = [&](int set, int i0, int i1) {  // rows wr0 + kRowsPerPass * [i0, i1) of the round whose slots are in `set`
        const SpecSlot s = s_slot[set * kRoundTasks + wcol / kFr];
        const int f = wcol % kFr;
        if (f < s.n_frames) {
            const bool zero = s.zero == 1 || (s.zero == 2 && f == 1);
            const long long off = s.row0 + f;
            for (int i = i0; i < min(i1, w_iters); ++i) {  // output row r = physical bin r * bin_step
                const int r = wr0 + i * kRowsPerPass;
                const long long o = off + (long long)r * p.T;
                const int ti = r * p.bin_step * kRowStride + wcol;
                p.out_a[o] = zero ? 0.f : tile_a[ti];
                if constexpr (kTwo) p.out_b[o] = zero ? 0.f : tile_b[ti];
            }
        }
    };

    // all warps of the CTA run the same number of rounds (idle warps still join the barriers)
    int round = 0;
    for (long long base = (long long)blockIdx.x * kRoundTasks; base < p.n_tasks; base += round_stride, ++round) {
        const int set = round & 1;          // slot records of this round; the previous round's are in set ^ 1
        const bool drain = round > 0;       // the tile still holds the previous round: write it out while computing
#pragma unroll
        for (int sub = 0; sub < kRT; ++sub) {
            const bool have = task < p.n_tasks;
            Task t;
            t.valid0 = t.valid1 = false;
            t.b = 0, t.t0 = 0;
            if (have) t = decode_task<kPair>(p, cb, cq);
            // coordinates of this warp's next task
            if (sub + 1 < kRT) task += kWarps, advance(cb, cq, sub_db, sub_dq);
            else task += round_stride - (kRT - 1) * kWarps, advance(cb, cq, p.stride_b, p.stride_q);
            float2 a[32];
            if (sub == 0 && drain) writeout(set ^ 1, 0, w_chunk);
            if (t.valid0) {
                const CopyGeom g = copy_geom(p, t.b, t.s_first, t.span, t.Li);
                mbar_wait(bar, parity);
                parity ^= 1;
                if (g.patch) patch_stage(p, t.b, t.s_first, t.span, t.Li, g, stage, lane);
                const float *x0 = stage + g.delta + lane;
                if (kPair) {
                    const float *x1 = x0 + (t.valid1 ? p.hop : 0);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float w = s_win[32 * j + lane];
                        a[j] = __fmul2_rn(make_float2(x0[32 * j], t.valid1 ? x1[32 * j] : 0.f), make_float2(w, w));
                    }
                } else {
                    const float2 *w2 = reinterpret_cast<const float2 *>(s_win);
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        a[j] = __fmul2_rn(make_float2(x0[64 * j + lane], x0[64 * j + lane + 1]), w2[32 * j + lane]);
                }
                fft32(a);
                __syncwarp();  // the stage has been consumed by every lane: the transpose may overwrite it
                static_for<0, 32>([&](auto k1_) {
                    constexpr int k1 = decltype(k1_)::value;
                    buf[k1 * kBufStride + lane] = a[fft32_pos(k1)];
                });
                __syncwarp();
                if (sub == 0 && drain) writeout(set ^ 1, w_chunk, 2 * w_chunk);
                xpose_read_twiddle<16>(a, buf, s_tw, lane);
                __syncwarp();
            } else if (sub == 0 && drain) {
                writeout(set ^ 1, w_chunk, 2 * w_chunk);
            }
            // prefetch this warp's next task into the (now free) stage
            if (lane == 0 && task < p.n_tasks) {
                const Task n = decode_task<kPair>(p, cb, cq);
                if (n.valid0) issue_copy(copy_geom(p, n.b, n.s_first, n.span, n.Li), stage_s, bar);
            }
            const int slot = sub * kWarps + warp;
            if (lane == 0) {
                SpecSlot s;
                s.n_frames = have ? (kPair && p.pair_frames == 2 ? (t.t0 + 1 < p.T ? 2 : 1) : 1) : 0;
                s.zero = have && !t.valid0;
                s.row0 = have ? (t.b * p.n_freq_out) * (long long)p.T + t.t0 : 0;
                if (have && kPair && p.pair_frames == 2 && t.valid0 && !t.valid1 && t.t0 + 1 < p.T) s.zero = 2;  // second frame only
                s_slot[set * kRoundTasks + slot] = s;
            }
            if (sub == 0 && drain) writeout(set ^ 1, 2 * w_chunk, 3 * w_chunk);
            if (t.valid0) fft32(a);
            if (sub == 0) {
                if (drain) writeout(set ^ 1, 3 * w_chunk, 4 * w_chunk);
                __syncthreads();  // the previous round has left the tile: deposits may begin
            }
            if (t.valid0) {
                const int col = slot * kFr;
                static_for<0, 16>([&](auto k2_) {
                    constexpr int k2 = decltype(k2_)::value;
                    const float2 A = a[fft32_pos(k2)];
                    const float2 g0 = a[fft32_pos((32 - k2) & 31)];
                    const float2 g1 = a[fft32_pos(31 - k2)];
                    float2 Bv;
                    Bv.x = __shfl_sync(0xffffffffu, lane == 0 ? g0.x : g1.x, partner);
                    Bv.y = __shfl_sync(0xffffffffu, lane == 0 ? g0.y : g1.y, partner);
                    const int k = lane + 32 * k2;
                    const float2 Bc = make_float2(Bv.x, -Bv.y);                     // conj(B): an operand sign pattern
                    const float2 E = __fadd2_rn(A, Bc);                             // frame t
                    const float2 D = __fadd2_rn(A, make_float2(-Bc.x, -Bc.y));
                    const float2 O = make_float2(D.y, -D.x);                        // frame t+1 = -i (A - conj(B))
                    if constexpr (kPair) {
                        spec_deposit<kSpec>(tile_a, tile_b, k * kRowStride + col, E.x, E.y, p.mag_eps);
                        spec_deposit<kSpec>(tile_a, tile_b, k * kRowStride + col + 1, O.x, O.y, p.mag_eps);
                    } else {
                        constexpr float w64c = TwConst::c64[k2], w64s = TwConst::s64[k2];
                        const float2 P = cmul(O, cmul(wl, make_float2(w64c, w64s)));
                        const float2 X0 = cadd(E, P), X1 = csub(E, P);
                        spec_deposit<kSpec>(tile_a, tile_b, k * kRowStride + col, X0.x, X0.y, p.mag_eps);
                        spec_deposit<kSpec>(tile_a, tile_b, (1024 - k) * kRowStride + col, X1.x, -X1.y, p.mag_eps);
                    }
                });
                if (lane == 0) {
                    const float2 A = a[fft32_pos(16)];
                    if constexpr (kPair) {
                        spec_deposit<kSpec>(tile_a, tile_b, 512 * kRowStride + col, 2.f * A.x, 0.f, p.mag_eps);
                        spec_deposit<kSpec>(tile_a, tile_b, 512 * kRowStride + col + 1, 2.f * A.y, 0.f, p.mag_eps);
                    } else {
                        spec_deposit<kSpec>(tile_a, tile_b, 512 * kRowStride + col, 2.f * A.x, -2.f * A.y, p.mag_eps);
                    }
                }
            }
        }
        __syncthreads();  // every task's columns (and slot records) are in the tile
    }
    if (round > 0) writeout((round - 1) & 1, 0, 4 * w_chunk);  // the last round has no successor to hide behind
}

}  // namespace b200mel
