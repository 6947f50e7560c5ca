// spec_kernel.cuh — spectrum-output variant of the fused STFT kernel (sm_100a): STFT.transform (mag, phase),
// STFTTorchAudio.forward (re, im) and magnitude-only.  Replaces models/transforms.py:53-69 and :297-311.
//
// These operators write (B, n_fft/2+1, T) tensors — 6.4x (one array) to 12.8x (two arrays) the bytes of the
// mel output — so how the stores coalesce matters as much as the FFT.  A warp only ever holds two frames of a bin
// (8 contiguous bytes of an output row), which made per-warp stores crawl (32 partial sectors per instruction).
// Warps therefore cooperate in GROUPS of 4 (pair mode; 8 in split mode): a group works on consecutive tasks — 8
// consecutive frames of a clip — pools its spectra in a shared tile [bin][8 frame columns] and writes it out
// row-wise, 8 lanes storing the 32 contiguous bytes (one sector) of a row.  Groups synchronise only among
// themselves (named barriers), so the 2-4 groups of a CTA drift apart and the SM always has warps in different
// phases — the CTA-wide lock step of the first two versions (16 x 1 / 8 x 2 warps on one 32-column tile, two
// __syncthreads per round) serialised the FMA-bound and the store-bound phases of all 16 warps.
//
// FFT pipeline per warp identical to logmel_kernel.cuh (TMA stage -> window -> radix-32 pass -> transpose +
// twiddle -> radix-32 pass -> real-input separation); see there for the index algebra.
#pragma once
#include "logmel_kernel.cuh"
#include "logmel_fast.cuh"

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

// kGroups groups of kGW warps per CTA; a group's tile holds kGW * (2 | 1) = 8 frame columns per output array.
// Tile element (bin k, column c) sits at k * 8 + (c ^ ((k >> 2) & 7)): 32 consecutive bins of one column (a deposit)
// and 4 consecutive rows x 8 columns (a write-out instruction) both touch 32 different banks, without row padding.
// Round of a group = {every warp: TMA stage -> window -> radix-32 -> transpose + twiddle -> radix-32 -> separation ->
// deposit its columns}.  The row-wise write-out of round r is NOT a phase of its own: every thread's share of it
// is cut into four chunks that are issued between the FFT steps of round r + 1, so the stores drain in the background
// of the arithmetic.  Two group barriers per round remain: "every chunk of round r has been read out of the tile"
// before the first deposit of round r + 1, and "every deposit is in" after the last.
// Shared layout (bytes): tw 8192 | window 4 n_fft | mbarriers | slots | warp regions | per group: tile A (| tile B).
// A warp region is the transpose buffer with the sample stage overlaid at offset 0 (the stage is consumed before the
// transpose is written, and the next TMA is issued only after the transpose has been read back).
template <bool kPair>
struct SpecShape {
    static constexpr int kGW = kPair ? 4 : 8;   // warps per group
    static constexpr int kFr = kPair ? 2 : 1;   // frames per task
    static constexpr int kCols = 8;             // kGW * kFr
};
__device__ __forceinline__ int spec_tile_index(int k, int c) { return k * 8 + (c ^ ((k >> 2) & 7)); }

// kFast (pair mode only): the common geometry fixed at compile time exactly as in logmel_fast.cuh — hop 256, full
// periodic Hann generated in registers, no per-clip `lengths` — so a task is located by one FastDesc computed when
// its samples are requested, the two overlapping frames share their stage loads, and the zero-fill paths are gone.
template <bool kPair, int kSpec, int kGroups, bool kFast = false>
__global__ void __launch_bounds__(kGroups * SpecShape<kPair>::kGW * 32, 1) spec_kernel(const KParams p) {
    static_assert(!kFast || kPair, "the fast instantiation is a pair-mode kernel");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int kGW = SpecShape<kPair>::kGW, kFr = SpecShape<kPair>::kFr, kCols = SpecShape<kPair>::kCols;
    constexpr bool kTwo = kSpec != B200MEL_SPEC_MAG;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int grp = warp / kGW, gw = warp % kGW, gtid = tid - grp * kGW * 32;   // group, warp / thread inside the group

    float2 *s_tw = reinterpret_cast<float2 *>(smem_raw);
    float *s_win = reinterpret_cast<float *>(smem_raw + p.off_window);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + p.off_bar);
    SpecSlot *s_slot = reinterpret_cast<SpecSlot *>(smem_raw + p.off_entries) + grp * 2 * kGW;  // [2 sets][kGW]
    unsigned char *region = smem_raw + p.off_regions + warp * p.region_bytes;
    float2 *buf = reinterpret_cast<float2 *>(region);
    float *stage = reinterpret_cast<float *>(region);
    const int tile_len = p.n_freq * kCols;  // floats per tile
    float *tile_a = reinterpret_cast<float *>(smem_raw + p.off_melw) + grp * (kTwo ? 2 : 1) * tile_len;
    float *tile_b = tile_a + tile_len;
    const uint32_t bar = smem_u32(s_bar + warp);
    const uint32_t stage_s = smem_u32(stage);
    uint32_t parity = 0;
    auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(kGW * 32) : "memory"); };

    // task sequence of this warp: round R of group gg = blockIdx * kGroups + grp -> task (gg + R * gridDim * kGroups) * kGW + gw;
    // the step between a warp's tasks is pre-split on the host as stride_b / stride_q
    long long task = ((long long)blockIdx.x * kGroups + grp) * kGW + gw;
    const long long step = (long long)gridDim.x * kGroups * kGW;
    long long cb = task / p.tasks_per_clip;
    int cq = (int)(task - cb * p.tasks_per_clip);

    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 32 * 32 * 8 / 16; i += blockDim.x)
        reinterpret_cast<int4 *>(s_tw)[i] = __ldg(reinterpret_cast<const int4 *>(p.tw) + i);
    for (int i = tid; i < p.n_fft / 4; i += blockDim.x)
        reinterpret_cast<int4 *>(s_win)[i] = __ldg(reinterpret_cast<const int4 *>(p.window) + i);
    asm volatile("griddepcontrol.wait;" ::: "memory");  // PDL: caller memory is only touched below this line
    FastDesc cur;
    cur.b = cur.t0 = cur.delta = cur.Li = 0;
    cur.flags = 0;
    if constexpr (kFast) {
        __syncwarp();
        if (task < p.n_tasks) cur = fast_request(p, (int)cb, cq, lane, stage_s, bar);
    } else if (lane == 0 && task < p.n_tasks) {
        const Task t = decode_task<kPair>(p, cb, cq);
        if (t.valid0) issue_copy(copy_geom(p, t.b, t.s_first, t.span, t.Li), stage_s, bar);
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __syncthreads();
    float2 hann_cs = make_float2(0.f, 0.f);
    if constexpr (kFast) {
        float sn, cs;
        sincospif((float)lane * (1.0f / 512.0f), &sn, &cs);
        hann_cs = make_float2(0.25f * cs, 0.25f * sn);
    }

    float2 wl = make_float2(1.f, 0.f);
    if (!kPair) wl = __ldg(p.tw_post + lane);
    const int partner = (32 - lane) & 31;

    // this thread's share of a round's write-out: column `wcol` of the group's tile, rows wr0, wr0 + kRowsPerPass, ...
    constexpr int kRowsPerPass = kGW * 32 / kCols;
    const int wcol = gtid % kCols, wr0 = gtid / kCols;
    const int w_iters = (p.n_freq_out - wr0 + kRowsPerPass - 1) / kRowsPerPass;
    const int w_chunk = (((w_iters + 3) / 4) + 1) & ~1;   // even, so a chunk starts on an even iteration
    // bin_step == 1 (n_fft = 1024 / 2048): iteration i reads tile[(wr0 + P i) * 8 + (wcol ^ swz)] with P = kRowsPerPass, and
    // swz = ((wr0 + P i) >> 2) & 7 alternates between two values (P = 16) or is constant (P = 32): two precomputed tile
    // offsets, compile-time strides and one 64-bit pointer bump per pair of stores instead of a multiply-add chain per store.
    const int w_off0 = wr0 * 8 + (wcol ^ ((wr0 >> 2) & 7));
    const int w_off1 = (wr0 + kRowsPerPass) * 8 + (wcol ^ (((wr0 + kRowsPerPass) >> 2) & 7));
    const long long w_rowstep = (long long)kRowsPerPass * p.T;
    auto writeout = [&](int set, int i0, int i1) {  // rows wr0 + kRowsPerPass * [i0, i1) of the round whose slots are in `set`
        const SpecSlot s = s_slot[set * kGW + wcol / kFr];
        const int f = wcol % kFr;
        if (f >= s.n_frames) return;
        const bool zero = !kFast && (s.zero == 1 || (s.zero == 2 && f == 1));
        const long long off = s.row0 + f;
        i1 = min(i1, w_iters);
        if (kFast || (p.bin_step == 1 && !zero)) {
            float *pa = p.out_a + off + (long long)(wr0 + i0 * kRowsPerPass) * p.T;
            float *pb = kTwo ? p.out_b + off + (long long)(wr0 + i0 * kRowsPerPass) * p.T : nullptr;
            const float *ta = tile_a + i0 * (kRowsPerPass * 8), *tb = tile_b + i0 * (kRowsPerPass * 8);
            int i = i0;   // even
#pragma unroll 2
            for (; i + 1 < i1; i += 2) {
                const float a0 = ta[w_off0], a1 = ta[w_off1];
                pa[0] = a0, pa[w_rowstep] = a1;
                pa += 2 * w_rowstep;
                if constexpr (kTwo) {
                    const float b0 = tb[w_off0], b1 = tb[w_off1];
                    pb[0] = b0, pb[w_rowstep] = b1;
                    pb += 2 * w_rowstep, tb += 2 * kRowsPerPass * 8;
                }
                ta += 2 * kRowsPerPass * 8;
            }
            if (i < i1) {
                pa[0] = ta[w_off0];
                if constexpr (kTwo) pb[0] = tb[w_off0];
            }
            return;
        }
        for (int i = i0; i < i1; ++i) {  // output row r = physical bin r * bin_step
            const int r = wr0 + i * kRowsPerPass;
            const long long o = off + (long long)r * p.T;
            const int ti = spec_tile_index(r * p.bin_step, wcol);
            p.out_a[o] = zero ? 0.f : tile_a[ti];
            if constexpr (kTwo) p.out_b[o] = zero ? 0.f : tile_b[ti];
        }
    };

    // all warps of a group run the same number of rounds (idle warps still join the group's barriers)
    int round = 0;
    for (long long base = task - gw; base < p.n_tasks; base += step, ++round) {
        const int set = round & 1;          // slot records of this round; the previous round's are in set ^ 1
        const bool drain = round > 0;       // the tile still holds the previous round: write it out while computing
        const bool have = task < p.n_tasks;
        Task t;
        t.valid0 = t.valid1 = false;
        t.b = 0, t.t0 = 0;
        const FastDesc d = cur;
        if constexpr (kFast) {
            t.valid0 = have, t.valid1 = have && (d.flags & 2u);
            t.b = d.b, t.t0 = d.t0;
        } else if (have) {
            t = decode_task<kPair>(p, cb, cq);
        }
        // coordinates of this warp's next task
        task += step;
        cb += p.stride_b, cq += p.stride_q;
        if (cq >= p.tasks_per_clip) cq -= p.tasks_per_clip, ++cb;
        float2 a[32];
        if (drain) writeout(set ^ 1, 0, w_chunk);
        if (t.valid0) {
            if constexpr (kFast) {
                mbar_wait(bar, parity);
                parity ^= 1;
                if (d.flags & 4u) fast_patch_halo(p, d, stage, lane);
                fast_load_windowed(a, stage + d.delta + lane, hann_cs);
            } else {
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
            }
            fft32(a);
            __syncwarp();  // the stage has been consumed by every lane: the transpose may overwrite it
            static_for<0, 32>([&](auto k1_) {
                constexpr int k1 = decltype(k1_)::value;
                buf[k1 * kBufStride + lane] = a[fft32_pos(k1)];
            });
            __syncwarp();
            if (drain) writeout(set ^ 1, w_chunk, 2 * w_chunk);
            xpose_read_twiddle<16>(a, buf, s_tw, lane);
            __syncwarp();
        } else if (drain) {
            writeout(set ^ 1, w_chunk, 2 * w_chunk);
        }
        // prefetch this warp's next task into the (now free) stage
        if constexpr (kFast) {
            if (task < p.n_tasks) cur = fast_request(p, (int)cb, cq, lane, stage_s, bar);
        } else if (lane == 0 && task < p.n_tasks) {
            const Task n = decode_task<kPair>(p, cb, cq);
            if (n.valid0) issue_copy(copy_geom(p, n.b, n.s_first, n.span, n.Li), stage_s, bar);
        }
        if (lane == 0) {
            SpecSlot s;
            if constexpr (kFast) {
                s.n_frames = have ? ((d.flags & 2u) ? 2 : 1) : 0;
                s.zero = 0;
            } else {
                s.n_frames = have ? (kPair && p.pair_frames == 2 ? (t.t0 + 1 < p.T ? 2 : 1) : 1) : 0;
                s.zero = have && !t.valid0;
                if (have && kPair && p.pair_frames == 2 && t.valid0 && !t.valid1 && t.t0 + 1 < p.T) s.zero = 2;  // second frame only
            }
            s.row0 = have ? (t.b * p.n_freq_out) * (long long)p.T + t.t0 : 0;
            s_slot[set * kGW + gw] = s;
        }
        if (drain) writeout(set ^ 1, 2 * w_chunk, 3 * w_chunk);
        if (t.valid0) fft32(a);
        if (drain) writeout(set ^ 1, 3 * w_chunk, 4 * w_chunk);
        group_sync();  // the previous round has left the tile: deposits may begin
        if (t.valid0) {
            const int col = gw * kFr;
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
                    spec_deposit<kSpec>(tile_a, tile_b, spec_tile_index(k, col), E.x, E.y, p.mag_eps);
                    spec_deposit<kSpec>(tile_a, tile_b, spec_tile_index(k, col + 1), O.x, O.y, p.mag_eps);
                } else {
                    constexpr float w64c = TwConst::c64[k2], w64s = TwConst::s64[k2];
                    const float2 P = cmul(O, cmul(wl, make_float2(w64c, w64s)));
                    const float2 X0 = cadd(E, P), X1 = csub(E, P);
                    spec_deposit<kSpec>(tile_a, tile_b, spec_tile_index(k, col), X0.x, X0.y, p.mag_eps);
                    spec_deposit<kSpec>(tile_a, tile_b, spec_tile_index(1024 - k, col), X1.x, -X1.y, p.mag_eps);
                }
            });
            if (lane == 0) {
                const float2 A = a[fft32_pos(16)];
                if constexpr (kPair) {
                    spec_deposit<kSpec>(tile_a, tile_b, spec_tile_index(512, col), 2.f * A.x, 0.f, p.mag_eps);
                    spec_deposit<kSpec>(tile_a, tile_b, spec_tile_index(512, col + 1), 2.f * A.y, 0.f, p.mag_eps);
                } else {
                    spec_deposit<kSpec>(tile_a, tile_b, spec_tile_index(512, col), 2.f * A.x, -2.f * A.y, p.mag_eps);
                }
            }
        }
        group_sync();  // every task's columns (and slot records) are in the tile
    }
    if (round > 0) writeout((round - 1) & 1, 0, 4 * w_chunk);  // the last round has no successor to hide behind
}

}  // namespace b200mel
