// stft_tc.cuh — the fused STFT -> |.| -> mel -> log kernel with BOTH radix-32 DFT stages on the tcgen05 tensor cores
// (sm_100a), accumulators in TMEM.  Same operator as logmel_kernel.cuh (LogMelSpectrogram.forward,
// models/transforms.py:231-244 on top of STFT.transform :53-69) for the geometry every default of the reference uses:
// n_fft = win_length = 1024 (periodic Hann), hop 256, filterbank below bin 384.
//
// Algebra (tests/tc_model.py is the numpy restatement, checked against numpy.fft): n = 32 n1 + n2, k = k1 + 32 k2,
//
//   stage 1  A[k1, n2]   = sum_n1 W32^(k1 n1) x[32 n1 + n2]       GEMM  [rows (n2, frame)] x [32 n1] x [32 cols]
//   twiddle  A'[k1, n2]  = W1024^(k1 n2) A[k1, n2]                 CUDA cores, between the GEMMs
//   Hann     Aw'[k1]     = A'[k1]/2 - (A'[k1-1] + A'[k1+1])/4      the window as a 3-tap filter over k1 (time-domain
//                                                                  windowing would break the sample sharing below)
//   stage 2  X[k1+32 k2] = sum_n2 W32^(n2 k2) Aw'[k1, n2]          GEMM  [rows (k1, frame)] x [64 (n2, re/im)] x [48]
//
// Real input: only k1 = 0..16 exist; rows k1 = 1..15 of stage 2 produce k2 in {0..11} u {20..31}, the upper twelve
// being conjugates of bins (32 - k1) + 32 (31 - k2); rows 0 and 16 (real stage-1 outputs) share ONE packed row that
// is multiplied by a second matrix B' — 16 rows per frame, 128 = one UMMA M tile per 8 frames.
//
// Precision: every operand is split into two fp16 limbs (x = hi + lo, 22 significant bits; a power-of-two scale per
// batch keeps both limbs in fp16's normal range) and a product is hi hi + lo hi + hi lo, accumulated in fp32 in TMEM:
// the spectrum is as accurate as an fp32 FFT (tests/test_tc_algebra.py).  The two constant limbs are concatenated along
// N, so one MMA A_hi x [B_hi | B_lo] and one A_lo x B_hi do the three products and the epilogue adds the two halves.
//
// Stage-1 operand ("Hankel" layout): the fp16 samples of 8 consecutive frames are stored ONCE, transposed,
// S[n2][q] = x[32 q + n2]; frame t's row (n2, t) is S[n2][8 t .. 8 t + 31], so the rows of one UMMA core matrix
// (8 frames) are the same bytes shifted by 16 — the descriptor's K-direction stride is 16 bytes and its row-group
// stride is the pitch of S.  Overlapping frames (hop = n_fft / 4) are converted and stored once instead of four times.
//
// One CTA per SM, 16 warps, batches of 8 frames of one clip.  Per batch:
//   P1  stage (TMA bulk copy, fp32) -> max |x| -> scale -> fp16 hi / lo -> S           warps 0..10
//   M1  8 tcgen05.mma (2 row tiles x 2 K steps x {hi, lo})  -> D1 (TMEM, 2 x 64 columns)
//   P2  D1 -> registers, hi + lo halves, twiddle, Hann 3-tap, fp16 hi / lo -> A2 (K-major core matrices)
//   M2  8 tcgen05.mma (4 K steps x {hi, lo}), N = 192 = [B | B'] x {hi, lo}   -> D2 (TMEM, 2 x 192 columns)
//   P3  D2 -> registers, |.| -> magnitude tile [bin][frame]
//   P4  banded mel (4 rows x 8 frames per warp instruction), log / clamp / norm, stores
// The loop is software-pipelined: M2 of batch i and M1 of batch i + 1 run on the tensor pipe while the CUDA cores do
// P1 of batch i + 1 and P3 / P4 of batch i - 1.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "logmel_kernel.cuh"
#include "mel_tc.cuh"

namespace b200mel {

constexpr int kStcThreads = 512;
constexpr int kStcGroup = 8;                         // frames per batch
constexpr int kStcHop = 256, kStcNfft = 1024;
constexpr int kStcSpan = (kStcGroup - 1) * kStcHop + kStcNfft;  // 2816 samples per batch
constexpr int kStcQ = kStcSpan / 32;                 // 88 fp16 per Hankel row
constexpr int kStcPitch = kStcQ * 2;                 // 176 bytes = 11 x 16: odd multiple of 16, so 8 rows hit 8 bank groups
constexpr int kStcBins = 384;                        // spectrum bins produced
constexpr int kStcMagPad = 4;                        // zero rows in front of the magnitude tile (mel windows may start at -3)
constexpr int kStcShift = 6;                         // stage-2 operands are scaled by 2^-6
constexpr int kStcMaxGroups = 32;                    // mel rows / 4
constexpr int kStcMelSlots = 4;                      // row groups a warp may own

// shared-memory carve-up (bytes)
constexpr int kStcStageBytes = 11392;                // (2816 + 8) floats rounded to 128
constexpr int kStcOffStage = 0;
constexpr int kStcOffHank = 2 * kStcStageBytes;                  // hi limb, then lo limb
constexpr int kStcHankBytes = 32 * kStcPitch;                    // 5632
constexpr int kStcOffB1 = kStcOffHank + 2 * kStcHankBytes;       // [4 k chunks][64 n][8 k] fp16
constexpr int kStcB1Bytes = 4096;
constexpr int kStcOffA2 = kStcOffB1 + kStcB1Bytes;               // hi tile, then lo tile: [8 k chunks][16 slots][8 frames][8 k]
constexpr int kStcA2Bytes = 16384;
constexpr int kStcOffB2 = kStcOffA2 + 2 * kStcA2Bytes;           // [8 k chunks][192 n][8 k] fp16
constexpr int kStcB2Bytes = 24576;
constexpr int kStcOffTw = kStcOffB2 + kStcB2Bytes;               // float2 [17][32]
constexpr int kStcTwBytes = 17 * 32 * 8;
constexpr int kStcOffMag = kStcOffTw + kStcTwBytes;              // float [4 + 384][8]
constexpr int kStcMagBytes = (kStcMagPad + kStcBins) * kStcGroup * 4;
constexpr int kStcOffMisc = kStcOffMag + kStcMagBytes;           // mbarriers, TMEM address, per-warp maxima, descale ring
constexpr int kStcMiscBytes = 256;
constexpr int kStcOffMel = kStcOffMisc + kStcMiscBytes;          // mel schedule (header + weights), size from the plan
constexpr int kStcMelHeader = (16 * kStcMelSlots + 2 * kStcMaxGroups) * 4 + kStcMaxGroups * 4 * 8;  // grp | glen | gwoff | ent

struct StcParams {
    KParams k;                   // waveform geometry, outputs, epilogue (the fields copy_geom / epilogue() read)
    const unsigned char *tables; // global blob: B1 | B2 | tw | mel schedule, copied verbatim to kStcOffB1.. / kStcOffMel
    int mel_bytes;
    int groups_per_clip;         // ceil(T / 8)
    long long n_batches;         // B * groups_per_clip
    int power;
    // debug taps (tests only; null in production): magnitudes (B, 384, T) and, for batch 0, the raw operands
    float *dbg_mag, *dbg_d1, *dbg_d2;
    unsigned char *dbg_a2;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float *v) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// (x, y) -> fp16 limbs packed as {x, y} (x in the low half): v = hi + lo to 22 bits
__device__ __forceinline__ void split_f16x2(float x, float y, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

template <int kPower>
__global__ void __launch_bounds__(kStcThreads, 1) stft_tc_kernel(const StcParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *s_mag = reinterpret_cast<float *>(smem + kStcOffMag);
    const float2 *s_tw = reinterpret_cast<const float2 *>(smem + kStcOffTw);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + kStcOffMisc);  // [0,1] TMA, [2] M1, [3,4] M2
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + kStcOffMisc + 48);
    uint32_t *s_wmax = reinterpret_cast<uint32_t *>(smem + kStcOffMisc + 64);  // [16]
    float *s_descale = reinterpret_cast<float *>(smem + kStcOffMisc + 128);    // [4]
    const int *s_grp = reinterpret_cast<const int *>(smem + kStcOffMel);                       // [16][kStcMelSlots]
    const int *s_glen = s_grp + 16 * kStcMelSlots;                                              // [kStcMaxGroups]
    const int *s_gwoff = s_glen + kStcMaxGroups;                                                // [kStcMaxGroups]
    const int2 *s_ent = reinterpret_cast<const int2 *>(s_gwoff + kStcMaxGroups);                // [kStcMaxGroups * 4] {m, lo}
    const float *s_melw = reinterpret_cast<const float *>(smem + kStcOffMel + kStcMelHeader);

    // ------------------------------------------------------------------ one-time setup
    {   // constant operands and schedules: one contiguous blob per destination
        const int4 *src = reinterpret_cast<const int4 *>(p.tables);
        int4 *d1 = reinterpret_cast<int4 *>(smem + kStcOffB1);
        for (int i = tid; i < kStcB1Bytes / 16; i += kStcThreads) d1[i] = __ldg(src + i);
        src += kStcB1Bytes / 16;
        int4 *d2 = reinterpret_cast<int4 *>(smem + kStcOffB2);
        for (int i = tid; i < kStcB2Bytes / 16; i += kStcThreads) d2[i] = __ldg(src + i);
        src += kStcB2Bytes / 16;
        int4 *d3 = reinterpret_cast<int4 *>(smem + kStcOffTw);
        for (int i = tid; i < kStcTwBytes / 16; i += kStcThreads) d3[i] = __ldg(src + i);
        src += kStcTwBytes / 16;
        int4 *d4 = reinterpret_cast<int4 *>(smem + kStcOffMel);
        for (int i = tid; i < p.mel_bytes / 16; i += kStcThreads) d4[i] = __ldg(src + i);
        for (int i = tid; i < kStcMagPad * kStcGroup; i += kStcThreads) s_mag[i] = 0.f;
    }
    if (tid == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(smem_u32(s_bar + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t tm_d1 = tmem;            // 2 tiles x 64 columns
    const uint32_t tm_d2 = tmem + 128;      // 2 buffers x 192 columns

    // instruction descriptors: D fp32, A / B fp16, both K-major, M = 128
    constexpr uint32_t kIdesc = (1u << 4) | (8u << 24);
    constexpr uint32_t kI64 = kIdesc | (8u << 17), kI32 = kIdesc | (4u << 17), kI192 = kIdesc | (24u << 17), kI96 = kIdesc | (12u << 17);
    const uint32_t sm_base = smem_u32(smem);

    const long long nb_total = p.n_batches;
    const long long first = blockIdx.x, stride = gridDim.x;
    const int n_mine = first < nb_total ? (int)((nb_total - first + stride - 1) / stride) : 0;
    const int Gc = p.groups_per_clip;

    asm volatile("griddepcontrol.wait;" ::: "memory");  // PDL: nothing above touched caller memory

    // geometry of my batch number i
    auto batch_geom = [&](int i, long long &b, int &t0, int &s_first) {
        const long long g = first + (long long)i * stride;
        b = g / Gc;
        t0 = (int)(g - b * Gc) * kStcGroup;
        s_first = t0 * kStcHop - p.k.pad;
    };
    auto issue_tma = [&](int i) {  // one thread
        long long b; int t0, s_first;
        batch_geom(i, b, t0, s_first);
        const CopyGeom g = copy_geom(p.k, b, s_first, kStcSpan, p.k.L);
        issue_copy(g, sm_base + kStcOffStage + (i & 1) * kStcStageBytes, smem_u32(s_bar + (i & 1)));
    };

    // ---- P1: staged fp32 samples of batch i -> scaled fp16 limbs in the Hankel layout
    auto phase1 = [&](int i) {
        long long b; int t0, s_first;
        batch_geom(i, b, t0, s_first);
        float v[8];
        uint32_t amax = 0u;
        if (warp < 11) {
            const CopyGeom g = copy_geom(p.k, b, s_first, kStcSpan, p.k.L);
            mbar_wait(smem_u32(s_bar + (i & 1)), (uint32_t)(i >> 1) & 1u);
            const float *st = reinterpret_cast<const float *>(smem + kStcOffStage + (i & 1) * kStcStageBytes) + g.delta - s_first;
            const int s0 = s_first + 256 * warp + lane;  // q = 8 warp + j
            if (!g.patch) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = st[s0 + 32 * j];
            } else {
                const float *row = p.k.wav + b * p.k.row_stride;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int s = s0 + 32 * j;
                    v[j] = (s >= g.c_lo && s < g.c_hi) ? st[s] : __ldg(row + reflect_index(s, p.k.L));
                }
            }
            float m = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) m = fmaxf(m, fabsf(v[j]));  // NaN samples are caught below (fabsf keeps them out of fmaxf)
#pragma unroll
            for (int j = 0; j < 8; ++j) if (!(fabsf(v[j]) <= 3.0e38f)) m = __uint_as_float(0x7f800000u);
            amax = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
        }
        if (lane == 0) s_wmax[warp] = amax;
        __syncthreads();
        uint32_t mx = 0u;
#pragma unroll
        for (int w = 0; w < 16; ++w) mx = max(mx, s_wmax[w]);
        // power-of-two scale: max |x| * sc in [2^13, 2^14); 1 for silence / non-finite input
        int sexp = 127;
        if (mx != 0u && mx < 0x7f800000u) {
            sexp = 267 - (int)(mx >> 23);
            sexp = min(max(sexp, 1), 254);
            if (p.k.mag_eps > 0.f) sexp = min(sexp, 127 + 40);
        }
        const float sc = __uint_as_float((uint32_t)sexp << 23);
        if (tid == 0) s_descale[i & 3] = __uint_as_float((uint32_t)(260 - sexp) << 23);  // 2^6 / sc
        if (warp < 11) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) split_f16x2(v[2 * j] * sc, v[2 * j + 1] * sc, hi[j], lo[j]);
            unsigned char *dst = smem + kStcOffHank + lane * kStcPitch + warp * 16;
            *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4 *>(dst + kStcHankBytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async();
    };
    auto issue_m1 = [&]() {  // one thread, after the barrier that follows P1
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const uint32_t a_off = (uint32_t)(16 * h * kStcPitch + 32 * s);
                const uint64_t a_hi = tc_desc(sm_base + kStcOffHank + a_off, 16u, (uint32_t)kStcPitch);
                const uint64_t a_lo = tc_desc(sm_base + kStcOffHank + kStcHankBytes + a_off, 16u, (uint32_t)kStcPitch);
                const uint64_t bd = tc_desc(sm_base + kStcOffB1 + 2048u * s, 1024u, 128u);
                tc_mma(tm_d1 + 64u * h, a_hi, bd, kI64, s);
                tc_mma(tm_d1 + 64u * h, a_lo, bd, kI32, 1u);
            }
        tc_commit(smem_u32(s_bar + 2));
    };

    // ---- P2: D1 -> twiddle, Hann 3-tap -> fp16 limbs of the stage-2 operand
    auto phase2 = [&](int i) {
        const int quad = warp & 3, h = (warp >> 2) & 1, kh = warp >> 3;
        const int n2 = 16 * h + 4 * quad + (lane >> 3), t = lane & 7;
        const uint32_t ta = tm_d1 + ((uint32_t)(32 * quad) << 16) + 64u * h;
        float a[20], l[20];   // kh 0: columns 0..19 (A0, A16, A1..A9); kh 1: [0..15] = columns 16..31 (A8..A15), [16..19] = columns 0..3
        if (kh == 0) {
            tmem_ld16(ta, a), tmem_ld4(ta + 16, a + 16), tmem_ld16(ta + 32, l), tmem_ld4(ta + 48, l + 16);
        } else {
            tmem_ld16(ta + 16, a), tmem_ld4(ta, a + 16), tmem_ld16(ta + 48, l), tmem_ld4(ta + 32, l + 16);
        }
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 20; ++c) a[c] += l[c];
        if (p.dbg_d1 && first == 0 && i == 0) {
            float *d = p.dbg_d1 + (n2 * 8 + t) * 32;
#pragma unroll
            for (int c = 0; c < 20; ++c) d[kh == 0 ? c : (c < 16 ? 16 + c : c - 16)] = a[c];
        }
        const float S = 1.0f / (float)(1 << kStcShift);
        float2 ap[10];    // twiddled values A'[k]: kh 0: k = 0..9; kh 1: k = 8..16 in [0..8]
        float2 out0;      // kh 1: the packed row
        const float2 *tw = s_tw + n2;
        if (kh == 0) {
            ap[0] = make_float2(a[0] * S, 0.f);
#pragma unroll
            for (int k = 1; k < 10; ++k) ap[k] = cmul(make_float2(a[2 * k], a[2 * k + 1]), tw[32 * k]);
        } else {
#pragma unroll
            for (int k = 8; k < 16; ++k) ap[k - 8] = cmul(make_float2(a[2 * (k - 8)], a[2 * (k - 8) + 1]), tw[32 * k]);
            const float2 w16 = tw[32 * 16], w1 = tw[32];
            ap[8] = make_float2(a[17] * w16.x, a[17] * w16.y);                 // A'[16] = W^(16 n2) A[16], A[16] real
            const float2 ap1 = cmul(make_float2(a[18], a[19]), w1);            // A'[1]
            out0.x = 0.5f * S * a[16] - 0.5f * ap1.x;                          // Aw'[0] (real)
            out0.y = 0.5f * S * a[17] - 0.5f * fmaf(w1.x, a[14], w1.y * a[15]);  // r16: Aw'[16] = W64^n2 r16
        }
        unsigned char *dst = smem + kStcOffA2 + t * 16 + (n2 >> 2) * 2048 + (n2 & 3) * 4;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            // kh 0: slots 1..8 (= k1); kh 1: slots 9..15 and the packed slot 0
            const int slot = kh == 0 ? o + 1 : (o < 7 ? o + 9 : 0);
            float2 w;
            if (kh == 1 && o == 7) w = out0;
            else {
                const int c = kh == 0 ? o + 1 : o + 1;   // index of A'[slot] inside ap[]
                w.x = fmaf(-0.25f, ap[c - 1].x + ap[c + 1].x, 0.5f * ap[c].x);
                w.y = fmaf(-0.25f, ap[c - 1].y + ap[c + 1].y, 0.5f * ap[c].y);
            }
            uint32_t hi, lo;
            split_f16x2(w.x, w.y, hi, lo);
            *reinterpret_cast<uint32_t *>(dst + slot * 128) = hi;
            *reinterpret_cast<uint32_t *>(dst + kStcA2Bytes + slot * 128) = lo;
        }
        fence_proxy_async();
        tc_fence_before();
    };
    auto issue_m2 = [&](int i) {
        tc_fence_after();
        const uint32_t d = tm_d2 + 192u * (uint32_t)(i & 1);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const uint64_t a_hi = tc_desc(sm_base + kStcOffA2 + 4096u * s, 2048u, 128u);
            const uint64_t a_lo = tc_desc(sm_base + kStcOffA2 + kStcA2Bytes + 4096u * s, 2048u, 128u);
            const uint64_t bd = tc_desc(sm_base + kStcOffB2 + 6144u * s, 3072u, 128u);
            tc_mma(d, a_hi, bd, kI192, s);
            tc_mma(d, a_lo, bd, kI96, 1u);
        }
        tc_commit(smem_u32(s_bar + 3 + (i & 1)));
    };

    // ---- P3: D2 -> magnitudes [bin][frame]
    auto phase3 = [&](int i) {
        const int quad = warp & 3, jq = warp >> 2;
        const int slot = 4 * quad + (lane >> 3), t = lane & 7;
        const uint32_t ta = tm_d2 + 192u * (uint32_t)(i & 1) + ((uint32_t)(32 * quad) << 16) + 12u * jq;
        float a[12], l[12];
        tmem_ld8(ta, a), tmem_ld4(ta + 8, a + 8), tmem_ld8(ta + 96, l), tmem_ld4(ta + 104, l + 8);
        if (quad == 0) {  // rows 0..7 are the packed rows: their results are in the B' columns
            float a2[12], l2[12];
            tmem_ld8(ta + 48, a2), tmem_ld4(ta + 56, a2 + 8), tmem_ld8(ta + 144, l2), tmem_ld4(ta + 152, l2 + 8);
            tmem_ld_wait();
            if (lane < 8) {
#pragma unroll
                for (int c = 0; c < 12; ++c) a[c] = a2[c], l[c] = l2[c];
            }
        } else {
            tmem_ld_wait();
        }
#pragma unroll
        for (int c = 0; c < 12; ++c) a[c] += l[c];
        if (p.dbg_d2 && first == 0 && i == 0) {
            float *d = p.dbg_d2 + (slot * 8 + t) * 48 + 12 * jq;
#pragma unroll
            for (int c = 0; c < 12; ++c) d[c] = a[c];
        }
        const float descale = s_descale[i & 3];
        const float eps = p.k.mag_eps > 0.f ? (p.k.mag_eps / descale) / descale : 0.f;
        const int off2 = slot ? 32 - slot : 16;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const int j = 6 * jq + c;
            const int bin = jq < 2 ? slot + 32 * j : off2 + 32 * (23 - j);
            s_mag[(kStcMagPad + bin) * kStcGroup + t] = magnitude<kPower>(a[2 * c], a[2 * c + 1], eps);
        }
        tc_fence_before();
    };

    // ---- P4: banded mel on the magnitude tile, log epilogue, stores
    auto phase4 = [&](int i) {
        long long b; int t0, s_first;
        batch_geom(i, b, t0, s_first);
        const int r = lane >> 3, t = lane & 7;
        const float descale = s_descale[i & 3];
        if (p.dbg_mag) {
            for (int e = tid; e < kStcBins * kStcGroup; e += kStcThreads) {
                const int bin = e >> 3, tt = e & 7;
                if (t0 + tt < p.k.T) {
                    float m = s_mag[(kStcMagPad + bin) * kStcGroup + tt] * descale;
                    if (kPower == 2) m *= descale;
                    p.dbg_mag[((long long)b * kStcBins + bin) * p.k.T + t0 + tt] = m;
                }
            }
        }
#pragma unroll 1
        for (int s = 0; s < kStcMelSlots; ++s) {
            const int g = s_grp[warp * kStcMelSlots + s];
            if (g < 0) break;
            const int2 ent = s_ent[g * 4 + r];
            const int len = s_glen[g];
            const float *w = s_melw + s_gwoff[g] + r;
            const float *mg = s_mag + (kStcMagPad + ent.y) * kStcGroup + t;
            float acc0 = 0.f, acc1 = 0.f;
            int q = 0;
#pragma unroll 4
            for (; q + 1 < len; q += 2) {
                acc0 = fmaf(w[4 * q], mg[kStcGroup * q], acc0);
                acc1 = fmaf(w[4 * q + 4], mg[kStcGroup * q + kStcGroup], acc1);
            }
            if (q < len) acc0 = fmaf(w[4 * q], mg[kStcGroup * q], acc0);
            float x = (acc0 + acc1) * descale;
            if (kPower == 2) x *= descale;
            if (ent.x >= 0 && t0 + t < p.k.T)
                p.k.out_mel[((long long)b * p.k.n_mels + ent.x) * p.k.T + t0 + t] = epilogue(x, p.k);
        }
    };

#ifdef B200MEL_STC_TIMING
    // per-CTA phase accounting (tools/tc_bench.py --timing): thread 0's clock at the phase boundaries, summed over CTAs
    long long stc_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long stc_mark = clock64();
#define STC_MARK(i)                              \
    do {                                         \
        const long long now_ = clock64();        \
        stc_acc[i] += now_ - stc_mark;           \
        stc_mark = now_;                         \
    } while (0)
#else
#define STC_MARK(i) \
    do {            \
    } while (0)
#endif
    // ------------------------------------------------------------------ the pipelined batch loop
    if (n_mine > 0) {
        if (tid == 0) {
            issue_tma(0);
            if (n_mine > 1) issue_tma(1);
        }
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        phase1(0);
        __syncthreads();
        if (tid == 0) {
            issue_m1();
            if (n_mine > 2) issue_tma(2);
        }
        STC_MARK(0);
        for (int i = 0; i < n_mine; ++i) {
            mbar_wait(smem_u32(s_bar + 2), (uint32_t)i & 1u);                                 // D1(i) complete, S free
            STC_MARK(1);
            if (i > 0) mbar_wait(smem_u32(s_bar + 3 + ((i - 1) & 1)), (uint32_t)((i - 1) >> 1) & 1u);  // M2(i-1) done: A2 free
            tc_fence_after();
            STC_MARK(2);
            phase2(i);
            __syncthreads();
            STC_MARK(3);
            if (tid == 0) issue_m2(i);
            STC_MARK(4);
            if (i + 1 < n_mine) {
                phase1(i + 1);
                __syncthreads();
                STC_MARK(5);
                if (tid == 0) {
                    issue_m1();
                    if (i + 3 < n_mine) issue_tma(i + 3);
                }
                STC_MARK(6);
            }
            if (i > 0) {
                phase3(i - 1);
                __syncthreads();
                STC_MARK(7);
                phase4(i - 1);
                STC_MARK(8);
            }
        }
        mbar_wait(smem_u32(s_bar + 3 + ((n_mine - 1) & 1)), (uint32_t)((n_mine - 1) >> 1) & 1u);
        tc_fence_after();
        __syncthreads();   // P4(n-2) readers of the magnitude tile are done
        phase3(n_mine - 1);
        __syncthreads();
        phase4(n_mine - 1);
    } else {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    // ------------------------------------------------------------------ teardown
#ifdef B200MEL_STC_TIMING
    STC_MARK(9);
    if (tid == 0 && p.k.dbg)
        for (int i = 0; i < 10; ++i) atomicAdd(reinterpret_cast<unsigned long long *>(p.k.dbg) + i, (unsigned long long)stc_acc[i]);
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace b200mel
