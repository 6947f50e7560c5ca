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

constexpr int kStcTwWarps = 8, kStcSpecWarps = 4, kStcEdgeWarps = 4, kStcMelWarps = 4;
constexpr int kStcWorkWarps = kStcTwWarps + kStcSpecWarps + kStcEdgeWarps + kStcMelWarps;   // 20
constexpr int kStcThreads = (kStcWorkWarps + 1) * 32;                        // + the issuer warp
constexpr int kStcGroup = 8;                         // frames per batch
constexpr int kStcHop = 256, kStcNfft = 1024;
constexpr int kStcSpan = (kStcGroup - 1) * kStcHop + kStcNfft;  // 2816 samples per batch
constexpr int kStcQ = kStcSpan / 32;                 // 88 fp16 per Hankel row
constexpr int kStcPitch = kStcQ * 2;                 // 176 bytes = 11 x 16: odd multiple of 16, so 8 rows hit 8 bank groups
constexpr int kStcBins = 384;                        // spectrum bins produced
constexpr int kStcMagPad = 4;                        // zero rows in front of the magnitude tile (mel windows may start at -3)
constexpr int kStcShift = 6;                         // stage-2 operands are scaled by 2^-6
constexpr int kStcMaxMels = kStcMelWarps * 32;       // one mel row per thread of the mel warps
constexpr int kStcOutPitch = 12;                     // floats per row of a mel warp's transpose tile (8 used)

// shared-memory carve-up (bytes); S, A2 and the magnitude tile are double-buffered
constexpr int kStcStageBytes = 11392;                // (2816 + 8) floats rounded to 128
constexpr int kStcOffStage = 0;
constexpr int kStcOffHank = 2 * kStcStageBytes;                  // per buffer: hi limb, then lo limb
constexpr int kStcHankBytes = 32 * kStcPitch;                    // 5632
constexpr int kStcOffB1 = kStcOffHank + 4 * kStcHankBytes;       // [4 k chunks][64 n][8 k] fp16
constexpr int kStcB1Bytes = 4096;
constexpr int kStcOffA2 = kStcOffB1 + kStcB1Bytes;               // per buffer: hi tile, lo tile: [8 k chunks][16 slots][8 frames][8 k]
constexpr int kStcA2Bytes = 16384;
constexpr int kStcOffB2 = kStcOffA2 + 4 * kStcA2Bytes;           // [8 k chunks][192 n][8 k] fp16
constexpr int kStcB2Bytes = 24576;
constexpr int kStcOffTw = kStcOffB2 + kStcB2Bytes;               // float2 [17][32]
constexpr int kStcTwBytes = 17 * 32 * 8;
constexpr int kStcOffMag = kStcOffTw + kStcTwBytes;              // 2 x float [4 + 384][8]
constexpr int kStcMagBytes = (kStcMagPad + kStcBins) * kStcGroup * 4;
constexpr int kStcOffMisc = kStcOffMag + 2 * kStcMagBytes;       // mbarriers, TMEM address, per-warp maxima, descale ring
constexpr int kStcMiscBytes = 384;
constexpr int kStcOffOut = kStcOffMisc + kStcMiscBytes;          // per mel warp: float [32 rows][kStcOutPitch] transpose tile
constexpr int kStcOutBytes = 32 * kStcOutPitch * 4;
constexpr int kStcOffMel = kStcOffOut + kStcMelWarps * kStcOutBytes;  // mel schedule (header + weights), size from the plan
constexpr int kStcMelHeader = 2 * kStcMelWarps * 4 + kStcMaxMels * 8;  // trip[4] | woff[4] | ent[128] {m, lo}

// mbarriers (index into the array at kStcOffMisc)
enum {
    kBarStageFull = 0,    // [2] TMA landed                       -> edge warps
    kBarStageEmpty = 2,   // [2] edge warps read the stage        -> issuer (next TMA into it)
    kBarSFull = 4,        // [2] Hankel buffer written            -> issuer (M1)
    kBarD1Full = 6,       // [2] M1 complete (commit)             -> twiddle warps; also: Hankel buffer free -> edge warps
    kBarD1Empty = 8,      // [2] twiddle warps read D1            -> issuer (M1 two batches on)
    kBarA2Full = 10,      // [2] stage-2 operand written          -> issuer (M2)
    kBarD2Full = 12,      // [2] M2 complete (commit)             -> spectrum warps; also: A2 buffer free -> twiddle warps
    kBarD2Empty = 14,     // [1] spectrum warps read D2           -> issuer (next M2)
    kBarMagFull = 15,     // [2] magnitude tile written           -> mel warps
    kBarMagEmpty = 17,    // [2] mel warps read the tile          -> spectrum warps (two batches on)
    kStcNumBars = 19
};

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
    const float2 d = __fadd2_rn(make_float2(x, y), make_float2(-hf.x, -hf.y));
    const __half2 l = __floats2half2_rn(d.x, d.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0u;
}
// mbarrier wait that parks the warp in hardware (suspend-time hint, as CUTLASS' ClusterBarrier::wait does) instead of
// spinning: the role warps of this kernel spend most of their time waiting for each other, and a spinning warp
// competes for issue slots with the working warps of its scheduler (measured: 45 % of all executed instructions
// were try_wait loops before this).
__device__ __forceinline__ void mbar_park(uint32_t bar, uint32_t parity) {
    uint32_t done, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(0x989680u)
            : "memory");
        if (!done && ++spins > (1u << 20)) __trap();
    } while (!done);
}
// one arrival per warp: every lane's prior work is ordered before it by the __syncwarp
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// Per-role cycle accounting (-DB200MEL_STC_TIMING, tools/tc_bench.py timing): lane 0 of one warp per role adds its
// clock64() deltas to dbg[role * 8 + slot]; roles: 0 issuer, 1 twiddle, 2 spectrum, 3 edge.
#ifdef B200MEL_STC_TIMING
#define STC_TIMER() long long stc_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, stc_mark = clock64()
#define STC_MARK(i)                        \
    do {                                   \
        const long long now_ = clock64();  \
        stc_acc[i] += now_ - stc_mark;     \
        stc_mark = now_;                   \
    } while (0)
#define STC_FLUSH(role)                                                                                        \
    do {                                                                                                       \
        if (lane == 0 && p.k.dbg)                                                                              \
            for (int i_ = 0; i_ < 8; ++i_)                                                                     \
                atomicAdd(reinterpret_cast<unsigned long long *>(p.k.dbg) + (role) * 8 + i_, (unsigned long long)stc_acc[i_]); \
    } while (0)
#else
#define STC_TIMER() do { } while (0)
#define STC_MARK(i) do { } while (0)
#define STC_FLUSH(role) do { } while (0)
#endif

template <int kPower>
__global__ void __launch_bounds__(kStcThreads, 1) stft_tc_kernel(const StcParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float2 *s_tw = reinterpret_cast<const float2 *>(smem + kStcOffTw);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + kStcOffMisc);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + kStcOffMisc + 160);
    uint32_t *s_wmax = reinterpret_cast<uint32_t *>(smem + kStcOffMisc + 176);  // [2][4]
    float *s_descale = reinterpret_cast<float *>(smem + kStcOffMisc + 208);     // [8]
    int *s_geo = reinterpret_cast<int *>(smem + kStcOffMisc + 256);             // [2][8] per stage buffer: what the issuer staged
    const int *s_trip = reinterpret_cast<const int *>(smem + kStcOffMel);                        // [kStcMelWarps]
    const int *s_woff = s_trip + kStcMelWarps;                                                    // [kStcMelWarps]
    const int2 *s_ent = reinterpret_cast<const int2 *>(s_woff + kStcMelWarps);                    // [kStcMaxMels] {m, lo}
    const float *s_melw = reinterpret_cast<const float *>(smem + kStcOffMel + kStcMelHeader);     // per warp [trip][32 lanes]
    const uint32_t sm_base = smem_u32(smem);
    auto bar = [&](int i) { return smem_u32(s_bar + i); };

    // ------------------------------------------------------------------ one-time setup (all 17 warps)
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
        float *mg = reinterpret_cast<float *>(smem + kStcOffMag);
        for (int i = tid; i < kStcMagPad * kStcGroup; i += kStcThreads) mg[i] = 0.f, mg[kStcMagBytes / 4 + i] = 0.f;
    }
    if (tid == 0) {
        static_assert(kStcNumBars * 8 <= 160, "mbarriers overlap the scalars behind them");
        const int counts[kStcNumBars] = {1, 1, kStcEdgeWarps, kStcEdgeWarps, kStcEdgeWarps, kStcEdgeWarps, 1, 1, kStcTwWarps, kStcTwWarps,
                                         kStcTwWarps, kStcTwWarps, 1, 1, kStcSpecWarps, kStcSpecWarps, kStcSpecWarps, kStcMelWarps, kStcMelWarps};
        static_assert(kBarMagEmpty + 2 == kStcNumBars, "barrier table out of date");
        for (int i = 0; i < kStcNumBars; ++i) mbar_init(bar(i), counts[i]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kStcWorkWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t tm_d1 = tmem;            // 2 buffers x (2 tiles x 64 columns)
    const uint32_t tm_d2 = tmem + 256;      // 192 columns

    // instruction descriptors: D fp32, A / B fp16, both K-major, M = 128
    constexpr uint32_t kIdesc = (1u << 4) | (8u << 24);
    constexpr uint32_t kI64 = kIdesc | (8u << 17), kI32 = kIdesc | (4u << 17), kI192 = kIdesc | (24u << 17), kI96 = kIdesc | (12u << 17);

    const unsigned nb_total = (unsigned)p.n_batches;   // < 2^31 (checked by the host)
    const unsigned first = blockIdx.x, stride = gridDim.x;
    const int n_mine = first < nb_total ? (int)((nb_total - first + stride - 1) / stride) : 0;
    const unsigned Gc = (unsigned)p.groups_per_clip;

    asm volatile("griddepcontrol.wait;" ::: "memory");  // PDL: nothing above touched caller memory
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // geometry of my batch number j
    auto batch_geom = [&](int j, unsigned &b, int &t0, int &s_first) {
        const unsigned g = first + (unsigned)j * stride;
        b = g / Gc;
        t0 = (int)(g - b * Gc) * kStcGroup;
        s_first = t0 * kStcHop - p.k.pad;
    };

    if (warp == kStcWorkWarps) {
        // ================================================================== issuer: TMA + tcgen05.mma, one thread
        if (lane == 0 && n_mine > 0) {
            int j_tma = 0, j_m1 = 0, j_m2 = 0;
            uint32_t idle = 0;
            STC_TIMER();
            while (j_m2 < n_mine) {
                bool progress = false;
                STC_MARK(0);
                if (j_tma < n_mine && (j_tma < 2 || mbar_test(bar(kBarStageEmpty + (j_tma & 1)), (uint32_t)((j_tma >> 1) - 1) & 1u))) {
                    unsigned b; int t0, s_first;
                    batch_geom(j_tma, b, t0, s_first);
                    const CopyGeom g = copy_geom(p.k, (long long)b, s_first, kStcSpan, p.k.L);
                    int *geo = s_geo + (j_tma & 1) * 8;   // read by the edge warps once the copy has landed
                    geo[0] = (int)b, geo[1] = s_first, geo[2] = g.delta, geo[3] = g.c_lo, geo[4] = g.c_hi, geo[5] = g.patch ? 1 : 0;
                    issue_copy(g, sm_base + kStcOffStage + (j_tma & 1) * kStcStageBytes, bar(kBarStageFull + (j_tma & 1)));
                    ++j_tma;
                    progress = true;
                    STC_MARK(1);
                }
                // the older batch first: M2(j) before M1(j + 2)
                if (j_m2 < j_m1 && mbar_test(bar(kBarA2Full + (j_m2 & 1)), (uint32_t)(j_m2 >> 1) & 1u) &&
                    (j_m2 == 0 || mbar_test(bar(kBarD2Empty), (uint32_t)(j_m2 - 1) & 1u))) {
                    tc_fence_after();
                    const uint32_t a2 = sm_base + kStcOffA2 + (uint32_t)(j_m2 & 1) * 2u * kStcA2Bytes;
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const uint64_t a_hi = tc_desc(a2 + 4096u * s, 2048u, 128u);
                        const uint64_t a_lo = tc_desc(a2 + kStcA2Bytes + 4096u * s, 2048u, 128u);
                        const uint64_t bd = tc_desc(sm_base + kStcOffB2 + 6144u * s, 3072u, 128u);
                        tc_mma(tm_d2, a_hi, bd, kI192, s);
                        tc_mma(tm_d2, a_lo, bd, kI96, 1u);
                    }
                    tc_commit(bar(kBarD2Full + (j_m2 & 1)));
                    ++j_m2;
                    progress = true;
                    STC_MARK(2);
                }
                if (j_m1 < n_mine && mbar_test(bar(kBarSFull + (j_m1 & 1)), (uint32_t)(j_m1 >> 1) & 1u) &&
                    (j_m1 < 2 || mbar_test(bar(kBarD1Empty + (j_m1 & 1)), (uint32_t)((j_m1 >> 1) - 1) & 1u))) {
                    tc_fence_after();
                    const uint32_t hk = sm_base + kStcOffHank + (uint32_t)(j_m1 & 1) * 2u * kStcHankBytes;
                    const uint32_t d1 = tm_d1 + 128u * (uint32_t)(j_m1 & 1);
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
                            const uint32_t a_off = (uint32_t)(16 * h * kStcPitch + 32 * s);
                            const uint64_t a_hi = tc_desc(hk + a_off, 16u, (uint32_t)kStcPitch);
                            const uint64_t a_lo = tc_desc(hk + kStcHankBytes + a_off, 16u, (uint32_t)kStcPitch);
                            const uint64_t bd = tc_desc(sm_base + kStcOffB1 + 2048u * s, 1024u, 128u);
                            tc_mma(d1 + 64u * h, a_hi, bd, kI64, s);
                            tc_mma(d1 + 64u * h, a_lo, bd, kI32, 1u);
                        }
                    tc_commit(bar(kBarD1Full + (j_m1 & 1)));
                    ++j_m1;
                    progress = true;
                    STC_MARK(3);
                }
                if (progress) idle = 0;
                else {
                    __nanosleep(64);                      // nothing ready: stay out of the working warps' issue slots
                    if (++idle > (1u << 24)) __trap();    // protocol bug: a CUDA error instead of a hung GPU
                }
            }
            STC_FLUSH(0);
        }
    } else if (warp < kStcTwWarps) {
        // ================================================================== twiddle warps: P2
        // D1 -> registers, hi + lo halves, twiddle, Hann 3-tap over k1, fp16 limbs -> stage-2 operand
        const int quad = warp & 3, h = warp >> 2;
        const int n2 = 16 * h + 4 * quad + (lane >> 3), t = lane & 7;
        const float2 *tw = s_tw + n2;
        const float S = 1.0f / (float)(1 << kStcShift);
        STC_TIMER();
        for (int j = 0; j < n_mine; ++j) {
            STC_MARK(0);
            mbar_park(bar(kBarD1Full + (j & 1)), (uint32_t)(j >> 1) & 1u);
            STC_MARK(1);
            if (j >= 2) mbar_park(bar(kBarD2Full + (j & 1)), (uint32_t)((j - 2) >> 1) & 1u);  // M2(j - 2) done: A2 buffer free
            tc_fence_after();
            STC_MARK(2);
            const uint32_t ta = tm_d1 + 128u * (uint32_t)(j & 1) + ((uint32_t)(32 * quad) << 16) + 64u * h;
            unsigned char *dst = smem + kStcOffA2 + (j & 1) * 2 * kStcA2Bytes + t * 16 + (n2 >> 2) * 2048 + (n2 & 3) * 4;
            float a[20], b2[20];
            {   // a: columns 0..19 (A0, A16, A1..A9); b2: columns 16..31 (A8..A15) and 0..3 (A0, A16, A1); each as hi part + lo part
                float l[20], l2[20];
                tmem_ld16(ta, a), tmem_ld4(ta + 16, a + 16), tmem_ld16(ta + 32, l), tmem_ld4(ta + 48, l + 16);
                tmem_ld16(ta + 16, b2), tmem_ld4(ta, b2 + 16), tmem_ld16(ta + 48, l2), tmem_ld4(ta + 32, l2 + 16);
                tmem_ld_wait();
                tc_fence_before();
                warp_arrive(bar(kBarD1Empty + (j & 1)), lane);
#pragma unroll
                for (int c = 0; c < 20; c += 2) {
                    const float2 u = __fadd2_rn(make_float2(a[c], a[c + 1]), make_float2(l[c], l[c + 1]));
                    const float2 v = __fadd2_rn(make_float2(b2[c], b2[c + 1]), make_float2(l2[c], l2[c + 1]));
                    a[c] = u.x, a[c + 1] = u.y, b2[c] = v.x, b2[c + 1] = v.y;
                }
            }
            if (p.dbg_d1 && first == 0 && j == 0) {
                float *d = p.dbg_d1 + (n2 * 8 + t) * 32;
#pragma unroll
                for (int c = 0; c < 20; ++c) d[c] = a[c];
#pragma unroll
                for (int c = 0; c < 16; ++c) d[16 + c] = b2[c];
            }
            {   // slots 1..8 from A'[0..9]
                float2 ap[10];
                ap[0] = make_float2(a[0] * S, 0.f);
#pragma unroll
                for (int k = 1; k < 10; ++k) ap[k] = cmul(make_float2(a[2 * k], a[2 * k + 1]), tw[32 * k]);
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    const float2 sum = __fadd2_rn(ap[o], ap[o + 2]);
                    const float2 w = __ffma2_rn(sum, make_float2(-0.25f, -0.25f), __fmul2_rn(ap[o + 1], make_float2(0.5f, 0.5f)));
                    uint32_t hi, lo;
                    split_f16x2(w.x, w.y, hi, lo);
                    *reinterpret_cast<uint32_t *>(dst + (o + 1) * 128) = hi;
                    *reinterpret_cast<uint32_t *>(dst + kStcA2Bytes + (o + 1) * 128) = lo;
                }
            }
            {   // slots 9..15 from A'[8..16], and the packed slot 0 = (Aw'[0], r16)
                float2 ap[9];
#pragma unroll
                for (int k = 8; k < 16; ++k) ap[k - 8] = cmul(make_float2(b2[2 * (k - 8)], b2[2 * (k - 8) + 1]), tw[32 * k]);
                const float2 w16 = tw[32 * 16], w1 = tw[32];
                ap[8] = make_float2(b2[17] * w16.x, b2[17] * w16.y);                 // A'[16] = W^(16 n2) A[16], A[16] real
                const float2 ap1 = cmul(make_float2(b2[18], b2[19]), w1);            // A'[1]
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    float2 w;
                    if (o < 7) {
                        const float2 sum = __fadd2_rn(ap[o], ap[o + 2]);
                        w = __ffma2_rn(sum, make_float2(-0.25f, -0.25f), __fmul2_rn(ap[o + 1], make_float2(0.5f, 0.5f)));
                    } else {
                        w.x = 0.5f * S * b2[16] - 0.5f * ap1.x;                                 // Aw'[0] (real)
                        w.y = 0.5f * S * b2[17] - 0.5f * fmaf(w1.x, b2[14], w1.y * b2[15]);     // r16: Aw'[16] = W64^n2 r16
                    }
                    const int slot = o < 7 ? o + 9 : 0;
                    uint32_t hi, lo;
                    split_f16x2(w.x, w.y, hi, lo);
                    *reinterpret_cast<uint32_t *>(dst + slot * 128) = hi;
                    *reinterpret_cast<uint32_t *>(dst + kStcA2Bytes + slot * 128) = lo;
                }
            }
            fence_proxy_async();
            warp_arrive(bar(kBarA2Full + (j & 1)), lane);
            STC_MARK(3);
        }
        if (warp == 0) STC_FLUSH(1);
    } else if (warp < kStcTwWarps + kStcSpecWarps) {
        // ================================================================== spectrum warps: P3
        const int quad = warp & 3;
        const int slot = 4 * quad + (lane >> 3), t = lane & 7;
        const int off2 = slot ? 32 - slot : 16;
        STC_TIMER();
        for (int j = 0; j < n_mine; ++j) {
            STC_MARK(0);
            mbar_park(bar(kBarD2Full + (j & 1)), (uint32_t)(j >> 1) & 1u);
            STC_MARK(1);
            if (j >= 2) mbar_park(bar(kBarMagEmpty + (j & 1)), (uint32_t)((j - 2) >> 1) & 1u);
            tc_fence_after();
            STC_MARK(2);
            const uint32_t ta = tm_d2 + ((uint32_t)(32 * quad) << 16);
            float *tile = reinterpret_cast<float *>(smem + kStcOffMag + (j & 1) * kStcMagBytes);
            const float descale = s_descale[j & 7];
            const float eps = p.k.mag_eps > 0.f ? (p.k.mag_eps / descale) / descale : 0.f;
#pragma unroll
            for (int jq = 0; jq < 4; ++jq) {   // 6 complex outputs at a time: columns 12 jq .. 12 jq + 11 of the 48
                float a[12], l[12];
                const uint32_t col = 12u * jq;
                tmem_ld8(ta + col, a), tmem_ld4(ta + col + 8, a + 8), tmem_ld8(ta + 96 + col, l), tmem_ld4(ta + 96 + col + 8, l + 8);
                if (quad == 0) {  // rows 0..7 (lanes 0..7) are the packed rows: their results sit in the B' columns (+48)
                    float a2[12], l2[12];
                    tmem_ld8(ta + 48 + col, a2), tmem_ld4(ta + 48 + col + 8, a2 + 8);
                    tmem_ld8(ta + 144 + col, l2), tmem_ld4(ta + 144 + col + 8, l2 + 8);
                    tmem_ld_wait();
                    if (lane < 8) {
#pragma unroll
                        for (int c = 0; c < 12; ++c) a[c] = a2[c], l[c] = l2[c];
                    }
                } else {
                    tmem_ld_wait();
                }
                if (jq == 3) {
                    tc_fence_before();
                    warp_arrive(bar(kBarD2Empty), lane);
                }
#pragma unroll
                for (int c = 0; c < 12; c += 2) {
                    const float2 u = __fadd2_rn(make_float2(a[c], a[c + 1]), make_float2(l[c], l[c + 1]));
                    a[c] = u.x, a[c + 1] = u.y;
                }
                if (p.dbg_d2 && first == 0 && j == 0) {
                    float *d = p.dbg_d2 + (slot * 8 + t) * 48 + 12 * jq;
#pragma unroll
                    for (int c = 0; c < 12; ++c) d[c] = a[c];
                }
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const int jj = 6 * jq + c;
                    const int bin = jq < 2 ? slot + 32 * jj : off2 + 32 * (23 - jj);
                    tile[(kStcMagPad + bin) * kStcGroup + t] = magnitude<kPower>(a[2 * c], a[2 * c + 1], eps);
                }
            }
            warp_arrive(bar(kBarMagFull + (j & 1)), lane);
            STC_MARK(3);
        }
        if (warp == kStcTwWarps) STC_FLUSH(2);
    } else if (warp < kStcTwWarps + kStcSpecWarps + kStcEdgeWarps) {
        // ================================================================== edge warps: P1
        // staged fp32 samples -> max |x| -> power-of-two scale -> fp16 limbs in the Hankel layout
        const int e = warp - kStcTwWarps - kStcSpecWarps;
        STC_TIMER();
        for (int j = 0; j < n_mine; ++j) {
            STC_MARK(0);
            mbar_park(bar(kBarStageFull + (j & 1)), (uint32_t)(j >> 1) & 1u);
            STC_MARK(1);
            const int *geo = s_geo + (j & 1) * 8;
            const int s_first = geo[1], delta = geo[2];
            const float *st = reinterpret_cast<const float *>(smem + kStcOffStage + (j & 1) * kStcStageBytes) + delta - s_first;
            float v[3][8];   // q octets e, e + 4, e + 8 (the last one only for e < 3)
            if (geo[5] == 0) {
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    if (e + 4 * r < 11) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[r][i] = st[s_first + 256 * (e + 4 * r) + lane + 32 * i];
                    }
            } else {   // edge of the clip: reflected samples (and floats cut off by the tensor-edge clamp) come from global memory
                const int c_lo = geo[3], c_hi = geo[4];
                const float *row = p.k.wav + (long long)geo[0] * p.k.row_stride;
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    if (e + 4 * r < 11) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int sidx = s_first + 256 * (e + 4 * r) + lane + 32 * i;
                            v[r][i] = (sidx >= c_lo && sidx < c_hi) ? st[sidx] : __ldg(row + reflect_index(sidx, p.k.L));
                        }
                    }
            }
            // max |x| on the bit patterns: NaN / Inf compare above every finite value and switch the scaling off
            uint32_t m = 0u;
#pragma unroll
            for (int r = 0; r < 3; ++r)
                if (e + 4 * r < 11) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) m = max(m, __float_as_uint(v[r][i]) & 0x7fffffffu);
                }
            warp_arrive(bar(kBarStageEmpty + (j & 1)), lane);
            const uint32_t wm = __reduce_max_sync(0xffffffffu, m);
            if (lane == 0) s_wmax[(j & 1) * 4 + e] = wm;
            STC_MARK(2);
            asm volatile("bar.sync 1, %0;" ::"n"(kStcEdgeWarps * 32) : "memory");
            STC_MARK(3);
            const uint4 all = *reinterpret_cast<const uint4 *>(s_wmax + (j & 1) * 4);
            const uint32_t mx = max(max(all.x, all.y), max(all.z, all.w));
            // power-of-two scale: max |x| * sc in [2^13, 2^14); 1 for silence / non-finite input
            int sexp = 127;
            if (mx != 0u && mx < 0x7f800000u) {
                sexp = 267 - (int)(mx >> 23);
                sexp = min(max(sexp, 1), 254);
                if (p.k.mag_eps > 0.f) sexp = min(sexp, 127 + 40);
            }
            const float sc = __uint_as_float((uint32_t)sexp << 23);
            if (e == 0 && lane == 0) s_descale[j & 7] = __uint_as_float((uint32_t)(260 - sexp) << 23);  // 2^6 / sc
            if (j >= 2) mbar_park(bar(kBarD1Full + (j & 1)), (uint32_t)((j - 2) >> 1) & 1u);  // M1(j - 2) done: Hankel buffer free
            STC_MARK(4);
            unsigned char *hk = smem + kStcOffHank + (j & 1) * 2 * kStcHankBytes + lane * kStcPitch;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int oct = e + 4 * r;
                if (oct < 11) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) split_f16x2(v[r][2 * i] * sc, v[r][2 * i + 1] * sc, hi[i], lo[i]);
                    *reinterpret_cast<uint4 *>(hk + oct * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4 *>(hk + kStcHankBytes + oct * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            fence_proxy_async();
            warp_arrive(bar(kBarSFull + (j & 1)), lane);
            STC_MARK(5);
        }
        if (e == 0) STC_FLUSH(3);
    } else {
        // ================================================================== mel warps: P4
        // one mel row per thread, all 8 frames of the batch in registers (a weight is loaded once per 8 FMAs); the
        // results are transposed through a per-warp tile so that 8 lanes store the 32 contiguous bytes of a row
        const int mw = warp - kStcTwWarps - kStcSpecWarps - kStcEdgeWarps;
        const int2 ent = s_ent[mw * 32 + lane];                 // {mel row (-1: idle), first bin of the read window}
        const int trip = s_trip[mw];
        const float *wgt = s_melw + s_woff[mw] + lane;
        float *otile = reinterpret_cast<float *>(smem + kStcOffOut + mw * kStcOutBytes);
        const int r4 = lane >> 3, t = lane & 7;
        STC_TIMER();
        for (int j = 0; j < n_mine; ++j) {
            STC_MARK(0);
            unsigned b; int t0, s_first;
            batch_geom(j, b, t0, s_first);
            mbar_park(bar(kBarMagFull + (j & 1)), (uint32_t)(j >> 1) & 1u);
            STC_MARK(1);
            const float4 *tile = reinterpret_cast<const float4 *>(smem + kStcOffMag + (j & 1) * kStcMagBytes) + (kStcMagPad + ent.y) * 2;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
#pragma unroll 4
            for (int i = 0; i < trip; ++i) {
                const float w = wgt[32 * i];
                const float4 m0 = tile[2 * i], m1 = tile[2 * i + 1];
                a0.x = fmaf(w, m0.x, a0.x), a0.y = fmaf(w, m0.y, a0.y), a0.z = fmaf(w, m0.z, a0.z), a0.w = fmaf(w, m0.w, a0.w);
                a1.x = fmaf(w, m1.x, a1.x), a1.y = fmaf(w, m1.y, a1.y), a1.z = fmaf(w, m1.z, a1.z), a1.w = fmaf(w, m1.w, a1.w);
            }
            warp_arrive(bar(kBarMagEmpty + (j & 1)), lane);
            if (p.dbg_mag && mw == 0) {   // debug tap: the whole magnitude tile of this batch
                const float *tl = reinterpret_cast<const float *>(smem + kStcOffMag + (j & 1) * kStcMagBytes);
                const float ds = s_descale[j & 7];
                for (int el = lane; el < kStcBins * kStcGroup; el += 32) {
                    const int bin = el >> 3, tt = el & 7;
                    if (t0 + tt < p.k.T) {
                        float m = tl[(kStcMagPad + bin) * kStcGroup + tt] * ds;
                        if (kPower == 2) m *= ds;
                        p.dbg_mag[((long long)b * kStcBins + bin) * p.k.T + t0 + tt] = m;
                    }
                }
            }
            *reinterpret_cast<float4 *>(otile + lane * kStcOutPitch) = a0;
            *reinterpret_cast<float4 *>(otile + lane * kStcOutPitch + 4) = a1;
            __syncwarp();
            const float descale = s_descale[j & 7];
            float *obase = p.k.out_mel + (long long)b * p.k.n_mels * p.k.T + t0 + t;
            const bool t_ok = t0 + t < p.k.T;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int row = 4 * it + r4;
                const int m = __shfl_sync(0xffffffffu, ent.x, row);
                float x = otile[row * kStcOutPitch + t] * descale;
                if (kPower == 2) x *= descale;
                if (m >= 0 && t_ok) obase[(long long)m * p.k.T] = epilogue(x, p.k);
            }
            __syncwarp();
            STC_MARK(2);
        }
        if (mw == 0) STC_FLUSH(4);
    }
    // ------------------------------------------------------------------ teardown
    tc_fence_before();
    __syncthreads();
    if (warp == kStcWorkWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace b200mel
