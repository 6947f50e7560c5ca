// mel_tc.cuh — the mel filterbank as a tcgen05 tensor-core GEMM with the accumulator in TMEM (sm_100a).
//
//     out[b, m, t] = epilogue( sum_k W[m, k] * mag[b, k, t] )          mag (B, F, T) fp32 -> out (B, n_mels, T) fp32
//
// This is `torch.matmul(self.mel_filter, magnitude)` + log + clamp of the reference (LogMelScale.forward,
// models/transforms.py:261-268; the same contraction as LogMelSpectrogram.forward :235-243) taken literally as the
// dense contraction the north star names, on magnitudes that already sit in HBM.
//
// GEMM shape per tile:  D[128 frames, N mels] += A[128 frames, 16 bins] * B[16 bins, N mels]   (UMMA M = 128, K = 16)
//   A = magnitudes.  (b, t) pairs are flattened into columns c = b T + t; a tile is 128 consecutive columns, so there
//       is no per-clip padding.  The tensor is (F, T) row-major per clip, i.e. frames are contiguous: the natural
//       UMMA layout is MN-major (M = frames).  fp32 -> 2 x bf16 happens on the way into shared memory:
//       x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits together; magnitudes and weights are >= 0,
//       so there is no cancellation and the error stays relative: ~3 x 2^-18, measured max 6e-6 on the log-mel value).
//   B = filterbank, K-major, pre-split into bf16 hi / lo on the host in the canonical core-matrix layout and held in
//       shared memory for the whole kernel.
//   D = fp32 accumulator in TMEM (128 lanes = frames, N columns = mels); three MMAs per K step:
//       A_hi B_hi + A_lo B_hi + A_hi B_lo   (lo x lo is below 2^-18 relative and is dropped).
//   Epilogue: tcgen05.ld (lane = frame), log / clamp / norm in registers, stores of 32 consecutive frames per warp
//       and mel row — full 128-byte lines.
//
// Pipeline: a stage is 32 bins (two K steps); two A buffers alternate; the MMAs of stage s run on the tensor pipe
// while every warp loads and converts stage s + 1.  One thread issues tcgen05.mma; tcgen05.commit arrives on an
// mbarrier per buffer, which is what the producers wait on before overwriting that buffer.
//
// Canonical (no-swizzle) UMMA layouts used, in bytes (see cute/arch/mma_sm100_desc.hpp in any CUTLASS tree):
//   A, MN-major:  elem(m, k) at (m % 8) * 2 + (m / 8) * SBO + (k % 8) * 16 + (k / 8) * LBO     SBO = 144, LBO = 2304
//                 (8 frames = 16 contiguous bytes; SBO is 144 rather than 128 so that the four 16-byte segments a warp
//                  writes per bin land in different banks)
//   B, K-major:   elem(n, k) at (n % 8) * 16 + (n / 8) * SBO + (k % 8) * 2 + (k / 8) * LBO      SBO = 128, LBO = (N / 8) * 128
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "logmel_kernel.cuh"

namespace b200mel {

constexpr int kTcThreads = 512;
constexpr int kTcTileM = 128;                 // frames per tile (UMMA M)
constexpr int kTcStageBins = 32;              // bins per pipeline stage (two K = 16 steps)
constexpr int kTcASbo = 144, kTcALbo = 16 * kTcASbo;  // 2304
constexpr int kTcAStep = 2 * kTcALbo;         // bytes of one K step of one A limb: 4608
constexpr int kTcALimb = 2 * kTcAStep;        // one limb (hi or lo) of a stage: 9216
constexpr int kTcABuf = 2 * kTcALimb;         // hi + lo of a stage: 18432
constexpr int kTcTmemCols = 128;              // accumulator columns allocated (power of two >= N)

struct TcParams {
    const float *mag;
    float *out;
    long long n_cols;   // B * T
    int T, F, n_mels;
    int n_pad;          // N: n_mels rounded up to 16 (<= 128)
    int k_steps;        // K steps of 16 bins that cover every non-zero filterbank column (even: a stage is two steps)
    const uint16_t *b_hi, *b_lo;  // [k_steps][2][n_pad / 8][8][8] bf16 bit patterns
    int b_bytes;        // bytes of one limb table = k_steps * 2 * (n_pad / 8) * 128
    int use_log;
    float ep_floor, ep_offset, log_scale, lo, hi, norm_scale, norm_bias;
};

__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           (1ull << 46);  // version 1 (Blackwell), base offset 0, SWIZZLE_NONE
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float tc_epilogue(float x, const TcParams &p) {
    float y = x;
    if (p.use_log) {
        const float t = fmaxf(x, p.ep_floor) + p.ep_offset;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(t));
        y *= p.log_scale;
    }
    y = fminf(fmaxf(y, p.lo), p.hi);
    return fmaf(y, p.norm_scale, p.norm_bias);
}

// Shared memory: B_hi | B_lo | A buffer 0 | A buffer 1 | 2 mbarriers | TMEM base address
__global__ void __launch_bounds__(kTcThreads, 1) mel_tc_kernel(const TcParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char *s_b = smem_raw;
    unsigned char *s_a = smem_raw + 2 * p.b_bytes;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_a + 2 * kTcABuf);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 2);

    // ---- one-time setup: filterbank limbs into shared memory, mbarriers, TMEM allocation
    for (int i = tid; i < p.b_bytes / 16; i += kTcThreads) {
        reinterpret_cast<int4 *>(s_b)[i] = __ldg(reinterpret_cast<const int4 *>(p.b_hi) + i);
        reinterpret_cast<int4 *>(s_b + p.b_bytes)[i] = __ldg(reinterpret_cast<const int4 *>(p.b_lo) + i);
    }
    if (tid == 0) {
        mbar_init(smem_u32(s_bar), 1);
        mbar_init(smem_u32(s_bar + 1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(kTcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();  // the filterbank limbs were written with generic stores; the tensor core reads them through the async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;

    // instruction descriptor: D fp32, A / B bf16, A MN-major, B K-major, N = n_pad, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(kTcTileM >> 4) << 24);
    const uint32_t b_step = 2u * (uint32_t)(p.n_pad / 8) * 128u;  // bytes of one K step of one B limb
    const uint32_t b_lbo = (uint32_t)(p.n_pad / 8) * 128u;

    // producer role of this thread: frame fq * 32 + lane of the tile, bins bg * 8 .. bg * 8 + 7 of the stage
    const int fq = warp & 3, bg = warp >> 2;
    const int fi = fq * 32 + lane;
    unsigned char *a_dst0 = s_a + (bg >> 1) * kTcAStep + (bg & 1) * kTcALbo + (fi >> 3) * kTcASbo + (fi & 7) * 2;

    const long long n_tiles = (p.n_cols + kTcTileM - 1) / kTcTileM;
    const int n_stages = p.k_steps / 2;
    unsigned g = 0;  // stages produced so far by this CTA (buffer = g & 1, use count of that buffer = g >> 1)

    // this thread's column of a tile and the 8 magnitudes of one stage (issued as 8 independent coalesced loads)
    struct Col { bool ok; long long b; int t; const float *src; };
    auto locate = [&](long long tile) {
        Col c;
        const long long col = tile * kTcTileM + fi;
        c.ok = tile < n_tiles && col < p.n_cols;
        c.b = c.ok ? col / p.T : 0;
        c.t = c.ok ? (int)(col - c.b * p.T) : 0;
        c.src = p.mag + (c.b * p.F) * (long long)p.T + c.t;
        return c;
    };
    auto fetch = [&](const Col &c, int s, float *v) {
        const int k0 = s * kTcStageBins + bg * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (c.ok && k0 + j < p.F) ? __ldg(c.src + (long long)(k0 + j) * p.T) : 0.f;
    };

    Col col = locate(blockIdx.x);
    float v[8];
    fetch(col, 0, v);  // software pipeline: the loads of stage s + 1 are in flight while stage s is converted and issued
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const bool col_ok = col.ok;
        const long long cb = col.b;
        const int ct = col.t;
        const Col next_col = locate(tile + gridDim.x);

        for (int s = 0; s < n_stages; ++s, ++g) {
            const unsigned buf = g & 1u, use = g >> 1;
            // ---- produce: fp32 -> bf16 hi / lo, MN-major stores
            __nv_bfloat16 h[8], l[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                h[j] = __float2bfloat16_rn(v[j]);
                l[j] = __float2bfloat16_rn(v[j] - __bfloat162float(h[j]));
            }
            if (s + 1 < n_stages) fetch(col, s + 1, v);
            else fetch(next_col, 0, v);
            if (use > 0) mbar_wait(smem_u32(s_bar + buf), (use - 1) & 1u);  // the MMAs that read this buffer are done
            unsigned char *dst = a_dst0 + buf * kTcABuf;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                *reinterpret_cast<__nv_bfloat16 *>(dst + j * 16) = h[j];
                *reinterpret_cast<__nv_bfloat16 *>(dst + kTcALimb + j * 16) = l[j];
            }
            fence_proxy_async();
            __syncthreads();
            // ---- one thread feeds the tensor core: 2 K steps x (hi hi, lo hi, hi lo)
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_base = smem_u32(s_a + buf * kTcABuf);
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const int kstep = 2 * s + ks;
                    const uint64_t a_hi = tc_desc(a_base + ks * kTcAStep, kTcALbo, kTcASbo);
                    const uint64_t a_lo = tc_desc(a_base + kTcALimb + ks * kTcAStep, kTcALbo, kTcASbo);
                    const uint64_t bh = tc_desc(smem_u32(s_b) + kstep * b_step, b_lbo, 128u);
                    const uint64_t bl = tc_desc(smem_u32(s_b + p.b_bytes) + kstep * b_step, b_lbo, 128u);
                    tc_mma(tmem, a_hi, bh, idesc, kstep > 0 ? 1u : 0u);
                    tc_mma(tmem, a_lo, bh, idesc, 1u);
                    tc_mma(tmem, a_hi, bl, idesc, 1u);
                }
                tc_commit(smem_u32(s_bar + buf));  // arrives when every MMA issued so far has completed
            }
        }
        // ---- epilogue: the last commit covers all MMAs of the tile
        {
            const unsigned last = g - 1;
            mbar_wait(smem_u32(s_bar + (last & 1u)), (last >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // warp w reads TMEM lanes 32 (w % 4) .. + 31 (its frames) and a quarter of the mel columns
            const int cols_per = ((p.n_pad / 4) + 3) & ~3;  // columns per warp group, multiple of 4
            const int c_lo = bg * cols_per;
            float *orow = p.out + (cb * p.n_mels) * (long long)p.T + ct;
            for (int c0 = c_lo; c0 < min(c_lo + cols_per, p.n_pad); c0 += 4) {
                uint32_t r0, r1, r2, r3;
                const uint32_t taddr = tmem + ((uint32_t)(fq * 32) << 16) + (uint32_t)c0;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const float acc[4] = {__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3)};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (col_ok && c0 + j < p.n_mels) orow[(long long)(c0 + j) * p.T] = tc_epilogue(acc[j], p);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();  // the accumulator may be overwritten by the next tile's first MMA
        }
        col = next_col;
    }
    // ---- teardown
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcTmemCols) : "memory");
}

}  // namespace b200mel
