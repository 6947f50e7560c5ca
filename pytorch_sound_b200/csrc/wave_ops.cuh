// wave_ops.cuh — the small operators either side of the spectral path (sm_100a):
//   pre-emphasis        y[n] = x[n] - c x[n-1], y[0] = x[0] - c x[1]      models/sound.py:66-81 (PreEmphasis.forward)
//   RMS volume norm     y = x / (std(x) / 10^(dB/10)), std over the tensor utils/calculate.py:56-63 (volume_norm_log_torch)
//   mel -> MFCC         out[b, c, t] = sum_m dct[c, m] mel[b, m, t]       models/transforms.py:419-430 (MelToMFCC.forward)
//   STFT-loss terms     per-clip sums of multi_stft_loss                  models/sound.py:139-141
// All three are HBM-bound streaming kernels (8, 8 and 4 (M + C) / M bytes per element); grids are sized in
// multiples of the SM count and every global access is a full 128-byte line per warp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200mel {

// One CTA row-strip per (sample block, clip): thread i of a block handles samples n, n + 256, n + 512, n + 768 of a
// row, so every load / store instruction of a warp covers one contiguous 128-byte line whatever the row alignment
// (22050-sample rows are only 8-byte aligned).  The left neighbour x[n-1] comes from the same or the previous line
// (L1 hit), so DRAM sees each sample once: 4 bytes read + 4 bytes written per sample.
__global__ void __launch_bounds__(256) preemph_kernel(const float *__restrict__ x, float *__restrict__ y, long long B,
                                                       int L, long long xs, long long ys, float coef) {
    const int n0 = blockIdx.x * 1024 + threadIdx.x;
    for (long long b = blockIdx.y; b < B; b += gridDim.y) {
        const float *xr = x + b * xs;
        float *yr = y + b * ys;
        float v[4], p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + 256 * j;
            if (n < L) {
                v[j] = xr[n];
                p[j] = xr[n > 0 ? n - 1 : 1];  // F.pad(input, (1, 0), 'reflect'): the sample before x[0] is x[1]
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + 256 * j;
            if (n < L) yr[n] = fmaf(-coef, p[j], v[j]);
        }
    }
}

// sum and sum of squares of a contiguous array in double precision -> acc[0], acc[1] (atomics, one per block)
__global__ void __launch_bounds__(256) moments_kernel(const float *__restrict__ x, long long n, double *acc) {
    double s = 0.0, q = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = (double)x[i];
        s += v;
        q = fma(v, v, q);
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    __shared__ double ss[8], sq[8];
    if ((threadIdx.x & 31) == 0) ss[threadIdx.x >> 5] = s, sq[threadIdx.x >> 5] = q;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) s += ss[w], q += sq[w];
        atomicAdd(acc, s);
        atomicAdd(acc + 1, q);
    }
}

// y = x * gain / std, std = unbiased standard deviation from the accumulated moments (torch.std default)
__global__ void __launch_bounds__(256) scale_by_std_kernel(const float *__restrict__ x, float *__restrict__ y, long long n,
                                                            const double *acc, float gain) {
    const double mean = acc[0] / (double)n;
    const double var = (acc[1] - (double)n * mean * mean) / (double)(n > 1 ? n - 1 : 1);
    const float k = gain / (float)sqrt(var > 0.0 ? var : 0.0);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = x[i] * k;
}

// multi_stft_loss reductions (models/sound.py:139-141): for every clip b over its n = F * T magnitudes
//   acc[3 b + 0] = sum (t - p)^2,  acc[3 b + 1] = sum t^2,  acc[3 b + 2] = sum |ln(t + eps) - ln(p + eps)|
// One streaming pass over both magnitude tensors (8 bytes per element, HBM-bound); double accumulation, one
// atomicAdd per quantity per CTA.  grid = (chunks, B).
__global__ void __launch_bounds__(256) stft_loss_partial_kernel(const float *__restrict__ pm, const float *__restrict__ tm,
                                                                 long long B, long long n, float eps, double *acc) {
    for (long long b = blockIdx.y; b < B; b += gridDim.y) {
        const float *pr = pm + b * n, *tr = tm + b * n;
        double d2 = 0.0, t2 = 0.0, l1 = 0.0;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
            const float pv = pr[i], tv = tr[i];
            const float d = tv - pv;
            d2 += (double)(d * d);
            t2 += (double)(tv * tv);
            l1 += (double)fabsf(logf(tv + eps) - logf(pv + eps));
        }
        for (int o = 16; o > 0; o >>= 1) {
            d2 += __shfl_xor_sync(0xffffffffu, d2, o);
            t2 += __shfl_xor_sync(0xffffffffu, t2, o);
            l1 += __shfl_xor_sync(0xffffffffu, l1, o);
        }
        __shared__ double sh[3][8];
        const int w = threadIdx.x >> 5;
        __syncthreads();  // previous clip's partials consumed
        if ((threadIdx.x & 31) == 0) sh[0][w] = d2, sh[1][w] = t2, sh[2][w] = l1;
        __syncthreads();
        if (threadIdx.x < 3) {
            double v = 0.0;
            for (int k = 0; k < (int)(blockDim.x >> 5); ++k) v += sh[threadIdx.x][k];
            atomicAdd(acc + 3 * b + threadIdx.x, v);
        }
    }
}
// out[0] += mean_b sqrt(acc[3b] / acc[3b+1])      (spectral convergence, :139)
// out[1] += mean_b acc[3b+2] / n                   (log-magnitude L1, :140)
// Accumulates, so the resolutions of one multi_stft_loss call add up in the same two floats.
__global__ void stft_loss_final_kernel(const double *acc, long long B, long long n, float *out) {
    double sc = 0.0, mg = 0.0;
    for (long long b = threadIdx.x; b < B; b += 32) {
        sc += sqrt(acc[3 * b]) / sqrt(acc[3 * b + 1]);
        mg += acc[3 * b + 2];
    }
    for (int o = 16; o > 0; o >>= 1) {
        sc += __shfl_xor_sync(0xffffffffu, sc, o);
        mg += __shfl_xor_sync(0xffffffffu, mg, o);
    }
    if (threadIdx.x == 0) {
        out[0] += (float)(sc / (double)B);
        out[1] += (float)(mg / (double)B / (double)n);
    }
}

// All-gather of mel blocks over NVLink by PULLING from peer memory: every rank owns the same symmetric
// buffer layout (world blocks back to back); block r was just written by rank r's extraction kernel into ITS
// buffer.  This kernel copies every peer's own block from the peer's buffer (mapped peer pointers from
// torch.distributed._symmetric_memory) into the local buffer with 16-byte loads — the measured fastest peer path
// (kernel LDG.128 ~775 GB/s per direction vs ~725 for copy engines / NCCL at these sizes) — so after it the
// local buffer holds all world blocks.  Loads are volatile (peer lines may sit in L1 from the previous step);
// several independent loads per thread keep ~2 us of NVLink latency covered.  The caller orders it after a
// device-side barrier (all ranks have finished writing their block).
struct PullArgs {
    const float *peer[16];      // peer[r] = rank r's symmetric buffer as mapped in this process (peer[rank] unused)
    long long off[17];          // element offset of block r in every buffer; off[world] = total elements
    int *peer_sync[16];         // peer_sync[r] = rank r's sync words as mapped here: [0, world) arrival flags, [world] epoch,
                                // [world + 1] ticket; null = the caller orders the ranks itself
    int world, rank;
    int pull;                   // 0: barrier only (the copies are done by the copy engines, b200mel_gather_copy)
    int stages;                 // gather_tma_kernel: ring stages in use (<= kPullMaxStages)
};
// The cross-rank barrier lives INSIDE the kernel (no separate barrier launch, CUDA-graph replayable): the step
// number is epoch + 1, where the epoch word is advanced by the last CTA of every launch — so it is the same for
// every CTA of a launch and for the matching launches of all ranks, and a replayed graph keeps counting.  Block
// 0 publishes "my block of this step is written" (the extraction kernel precedes this launch in stream order) to
// every peer with a system-scope release store; every CTA then spins (bounded, then traps) with system-scope
// acquire loads on its OWN rank's flags until all peers have published this step.
__global__ void __launch_bounds__(512, 2) gather_pull_kernel(float *__restrict__ local, const PullArgs a) {
    int *sync = a.peer_sync[a.rank];
    __shared__ int s_step;
    if (sync) {
        if (threadIdx.x == 0) s_step = *reinterpret_cast<volatile int *>(sync + a.world) + 1;
        __syncthreads();
        const int step = s_step;
        const int r = threadIdx.x;
        if (r < a.world && r != a.rank) {
            if (blockIdx.x == 0) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(a.peer_sync[r] + a.rank), "r"(step) : "memory");
            int seen;
            unsigned spins = 0;
            do {
                asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(sync + r) : "memory");
                if (seen - step < 0 && ++spins > (1u << 26)) __trap();  // a peer never arrived: CUDA error, not a hung GPU
            } while (seen - step < 0);
        }
        __syncthreads();
    }
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    if (a.pull) {
        // Peer-interleaved: one 16-byte load per peer in flight per thread (groups of 8 peers), so every link is busy for
        // the whole kernel and nothing drains between peers.  Block heads / tails that are not 16-byte aligned go
        // through scalar loads.  Pointers and offsets are re-derived from the kernel parameters (constant bank) at every
        // use: the kernel must stay at <= 64 registers so that two 512-thread CTAs fit on an SM.
        auto body = [&](int r, long long &lo4, long long &n4) {  // 16-byte aligned body of block r, in floats / float4s
            const long long lo = a.off[r], hi = a.off[r + 1];
            long long b_lo = (lo + 3) & ~3LL, b_hi = hi & ~3LL;
            if (b_lo >= b_hi) b_lo = b_hi = hi;
            lo4 = b_lo;
            n4 = (b_hi - b_lo) >> 2;
        };
        long long n4_max = 0;
        for (int r = 0; r < a.world; ++r) {
            if (r == a.rank) continue;
            const long long lo = a.off[r], hi = a.off[r + 1];
            long long lo4, n4;
            body(r, lo4, n4);
            const float *src = a.peer[r];
            for (long long e = lo + tid; e < min(lo4, hi); e += nthreads) local[e] = *reinterpret_cast<const volatile float *>(src + e);
            for (long long e = lo4 + 4 * n4 + tid; e < hi; e += nthreads) local[e] = *reinterpret_cast<const volatile float *>(src + e);
            n4_max = max(n4_max, n4);
        }
#pragma unroll 1
        for (int g = 0; g < a.world; g += 8) {
            for (long long k = tid; k < n4_max; k += nthreads) {
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = g + i;
                    if (r < a.world && r != a.rank) {
                        long long lo4, n4;
                        body(r, lo4, n4);
                        if (k < n4)
                            asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                                         : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w)
                                         : "l"(reinterpret_cast<const float4 *>(a.peer[r] + lo4) + k));
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = g + i;
                    if (r < a.world && r != a.rank) {
                        long long lo4, n4;
                        body(r, lo4, n4);
                        if (k < n4) reinterpret_cast<float4 *>(local + lo4)[k] = v[i];
                    }
                }
            }
        }
    }
    if (sync) {  // the last CTA of the launch advances the epoch
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const int t = atomicAdd(sync + a.world + 1, 1);
            if (t == (int)gridDim.x - 1) {
                sync[a.world + 1] = 0;
                __threadfence();
                *reinterpret_cast<volatile int *>(sync + a.world) = s_step;
            }
        }
    }
}

// The same gather with the TMA engine instead of the load / store units: one thread per CTA streams the peers' blocks
// through a ring of shared-memory stages with bulk async copies — cp.async.bulk global (peer, over NVLink) -> shared,
// completion on an mbarrier, then cp.async.bulk shared -> global (local buffer) — kPullStages x kPullChunk bytes in
// flight per CTA with no registers and no load instructions involved, so a handful of CTAs covers the NVLink
// bandwidth-delay product (~770 GB/s x ~3 us = 2.3 MB) and the SMs the gather takes away from the concurrent
// extraction kernel shrink from 48 to 16.  Consecutive chunks of a CTA go to different peers (every link busy all
// the time).  Barrier, epoch and the unaligned heads / tails of a block are those of gather_pull_kernel.
constexpr int kPullStages = 4, kPullMaxStages = 7, kPullChunk = 32768;   // default / largest ring (7 x 32 KB = 224 KB)
__global__ void __launch_bounds__(128, 1) gather_tma_kernel(float *__restrict__ local, const PullArgs a) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ uint64_t s_full[kPullMaxStages];
    __shared__ int s_step;
    int *sync = a.peer_sync[a.rank];
    const int n_stages = a.stages;
    if (threadIdx.x == 0) {
        for (int i = 0; i < n_stages; ++i) mbar_init(smem_u32(s_full + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (sync) {
        if (threadIdx.x == 0) s_step = *reinterpret_cast<volatile int *>(sync + a.world) + 1;
        __syncthreads();
        const int step = s_step;
        const int r = threadIdx.x;
        if (r < a.world && r != a.rank) {
            if (blockIdx.x == 0) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(a.peer_sync[r] + a.rank), "r"(step) : "memory");
            int seen;
            unsigned spins = 0;
            do {
                asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(sync + r) : "memory");
                if (seen - step < 0 && ++spins > (1u << 26)) __trap();  // a peer never arrived: CUDA error, not a hung GPU
            } while (seen - step < 0);
        }
    }
    __syncthreads();
    auto body = [&](int r, long long &lo4, long long &bytes) {  // 16-byte aligned body of block r: first float, bytes
        const long long lo = a.off[r], hi = a.off[r + 1];
        long long b_lo = (lo + 3) & ~3LL, b_hi = hi & ~3LL;
        if (b_lo >= b_hi) b_lo = b_hi = hi;
        lo4 = b_lo;
        bytes = (b_hi - b_lo) * 4;
    };
    // unaligned heads / tails (<= 3 floats each side of a block): scalar, by the threads of CTA 0
    if (blockIdx.x == 0) {
        for (int r = 0; r < a.world; ++r) {
            if (r == a.rank) continue;
            const long long lo = a.off[r], hi = a.off[r + 1];
            long long lo4, bytes;
            body(r, lo4, bytes);
            const float *src = a.peer[r];
            for (long long e = lo + threadIdx.x; e < min(lo4, hi); e += blockDim.x) local[e] = *reinterpret_cast<const volatile float *>(src + e);
            for (long long e = lo4 + bytes / 4 + threadIdx.x; e < hi; e += blockDim.x) local[e] = *reinterpret_cast<const volatile float *>(src + e);
        }
    }
    if (threadIdx.x == 0) {
        // chunk id = c * (world - 1) + p: chunk c of the p-th peer after this rank; ids blockIdx, blockIdx + gridDim, ...
        const int peers = a.world - 1;
        long long max_chunks = 0;
        for (int r = 0; r < a.world; ++r) {
            long long lo4, bytes;
            body(r, lo4, bytes);
            if (r != a.rank) max_chunks = max(max_chunks, (bytes + kPullChunk - 1) / kPullChunk);
        }
        const long long n_ids = max_chunks * peers;
        struct Chunk { const unsigned char *src; unsigned char *dst; uint32_t bytes; };
        auto chunk_of = [&](long long id, Chunk &ck) -> bool {  // false: this id is past the end of its (shorter) block
            const long long c = id / peers;
            const int r = (a.rank + 1 + (int)(id - c * peers)) % a.world;
            long long lo4, bytes;
            body(r, lo4, bytes);
            const long long off = c * kPullChunk;
            if (off >= bytes) return false;
            ck.src = reinterpret_cast<const unsigned char *>(a.peer[r] + lo4) + off;
            ck.dst = reinterpret_cast<unsigned char *>(local + lo4) + off;
            ck.bytes = (uint32_t)min((long long)kPullChunk, bytes - off);
            return true;
        };
        long long next_load = blockIdx.x;   // next chunk id to request
        long long loaded = 0, stored = 0;   // chunks requested / written so far (ring positions)
        Chunk pending[kPullMaxStages];
        auto request = [&]() {              // issue the load of the next valid chunk into ring stage loaded % kPullStages
            Chunk ck;
            while (next_load < n_ids && !chunk_of(next_load, ck)) next_load += gridDim.x;
            if (next_load >= n_ids) return false;
            next_load += gridDim.x;
            const int st = (int)(loaded % n_stages);
            const uint32_t bar = smem_u32(s_full + st);
            pending[st] = ck;
            mbar_arrive_expect_tx(bar, ck.bytes);
            tma_load_1d(smem_u32(ring + st * kPullChunk), ck.src, ck.bytes, bar);
            ++loaded;
            return true;
        };
        for (int i = 0; i < n_stages; ++i)
            if (!request()) break;
        while (stored < loaded) {
            const int st = (int)(stored % n_stages);
            mbar_wait(smem_u32(s_full + st), (uint32_t)(stored / n_stages) & 1u);
            const Chunk ck = pending[st];
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(ck.dst)),
                         "r"(smem_u32(ring + st * kPullChunk)), "r"(ck.bytes)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            ++stored;
            // the stage may be refilled once the store has READ it; the other stages' loads stay in flight meanwhile
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            request();
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every store has landed before the epoch moves on
    }
    if (sync) {  // the last CTA of the launch advances the epoch
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const int t = atomicAdd(sync + a.world + 1, 1);
            if (t == (int)gridDim.x - 1) {
                sync[a.world + 1] = 0;
                __threadfence();
                *reinterpret_cast<volatile int *>(sync + a.world) = s_step;
            }
        }
    }
}

// mel (B, M, T) -> mfcc (B, C, T), DCT matrix in CONSTANT memory.  A thread owns one (b, t) column: its M mel values
// are read once into registers (a warp reads 32 consecutive frames of a mel row: full 128-byte lines) and every
// output row is M FFMAs whose weight operand comes straight from the constant bank — warp-uniform, so it costs no
// load instruction and no shared-memory wavefront (the shared-memory variant below is bound by exactly those:
// 25.5 us at C2 against 3.6 us here; the HBM time of 10.7 MB is 1.6 us).  Rows of the matrix are kM floats apart.
constexpr int kDctConstRows = 64, kDctConstCols = 128;
__constant__ float c_dct[kDctConstRows * kDctConstCols];
template <int kM>
__global__ void __launch_bounds__(128) dct_const_kernel(const float *__restrict__ mel, float *__restrict__ out, long long B,
                                                         int C, int T) {
    const long long cols = B * (long long)T;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cols; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / T;
        const int t = (int)(i - b * T);
        const float *src = mel + (b * kM) * (long long)T + t;
        float *dst = out + (b * C) * (long long)T + t;
        float v[kM];
#pragma unroll
        for (int m = 0; m < kM; ++m) v[m] = src[(long long)m * T];
#pragma unroll 2
        for (int c = 0; c < C; ++c) {
            const float *w = c_dct + c * kM;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;  // four chains, summed as the shared-memory kernel does
#pragma unroll
            for (int m = 0; m + 3 < kM; m += 4) {
                a0 = fmaf(w[m], v[m], a0);
                a1 = fmaf(w[m + 1], v[m + 1], a1);
                a2 = fmaf(w[m + 2], v[m + 2], a2);
                a3 = fmaf(w[m + 3], v[m + 3], a3);
            }
            dst[(long long)c * T] = (a0 + a1) + (a2 + a3);
        }
    }
}

// mel (B, M, T) -> mfcc (B, C, T), general shapes.  A warp owns 32 consecutive (b, t) columns (every global access a contiguous
// 128-byte line) and kDctRows output rows: the column's M mel values are read into registers once per warp (the
// C / kDctRows warps that share a column block re-read them from L1/L2, the mel tensor is small), and the DCT rows
// come from shared memory (rows padded to a multiple of 4) as 128-bit warp-wide broadcasts, one LDS.128 per 4 FFMAs.
// Splitting the rows over warps gives the grid enough warps to hide latency (C2: 22272 columns only).  Measured:
// 25 us at C2 — bound by the broadcast loads (a broadcast LDS.128 still costs four shared-memory wavefronts), not
// by HBM; a register-blocked 4-column variant spilled and was no faster.
constexpr int kDctRows = 8;
template <int kMaxM>
__global__ void __launch_bounds__(128) dct_kernel(const float *__restrict__ mel, const float *__restrict__ dct,
                                                   float *__restrict__ out, long long B, int M, int C, int T) {
    extern __shared__ __align__(16) float s_dct[];  // [C][Mp], Mp = M rounded up to 4, zero padded
    const int Mp = (M + 3) & ~3;
    for (int i = threadIdx.x; i < C * Mp; i += blockDim.x) {
        const int c = i / Mp, m = i - c * Mp;
        s_dct[i] = m < M ? dct[c * M + m] : 0.f;
    }
    __syncthreads();
    const long long cols = B * (long long)T;
    const int lane = threadIdx.x & 31;
    const int n_cg = (C + kDctRows - 1) / kDctRows;
    const long long n_units = ((cols + 31) / 32) * n_cg;  // (column block, row group) pairs, row group fastest
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long u = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; u < n_units; u += warps) {
        const long long blk = u / n_cg;
        const int c0 = (int)(u - blk * n_cg) * kDctRows;
        const long long i = blk * 32 + lane;
        const bool ok = i < cols;
        const long long b = ok ? i / T : 0;
        const int t = ok ? (int)(i - b * T) : 0;
        const float *src = mel + (b * M) * (long long)T + t;
        float *dst = out + (b * C) * (long long)T + t;
        float v[kMaxM];
#pragma unroll
        for (int m = 0; m < kMaxM; ++m) v[m] = (m < M && ok) ? src[(long long)m * T] : 0.f;
#pragma unroll 2
        for (int c = c0; c < min(c0 + kDctRows, C); ++c) {
            const float4 *w4 = reinterpret_cast<const float4 *>(s_dct + c * Mp);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int g = 0; g < kMaxM / 4; ++g) {
                if (4 * g < M) {
                    const float4 w = w4[g];
                    a0 = fmaf(w.x, v[4 * g], a0);
                    a1 = fmaf(w.y, v[4 * g + 1], a1);
                    a2 = fmaf(w.z, v[4 * g + 2], a2);
                    a3 = fmaf(w.w, v[4 * g + 3], a3);
                }
            }
            if (ok) dst[(long long)c * T] = (a0 + a1) + (a2 + a3);
        }
    }
}

}  // namespace b200mel
