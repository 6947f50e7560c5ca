// stc_probe.cuh — tcgen05.mma timing probe (tools/tc_bench.py probe).  Not part of the product path: one CTA issues
// chains of identical MMAs on zero-filled shared memory and reports clock64() cycles per MMA, for the operand layouts
// stft_tc.cuh uses (the overlapping "Hankel" rows of stage 1, the canonical K-major tiles of stage 2) and for
// reference shapes.  This is how the per-instruction costs quoted in DESIGN.md were measured.
#pragma once
#include "mel_tc.cuh"
#include "stft_tc.cuh"

namespace b200mel {

struct ProbeCfg {
    uint32_t a_off, a_lbo, a_sbo;   // A descriptor (bytes from the smem base)
    uint32_t b_off, b_lbo, b_sbo;
    uint32_t n;                     // UMMA N (M = 128, K = 16, fp16)
    uint32_t chain;                 // MMAs per commit
};

__global__ void __launch_bounds__(128, 1) stc_probe_kernel(const ProbeCfg *cfgs, int n_cfg, long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_addr;
    const int tid = threadIdx.x;
    for (int i = tid; i < 200 * 1024 / 16; i += 128) reinterpret_cast<int4 *>(smem)[i] = make_int4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_addr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_addr;
    if (tid == 0) {
        const uint32_t base = smem_u32(smem);
        uint32_t phase = 0;
        for (int c = 0; c < n_cfg; ++c) {
            const ProbeCfg cf = cfgs[c];
            const uint32_t idesc = (1u << 4) | (8u << 24) | ((cf.n >> 3) << 17);
            const uint64_t ad = tc_desc(base + cf.a_off, cf.a_lbo, cf.a_sbo), bd = tc_desc(base + cf.b_off, cf.b_lbo, cf.b_sbo);
            long long best_total = 1ll << 60, best_issue = 1ll << 60;
            for (int rep = 0; rep < 5; ++rep) {
                const long long t0 = clock64();
                for (uint32_t r = 0; r < cf.chain; ++r) tc_mma(tmem, ad, bd, idesc, r);
                const long long t1 = clock64();
                tc_commit(smem_u32(&bar));
                mbar_wait(smem_u32(&bar), phase);
                phase ^= 1u;
                const long long t2 = clock64();
                best_total = min(best_total, t2 - t0);
                best_issue = min(best_issue, t1 - t0);
            }
            out[2 * c] = best_issue;
            out[2 * c + 1] = best_total;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace b200mel
