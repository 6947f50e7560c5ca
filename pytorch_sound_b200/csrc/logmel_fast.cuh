// logmel_fast.cuh — the common-case instantiation of the fused STFT -> |.| -> mel -> log kernel (sm_100a).
//
// Same pipeline and index algebra as logmel_kernel.cuh (persistent CTA per SM, every warp an independent
// pipeline over (clip, frame-pair) tasks, one TMA bulk copy per task, register-resident 1024-point FFT as two
// radix-32 passes, conjugate-symmetry separation, banded mel from shared memory, log epilogue in registers) with
// everything the BASELINE geometries never vary fixed at COMPILE time:
//     n_fft = win_length = 1024 (pair mode, generated periodic Hann), hop = 256, no per-clip `lengths`,
//     a log epilogue, and a filterbank whose round signature (weight groups per mel round) is a template
//     parameter — so the mel rounds are straight-line code with compile-time trip counts and table offsets.
// The generic kernel spends ~23 % of its 1494 warp-instructions per task on run-time generality (validity flags,
// `lengths`, hop / window / round-count branches, constant-bank indexing of the round tables); this body has none
// of it.  Edge tasks (reflect halo, the odd last frame of a clip) are handled in place: the halo patch is the one
// warp-uniform branch, and the second frame of a pair is ALWAYS computed (on the reflected continuation of the
// clip) and only its store is predicated.  b200mel_forward picks this kernel when the plan and the call qualify
// and falls back to the generic body otherwise; the two are separately compiled instances of the same arithmetic
// and agree to a few ulp of the log-mel value (<= 1e-5, measured 2.4e-6; the parity bar is 1e-4) —
// tests/test_gpu_round2.py::test_fast_and_generic_kernels_agree.
#pragma once
#include "logmel_kernel.cuh"

namespace b200mel {

constexpr int kFastHop = 256;
constexpr int kFastSpan = 1024 + kFastHop;  // samples staged per task: two frames, 75 % overlapped

struct FastDesc {
    int b, t0;       // clip, first frame of the pair
    int delta;       // stage shift: sample s of the span sits at stage[s - s_first + delta]
    unsigned flags;  // 1: frame t0 exists (always, without `lengths`), 2: frame t0 + 1 exists, 4: the span leaves [0, Li) (reflect patch)
    int Li;          // valid samples of the clip (kLen instantiations: lengths[b]; otherwise L)
};

// Locate a task and (lane 0) request its samples: one bulk copy of the in-range part of the 1280-sample span
// (copy_geom in logmel_kernel.cuh: widened to 16-byte boundaries, clamped to the tensor).
// kLen: per-clip `lengths` (zero-padded batches, SpeechDataLoader.pad_collate_fn): the clip ends at lengths[b], reflection
// happens there, and a task whose first frame lies past the clip's last frame requests nothing.
template <bool kLen = false>
__device__ __forceinline__ FastDesc fast_request(const KParams &p, int b, int q, int lane, uint32_t stage_s, uint32_t bar) {
    FastDesc d;
    d.b = b;
    d.t0 = 2 * q;
    d.Li = p.L;
    int Ti = p.T;
    if constexpr (kLen) {
        if (p.lengths) {
            d.Li = min(__ldg(p.lengths + b), p.L);
            Ti = min(frames_of(d.Li, kFastSpan - kFastHop, kFastHop, p.pad), p.T);
        }
    }
    if (kLen && d.t0 >= Ti) {
        d.delta = 0, d.flags = 0;
        return d;
    }
    const CopyGeom g = copy_geom(p, b, d.t0 * kFastHop - p.pad, kFastSpan, d.Li);
    d.delta = g.delta;
    d.flags = 1u | (d.t0 + 1 < Ti ? 2u : 0u) | (g.patch ? 4u : 0u);
    if (lane == 0) issue_copy(g, stage_s, bar);
    return d;
}

// Reflected halo of an edge task; the span is always 1280 samples here (the second frame of a pair is computed
// even when it lies past the clip's last frame — on the reflected continuation — and simply not stored).
__device__ __forceinline__ void fast_patch_halo(const KParams &p, const FastDesc &d, float *stage, int lane) {
    const int s_first = d.t0 * kFastHop - p.pad;
    patch_stage(p, d.b, s_first, kFastSpan, d.Li, copy_geom(p, d.b, s_first, kFastSpan, d.Li), stage, lane);
}

// stage -> registers with the generated periodic Hann applied: a[j] = {x_t[32 j + lane], x_t+1[32 j + lane]} * w[32 j + lane]
// (hop = 8 * 32: element j of frame t+1 IS element j + 8 of frame t in the stage — 40 loads for the two frames).
__device__ __forceinline__ void fast_load_windowed(float2 *a, const float *x0, float2 hann_cs) {
    asm volatile("" : "+f"(hann_cs.x), "+f"(hann_cs.y));  // keep the window a per-task computation (see load_windowed_pair)
    float raw[40];
#pragma unroll
    for (int j = 0; j < 40; ++j) raw[j] = x0[32 * j];
    static_for<0, 16>([&](auto j_) {
        constexpr int j = decltype(j_)::value;
        const float t = fmaf(TwConst::s32[j], hann_cs.y, TwConst::c32[j] * hann_cs.x);  // 0.25 cos(theta_j + phi)
        const float w0 = 0.25f - t, w1 = 0.25f + t;                                     // slots j and j + 16
        a[j] = __fmul2_rn(make_float2(raw[j], raw[j + 8]), make_float2(w0, w0));
        a[j + 16] = __fmul2_rn(make_float2(raw[j + 16], raw[j + 24]), make_float2(w1, w1));
    });
}

__device__ __forceinline__ float fast_epilogue(float x, const KParams &p) {
    float y;
    const float t = fmaxf(x, p.ep_floor) + p.ep_offset;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(t));
    y = fminf(fmaxf(y * p.log_scale, p.lo), p.hi);
    return fmaf(y, p.norm_scale, p.norm_bias);
}

__host__ __device__ constexpr int fast_sig_rounds(unsigned sig) { return sig > 0xfffu ? 4 : sig > 0xffu ? 3 : sig > 0xfu ? 2 : 1; }
__host__ __device__ constexpr int fast_sig_groups(unsigned sig, int r) { return (int)((sig >> (4 * (fast_sig_rounds(sig) - 1 - r))) & 15u); }
__host__ __device__ constexpr int fast_sig_wbase(unsigned sig, int r) {  // float4 index of round r's weights ([group][lane] layout)
    int wb = 0;
    for (int i = 0; i < r; ++i) wb += fast_sig_groups(sig, i) * 32;
    return wb;
}

constexpr int kDctPitch = 64;  // floats per mel row of the shared DCT matrix: {coefficient lane, coefficient 32 + lane} at 2 * lane

// kSig: one hex digit per mel round = its float4 weight groups (e.g. 0x731: three rounds of 7, 3 and 1 groups — the
// settings.py filterbank 22050 Hz / 1024 / 80 mels / 0-8000 Hz).  kTop: 32-bin groups separated (12: bins < 384).
// kDct: MelToMFCC / MFCC (models/transforms.py:419-455) fused as an epilogue — the lanes leave their log-mel values
// of the two frames in a per-warp column, and lane c then forms coefficients c and 32 + c of both frames from the
// shared DCT matrix (one broadcast 128-bit load of two column entries per four FFMA2).
// kLen: per-clip `lengths` and / or the SpectrogramMasker frame mask (the data-path hook: GpuFeatureLoader passes both).
template <int kTop, unsigned kSig, int kPower, bool kDct = false, bool kLen = false>
__global__ void __launch_bounds__(kMaxWarps * 32, 1) logmel_fast_kernel(const KParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_warps = blockDim.x >> 5;
    constexpr int kRounds = fast_sig_rounds(kSig);

    float2 *s_tw = reinterpret_cast<float2 *>(smem_raw);
    const MelEntry *s_ent = reinterpret_cast<const MelEntry *>(smem_raw + p.off_entries);
    const float *s_melw = reinterpret_cast<const float *>(smem_raw + p.off_melw);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + p.off_bar);
    unsigned char *region = smem_raw + p.off_regions + warp * p.region_bytes;
    float2 *buf = reinterpret_cast<float2 *>(region);    // transpose buffer
    float2 *tile2 = reinterpret_cast<float2 *>(region);  // magnitude tile
    float *stage = reinterpret_cast<float *>(region + kStageOff);
    const uint32_t stage_s = smem_u32(stage);
    const uint32_t bar = smem_u32(s_bar + warp);
    uint32_t parity = 0;

    const long long stride = (long long)gridDim.x * n_warps;
    long long task = (long long)warp * gridDim.x + blockIdx.x;
    int cb = (int)(task / p.tasks_per_clip);
    int cq = (int)(task - (long long)cb * p.tasks_per_clip);

    // Prologue ordered for programmatic dependent launch, as in logmel_kernel: plan-owned tables first (four bulk
    // copies on one mbarrier), caller memory only after griddepcontrol.wait.
    const uint32_t tbar = smem_u32(s_bar + kMaxWarps);
    if (lane == 0) mbar_init(bar, 1);
    if (threadIdx.x == 0) {
        mbar_init(tbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t b_tw = 32 * 32 * 8;
        const uint32_t b_ent = (uint32_t)kRounds * 32u * (uint32_t)sizeof(MelEntry), b_w = (uint32_t)p.mel_w_len * 4u;
        mbar_arrive_expect_tx(tbar, b_tw + b_ent + b_w);  // the window table is not needed: the Hann is generated
        tma_load_1d(smem_u32(s_tw), p.tw, b_tw, tbar);
        tma_load_1d(smem_u32(smem_raw + p.off_entries), p.mel_entries, b_ent, tbar);
        tma_load_1d(smem_u32(smem_raw + p.off_melw), p.mel_w, b_w, tbar);
    } else if (lane == 0) {
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (lane == 0 && task < p.n_tasks) {  // hint: pull the first span towards L2 while the previous kernel may still run
        const int s0 = max(cq * 2 * kFastHop - p.pad, 0);
        const int s1 = min(s0 + kFastSpan, p.L);
        const float *row = p.wav + (long long)cb * p.row_stride;
        const uintptr_t a16 = reinterpret_cast<uintptr_t>(row + s0) & ~(uintptr_t)15;
        const uintptr_t e16 = reinterpret_cast<uintptr_t>(row + s1) & ~(uintptr_t)15;
        if (e16 > a16)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(reinterpret_cast<const void *>(a16))),
                         "r"((uint32_t)(e16 - a16))
                         : "memory");
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    FastDesc cur;
    cur.b = cur.t0 = cur.delta = cur.Li = 0;
    cur.flags = 0;
    __syncwarp();
    if (task < p.n_tasks) cur = fast_request<kLen>(p, cb, cq, lane, stage_s, bar);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int dct_rows = (p.n_mels + 1) & ~1;
    float *s_dct = reinterpret_cast<float *>(smem_raw + p.off_dct);
    float2 *col = reinterpret_cast<float2 *>(smem_raw + p.off_col + warp * p.col_bytes);
    if constexpr (kDct) {   // the DCT matrix is caller memory: after griddepcontrol.wait
        for (int i = threadIdx.x; i < dct_rows * kDctPitch; i += blockDim.x) {
            const int m = i / kDctPitch, c = ((i & 1) << 5) | ((i % kDctPitch) >> 1);   // pairs {lane, 32 + lane}
            s_dct[i] = (m < p.n_mels && c < p.n_mfcc) ? __ldg(p.dct + (long long)c * p.n_mels + m) : 0.f;
        }
        if (lane == 0) col[dct_rows - 1] = make_float2(0.f, 0.f);  // the padding entry of an odd mel count
    }
    __syncthreads();
    mbar_wait(tbar, 0);

    float2 hann_cs;
    {
        float sn, cs;
        sincospif((float)lane * (1.0f / 512.0f), &sn, &cs);
        hann_cs = make_float2(0.25f * cs, 0.25f * sn);
    }
    const int2 *ent2 = reinterpret_cast<const int2 *>(s_ent) + lane;
    const float4 *wbase = reinterpret_cast<const float4 *>(s_melw) + lane;
    const int partner = (32 - lane) & 31;

    for (; task < p.n_tasks; task += stride) {
        const FastDesc d = cur;
        cb += p.stride_b;
        cq += p.stride_q;
        if (cq >= p.tasks_per_clip) cq -= p.tasks_per_clip, ++cb;
        float2 a[32];

        if constexpr (kLen) {
            if (p.out_fmask && lane < 2 && d.t0 + lane < p.T)  // frame mask of this task's frames (models/transforms.py:397-416)
                p.out_fmask[(long long)d.b * p.T + d.t0 + lane] = ((d.t0 + lane) * kFastHop - p.win_half < d.Li) ? 1.f : 0.f;
            if (!(d.flags & 2u)) {   // frames past the clip's own end are zeros, as pad_collate_fn zero-pads per-item features
                for (int tt = d.t0 + ((d.flags & 1u) ? 1 : 0); tt <= d.t0 + 1 && tt < p.T; ++tt)
                    for (int m = lane; m < p.n_mels; m += 32) p.out_mel[((long long)d.b * p.n_mels + m) * (long long)p.T + tt] = 0.f;
            }
            if (!(d.flags & 1u)) {   // nothing was requested for this task
                if (task + stride < p.n_tasks) cur = fast_request<kLen>(p, cb, cq, lane, stage_s, bar);
                continue;
            }
        }
        mbar_wait(bar, parity);
        parity ^= 1;
        if (d.flags & 4u) fast_patch_halo(p, d, stage, lane);
        fast_load_windowed(a, stage + d.delta + lane, hann_cs);

        fft32(a);      // pass 1: lane = n2, FFT over n1
        __syncwarp();  // every lane has consumed the stage before the transpose buffer overwrites it
        static_for<0, 32>([&](auto k1_) {
            constexpr int k1 = decltype(k1_)::value;
            buf[k1 * kBufStride + lane] = a[fft32_pos(k1)];
        });
        __syncwarp();
        xpose_read_twiddle<B200MEL_XPOSE_BATCH>(a, buf, s_tw, lane);
        __syncwarp();  // transpose buffer is dead: magnitude tile and next stage may reuse it

        if (task + stride < p.n_tasks) cur = fast_request<kLen>(p, cb, cq, lane, stage_s, bar);

        int2 ents[kRounds];  // {lo, m} of every mel round, fetched here so the latency hides under pass 2
#pragma unroll
        for (int r = 0; r < kRounds; ++r) ents[r] = ent2[r * 32];

        fft32(a);  // pass 2: lane = k1, FFT over n2 -> Z[k1 + 32 k2] at a[pos(k2)]

        // real-input separation + magnitudes (see logmel_kernel.cuh)
        static_for<0, kTop>([&](auto k2_) {
            constexpr int k2 = decltype(k2_)::value;
            const float2 A = a[fft32_pos(k2)];
            const float2 g0 = a[fft32_pos((32 - k2) & 31)];
            const float2 g1 = a[fft32_pos(31 - k2)];
            float2 Bv;
            Bv.x = __shfl_sync(0xffffffffu, lane == 0 ? g0.x : g1.x, partner);
            Bv.y = __shfl_sync(0xffffffffu, lane == 0 ? g0.y : g1.y, partner);
            tile2[lane + 32 * k2] = pair_magnitudes<kPower>(A, Bv, p.mag_eps);
        });
        if constexpr (kTop == 16) {
            if (lane == 0) {  // bin 512 is its own partner
                const float2 A = a[fft32_pos(16)];
                tile2[512] = magnitude2<kPower>(make_float2(2.f * A.x, 0.f), make_float2(2.f * A.y, 0.f), p.mag_eps);
            }
            if (lane < 7) tile2[513 + lane] = make_float2(0.f, 0.f);  // padded tail the float4 weight groups may touch
        }
        __syncwarp();

        float *orow = p.out_mel + (long long)d.b * p.n_mels * (long long)p.T + d.t0;
        const bool valid1 = d.flags & 2u;
        // banded mel: compile-time rounds, each one loads -> FFMA chains -> log epilogue -> stores (hoisting every round's
        // loads above the first round's arithmetic measured the same 25.5 us at C2 and 2.3 us SLOWER inside the generic
        // kernel, where it cost registers the FFT phases want)
        static_for<0, kRounds>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            float acc0 = 0.f, acc1 = 0.f;
            mel_groups<true, B200MEL_MEL_CHUNK>(fast_sig_groups(kSig, r), wbase + fast_sig_wbase(kSig, r), region + ents[r].x * 8, acc0, acc1);
            const float y0 = fast_epilogue(acc0, p), y1 = fast_epilogue(acc1, p);
            if (ents[r].y >= 0) {
                if constexpr (kDct) col[ents[r].y] = make_float2(y0, y1);
                if (!kDct || p.out_mel) {
                    float *o = orow + (long long)ents[r].y * p.T;
                    o[0] = y0;
                    if (valid1) o[1] = y1;
                }
            }
        });
        __syncwarp();  // tile reads done before the next task's transpose overwrites the region
        if constexpr (kDct) {
            // coefficients lane (c0) and 32 + lane (c1) of frames t0, t0 + 1; two accumulators each (even / odd mel rows)
            // halve the dependent FFMA2 chains
            float2 c0 = make_float2(0.f, 0.f), c1 = c0, e0 = c0, e1 = c0;
            const float2 *dl = reinterpret_cast<const float2 *>(s_dct) + lane;
            if (p.n_mfcc > 32) {
#pragma unroll 4
                for (int m = 0; m < dct_rows; m += 2) {
                    const float4 y = *reinterpret_cast<const float4 *>(col + m);   // {mel m: t0, t0+1, mel m+1: t0, t0+1}
                    const float2 da = dl[m * (kDctPitch / 2)], db = dl[(m + 1) * (kDctPitch / 2)];
                    c0 = __ffma2_rn(make_float2(da.x, da.x), make_float2(y.x, y.y), c0);
                    c1 = __ffma2_rn(make_float2(da.y, da.y), make_float2(y.x, y.y), c1);
                    e0 = __ffma2_rn(make_float2(db.x, db.x), make_float2(y.z, y.w), e0);
                    e1 = __ffma2_rn(make_float2(db.y, db.y), make_float2(y.z, y.w), e1);
                }
            } else {
                const float *ds = s_dct + 2 * lane;
#pragma unroll 4
                for (int m = 0; m < dct_rows; m += 2) {
                    const float4 y = *reinterpret_cast<const float4 *>(col + m);
                    c0 = __ffma2_rn(make_float2(ds[m * kDctPitch], ds[m * kDctPitch]), make_float2(y.x, y.y), c0);
                    e0 = __ffma2_rn(make_float2(ds[(m + 1) * kDctPitch], ds[(m + 1) * kDctPitch]), make_float2(y.z, y.w), e0);
                }
            }
            c0 = __fadd2_rn(c0, e0);
            c1 = __fadd2_rn(c1, e1);
            float *oc = p.out_mfcc + (long long)d.b * p.n_mfcc * (long long)p.T + d.t0;
            if (lane < p.n_mfcc) {
                float *o = oc + (long long)lane * p.T;
                o[0] = c0.x;
                if (valid1) o[1] = c0.y;
            }
            if (lane + 32 < p.n_mfcc) {
                float *o = oc + (long long)(lane + 32) * p.T;
                o[0] = c1.x;
                if (valid1) o[1] = c1.y;
            }
            __syncwarp();  // column reads done before the next task's mel rounds rewrite it
        }
    }
}

}  // namespace b200mel
