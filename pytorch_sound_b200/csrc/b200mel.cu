// b200mel.cu — fused STFT -> |.| -> mel -> log kernel for sm_100a and its C ABI.
//
// Replaces the op chain of pytorch_sound's spectral modules
//   reflect-pad -> conv1d-DFT | torch.stft -> sqrt(re^2+im^2) -> mel matmul -> log -> clamp
// (models/transforms.py:53-69, 231-244, 297-311, 351-366; interface/hifi_gan.py:46-63)
// by ONE launch per clip batch.  See DESIGN.md for the data layout and roofline.
//
// Work decomposition (v1): one warp = one 1024-point complex FFT held entirely in
// registers (32 complex values per lane, two radix-32 passes, one shared-memory
// transpose):
//   n_fft = 1024 ("pair" mode):  two consecutive real frames t, t+1 are packed as
//        re/im of one complex FFT and separated with the conjugate-symmetry identity;
//   n_fft = 2048 ("split" mode): one real frame is packed even/odd into a 1024-point
//        complex FFT and finished with the real-input split pass.
// The magnitudes go to a per-warp shared-memory tile, the banded (CSR) mel filterbank
// is applied from there, the log/clamp epilogue runs in registers.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <new>
#include <string>
#include <vector>

#include "../../include/b200mel.h"
#include "fft32.cuh"

namespace b200mel {

// ----------------------------------------------------------------------------------------------
// kernel parameters
// ----------------------------------------------------------------------------------------------
struct KParams {
    const float *wav;
    long long row_stride;
    long long B;
    int L;
    const int *lengths;
    int T, hop, pad, n_fft;
    const float *window;    // [n_fft], periodic Hann centre-padded, pre-scaled by 0.5
    const float2 *tw;       // [32][32]  tw[k1*32 + lane] = exp(-2 pi i k1 lane / 1024)
    const float2 *tw_post;  // [32]      exp(-2 pi i lane / 2048)            (split mode)
    int n_mels, n_freq;
    const int *mel_lo;   // first bin of row m
    const int *mel_cnt;  // number of bins of row m
    const int *mel_ptr;  // offset of row m in mel_w
    const float *mel_w;
    float *out_mel, *out_a, *out_b;
    int spec_kind;
    float mag_eps;
    int power;
    int log_kind;
    float log_arg;
    int has_lo, has_hi, norm;
    float lo, hi, norm_scale;
    long long tasks_per_clip, n_tasks;
};

constexpr int kWarpsPerCta = 4;
constexpr int kBufStride = 33;                   // float2 per transposed row (+1 pad)
constexpr int kWarpBufElems = 32 * kBufStride;  // float2 per warp
constexpr int kMagStride = 520;                  // floats between the two frames of a pair tile

__device__ __forceinline__ int frames_of(int Li, int n_fft, int hop, int pad) {
    int span = Li + 2 * pad - n_fft;
    return span < 0 ? 0 : span / hop + 1;
}

__device__ __forceinline__ int reflect_index(int i, int Li) {
    if (i < 0) i = -i;
    if (i >= Li) i = 2 * (Li - 1) - i;
    return min(max(i, 0), Li - 1);
}

__device__ __forceinline__ float epilogue(float x, const KParams &p) {
    float y = x;
    if (p.log_kind == B200MEL_LOG_LN_OFFSET)
        y = logf(x + p.log_arg);
    else if (p.log_kind == B200MEL_LOG_LN_FLOOR)
        y = logf(fmaxf(x, p.log_arg));
    else if (p.log_kind == B200MEL_LOG_LOG10_FLOOR)
        y = log10f(fmaxf(x, p.log_arg));
    if (p.has_lo) y = fmaxf(y, p.lo);
    if (p.has_hi) y = fminf(y, p.hi);
    if (p.norm) y = (y - p.lo) * p.norm_scale - 1.0f;
    return y;
}

__device__ __forceinline__ float magnitude(float re, float im, const KParams &p) {
    float sq = fmaf(re, re, im * im);
    if (p.power == 2) return sq;
    return sqrtf(sq + p.mag_eps);
}

// Store one spectrum bin (frame t of clip b) according to spec_kind.
__device__ __forceinline__ void store_spec(const KParams &p, long long b, int k, int t, float re, float im,
                                           float mag) {
    long long o = (b * p.n_freq + k) * (long long)p.T + t;
    if (p.spec_kind == B200MEL_SPEC_RE_IM) {
        p.out_a[o] = re;
        p.out_b[o] = im;
    } else {
        p.out_a[o] = mag;
        if (p.spec_kind == B200MEL_SPEC_MAG_PHASE) p.out_b[o] = atan2f(im, re);
    }
}

// kPair = true : n_fft 1024, the warp transforms frames (2q, 2q+1) of its clip
// kPair = false: n_fft 2048, the warp transforms frame q of its clip
template <bool kPair>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 4) logmel_warp_kernel(const KParams p) {
    extern __shared__ float2 smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 *buf = smem + warp * kWarpBufElems;
    float *tile = reinterpret_cast<float *>(buf);

    const long long task = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (task >= p.n_tasks) return;  // warps are independent: no block-level barrier below
    const long long b = task / p.tasks_per_clip;
    const int q = (int)(task - b * p.tasks_per_clip);

    const int Li = p.lengths ? min(p.lengths[b], p.L) : p.L;
    const int Ti = p.lengths ? min(frames_of(Li, p.n_fft, p.hop, p.pad), p.T) : p.T;
    const int t0 = kPair ? 2 * q : q;
    const bool valid0 = t0 < Ti;
    const bool valid1 = kPair && (t0 + 1 < Ti);
    const float *row = p.wav + b * p.row_stride;

    float2 a[32];

    if (valid0) {
        // ------------------------------------------------------------------ load + window
        const int s0 = t0 * p.hop - p.pad;
        if (kPair) {
            const int s1 = s0 + p.hop;
            const int last = valid1 ? s1 : s0;
            const bool interior = (s0 >= 0) && (last + 1024 <= Li);
            if (interior) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = 32 * j + lane;
                    const float w = __ldg(p.window + n);
                    a[j].x = __ldg(row + s0 + n) * w;
                    a[j].y = valid1 ? __ldg(row + s1 + n) * w : 0.0f;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = 32 * j + lane;
                    const float w = __ldg(p.window + n);
                    a[j].x = __ldg(row + reflect_index(s0 + n, Li)) * w;
                    a[j].y = valid1 ? __ldg(row + reflect_index(s1 + n, Li)) * w : 0.0f;
                }
            }
        } else {
            const bool interior = (s0 >= 0) && (s0 + 2048 <= Li);
            if (interior) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = 64 * j + 2 * lane;
                    a[j].x = __ldg(row + s0 + n) * __ldg(p.window + n);
                    a[j].y = __ldg(row + s0 + n + 1) * __ldg(p.window + n + 1);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = 64 * j + 2 * lane;
                    a[j].x = __ldg(row + reflect_index(s0 + n, Li)) * __ldg(p.window + n);
                    a[j].y = __ldg(row + reflect_index(s0 + n + 1, Li)) * __ldg(p.window + n + 1);
                }
            }
        }

        // ------------------------------------------------------------------ 1024-point complex FFT
        // pass 1: lane = n2, FFT over n1 -> Y[k1] at a[pos(k1)]
        fft32(a);
        // inter-pass twiddle W_1024^{n2 k1}, then transpose through shared memory
        static_for<0, 32>([&](auto k1_) {
            constexpr int k1 = decltype(k1_)::value;
            float2 v = a[fft32_pos(k1)];
            if constexpr (k1 > 0) v = cmul(v, __ldg(p.tw + k1 * 32 + lane));
            buf[k1 * kBufStride + lane] = v;
        });
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) a[n2] = buf[lane * kBufStride + n2];
        __syncwarp();  // buf is reused as the magnitude tile below
        // pass 2: lane = k1, FFT over n2 -> Z[k1 + 32 k2] at a[pos(k2)]
        fft32(a);

        // ------------------------------------------------------------------ real-input separation
        const int partner = (32 - lane) & 31;
        float2 wl = make_float2(1.f, 0.f);
        if (!kPair) wl = __ldg(p.tw_post + lane);
        static_for<0, 16>([&](auto k2_) {
            constexpr int k2 = decltype(k2_)::value;
            const float2 A = a[fft32_pos(k2)];
            // value my reader needs: lane 0 is read by itself and wants Z[32*((32-k2)&31)],
            // lane s != 0 is read by lane 32-s which wants my slot 31-k2.
            const float2 give0 = a[fft32_pos((32 - k2) & 31)];
            const float2 give1 = a[fft32_pos(31 - k2)];
            float2 give = lane == 0 ? give0 : give1;
            float2 Bv;
            Bv.x = __shfl_sync(0xffffffffu, give.x, partner);
            Bv.y = __shfl_sync(0xffffffffu, give.y, partner);
            const int k = lane + 32 * k2;
            // E = A + conj(B), O = (A - conj(B)) / i   (the 1/2 is folded into the window)
            const float2 E = make_float2(A.x + Bv.x, A.y - Bv.y);
            const float2 O = make_float2(A.y + Bv.y, Bv.x - A.x);
            if (kPair) {
                const float m0 = magnitude(E.x, E.y, p), m1 = magnitude(O.x, O.y, p);
                tile[k] = m0;
                tile[kMagStride + k] = m1;
                if (p.spec_kind) {
                    store_spec(p, b, k, t0, E.x, E.y, m0);
                    if (valid1) store_spec(p, b, k, t0 + 1, O.x, O.y, m1);
                }
            } else {
                // X[k] = E + W_2048^k O,  X[1024-k] = conj(E - W_2048^k O),  W_2048^k = wl * W_64^{k2}
                constexpr float w64c = TwConst::c64[k2], w64s = TwConst::s64[k2];
                const float2 w64 = make_float2(w64c, w64s);
                const float2 P = cmul(O, cmul(wl, w64));
                const float2 X0 = cadd(E, P), X1 = csub(E, P);
                const float m0 = magnitude(X0.x, X0.y, p), m1 = magnitude(X1.x, X1.y, p);
                tile[k] = m0;
                tile[1024 - k] = m1;
                if (p.spec_kind) {
                    store_spec(p, b, k, t0, X0.x, X0.y, m0);
                    store_spec(p, b, 1024 - k, t0, X1.x, -X1.y, m1);
                }
            }
        });
        if (lane == 0) {  // bin 512 (k1 = 0, k2 = 16): its own partner
            const float2 A = a[fft32_pos(16)];
            if (kPair) {
                const float re0 = 2.f * A.x, re1 = 2.f * A.y;
                const float m0 = magnitude(re0, 0.f, p), m1 = magnitude(re1, 0.f, p);
                tile[512] = m0;
                tile[kMagStride + 512] = m1;
                if (p.spec_kind) {
                    store_spec(p, b, 512, t0, re0, 0.f, m0);
                    if (valid1) store_spec(p, b, 512, t0 + 1, re1, 0.f, m1);
                }
            } else {  // E = 2 Re A, O = 2 Im A, W_2048^512 = -i  ->  X[512] = 2 (Re A - i Im A)
                const float re = 2.f * A.x, im = -2.f * A.y;
                const float m0 = magnitude(re, im, p);
                tile[512] = m0;
                if (p.spec_kind) store_spec(p, b, 512, t0, re, im, m0);
            }
        }
        __syncwarp();
    }

    // ---------------------------------------------------------------------- frames past the clip's end
    if (!valid0 || (kPair && !valid1)) {
        // only reachable with `lengths` (or the odd last frame of a pair): zero-fill, as pad_collate_fn
        // zero-pads per-item features (data/dataset.py:230-250).
        const int tz0 = valid0 ? t0 + 1 : t0;
        const int tz1 = kPair ? t0 + 1 : t0;
        for (int t = tz0; t <= tz1 && t < p.T; ++t) {
            if (p.out_mel)
                for (int m = lane; m < p.n_mels; m += 32) p.out_mel[(b * p.n_mels + m) * (long long)p.T + t] = 0.f;
            if (p.spec_kind)
                for (int k = lane; k < p.n_freq; k += 32) {
                    long long o = (b * p.n_freq + k) * (long long)p.T + t;
                    p.out_a[o] = 0.f;
                    if (p.spec_kind != B200MEL_SPEC_MAG) p.out_b[o] = 0.f;
                }
        }
        if (!valid0) return;
    }

    // ---------------------------------------------------------------------- banded mel + log epilogue
    if (p.out_mel) {
        for (int m = lane; m < p.n_mels; m += 32) {
            const int lo = __ldg(p.mel_lo + m), cnt = __ldg(p.mel_cnt + m);
            const float *w = p.mel_w + __ldg(p.mel_ptr + m);
            float acc0 = 0.f, acc1 = 0.f;
            for (int j = 0; j < cnt; ++j) {
                const float wj = __ldg(w + j);
                acc0 = fmaf(wj, tile[lo + j], acc0);
                if (kPair) acc1 = fmaf(wj, tile[kMagStride + lo + j], acc1);
            }
            float *o = p.out_mel + (b * p.n_mels + m) * (long long)p.T + t0;
            o[0] = epilogue(acc0, p);
            if (kPair && valid1) o[1] = epilogue(acc1, p);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
static int cuda_fail(cudaError_t e, const char *what) {
    return fail(B200MEL_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// --- librosa 0.8.0 filters.mel restated (Slaney/HTK scales), all intermediate math in double ---
static double hz_to_mel(double f, bool htk) {
    if (htk) return 2595.0 * log10(1.0 + f / 700.0);
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0;
    const double min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m, bool htk) {
    if (htk) return 700.0 * (pow(10.0, m / 2595.0) - 1.0);
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0;
    const double min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}
// numpy.linspace(a, b, n): arange(n) * step + a with the last point pinned to b
static std::vector<double> linspace(double a, double b, int n) {
    std::vector<double> y(n);
    const double step = n > 1 ? (b - a) / (n - 1) : 0.0;
    for (int i = 0; i < n; ++i) y[i] = i * step + a;
    if (n > 1) y[n - 1] = b;
    return y;
}
static void build_filterbank(int sr, int n_fft, int n_mels, double fmin, double fmax, bool htk, bool slaney_norm,
                             float *out) {
    const int F = n_fft / 2 + 1;
    std::vector<double> fftfreqs = linspace(0.0, sr / 2.0, F);
    std::vector<double> mels = linspace(hz_to_mel(fmin, htk), hz_to_mel(fmax, htk), n_mels + 2);
    std::vector<double> mel_f(n_mels + 2);
    for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(mels[i], htk);
    for (int i = 0; i < n_mels; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        for (int k = 0; k < F; ++k) {
            const double lower = -(mel_f[i] - fftfreqs[k]) / fd0;
            const double upper = (mel_f[i + 2] - fftfreqs[k]) / fd1;
            double v = lower < upper ? lower : upper;
            if (!(v > 0.0)) v = 0.0;
            float w = (float)v;  // librosa stores the triangle in a float32 array first ...
            if (slaney_norm) w = (float)((double)w * enorm);  // ... then scales it in place
            out[(size_t)i * F + k] = w;
        }
    }
}
static void build_window(int win, int n_fft, float *out) {
    const double two_pi = 6.283185307179586476925286766559;
    for (int n = 0; n < n_fft; ++n) out[n] = 0.f;
    const int lpad = (n_fft - win) / 2;
    for (int n = 0; n < win; ++n) out[lpad + n] = (float)(0.5 - 0.5 * cos(two_pi * n / win));
}

}  // namespace b200mel

using namespace b200mel;

struct b200mel_plan {
    b200mel_config cfg;
    int device;
    int pad, n_freq;
    bool pair;
    // device tables
    float *d_window = nullptr;
    float2 *d_tw = nullptr, *d_tw_post = nullptr;
    int *d_mel_lo = nullptr, *d_mel_cnt = nullptr, *d_mel_ptr = nullptr;
    float *d_mel_w = nullptr;
    // staging for forward_host
    float *d_stage_in = nullptr, *d_stage_out = nullptr;
    size_t stage_in_bytes = 0, stage_out_bytes = 0;
};

static void free_mel_tables(b200mel_plan *pl) {
    cudaFree(pl->d_mel_lo);
    cudaFree(pl->d_mel_cnt);
    cudaFree(pl->d_mel_ptr);
    cudaFree(pl->d_mel_w);
    pl->d_mel_lo = pl->d_mel_cnt = pl->d_mel_ptr = nullptr;
    pl->d_mel_w = nullptr;
}

// dense (n_mels x F) -> banded rows [lo, lo+cnt) and upload
static int upload_filterbank(b200mel_plan *pl, const float *W, int n_mels, int F) {
    std::vector<int> lo(n_mels), cnt(n_mels), ptr(n_mels);
    std::vector<float> w;
    for (int m = 0; m < n_mels; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < F; ++k)
            if (W[(size_t)m * F + k] != 0.f) {
                if (first < 0) first = k;
                last = k;
            }
        lo[m] = first < 0 ? 0 : first;
        cnt[m] = first < 0 ? 0 : last - first + 1;
        ptr[m] = (int)w.size();
        for (int k = 0; k < cnt[m]; ++k) w.push_back(W[(size_t)m * F + lo[m] + k]);
    }
    if (w.empty()) w.push_back(0.f);
    free_mel_tables(pl);
    cudaError_t e;
    if ((e = cudaMalloc(&pl->d_mel_lo, n_mels * sizeof(int))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMalloc(&pl->d_mel_cnt, n_mels * sizeof(int))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMalloc(&pl->d_mel_ptr, n_mels * sizeof(int))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMalloc(&pl->d_mel_w, w.size() * sizeof(float))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    cudaMemcpy(pl->d_mel_lo, lo.data(), n_mels * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(pl->d_mel_cnt, cnt.data(), n_mels * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(pl->d_mel_ptr, ptr.data(), n_mels * sizeof(int), cudaMemcpyHostToDevice);
    e = cudaMemcpy(pl->d_mel_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(filterbank)");
    return B200MEL_OK;
}

extern "C" {

int b200mel_version(void) { return B200MEL_VERSION; }
const char *b200mel_last_error(void) { return g_err.c_str(); }
int64_t b200mel_launch_count(void) { return g_launches.load(); }

int b200mel_mel_filterbank(int32_t sr, int32_t n_fft, int32_t n_mels, double fmin, double fmax, int32_t mel_scale,
                           int32_t mel_norm, float *out) {
    if (!out || sr <= 0 || n_fft < 2 || n_mels <= 0) return fail(B200MEL_EINVAL, "mel_filterbank: bad argument");
    if (fmax <= 0.0) fmax = sr / 2.0;
    if (fmin < 0.0 || fmin >= fmax) return fail(B200MEL_EINVAL, "mel_filterbank: need 0 <= fmin < fmax");
    build_filterbank(sr, n_fft, n_mels, fmin, fmax, mel_scale == B200MEL_MEL_HTK, mel_norm == B200MEL_NORM_SLANEY,
                     out);
    return B200MEL_OK;
}

int b200mel_hann_window(int32_t win_length, int32_t n_fft, float *out) {
    if (!out || win_length <= 0 || n_fft < win_length) return fail(B200MEL_EINVAL, "hann_window: need 0 < win <= n_fft");
    build_window(win_length, n_fft, out);
    return B200MEL_OK;
}

static int frames_host(const b200mel_plan *pl, int64_t L, int64_t *T) {
    int64_t span = L + 2 * (int64_t)pl->pad - pl->cfg.n_fft;
    *T = span < 0 ? 0 : span / pl->cfg.hop_length + 1;
    return B200MEL_OK;
}

int b200mel_out_frames(const b200mel_plan *plan, int64_t L, int64_t *T) {
    if (!plan || !T || L < 0) return fail(B200MEL_EINVAL, "out_frames: bad argument");
    return frames_host(plan, L, T);
}

int b200mel_plan_create(const b200mel_config *cfg, b200mel_plan **out) {
    if (!cfg || !out) return fail(B200MEL_EINVAL, "plan_create: null argument");
    if (cfg->struct_size != (int32_t)sizeof(b200mel_config))
        return fail(B200MEL_EINVAL, "plan_create: struct_size mismatch (ABI version skew)");
    if (cfg->n_fft != 1024 && cfg->n_fft != 2048)
        return fail(B200MEL_EUNSUP, "plan_create: n_fft must be 1024 or 2048 in this build");
    if (cfg->win_length <= 0 || cfg->win_length > cfg->n_fft)
        return fail(B200MEL_EINVAL, "plan_create: need 0 < win_length <= n_fft (models/transforms.py:28)");
    if (cfg->hop_length <= 0) return fail(B200MEL_EINVAL, "plan_create: hop_length must be positive");
    if (cfg->sample_rate <= 0 || cfg->n_mels < 0 || cfg->n_mels > 4096)
        return fail(B200MEL_EINVAL, "plan_create: bad sample_rate / n_mels");
    if (cfg->pad_mode != B200MEL_PAD_CENTER && cfg->pad_mode != B200MEL_PAD_HIFI)
        return fail(B200MEL_EINVAL, "plan_create: bad pad_mode");
    if (cfg->power != 1 && cfg->power != 2) return fail(B200MEL_EINVAL, "plan_create: power must be 1 or 2");
    double fmax = cfg->fmax > 0.f ? (double)cfg->fmax : cfg->sample_rate / 2.0;
    if (cfg->n_mels > 0 && (cfg->fmin < 0.f || (double)cfg->fmin >= fmax))
        return fail(B200MEL_EINVAL, "plan_create: need 0 <= fmin < fmax");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(B200MEL_ENODEV, "plan_create: no CUDA device (there is no CPU fallback)");
    }
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail(B200MEL_ENODEV, "plan_create: device is not compute capability 10.x (library is built for sm_100a only)");

    b200mel_plan *pl = new (std::nothrow) b200mel_plan();
    if (!pl) return fail(B200MEL_ENOMEM, "plan_create: out of host memory");
    pl->cfg = *cfg;
    pl->device = dev;
    pl->pair = cfg->n_fft == 1024;
    pl->n_freq = cfg->n_fft / 2 + 1;
    pl->pad = cfg->pad_mode == B200MEL_PAD_CENTER ? cfg->n_fft / 2 : (cfg->n_fft - cfg->hop_length) / 2;
    if (pl->pad < 0) pl->pad = 0;

    const int N = cfg->n_fft;
    std::vector<float> win(N);
    build_window(cfg->win_length, N, win.data());
    for (auto &w : win) w *= 0.5f;  // exact; the separation pass omits its 1/2
    std::vector<float2> tw(32 * 32), twp(32);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k1 = 0; k1 < 32; ++k1)
        for (int l = 0; l < 32; ++l) {
            double ang = two_pi * (double)(k1 * l) / 1024.0;
            tw[k1 * 32 + l] = make_float2((float)cos(ang), (float)-sin(ang));
        }
    for (int l = 0; l < 32; ++l) {
        double ang = two_pi * l / 2048.0;
        twp[l] = make_float2((float)cos(ang), (float)-sin(ang));
    }
    int rc = B200MEL_OK;
    do {
        if ((e = cudaMalloc(&pl->d_window, N * sizeof(float))) != cudaSuccess) { rc = cuda_fail(e, "cudaMalloc"); break; }
        if ((e = cudaMalloc(&pl->d_tw, tw.size() * sizeof(float2))) != cudaSuccess) { rc = cuda_fail(e, "cudaMalloc"); break; }
        if ((e = cudaMalloc(&pl->d_tw_post, twp.size() * sizeof(float2))) != cudaSuccess) { rc = cuda_fail(e, "cudaMalloc"); break; }
        cudaMemcpy(pl->d_window, win.data(), N * sizeof(float), cudaMemcpyHostToDevice);
        cudaMemcpy(pl->d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice);
        e = cudaMemcpy(pl->d_tw_post, twp.data(), twp.size() * sizeof(float2), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMemcpy(tables)"); break; }
        if (cfg->n_mels > 0) {
            std::vector<float> W((size_t)cfg->n_mels * pl->n_freq);
            build_filterbank(cfg->sample_rate, N, cfg->n_mels, cfg->fmin, fmax, cfg->mel_scale == B200MEL_MEL_HTK,
                             cfg->mel_norm == B200MEL_NORM_SLANEY, W.data());
            rc = upload_filterbank(pl, W.data(), cfg->n_mels, pl->n_freq);
            if (rc) break;
        }
        const int smem = kWarpsPerCta * kWarpBufElems * (int)sizeof(float2);
        e = cudaFuncSetAttribute(logmel_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(logmel_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { rc = cuda_fail(e, "cudaFuncSetAttribute (is the library built for this GPU?)"); break; }
    } while (0);
    if (rc) {
        b200mel_plan_destroy(pl);
        return rc;
    }
    *out = pl;
    return B200MEL_OK;
}

int b200mel_plan_set_filterbank(b200mel_plan *plan, const float *weights, int32_t n_mels, int32_t n_freq) {
    if (!plan || !weights) return fail(B200MEL_EINVAL, "set_filterbank: null argument");
    if (n_freq != plan->n_freq || n_mels <= 0 || n_mels > 4096)
        return fail(B200MEL_EINVAL, "set_filterbank: shape must be (n_mels, n_fft/2+1)");
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(plan->device);
    int rc = upload_filterbank(plan, weights, n_mels, n_freq);
    if (!rc) plan->cfg.n_mels = n_mels;
    cudaSetDevice(prev);
    return rc;
}

int b200mel_plan_destroy(b200mel_plan *pl) {
    if (!pl) return B200MEL_OK;
    cudaFree(pl->d_window);
    cudaFree(pl->d_tw);
    cudaFree(pl->d_tw_post);
    free_mel_tables(pl);
    cudaFree(pl->d_stage_in);
    cudaFree(pl->d_stage_out);
    delete pl;
    return B200MEL_OK;
}

int b200mel_forward(const b200mel_plan *pl, const float *wav, int64_t B, int64_t L, int64_t row_stride,
                    const int32_t *lengths, const b200mel_epilogue *epi, float *out_mel, int32_t spec_kind,
                    float *out_a, float *out_b, void *stream) {
    if (!pl) return fail(B200MEL_EINVAL, "forward: null plan");
    if (B < 0 || L < 0) return fail(B200MEL_EINVAL, "forward: negative shape");
    if (B == 0) return B200MEL_OK;
    if (!wav) return fail(B200MEL_EINVAL, "forward: null wav");
    if (row_stride < L) return fail(B200MEL_EINVAL, "forward: row_stride < L");
    if (L > 0x3fffffff) return fail(B200MEL_EINVAL, "forward: clip longer than 2^30 samples");
    if (L <= pl->pad)
        return fail(B200MEL_EINVAL, "forward: reflect padding needs L > pad (torch raises the same for F.pad reflect)");
    if (spec_kind < 0 || spec_kind > 3) return fail(B200MEL_EINVAL, "forward: bad spec_kind");
    if (spec_kind && !out_a) return fail(B200MEL_EINVAL, "forward: spec_kind set but out_a is null");
    if ((spec_kind == B200MEL_SPEC_MAG_PHASE || spec_kind == B200MEL_SPEC_RE_IM) && !out_b)
        return fail(B200MEL_EINVAL, "forward: spec_kind needs out_b");
    if (!out_mel && !spec_kind) return fail(B200MEL_EINVAL, "forward: no output requested");
    if (out_mel && pl->cfg.n_mels == 0) return fail(B200MEL_EINVAL, "forward: plan has no filterbank (n_mels = 0)");
    if (out_mel && !epi) return fail(B200MEL_EINVAL, "forward: mel output needs an epilogue");
    if (epi && epi->struct_size != (int32_t)sizeof(b200mel_epilogue))
        return fail(B200MEL_EINVAL, "forward: epilogue struct_size mismatch");
    if (epi && (epi->log_kind < 0 || epi->log_kind > 3)) return fail(B200MEL_EINVAL, "forward: bad log_kind");
    if (epi && epi->norm_mel && !(epi->has_clamp_lo && epi->has_clamp_hi && epi->clamp_hi > epi->clamp_lo))
        return fail(B200MEL_EINVAL, "forward: norm_mel needs clamp_lo < clamp_hi");

    int64_t T = 0;
    frames_host(pl, L, &T);
    if (T <= 0) return fail(B200MEL_EINVAL, "forward: clip shorter than one frame");
    if (T > 0x7fffffff) return fail(B200MEL_EINVAL, "forward: too many frames");

    KParams p;
    memset(&p, 0, sizeof(p));
    p.wav = wav;
    p.row_stride = row_stride;
    p.B = B;
    p.L = (int)L;
    p.lengths = lengths;
    p.T = (int)T;
    p.hop = pl->cfg.hop_length;
    p.pad = pl->pad;
    p.n_fft = pl->cfg.n_fft;
    p.window = pl->d_window;
    p.tw = pl->d_tw;
    p.tw_post = pl->d_tw_post;
    p.n_mels = pl->cfg.n_mels;
    p.n_freq = pl->n_freq;
    p.mel_lo = pl->d_mel_lo;
    p.mel_cnt = pl->d_mel_cnt;
    p.mel_ptr = pl->d_mel_ptr;
    p.mel_w = pl->d_mel_w;
    p.out_mel = out_mel;
    p.out_a = out_a;
    p.out_b = out_b;
    p.spec_kind = spec_kind;
    p.mag_eps = pl->cfg.mag_eps;
    p.power = pl->cfg.power;
    if (epi) {
        p.log_kind = epi->log_kind;
        p.log_arg = epi->log_arg;
        p.has_lo = epi->has_clamp_lo;
        p.lo = epi->clamp_lo;
        p.has_hi = epi->has_clamp_hi;
        p.hi = epi->clamp_hi;
        p.norm = epi->norm_mel;
        if (p.norm) p.norm_scale = 2.0f / (p.hi - p.lo);
    }
    p.tasks_per_clip = pl->pair ? (T + 1) / 2 : T;
    p.n_tasks = p.tasks_per_clip * B;
    const long long n_cta = (p.n_tasks + kWarpsPerCta - 1) / kWarpsPerCta;
    if (n_cta > 0x7fffffffLL) return fail(B200MEL_EINVAL, "forward: batch too large for one launch");

    const int smem = kWarpsPerCta * kWarpBufElems * (int)sizeof(float2);
    cudaStream_t st = (cudaStream_t)stream;
    if (pl->pair)
        logmel_warp_kernel<true><<<(unsigned)n_cta, kWarpsPerCta * 32, smem, st>>>(p);
    else
        logmel_warp_kernel<false><<<(unsigned)n_cta, kWarpsPerCta * 32, smem, st>>>(p);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    return B200MEL_OK;
}

int b200mel_forward_host(b200mel_plan *pl, const float *wav_host, int64_t B, int64_t L, int64_t row_stride,
                         const b200mel_epilogue *epi, float *out_mel_host, void *stream) {
    if (!pl || !wav_host || !out_mel_host) return fail(B200MEL_EINVAL, "forward_host: null argument");
    if (B <= 0 || L <= 0 || row_stride < L) return fail(B200MEL_EINVAL, "forward_host: bad shape");
    int64_t T = 0;
    frames_host(pl, L, &T);
    if (T <= 0) return fail(B200MEL_EINVAL, "forward_host: clip shorter than one frame");
    const size_t in_bytes = (size_t)B * row_stride * sizeof(float);
    const size_t out_bytes = (size_t)B * pl->cfg.n_mels * T * sizeof(float);
    cudaError_t e;
    cudaStream_t st = (cudaStream_t)stream;
    if (in_bytes > pl->stage_in_bytes) {
        cudaStreamSynchronize(st);
        cudaFree(pl->d_stage_in);
        pl->d_stage_in = nullptr;
        pl->stage_in_bytes = 0;
        if ((e = cudaMalloc(&pl->d_stage_in, in_bytes)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(stage_in)");
        pl->stage_in_bytes = in_bytes;
    }
    if (out_bytes > pl->stage_out_bytes) {
        cudaStreamSynchronize(st);
        cudaFree(pl->d_stage_out);
        pl->d_stage_out = nullptr;
        pl->stage_out_bytes = 0;
        if ((e = cudaMalloc(&pl->d_stage_out, out_bytes)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(stage_out)");
        pl->stage_out_bytes = out_bytes;
    }
    if ((e = cudaMemcpyAsync(pl->d_stage_in, wav_host, in_bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess)
        return cuda_fail(e, "cudaMemcpyAsync(H2D)");
    int rc = b200mel_forward(pl, pl->d_stage_in, B, L, row_stride, nullptr, epi, pl->d_stage_out, B200MEL_SPEC_NONE,
                             nullptr, nullptr, stream);
    if (rc) return rc;
    if ((e = cudaMemcpyAsync(out_mel_host, pl->d_stage_out, out_bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
        return cuda_fail(e, "cudaMemcpyAsync(D2H)");
    return B200MEL_OK;
}

}  // extern "C"
