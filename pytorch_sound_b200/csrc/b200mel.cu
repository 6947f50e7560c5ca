// b200mel.cu — fused STFT -> |.| -> mel -> log kernel for sm_100a and its C ABI.
//
// Replaces the op chain of pytorch_sound's spectral modules
//   reflect-pad -> conv1d-DFT | torch.stft -> sqrt(re^2+im^2) -> mel matmul -> log -> clamp
// (models/transforms.py:53-69, 231-244, 297-311, 351-366; interface/hifi_gan.py:46-63)
// by ONE launch per clip batch.  See DESIGN.md for the data layout and roofline.
//
// The kernel itself (persistent warps, TMA-staged samples, register-resident radix-32 FFT passes,
// banded mel from shared memory, log epilogue in registers) lives in logmel_kernel.cuh; this file is the
// host side: librosa/scipy restatements, plan construction, launch and the extern "C" surface.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/b200mel.h"
#include "fft32.cuh"
#include "logmel_kernel.cuh"
#include "logmel_fast.cuh"
#include "mel_tc.cuh"
#include "spec_kernel.cuh"
#include "stft_tc.cuh"
#include "stft_tc_tables.h"
#include "stc_probe.cuh"
#include "wave_ops.cuh"

namespace b200mel {

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};
static const bool g_pdl = [] {  // B200MEL_PDL=0 disables programmatic dependent launch (A/B measurements)
    const char *e = getenv("B200MEL_PDL");
    return !(e && e[0] == '0');
}();
static const bool g_table_window = getenv("B200MEL_TABLE_WINDOW") != nullptr;  // A/B: read the Hann table instead
#ifndef B200MEL_TC_DEFAULT
#define B200MEL_TC_DEFAULT 0  // 2: auto (eligible plans, launches that fill the GPU); 0: only when B200MEL_TC is set
#endif
static long long *g_dbg = nullptr;  // device buffer for -DB200MEL_PHASE_TIMING builds (b200mel_debug_set_buffer)

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
    int num_sms;
    int pad, n_freq;  // n_freq = logical n_fft / 2 + 1 (rows of the filterbank and of the spectrum outputs)
    int phys_n_fft;   // transform the kernels run: 1024 for every power-of-two n_fft <= 1024, else 2048
    int bin_step;     // phys_n_fft / n_fft: logical bin k is physical bin k * bin_step
    int phys_n_freq;  // phys_n_fft / 2 + 1
    bool pair;        // phys_n_fft == 1024: two frames per complex FFT
    int pair_frames;  // frames per task (2 in pair mode unless hop > n_fft, else 1)
    // device tables
    float *d_window = nullptr;
    float2 *d_tw = nullptr, *d_tw_post = nullptr;
    MelEntry *d_mel_entries = nullptr;
    float *d_mel_w = nullptr;
    int mel_rounds = 0, mel_w_len = 0;
    int top_groups = 16;  // 32-bin groups the pair kernel separates: 12 when the filterbank ends below bin 384
    int round_groups[kMaxMelRounds] = {0}, round_wbase[kMaxMelRounds] = {0};
    // dense filterbank as uploaded (logical bins), kept for the tensor-core operator
    std::vector<float> W_dense;
    // tcgen05 mel GEMM (mel_tc.cuh): bf16 hi / lo limbs of the filterbank in the UMMA K-major core-matrix layout
    uint16_t *d_tc_bhi = nullptr, *d_tc_blo = nullptr;
    int tc_k_steps = 0, tc_n_pad = 0, tc_b_bytes = 0, tc_smem = 0;
    std::mutex tc_mu;
    // tcgen05 STFT kernel (stft_tc.cuh): constant operands + mel schedule blob; null when the plan is not eligible
    // (n_fft = win_length = 1024, hop 256, filterbank below bin 384, <= 128 mel rows)
    unsigned char *d_stc = nullptr;
    int stc_mel_bytes = 0, stc_smem = 0;
    // shared-memory layout
    int off_window = 0, off_entries = 0, off_melw = 0, off_bar = 0, off_regions = 0, region_bytes = 0, stage_bytes = 0;
    int n_warps = 0, smem_bytes = 0;
    // spectrum-output kernel (spec_kernel.cuh): tw | window | mbarriers | slots | warp regions | per group: tile A (| tile B);
    // one layout per output kind (index = B200MEL_SPEC_*): as many warp groups (4 warps in pair mode, 8 in split mode)
    // as fit next to their tiles — |X| only: 4 groups, two outputs: 3 (pair mode, hop 256)
    struct SpecLayout {
        int groups = 0, warps = 0, region = 0, off_bar = 0, off_slots = 0, off_regions = 0, off_tiles = 0, smem = 0;
    } sp[4];
    // staging for forward_host: one (input, output) pair per CUDA stream that has called it, so calls on different
    // streams overlap (copy of one batch under the kernel / read-back of another); calls on one stream are ordered
    // by the stream itself.  Guarded by host_mu.
    struct HostStage {
        void *stream = nullptr;
        float *d_in = nullptr, *d_out = nullptr;
        size_t in_bytes = 0, out_bytes = 0;
        bool used = false;
    };
    static constexpr int kHostStages = 8;
    HostStage host_stage[kHostStages];
    std::mutex host_mu;
};

constexpr int kMaxSmem = 232448;
constexpr int kDefaultWarps = kMaxWarps;  // 227 KB opt-in limit per CTA on sm_100

static void free_mel_tables(b200mel_plan *pl) {
    cudaFree(pl->d_mel_entries);
    cudaFree(pl->d_mel_w);
    pl->d_mel_entries = nullptr;
    pl->d_mel_w = nullptr;
    pl->mel_rounds = pl->mel_w_len = 0;
}

// Shared-memory carve-up (must match logmel_kernel.cuh): tw | window | mel entries | mel weights | mbarriers | warp regions
static int layout_smem(b200mel_plan *pl) {
    const int n_fft = pl->phys_n_fft;
    const int span = n_fft + (pl->pair_frames == 2 ? pl->cfg.hop_length : 0);
    pl->stage_bytes = ((span + 8) * 4 + 15) & ~15;
    int region = kStageOff + pl->stage_bytes;
    if (region < kXposeBytes) region = kXposeBytes;
    pl->region_bytes = (region + 127) & ~127;
    pl->off_window = 32 * 32 * 8;
    pl->off_entries = pl->off_window + n_fft * 4;
    pl->off_melw = pl->off_entries + pl->mel_rounds * 32 * (int)sizeof(MelEntry);
    pl->off_bar = pl->off_melw + pl->mel_w_len * 4;
    pl->off_regions = (pl->off_bar + (kMaxWarps + 1) * 8 + 127) & ~127;  // per-warp mbarriers + the table mbarrier
    int n_warps = (kMaxSmem - pl->off_regions) / pl->region_bytes;
    int cap = kDefaultWarps;
    if (const char *env = getenv("B200MEL_WARPS")) cap = atoi(env);  // debugging knob: fewer warps per CTA
    if (cap < 1) cap = 1;
    if (cap > kMaxWarps) cap = kMaxWarps;
    if (n_warps > cap) n_warps = cap;
    if (n_warps < 1)
        return fail(B200MEL_EUNSUP, "plan: hop_length / filterbank too large for the shared-memory staging of this build");
    pl->n_warps = n_warps;
    pl->smem_bytes = pl->off_regions + n_warps * pl->region_bytes;
    // spectrum-output kernel: per output kind as many warp groups as fit in shared memory (at most 16 warps)
    const int sp_region = (std::max(kXposeBytes, pl->stage_bytes) + 127) & ~127;  // stage overlaid on the transpose buffer
    const int gw = pl->pair ? 4 : 8;
    for (int kind = 1; kind <= 3; ++kind) {
        const int n_tiles = kind == B200MEL_SPEC_MAG ? 1 : 2;
        b200mel_plan::SpecLayout best;
        for (int groups = 16 / gw; groups >= 1; --groups) {
            b200mel_plan::SpecLayout L;
            L.groups = groups, L.warps = groups * gw, L.region = sp_region;
            L.off_bar = pl->off_window + n_fft * 4;
            L.off_slots = L.off_bar + L.warps * 8;
            L.off_regions = (L.off_slots + 2 * L.warps * (int)sizeof(SpecSlot) + 127) & ~127;  // two sets: this round's and the one being written out
            L.off_tiles = L.off_regions + L.warps * L.region;
            L.smem = L.off_tiles + groups * n_tiles * pl->phys_n_freq * 8 * 4;
            if (L.smem <= kMaxSmem) {
                best = L;
                break;
            }
        }
        if (!best.warps)
            return fail(B200MEL_EUNSUP, "plan: hop_length too large for the shared-memory staging of the spectrum kernel");
        pl->sp[kind] = best;
    }
    return B200MEL_OK;
}

// dense (n_mels x F) -> lane-balanced, bank-conflict-free banded schedule, uploaded to the device.
//
// Rows are cut to [first non-zero, last non-zero] and sorted by length (longest first); round r holds rows
// 32r .. 32r+31, one per lane, and every lane of the round runs the same number of float4 weight groups G_r
// (rows padded with zero weights), so the inner loop has a warp-uniform trip count.  The kernel reads the
// magnitude tile with 128-bit loads, so a row's read window must start on a 16-byte boundary (`align` tile
// elements) — the slack of the padding is used to slide each window so that the 8 lanes of every quarter
// warp hit 8 different 16-byte bank groups (greedy placement with restarts; residual conflicts only cost
// replays, never correctness).  Weights are stored [round group][lane] so their 128-bit loads are
// conflict-free by construction.
struct MelSchedule {  // what upload_filterbank puts on the device (pure host data, see build_mel_schedule)
    int top_groups = 16, rounds = 0, tile_len = 0, conflict_cost = 0;
    int round_groups[kMaxMelRounds] = {0}, round_wbase[kMaxMelRounds] = {0};
    std::vector<MelEntry> ent;
    std::vector<float> w;
};

static int build_mel_schedule(bool pair, const float *W, int n_mels, int F, MelSchedule *out) {
    struct Row { int m, lo, cnt; };
    const int align = pair ? 2 : 4;  // tile elements per 16 bytes (float2 pairs vs float)
    int tile_len = pair ? kPairTileLen : kSplitTileLen;
    std::vector<Row> rows(n_mels);
    int top = 0;  // one past the highest bin any row touches
    for (int m = 0; m < n_mels; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < F; ++k)
            if (W[(size_t)m * F + k] != 0.f) {
                if (first < 0) first = k;
                last = k;
            }
        rows[m] = {m, first < 0 ? 0 : first, first < 0 ? 0 : last - first + 1};
        top = std::max(top, last + 1);
    }
    // Pair mode with a filterbank that ends below bin 384 (e.g. fmax = 8000 Hz at 22050 Hz): the kernel variant
    // that separates only the first 12 groups of 32 bins is used, and every read window is kept below bin 384.
    out->top_groups = 16;
    if (pair && top <= 384 && !getenv("B200MEL_NO_PRUNE")) {
        bool fits = true;
        for (int m = 0; m < n_mels; ++m) fits = fits && (rows[m].cnt + 8 <= 384);
        if (fits) out->top_groups = 12, tile_len = 384;
    }
    out->tile_len = tile_len;
    auto need = [&](const Row &r) { return (r.cnt + r.lo % align + 3) / 4; };  // groups incl. alignment lead-in
    std::stable_sort(rows.begin(), rows.end(), [&](const Row &x, const Row &y) { return need(x) > need(y); });
    const int rounds = (n_mels + 31) / 32;
    if (rounds > kMaxMelRounds) return fail(B200MEL_EUNSUP, "filterbank: more than 256 mel rows");
    out->rounds = rounds;
    out->conflict_cost = 0;
    std::vector<MelEntry> &ent = out->ent;
    std::vector<float> &w = out->w;
    ent.assign((size_t)rounds * 32, MelEntry{0, -1});
    w.clear();
    uint32_t rng = 12345u;
    auto rnd = [&]() { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
    for (int r = 0; r < rounds; ++r) {
        const int n = std::min(32, n_mels - r * 32);
        const Row *rr = &rows[r * 32];
        int G = 1;
        for (int i = 0; i < n; ++i) G = std::max(G, need(rr[i]));
        if (4 * G > tile_len) return fail(B200MEL_EUNSUP, "filterbank: row longer than the magnitude tile");
        out->round_groups[r] = G;
        out->round_wbase[r] = (int)(w.size() / 4);
        // candidate window starts of every row
        std::vector<std::vector<int>> opts(n);
        for (int i = 0; i < n; ++i) {
            const int lo_min = std::max(0, rr[i].lo + rr[i].cnt - 4 * G), lo_max = std::min(rr[i].lo, tile_len - 4 * G);
            for (int l = (lo_min + align - 1) / align * align; l <= lo_max; l += align) opts[i].push_back(l);
            if (opts[i].empty()) opts[i].push_back(std::max(0, lo_max / align * align));  // cannot happen (G covers it)
        }
        int best_cost = 1 << 30;
        std::vector<int> best_lane(n), best_lo(n);
        for (int trial = 0; trial < 200 && best_cost > 4; ++trial) {
            std::vector<int> order(n);
            for (int i = 0; i < n; ++i) order[i] = i;
            for (int i = n - 1; i > 0; --i) std::swap(order[i], order[rnd() % (i + 1)]);
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return opts[x].size() < opts[y].size(); });
            int used[4][8] = {{0}}, qsize[4] = {0};
            std::vector<int> lane(n), lo(n);
            for (int oi = 0; oi < n; ++oi) {
                const int i = order[oi];
                int bq = -1, bl = 0, bkey = 1 << 30;
                const int q0 = rnd() % 4;
                for (int dq = 0; dq < 4; ++dq) {
                    const int q = (q0 + dq) % 4;
                    if (qsize[q] >= 8) continue;
                    for (int l : opts[i]) {
                        const int key = used[q][(l / align) % 8] * 16 + qsize[q];
                        if (key < bkey) bkey = key, bq = q, bl = l;
                    }
                }
                used[bq][(bl / align) % 8]++;
                lane[i] = bq * 8 + qsize[bq]++;
                lo[i] = bl;
            }
            int cost = 0;
            for (int q = 0; q < 4; ++q) {
                int mx = 0;
                for (int k = 0; k < 8; ++k) mx = std::max(mx, used[q][k]);
                cost += mx;
            }
            if (cost < best_cost) best_cost = cost, best_lane = lane, best_lo = lo;
        }
        out->conflict_cost = std::max(out->conflict_cost, best_cost);
        // weights of the round: [g][lane] float4
        const size_t base = w.size();
        w.resize(base + (size_t)G * 32 * 4, 0.f);
        for (int i = 0; i < n; ++i) {
            const Row &row = rr[i];
            MelEntry &e = ent[(size_t)r * 32 + best_lane[i]];
            e.lo = best_lo[i];
            e.m = row.m;
            for (int k = 0; k < 4 * G; ++k) {
                const int bin = e.lo + k;
                if (bin >= row.lo && bin < row.lo + row.cnt)
                    w[base + ((size_t)(k / 4) * 32 + best_lane[i]) * 4 + k % 4] = W[(size_t)row.m * F + bin];
            }
        }
    }
    if (w.empty()) w.resize(4, 0.f);
    return B200MEL_OK;
}

// Tensor-core STFT kernel: (re)build its operand blob for the plan's current filterbank.  Not an error when the plan
// is outside the kernel's geometry — the plan then simply keeps using the CUDA-core kernels.
static int stc_prepare(b200mel_plan *pl) {
    cudaFree(pl->d_stc);
    pl->d_stc = nullptr;
    pl->stc_mel_bytes = pl->stc_smem = 0;
    if (pl->cfg.n_fft != kStcNfft || pl->cfg.win_length != kStcNfft || pl->cfg.hop_length != kStcHop || pl->W_dense.empty())
        return B200MEL_OK;
    StcTables tb;
    const char *why = nullptr;
    if (!stc_build_tables(pl->W_dense.data(), pl->cfg.n_mels > 0 ? (int)(pl->W_dense.size() / pl->n_freq) : 0, pl->n_freq, &tb, &why))
        return B200MEL_OK;
    const int smem = kStcOffMel + tb.mel_bytes;
    if (smem > kMaxSmem) return B200MEL_OK;
    cudaError_t e;
    if ((e = cudaMalloc(&pl->d_stc, tb.blob.size())) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMemcpy(pl->d_stc, tb.blob.data(), tb.blob.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
        return cuda_fail(e, "cudaMemcpy(tensor-core tables)");
    for (int power = 1; power <= 2; ++power) {
        e = cudaFuncSetAttribute(power == 2 ? stft_tc_kernel<2> : stft_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(stft_tc_kernel)");
    }
    pl->stc_mel_bytes = tb.mel_bytes;
    pl->stc_smem = smem;
    return B200MEL_OK;
}

static int upload_filterbank(b200mel_plan *pl, const float *W_log, int n_mels, int F_log) {
    MelSchedule sch;
    pl->W_dense.assign(W_log, W_log + (size_t)n_mels * F_log);
    cudaFree(pl->d_tc_bhi);
    cudaFree(pl->d_tc_blo);
    pl->d_tc_bhi = pl->d_tc_blo = nullptr;  // rebuilt on demand from W_dense
    pl->tc_k_steps = 0;
    // logical bin k lives at physical bin k * bin_step of the transform the kernel runs
    const int F = pl->phys_n_freq;
    std::vector<float> W_phys;
    const float *W = W_log;
    if (pl->bin_step != 1) {
        W_phys.assign((size_t)n_mels * F, 0.f);
        for (int m = 0; m < n_mels; ++m)
            for (int k = 0; k < F_log; ++k) W_phys[(size_t)m * F + (size_t)k * pl->bin_step] = W_log[(size_t)m * F_log + k];
        W = W_phys.data();
    }
    if (int rc = build_mel_schedule(pl->pair, W, n_mels, F, &sch)) return rc;
    pl->top_groups = sch.top_groups;
    for (int r = 0; r < kMaxMelRounds; ++r) pl->round_groups[r] = sch.round_groups[r], pl->round_wbase[r] = sch.round_wbase[r];
    free_mel_tables(pl);
    cudaError_t e;
    if ((e = cudaMalloc(&pl->d_mel_entries, sch.ent.size() * sizeof(MelEntry))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMalloc(&pl->d_mel_w, sch.w.size() * sizeof(float))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    cudaMemcpy(pl->d_mel_entries, sch.ent.data(), sch.ent.size() * sizeof(MelEntry), cudaMemcpyHostToDevice);
    e = cudaMemcpy(pl->d_mel_w, sch.w.data(), sch.w.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(filterbank)");
    pl->mel_rounds = sch.rounds;
    pl->mel_w_len = (int)sch.w.size();
    if (int rc = layout_smem(pl)) return rc;
    return stc_prepare(pl);
}

typedef void (*kernel_fn)(const KParams);
// Mel kernels: {pair, split} x {magnitude, power} x {all bins, bins < 384 (pair only)}, 16 warps per CTA; the
// spectrum-output operators use the 8-warp cooperative kernel of spec_kernel.cuh.
template <int kTop, bool kPre>
static kernel_fn pick_mel_kernel(bool pair, int power) {
    if (pair) return power == 2 ? logmel_kernel<true, 2, kTop, kPre> : logmel_kernel<true, 1, kTop, kPre>;
    return power == 2 ? logmel_kernel<false, 2, 16, kPre> : logmel_kernel<false, 1, 16, kPre>;
}
// pre: the instantiation with the fused pre-emphasis prologue (io.preemphasis != 0)
static kernel_fn pick_kernel(bool pair, int spec, bool mel, int power, int top_groups, bool pre = false) {
    (void)spec;
    if (!mel) return nullptr;
    if (pre) return top_groups == 12 ? pick_mel_kernel<12, true>(pair, power) : pick_mel_kernel<16, true>(pair, power);
    return top_groups == 12 ? pick_mel_kernel<12, false>(pair, power) : pick_mel_kernel<16, false>(pair, power);
}
template <bool kPair, int kSpec>
static kernel_fn pick_spec_shape(int groups) {
    if constexpr (kPair) {
        switch (groups) {
            case 4: return spec_kernel<true, kSpec, 4>;
            case 3: return spec_kernel<true, kSpec, 3>;
            case 2: return spec_kernel<true, kSpec, 2>;
            default: return spec_kernel<true, kSpec, 1>;
        }
    } else {
        return groups >= 2 ? spec_kernel<false, kSpec, 2> : spec_kernel<false, kSpec, 1>;
    }
}
// compile-time specialised instance for the common geometry (see pick_fast_kernel) of the |X|-only kernel at the
// group count the shared-memory layout gives that geometry.  Measured at C2: 36.8 -> 34.8 us; the same
// specialisation of the two-output kernels (3 groups, 168 registers) measured SLOWER (61 -> 68 us, 54 -> 68 us)
// and is not instantiated.
static kernel_fn pick_spec_fast(int spec, int groups) {
    if (spec == B200MEL_SPEC_MAG && groups == 4) return spec_kernel<true, B200MEL_SPEC_MAG, 4, true>;
    return nullptr;
}
static kernel_fn pick_spec_kernel(bool pair, int spec, int groups) {
    switch (spec) {
        case B200MEL_SPEC_MAG_PHASE: return pair ? pick_spec_shape<true, 1>(groups) : pick_spec_shape<false, 1>(groups);
        case B200MEL_SPEC_RE_IM: return pair ? pick_spec_shape<true, 2>(groups) : pick_spec_shape<false, 2>(groups);
        default: return pair ? pick_spec_shape<true, 3>(groups) : pick_spec_shape<false, 3>(groups);
    }
}

// Fast-path instantiations (logmel_fast.cuh): {bins < 384, all bins} x round signatures of the filterbanks the
// reference's defaults and the BASELINE configs produce.  A plan whose signature is not listed uses the generic kernel.
struct FastEntry { int top; unsigned sig; int power; kernel_fn fn, fn_dct, fn_len; };
#define B200MEL_FAST(top, sig) \
    {top, sig, 1, logmel_fast_kernel<top, sig, 1>, logmel_fast_kernel<top, sig, 1, true>, logmel_fast_kernel<top, sig, 1, false, true>}
static const FastEntry g_fast[] = {
    B200MEL_FAST(12, 0x731u),  // 22050 Hz / 1024 / 80 mels / 0-8000 Hz: settings.py, C2, C3, HiFi-GAN front-end
    B200MEL_FAST(16, 0xa32u),  // 16000 Hz / 1024 / 80 mels / 0-8000 Hz: C5
    B200MEL_FAST(16, 0xa42u),  // 22050 Hz / 1024 / 80 mels / 0-11025 Hz: Audio2Mel defaults (mel_fmax = None)
};
#undef B200MEL_FAST
static unsigned plan_signature(const b200mel_plan *pl) {
    if (pl->mel_rounds < 1 || pl->mel_rounds > 4) return 0;
    unsigned sig = 0;
    for (int r = 0; r < pl->mel_rounds; ++r) {
        if (pl->round_groups[r] < 1 || pl->round_groups[r] > 15) return 0;
        sig = (sig << 4) | (unsigned)pl->round_groups[r];
    }
    return sig;
}
static const bool g_no_fast = getenv("B200MEL_NO_FAST") != nullptr;  // A/B: always run the generic kernel
static kernel_fn pick_fast_kernel(const b200mel_plan *pl, bool dct = false, bool len = false) {
    if ((g_no_fast && !dct) || !pl->pair || pl->pair_frames != 2 || pl->cfg.hop_length != kFastHop || pl->cfg.win_length != pl->phys_n_fft ||
        pl->cfg.n_mels <= 0 || g_table_window || pl->n_warps != kMaxWarps)
        return nullptr;
    const unsigned sig = plan_signature(pl);
    for (const FastEntry &e : g_fast)
        if (e.top == pl->top_groups && e.sig == sig && e.power == pl->cfg.power) return dct ? e.fn_dct : (len ? e.fn_len : e.fn);
    return nullptr;
}

// The tensor-core STFT kernel is used for every eligible plan and launch unless B200MEL_TC=0; B200MEL_TC=1 also uses
// it for launches too small to fill the GPU (fewer 8-frame batches than half the SMs), which the per-warp kernel serves better.
static std::atomic<int> g_stc_mode{[] {
    const char *e = getenv("B200MEL_TC");
    return e ? atoi(e) : B200MEL_TC_DEFAULT;
}()};
static void *g_stc_taps[4] = {nullptr, nullptr, nullptr, nullptr};
static std::atomic<long long> g_stc_launches{0};
static bool stc_enabled(const b200mel_plan *pl, int64_t T, int64_t B) {
    const int mode = g_stc_mode.load();
    if (!pl->d_stc || mode == 0) return false;
    const int64_t batches = ((T + kStcGroup - 1) / kStcGroup) * B;
    if (batches > 0x7fffffff) return false;   // the kernel walks its batches with 32-bit arithmetic
    if (mode == 1) return true;
    return batches >= pl->num_sms / 2;
}

extern "C" {

// debug hooks, not part of include/b200mel.h.  Taps of the tensor-core STFT kernel for the NEXT launches (device
// pointers, null = off): [0] magnitudes (B, 384, T), [1] summed stage-1 accumulators of batch 0 [(n2, t)][32],
// [2] summed stage-2 accumulators of batch 0 [(slot, t)][48], [3] unused.
void b200mel_debug_set_tc_taps(void *mag, void *d1, void *d2, void *a2) {
    g_stc_taps[0] = mag, g_stc_taps[1] = d1, g_stc_taps[2] = d2, g_stc_taps[3] = a2;
}
int64_t b200mel_debug_tc_launch_count(void) { return g_stc_launches.load(); }
// 0: never, 1: every eligible launch, 2: eligible launches that fill the GPU; returns the previous mode
int b200mel_debug_set_tc_mode(int mode) { return g_stc_mode.exchange(mode); }
// host only, no GPU: the operand blob stft_tc_kernel would use for a dense (n_mels, 513) filterbank.  Returns the
// number of bytes written (<= cap), -1 if the filterbank is not eligible.  info[0] = mel bytes, [1] = row groups,
// [2] = longest group, [3] = cost of the busiest warp.
int64_t b200mel_debug_tc_tables(const float *W, int32_t n_mels, int32_t n_freq, unsigned char *out, int64_t cap, int32_t *info) {
    StcTables tb;
    const char *why = nullptr;
    if (!W || !stc_build_tables(W, n_mels, n_freq, &tb, &why)) return -1;
    if (info) info[0] = tb.mel_bytes, info[1] = tb.n_groups, info[2] = tb.max_len, info[3] = tb.warp_cost_max;
    if (out && cap >= (int64_t)tb.blob.size()) memcpy(out, tb.blob.data(), tb.blob.size());
    return (int64_t)tb.blob.size();
}

// debug hook (tools/tc_bench.py probe): times chains of tcgen05.mma for n_cfg operand layouts, 8 uint32 per config
// {a_off, a_lbo, a_sbo, b_off, b_lbo, b_sbo, n, chain}; out = 2 int64 per config {issue cycles, issue + completion cycles}
int b200mel_debug_mma_probe(const uint32_t *cfgs, int32_t n_cfg, long long *out_host) {
    ProbeCfg *d_cfg = nullptr;
    long long *d_out = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&d_cfg, n_cfg * sizeof(ProbeCfg))) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMalloc(&d_out, n_cfg * 16)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    cudaMemcpy(d_cfg, cfgs, n_cfg * sizeof(ProbeCfg), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(stc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    stc_probe_kernel<<<1, 128, 200 * 1024>>>(d_cfg, n_cfg, d_out);
    e = cudaMemcpy(out_host, d_out, n_cfg * 16, cudaMemcpyDeviceToHost);
    cudaFree(d_cfg);
    cudaFree(d_out);
    return e == cudaSuccess ? B200MEL_OK : cuda_fail(e, "mma probe");
}

int b200mel_version(void) { return B200MEL_VERSION; }
const char *b200mel_last_error(void) { return g_err.c_str(); }
int64_t b200mel_launch_count(void) { return g_launches.load(); }
// debug hook, not part of include/b200mel.h: device pointer to >= 16 int64 accumulators (phase-timing builds)
void b200mel_debug_set_buffer(long long *dev_ptr) { g_dbg = dev_ptr; }
// debug hook, not part of include/b200mel.h (host only, no GPU): builds the banded mel schedule for a dense filterbank
// and REPLAYS it the way the kernel's mel loop walks it — round by round, lane by lane, weight group by weight group —
// into dense_out[n_mels][n_freq].  info[0] = bin groups the kernel separates (12 / 16), [1] = rounds, [2] = tile
// length the read windows stay below, [3] = worst quarter-warp bank-conflict cost (4 = conflict-free),
// [4] = 1 if any read window is misaligned or leaves the tile, [5 + r] = weight groups of round r.
int b200mel_debug_mel_schedule(const float *W, int32_t n_mels, int32_t n_freq, int32_t pair, float *dense_out, int32_t *info) {
    if (!W || !dense_out || !info || n_mels <= 0 || n_freq <= 0) return fail(B200MEL_EINVAL, "debug_mel_schedule: bad argument");
    MelSchedule sch;
    if (int rc = build_mel_schedule(pair != 0, W, n_mels, n_freq, &sch)) return rc;
    const int align = pair ? 2 : 4;
    info[0] = sch.top_groups, info[1] = sch.rounds, info[2] = sch.tile_len, info[3] = sch.conflict_cost, info[4] = 0;
    for (size_t i = 0; i < (size_t)n_mels * n_freq; ++i) dense_out[i] = 0.f;
    for (int r = 0; r < sch.rounds; ++r) {
        info[5 + r] = sch.round_groups[r];
        for (int lane = 0; lane < 32; ++lane) {
            const MelEntry &e = sch.ent[(size_t)r * 32 + lane];
            if (e.m < 0) continue;
            if (e.lo % align || e.lo < 0 || e.lo + 4 * sch.round_groups[r] > sch.tile_len) info[4] = 1;
            for (int g = 0; g < sch.round_groups[r]; ++g)
                for (int k = 0; k < 4; ++k) {
                    const float w = sch.w[((size_t)sch.round_wbase[r] + (size_t)g * 32 + lane) * 4 + k];  // kernel: wbase + round_wbase + g * 32 (float4 units)
                    const int bin = e.lo + 4 * g + k;
                    if (w != 0.f) {
                        if (bin >= n_freq) info[4] = 1;
                        else dense_out[(size_t)e.m * n_freq + bin] += w;
                    }
                }
        }
    }
    return B200MEL_OK;
}

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
    if (cfg->n_fft < 32 || cfg->n_fft > 2048 || (cfg->n_fft & (cfg->n_fft - 1)))
        return fail(B200MEL_EUNSUP, "plan_create: n_fft must be a power of two in [32, 2048] in this build");
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
    pl->num_sms = prop.multiProcessorCount;
    pl->phys_n_fft = cfg->n_fft <= 1024 ? 1024 : 2048;
    pl->bin_step = pl->phys_n_fft / cfg->n_fft;
    pl->phys_n_freq = pl->phys_n_fft / 2 + 1;
    pl->pair = pl->phys_n_fft == 1024;
    pl->pair_frames = (pl->pair && cfg->hop_length <= pl->phys_n_fft) ? 2 : 1;
    pl->n_freq = cfg->n_fft / 2 + 1;
    pl->pad = cfg->pad_mode == B200MEL_PAD_CENTER ? cfg->n_fft / 2 : (cfg->n_fft - cfg->hop_length) / 2;
    if (pl->pad < 0) pl->pad = 0;

    const int N = pl->phys_n_fft;
    std::vector<float> win(N, 0.f);  // logical window (centre-padded to n_fft), zero-extended to the physical transform
    build_window(cfg->win_length, cfg->n_fft, win.data());
    for (auto &w : win) w *= 0.5f;  // exact; the separation pass omits its 1/2
    std::vector<float2> tw(32 * 32), twp(32);
    const double two_pi = 6.283185307179586476925286766559;
    for (int j = 0; j < 32; ++j)  // [j / 2][lane][j & 1]: one 128-bit load fetches the twiddles of slots j, j + 1
        for (int l = 0; l < 32; ++l) {
            double ang = two_pi * (double)(j * l) / 1024.0;
            tw[((j >> 1) * 32 + l) * 2 + (j & 1)] = make_float2((float)cos(ang), (float)-sin(ang));
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
            build_filterbank(cfg->sample_rate, cfg->n_fft, cfg->n_mels, cfg->fmin, fmax, cfg->mel_scale == B200MEL_MEL_HTK,
                             cfg->mel_norm == B200MEL_NORM_SLANEY, W.data());
            rc = upload_filterbank(pl, W.data(), cfg->n_mels, pl->n_freq);
            if (rc) break;
        } else if ((rc = layout_smem(pl)) != B200MEL_OK)
            break;
        for (int power = 1; power <= 2 && e == cudaSuccess; ++power)
            for (int top = 12; top <= 16 && e == cudaSuccess; top += 4)
                for (int pre = 0; pre <= 1 && e == cudaSuccess; ++pre)
                    e = cudaFuncSetAttribute(pick_kernel(pl->pair, 0, true, power, top, pre != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        for (int spec = 1; spec <= 3 && e == cudaSuccess; ++spec)
        {
            e = cudaFuncSetAttribute(pick_spec_kernel(pl->pair, spec, pl->sp[spec].groups),
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
            if (kernel_fn ff = pl->pair ? pick_spec_fast(spec, pl->sp[spec].groups) : nullptr)
                if (e == cudaSuccess) e = cudaFuncSetAttribute(ff, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        }
        for (const FastEntry &fe : g_fast) {
            if (e == cudaSuccess) e = cudaFuncSetAttribute(fe.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(fe.fn_dct, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(fe.fn_len, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        }
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
    cudaFree(pl->d_tc_bhi);
    cudaFree(pl->d_tc_blo);
    cudaFree(pl->d_stc);
    for (auto &hs : pl->host_stage) {
        cudaFree(hs.d_in);
        cudaFree(hs.d_out);
    }
    delete pl;
    return B200MEL_OK;
}

int b200mel_forward(const b200mel_plan *pl, const float *wav, int64_t B, int64_t L, int64_t row_stride,
                    const int32_t *lengths, const b200mel_epilogue *epi, float *out_mel, int32_t spec_kind,
                    float *out_a, float *out_b, void *stream) {
    b200mel_io io;
    memset(&io, 0, sizeof(io));
    io.struct_size = (int32_t)sizeof(io);
    io.spec_kind = spec_kind;
    io.wav = wav, io.B = B, io.L = L, io.row_stride = row_stride, io.lengths = lengths;
    io.out_mel = out_mel, io.out_a = out_a, io.out_b = out_b;
    return b200mel_forward_io(pl, &io, epi, stream);
}

int b200mel_forward_io(const b200mel_plan *pl, const b200mel_io *io, const b200mel_epilogue *epi, void *stream) {
    if (!pl) return fail(B200MEL_EINVAL, "forward: null plan");
    if (!io || io->struct_size != (int32_t)sizeof(b200mel_io)) return fail(B200MEL_EINVAL, "forward: io struct_size mismatch");
    const float *wav = io->wav;
    const int64_t B = io->B, L = io->L, row_stride = io->row_stride;
    const int32_t *lengths = io->lengths;
    float *out_mel = io->out_mel, *out_a = io->out_a, *out_b = io->out_b;
    const int32_t spec_kind = io->spec_kind;
    if (io->reserve_sms < 0) return fail(B200MEL_EINVAL, "forward: negative reserve_sms");
    if (io->preemphasis != 0.f && (!io->out_mel || io->spec_kind))
        return fail(B200MEL_EINVAL, "forward: fused pre-emphasis is a prologue of the mel launch (no spectrum outputs)");
    if (io->preemphasis != 0.f && io->L < 2)
        return fail(B200MEL_EINVAL, "forward: pre-emphasis needs L >= 2 (reflect pad, models/sound.py:80)");
    if (io->out_frame_mask && pl->cfg.pad_mode != B200MEL_PAD_CENTER)
        return fail(B200MEL_EINVAL, "forward: out_frame_mask needs centre framing (SpectrogramMasker geometry)");
    if (io->out_frame_mask && !out_mel) return fail(B200MEL_EINVAL, "forward: out_frame_mask is written by the mel launch");
    if (B < 0 || L < 0) return fail(B200MEL_EINVAL, "forward: negative shape");
    if (B > 0x7fffffff) return fail(B200MEL_EINVAL, "forward: more than 2^31 - 1 clips");
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
    if (io->reserved0 != 0) return fail(B200MEL_EINVAL, "forward: io.reserved0 must be 0");
    if (io->out_mfcc) {
        if (!io->dct_mat || io->n_mfcc < 1) return fail(B200MEL_EINVAL, "forward: out_mfcc needs dct_mat and n_mfcc >= 1");
        if (pl->cfg.n_mels == 0) return fail(B200MEL_EINVAL, "forward: plan has no filterbank (n_mels = 0)");
        if (!epi) return fail(B200MEL_EINVAL, "forward: mel output needs an epilogue");
    }
    if (!out_mel && !spec_kind && !io->out_mfcc) return fail(B200MEL_EINVAL, "forward: no output requested");
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
    if (out_mel && (int64_t)pl->cfg.n_mels * T > 0x7fffffff)
        return fail(B200MEL_EINVAL, "forward: n_mels * frames per clip exceeds 2^31 - 1 (the kernels index a clip's mel block with 32 bits)");

    KParams p;
    memset(&p, 0, sizeof(p));
    p.wav = wav;
    {
        const uintptr_t lo = reinterpret_cast<uintptr_t>(wav), hi = reinterpret_cast<uintptr_t>(wav + (B - 1) * row_stride + L);
        p.wav_lo16 = (unsigned long long)((lo + 15) & ~(uintptr_t)15);
        p.wav_hi16 = (unsigned long long)(hi & ~(uintptr_t)15);
    }
    p.row_stride = row_stride;
    p.B = B;
    p.L = (int)L;
    p.lengths = lengths;
    p.T = (int)T;
    p.hop = pl->cfg.hop_length;
    p.pad = pl->pad;
    p.n_fft = pl->phys_n_fft;
    p.n_fft_log = pl->cfg.n_fft;
    p.bin_step = pl->bin_step;
    p.n_freq_out = pl->n_freq;
    p.pair_frames = pl->pair_frames;
    p.hann_full = (pl->pair && pl->cfg.win_length == pl->phys_n_fft && !g_table_window) ? 1 : 0;
    p.window = pl->d_window;
    p.tw = pl->d_tw;
    p.tw_post = pl->d_tw_post;
    p.mel_entries = pl->d_mel_entries;
    p.mel_w = pl->d_mel_w;
    p.n_mels = pl->cfg.n_mels;
    p.n_freq = pl->phys_n_freq;
    p.mel_rounds = pl->mel_rounds;
    p.mel_w_len = pl->mel_w_len;
    for (int r = 0; r < kMaxMelRounds; ++r) p.round_groups[r] = pl->round_groups[r], p.round_wbase[r] = pl->round_wbase[r];
    p.off_window = pl->off_window;
    p.off_entries = pl->off_entries;
    p.off_melw = pl->off_melw;
    p.off_bar = pl->off_bar;
    p.off_regions = pl->off_regions;
    p.region_bytes = pl->region_bytes;
    p.stage_bytes = pl->stage_bytes;
    p.out_mel = out_mel;
    p.out_fmask = io->out_frame_mask;
    p.preemph = io->preemphasis;
    p.win_half = pl->cfg.win_length / 2;
    p.out_a = out_a;
    p.out_b = out_b;
    p.mag_eps = pl->cfg.mag_eps;
    p.dbg = g_dbg;
    p.lo = -INFINITY;
    p.hi = INFINITY;
    p.norm_scale = 1.f;
    p.norm_bias = 0.f;
    p.ep_floor = -INFINITY;
    if (epi) {
        p.use_log = epi->log_kind != B200MEL_LOG_NONE;
        if (epi->log_kind == B200MEL_LOG_LN_OFFSET) p.ep_offset = epi->log_arg;
        else if (p.use_log) p.ep_floor = epi->log_arg;
        p.log_scale = epi->log_kind == B200MEL_LOG_LOG10_FLOOR ? 0.301029995663981195f : 0.693147180559945309f;
        if (epi->has_clamp_lo) p.lo = epi->clamp_lo;
        if (epi->has_clamp_hi) p.hi = epi->clamp_hi;
        if (epi->norm_mel) {  // (y - lo) / (hi - lo) * 2 - 1
            p.norm_scale = 2.0f / (p.hi - p.lo);
            p.norm_bias = -p.lo * p.norm_scale - 1.0f;
        }
    }
    const long long tpc = (T + pl->pair_frames - 1) / pl->pair_frames;
    p.tasks_per_clip = (int)tpc;
    p.n_tasks = tpc * B;
    long long n_cta = (p.n_tasks + pl->n_warps - 1) / pl->n_warps;
    int usable_sms = pl->num_sms;
    if (io->reserve_sms > 0) usable_sms = std::max(1, pl->num_sms - io->reserve_sms);
    if (n_cta > usable_sms) n_cta = usable_sms;  // persistent: one CTA per SM, warps stride over the tasks
    const long long stride = n_cta * pl->n_warps;
    p.stride_b = (int)(stride / tpc);
    p.stride_q = (int)(stride % tpc);

    cudaStream_t st = (cudaStream_t)stream;
    // Launch with programmatic stream serialization (PDL): the kernel's table prologue may overlap the tail of the
    // previous kernel in the stream; it executes griddepcontrol.wait before touching wav / outputs.
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.dynamicSmemBytes = pl->smem_bytes;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaSuccess;
    // mel and spectrum outputs come from separately specialised kernels (no reference module needs both at once)
    if (io->out_mfcc) {
        // MFCC as a fused epilogue of the compile-time specialised kernel; everything else is the caller's two launches
        kernel_fn fn = (!lengths && p.use_log && !p.out_fmask && p.preemph == 0.f && !spec_kind && io->n_mfcc <= kDctPitch)
                           ? pick_fast_kernel(pl, true) : nullptr;
        const int dct_rows = (pl->cfg.n_mels + 1) & ~1;
        p.dct = io->dct_mat, p.out_mfcc = io->out_mfcc, p.n_mfcc = io->n_mfcc;
        p.off_dct = (pl->smem_bytes + 127) & ~127;
        p.off_col = p.off_dct + dct_rows * kDctPitch * 4;
        p.col_bytes = (dct_rows * 8 + 127) & ~127;
        const int smem = p.off_col + pl->n_warps * p.col_bytes;
        if (!fn || smem > kMaxSmem)
            return fail(B200MEL_EUNSUP, "forward: the fused MFCC epilogue serves n_fft = win_length = 1024, hop 256, no lengths / "
                                        "frame mask / pre-emphasis, n_mfcc <= 64 — run b200mel_mel_to_mfcc on the mel output instead");
        cfg.gridDim = dim3((unsigned)n_cta);
        cfg.blockDim = dim3(pl->n_warps * 32);
        cfg.dynamicSmemBytes = smem;
        le = cudaLaunchKernelEx(&cfg, fn, p);
        g_launches.fetch_add(1);
    } else if (out_mel && stc_enabled(pl, T, B) && !lengths && !p.out_fmask && p.preemph == 0.f && !spec_kind) {
        // both DFT stages on the tensor cores (stft_tc.cuh)
        StcParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.k = p;
        sp.tables = pl->d_stc;
        sp.mel_bytes = pl->stc_mel_bytes;
        sp.groups_per_clip = (int)((T + kStcGroup - 1) / kStcGroup);
        sp.n_batches = (long long)sp.groups_per_clip * B;
        sp.power = pl->cfg.power;
        sp.dbg_mag = g_stc_taps[0] ? (float *)g_stc_taps[0] : nullptr;
        sp.dbg_d1 = (float *)g_stc_taps[1];
        sp.dbg_d2 = (float *)g_stc_taps[2];
        sp.dbg_a2 = (unsigned char *)g_stc_taps[3];
        cfg.gridDim = dim3((unsigned)std::min<long long>(sp.n_batches, usable_sms));
        cfg.blockDim = dim3(kStcThreads);
        cfg.dynamicSmemBytes = pl->stc_smem;
        le = pl->cfg.power == 2 ? cudaLaunchKernelEx(&cfg, stft_tc_kernel<2>, sp) : cudaLaunchKernelEx(&cfg, stft_tc_kernel<1>, sp);
        g_launches.fetch_add(1);
        g_stc_launches.fetch_add(1);
    } else if (out_mel) {
        cfg.gridDim = dim3((unsigned)n_cta);
        cfg.blockDim = dim3(pl->n_warps * 32);
        kernel_fn fn = (p.use_log && p.preemph == 0.f) ? pick_fast_kernel(pl, false, lengths || p.out_fmask) : nullptr;
        if (!fn) fn = pick_kernel(pl->pair, 0, true, pl->cfg.power, pl->top_groups, p.preemph != 0.f);
        le = cudaLaunchKernelEx(&cfg, fn, p);
        g_launches.fetch_add(1);
    }
    if (spec_kind && le == cudaSuccess) {
        // cooperative kernel with its own shared-memory carve-up and launch shape (per output kind)
        const b200mel_plan::SpecLayout &L = pl->sp[spec_kind];
        const int cta_tasks = L.warps;  // one task per warp and round
        long long s_cta = (p.n_tasks + cta_tasks - 1) / cta_tasks;
        if (s_cta > pl->num_sms) s_cta = pl->num_sms;
        const long long step = s_cta * cta_tasks;  // between a warp's consecutive tasks
        p.stride_b = (int)(step / tpc);
        p.stride_q = (int)(step % tpc);
        p.off_bar = L.off_bar;
        p.off_entries = L.off_slots;
        p.off_regions = L.off_regions;
        p.off_melw = L.off_tiles;
        p.region_bytes = L.region;
        cfg.gridDim = dim3((unsigned)s_cta);
        cfg.blockDim = dim3(L.warps * 32);
        cfg.dynamicSmemBytes = L.smem;
        kernel_fn sfn = nullptr;
        if (!g_no_fast && !lengths && pl->pair && pl->pair_frames == 2 && pl->cfg.hop_length == kFastHop && pl->cfg.n_fft == pl->phys_n_fft &&
            pl->cfg.win_length == pl->phys_n_fft && !g_table_window)
            sfn = pick_spec_fast(spec_kind, L.groups);
        if (!sfn) sfn = pick_spec_kernel(pl->pair, spec_kind, L.groups);
        le = cudaLaunchKernelEx(&cfg, sfn, p);
        g_launches.fetch_add(1);
    }
    if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx");
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
    const size_t in_bytes = ((size_t)(B - 1) * row_stride + L) * sizeof(float);
    const size_t out_bytes = (size_t)B * pl->cfg.n_mels * T * sizeof(float);
    cudaError_t e;
    cudaStream_t st = (cudaStream_t)stream;
    b200mel_plan::HostStage *hs = nullptr;
    {
        std::lock_guard<std::mutex> lock(pl->host_mu);
        for (auto &c : pl->host_stage)
            if (c.used && c.stream == stream) hs = &c;
        if (!hs)
            for (auto &c : pl->host_stage)
                if (!c.used) {
                    hs = &c;
                    hs->used = true;
                    hs->stream = stream;
                    break;
                }
        if (!hs) return fail(B200MEL_EUNSUP, "forward_host: more than 8 distinct streams on one plan");
        // growing a staging buffer frees the old one: wait for the work this stream still has on it (rare: first call /
        // larger batch), never inside the steady state
        if (in_bytes > hs->in_bytes) {
            cudaStreamSynchronize(st);
            cudaFree(hs->d_in);
            hs->d_in = nullptr, hs->in_bytes = 0;
            if ((e = cudaMalloc(&hs->d_in, in_bytes)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(stage_in)");
            hs->in_bytes = in_bytes;
        }
        if (out_bytes > hs->out_bytes) {
            cudaStreamSynchronize(st);
            cudaFree(hs->d_out);
            hs->d_out = nullptr, hs->out_bytes = 0;
            if ((e = cudaMalloc(&hs->d_out, out_bytes)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(stage_out)");
            hs->out_bytes = out_bytes;
        }
    }
    if ((e = cudaMemcpyAsync(hs->d_in, wav_host, in_bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess)
        return cuda_fail(e, "cudaMemcpyAsync(H2D)");
    int rc = b200mel_forward(pl, hs->d_in, B, L, row_stride, nullptr, epi, hs->d_out, B200MEL_SPEC_NONE, nullptr, nullptr,
                             stream);
    if (rc) return rc;
    if ((e = cudaMemcpyAsync(out_mel_host, hs->d_out, out_bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
        return cuda_fail(e, "cudaMemcpyAsync(D2H)");
    return B200MEL_OK;
}

static int grid_for(long long work_items, int block, int sms) {
    long long g = (work_items + block - 1) / block;
    const long long cap = (long long)sms * 8;  // a few waves of resident CTAs, grid-stride beyond that
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}
static int current_sms(int *sms) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B200MEL_ENODEV, "no CUDA device (there is no CPU fallback)");
    }
    return B200MEL_OK;
}

int b200mel_preemphasis(const float *x, int64_t B, int64_t L, int64_t x_row_stride, float coef, float *y,
                        int64_t y_row_stride, void *stream) {
    if (B < 0 || L < 0) return fail(B200MEL_EINVAL, "preemphasis: negative shape");
    if (B == 0 || L == 0) return B200MEL_OK;
    if (!x || !y) return fail(B200MEL_EINVAL, "preemphasis: null pointer");
    if (L < 2) return fail(B200MEL_EINVAL, "preemphasis: reflect padding needs L >= 2 (models/sound.py:80)");
    if (x_row_stride < L || y_row_stride < L || L > 0x7ffffff0) return fail(B200MEL_EINVAL, "preemphasis: bad stride / length");
    int sms = 0;
    if (int rc = current_sms(&sms)) return rc;
    const dim3 grid((unsigned)((L + 1023) / 1024), (unsigned)(B < 65535 ? B : 65535));
    (void)sms;
    preemph_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, B, (int)L, x_row_stride, y_row_stride, coef);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200MEL_OK : cuda_fail(e, "preemphasis launch");
}

int b200mel_volume_norm(const float *x, int64_t n, float target_db, float *y, double *scratch, void *stream) {
    if (n < 0) return fail(B200MEL_EINVAL, "volume_norm: negative size");
    if (n == 0) return B200MEL_OK;
    if (!x || !y || !scratch) return fail(B200MEL_EINVAL, "volume_norm: null pointer");
    int sms = 0;
    if (int rc = current_sms(&sms)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(scratch, 0, 2 * sizeof(double), st);
    if (e != cudaSuccess) return cuda_fail(e, "volume_norm memset");
    const int grid = grid_for(n, 256, sms);
    moments_kernel<<<grid, 256, 0, st>>>(x, n, scratch);
    // x / (std / 10^(dB/10)): the reference applies the 10^(dB/10) power ratio to an amplitude, kept as is
    scale_by_std_kernel<<<grid, 256, 0, st>>>(x, y, n, scratch, powf(10.f, target_db / 10.f));
    g_launches.fetch_add(2);
    e = cudaGetLastError();
    return e == cudaSuccess ? B200MEL_OK : cuda_fail(e, "volume_norm launch");
}

int b200mel_stft_loss_terms(const float *pred_mag, const float *target_mag, int64_t B, int64_t n_per_clip, float eps,
                            double *scratch, float *out2, void *stream) {
    if (B < 0 || n_per_clip < 0) return fail(B200MEL_EINVAL, "stft_loss_terms: negative shape");
    if (B == 0 || n_per_clip == 0) return fail(B200MEL_EINVAL, "stft_loss_terms: empty input (the reference divides by zero)");
    if (!pred_mag || !target_mag || !scratch || !out2) return fail(B200MEL_EINVAL, "stft_loss_terms: null pointer");
    int sms = 0;
    if (int rc = current_sms(&sms)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t)B * 3 * sizeof(double), st);
    if (e != cudaSuccess) return cuda_fail(e, "stft_loss_terms memset");
    // a few waves of CTAs over (chunks, clips)
    long long by = B < 65535 ? B : 65535;
    long long bx = (n_per_clip + 256 * 8 - 1) / (256 * 8);
    const long long cap = std::max<long long>(1, (long long)sms * 8 / by);
    if (bx > cap) bx = cap;
    stft_loss_partial_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, st>>>(pred_mag, target_mag, B, n_per_clip, eps, scratch);
    stft_loss_final_kernel<<<1, 32, 0, st>>>(scratch, B, n_per_clip, out2);
    g_launches.fetch_add(2);
    e = cudaGetLastError();
    return e == cudaSuccess ? B200MEL_OK : cuda_fail(e, "stft_loss_terms launch");
}

// fp32 -> bf16 (round to nearest even) on the host, as __float2bfloat16_rn does on the device
static uint16_t bf16_rn(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static float bf16_to_float(uint16_t h) {
    const uint32_t u = (uint32_t)h << 16;
    float x;
    memcpy(&x, &u, 4);
    return x;
}
// Build (once per plan / filterbank) the bf16 hi / lo limbs of the filterbank in the layout mel_tc_kernel's B
// descriptors address: [k_step][k / 8][n / 8][n % 8][k % 8].
static int tc_prepare(b200mel_plan *pl) {
    std::lock_guard<std::mutex> lock(pl->tc_mu);
    if (pl->tc_k_steps) return B200MEL_OK;
    const int n_mels = pl->cfg.n_mels, F = pl->n_freq;
    if (n_mels <= 0 || pl->W_dense.empty()) return fail(B200MEL_EINVAL, "logmel_from_magnitude: plan has no filterbank");
    int top = 0;
    for (int m = 0; m < n_mels; ++m)
        for (int k = 0; k < F; ++k)
            if (pl->W_dense[(size_t)m * F + k] != 0.f) top = std::max(top, k + 1);
    const int n_pad = (n_mels + 15) & ~15;
    int k_steps = (std::max(top, 1) + 15) / 16;
    k_steps += k_steps & 1;  // a pipeline stage is two K steps
    const int b_bytes = k_steps * 2 * (n_pad / 8) * 128;
    const int smem = 2 * b_bytes + 2 * kTcABuf + 2 * 8 + 16;
    if (n_pad > kTcTmemCols || smem > kMaxSmem)
        return fail(B200MEL_EUNSUP, "logmel_from_magnitude: filterbank too large for the shared-memory-resident tensor-core "
                                    "kernel of this build (n_mels <= 128 and bins x mels x 4 bytes <= ~190 KB)");
    std::vector<uint16_t> hi((size_t)b_bytes / 2, 0), lo((size_t)b_bytes / 2, 0);
    for (int ks = 0; ks < k_steps; ++ks)
        for (int kk = 0; kk < 16; ++kk)
            for (int n = 0; n < n_pad; ++n) {
                const int k = ks * 16 + kk;
                const float w = (n < n_mels && k < F) ? pl->W_dense[(size_t)n * F + k] : 0.f;
                const uint16_t h = bf16_rn(w);
                const uint16_t l = bf16_rn(w - bf16_to_float(h));
                const size_t off = ((size_t)ks * 2 * (n_pad / 8) * 128 + (size_t)(kk / 8) * (n_pad / 8) * 128 + (size_t)(n / 8) * 128 +
                                    (size_t)(n % 8) * 16 + (size_t)(kk % 8) * 2) / 2;
                hi[off] = h, lo[off] = l;
            }
    cudaError_t e;
    if ((e = cudaMalloc(&pl->d_tc_bhi, b_bytes)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMalloc(&pl->d_tc_blo, b_bytes)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    cudaMemcpy(pl->d_tc_bhi, hi.data(), b_bytes, cudaMemcpyHostToDevice);
    if ((e = cudaMemcpy(pl->d_tc_blo, lo.data(), b_bytes, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
    if ((e = cudaFuncSetAttribute(mel_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess)
        return cuda_fail(e, "cudaFuncSetAttribute(mel_tc_kernel)");
    pl->tc_n_pad = n_pad, pl->tc_b_bytes = b_bytes, pl->tc_smem = smem;
    pl->tc_k_steps = k_steps;
    return B200MEL_OK;
}

int b200mel_logmel_from_magnitude(b200mel_plan *pl, const float *mag, int64_t B, int64_t T, const b200mel_epilogue *epi,
                                  float *out_mel, void *stream) {
    if (!pl) return fail(B200MEL_EINVAL, "logmel_from_magnitude: null plan");
    if (B < 0 || T < 0) return fail(B200MEL_EINVAL, "logmel_from_magnitude: negative shape");
    if (B == 0 || T == 0) return B200MEL_OK;
    if (!mag || !out_mel || !epi) return fail(B200MEL_EINVAL, "logmel_from_magnitude: null pointer");
    if (epi->struct_size != (int32_t)sizeof(b200mel_epilogue)) return fail(B200MEL_EINVAL, "logmel_from_magnitude: epilogue struct_size mismatch");
    if (epi->log_kind < 0 || epi->log_kind > 3) return fail(B200MEL_EINVAL, "logmel_from_magnitude: bad log_kind");
    if (epi->norm_mel && !(epi->has_clamp_lo && epi->has_clamp_hi && epi->clamp_hi > epi->clamp_lo))
        return fail(B200MEL_EINVAL, "logmel_from_magnitude: norm_mel needs clamp_lo < clamp_hi");
    if (T > 0x7fffffff || B * T / kTcTileM > 0x7fffffff) return fail(B200MEL_EINVAL, "logmel_from_magnitude: too many frames");
    if (int rc = tc_prepare(pl)) return rc;
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.mag = mag, p.out = out_mel;
    p.n_cols = B * T;
    p.T = (int)T, p.F = pl->n_freq, p.n_mels = pl->cfg.n_mels;
    p.n_pad = pl->tc_n_pad, p.k_steps = pl->tc_k_steps;
    p.b_hi = pl->d_tc_bhi, p.b_lo = pl->d_tc_blo, p.b_bytes = pl->tc_b_bytes;
    p.lo = -INFINITY, p.hi = INFINITY, p.norm_scale = 1.f, p.norm_bias = 0.f, p.ep_floor = -INFINITY;
    p.use_log = epi->log_kind != B200MEL_LOG_NONE;
    if (epi->log_kind == B200MEL_LOG_LN_OFFSET) p.ep_offset = epi->log_arg;
    else if (p.use_log) p.ep_floor = epi->log_arg;
    p.log_scale = epi->log_kind == B200MEL_LOG_LOG10_FLOOR ? 0.301029995663981195f : 0.693147180559945309f;
    if (epi->has_clamp_lo) p.lo = epi->clamp_lo;
    if (epi->has_clamp_hi) p.hi = epi->clamp_hi;
    if (epi->norm_mel) {
        p.norm_scale = 2.0f / (p.hi - p.lo);
        p.norm_bias = -p.lo * p.norm_scale - 1.0f;
    }
    long long tiles = (p.n_cols + kTcTileM - 1) / kTcTileM;
    const int grid = (int)std::min<long long>(tiles, pl->num_sms);
    mel_tc_kernel<<<grid, kTcThreads, pl->tc_smem, (cudaStream_t)stream>>>(p);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200MEL_OK : cuda_fail(e, "logmel_from_magnitude launch");
}

// Internal streams / events of b200mel_gather_copy, one set per device, created on first use.
struct GatherCtx {
    bool ready = false;
    cudaStream_t streams[15];
    cudaEvent_t fork, join[15];
};
static GatherCtx g_gather[64];
static std::mutex g_gather_mu;

static int gather_impl(float *local_buf, const float *const *peer_bufs, int32_t *const *peer_sync, int32_t world,
                       int32_t rank, const int64_t *block_offsets, void *stream, bool use_copy_engines, int n_ctas,
                       bool use_tma = false);

int b200mel_gather_pull(float *local_buf, const float *const *peer_bufs, int32_t *const *peer_sync, int32_t world,
                        int32_t rank, const int64_t *block_offsets, int32_t n_ctas, void *stream) {
    if (n_ctas < 0) return fail(B200MEL_EINVAL, "gather_pull: negative n_ctas");
    return gather_impl(local_buf, peer_bufs, peer_sync, world, rank, block_offsets, stream, false, n_ctas);
}
int b200mel_gather_tma(float *local_buf, const float *const *peer_bufs, int32_t *const *peer_sync, int32_t world,
                       int32_t rank, const int64_t *block_offsets, int32_t n_ctas, void *stream) {
    if (n_ctas < 0) return fail(B200MEL_EINVAL, "gather_tma: negative n_ctas");
    return gather_impl(local_buf, peer_bufs, peer_sync, world, rank, block_offsets, stream, false, n_ctas, true);
}
int b200mel_gather_copy(float *local_buf, const float *const *peer_bufs, int32_t *const *peer_sync, int32_t world,
                        int32_t rank, const int64_t *block_offsets, void *stream) {
    if (!peer_sync) return fail(B200MEL_EINVAL, "gather_copy: peer_sync is required (the barrier kernel orders the copies)");
    return gather_impl(local_buf, peer_bufs, peer_sync, world, rank, block_offsets, stream, true, 1);
}

static int gather_impl(float *local_buf, const float *const *peer_bufs, int32_t *const *peer_sync, int32_t world,
                       int32_t rank, const int64_t *block_offsets, void *stream, bool use_copy_engines, int n_ctas,
                       bool use_tma) {
    if (!local_buf || !peer_bufs || !block_offsets) return fail(B200MEL_EINVAL, "gather_pull: null pointer");
    if (world < 1 || world > 16 || rank < 0 || rank >= world) return fail(B200MEL_EINVAL, "gather_pull: need 1 <= world <= 16, 0 <= rank < world");
    PullArgs a;
    memset(&a, 0, sizeof(a));
    a.world = world, a.rank = rank;
    for (int r = 0; r <= world; ++r) {
        a.off[r] = block_offsets[r];
        if (r && a.off[r] < a.off[r - 1]) return fail(B200MEL_EINVAL, "gather_pull: block offsets must be non-decreasing");
    }
    for (int r = 0; r < world; ++r) {
        a.peer[r] = peer_bufs[r];
        if (r != rank && !a.peer[r] && a.off[r + 1] > a.off[r]) return fail(B200MEL_EINVAL, "gather_pull: null peer buffer");
        a.peer_sync[r] = peer_sync ? reinterpret_cast<int *>(peer_sync[r]) : nullptr;
        if (peer_sync && !a.peer_sync[r]) return fail(B200MEL_EINVAL, "gather_pull: null sync pointer");
    }
    if (world == 1) return B200MEL_OK;  // (an empty gather still runs the barrier: every rank must launch every step)
    int sms = 0;
    if (int rc = current_sms(&sms)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    a.pull = use_copy_engines ? 0 : 1;
    if (use_tma) {
        static const int stages = [] {   // B200MEL_PULL_STAGES: ring depth of the TMA pull (A/B measurements), default 4
            const char *e = getenv("B200MEL_PULL_STAGES");
            const int v = e ? atoi(e) : kPullStages;
            return v < 2 ? 2 : (v > kPullMaxStages ? kPullMaxStages : v);
        }();
        a.stages = stages;
        cudaFuncSetAttribute(gather_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPullMaxStages * kPullChunk);  // per device
        gather_tma_kernel<<<n_ctas > 0 ? n_ctas : 32, 128, stages * kPullChunk, st>>>(local_buf, a);
    } else
        gather_pull_kernel<<<use_copy_engines ? 1 : (n_ctas > 0 ? n_ctas : sms * 2), 512, 0, st>>>(local_buf, a);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "gather launch");
    if (!use_copy_engines) return B200MEL_OK;
    // one device-to-device copy per peer, each on its own internal stream forked from (and joined back into) `stream`:
    // the copy engines move the blocks while the SMs stay with the next extraction kernel
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (dev < 0 || dev >= 64) return fail(B200MEL_EUNSUP, "gather_copy: device index out of range");
    GatherCtx &gc = g_gather[dev];
    {
        std::lock_guard<std::mutex> lock(g_gather_mu);
        if (!gc.ready) {
            for (int i = 0; i < 15 && e == cudaSuccess; ++i) {
                e = cudaStreamCreateWithFlags(&gc.streams[i], cudaStreamNonBlocking);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&gc.join[i], cudaEventDisableTiming);
            }
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&gc.fork, cudaEventDisableTiming);
            if (e != cudaSuccess) return cuda_fail(e, "gather_copy: creating streams / events");
            gc.ready = true;
        }
    }
    if ((e = cudaEventRecord(gc.fork, st)) != cudaSuccess) return cuda_fail(e, "cudaEventRecord");
    for (int i = 1; i < world && e == cudaSuccess; ++i) {
        const int r = (rank + i) % world;
        const size_t bytes = (size_t)(a.off[r + 1] - a.off[r]) * sizeof(float);
        if (!bytes) continue;
        cudaStream_t cs = gc.streams[i - 1];
        e = cudaStreamWaitEvent(cs, gc.fork, 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(local_buf + a.off[r], a.peer[r] + a.off[r], bytes, cudaMemcpyDeviceToDevice, cs);
        if (e == cudaSuccess) e = cudaEventRecord(gc.join[i - 1], cs);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, gc.join[i - 1], 0);
    }
    return e == cudaSuccess ? B200MEL_OK : cuda_fail(e, "gather_copy: peer copies");
}

int b200mel_mel_to_mfcc(const float *mel, const float *dct, int64_t B, int32_t n_mels, int32_t n_mfcc, int64_t T,
                        float *out, void *stream) {
    if (B < 0 || T < 0 || n_mels <= 0 || n_mfcc <= 0) return fail(B200MEL_EINVAL, "mel_to_mfcc: bad shape");
    if (B == 0 || T == 0) return B200MEL_OK;
    if (!mel || !dct || !out) return fail(B200MEL_EINVAL, "mel_to_mfcc: null pointer");
    const size_t sm = (size_t)((n_mels + 3) & ~3) * n_mfcc * 4;
    if (n_mels > 128 || sm > 48 * 1024 || T > 0x7fffffff)
        return fail(B200MEL_EUNSUP, "mel_to_mfcc: supports n_mels <= 128 and a DCT matrix of at most 48 KB");
    int sms = 0;
    if (int rc = current_sms(&sms)) return rc;
    // Common shapes (the reference's MFCC_SIZE x MEL_SIZE = 40 x 80, settings.py:15-16, and any n_mfcc <= 64 over 40 / 80 /
    // 128 mels): DCT matrix staged in constant memory by a stream-ordered device-to-device copy, weights as
    // constant-bank operands.  (Calls on DIFFERENT streams with DIFFERENT matrices must not overlap: one staging buffer.)
    if (!getenv("B200MEL_DCT_SMEM") && n_mfcc <= kDctConstRows && (n_mels == 40 || n_mels == 80 || n_mels == 128)) {
        cudaError_t ce = cudaMemcpyToSymbolAsync(c_dct, dct, (size_t)n_mfcc * n_mels * sizeof(float), 0,
                                                 cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
        if (ce != cudaSuccess) return cuda_fail(ce, "mel_to_mfcc: staging the DCT matrix");
        const int g = grid_for(B * T, 128, sms);
        if (n_mels == 40) dct_const_kernel<40><<<g, 128, 0, (cudaStream_t)stream>>>(mel, out, B, n_mfcc, (int)T);
        else if (n_mels == 80) dct_const_kernel<80><<<g, 128, 0, (cudaStream_t)stream>>>(mel, out, B, n_mfcc, (int)T);
        else dct_const_kernel<128><<<g, 128, 0, (cudaStream_t)stream>>>(mel, out, B, n_mfcc, (int)T);
        g_launches.fetch_add(1);
        cudaError_t le = cudaGetLastError();
        return le == cudaSuccess ? B200MEL_OK : cuda_fail(le, "mel_to_mfcc launch");
    }
    const int grid = grid_for(B * T * ((n_mfcc + kDctRows - 1) / kDctRows), 128, sms);  // a thread = one column x 8 rows
    if (n_mels <= 80)
        dct_kernel<80><<<grid, 128, sm, (cudaStream_t)stream>>>(mel, dct, out, B, n_mels, n_mfcc, (int)T);
    else
        dct_kernel<128><<<grid, 128, sm, (cudaStream_t)stream>>>(mel, dct, out, B, n_mels, n_mfcc, (int)T);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200MEL_OK : cuda_fail(e, "mel_to_mfcc launch");
}

}  // extern "C"
