// stft_tc_tables.h — host-side construction of the constant operands of stft_tc_kernel (stft_tc.cuh): the two DFT
// matrices as fp16 hi / lo limbs in the UMMA K-major core-matrix layout, the inter-stage twiddles and the banded mel
// schedule.  Pure host code (no CUDA calls): b200mel_debug_tc_tables exports the blob so tests/test_tc_algebra.py can
// check it against the numpy model (tests/tc_model.py) without a GPU.
#pragma once
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "stft_tc.cuh"

namespace b200mel {

struct StcTables {
    std::vector<unsigned char> blob;  // B1 | B2 | tw | mel schedule  (the order stft_tc_kernel's prologue copies them in)
    int mel_bytes = 0;                // header + weights, a multiple of 16
    int n_groups = 0, max_len = 0, warp_cost_max = 0;
};

static inline uint16_t stc_f16_bits(float x) {
    const __half h = __float2half_rn(x);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
}
static inline float stc_f16_value(uint16_t u) {
    __half h;
    memcpy(&h, &u, 2);
    return __half2float(h);
}
// the two limbs of a constant: hi = fp16(x), lo = fp16(x - hi)
static inline void stc_split(double x, uint16_t *hi, uint16_t *lo) {
    const float xf = (float)x;
    *hi = stc_f16_bits(xf);
    *lo = stc_f16_bits((float)(x - (double)stc_f16_value(*hi)));
}

static inline int stc_k2_of_j(int j) { return j < 12 ? j : j + 8; }

// Returns false (with *why set) when the filterbank does not fit the kernel: more than 128 rows or a non-zero at a
// bin >= 384.  W is (n_mels, F) row-major in LOGICAL bins of the 1024-point transform (F = 513).
static bool stc_build_tables(const float *W, int n_mels, int F, StcTables *out, const char **why) {
    const double two_pi = 6.283185307179586476925286766559;
    if (n_mels < 1 || n_mels > kStcMaxMels) { *why = "more than 128 mel rows"; return false; }
    struct Row { int m, first, last; };
    std::vector<Row> rows;
    for (int m = 0; m < n_mels; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < F; ++k)
            if (W[(size_t)m * F + k] != 0.f) {
                if (first < 0) first = k;
                last = k;
            }
        if (last >= kStcBins) { *why = "filterbank reaches bin 384 or above"; return false; }
        if (first < 0) first = last = 0;
        rows.push_back({m, first, last});
    }
    std::stable_sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) { return a.last - a.first > b.last - b.first; });

    out->blob.assign(kStcB1Bytes + kStcB2Bytes + kStcTwBytes, 0);
    uint16_t *b1 = reinterpret_cast<uint16_t *>(out->blob.data());
    uint16_t *b2 = reinterpret_cast<uint16_t *>(out->blob.data() + kStcB1Bytes);
    float *tw = reinterpret_cast<float *>(out->blob.data() + kStcB1Bytes + kStcB2Bytes);

    // ---- stage 1: [n1 (K = 32)] x [col]: col 0 = A[0].re, col 1 = A[16].re, col 2 k1 / 2 k1 + 1 = A[k1].re / .im
    //      stored as B operand rows n = 0..31 (hi limb) and 32..63 (lo limb): elem(n, k) at ((k / 8) * 64 + n) * 8 + k % 8
    for (int k = 0; k < 32; ++k)
        for (int col = 0; col < 32; ++col) {
            double v;
            if (col == 0) v = 1.0;
            else if (col == 1) v = (k & 1) ? -1.0 : 1.0;
            else {
                const int k1 = col >> 1;
                const double th = two_pi * (double)((k1 * k) % 32) / 32.0;
                v = (col & 1) ? -sin(th) : cos(th);
            }
            uint16_t hi, lo;
            stc_split(v, &hi, &lo);
            b1[((k / 8) * 64 + col) * 8 + k % 8] = hi;
            b1[((k / 8) * 64 + 32 + col) * 8 + k % 8] = lo;
        }
    // ---- stage 2: K = (n2, c) = 64, N = 192 = [B hi | B' hi | B lo | B' lo], 48 columns (j, re / im) each
    for (int n2 = 0; n2 < 32; ++n2)
        for (int c = 0; c < 2; ++c)
            for (int j = 0; j < 24; ++j)
                for (int ri = 0; ri < 2; ++ri) {
                    const double th = two_pi * (double)((n2 * stc_k2_of_j(j)) % 32) / 32.0;
                    // (cos - i sin)(a_re + i a_im): re = cos a_re + sin a_im, im = -sin a_re + cos a_im
                    const double vb = ri == 0 ? (c == 0 ? cos(th) : sin(th)) : (c == 0 ? -sin(th) : cos(th));
                    double vp = 0.0;  // B': the packed row — re slot -> X[32 j] (j < 12), im slot -> X[16 + 32 (23 - j)] (j >= 12)
                    if (j < 12 && c == 0) vp = ri == 0 ? cos(th) : -sin(th);
                    if (j >= 12 && c == 1) {
                        const double ph = two_pi * (double)((n2 * (16 + 32 * (23 - j))) % 1024) / 1024.0;
                        vp = ri == 0 ? cos(ph) : -sin(ph);
                    }
                    const int k = 2 * n2 + c, col = 2 * j + ri;
                    uint16_t hi, lo;
                    stc_split(vb, &hi, &lo);
                    b2[((k / 8) * 192 + col) * 8 + k % 8] = hi;
                    b2[((k / 8) * 192 + 96 + col) * 8 + k % 8] = lo;
                    stc_split(vp, &hi, &lo);
                    b2[((k / 8) * 192 + 48 + col) * 8 + k % 8] = hi;
                    b2[((k / 8) * 192 + 144 + col) * 8 + k % 8] = lo;
                }
    // ---- twiddles tw[k1][n2] = 2^-kStcShift W1024^(k1 n2)
    const double S = 1.0 / (double)(1 << kStcShift);
    for (int k1 = 0; k1 < 17; ++k1)
        for (int n2 = 0; n2 < 32; ++n2) {
            const double ang = two_pi * (double)((k1 * n2) % 1024) / 1024.0;
            tw[(k1 * 32 + n2) * 2] = (float)(S * cos(ang));
            tw[(k1 * 32 + n2) * 2 + 1] = (float)(-S * sin(ang));
        }
    // ---- mel schedule: one row per thread of the mel warps, rows sorted by length (longest first) so that the rows of
    //      a warp have similar lengths; every lane of warp v runs trip[v] bins.  Read windows are slid down so that
    //      lane l starts at a bin = l (mod 4) — the 32-byte tile rows of neighbouring lanes then spread over the banks —
    //      and kept below bin 384; weights are stored [warp][bin step][lane].
    std::vector<int> header(kStcMelHeader / 4, 0);
    int *trip = header.data(), *woff = trip + kStcMelWarps, *ent = woff + kStcMelWarps;  // ent: {m, lo} pairs
    std::vector<float> w;
    while ((int)rows.size() < kStcMaxMels) rows.push_back({-1, 0, 0});
    for (int v = 0; v < kStcMelWarps; ++v) {
        int lo[32], len = 0;
        bool any = false;
        for (int l = 0; l < 32; ++l) {
            const Row &row = rows[v * 32 + l];
            lo[l] = row.first - (((row.first - l) % 4) + 4) % 4;  // <= first, = l (mod 4), >= -3
            if (row.m >= 0) any = true, len = std::max(len, row.last - lo[l] + 1);
        }
        for (bool moved = any; moved;) {  // windows must end inside the tile: slide down in steps of 4
            moved = false;
            for (int l = 0; l < 32; ++l)
                while (lo[l] + len > kStcBins) {
                    lo[l] -= 4;
                    if (rows[v * 32 + l].m >= 0) len = std::max(len, rows[v * 32 + l].last - lo[l] + 1);
                    moved = true;
                }
        }
        trip[v] = len;
        woff[v] = (int)w.size();
        for (int i = 0; i < len; ++i)
            for (int l = 0; l < 32; ++l) {
                const Row &row = rows[v * 32 + l];
                const int bin = lo[l] + i;
                w.push_back((row.m >= 0 && bin >= row.first && bin <= row.last) ? W[(size_t)row.m * F + bin] : 0.f);
            }
        for (int l = 0; l < 32; ++l) ent[(v * 32 + l) * 2] = rows[v * 32 + l].m, ent[(v * 32 + l) * 2 + 1] = lo[l];
        out->max_len = std::max(out->max_len, len);
        out->warp_cost_max = std::max(out->warp_cost_max, len);
    }
    out->n_groups = kStcMelWarps;
    while (w.size() % 4) w.push_back(0.f);
    out->mel_bytes = kStcMelHeader + (int)w.size() * 4;
    const size_t base = out->blob.size();
    out->blob.resize(base + (size_t)out->mel_bytes);
    memcpy(out->blob.data() + base, header.data(), kStcMelHeader);
    memcpy(out->blob.data() + base + kStcMelHeader, w.data(), w.size() * 4);
    return true;
}

}  // namespace b200mel
