// fft32.cuh — register-resident 32-point complex FFT building block (sm_100a).
//
// One thread owns 32 complex values in registers; every index below is a
// compile-time constant (static_for over std::integral_constant), so the
// arrays never touch local memory and the trivial twiddles (1, -i, -1,
// (+-1+-i)/sqrt2) are removed at compile time.
//
// Forward transform, engineering sign:  A[k] = sum_n a[n] * exp(-2*pi*i*n*k/32).
//
// Decomposition 32 = 4 x 8 (decimation in frequency), 8 = 2 x 4:
//   n = 8p + q, k = r + 4s:  A[r+4s] = sum_q W32^{qr} [ sum_p a[8p+q] W4^{pr} ] W8^{qs}
// computed in place; output A[k] ends at slot fft32_pos(k) (a digit reversal
// that is its own inverse), so callers index results with fft32_pos().
#pragma once
#include <cuda_runtime.h>
#include <type_traits>

namespace b200mel {

template <int I>
using IC = std::integral_constant<int, I>;

template <int B, int E, class F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (B < E) {
        f(IC<B>{});
        static_for<B + 1, E>(f);
    }
}

struct TwConst {
    // cos(2*pi*j/32), -sin(2*pi*j/32) for j = 0..23
    static constexpr float c32[24] = {
        1.0f, 0.9807852506637573f, 0.9238795042037964f, 0.8314695954322815f, 0.7071067690849304f,
        0.5555702447891235f, 0.3826834261417389f, 0.19509032368659973f, 0.0f, -0.19509032368659973f,
        -0.3826834261417389f, -0.5555702447891235f, -0.7071067690849304f, -0.8314695954322815f,
        -0.9238795042037964f, -0.9807852506637573f, -1.0f, -0.9807852506637573f, -0.9238795042037964f,
        -0.8314695954322815f, -0.7071067690849304f, -0.5555702447891235f, -0.3826834261417389f,
        -0.19509032368659973f};
    static constexpr float s32[24] = {
        0.0f, -0.19509032368659973f, -0.3826834261417389f, -0.5555702447891235f, -0.7071067690849304f,
        -0.8314695954322815f, -0.9238795042037964f, -0.9807852506637573f, -1.0f, -0.9807852506637573f,
        -0.9238795042037964f, -0.8314695954322815f, -0.7071067690849304f, -0.5555702447891235f,
        -0.3826834261417389f, -0.19509032368659973f, 0.0f, 0.19509032368659973f, 0.3826834261417389f,
        0.5555702447891235f, 0.7071067690849304f, 0.8314695954322815f, 0.9238795042037964f,
        0.9807852506637573f};
    // cos(2*pi*j/64), -sin(2*pi*j/64) for j = 0..16 (split-mode post twiddle W_2048^{32*k2} = W_64^{k2})
    static constexpr float c64[17] = {1.0f, 0.9951847195625305f, 0.9807852506637573f, 0.9569403529167175f,
                                      0.9238795042037964f, 0.8819212913513184f, 0.8314695954322815f,
                                      0.7730104327201843f, 0.7071067690849304f, 0.6343932747840881f,
                                      0.5555702447891235f, 0.4713967442512512f, 0.3826834261417389f,
                                      0.290284663438797f, 0.19509032368659973f, 0.0980171412229538f, 0.0f};
    static constexpr float s64[17] = {0.0f, -0.0980171412229538f, -0.19509032368659973f, -0.290284663438797f,
                                      -0.3826834261417389f, -0.4713967442512512f, -0.5555702447891235f,
                                      -0.6343932747840881f, -0.7071067690849304f, -0.7730104327201843f,
                                      -0.8314695954322815f, -0.8819212913513184f, -0.9238795042037964f,
                                      -0.9569403529167175f, -0.9807852506637573f, -0.9951847195625305f, -1.0f};
};

// Complex arithmetic on Blackwell's packed FP32x2 pipe: one FADD2 / FMUL2 / FFMA2 (sm_100 SASS) handles the real
// and imaginary halves of a register pair in ONE issue slot, and ptxas folds half swaps (.LO_HI), per-half sign
// patterns (.NP / .PN) and scalar broadcasts (.F32) into operand modifiers — so conj(), multiplication by -i and
// the (c, c) / (-s, s) operands of a complex multiply cost no instructions of their own.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// a * w = a * (w.x, w.x) + (a.y, a.x) * (-w.y, w.y): FMUL2 + FFMA2
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    return __ffma2_rn(a, make_float2(w.x, w.x), __fmul2_rn(make_float2(a.y, a.x), make_float2(-w.y, w.y)));
}
// a * (-i)
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }

// a *= W32^J with the trivial cases folded
template <int J>
__device__ __forceinline__ float2 twiddle32(float2 a) {
    constexpr float h = 0.7071067690849304f;
    if constexpr (J == 0) {
        return a;
    } else if constexpr (J == 8) {
        return cmul_mi(a);
    } else if constexpr (J == 16) {
        return make_float2(-a.x, -a.y);
    } else if constexpr (J == 4) {  // (1 - i)/sqrt2:  ((a.x + a.y) h, (a.y - a.x) h)
        return __fmul2_rn(__fadd2_rn(a, make_float2(a.y, -a.x)), make_float2(h, h));
    } else if constexpr (J == 12) {  // (-1 - i)/sqrt2: ((a.y - a.x) h, -(a.x + a.y) h)
        return __fmul2_rn(__fadd2_rn(make_float2(a.y, a.x), make_float2(-a.x, a.y)), make_float2(h, -h));
    } else {
        constexpr float c = TwConst::c32[J];
        constexpr float s = TwConst::s32[J];
        return cmul(a, make_float2(c, s));
    }
}

// 4-point forward FFT in place: (e0,e1,e2,e3) -> (E0,E1,E2,E3), natural order
__device__ __forceinline__ void fft4(float2 &e0, float2 &e1, float2 &e2, float2 &e3) {
    float2 t0 = cadd(e0, e2), t1 = csub(e0, e2), t2 = cadd(e1, e3), t3 = cmul_mi(csub(e1, e3));
    e0 = cadd(t0, t2);
    e2 = csub(t0, t2);
    e1 = cadd(t1, t3);
    e3 = csub(t1, t3);
}

// 8-point forward FFT in place on c[0..7]; output C[g + 2h] lands at slot 4g + h.
__device__ __forceinline__ void fft8(float2 *c) {
    static_for<0, 4>([&](auto v_) {
        constexpr int v = decltype(v_)::value;
        float2 s = cadd(c[v], c[4 + v]);
        float2 d = csub(c[v], c[4 + v]);
        c[v] = s;
        c[4 + v] = twiddle32<4 * v>(d);  // W8^v = W32^{4v}
    });
    fft4(c[0], c[1], c[2], c[3]);
    fft4(c[4], c[5], c[6], c[7]);
}

// slot of output bin k after fft32 (involution)
__host__ __device__ constexpr int fft32_pos(int k) { return 8 * (k & 3) + (k & 4) + (k >> 3); }

// 32-point forward FFT in place.  Output A[k] is at a[fft32_pos(k)].
__device__ __forceinline__ void fft32(float2 *a) {
    // radix-4 over p for each q; b_q[r] -> a[q + 8r]
    static_for<0, 8>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        fft4(a[q], a[q + 8], a[q + 16], a[q + 24]);
        a[q + 8] = twiddle32<q * 1>(a[q + 8]);
        a[q + 16] = twiddle32<q * 2>(a[q + 16]);
        a[q + 24] = twiddle32<q * 3>(a[q + 24]);
    });
    // 8-point FFT over q for each r
    fft8(a + 0);
    fft8(a + 8);
    fft8(a + 16);
    fft8(a + 24);
}

}  // namespace b200mel
