// FFT v2 core: Stockham autosort, radix 16 first (two radix-4 layers in registers), 16 points
// per thread, remainder radix (2/4/8) last.  Written as host/device code so that the index
// arithmetic (pass structure, digit order, shared-memory padding, bank mapping) is also
// executed on the CPU by tests/emul/fft2_emul.cpp, thread by thread, without a GPU.
//
// Reference semantics (src/fft/mod.rs:73-96): X[k] = sum_n x[n] e^{-/+ j 2 pi k n / N},
// unnormalised in both directions.
//
// Pass p (radix R, Ns = product of the earlier radices), butterfly index jj in [0, N/R):
//   inputs   a[jj + r*N/R]                       r = 0..R-1
//   twiddle  w^r,  w = e^{-/+ 2 pi i (jj mod Ns) / (R*Ns)}
//   outputs  b[(jj / Ns)*Ns*R + (jj mod Ns) + q*Ns]   q = 0..R-1
// A thread (index j in [0, T), T = N/16) holds v[m] = a[j + m*T], m = 0..15, which is exactly
// the input set of 16/R butterflies jj = j + h*T (h = 0..16/R-1; input r of butterfly h is
// v[h + (16/R)*r]).  Shared memory index i is stored at i + (i >> 4): every access pattern
// below is then conflict-free per half-warp for 8-byte words.
#pragma once

#if defined(__CUDACC__)
#define CB_HD __host__ __device__ __forceinline__
#define CB_HDC __host__ __device__ constexpr
#else
#include <cuda_runtime.h>  // float2 / make_float2 only
#define CB_HD inline
#define CB_HDC constexpr
#endif

namespace cb {
namespace fft2 {

// ------------------------------------------------------------------ complex helpers
CB_HD float2 cadd(float2 a, float2 b)
{
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
CB_HD float2 csub(float2 a, float2 b)
{
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(b, make_float2(-1.f, -1.f), a);
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
CB_HD float2 cmul(float2 a, float2 b)
{
#if defined(__CUDA_ARCH__)
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
#else
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
#endif
}
CB_HD float2 csqr(float2 a)
{
#if defined(__CUDA_ARCH__)
    return make_float2(fmaf(a.x, a.x, -a.y * a.y), (a.x + a.x) * a.y);
#else
    return make_float2(a.x * a.x - a.y * a.y, (a.x + a.x) * a.y);
#endif
}
template <bool INV>
CB_HD float2 mul_mi(float2 a)  // a * (-i) forward, a * (+i) inverse
{
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// a * e^{-/+ i pi/4} * sqrt(2)/sqrt(2): (1 -/+ i)/sqrt2
template <bool INV>
CB_HD float2 mul_w8(float2 a)
{
    const float c = 0.70710678118654752440f;
    return INV ? make_float2(c * (a.x - a.y), c * (a.x + a.y)) : make_float2(c * (a.x + a.y), c * (a.y - a.x));
}
// a * e^{-/+ 3 i pi/4}: (-1 -/+ i)/sqrt2
template <bool INV>
CB_HD float2 mul_w8_3(float2 a)
{
    const float c = 0.70710678118654752440f;
    return INV ? make_float2(-c * (a.x + a.y), c * (a.x - a.y)) : make_float2(c * (a.y - a.x), -c * (a.x + a.y));
}
// a * (cr -/+ i ci): constant rotation by e^{-/+ i theta}, (cr, ci) = (cos theta, sin theta)
template <bool INV>
CB_HD float2 mul_c(float2 a, float cr, float ci)
{
    return INV ? make_float2(a.x * cr - a.y * ci, a.y * cr + a.x * ci) : make_float2(a.x * cr + a.y * ci, a.y * cr - a.x * ci);
}

template <bool INV>
CB_HD void bfly2(float2 &x0, float2 &x1)
{
    const float2 a = x0;
    x0 = cadd(a, x1);
    x1 = csub(a, x1);
}

template <bool INV>
CB_HD void bfly4(float2 &x0, float2 &x1, float2 &x2, float2 &x3)  // natural order in and out
{
    const float2 b0 = cadd(x0, x2), b2 = csub(x0, x2), b1 = cadd(x1, x3), b3 = mul_mi<INV>(csub(x1, x3));
    x0 = cadd(b0, b1);
    x2 = csub(b0, b1);
    x1 = cadd(b2, b3);
    x3 = csub(b2, b3);
}

template <bool INV>
CB_HD void bfly8(float2 &v0, float2 &v1, float2 &v2, float2 &v3, float2 &v4, float2 &v5, float2 &v6, float2 &v7)
{
    float2 a0 = cadd(v0, v4), a4 = csub(v0, v4);
    float2 a1 = cadd(v1, v5), a5 = mul_w8<INV>(csub(v1, v5));
    float2 a2 = cadd(v2, v6), a6 = mul_mi<INV>(csub(v2, v6));
    float2 a3 = cadd(v3, v7), a7 = mul_w8_3<INV>(csub(v3, v7));
    bfly4<INV>(a0, a1, a2, a3);  // X0 X2 X4 X6
    bfly4<INV>(a4, a5, a6, a7);  // X1 X3 X5 X7
    v0 = a0; v2 = a1; v4 = a2; v6 = a3;
    v1 = a4; v3 = a5; v5 = a6; v7 = a7;
}

// 16-point DFT in registers (two radix-4 layers).  In: v[m] = x[m].  Layer 1 (over m1, m = 4*m1 + m0)
// leaves Y[m0][a] in v[m0 + 4a]; after the W16^{m0*a} rotations layer 2 (over m0) leaves
// X[a + 4b] in v[b + 4a] -- callers index the result with q16(slot).
template <bool INV>
CB_HD void bfly16(float2 *v)
{
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
#pragma unroll
    for (int m0 = 0; m0 < 4; ++m0) bfly4<INV>(v[m0], v[m0 + 4], v[m0 + 8], v[m0 + 12]);
    // now v[m0 + 4a] = Y[m0][a]
    v[1 + 4 * 1] = mul_c<INV>(v[1 + 4 * 1], c1, s1);      // m0=1,a=1: W^1
    v[1 + 4 * 2] = mul_w8<INV>(v[1 + 4 * 2]);             // m0=1,a=2: W^2
    v[1 + 4 * 3] = mul_c<INV>(v[1 + 4 * 3], s1, c1);      // m0=1,a=3: W^3
    v[2 + 4 * 1] = mul_w8<INV>(v[2 + 4 * 1]);             // m0=2,a=1: W^2
    v[2 + 4 * 2] = mul_mi<INV>(v[2 + 4 * 2]);             // m0=2,a=2: W^4
    v[2 + 4 * 3] = mul_w8_3<INV>(v[2 + 4 * 3]);           // m0=2,a=3: W^6
    v[3 + 4 * 1] = mul_c<INV>(v[3 + 4 * 1], s1, c1);      // m0=3,a=1: W^3
    v[3 + 4 * 2] = mul_w8_3<INV>(v[3 + 4 * 2]);           // m0=3,a=2: W^6
    v[3 + 4 * 3] = mul_c<INV>(v[3 + 4 * 3], -c1, -s1);    // m0=3,a=3: W^9
    // second layer: for each a, radix-4 over m0 (inputs v[m0 + 4a]); output b replaces input m0 = b,
    // so X[a + 4b] ends in v[b + 4a].
#pragma unroll
    for (int a = 0; a < 4; ++a) bfly4<INV>(v[4 * a], v[4 * a + 1], v[4 * a + 2], v[4 * a + 3]);
}
// output index held by register slot s after bfly16: slot s = b + 4a holds X[a + 4b]
CB_HDC int q16(int s) { return (s >> 2) + 4 * (s & 3); }

// powers w^1..w^15 of a unit twiddle, applied to v[1..15]
CB_HD void twiddle16(float2 *v, float2 w1)
{
    const float2 w2 = csqr(w1), w3 = cmul(w2, w1), w4 = csqr(w2);
    v[1] = cmul(v[1], w1);
    v[2] = cmul(v[2], w2);
    v[3] = cmul(v[3], w3);
    v[4] = cmul(v[4], w4);
    const float2 w5 = cmul(w4, w1), w6 = csqr(w3), w7 = cmul(w4, w3), w8 = csqr(w4);
    v[5] = cmul(v[5], w5);
    v[6] = cmul(v[6], w6);
    v[7] = cmul(v[7], w7);
    v[8] = cmul(v[8], w8);
    v[9] = cmul(v[9], cmul(w8, w1));
    v[10] = cmul(v[10], csqr(w5));
    v[11] = cmul(v[11], cmul(w8, w3));
    v[12] = cmul(v[12], csqr(w6));
    v[13] = cmul(v[13], cmul(w8, w5));
    v[14] = cmul(v[14], csqr(w7));
    v[15] = cmul(v[15], cmul(w8, w7));
}

// v[m] *= c * w^m, m = 0..15 (power tree of depth <= 4)
CB_HD void twiddle16c(float2 *v, float2 c, float2 w1)
{
    const float2 w2 = csqr(w1), w4 = csqr(w2), w8 = csqr(w4);
    float2 u[8];
    u[0] = c;
    u[1] = cmul(c, w1);
    u[2] = cmul(c, w2);
    u[3] = cmul(u[1], w2);
    for (int i = 0; i < 4; ++i) u[4 + i] = cmul(u[i], w4);
    for (int i = 0; i < 8; ++i) {
        v[i] = cmul(v[i], u[i]);
        v[8 + i] = cmul(v[8 + i], cmul(u[i], w8));
    }
}

// ------------------------------------------------------------------ plan constants
template <int LOG2N>
struct Plan {
    static constexpr int N = 1 << LOG2N;
    static constexpr int T = N / 16;            // threads per frame
    static constexpr int P16 = LOG2N / 4;       // radix-16 passes
    static constexpr int REM = LOG2N % 4;       // log2 of the last-pass radix (0 = none)
    static constexpr int PASSES = P16 + (REM ? 1 : 0);
    static constexpr int PADN = N + N / 16;     // padded frame pitch in shared memory (float2)
    // twiddle table layout (float2 entries): pass p >= 1 of radix 16 uses Ns = 16^p entries
    // e^{-/+ 2 pi i s / 16^(p+1)} at offset tw_off(p); the remainder pass uses N/R entries
    // e^{-/+ 2 pi i s / N} at tw_off(P16).
    static CB_HDC int tw_off(int p) { return p <= 1 ? 0 : ((1 << (4 * p)) - 16) / 15; }  // sum_{q=1}^{p-1} 16^q
    static constexpr int TW_TOTAL = tw_off(P16) + (REM ? (N >> REM) : 0);
};

CB_HD int pad16(int i) { return i + (i >> 4); }

// One radix-16 pass (pass number P, Ns = 16^P) for thread j.
//   load(m)  -> a[j + m*T];  mid() is called once between the last load and the first store
//   store(idx, val) writes b[idx]
template <int LOG2N, bool INV, int P, typename LOAD, typename MID, typename STORE>
CB_HD void pass16(int j, const float2 *tw, LOAD load, MID mid, STORE store)
{
    using PL = Plan<LOG2N>;
    constexpr int Ns = 1 << (4 * P);
    constexpr int off = PL::tw_off(P);
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = load(m);
    mid();
    if (P > 0) twiddle16(v, tw[off + (j & (Ns - 1))]);
    bfly16<INV>(v);
    const int d = ((j >> (4 * P)) << (4 * P + 4)) + (j & (Ns - 1));
#pragma unroll
    for (int sl = 0; sl < 16; ++sl) store(d + q16(sl) * Ns, v[sl]);
}

// Remainder pass (radix R = 2^REM, Ns = N/R, always the last pass): 16/R butterflies
// jj = j + h*T per thread; X[jj + q*N/R] = X[j + (h + G*q)*T] -> store(j + m*T) with m = h + G*q.
template <int LOG2N, bool INV, typename LOAD, typename MID, typename STORE>
CB_HD void pass_rem(int j, const float2 *tw, LOAD load, MID mid, STORE store)
{
    using PL = Plan<LOG2N>;
    constexpr int R = 1 << PL::REM;
    constexpr int G = 16 / R;
    constexpr int off = PL::tw_off(PL::P16);
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = load(m);
    mid();
#pragma unroll
    for (int h = 0; h < G; ++h) {
        const float2 w1 = tw[off + j + h * PL::T];
        if constexpr (R == 2) {
            v[h + G] = cmul(v[h + G], w1);
            bfly2<INV>(v[h], v[h + G]);
        } else if constexpr (R == 4) {
            const float2 w2 = csqr(w1), w3 = cmul(w2, w1);
            v[h + G] = cmul(v[h + G], w1);
            v[h + 2 * G] = cmul(v[h + 2 * G], w2);
            v[h + 3 * G] = cmul(v[h + 3 * G], w3);
            bfly4<INV>(v[h], v[h + G], v[h + 2 * G], v[h + 3 * G]);
        } else if constexpr (R == 8) {
            const float2 w2 = csqr(w1), w3 = cmul(w2, w1), w4 = csqr(w2);
            v[h + G] = cmul(v[h + G], w1);
            v[h + 2 * G] = cmul(v[h + 2 * G], w2);
            v[h + 3 * G] = cmul(v[h + 3 * G], w3);
            v[h + 4 * G] = cmul(v[h + 4 * G], w4);
            v[h + 5 * G] = cmul(v[h + 5 * G], cmul(w4, w1));
            v[h + 6 * G] = cmul(v[h + 6 * G], csqr(w3));
            v[h + 7 * G] = cmul(v[h + 7 * G], cmul(w4, w3));
            bfly8<INV>(v[h], v[h + G], v[h + 2 * G], v[h + 3 * G], v[h + 4 * G], v[h + 5 * G], v[h + 6 * G],
                       v[h + 7 * G]);
        }
    }
#pragma unroll
    for (int m = 0; m < 16; ++m) store(j + m * PL::T, v[m]);
}

// Pass number PASS of the transform for thread j of one frame.
//   gld(i) / gst(i, v): the frame in global memory (first pass reads it, last pass writes it)
//   sin.ld(i) / sout.st(i, v): shared-memory accessors of the frame (logical index i; they apply
//   pad16 themselves).  sin and sout may be the same buffer; then mid() must be a barrier.
template <int LOG2N, bool INV, int PASS, typename GLD, typename GST, typename SIN, typename SOUT, typename MID>
CB_HD void run_pass(int j, const float2 *tw, GLD gld, GST gst, SIN sin, SOUT sout, MID mid)
{
    using PL = Plan<LOG2N>;
    constexpr int T = PL::T;
    constexpr bool first = PASS == 0, last = PASS == PL::PASSES - 1;
    auto nomid = [] {};
    auto ld_g = [&](int m) { return gld(j + m * T); };
    auto ld_s = [&](int m) { return sin.ld(j + m * T); };
    auto st_s = [&](int idx, float2 val) { sout.st(idx, val); };
    auto st_g = [&](int idx, float2 val) { gst(idx, val); };
    if constexpr (PASS < PL::P16) {
        if constexpr (first && last) pass16<LOG2N, INV, PASS>(j, tw, ld_g, nomid, st_g);
        else if constexpr (first) pass16<LOG2N, INV, PASS>(j, tw, ld_g, nomid, st_s);
        else if constexpr (last) pass16<LOG2N, INV, PASS>(j, tw, ld_s, nomid, st_g);
        else pass16<LOG2N, INV, PASS>(j, tw, ld_s, mid, st_s);
    } else {
        if constexpr (first) pass_rem<LOG2N, INV>(j, tw, ld_g, nomid, st_g);
        else pass_rem<LOG2N, INV>(j, tw, ld_s, nomid, st_g);
    }
}

}  // namespace fft2
}  // namespace cb
