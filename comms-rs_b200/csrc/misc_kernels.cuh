// Element-wise kernels and the device math shared with the fused chain kernel.
#pragma once
#include "common.cuh"

namespace cb {

int launch_mixer(const float2 *x, float2 *y, size_t n, double phase0, double dphase, cudaStream_t s);
int launch_fm(const float2 *x, float *out, size_t n, const float2 *prev_in, float2 *prev_out, cudaStream_t s);
int launch_decimate(const void *in, void *out, size_t n_out, size_t elem, size_t rate, cudaStream_t s);
int launch_upsample(const void *in, void *out, size_t n_out, size_t elem, size_t rate, cudaStream_t s);
int launch_bits_to_symbols(const uint8_t *bits, float2 *sym, size_t nsym, int mode, cudaStream_t s);
int launch_quantize_i16(const float *in, int16_t *out, size_t n, float scale, cudaStream_t s);
int launch_real_to_complex(const float *in, float2 *out, size_t n, cudaStream_t s);
int launch_complex_real(const float2 *in, float *out, size_t n, cudaStream_t s);
int launch_convert_u8(const uint8_t *in, float *out, size_t nfloats, cudaStream_t s);
int launch_convert_i16(const int16_t *in, float *out, size_t nfloats, float scale, cudaStream_t s);
int launch_synth(float *out, size_t nfloats, unsigned long long base, cudaStream_t s);
// Nco::push over a batch (nco_kernel.cu): tile_scratch holds nco_scratch_doubles(n) doubles
size_t nco_scratch_doubles(size_t n);
int launch_nco(const double *perr, size_t n, double *tile_scratch, const double *phase_in, double *phase_out, double dphase,
               double2 *out, cudaStream_t s);

#ifdef __CUDACC__
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// (cos, sin) of an f64 phase: Cody-Waite reduction to [-pi, pi] in f64, then a
// two-term f32 evaluation (sincosf of the high part, first-order correction by
// the low part) -- absolute error ~1e-7, i.e. at the rounding level of the f32
// output the reference produces after its f64 multiply (src/mixer.rs:77-83).
__device__ __forceinline__ float2 phase_rotation(double theta)
{
    const double inv2pi = 0.15915494309189534561;
    const double twopi_hi = 6.283185307179586232;    // fl(2*pi)
    const double twopi_lo = 2.4492935982947064e-16;  // 2*pi - twopi_hi
    const double k = rint(theta * inv2pi);
    double r = fma(-k, twopi_hi, theta);
    r = fma(-k, twopi_lo, r);
    const float hi = (float)r;
    const float lo = (float)(r - (double)hi);
    float s, c;
    sincosf(hi, &s, &c);
    return make_float2(fmaf(-lo, s, c), fmaf(lo, c, s));
}

// FM discriminator step: theta = s * conj(p); atan2(theta.im, theta.re), with the
// reference's operation order and no FMA contraction (src/modulation/analog.rs:27-29).
__device__ __forceinline__ float fm_angle(float2 s, float2 p)
{
    const float br = p.x, bi = -p.y;
    const float tr = __fsub_rn(__fmul_rn(s.x, br), __fmul_rn(s.y, bi));
    const float ti = __fadd_rn(__fmul_rn(s.x, bi), __fmul_rn(s.y, br));
    return atan2f(ti, tr);
}

// Same discriminator step with a branch-free atan2: a = min/max in [0, 1], odd polynomial of degree 17 in a
// (max error 1.1e-7 rad in f32), quadrant from |im| > |re| and the SIGN BITS of re / im, so that signed zeros pick
// the branch libm's atan2 picks (atan2(+0, -0) = pi, atan2(-0, +0) = -0).  About half the instructions of atan2f and
// no divergence; used by the fused chain kernel where the FM step shares issue slots with the filter.
__device__ __forceinline__ float fm_angle_fast(float2 s, float2 p)
{
    const float br = p.x, bi = -p.y;
    const float tr = __fsub_rn(__fmul_rn(s.x, br), __fmul_rn(s.y, bi));
    const float ti = __fadd_rn(__fmul_rn(s.x, bi), __fmul_rn(s.y, br));
    const float ax = fabsf(tr), ay = fabsf(ti);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = mx > 0.f ? __fdividef(mn, mx) : 0.f;
    const float z = a * a;
    float q = 0.0028340641874819994f;
    q = fmaf(q, z, -0.016005029901862144f);
    q = fmaf(q, z, 0.042587608098983765f);
    q = fmaf(q, z, -0.07495445758104324f);
    q = fmaf(q, z, 0.10636754333972931f);
    q = fmaf(q, z, -0.14202570915222168f);
    q = fmaf(q, z, 0.19992484152317047f);
    q = fmaf(q, z, -0.3333306610584259f);
    q = fmaf(q, z, 1.0f);
    float r = a * q;
    r = ay > ax ? 1.57079632679489662f - r : r;
    r = (__float_as_uint(tr) >> 31) ? 3.14159265358979324f - r : r;
    return __uint_as_float(__float_as_uint(r) | (__float_as_uint(ti) & 0x80000000u));
}
// Two discriminator steps at once with the polynomial in packed f32 (FFMA2): out.x = angle(s0, p0), out.y = angle(s1, p1).
// Same products, rounding points, polynomial and quadrant logic as fm_angle_fast: bit-identical results.
__device__ __forceinline__ float2 fm_angle_fast2(float2 s0, float2 p0, float2 s1, float2 p1)
{
    const float tr0 = __fsub_rn(__fmul_rn(s0.x, p0.x), __fmul_rn(s0.y, -p0.y));
    const float ti0 = __fadd_rn(__fmul_rn(s0.x, -p0.y), __fmul_rn(s0.y, p0.x));
    const float tr1 = __fsub_rn(__fmul_rn(s1.x, p1.x), __fmul_rn(s1.y, -p1.y));
    const float ti1 = __fadd_rn(__fmul_rn(s1.x, -p1.y), __fmul_rn(s1.y, p1.x));
    const float ax0 = fabsf(tr0), ay0 = fabsf(ti0), ax1 = fabsf(tr1), ay1 = fabsf(ti1);
    const float mx0 = fmaxf(ax0, ay0), mn0 = fminf(ax0, ay0), mx1 = fmaxf(ax1, ay1), mn1 = fminf(ax1, ay1);
    const float2 a = make_float2(mx0 > 0.f ? __fdividef(mn0, mx0) : 0.f, mx1 > 0.f ? __fdividef(mn1, mx1) : 0.f);
    const float2 z = __fmul2_rn(a, a);
    float2 q = make_float2(0.0028340641874819994f, 0.0028340641874819994f);
    q = __ffma2_rn(q, z, make_float2(-0.016005029901862144f, -0.016005029901862144f));
    q = __ffma2_rn(q, z, make_float2(0.042587608098983765f, 0.042587608098983765f));
    q = __ffma2_rn(q, z, make_float2(-0.07495445758104324f, -0.07495445758104324f));
    q = __ffma2_rn(q, z, make_float2(0.10636754333972931f, 0.10636754333972931f));
    q = __ffma2_rn(q, z, make_float2(-0.14202570915222168f, -0.14202570915222168f));
    q = __ffma2_rn(q, z, make_float2(0.19992484152317047f, 0.19992484152317047f));
    q = __ffma2_rn(q, z, make_float2(-0.3333306610584259f, -0.3333306610584259f));
    q = __ffma2_rn(q, z, make_float2(1.0f, 1.0f));
    const float2 r2 = __fmul2_rn(a, q);
    float r0 = r2.x, r1 = r2.y;
    r0 = ay0 > ax0 ? 1.57079632679489662f - r0 : r0;
    r1 = ay1 > ax1 ? 1.57079632679489662f - r1 : r1;
    r0 = (__float_as_uint(tr0) >> 31) ? 3.14159265358979324f - r0 : r0;
    r1 = (__float_as_uint(tr1) >> 31) ? 3.14159265358979324f - r1 : r1;
    return make_float2(__uint_as_float(__float_as_uint(r0) | (__float_as_uint(ti0) & 0x80000000u)),
                       __uint_as_float(__float_as_uint(r1) | (__float_as_uint(ti1) & 0x80000000u)));
}
#endif

}  // namespace cb
