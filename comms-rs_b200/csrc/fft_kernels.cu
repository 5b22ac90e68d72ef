// Batched FFT / IFFT kernels (K5) for sm_100a.
//
// Reference semantics (src/fft/mod.rs:73-96): X[k] = sum_n x[n] e^{-/+ j 2 pi k n / N},
// unnormalised in both directions.  The reference computes in f64 and rounds to f32;
// here the transform runs in f32 with twiddles rounded from f64 (rel-L2 error ~1e-7,
// tolerance 1e-4 per BASELINE north_star).
//
// Structure: Stockham autosort, radix 8 (+ one radix-4/2 pass), 8 points per thread
// held in registers, exchanges through padded shared memory.  The first pass reads
// global memory coalesced (x[j + r*N/8]) and the last pass writes it coalesced
// (X[j + q*N/R]), so a frame crosses HBM exactly once each way: 16 B per sample.
// Frames longer than 8192 use a four-step split N = N1*N2 with both steps done as
// 16-column batches (128-byte row segments) of the same butterfly core.
#include "fft_kernels.cuh"
#include "fft2_core.cuh"
#include "misc_kernels.cuh"

namespace cb {

// ---------------------------------------------------------------- butterflies
template <bool INV>
__device__ __forceinline__ float2 mul_mi(float2 a)  // a * (-i) forward, a * (+i) inverse
{
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

template <bool INV>
__device__ __forceinline__ void fft4(float2 &x0, float2 &x1, float2 &x2, float2 &x3)
{
    const float2 b0 = cadd(x0, x2), b2 = csub(x0, x2), b1 = cadd(x1, x3), b3 = mul_mi<INV>(csub(x1, x3));
    x0 = cadd(b0, b1);
    x2 = csub(b0, b1);
    x1 = cadd(b2, b3);
    x3 = csub(b2, b3);
}

// in-place size-8 DFT, natural order in and out
template <bool INV>
__device__ __forceinline__ void fft8(float2 *v)
{
    const float c = 0.70710678118654752440f;
    float2 a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
    float2 a1 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
    float2 a2 = cadd(v[2], v[6]), a6 = csub(v[2], v[6]);
    float2 a3 = cadd(v[3], v[7]), a7 = csub(v[3], v[7]);
    // a5 *= w8, a6 *= w8^2, a7 *= w8^3 with w8 = e^{-/+ j pi/4}
    if (INV) {
        a5 = make_float2(c * (a5.x - a5.y), c * (a5.x + a5.y));
        a7 = make_float2(-c * (a7.x + a7.y), c * (a7.x - a7.y));
    } else {
        a5 = make_float2(c * (a5.x + a5.y), c * (a5.y - a5.x));
        a7 = make_float2(c * (a7.y - a7.x), -c * (a7.x + a7.y));
    }
    a6 = mul_mi<INV>(a6);
    fft4<INV>(a0, a1, a2, a3);  // X[0], X[2], X[4], X[6]
    fft4<INV>(a4, a5, a6, a7);  // X[1], X[3], X[5], X[7]
    v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
    v[1] = a4; v[3] = a5; v[5] = a6; v[7] = a7;
}

// ---------------------------------------------------------------- FFT core
// One length-N transform by NT = N/8 cooperating threads (index j), 8 points each.
// SM(i) maps a logical point index to this transform's shared-memory slot.
// `v` holds x[j + r*NT] on entry; on exit v[q] holds the LAST pass's outputs:
//   REM == 0: X[j + q*NT]                      (q = 0..7)
//   REM == 2: X[jj + q*N/4], jj = j + h*NT     (stored at v[h + 2q])
//   REM == 1: X[jj + q*N/2], jj = j + h*NT     (stored at v[h + 4q])
// i.e. in every case v[r] = X[j + r*NT]: natural order with the same mapping as the input.
template <int LOG2N, bool INV, typename SM>
__device__ __forceinline__ void fft_core(float2 *v, const int j, const float2 *__restrict__ tw, SM sm)
{
    constexpr int N = 1 << LOG2N;
    constexpr int NT = N / 8;
    constexpr int P8 = LOG2N / 3;
    constexpr int REM = LOG2N % 3;
    int Ns = 1;
#pragma unroll
    for (int p = 0; p < P8; ++p) {
        if (p > 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r] = sm(j + r * NT);
            __syncthreads();
            const int idx = (j & (Ns - 1)) * (N / (Ns * 8));
#pragma unroll
            for (int q = 1; q < 8; ++q) v[q] = cmul(v[q], __ldg(tw + q * idx));
        }
        fft8<INV>(v);
        if (p == P8 - 1 && REM == 0) break;
        const int d = (j / Ns) * Ns * 8 + (j & (Ns - 1));
#pragma unroll
        for (int q = 0; q < 8; ++q) sm(d + q * Ns) = v[q];
        __syncthreads();
        Ns *= 8;
    }
    if (REM == 2) {  // final radix-4 pass: two butterflies per thread
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = sm(j + r * NT);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int jj = j + h * NT;  // Ns = N/4, idx = jj
            float2 x0 = v[h], x1 = cmul(v[h + 2], __ldg(tw + jj)), x2 = cmul(v[h + 4], __ldg(tw + 2 * jj)),
                   x3 = cmul(v[h + 6], __ldg(tw + 3 * jj));
            fft4<INV>(x0, x1, x2, x3);
            v[h] = x0; v[h + 2] = x1; v[h + 4] = x2; v[h + 6] = x3;
        }
    } else if (REM == 1) {  // final radix-2 pass: four butterflies per thread
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = sm(j + r * NT);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int jj = j + h * NT;  // Ns = N/2, idx = jj
            const float2 b = cmul(v[h + 4], __ldg(tw + jj));
            const float2 a = v[h];
            v[h] = cadd(a, b);
            v[h + 4] = csub(a, b);
        }
    }
}

// ---------------------------------------------------------------- one CTA, whole frames
struct SmPadded {
    float2 *base;
    __device__ __forceinline__ float2 &operator()(int i) const { return base[i + (i >> 4)]; }
};

template <int LOG2N, bool INV>
__global__ void __launch_bounds__((1 << LOG2N) / 8 >= 128 ? (1 << LOG2N) / 8 : 128)
fft_frames_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, const float2 *__restrict__ tw,
                  size_t nframes)
{
    constexpr int N = 1 << LOG2N;
    constexpr int NT = N / 8;
    constexpr int FPB = NT >= 128 ? 1 : 128 / NT;  // frames per block
    constexpr int PADN = N + N / 16 + 1;
    extern __shared__ __align__(16) float2 fsm[];
    const int f = threadIdx.x / NT, j = threadIdx.x % NT;
    const size_t frame = (size_t)blockIdx.x * FPB + f;
    const bool live = frame < nframes;
    const float2 *src = in + frame * N;
    float2 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = live ? src[j + r * NT] : make_float2(0.f, 0.f);
    fft_core<LOG2N, INV>(v, j, tw, SmPadded{fsm + f * PADN});
    if (live) {
        float2 *dst = out + frame * N;
#pragma unroll
        for (int r = 0; r < 8; ++r) dst[j + r * NT] = v[r];
    }
}

// ---------------------------------------------------------------- v2: radix-16 frames kernel
// One frame = T = N/16 threads, 16 points per thread (fft2_core.cuh); FPB frames per CTA.
// Passes exchange through ping-pong shared-memory buffers (one __syncthreads per exchange) or,
// when two buffers would not fit, in place with a barrier between a pass's loads and stores.
// First pass reads HBM and last pass writes HBM with lanes contiguous: 16 B per sample.
struct DevSm {
    float2 *base;
    __device__ __forceinline__ float2 ld(int i) const { return base[fft2::pad16(i)]; }
    __device__ __forceinline__ void st(int i, float2 v) const { base[fft2::pad16(i)] = v; }
};

template <int LOG2N>
struct Fft2Cfg {
    using PL = fft2::Plan<LOG2N>;
    static constexpr int THREADS = PL::T >= 256 ? PL::T : 256;
    static constexpr int FPB = THREADS / PL::T;
    // one in-place buffer + a barrier between the reads and the writes of a pass: 35 KiB per CTA, so that four
    // 64-register CTAs fit per SM (measured on 4096 points: 80 % -> 106 % of the copy-kernel roof vs ping-pong
    // buffers at three CTAs per SM)
    static constexpr bool PINGPONG = false;
    static constexpr int NBUF = PL::PASSES <= 1 ? 0 : (PINGPONG ? 2 : 1);
    static constexpr int SMEM = NBUF * FPB * PL::PADN * (int)sizeof(float2);
    static constexpr int MINB = THREADS >= 1024 ? 1 : (THREADS >= 512 ? 2 : 4);
};

template <int LOG2N, bool INV, int PASS, typename GLD, typename GST>
__device__ __forceinline__ void fft2_passes(int j, const float2 *tw, GLD gld, GST gst, float2 *b0, float2 *b1)
{
    using PL = fft2::Plan<LOG2N>;
    if constexpr (PASS < PL::PASSES) {
        if constexpr (PASS > 0) __syncthreads();
        DevSm sin{((PASS + 1) & 1) ? b1 : b0}, sout{(PASS & 1) ? b1 : b0};
        auto mid = [] { __syncthreads(); };
        auto nomid = [] {};
        if (b0 == b1) fft2::run_pass<LOG2N, INV, PASS>(j, tw, gld, gst, sin, sout, mid);
        else fft2::run_pass<LOG2N, INV, PASS>(j, tw, gld, gst, sin, sout, nomid);
        fft2_passes<LOG2N, INV, PASS + 1>(j, tw, gld, gst, b0, b1);
    }
}

// ROWS16: frame fr is row k1 = fr % 16 of the 16 x 4096 intermediate of a 65536-point transform
// (fft65536_stepA_kernel); output element k2 of the row is X[k1 + 16 k2] of big frame fr / 16.
template <int LOG2N, bool INV, bool ROWS16 = false>
__global__ void __launch_bounds__(Fft2Cfg<LOG2N>::THREADS, Fft2Cfg<LOG2N>::MINB)
fft2_frames_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, const float2 *__restrict__ tw,
                   size_t nframes)
{
    using PL = fft2::Plan<LOG2N>;
    using CF = Fft2Cfg<LOG2N>;
    extern __shared__ __align__(16) float2 fsm[];
    const int f = threadIdx.x / PL::T, j = threadIdx.x % PL::T;
    const size_t frame = (size_t)blockIdx.x * CF::FPB + f;
    const bool live = frame < nframes;
    const float2 *src = in + frame * PL::N;
    float2 *dst = ROWS16 ? out + (frame >> 4) * (PL::N * 16) + (frame & 15) : out + frame * PL::N;
    float2 *b0 = fsm + f * PL::PADN;
    float2 *b1 = CF::PINGPONG ? b0 + CF::FPB * PL::PADN : b0;
    auto gld = [&](int i) { return live ? ldg_stream2(src + i) : make_float2(0.f, 0.f); };
    auto gst = [&](int i, float2 v) {
        if (live) stg_stream2(dst + (ROWS16 ? 16 * i : i), v);
    };
    if constexpr (CF::PINGPONG) {
        fft2_passes<LOG2N, INV, 0>(j, tw, gld, gst, b0, b1);
    } else {
        fft2_passes<LOG2N, INV, 0>(j, tw, gld, gst, b0, b0);
    }
}

// i16 IQ input (src/io/raw_iq.rs:78-140): x = in_scale * (i16 as f32) folded into the first pass's loads -- half the
// input bytes and no widening pass (convert_i16_kernel's arithmetic, so the spectra equal convert -> transform bit for bit)
__device__ __forceinline__ float2 ldg_iq16(const uint32_t *p, float scale)
{
    uint32_t w;
    asm volatile("ld.global.cs.b32 %0, [%1];" : "=r"(w) : "l"(p));
    return make_float2(__fmul_rn(scale, (float)(int16_t)(w & 0xFFFFu)), __fmul_rn(scale, (float)(int16_t)(w >> 16)));
}

template <int LOG2N, bool INV>
__global__ void __launch_bounds__(Fft2Cfg<LOG2N>::THREADS, Fft2Cfg<LOG2N>::MINB)
fft2_frames_iq16_kernel(const uint32_t *__restrict__ in, float in_scale, float2 *__restrict__ out, const float2 *__restrict__ tw,
                        size_t nframes)
{
    using PL = fft2::Plan<LOG2N>;
    using CF = Fft2Cfg<LOG2N>;
    extern __shared__ __align__(16) float2 fsm[];
    const int f = threadIdx.x / PL::T, j = threadIdx.x % PL::T;
    const size_t frame = (size_t)blockIdx.x * CF::FPB + f;
    const bool live = frame < nframes;
    const uint32_t *src = in + frame * PL::N;
    float2 *dst = out + frame * PL::N;
    float2 *b0 = fsm + f * PL::PADN;
    float2 *b1 = CF::PINGPONG ? b0 + CF::FPB * PL::PADN : b0;
    auto gld = [&](int i) { return live ? ldg_iq16(src + i, in_scale) : make_float2(0.f, 0.f); };
    auto gst = [&](int i, float2 v) {
        if (live) stg_stream2(dst + i, v);
    };
    if constexpr (CF::PINGPONG) {
        fft2_passes<LOG2N, INV, 0>(j, tw, gld, gst, b0, b1);
    } else {
        fft2_passes<LOG2N, INV, 0>(j, tw, gld, gst, b0, b0);
    }
}

// ---------------------------------------------------------------- four-step column kernels
// Column batch: COLS adjacent columns of a (ROWS x ld) matrix, FFT along the rows
// index.  Thread t: column f = t % COLS, butterfly index j = t / COLS.
template <int COLS>
struct SmCols {
    float2 *base;
    int f;
    __device__ __forceinline__ float2 &operator()(int i) const { return base[i * (COLS + 1) + f]; }
};

// step 1 of N = N1*N2: for each n2, FFT over n1 of x[n1*N2 + n2], times
// twN[n2*k1], written TRANSPOSED: a[n2*N1 + k1]  (rows of k1 contiguous).
template <int LOG2N1, bool INV, int COLS>
__global__ void __launch_bounds__(((1 << LOG2N1) / 8) * COLS)
fft_step1_kernel(const float2 *__restrict__ in, float2 *__restrict__ scratch, const float2 *__restrict__ tw1,
                 const float2 *__restrict__ twN, int N2)
{
    constexpr int N1 = 1 << LOG2N1;
    constexpr int NT = N1 / 8;
    extern __shared__ __align__(16) float2 fsm[];
    const int f = threadIdx.x % COLS, j = threadIdx.x / COLS;
    const int colblocks = N2 / COLS;
    const size_t frame = blockIdx.x / colblocks;
    const int n2 = (blockIdx.x % colblocks) * COLS + f;
    const float2 *src = in + frame * (size_t)N1 * N2 + n2;
    float2 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = src[(size_t)(j + r * NT) * N2];
    SmCols<COLS> sm{fsm, f};
    fft_core<LOG2N1, INV>(v, j, tw1, sm);
    // twiddle by e^{-/+ 2 pi i n2 k1 / N}, k1 = j + r*NT, then stage for the transposed write
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int k1 = j + r * NT;
        sm(k1) = cmul(v[r], __ldg(twN + (size_t)n2 * k1));
    }
    __syncthreads();
    // write: for each column f' a contiguous run of N1 values (lanes along k1)
    float2 *dst = scratch + frame * (size_t)N1 * N2 + (size_t)((blockIdx.x % colblocks) * COLS) * N1;
    for (int e = threadIdx.x; e < N1 * COLS; e += NT * COLS) {
        const int ff = e / N1, k1 = e % N1;
        dst[(size_t)ff * N1 + k1] = fsm[k1 * (COLS + 1) + ff];
    }
}

// step 2: for each k1, FFT over n2 of a[n2*N1 + k1]; X[k2*N1 + k1].
template <int LOG2N2, bool INV, int COLS>
__global__ void __launch_bounds__(((1 << LOG2N2) / 8) * COLS)
fft_step2_kernel(const float2 *__restrict__ scratch, float2 *__restrict__ out, const float2 *__restrict__ tw2, int N1)
{
    constexpr int N2 = 1 << LOG2N2;
    constexpr int NT = N2 / 8;
    extern __shared__ __align__(16) float2 fsm[];
    const int f = threadIdx.x % COLS, j = threadIdx.x / COLS;
    const int colblocks = N1 / COLS;
    const size_t frame = blockIdx.x / colblocks;
    const int k1 = (blockIdx.x % colblocks) * COLS + f;
    const float2 *src = scratch + frame * (size_t)N1 * N2 + k1;
    float2 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = src[(size_t)(j + r * NT) * N1];
    fft_core<LOG2N2, INV>(v, j, tw2, SmCols<COLS>{fsm, f});
    float2 *dst = out + frame * (size_t)N1 * N2 + k1;
#pragma unroll
    for (int r = 0; r < 8; ++r) dst[(size_t)(j + r * NT) * N1] = v[r];
}

// ---------------------------------------------------------------- any-N direct DFT
// O(N^2) with an exact table: tw[(k*n) mod N].  Used for non power-of-two and tiny N
// (the reference accepts any size; its own golden test is N = 10).
__global__ void __launch_bounds__(128)
dft_direct_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, const float2 *__restrict__ tw, int N,
                  size_t nframes)
{
    extern __shared__ __align__(16) float2 fsm[];  // one frame
    for (size_t frame = blockIdx.x; frame < nframes; frame += gridDim.x) {
        const float2 *src = in + frame * N;
        for (int i = threadIdx.x; i < N; i += blockDim.x) fsm[i] = src[i];
        __syncthreads();
        for (int k = threadIdx.x; k < N; k += blockDim.x) {
            float ar = 0.f, ai = 0.f, cr = 0.f, ci = 0.f;  // Kahan-compensated: long sums in f32
            int m = 0;
            for (int n = 0; n < N; ++n) {
                const float2 w = __ldg(tw + m);
                const float2 p = cmul(fsm[n], w);
                float y = p.x - cr, t = ar + y;
                cr = (t - ar) - y; ar = t;
                y = p.y - ci; t = ai + y;
                ci = (t - ai) - y; ai = t;
                m += k;
                if (m >= N) m -= N;
            }
            out[frame * N + k] = make_float2(ar, ai);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- any-N chirp-z (Bluestein) glue
// X[k] = w[k] sum_n (x[n] w[n]) conj(w[k - n]),  w[n] = e^{-/+ j pi n^2 / N}: a length-N transform as a circular
// convolution of length M = 2^m >= 2N - 1, done with the power-of-two kernels above.  The three element-wise steps:
//   pre:  a[f][i] = x[f][i] w[i] for i < N, 0 for N <= i < M
//   mul:  A[f][i] *= B[i]                (B = M-point transform of the wrapped conj chirp, computed on the host in f64)
//   post: X[f][k] = c[f][k] w[k] / M
// One CTA = 1024 consecutive elements of one frame (256 threads x 4, lanes contiguous).
__global__ void __launch_bounds__(256)
bluestein_pre_kernel(const float2 *__restrict__ x, const float2 *__restrict__ chirp, float2 *__restrict__ a, unsigned N,
                     unsigned M, unsigned nseg)
{
    const size_t frame = blockIdx.x / nseg;
    const unsigned i0 = (blockIdx.x % nseg) * 1024 + threadIdx.x;
    const float2 *src = x + frame * N;
    float2 *dst = a + frame * M;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const unsigned i = i0 + 256 * u;
        if (i < M) dst[i] = i < N ? cmul(ldg_stream2(src + i), __ldg(chirp + i)) : make_float2(0.f, 0.f);
    }
}

__global__ void __launch_bounds__(256)
bluestein_mul_kernel(float2 *__restrict__ spec, const float2 *__restrict__ bspec, unsigned M, size_t total)
{
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += stride)
        spec[i] = cmul(spec[i], __ldg(bspec + (i & (M - 1))));
}

__global__ void __launch_bounds__(256)
bluestein_post_kernel(const float2 *__restrict__ c, const float2 *__restrict__ chirp, float2 *__restrict__ out, unsigned N,
                      unsigned M, unsigned nseg, float scale)
{
    const size_t frame = blockIdx.x / nseg;
    const unsigned i0 = (blockIdx.x % nseg) * 1024 + threadIdx.x;
    const float2 *src = c + frame * M;
    float2 *dst = out + frame * N;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const unsigned i = i0 + 256 * u;
        if (i < N) {
            const float2 v = cmul(src[i], __ldg(chirp + i));
            stg_stream2(dst + i, make_float2(v.x * scale, v.y * scale));
        }
    }
}

// Chirp-z with M <= 16384: the element-wise steps folded into the first-pass loads and last-pass stores of the two
// M-point transforms (fft2_frames_kernel's passes): stage 0 = chirp, zero fill, forward transform, product with B;
// stage 1 = inverse transform, chirp, 1/M, first N outputs.  The spectra make one round trip through a scratch that the
// caller keeps L2-sized, so HBM carries 8 N bytes in and 8 N bytes out per frame.
struct BlArgs {
    const float2 *x, *chirp, *bspec;
    float2 *spec, *out;
    unsigned N;
    size_t nframes;
    float scale;
};

template <int LOG2M, int STAGE>
__global__ void __launch_bounds__(Fft2Cfg<LOG2M>::THREADS, Fft2Cfg<LOG2M>::MINB)
bluestein_frames_kernel(const __grid_constant__ BlArgs a, const float2 *__restrict__ tw)
{
    using PL = fft2::Plan<LOG2M>;
    using CF = Fft2Cfg<LOG2M>;
    extern __shared__ __align__(16) float2 fsm[];
    const int f = threadIdx.x / PL::T, j = threadIdx.x % PL::T;
    const size_t frame = (size_t)blockIdx.x * CF::FPB + f;
    const bool live = frame < a.nframes;
    float2 *b0 = fsm + f * PL::PADN;
    if constexpr (STAGE == 0) {
        const float2 *src = a.x + frame * a.N;
        float2 *dst = a.spec + frame * PL::N;
        auto gld = [&](int i) {
            return (live && i < (int)a.N) ? cmul(ldg_stream2(src + i), __ldg(a.chirp + i)) : make_float2(0.f, 0.f);
        };
        auto gst = [&](int i, float2 v) {
            if (live) dst[i] = cmul(v, __ldg(a.bspec + i));
        };
        fft2_passes<LOG2M, false, 0>(j, tw, gld, gst, b0, b0);
    } else {
        const float2 *src = a.spec + frame * PL::N;
        float2 *dst = a.out + frame * a.N;
        auto gld = [&](int i) { return live ? src[i] : make_float2(0.f, 0.f); };
        auto gst = [&](int i, float2 v) {
            if (live && i < (int)a.N) {
                const float2 r = cmul(v, __ldg(a.chirp + i));
                stg_stream2(dst + i, make_float2(r.x * a.scale, r.y * a.scale));
            }
        };
        fft2_passes<LOG2M, true, 0>(j, tw, gld, gst, b0, b0);
    }
}

template <int LOG2M>
static int launch_bluestein_fused_m(const BlArgs &a, const float2 *tw_fwd, const float2 *tw_inv, cudaStream_t s)
{
    using CF = Fft2Cfg<LOG2M>;
    auto k0 = bluestein_frames_kernel<LOG2M, 0>;
    auto k1 = bluestein_frames_kernel<LOG2M, 1>;
    CB_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM));
    CB_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM));
    const unsigned grid = (unsigned)ceil_div(a.nframes, (size_t)CF::FPB);
    k0<<<grid, CF::THREADS, CF::SMEM, s>>>(a, tw_fwd);
    count_launch();
    k1<<<grid, CF::THREADS, CF::SMEM, s>>>(a, tw_inv);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

int launch_bluestein_fused(const float2 *x, const float2 *chirp, const float2 *bspec, float2 *spec, float2 *out, uint32_t N,
                           int log2m, const float2 *tw_fwd, const float2 *tw_inv, size_t frames, cudaStream_t s)
{
    BlArgs a{x, chirp, bspec, spec, out, N, frames, 1.0f / (float)(1u << log2m)};
    switch (log2m) {
    case 8: return launch_bluestein_fused_m<8>(a, tw_fwd, tw_inv, s);
    case 9: return launch_bluestein_fused_m<9>(a, tw_fwd, tw_inv, s);
    case 10: return launch_bluestein_fused_m<10>(a, tw_fwd, tw_inv, s);
    case 11: return launch_bluestein_fused_m<11>(a, tw_fwd, tw_inv, s);
    case 12: return launch_bluestein_fused_m<12>(a, tw_fwd, tw_inv, s);
    case 13: return launch_bluestein_fused_m<13>(a, tw_fwd, tw_inv, s);
    case 14: return launch_bluestein_fused_m<14>(a, tw_fwd, tw_inv, s);
    default: set_error("fft: no fused chirp-z kernel for 2^%d", log2m); return CB_ERR_UNSUPPORTED;
    }
}

int launch_bluestein_pre(const float2 *x, const float2 *chirp, float2 *a, uint32_t N, uint32_t M, size_t frames, cudaStream_t s)
{
    const unsigned nseg = (M + 1023) / 1024;
    bluestein_pre_kernel<<<(unsigned)(frames * nseg), 256, 0, s>>>(x, chirp, a, N, M, nseg);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

int launch_bluestein_mul(float2 *spec, const float2 *bspec, uint32_t M, size_t frames, cudaStream_t s)
{
    const size_t total = frames * M;
    const size_t blocks = ceil_div(total, (size_t)256);
    bluestein_mul_kernel<<<(unsigned)(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, s>>>(spec, bspec, M, total);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

int launch_bluestein_post(const float2 *c, const float2 *chirp, float2 *out, uint32_t N, uint32_t M, size_t frames,
                          cudaStream_t s)
{
    const unsigned nseg = (N + 1023) / 1024;
    bluestein_post_kernel<<<(unsigned)(frames * nseg), 256, 0, s>>>(c, chirp, out, N, M, nseg, 1.0f / (float)M);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// ---------------------------------------------------------------- launchers
template <int LOG2N, bool INV>
static int launch_frames(const float2 *in, float2 *out, const float2 *tw, size_t nframes, cudaStream_t s)
{
    constexpr int N = 1 << LOG2N;
    constexpr int NT = N / 8;
    constexpr int FPB = NT >= 128 ? 1 : 128 / NT;
    constexpr int THREADS = NT * FPB;
    constexpr int SMEM = FPB * (N + N / 16 + 1) * (int)sizeof(float2);
    auto kern = fft_frames_kernel<LOG2N, INV>;
        CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    kern<<<(unsigned)ceil_div(nframes, (size_t)FPB), THREADS, SMEM, s>>>(in, out, tw, nframes);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

template <int LOG2N, bool INV, bool IN16>
static int launch_small_frames(const void *in, float in_scale, float2 *out, const float2 *tw2, size_t nframes, cudaStream_t s);
static bool small_frames_enabled(int log2n);

template <int LOG2N, bool INV>
static int launch_frames2(const float2 *in, float2 *out, const float2 *tw2, size_t nframes, cudaStream_t s)
{
    if constexpr (LOG2N <= 7)
        if (small_frames_enabled(LOG2N)) return launch_small_frames<LOG2N, INV, false>(in, 1.f, out, tw2, nframes, s);
    using CF = Fft2Cfg<LOG2N>;
    auto kern = fft2_frames_kernel<LOG2N, INV>;
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM));
    kern<<<(unsigned)ceil_div(nframes, (size_t)CF::FPB), CF::THREADS, CF::SMEM, s>>>(in, out, tw2, nframes);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

template <bool INV>
static int launch_frames2_dir(int log2n, const float2 *in, float2 *out, const float2 *tw2, size_t nframes, cudaStream_t s)
{
    switch (log2n) {
    case 4: return launch_frames2<4, INV>(in, out, tw2, nframes, s);
    case 5: return launch_frames2<5, INV>(in, out, tw2, nframes, s);
    case 6: return launch_frames2<6, INV>(in, out, tw2, nframes, s);
    case 7: return launch_frames2<7, INV>(in, out, tw2, nframes, s);
    case 8: return launch_frames2<8, INV>(in, out, tw2, nframes, s);
    case 9: return launch_frames2<9, INV>(in, out, tw2, nframes, s);
    case 10: return launch_frames2<10, INV>(in, out, tw2, nframes, s);
    case 11: return launch_frames2<11, INV>(in, out, tw2, nframes, s);
    case 12: return launch_frames2<12, INV>(in, out, tw2, nframes, s);
    case 13: return launch_frames2<13, INV>(in, out, tw2, nframes, s);
    case 14: return launch_frames2<14, INV>(in, out, tw2, nframes, s);
    default: set_error("fft: no v2 kernel for 2^%d", log2n); return CB_ERR_UNSUPPORTED;
    }
}

// 16 .. 128 points: the generic kernel gives every thread 16 points of ONE frame, so a warp's accesses are 8 (N = 16)
// to 64 bytes per lane, a frame apart -- 27 % of the HBM roof at 16 points, 76 % at 64.  Here a CTA moves 4096 contiguous points
// (256 / 64 / ... frames) with coalesced 16-byte accesses through a shared-memory image in the passes' padded layout
// and transforms them in place there (a thread only ever overwrites points of its own frame; warp barriers separate
// the reads and the writes of a pass when a frame has more than one thread -- a warp owns 512 contiguous points from
// the first load to the last store, so the CTA never synchronises).
template <int LOG2N, bool INV, bool IN16>
__global__ void __launch_bounds__(256, 4)
fft2_small_frames_kernel(const void *__restrict__ in_, float in_scale, float2 *__restrict__ out, const float2 *__restrict__ tw,
                         size_t nframes)
{
    using PL = fft2::Plan<LOG2N>;
    constexpr int N = PL::N, T = PL::T, PADN = PL::PADN, PTS = 4096;
    static_assert(LOG2N >= 4 && LOG2N <= 7 && PL::P16 == 1, "one radix-16 pass plus at most a remainder pass");
    extern __shared__ __align__(16) float2 fsm[];
    const int tid = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * PTS, total = nframes * (size_t)N;
    const int live = (int)(total - base < (size_t)PTS ? total - base : (size_t)PTS);  // whole frames: a multiple of N
    const float2 zero = make_float2(0.f, 0.f);
    // every warp stages, transforms and writes back its own 512 contiguous points (32 / T whole frames): warp barriers
    // only, the eight warps of the CTA never wait for each other
    const int w0 = (tid >> 5) * 512, lane = tid & 31;
    if constexpr (IN16) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(in_) + base;
        for (int i = w0 + lane; i < w0 + 512; i += 32)
            fsm[(i >> LOG2N) * PADN + fft2::pad16(i & (N - 1))] = i < live ? ldg_iq16(src + i, in_scale) : zero;
    } else {
        const float2 *src = reinterpret_cast<const float2 *>(in_) + base;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {  // pairs never straddle a frame (N is even)
                const int i = w0 + 2 * lane + 64 * k;
                const float4 v = i < live ? ldg_stream(reinterpret_cast<const float4 *>(src + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                float2 *fr = fsm + (i >> LOG2N) * PADN;
                fr[fft2::pad16(i & (N - 1))] = make_float2(v.x, v.y);
                fr[fft2::pad16((i & (N - 1)) + 1)] = make_float2(v.z, v.w);
            }
        } else {
            for (int i = w0 + lane; i < w0 + 512; i += 32)
                fsm[(i >> LOG2N) * PADN + fft2::pad16(i & (N - 1))] = i < live ? ldg_stream2(src + i) : zero;
        }
    }
    __syncwarp();
    {
        const int j = tid % T;
        float2 *S = fsm + (tid / T) * PADN;
        auto ld = [&](int m) { return S[fft2::pad16(j + m * T)]; };
        auto st = [&](int idx, float2 v) { S[fft2::pad16(idx)] = v; };
        auto bar = [] { __syncwarp(); };  // a frame's T <= 8 threads sit in one warp
        auto nobar = [] {};
        if constexpr (T == 1) {
            fft2::pass16<LOG2N, INV, 0>(j, tw, ld, nobar, st);  // the thread owns its frame
        } else {
            fft2::pass16<LOG2N, INV, 0>(j, tw, ld, bar, st);
            __syncwarp();
            fft2::pass_rem<LOG2N, INV>(j, tw, ld, bar, st);
        }
    }
    __syncwarp();
    float2 *dst = out + base;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = w0 + 2 * lane + 64 * k;
            if (i < live) {
                const float2 *fr = fsm + (i >> LOG2N) * PADN;
                const float2 a = fr[fft2::pad16(i & (N - 1))], b = fr[fft2::pad16((i & (N - 1)) + 1)];
                stg_stream(reinterpret_cast<float4 *>(dst + i), make_float4(a.x, a.y, b.x, b.y));
            }
        }
    } else {
        for (int i = w0 + lane; i < w0 + 512 && i < live; i += 32)
            stg_stream2(dst + i, fsm[(i >> LOG2N) * PADN + fft2::pad16(i & (N - 1))]);
    }
}

template <int LOG2N, bool INV, bool IN16>
static int launch_small_frames(const void *in, float in_scale, float2 *out, const float2 *tw2, size_t nframes, cudaStream_t s)
{
    using PL = fft2::Plan<LOG2N>;
    constexpr int SMEM = (4096 / PL::N) * PL::PADN * (int)sizeof(float2);
    auto kern = fft2_small_frames_kernel<LOG2N, INV, IN16>;
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    kern<<<(unsigned)ceil_div(nframes * (size_t)PL::N, (size_t)4096), 256, SMEM, s>>>(in, in_scale, out, tw2, nframes);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// Measured on 2^28 samples (% of the 16 B/sample roof, this kernel / the generic one): 16 points 106 / 27, 32 points
// 107 / 51, 64 points 106 / 76, 128 points 106 / 107.  (A first form with block barriers around the staging reached
// only 83 / 68 / 69 / 69: what matters is that the warps never wait for each other.)
// COMMS_B200_FFT_SMALL = generic: the generic frames kernel for these sizes too (for comparison).
static bool small_frames_enabled(int log2n)
{
    const char *e = getenv("COMMS_B200_FFT_SMALL");  // read per launch: a test switches it within one process
    return !(e && strcmp(e, "generic") == 0) && log2n <= 7;
}

template <int LOG2N, bool INV>
static int launch_frames2_iq16(const uint32_t *in, float in_scale, float2 *out, const float2 *tw2, size_t nframes, cudaStream_t s)
{
    if constexpr (LOG2N <= 7)
        if (small_frames_enabled(LOG2N)) return launch_small_frames<LOG2N, INV, true>(in, in_scale, out, tw2, nframes, s);
    using CF = Fft2Cfg<LOG2N>;
    auto kern = fft2_frames_iq16_kernel<LOG2N, INV>;
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM));
    kern<<<(unsigned)ceil_div(nframes, (size_t)CF::FPB), CF::THREADS, CF::SMEM, s>>>(in, in_scale, out, tw2, nframes);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

template <bool INV>
static int launch_frames2_iq16_dir(int log2n, const uint32_t *in, float sc, float2 *out, const float2 *tw2, size_t nframes,
                                   cudaStream_t s)
{
    switch (log2n) {
    case 4: return launch_frames2_iq16<4, INV>(in, sc, out, tw2, nframes, s);
    case 5: return launch_frames2_iq16<5, INV>(in, sc, out, tw2, nframes, s);
    case 6: return launch_frames2_iq16<6, INV>(in, sc, out, tw2, nframes, s);
    case 7: return launch_frames2_iq16<7, INV>(in, sc, out, tw2, nframes, s);
    case 8: return launch_frames2_iq16<8, INV>(in, sc, out, tw2, nframes, s);
    case 9: return launch_frames2_iq16<9, INV>(in, sc, out, tw2, nframes, s);
    case 10: return launch_frames2_iq16<10, INV>(in, sc, out, tw2, nframes, s);
    case 11: return launch_frames2_iq16<11, INV>(in, sc, out, tw2, nframes, s);
    case 12: return launch_frames2_iq16<12, INV>(in, sc, out, tw2, nframes, s);
    case 13: return launch_frames2_iq16<13, INV>(in, sc, out, tw2, nframes, s);
    case 14: return launch_frames2_iq16<14, INV>(in, sc, out, tw2, nframes, s);
    default: set_error("fft: no v2 kernel for 2^%d", log2n); return CB_ERR_UNSUPPORTED;
    }
}

// host side of fft2::Plan<LOG2N>::tw_off / TW_TOTAL (runtime log2n)
size_t fft2_table_len(int log2n)
{
    const int p16 = log2n / 4, rem = log2n % 4;
    size_t off = 0;
    for (int q = 1; q < p16; ++q) off += (size_t)1 << (4 * q);
    return off + (rem ? ((size_t)1 << log2n) >> rem : 0) + 1;
}

void fft2_fill_table(int log2n, int inverse, float2 *t)
{
    const int p16 = log2n / 4, rem = log2n % 4;
    const long double sgn = inverse ? 2.0L : -2.0L, pi = 3.14159265358979323846264338327950288L;
    size_t off = 0;
    for (int p = 1; p < p16; ++p) {
        const size_t ns = (size_t)1 << (4 * p);
        for (size_t s = 0; s < ns; ++s) {
            const long double a = sgn * pi * (long double)s / (16.0L * (long double)ns);
            t[off + s] = make_float2((float)cosl(a), (float)sinl(a));
        }
        off += ns;
    }
    if (rem) {
        const size_t n = (size_t)1 << log2n, m = n >> rem;
        for (size_t s = 0; s < m; ++s) {
            const long double a = sgn * pi * (long double)s / (long double)n;
            t[off + s] = make_float2((float)cosl(a), (float)sinl(a));
        }
        off += m;
    }
    t[off] = make_float2(1.f, 0.f);
}

template <bool INV>
static int launch_frames_dir(int log2n, const float2 *in, float2 *out, const float2 *tw, size_t nframes, cudaStream_t s)
{
    switch (log2n) {
    case 3: return launch_frames<3, INV>(in, out, tw, nframes, s);
    case 4: return launch_frames<4, INV>(in, out, tw, nframes, s);
    case 5: return launch_frames<5, INV>(in, out, tw, nframes, s);
    case 6: return launch_frames<6, INV>(in, out, tw, nframes, s);
    case 7: return launch_frames<7, INV>(in, out, tw, nframes, s);
    case 8: return launch_frames<8, INV>(in, out, tw, nframes, s);
    case 9: return launch_frames<9, INV>(in, out, tw, nframes, s);
    case 10: return launch_frames<10, INV>(in, out, tw, nframes, s);
    case 11: return launch_frames<11, INV>(in, out, tw, nframes, s);
    case 12: return launch_frames<12, INV>(in, out, tw, nframes, s);
    case 13: return launch_frames<13, INV>(in, out, tw, nframes, s);
    default: set_error("fft: no single-CTA kernel for 2^%d", log2n); return CB_ERR_UNSUPPORTED;
    }
}

template <int LOG2M, bool INV, int COLS>
static int launch_step(int which, const float2 *in, float2 *out, const float2 *twM, const float2 *twN, int other,
                       size_t nframes, cudaStream_t s)
{
    constexpr int M = 1 << LOG2M;
    constexpr int THREADS = (M / 8) * COLS;
    constexpr int SMEM = M * (COLS + 1) * (int)sizeof(float2);
    const unsigned grid = (unsigned)(nframes * (size_t)(other / COLS));
    if (which == 1) {
        auto kern = fft_step1_kernel<LOG2M, INV, COLS>;
            CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        kern<<<grid, THREADS, SMEM, s>>>(in, out, twM, twN, other);
    } else {
        auto kern = fft_step2_kernel<LOG2M, INV, COLS>;
            CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        kern<<<grid, THREADS, SMEM, s>>>(in, out, twM, other);
    }
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

template <bool INV>
static int launch_step_dir(int which, int log2m, const float2 *in, float2 *out, const float2 *twM, const float2 *twN,
                           int other, size_t nframes, cudaStream_t s)
{
    switch (log2m) {
    case 7: return launch_step<7, INV, 16>(which, in, out, twM, twN, other, nframes, s);
    case 8: return launch_step<8, INV, 16>(which, in, out, twM, twN, other, nframes, s);
    case 9: return launch_step<9, INV, 16>(which, in, out, twM, twN, other, nframes, s);
    case 10: return launch_step<10, INV, 8>(which, in, out, twM, twN, other, nframes, s);
    default: set_error("fft: no four-step kernel for 2^%d", log2m); return CB_ERR_UNSUPPORTED;
    }
}

// ---------------------------------------------------------------- overlap-save FIR for long filters
// y = h * x for 129 .. 1025 taps by fast convolution (src/filter/fir.rs:87-102 semantics, state carried):
//   frame f = input samples [f hop - (K-1), f hop - (K-1) + 4096),  hop = 4096 - (K-1)
//   stage 0: forward 4096-point FFT of every frame (history / zero fill handled in the loads)      -> spec
//   stage 1: spec * Hf (FFT of the zero-padded taps, evaluated in f64 on the host), inverse FFT, scale by 1/4096,
//            keep outputs K-1 .. 4095 of the frame                                                  -> y
// Both stages are fft2_frames_kernel<12>'s passes with other first-pass loads / last-pass stores.  ~40 B of HBM
// traffic per sample instead of 16, but O(log N) work per sample where the direct form needs K MACs.
struct OlsArgs {
    const float2 *x;
    const float2 *hist;   // hist_len samples preceding x[0], chronological
    float2 *spec;         // frames x 4096
    const float2 *hf;     // 4096
    float2 *y;
    unsigned long long n;
    unsigned hist_len, K, hop;
};

template <int STAGE>
__global__ void __launch_bounds__(Fft2Cfg<12>::THREADS, Fft2Cfg<12>::MINB)
fir_ols_kernel(const __grid_constant__ OlsArgs a, const float2 *__restrict__ tw, const float2 *__restrict__ tw2)
{
    using PL = fft2::Plan<12>;
    extern __shared__ __align__(16) float2 fsm[];
    const int j = threadIdx.x;
    const long long frame = blockIdx.x;
    float2 *b0 = fsm;
    if constexpr (STAGE == 0) {
        const long long g0 = frame * a.hop - (long long)(a.K - 1);
        auto gld = [&](int i) {
            const long long g = g0 + i;
            if (g >= (long long)a.n) return make_float2(0.f, 0.f);
            if (g >= 0) return ldg_stream2(a.x + g);
            const long long hi = (long long)a.hist_len + g;
            return hi >= 0 ? a.hist[hi] : make_float2(0.f, 0.f);
        };
        float2 *dst = a.spec + frame * PL::N;
        auto gst = [&](int i, float2 v) { dst[i] = v; };
        fft2_passes<12, false, 0>(j, tw, gld, gst, b0, b0);
    } else if constexpr (STAGE == 2) {
        // fused: forward transform, product with Hf and inverse transform without leaving shared memory
        // (tw = forward tables, tw2 = inverse tables); HBM traffic 8 * 4096 / hop + 8 bytes per sample
        constexpr int T = PL::T;
        const long long g0 = frame * a.hop - (long long)(a.K - 1);
        auto gld = [&](int i) {
            const long long g = g0 + i;
            if (g >= (long long)a.n) return make_float2(0.f, 0.f);
            if (g >= 0) return ldg_stream2(a.x + g);
            const long long hi = (long long)a.hist_len + g;
            return hi >= 0 ? a.hist[hi] : make_float2(0.f, 0.f);
        };
        auto gst = [&](int i, float2 v) {
            const long long o = g0 + i;
            if (i >= (int)(a.K - 1) && o < (long long)a.n) stg_stream2(a.y + o, make_float2(v.x * (1.f / 4096.f), v.y * (1.f / 4096.f)));
        };
        DevSm sm{b0};
        auto nomid = [] {};
        auto bar = [] { __syncthreads(); };
        auto nogst = [](int, float2) {};
        auto st_s = [&](int idx, float2 v) { sm.st(idx, v); };
        fft2::run_pass<12, false, 0>(j, tw, gld, nogst, sm, sm, nomid);
        __syncthreads();
        fft2::run_pass<12, false, 1>(j, tw, gld, nogst, sm, sm, bar);
        __syncthreads();
        auto ld_s = [&](int m) { return sm.ld(j + m * T); };
        // the spectrum stays in place, multiplied by Hf on the way back into shared memory
        fft2::pass16<12, false, 2>(j, tw, ld_s, bar, [&](int idx, float2 v) { sm.st(idx, fft2::cmul(v, __ldg(a.hf + idx))); });
        __syncthreads();
        fft2::pass16<12, true, 0>(j, tw2, ld_s, bar, st_s);
        __syncthreads();
        fft2::run_pass<12, true, 1>(j, tw2, gld, nogst, sm, sm, bar);
        __syncthreads();
        fft2::run_pass<12, true, 2>(j, tw2, gld, gst, sm, sm, nomid);
    } else {
        const float2 *src = a.spec + frame * PL::N;
        auto gld = [&](int i) { return fft2::cmul(src[i], __ldg(a.hf + i)); };
        const long long o0 = frame * a.hop - (long long)(a.K - 1);
        auto gst = [&](int i, float2 v) {
            const long long o = o0 + i;
            if (i >= (int)(a.K - 1) && o < (long long)a.n) stg_stream2(a.y + o, make_float2(v.x * (1.f / 4096.f), v.y * (1.f / 4096.f)));
        };
        fft2_passes<12, true, 0>(j, tw, gld, gst, b0, b0);
    }
}

__global__ void __launch_bounds__(256)
fir_hist_update_kernel(const float2 *__restrict__ x, unsigned long long n, const float2 *__restrict__ hist_in,
                       float2 *__restrict__ hist_out, unsigned H)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < H; i += gridDim.x * blockDim.x) {
        const long long g = (long long)n - H + i;
        hist_out[i] = g >= 0 ? x[g] : hist_in[H + g];
    }
}

size_t fir_ols_frames(size_t n, uint32_t ntaps) { return ceil_div(n, (size_t)(4096 - (ntaps - 1))); }

int launch_fir_ols(const float2 *x, size_t n, const float2 *hist_in, float2 *hist_out, uint32_t hist_len, uint32_t ntaps,
                   const float2 *hf, const float2 *tw_fwd, const float2 *tw_inv, float2 *spec, float2 *y, cudaStream_t s)
{
    using CF = Fft2Cfg<12>;
    if (n == 0) return CB_OK;
    OlsArgs a;
    a.x = x;
    a.hist = hist_in;
    a.spec = spec;
    a.hf = hf;
    a.y = y;
    a.n = n;
    a.hist_len = hist_len;
    a.K = ntaps;
    a.hop = 4096 - (ntaps - 1);
    const unsigned frames = (unsigned)fir_ols_frames(n, ntaps);
    if (spec == nullptr) {
        auto kf = fir_ols_kernel<2>;
        CB_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM));
        kf<<<frames, CF::THREADS, CF::SMEM, s>>>(a, tw_fwd, tw_inv);
        count_launch();
    } else {
        auto k0 = fir_ols_kernel<0>;
        auto k1 = fir_ols_kernel<1>;
        CB_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM));
        CB_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM));
        k0<<<frames, CF::THREADS, CF::SMEM, s>>>(a, tw_fwd, tw_inv);
        count_launch();
        k1<<<frames, CF::THREADS, CF::SMEM, s>>>(a, tw_inv, tw_fwd);
        count_launch();
    }
    if (hist_out != nullptr) {
        fir_hist_update_kernel<<<(hist_len + 255) / 256, 256, 0, s>>>(x, n, hist_in, hist_out, hist_len);
        count_launch();
    }
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// ---------------------------------------------------------------- 65536 = 16 x 4096, two streaming passes
// step A: for every n2, the 16-point DFT over n1 of x[4096 n1 + n2], times W_N^{n2 k1}, stored as row k1
// (no shared memory: 16 coalesced loads, one radix-16 butterfly, 16 coalesced stores per thread);
// step B: fft2_frames_kernel<12, INV, true> on the 16 rows, scattering X[k1 + 16 k2].
template <bool INV>
__global__ void __launch_bounds__(256)
fft65536_stepA_kernel(const float2 *__restrict__ in, float2 *__restrict__ mid, const float2 *__restrict__ twN, size_t nframes)
{
    const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t frame = idx >> 12;
    const int n2 = (int)(idx & 4095);
    if (frame >= nframes) return;
    const float2 *src = in + frame * 65536 + n2;
    float2 *dst = mid + frame * 65536 + n2;
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = ldg_stream2(src + 4096 * m);
    fft2::bfly16<INV>(v);  // slot sl holds k1 = fft2::q16(sl)
    float2 p[16];    // p[k] = W_N^{n2 k}
    p[0] = make_float2(1.f, 0.f);
    p[1] = __ldg(twN + n2);
    p[2] = fft2::csqr(p[1]);
    p[3] = fft2::cmul(p[2], p[1]);
    p[4] = fft2::csqr(p[2]);
    p[5] = fft2::cmul(p[4], p[1]);
    p[6] = fft2::csqr(p[3]);
    p[7] = fft2::cmul(p[4], p[3]);
    p[8] = fft2::csqr(p[4]);
#pragma unroll
    for (int k = 9; k < 16; ++k) p[k] = (k & 1) ? fft2::cmul(p[8], p[k - 8]) : fft2::csqr(p[k / 2]);
#pragma unroll
    for (int sl = 0; sl < 16; ++sl) {
        const int k1 = fft2::q16(sl);
        stg_stream2(dst + 4096 * k1, k1 ? fft2::cmul(v[sl], p[k1]) : v[sl]);
    }
}

// step B, four rows per CTA: rows k1 = 4a .. 4a+3 of one big frame are transformed side by side (256 threads
// each, passes 0 and 1 as in fft2_frames_kernel); for the last pass the 1024 threads are re-partitioned so that
// four consecutive lanes hold the same output index k2 of the four rows: X[4a + r + 16 k2], r = 0..3, is one
// full 32-byte sector per quad of lanes instead of four 8-byte pieces 128 bytes apart.
template <bool INV>
__global__ void __launch_bounds__(1024, 1)
fft65536_stepB_kernel(const float2 *__restrict__ mid, float2 *__restrict__ out, const float2 *__restrict__ tw, size_t nquads)
{
    using PL = fft2::Plan<12>;
    constexpr int RPITCH = PL::PADN + 8;  // + 64 bytes: the four rows of a lane quad fall into different banks
    extern __shared__ __align__(16) float2 fsm[];
    const size_t quad = blockIdx.x;
    if (quad >= nquads) return;
    const size_t F = quad >> 2;
    const int a4 = (int)(quad & 3) * 4;
    const int f = threadIdx.x >> 8, j = threadIdx.x & 255;
    const float2 *src = mid + (F * 16 + a4 + f) * PL::N;
    DevSm sm{fsm + f * RPITCH};
    auto gld = [&](int i) { return ldg_stream2(src + i); };
    auto nogst = [](int, float2) {};
    auto nomid = [] {};
    auto mid_bar = [] { __syncthreads(); };
    fft2::run_pass<12, INV, 0>(j, tw, gld, nogst, sm, sm, nomid);
    __syncthreads();
    fft2::run_pass<12, INV, 1>(j, tw, gld, nogst, sm, sm, mid_bar);
    __syncthreads();
    const int r = threadIdx.x & 3, jq = threadIdx.x >> 2;
    DevSm smr{fsm + r * RPITCH};
    float2 *dst = out + F * 65536 + a4 + r;
    auto gst = [&](int i, float2 v) { stg_stream2(dst + 16 * i, v); };
    fft2::run_pass<12, INV, 2>(jq, tw, gld, gst, smr, smr, nomid);
}

int launch_fft65536_two_pass(const float2 *in, float2 *out, float2 *scratch, size_t scratch_frames, const float2 *twN,
                             const float2 *tw16_12, size_t nframes, bool inverse, cudaStream_t s)
{
    using CF = Fft2Cfg<12>;
    size_t done = 0;
    while (done < nframes) {
        const size_t g = nframes - done < scratch_frames ? nframes - done : scratch_frames;
        const float2 *gi = in + done * 65536;
        float2 *go = out + done * 65536;
        const unsigned gridA = (unsigned)(g * 4096 / 256);
        constexpr int SMEMB = 4 * (fft2::Plan<12>::PADN + 8) * (int)sizeof(float2);
        (void)sizeof(CF);
        if (inverse) {
            auto kb = fft65536_stepB_kernel<true>;
            CB_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEMB));
            fft65536_stepA_kernel<true><<<gridA, 256, 0, s>>>(gi, scratch, twN, g);
            kb<<<(unsigned)(g * 4), 1024, SMEMB, s>>>(scratch, go, tw16_12, g * 4);
        } else {
            auto kb = fft65536_stepB_kernel<false>;
            CB_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEMB));
            fft65536_stepA_kernel<false><<<gridA, 256, 0, s>>>(gi, scratch, twN, g);
            kb<<<(unsigned)(g * 4), 1024, SMEMB, s>>>(scratch, go, tw16_12, g * 4);
        }
        count_launch();
        count_launch();
        CB_CUDA(cudaGetLastError());
        done += g;
    }
    return CB_OK;
}

int fft_plan_split(size_t n, int *log2n1, int *log2n2)
{
    int l = 0;
    while (((size_t)1 << l) < n) ++l;
    if (((size_t)1 << l) != n || l < 14 || l > 20) return CB_ERR_UNSUPPORTED;
    *log2n1 = l / 2;
    *log2n2 = l - l / 2;
    return CB_OK;
}

// ---------------------------------------------------------------- 8192 / 16384 points: two levels inside one CTA
// N = 256 x N2 (N2 = 32 / 64), n = N2 a + b, k = k1 + 256 k2:
//   X[k1 + 256 k2] = sum_b W_N2^{b k2} { W_N^{b k1} sum_a x[N2 a + b] W_256^{a k1} }
// The generic kernel runs four radix passes over the whole frame with two block barriers each.  Here the frame is
// transposed into shared memory while it is loaded (row b = the 256 points x[N2 a + b], pitch 304 + a rotation), the 256-point
// transform of a row is done by 16 threads of ONE warp (two radix-16 passes, warp barriers), and after one block
// barrier the N2-point transform of a column k1 is done by 2 / 4 adjacent threads (radix 16 + radix 2 / 4, warp
// barriers) which store X[k1 + 256 k2] straight to global memory, 16 / 8 consecutive k1 per store instruction.
// Two block barriers per frame instead of seven -- and slower all the same (see two_level_enabled below): kept selectable.
template <int LOG2N, bool INV, bool IN16>
__global__ void __launch_bounds__((1 << LOG2N) / 16, LOG2N == 13 ? 2 : 1)
fft2_two_level_kernel(const void *__restrict__ in_, float in_scale, float2 *__restrict__ out, const float2 *__restrict__ twN,
                      size_t nframes)
{
    using namespace fft2;
    constexpr int N = 1 << LOG2N, L2 = LOG2N - 8, N2 = 1 << L2, NT = N / 16, T2 = N2 / 16, RP = 304;
    // row b starts at b * RP + (b / T2) + (16 / T2) * (b % T2): a multiple-of-16 pitch plus a rotation that puts the rows
    // of one staging store (b = 2 i or 2 i + 1) AND the T2 rows of one column access (b = j2 + T2 m) into different
    // banks -- with a plain odd pitch the 2 / 4 adjacent threads of a column sat one row = two banks apart and every
    // half-warp access took two wavefronts (profiles/r04s_ncu_full_fft8192_two_level.csv)
    auto rb = [](int b) { return b * RP + (b / T2) + (16 / T2) * (b % T2); };
    using P2 = Plan<L2>;
    extern __shared__ __align__(16) float2 fsm[];  // A[N2][RP], then twS[N2]
    float2 *twS = fsm + N2 * RP;
    const int tid = threadIdx.x;
    const size_t frame = blockIdx.x;
    if (frame >= nframes) return;
    const float2 *src = reinterpret_cast<const float2 *>(in_) + frame * N;
    float2 *dst = out + frame * N;
    if (tid < N2) twS[tid] = __ldg(twN + 256 * tid);  // W_N2^s = W_N^{256 s}: the remainder pass's table
    // coalesced load, transposed store: n = N2 a + b -> A[b][pad16(a)]
    if constexpr (IN16) {  // i16 IQ words, widened here (src/io/raw_iq.rs:78-140)
        const uint32_t *src16 = reinterpret_cast<const uint32_t *>(in_) + frame * N;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int n = tid + NT * k;
            fsm[rb(n & (N2 - 1)) + pad16(n >> L2)] = ldg_iq16(src16 + n, in_scale);
        }
    } else if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int n = 2 * tid + 2 * NT * k;
            const float4 v = ldg_stream(reinterpret_cast<const float4 *>(src + n));
            const int b = n & (N2 - 1), a = n >> L2;
            fsm[rb(b) + pad16(a)] = make_float2(v.x, v.y);
            fsm[rb(b + 1) + pad16(a)] = make_float2(v.z, v.w);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int n = tid + NT * k;
            fsm[rb(n & (N2 - 1)) + pad16(n >> L2)] = ldg_stream2(src + n);
        }
    }
    __syncthreads();
    {   // level 1: row b, 16 threads of one warp
        const int b = tid >> 4, j = tid & 15;
        float2 *row = fsm + rb(b);
        float2 v[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = row[pad16(j + 16 * m)];
        __syncwarp();
        bfly16<INV>(v);
#pragma unroll
        for (int sl = 0; sl < 16; ++sl) row[pad16(16 * j + q16(sl))] = v[sl];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = row[pad16(j + 16 * m)];
        twiddle16(v, __ldg(twN + N2 * j));  // W_256^j
        bfly16<INV>(v);
#pragma unroll
        for (int sl = 0; sl < 16; ++sl) row[pad16(j + 16 * q16(sl))] = v[sl];  // Y_b[k1] at k1: the slots this thread read
    }
    __syncthreads();
    {   // level 2: column k1, T2 adjacent threads
        const int k1 = tid / T2, j2 = tid % T2;
        float2 *col = fsm + pad16(k1);
        float2 v[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = col[rb(j2 + T2 * m)];
        twiddle16c(v, __ldg(twN + k1 * j2), __ldg(twN + k1 * T2));  // W_N^{k1 b}, b = j2 + T2 m
        auto ld0 = [&](int m) { return v[m]; };
        auto ld = [&](int m) { return col[rb(j2 + T2 * m)]; };
        auto st = [&](int idx, float2 val) { col[rb(idx)] = val; };
        auto bar = [] { __syncwarp(); };
        auto gst = [&](int idx, float2 val) { stg_stream2(dst + k1 + 256 * idx, val); };
        pass16<L2, INV, 0>(j2, twS, ld0, bar, st);
        __syncwarp();
        pass_rem<L2, INV>(j2, twS, ld, bar, gst);
    }
}

template <int LOG2N, bool INV, bool IN16 = false>
static int launch_two_level(const void *in, float2 *out, const float2 *twN, size_t nframes, cudaStream_t s, float in_scale = 1.f)
{
    constexpr int N2 = 1 << (LOG2N - 8), SMEM = (N2 * 304 + N2) * (int)sizeof(float2);
    auto kern = fft2_two_level_kernel<LOG2N, INV, IN16>;
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    kern<<<(unsigned)nframes, (1 << LOG2N) / 16, SMEM, s>>>(in, in_scale, out, twN, nframes);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// Measured (2^28 samples, % of the 16 B/sample roof, two-level / generic): 8192 points 70 / 77, 16384 points 53 / 58
// (58 / 39 before the rotated row layout removed the column accesses' bank conflicts).  Still behind: more shared-memory
// passes per point than the four-pass kernel, one frame per CTA, 64-byte store pieces at 16384 points.  The generic
// kernel stays the default; COMMS_B200_FFT_8K16K = two selects this one.
static bool two_level_enabled()
{
    const char *e = getenv("COMMS_B200_FFT_8K16K");  // read per launch: a test switches it within one process
    return e && strcmp(e, "two") == 0;
}

// true when launch_fft_iq16 reads the i16 IQ samples itself (otherwise the caller widens them first)
bool fft_fuses_iq16(const FftPlanDev &p, size_t nframes)
{
    if (p.kind == FFT_SINGLE && p.tw16 != nullptr && p.log2n >= 4) return true;
    return p.kind == FFT_FOURSTEP && p.n == 65536 && p.cluster_tpt == 6 && p.flags != nullptr && p.flags_frames >= nframes &&
           p.scratch_frames >= 4;
}

int launch_fft_iq16(const FftPlanDev &p, const int16_t *in, float in_scale, float2 *out, size_t nframes, cudaStream_t s)
{
    if (nframes == 0) return CB_OK;
    const uint32_t *in32 = reinterpret_cast<const uint32_t *>(in);
    if (p.kind == FFT_SINGLE && p.tw != nullptr && (p.log2n == 13 || p.log2n == 14) && nframes < (1ull << 31) && two_level_enabled()) {
        if (p.log2n == 13)
            return p.inverse ? launch_two_level<13, true, true>(in, out, p.tw, nframes, s, in_scale)
                             : launch_two_level<13, false, true>(in, out, p.tw, nframes, s, in_scale);
        return p.inverse ? launch_two_level<14, true, true>(in, out, p.tw, nframes, s, in_scale)
                         : launch_two_level<14, false, true>(in, out, p.tw, nframes, s, in_scale);
    }
    if (p.kind == FFT_SINGLE)
        return p.inverse ? launch_frames2_iq16_dir<true>(p.log2n, in32, in_scale, out, p.tw16, nframes, s)
                         : launch_frames2_iq16_dir<false>(p.log2n, in32, in_scale, out, p.tw16, nframes, s);
    return launch_fft65536_rows_iq16(p, in32, in_scale, out, nframes, s);
}

int launch_fft(const FftPlanDev &p, const float2 *in, float2 *out, size_t nframes, cudaStream_t s)
{
    if (nframes == 0) return CB_OK;
    if (p.kind == FFT_SINGLE && p.tw != nullptr && (p.log2n == 13 || p.log2n == 14) && nframes < (1ull << 31) && two_level_enabled()) {
        if (p.log2n == 13)
            return p.inverse ? launch_two_level<13, true>(in, out, p.tw, nframes, s) : launch_two_level<13, false>(in, out, p.tw, nframes, s);
        return p.inverse ? launch_two_level<14, true>(in, out, p.tw, nframes, s) : launch_two_level<14, false>(in, out, p.tw, nframes, s);
    }
    if (p.kind == FFT_SINGLE && p.tw16 != nullptr && p.log2n >= 4) {
        return p.inverse ? launch_frames2_dir<true>(p.log2n, in, out, p.tw16, nframes, s)
                         : launch_frames2_dir<false>(p.log2n, in, out, p.tw16, nframes, s);
    }
    if (p.kind == FFT_SINGLE) {
        return p.inverse ? launch_frames_dir<true>(p.log2n, in, out, p.tw, nframes, s)
                         : launch_frames_dir<false>(p.log2n, in, out, p.tw, nframes, s);
    }
    if (fft_big_applicable(p, nframes) && (p.n != 65536 || p.cluster_tpt == 8)) return launch_fft_big(p, in, out, nframes, s);
    if (p.kind == FFT_FOURSTEP && p.n == 65536 && p.cluster_tpt == 9)
        return launch_fft65536_cpipe(in, out, p.tw, nframes, p.inverse != 0, s);
    if (p.kind == FFT_FOURSTEP && p.n == 65536 && (p.cluster_tpt == 6 || p.cluster_tpt == 7 || p.cluster_tpt == 10))
        return launch_fft65536_rows(p, in, out, nframes, s);
    if (p.kind == FFT_FOURSTEP && p.n == 65536 && p.cluster_tpt == 5 && p.tw16 != nullptr)
        return launch_fft65536_two_pass(in, out, p.scratch, p.scratch_frames, p.tw, p.tw16, nframes, p.inverse != 0, s);
    if (p.kind == FFT_FOURSTEP && p.n == 65536 && p.cluster_tpt > 0)
        return launch_fft65536_cluster(in, out, p.tw, nframes, p.inverse != 0, p.cluster_tpt, s);
    if (p.kind == FFT_FOURSTEP) {
        const int N1 = 1 << p.log2n1, N2 = 1 << p.log2n2;
        // scratch holds as many frames as fit; process in groups
        size_t done = 0;
        while (done < nframes) {
            const size_t g = nframes - done < p.scratch_frames ? nframes - done : p.scratch_frames;
            const float2 *gi = in + done * p.n;
            float2 *go = out + done * p.n;
            int rc = p.inverse ? launch_step_dir<true>(1, p.log2n1, gi, p.scratch, p.tw1, p.tw, N2, g, s)
                               : launch_step_dir<false>(1, p.log2n1, gi, p.scratch, p.tw1, p.tw, N2, g, s);
            if (rc) return rc;
            rc = p.inverse ? launch_step_dir<true>(2, p.log2n2, p.scratch, go, p.tw2, nullptr, N1, g, s)
                           : launch_step_dir<false>(2, p.log2n2, p.scratch, go, p.tw2, nullptr, N1, g, s);
            if (rc) return rc;
            done += g;
        }
        return CB_OK;
    }
    // direct
    const int smem = (int)(p.n * sizeof(float2));
    size_t blocks = nframes < 148 * 8 ? nframes : 148 * 8;
    dft_direct_kernel<<<(unsigned)blocks, 128, smem, s>>>(in, out, p.tw, (int)p.n, nframes);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

}  // namespace cb
