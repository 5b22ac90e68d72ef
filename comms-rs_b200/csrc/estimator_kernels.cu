// K7: the f64 feed-forward estimators that call batch_fir in the reference, as fused
// filter + product + reduction kernels (sm_100a).
//
//   frequency_offset_estimate (src/demodulation/frequency_estimator.rs:27-42):
//       arg( sum_{i < n-1} x[i+1] conj(x[i]) )
//   TimingEstimator::push (src/demodulation/timing_estimator.rs:85-112), N samples per symbol, D symbols of delay:
//       r[i] = e^{-j pi i / N},  qin = conj(x) r,  din = x r
//       qout = batch_fir(qin, q) with the 2ND+1 real taps q and a zero state,  dout[i] = din[i - ND] (0 before)
//       -N arg( sum_i qout[i] dout[i] ) / (2 pi)
//
// Both are one pass over the samples (16 bytes each, complex f64) ending in a complex f64 sum.  Each CTA
// accumulates its tiles in registers, reduces through shared memory in a fixed order and writes one partial; a
// second one-CTA kernel adds the partials in index order: no atomics, so a given launch shape always returns the
// same bits.  The frequency estimator is HBM bound; the timing estimator does 2 (2ND+1) f64 FMAs per sample from
// a shared-memory tile of qin and is bound by the FP64 pipe.
#include "estimator_kernels.cuh"

namespace cb {

namespace {

constexpr int NT = 256;

__device__ __forceinline__ double2 cmul64(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// sum over the CTA in a fixed order; result valid in thread 0
__device__ __forceinline__ double2 block_sum(double2 v, double2 *red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_down_sync(0xffffffffu, v.x, o);
        v.y += __shfl_down_sync(0xffffffffu, v.y, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < NT / 32; ++w) {
            v.x += red[w].x;
            v.y += red[w].y;
        }
    }
    return v;
}

__global__ void __launch_bounds__(NT)
freq_partial_kernel(const double2 *__restrict__ x, size_t n, double2 *__restrict__ partial)
{
    __shared__ double2 red[NT / 32];
    double2 acc = make_double2(0.0, 0.0);
    const size_t stride = (size_t)gridDim.x * NT;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i + 1 < n; i += stride) {
        const double2 a = x[i + 1], b = x[i];  // latest * conj(delayed)
        acc.x += a.x * b.x + a.y * b.y;
        acc.y += a.y * b.x - a.x * b.y;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// one tile = TILE consecutive outputs i; shared memory: taps[ntaps] (f64) then qin[i0 - (ntaps-1) .. i0 + TILE)
constexpr int TILE = 4 * NT;

__global__ void __launch_bounds__(NT)
timing_partial_kernel(const double2 *__restrict__ x, size_t n, const double *__restrict__ taps, unsigned ntaps,
                      unsigned nd, double sps, double2 *__restrict__ partial)
{
    extern __shared__ __align__(16) unsigned char esm[];
    __shared__ double2 red[NT / 32];
    double *tsm = reinterpret_cast<double *>(esm);
    double2 *qsm = reinterpret_cast<double2 *>(esm + (((size_t)ntaps * sizeof(double) + 15) & ~(size_t)15));
    const double pi = 3.14159265358979323846;
    for (unsigned k = threadIdx.x; k < ntaps; k += NT) tsm[k] = taps[k];
    double2 acc = make_double2(0.0, 0.0);
    const size_t ntiles = (n + TILE - 1) / TILE;
    const unsigned halo = ntaps - 1;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long i0 = (long long)tile * TILE;
        __syncthreads();  // the previous tile's reads are done (and the taps are in)
        for (unsigned e = threadIdx.x; e < TILE + halo; e += NT) {
            const long long m = i0 - (long long)halo + e;
            double2 q = make_double2(0.0, 0.0);
            if (m >= 0 && m < (long long)n) {
                double sn, cs;
                sincos(-pi * (double)m / sps, &sn, &cs);  // r = e^{-j pi m / N}, the reference's operation order
                const double2 s = x[m];
                q = cmul64(make_double2(s.x, -s.y), make_double2(cs, sn));
            }
            qsm[e] = q;
        }
        __syncthreads();
        double2 y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) y[u] = make_double2(0.0, 0.0);
        const double2 *q0 = qsm + halo + threadIdx.x;  // qin[i0 + tid]
        for (unsigned k = 0; k < ntaps; ++k) {
            const double h = tsm[k];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double2 q = q0[u * NT - (int)k];
                y[u].x = fma(h, q.x, y[u].x);
                y[u].y = fma(h, q.y, y[u].y);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long i = i0 + threadIdx.x + u * NT;
            const long long m = i - (long long)nd;  // dout[i] = din[i - ND]
            if (i < (long long)n && m >= 0) {
                double sn, cs;
                sincos(-pi * (double)m / sps, &sn, &cs);
                const double2 d = cmul64(x[m], make_double2(cs, sn));
                const double2 p = cmul64(y[u], d);
                acc.x += p.x;
                acc.y += p.y;
            }
        }
    }
    __syncthreads();
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(NT)
final_sum_kernel(const double2 *__restrict__ partial, unsigned count, double2 *__restrict__ out)
{
    __shared__ double2 red[NT / 32];
    double2 acc = make_double2(0.0, 0.0);
    for (unsigned i = threadIdx.x; i < count; i += NT) {
        acc.x += partial[i].x;
        acc.y += partial[i].y;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) *out = acc;
}

}  // namespace

unsigned estimator_max_partials() { return 148 * 8; }

int launch_freq_sum(const double2 *x, size_t n, double2 *partial, double2 *out, cudaStream_t s)
{
    size_t blocks = ceil_div(n > 0 ? n : (size_t)1, (size_t)NT * 8);
    if (blocks > estimator_max_partials()) blocks = estimator_max_partials();
    freq_partial_kernel<<<(unsigned)blocks, NT, 0, s>>>(x, n, partial);
    count_launch();
    final_sum_kernel<<<1, NT, 0, s>>>(partial, (unsigned)blocks, out);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

size_t timing_smem_bytes(unsigned ntaps)
{
    return (((size_t)ntaps * sizeof(double) + 15) & ~(size_t)15) + ((size_t)TILE + ntaps - 1) * sizeof(double2);
}

int launch_timing_sum(const double2 *x, size_t n, const double *taps, unsigned ntaps, unsigned nd, unsigned sps,
                      double2 *partial, double2 *out, cudaStream_t s)
{
    const size_t smem = timing_smem_bytes(ntaps);
    CB_CUDA(cudaFuncSetAttribute(timing_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    size_t blocks = ceil_div(n > 0 ? n : (size_t)1, (size_t)TILE);
    if (blocks > estimator_max_partials()) blocks = estimator_max_partials();
    timing_partial_kernel<<<(unsigned)blocks, NT, smem, s>>>(x, n, taps, ntaps, nd, (double)sps, partial);
    count_launch();
    final_sum_kernel<<<1, NT, 0, s>>>(partial, (unsigned)blocks, out);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

}  // namespace cb
