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
// a shared-memory tile of qin (register-blocked: 8 FMAs per 16-byte shared-memory load).
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

// One tile = TILE consecutive outputs; thread t owns outputs 4t .. 4t+3 of the tile and slides a 7-sample register
// window over qin, four taps per step: 4 shared-memory loads of 16 bytes feed 32 FMAs.
// Shared memory: taps, zero-padded to a multiple of four, then qin[i0 - H4 .. i0 + TILE) with H4 = halo + lead a
// multiple of four (lead = 4 .. 7 zeros that only the padding taps touch), stored de-interleaved by index mod 4 --
// L[e] at (e & 3) * pitch + (e >> 2) -- so that the threads of a warp, whose windows are 4 samples apart, read
// consecutive 16-byte words.  pitch = 2 (mod 8) keeps the fill phase (consecutive e per thread) conflict-free too.
constexpr int TILE = 4 * NT;

struct TimingGeo {
    unsigned ntaps4, lead, h4, pitch;
    size_t smem;
};

static TimingGeo timing_geo(unsigned ntaps)
{
    TimingGeo g;
    const unsigned halo = ntaps - 1;
    g.ntaps4 = (ntaps + 3) & ~3u;
    g.lead = 4 + ((4 - halo % 4) % 4);
    g.h4 = halo + g.lead;
    unsigned q = (TILE + g.h4) / 4;  // h4 and TILE are multiples of four
    while (q % 8 != 2) ++q;
    g.pitch = q;
    // taps, the de-interleaved qin tile, and din = x r for the same samples (linear; read once per output)
    g.smem = (size_t)g.ntaps4 * sizeof(double) + (size_t)4 * g.pitch * sizeof(double2) + (size_t)(TILE + g.h4) * sizeof(double2);
    return g;
}

__global__ void __launch_bounds__(NT)
timing_partial_kernel(const double2 *__restrict__ x, size_t n, const double *__restrict__ taps, unsigned ntaps,
                      unsigned ntaps4, unsigned lead, unsigned h4, unsigned pitch, unsigned nd, double sps,
                      double2 *__restrict__ partial)
{
    extern __shared__ __align__(16) unsigned char esm[];
    __shared__ double2 red[NT / 32];
    double *tsm = reinterpret_cast<double *>(esm);
    double2 *qsm = reinterpret_cast<double2 *>(esm + (size_t)ntaps4 * sizeof(double));
    double2 *dsm = qsm + 4 * (size_t)pitch;  // din[i0 - h4 + e] = x r
    const double pi = 3.14159265358979323846;
    for (unsigned k = threadIdx.x; k < ntaps4; k += NT) tsm[k] = k < ntaps ? taps[k] : 0.0;
    double2 acc = make_double2(0.0, 0.0);
    const size_t ntiles = (n + TILE - 1) / TILE;
    const int t = threadIdx.x;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long i0 = (long long)tile * TILE;
        __syncthreads();  // the previous tile's reads are done (and the taps are in)
        for (unsigned e = threadIdx.x; e < TILE + h4; e += NT) {
            const long long m = i0 - (long long)h4 + e;
            double2 q = make_double2(0.0, 0.0), d = q;
            if (e >= lead && m >= 0 && m < (long long)n) {
                double sn, cs;
                sincos(-pi * (double)m / sps, &sn, &cs);  // r = e^{-j pi m / N}, the reference's operation order
                const double2 s = x[m], r = make_double2(cs, sn);
                q = cmul64(make_double2(s.x, -s.y), r);
                d = cmul64(s, r);
            }
            qsm[(e & 3) * pitch + (e >> 2)] = q;
            dsm[e] = d;
        }
        __syncthreads();
        // window W[w + 3] = L[base + w], w = -3 .. 3, base = h4 + 4 t - 4 j for tap group j
        const double2 *c0 = qsm + (h4 >> 2) + t;  // class 0: L[base]
        const double2 *c1 = c0 + pitch, *c2 = c0 + 2 * pitch, *c3 = c0 + 3 * pitch;
        double2 W[7];
        W[0] = c1[-1];
        W[1] = c2[-1];
        W[2] = c3[-1];
        W[3] = c0[0];
        W[4] = c1[0];
        W[5] = c2[0];
        W[6] = c3[0];
        double2 y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) y[u] = make_double2(0.0, 0.0);
        for (unsigned k = 0; k < ntaps4; k += 4) {
            const double2 h01 = *reinterpret_cast<const double2 *>(tsm + k), h23 = *reinterpret_cast<const double2 *>(tsm + k + 2);
            const double h[4] = {h01.x, h01.y, h23.x, h23.y};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    y[u].x = fma(h[r], W[u - r + 3].x, y[u].x);
                    y[u].y = fma(h[r], W[u - r + 3].y, y[u].y);
                }
            // next group's base is 4 samples earlier (after the last group these loads land on the lead zeros or the
            // tap table just below them, and are not used)
            const int back = (int)(k >> 2) + 1;
            W[4] = W[0];
            W[5] = W[1];
            W[6] = W[2];
            W[0] = c1[-1 - back];
            W[1] = c2[-1 - back];
            W[2] = c3[-1 - back];
            W[3] = c0[-back];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long i = i0 + 4 * t + u;
            const long long m = i - (long long)nd;  // dout[i] = din[i - ND]
            if (i < (long long)n && m >= 0) {
                const double2 d = dsm[h4 + 4 * t + u - nd];  // nd <= halo: inside the tile's look-back
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

size_t timing_smem_bytes(unsigned ntaps) { return timing_geo(ntaps).smem; }

int launch_timing_sum(const double2 *x, size_t n, const double *taps, unsigned ntaps, unsigned nd, unsigned sps,
                      double2 *partial, double2 *out, cudaStream_t s)
{
    const TimingGeo g = timing_geo(ntaps);
    const size_t smem = g.smem;
    CB_CUDA(cudaFuncSetAttribute(timing_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    size_t blocks = ceil_div(n > 0 ? n : (size_t)1, (size_t)TILE);
    if (blocks > estimator_max_partials()) blocks = estimator_max_partials();
    timing_partial_kernel<<<(unsigned)blocks, NT, smem, s>>>(x, n, taps, ntaps, g.ntaps4, g.lead, g.h4, g.pitch, nd, (double)sps, partial);
    count_launch();
    final_sum_kernel<<<1, NT, 0, s>>>(partial, (unsigned)blocks, out);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

}  // namespace cb
