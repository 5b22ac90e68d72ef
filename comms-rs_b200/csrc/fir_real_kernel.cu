// K1-R: real-input, real-output decimating FIR -- the second stage of examples/fm_radio.rs
// (Convert2Node -> filt2 -> Convert3Node -> dec2, fm_radio.rs:98-164) as ONE kernel (sm_100a).
//
// Reference semantics: x -> Complex(x, 0) (fm_radio.rs:107-113), batch_fir with the node's taps
// (src/filter/fir.rs:87-102), .re (fm_radio.rs:131-137), DecimateNode (src/util/resample_node.rs:53-65):
//   y[J] = sum_k Re(h[k]) x[J D - k],   x[-1-k] = Re(state[k])
// (the imaginary parts of the taps only ever meet the zero imaginary parts of the samples).  HBM carries 4 bytes per
// input sample and 4 / D per output instead of the 24 bytes per sample of the four separate nodes.
//
// Polyphase form, so that the decimation stride never reaches shared memory: with k = q D + p,
//   y[J] = sum_p sum_q h[q D + p] A_p[J - q],   A_p[m] = x[m D - p]
// A tile of TO = 4 NT outputs de-interleaves its input range into the D arrays A_p (coalesced global reads, one
// pass); thread t owns the four consecutive outputs J0 + 4t .. 4t+3 and, per phase, reads the LB + 4 consecutive
// values A_p[J0 + 4t - LB .. J0 + 4t + 3] as aligned 128-bit words (consecutive threads -> consecutive words:
// conflict-free) and does 4 Q FMAs on them with the taps as constant-bank operands: 4 (LB/4 + 1) D shared-memory
// loads for 4 Q D FMAs per thread, which leaves the kernel HBM-bound.
#include "fir_kernels.cuh"

namespace cb {

namespace {

constexpr int NT = 256;
constexpr int TO = 4 * NT;

template <int D>
struct RGeo {
    static constexpr int Q = (64 + D - 1) / D;          // taps per phase (64 taps, zero-padded)
    static constexpr int LB = (Q - 1 + 3) / 4 * 4;      // look-back in outputs, a multiple of four
    static constexpr int NW = LB / 4 + 1;               // 128-bit words per phase and thread
    static constexpr int LEN = TO + LB;                 // entries per phase array
    // multiple of four (aligned words); = 12 (mod 32) spreads the D phases of consecutive input samples over the banks
    static constexpr int PITCH = (LEN + 3) / 4 * 4 + ((12 - (LEN + 3) / 4 * 4 % 32 + 32) % 32);
    static constexpr int SMEM = D * PITCH * (int)sizeof(float);
};

struct RealTaps {
    float t[80];  // h[q D + p], zero beyond the filter
};

struct RealArgs {
    const float *x;          // n_in real samples
    const float2 *hist_in;   // hist_len complex samples preceding x[0], chronological; real parts used
    float *y;                // n_out = ceil(n_in / D) outputs
    unsigned long long n_in, n_out;
    unsigned hist_len;
};

template <int D>
__global__ void __launch_bounds__(NT, 4)
fir_real_decim_kernel(const __grid_constant__ RealArgs a, const __grid_constant__ RealTaps taps)
{
    using G = RGeo<D>;
    extern __shared__ __align__(16) float rsm[];
    const int t = threadIdx.x;
    const long long ntiles = ((long long)a.n_out + TO - 1) / TO;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long J0 = tile * TO;
        // A_p[i] = x[(J0 - LB + i) D - p]: linear input index g = G0 + e  <->  i = e / D, p = D - 1 - e % D
        const long long G0 = (J0 - G::LB) * D - (D - 1);
        __syncthreads();  // the previous tile's reads are done
        if (G0 < 0) {  // first tile: the look-back reaches into the carried history
            for (int e = t; e < G::LEN * D; e += NT) {
                const long long g = G0 + e;
                float v = 0.f;
                if (g >= 0) {
                    if (g < (long long)a.n_in) v = __ldg(a.x + g);
                } else if (g + (long long)a.hist_len >= 0) {
                    v = a.hist_in[(long long)a.hist_len + g].x;
                }
                rsm[(D - 1 - e % D) * G::PITCH + e / D] = v;
            }
        } else if ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0) {
            // 128-bit loads from the 16-byte boundary at or below G0, groups of up to 8 per thread in flight before
            // their shared-memory stores
            const long long G0a = G0 & ~3LL;
            const int skip = (int)(G0 - G0a);
            constexpr int NV = (G::LEN * D + 3 + 3) / 4;
            constexpr int NLD = (NV + NT - 1) / NT;
            constexpr int GRP = NLD < 8 ? NLD : 8;
            const float *src = a.x + G0a;
            const long long left = (long long)a.n_in - G0a;  // samples available from G0a on
#pragma unroll 1
            for (int b = 0; b < NLD; b += GRP) {
                float4 v[GRP];
#pragma unroll
                for (int it = 0; it < GRP; ++it) {
                    const int idx = t + (b + it) * NT;
                    v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx < NV) {
                        if (4 * idx + 3 < left) {
                            v[it] = __ldg(reinterpret_cast<const float4 *>(src) + idx);
                        } else {  // the vector that straddles the end of the batch
                            if (4 * idx + 0 < left) v[it].x = __ldg(src + 4 * idx);
                            if (4 * idx + 1 < left) v[it].y = __ldg(src + 4 * idx + 1);
                            if (4 * idx + 2 < left) v[it].z = __ldg(src + 4 * idx + 2);
                        }
                    }
                }
#pragma unroll
                for (int it = 0; it < GRP; ++it) {
                    const int e0 = 4 * (t + (b + it) * NT) - skip;
                    const float c[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int e = e0 + k;
                        if (e >= 0 && e < G::LEN * D) rsm[(D - 1 - e % D) * G::PITCH + e / D] = c[k];
                    }
                }
            }
        } else {  // unaligned input: scalar loads, groups of up to 16 per thread in flight
            constexpr int NLD = (G::LEN * D + NT - 1) / NT;
            constexpr int GRP = 16;
            const float *src = a.x + G0;
            const long long left = (long long)a.n_in - G0;  // samples available from G0 on
#pragma unroll 1
            for (int b = 0; b < NLD; b += GRP) {
                float v[GRP];
#pragma unroll
                for (int it = 0; it < GRP; ++it) {
                    const int e = t + (b + it) * NT;
                    v[it] = (e < G::LEN * D && e < left) ? __ldg(src + e) : 0.f;
                }
#pragma unroll
                for (int it = 0; it < GRP; ++it) {
                    const int e = t + (b + it) * NT;
                    if (e < G::LEN * D) rsm[(D - 1 - e % D) * G::PITCH + e / D] = v[it];
                }
            }
        }
        __syncthreads();
        float y[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int p = 0; p < D; ++p) {
            float w[G::NW * 4];
            const float4 *src = reinterpret_cast<const float4 *>(rsm + p * G::PITCH) + t;
#pragma unroll
            for (int j = 0; j < G::NW; ++j) {
                const float4 q4 = src[j];
                w[4 * j] = q4.x;
                w[4 * j + 1] = q4.y;
                w[4 * j + 2] = q4.z;
                w[4 * j + 3] = q4.w;
            }
#pragma unroll
            for (int q = 0; q < G::Q; ++q) {
                if (q * D + p < 64) {
                    const float h = taps.t[q * D + p];
#pragma unroll
                    for (int u = 0; u < 4; ++u) y[u] = fmaf(h, w[u - q + G::LB], y[u]);
                }
            }
        }
        const long long J = J0 + 4 * t;
        if (J + 3 < (long long)a.n_out && (reinterpret_cast<uintptr_t>(a.y) & 15) == 0) {
            *reinterpret_cast<float4 *>(a.y + J) = make_float4(y[0], y[1], y[2], y[3]);
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (J + u < (long long)a.n_out) a.y[J + u] = y[u];
        }
    }
}

// hist_out[i] = (x[n - H + i], 0), older entries from hist_in
__global__ void __launch_bounds__(256)
real_hist_update_kernel(const float *__restrict__ x, unsigned long long n, const float2 *__restrict__ hist_in,
                        float2 *__restrict__ hist_out, unsigned H)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < H; i += gridDim.x * blockDim.x) {
        const long long g = (long long)n - H + i;
        hist_out[i] = g >= 0 ? make_float2(x[g], 0.f) : hist_in[H + g];
    }
}

template <int D>
int launch_d(const RealArgs &a, const RealTaps &taps, cudaStream_t s)
{
    using G = RGeo<D>;
    auto kern = fir_real_decim_kernel<D>;
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));
    const size_t tiles = ceil_div((size_t)a.n_out, (size_t)TO);
    const size_t cap = 148 * 4 * 4;  // a few waves of resident CTAs; tiles are strided over the grid
    kern<<<(unsigned)(tiles < cap ? tiles : cap), NT, G::SMEM, s>>>(a, taps);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

}  // namespace

bool fir_real_applicable(uint32_t ntaps, uint32_t interp, uint32_t decim)
{
    return interp == 1 && ntaps >= 1 && ntaps <= 64 && (decim == 2 || decim == 4 || decim == 5 || decim == 8 || decim == 10);
}

int launch_fir_real(const float *x, size_t n_in, const float2 *hist_in, float2 *hist_out, uint32_t hist_len,
                    const float2 *taps_host, uint32_t ntaps, uint32_t decim, float *y, cudaStream_t s)
{
    if (n_in == 0) return CB_OK;
    RealTaps taps;
    for (int k = 0; k < 80; ++k) taps.t[k] = k < (int)ntaps ? taps_host[k].x : 0.f;
    RealArgs a{x, hist_in, y, n_in, ceil_div(n_in, (size_t)decim), hist_len};
    int rc;
    switch (decim) {
    case 2: rc = launch_d<2>(a, taps, s); break;
    case 4: rc = launch_d<4>(a, taps, s); break;
    case 5: rc = launch_d<5>(a, taps, s); break;
    case 8: rc = launch_d<8>(a, taps, s); break;
    case 10: rc = launch_d<10>(a, taps, s); break;
    default: set_error("fir: no real-input kernel for decimation %u", decim); return CB_ERR_UNSUPPORTED;
    }
    if (rc) return rc;
    if (hist_out != nullptr) {
        real_hist_update_kernel<<<(hist_len + 255) / 256, 256, 0, s>>>(x, n_in, hist_in, hist_out, hist_len);
        count_launch();
        CB_CUDA(cudaGetLastError());
    }
    return CB_OK;
}

}  // namespace cb
