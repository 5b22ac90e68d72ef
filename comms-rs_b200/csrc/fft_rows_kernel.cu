// K5-R: 65536-point FFT / IFFT as two streaming kernels over a 256 x 256 split, with the intermediate kept in
// an L2-sized scratch (sm_100a).
//
// Reference semantics (src/fft/mod.rs:73-96): X[k] = sum_n x[n] e^{-/+ j 2 pi k n / N}, unnormalised.
//
//   n = 256 n1 + n2,  k = k1 + 256 k2
//   X[k1 + 256 k2] = sum_{n2} W256^{n2 k2} { W65536^{n2 k1} sum_{n1} x[256 n1 + n2] W256^{n1 k1} }
//
// step A (fft65536_cols_kernel): a CTA takes 16 columns n2 of one frame (16 x 256 points, 35 KiB of shared
//   memory), does the 16 column FFTs of length 256 (two radix-16 passes in registers, one shared-memory
//   exchange) and writes Y[k1][n2] transposed into the scratch: 128 contiguous bytes per half-warp on both sides.
// step B (fft65536_rows_kernel): a CTA takes 16 rows k1 of the scratch (2 KiB contiguous each), applies
//   W^{k1 n2} on the way in, does the 16 row FFTs and scatters X[k1 + 256 k2], again 128 bytes per half-warp.
// Both kernels use 256 threads, 64 registers and one 35 KiB buffer, so four CTAs (32 warps) are resident per SM
// -- the configuration that lets fft2_frames_kernel stream at 7 TB/s; each step runs at ~6.7 TB/s.
// Algorithmic HBM traffic: 8 B read + 8 B written per sample; actual traffic is twice that (the intermediate makes
// one round trip through a scratch as large as the batch), which still beats the one-pass cluster kernel
// (1.29 ms vs 1.55 ms for 2^28 samples) because that one is bound by its DSMEM all-to-all, not by HBM.
#include "fft2_core.cuh"
#include "fft_kernels.cuh"

namespace cb {

namespace fftr {

constexpr int NF = 65536;
constexpr int RP = 273;  // row pitch (float2): pad16(256) = 272, + 1 so that rows fall into different banks
constexpr int SMEM = 16 * RP * (int)sizeof(float2);

__device__ __forceinline__ float2 ld_cs(const float2 *p)
{
    float2 v;
    asm volatile("ld.global.cs.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_cs(float2 *p, float2 v)
{
    asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// thread (c = lane & 15, j = 2 warp + lane / 16): column (or row) c of the CTA's 16, radix-16 task j
template <bool INV>
__global__ void __launch_bounds__(256, 4)
fft65536_cols_kernel(const float2 *__restrict__ in, float2 *__restrict__ mid, const float2 *__restrict__ twN)
{
    using namespace fft2;
    extern __shared__ __align__(16) float2 rsm[];
    const size_t frame = blockIdx.x >> 4;
    const int cb16 = (int)(blockIdx.x & 15) * 16;
    const int lane = threadIdx.x & 31, c = lane & 15, j = 2 * (threadIdx.x >> 5) + (lane >> 4);
    const float2 *src = in + frame * NF + cb16 + c;
    float2 *dst = mid + frame * NF + cb16 + c;
    float2 *row = rsm + c * RP;
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = ld_cs(src + 256 * (j + 16 * m));
    bfly16<INV>(v);
#pragma unroll
    for (int sl = 0; sl < 16; ++sl) row[pad16(16 * j + q16(sl))] = v[sl];
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = row[pad16(j + 16 * m)];
    twiddle16(v, __ldg(twN + 256 * j));
    bfly16<INV>(v);  // slot sl: Y[k1 = j + 16 q16(sl)][n2 = cb16 + c]
#pragma unroll
    for (int sl = 0; sl < 16; ++sl) dst[256 * (j + 16 * q16(sl))] = v[sl];
}

template <bool INV>
__global__ void __launch_bounds__(256, 4)
fft65536_rows_kernel(const float2 *__restrict__ mid, float2 *__restrict__ out, const float2 *__restrict__ twN)
{
    using namespace fft2;
    extern __shared__ __align__(16) float2 rsm[];
    const size_t frame = blockIdx.x >> 4;
    const int rb16 = (int)(blockIdx.x & 15) * 16;
    const int lane = threadIdx.x & 31, lo = lane & 15, hi = 2 * (threadIdx.x >> 5) + (lane >> 4);
    float2 v[16];
    {   // pass 3: thread (n2 digit j = lo, row r = hi): 128 contiguous bytes of one row per half-warp
        const int k1 = rb16 + hi;
        const float2 *src = mid + frame * NF + (size_t)k1 * 256 + lo;
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = src[16 * m];
        twiddle16c(v, __ldg(twN + k1 * lo), __ldg(twN + 16 * k1));  // W^{k1 (lo + 16 m)}
        bfly16<INV>(v);
        float2 *row = rsm + hi * RP;
#pragma unroll
        for (int sl = 0; sl < 16; ++sl) row[pad16(16 * lo + q16(sl))] = v[sl];
    }
    __syncthreads();
    {   // pass 4: thread (row r = lo, task j = hi): the 16 rows of one output index per half-warp
        const float2 *row = rsm + lo * RP;
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = row[pad16(hi + 16 * m)];
        twiddle16(v, __ldg(twN + 256 * hi));
        bfly16<INV>(v);  // slot sl: X[k1 = rb16 + lo + 256 (hi + 16 q16(sl))]
        float2 *dst = out + frame * NF + rb16 + lo;
#pragma unroll
        for (int sl = 0; sl < 16; ++sl) st_cs(dst + 256 * (hi + 16 * q16(sl)), v[sl]);
    }
}

// Frames are processed in groups of p.scratch_frames (the caller sizes the scratch to the whole batch when it
// can: one group = two launches measured fastest; L2-sized groups, with or without overlapping step B of one
// group with step A of the next on a second stream, or a persisting-L2 window on the scratch, all measured
// slower -- the bubbles between dependent launches cost more than the saved HBM round trip).
template <bool INV>
static int launch(const FftPlanDev &p, const float2 *in, float2 *out, size_t nframes, cudaStream_t s)
{
    auto ka = fft65536_cols_kernel<INV>;
    auto kb = fft65536_rows_kernel<INV>;
    CB_CUDA(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    CB_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const size_t gf = p.scratch_frames ? p.scratch_frames : 1;
    for (size_t done = 0; done < nframes; done += gf) {
        const size_t n = nframes - done < gf ? nframes - done : gf;
        ka<<<(unsigned)(n * 16), 256, SMEM, s>>>(in + done * NF, p.scratch, p.tw);
        count_launch();
        kb<<<(unsigned)(n * 16), 256, SMEM, s>>>(p.scratch, out + done * NF, p.tw);
        count_launch();
    }
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

}  // namespace fftr

int launch_fft65536_rows(const FftPlanDev &p, const float2 *in, float2 *out, size_t nframes, cudaStream_t s)
{
    if (nframes == 0) return CB_OK;
    return p.inverse ? fftr::launch<true>(p, in, out, nframes, s) : fftr::launch<false>(p, in, out, nframes, s);
}

}  // namespace cb
