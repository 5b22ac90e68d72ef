// K5-R: 65536-point FFT / IFFT over a 256 x 256 split: two streaming steps, run either as ONE persistent kernel
// with the intermediate in an L2-resident ring (fft65536_fused_kernel, the default) or as two launches with a
// batch-sized scratch (sm_100a).
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

// ---------------------------------------------------------------------------------------------------------
// Fused form: ONE persistent kernel runs both steps, step A running `lag` frames ahead of step B, with the
// intermediate in a small ring of scratch frames that stays resident in L2 (written by step A, read by step B a
// few microseconds later, overwritten `ring` frames later while still dirty in L2), so that it never travels to
// HBM.  Work items (16 per frame and step) are taken in a fixed order -- A(0..lag-1), then A(lag+u), B(u)
// alternating -- from a global ticket counter, so an item only ever waits for items with smaller tickets, and a
// ticket is only ever held by a CTA that is running: the earliest unfinished item can always proceed, whatever
// share of the grid is resident (two such kernels on different streams cannot deadlock each other, as a static
// item-to-CTA assignment could).  Per-frame counters in global memory carry the
// dependencies (release: st + __threadfence + atomicAdd; acquire: ld.acquire.gpu polled by warp 0 as a whole, then the
// block barrier; the scratch is read with
// ld.global.cg so that no stale L1 line of an earlier use of the slot is seen).
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// A whole warp waits, converged, until *p >= need (lane 0 reads, the value is broadcast).  A single thread spinning
// ahead of a block barrier while the other 31 lanes of its warp run on was observed to let the barrier release
// without it; keeping the warp together avoids depending on that.
__device__ __forceinline__ void poll_at_least(const unsigned *p, unsigned need, int lane)
{
    for (;;) {
        unsigned f = lane == 0 ? ld_acquire(p) : 0u;
        f = __shfl_sync(0xffffffffu, f, 0);
        if (f >= need) break;
        __nanosleep(64);
    }
}

template <bool INV>
__global__ void __launch_bounds__(256, 4)
fft65536_fused_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, float2 *__restrict__ scratch,
                      const float2 *__restrict__ twN, unsigned *ticket, unsigned *flags_a, unsigned *flags_b,
                      unsigned long long nframes, unsigned lag, unsigned ring)
{
    using namespace fft2;
    extern __shared__ __align__(16) float2 rsm[];
    __shared__ unsigned s_item;
    const int lane = threadIdx.x & 31, lo = lane & 15, hi = 2 * (threadIdx.x >> 5) + (lane >> 4);
    const unsigned long long nchunks = lag + 2ull * nframes;
    for (;;) {
        __syncthreads();  // the previous item's shared-memory reads (and its read of s_item) are done
        if (threadIdx.x == 0) s_item = atomicAdd(ticket, 1u);
        __syncthreads();
        const unsigned long long item = s_item;
        if (item >= nchunks * 16) break;
        const unsigned long long chunk = item >> 4;
        const int part16 = (int)(item & 15) * 16;
        bool is_a;
        unsigned long long frame;
        if (chunk < lag) {
            is_a = true;
            frame = chunk;
        } else {
            const unsigned long long t = chunk - lag;
            is_a = (t & 1) == 0;
            frame = is_a ? lag + (t >> 1) : (t >> 1);
        }
        if (frame >= nframes) continue;
        float2 *slot = scratch + (frame % ring) * NF;
        float2 v[16];
        if (is_a) {
            const float2 *src = in + frame * NF + part16 + lo;
            float2 *row = rsm + lo * RP;
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = ld_cs(src + 256 * (hi + 16 * m));
            // the slot must have been consumed by its last reader before anything is stored into it: warp 0 polls
            // (all 32 lanes together) while the input loads are in flight; the block barrier below orders the
            // observation before every thread's stores
            if (frame >= ring && threadIdx.x < 32) poll_at_least(flags_b + (frame - ring), 16u, lane);
            bfly16<INV>(v);
#pragma unroll
            for (int sl = 0; sl < 16; ++sl) row[pad16(16 * hi + q16(sl))] = v[sl];
            __syncthreads();  // (also orders thread 0's slot wait before every thread's stores below)
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = row[pad16(hi + 16 * m)];
            twiddle16(v, __ldg(twN + 256 * hi));
            bfly16<INV>(v);
            float2 *dst = slot + part16 + lo;
#pragma unroll
            for (int sl = 0; sl < 16; ++sl) __stcg(dst + 256 * (hi + 16 * q16(sl)), v[sl]);
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                atomicAdd(flags_a + frame, 1u);
            }
        } else {
            // all 16 column blocks of the frame are in
            if (threadIdx.x < 32) poll_at_least(flags_a + frame, 16u, lane);
            __syncthreads();
            const int k1 = part16 + hi;
            const float2 *src = slot + (size_t)k1 * 256 + lo;
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = __ldcg(src + 16 * m);
            twiddle16c(v, __ldg(twN + k1 * lo), __ldg(twN + 16 * k1));
            bfly16<INV>(v);
            float2 *rowp = rsm + hi * RP;
#pragma unroll
            for (int sl = 0; sl < 16; ++sl) rowp[pad16(16 * lo + q16(sl))] = v[sl];
            __syncthreads();  // every thread has consumed its scratch reads
            if (threadIdx.x == 0) {
                __threadfence();
                atomicAdd(flags_b + frame, 1u);
            }
            const float2 *row = rsm + lo * RP;
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = row[pad16(hi + 16 * m)];
            twiddle16(v, __ldg(twN + 256 * hi));
            bfly16<INV>(v);
            float2 *dst = out + frame * NF + part16 + lo;
#pragma unroll
            for (int sl = 0; sl < 16; ++sl) st_cs(dst + 256 * (hi + 16 * q16(sl)), v[sl]);
        }
    }
}

template <bool INV>
static int launch_fused(const FftPlanDev &p, const float2 *in, float2 *out, size_t nframes, cudaStream_t s)
{
    auto kf = fft65536_fused_kernel<INV>;
    CB_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kf, 256, SMEM);
    if (per_sm < 1) per_sm = 1;
    const unsigned ring = (unsigned)p.scratch_frames;
    unsigned lag = ring / 2;
    if (lag < 1) lag = 1;
    const unsigned long long items = (lag + 2ull * nframes) * 16;
    const unsigned long long cap = (unsigned long long)sms * per_sm;  // one CTA per resident slot
    const unsigned grid = (unsigned)(items < cap ? items : cap);
    CB_CUDA(cudaMemsetAsync(p.flags, 0, (4 + 2 * nframes) * sizeof(unsigned), s));
    kf<<<grid, 256, SMEM, s>>>(in, out, p.scratch, p.tw, p.flags, p.flags + 4, p.flags + 4 + nframes, nframes, lag, ring);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
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
    if (p.cluster_tpt == 6 && p.flags != nullptr && p.flags_frames >= nframes && p.scratch_frames >= 4)
        return p.inverse ? fftr::launch_fused<true>(p, in, out, nframes, s) : fftr::launch_fused<false>(p, in, out, nframes, s);
    return p.inverse ? fftr::launch<true>(p, in, out, nframes, s) : fftr::launch<false>(p, in, out, nframes, s);
}

}  // namespace cb
