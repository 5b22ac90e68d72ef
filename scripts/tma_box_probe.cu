// What the TMA engine delivers for the two access shapes of the prefetching 65536-point FFT (fft_rows_kernel.cu, K5-R2),
// with nothing else going on: persistent CTAs (3 per SM, two 33 KiB buffers each), one thread keeps two copies of 32 KiB
// in flight and re-issues as soon as one lands; no compute, no stores.
//   mode 0: step-A shape -- tensor-map box of 256 rows x 128 B out of a [256 n][2 KiB] matrix (HBM, read once)
//   mode 1: step-B shape -- 16 bulk copies of 2 KiB rows out of a ring of `ring` frames (L2 resident)
//   mode 2: one contiguous 32 KiB bulk copy (HBM, read once)
//   mode 3: step-A shape with 256-B rows (box 128 rows x 256 B)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tma_probe scripts/tma_box_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_box_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int BUF = 33024;

__global__ void __launch_bounds__(32, 3)
probe(const __grid_constant__ CUtensorMap map, const float2 *base, unsigned long long nitems, unsigned ring, int mode)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t full[2];
    const int lane = threadIdx.x;
    if (lane == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](unsigned long long item, int b) {
        const unsigned long long frame = item >> 4;
        const int part16 = (int)(item & 15) * 16;
        if (lane == 0) mbar_expect_tx(&full[b], 32768u);
        __syncwarp();
        if (mode == 0) {
            if (lane == 0) tma_box_2d(sm + b * BUF, &map, 2 * part16, (int)(frame * 256), &full[b]);
        } else if (mode == 3) {
            if (lane == 0) tma_box_2d(sm + b * BUF, &map, 4 * (part16 & ~16) , (int)(frame * 256 + (part16 & 16) * 8), &full[b]);
        } else if (mode == 1) {
            if (lane < 16) tma_bulk_1d(sm + b * BUF + lane * 2064, base + (size_t)(frame % ring) * 65536 + (size_t)(part16 + lane) * 256, 2048u, &full[b]);
        } else {
            if (lane == 0) tma_bulk_1d(sm + b * BUF, base + item * 4096, 32768u, &full[b]);
        }
    };
    unsigned long long it0 = blockIdx.x, it1 = it0 + gridDim.x;
    if (it0 < nitems) issue(it0, 0);
    if (it1 < nitems) issue(it1, 1);
    unsigned ph = 0;
    for (unsigned long long it = it0; it < nitems; it += 2ull * gridDim.x) {
        mbar_wait(&full[0], ph);
        if (it + 2ull * gridDim.x < nitems) issue(it + 2ull * gridDim.x, 0);
        if (it + gridDim.x < nitems) {
            mbar_wait(&full[1], ph);
            if (it + 3ull * gridDim.x < nitems) issue(it + 3ull * gridDim.x, 1);
        }
        ph ^= 1;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const size_t nframes = 4096, N = 65536;
    const unsigned ring = argc > 1 ? atoi(argv[1]) : 96;
    float2 *in;
    cudaMalloc(&in, nframes * N * 8);
    cudaMemset(in, 0, nframes * N * 8);
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * BUF);
    const char *names[4] = {"step-A box, 256 rows x 128 B (HBM)", "step-B, 16 x 2 KiB rows of an L2-resident ring", "one contiguous 32 KiB copy (HBM)",
                            "box of 128 rows x 256 B (HBM)"};
    for (int per_sm = 1; per_sm <= 3; per_sm += 2)
        for (int mode = 0; mode < 4; ++mode) {
            CUtensorMap map;
            const cuuint64_t dims[2] = {512, (cuuint64_t)256 * nframes};
            const cuuint64_t strides[1] = {2048};
            const cuuint32_t box[2] = {mode == 3 ? 64u : 32u, mode == 3 ? 128u : 256u}, estr[2] = {1, 1};
            if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return 1;
            const unsigned long long nitems = nframes * 16;
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                probe<<<148 * per_sm, 32, 2 * BUF>>>(map, in, nitems, ring, mode);
                cudaEventRecord(e1);
                if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("%d CTA/SM  %-52s %.3f ms  %.0f GB/s  (%.1f B/cycle/SM at 1.9 GHz)\n", per_sm, names[mode], best, nitems * 32768.0 / best / 1e6,
                   nitems * 32768.0 / best / 1e-3 / 148 / 1.9e9);
        }
    return 0;
}
