// Is the SM<->L2 fabric a separate ceiling from HBM on this B200?  (FFT 65536: "fabric-bound at ~7.3 TB/s")
//   A: stream an L2-resident buffer (32 MiB, read over and over): pure L2-hit read bandwidth
//   B: copy within an L2-resident pair of buffers (16 MiB -> 16 MiB): L2-hit read + write
//   C: HBM stream copy (2 GiB -> 2 GiB): the roofline kernel's traffic
//   D: C and B at the same time in ONE kernel (each thread does both): HBM traffic + as many L2-only bytes, the mix of the
//      fused two-step FFT (8 B HBM read + 8 B ring write + 8 B ring read + 8 B HBM write per sample)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/l2_bw_probe scripts/l2_bw_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) read_kernel(const float4 *__restrict__ p, size_t n4, int reps, float *sink)
{
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 v;
            asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i));
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 1.2345f) *sink = acc;
}
__global__ void __launch_bounds__(256) copy_kernel(const float4 *__restrict__ a, float4 *__restrict__ b, size_t n4, int reps)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 v;
            asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a + i));
            asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(b + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
}
// HBM copy (big -> big2) interleaved with an L2-resident copy (ring -> ring2) of the same byte count
__global__ void __launch_bounds__(256) mix_kernel(const float4 *__restrict__ big, float4 *__restrict__ big2, size_t n4,
                                                  const float4 *__restrict__ ring, float4 *__restrict__ ring2, size_t r4)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v, w;
        asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(big + i));
        const size_t j = i % r4;
        asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w) : "l"(ring + j));
        asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(ring2 + j), "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w) : "memory");
        asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(big2 + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
}

int main()
{
    const size_t big = (size_t)2 << 30, ring = (size_t)16 << 20;
    float4 *a, *b, *r1, *r2;
    float *sink;
    cudaMalloc(&a, big); cudaMalloc(&b, big); cudaMalloc(&r1, 2 * ring); cudaMalloc(&r2, ring); cudaMalloc(&sink, 4);
    cudaMemset(a, 0, big); cudaMemset(r1, 0, 2 * ring); cudaMemset(r2, 0, ring);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int grid = 148 * 8;
    for (int pass = 0; pass < 2; ++pass) {
        cudaEventRecord(e0); read_kernel<<<grid, 256>>>(r1, 2 * ring / 16, 64, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (pass) printf("A  L2-resident read  (32 MiB x 64):            %8.1f GB/s\n", 64.0 * 2 * ring / ms / 1e6);
        cudaEventRecord(e0); copy_kernel<<<grid, 256>>>(r1, r2, ring / 16, 64); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (pass) printf("B  L2-resident copy  (16 MiB -> 16 MiB x 64):  %8.1f GB/s (read + write)\n", 64.0 * 2 * ring / ms / 1e6);
        cudaEventRecord(e0); copy_kernel<<<grid, 256>>>(a, b, big / 16, 1); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (pass) printf("C  HBM copy          (2 GiB -> 2 GiB):          %8.1f GB/s (read + write)\n", 2.0 * big / ms / 1e6);
        cudaEventRecord(e0); mix_kernel<<<grid, 256>>>(a, b, big / 16, r1, r2, ring / 16); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (pass) printf("D  HBM copy + L2-resident copy of as many bytes: %7.3f ms: HBM %8.1f GB/s, through the SM<->L2 fabric %8.1f GB/s\n", ms,
                         2.0 * big / ms / 1e6, 4.0 * big / ms / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
