// Phase sums of the prefetching fused 65536-point FFT (fft_rows_kernel.cu, K5-R2): compiles the kernel source with
// CB_FFTR_STATS (clock64 sums per phase over every item of a CTA, thread 0's view, plus prefetch / poll counts).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/fftr_timeline scripts/fftr_timeline.cu
//   /tmp/fftr_timeline [frames] [ring frames] [path: 10 = prefetching, 6 = plain]
#define CB_FFTR_STATS
#include <cstdarg>
#include <cmath>
#include <algorithm>
#include <vector>
#include "../comms-rs_b200/csrc/fft_rows_kernel.cu"

namespace cb {
void set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    fprintf(stderr, "%s:%d %s: %s\n", file, line, what, cudaGetErrorString(e));
    return CB_ERR_CUDA;
}
void count_launch() {}
}  // namespace cb

int main(int argc, char **argv)
{
    const size_t nframes = argc > 1 ? atol(argv[1]) : 4096, N = 65536;
    const size_t ring = argc > 2 ? atol(argv[2]) : 64;
    const int path = argc > 3 ? atoi(argv[3]) : 10;
    float2 *in, *out, *tw, *scratch;
    unsigned *flags;
    cudaMalloc(&in, nframes * N * 8);
    cudaMalloc(&out, nframes * N * 8);
    cudaMalloc(&tw, N * 8);
    cudaMalloc(&scratch, ring * N * 8);
    cudaMalloc(&flags, (4 + 2 * nframes) * 4);
    cudaMemset(in, 0, nframes * N * 8);
    std::vector<float2> htw(N);
    for (size_t k = 0; k < N; ++k) htw[k] = make_float2((float)cos(-2 * M_PI * k / N), (float)sin(-2 * M_PI * k / N));
    cudaMemcpy(tw, htw.data(), N * 8, cudaMemcpyHostToDevice);
    const size_t ncta_max = 148 * 4;
    unsigned long long *dbg;
    cudaMalloc(&dbg, ncta_max * 12 * 8);
    cudaMemset(dbg, 0, ncta_max * 12 * 8);
    cudaMemcpyToSymbol(cb::fftr::g_fftr_dbg, &dbg, sizeof(dbg));
    cb::FftPlanDev p = {};
    p.kind = 0;
    p.n = N;
    p.tw = tw;
    p.scratch = scratch;
    p.scratch_frames = ring;
    p.flags = flags;
    p.flags_frames = nframes;
    p.cluster_tpt = path;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        if (cb::launch_fft65536_rows(p, in, out, nframes, 0)) return 1;
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("run %d: %.3f ms  %.1f Gsamples/s (memset of the counters included)\n", it, ms, nframes * N / ms / 1e6);
    }
    if (path != 10) return 0;
    std::vector<unsigned long long> d(ncta_max * 12);
    cudaMemcpy(d.data(), dbg, d.size() * 8, cudaMemcpyDeviceToHost);
    const char *nm[6] = {"-", "wait for the item's points (mbarrier)", "first pass (LDS, radix 16, STS)", "mid barrier",
                         "global stores + proxy fence", "end barrier"};
    double sum[12] = {0};
    size_t ncta = 0;
    for (size_t c = 0; c < ncta_max; ++c) {
        if (d[c * 12 + 6] == 0) continue;
        ++ncta;
        for (int k = 0; k < 12; ++k) sum[k] += (double)d[c * 12 + k];
    }
    double tot = sum[8];
    for (int k = 0; k < 6; ++k) tot += sum[k];
    printf("CTAs %zu, items %.0f (%.1f per CTA), copies started a round ahead: %.1f %% of items\n", ncta, sum[6], sum[6] / ncta,
           100 * sum[7] / sum[6]);
    printf("cycles per item (thread 0): %.0f\n", tot / sum[6]);
    for (int k = 1; k < 6; ++k) {
        if (k == 4) printf("  %5.1f %%  %7.0f cyc  second pass (LDS, twiddles, radix 16)\n", 100 * sum[8] / tot, sum[8] / sum[6]);
        printf("  %5.1f %%  %7.0f cyc  %s\n", 100 * sum[k] / tot, sum[k] / sum[6], nm[k]);
    }
    return 0;
}
