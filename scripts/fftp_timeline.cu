// Phase timeline of the pipelined-cluster 65536-point FFT (fft_cpipe_kernel.cu): compiles the kernel source with
// CB_FFTP_TIMELINE (clock64 stamps of one steady-state iteration, compute thread 0 and DMA lane 0 of every CTA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/fftp_timeline scripts/fftp_timeline.cu
#define CB_FFTP_TIMELINE
#include <cstdarg>
#include <cmath>
#include <algorithm>
#include <vector>
#include "../comms-rs_b200/csrc/fft_cpipe_kernel.cu"

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
    float2 *in, *out, *tw;
    cudaMalloc(&in, nframes * N * 8);
    cudaMalloc(&out, nframes * N * 8);
    cudaMalloc(&tw, N * 8);
    cudaMemset(in, 0, nframes * N * 8);
    std::vector<float2> htw(N);
    for (size_t k = 0; k < N; ++k) htw[k] = make_float2((float)cos(-2 * M_PI * k / N), (float)sin(-2 * M_PI * k / N));
    cudaMemcpy(tw, htw.data(), N * 8, cudaMemcpyHostToDevice);
    const size_t ncta_max = 148 * 2;
    cudaMalloc(&cb::fftp::g_fftp_dbg, ncta_max * 16 * 8);
    cudaMemset(cb::fftp::g_fftp_dbg, 0, ncta_max * 16 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        if (cb::launch_fft65536_cpipe(in, out, tw, nframes, false, 0)) return 1;
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("run %d: %.3f ms  %.1f Gsamples/s  resident clusters %d\n", it, ms, nframes * N / ms / 1e6, cb::fftp::g_fftp_resident);
    }
    const size_t ncta = (size_t)cb::fftp::g_fftp_resident * 8;
    std::vector<unsigned long long> d(ncta_max * 16);
    cudaMemcpy(d.data(), cb::fftp::g_fftp_dbg, d.size() * 8, cudaMemcpyDeviceToHost);
    const char *cn[5] = {"compute: wait full[b] (TMA load landed)", "compute: step A (2 passes, 4 barriers)", "compute: wait landed (pushes of it-1)",
                         "compute: step B pass 3 + reads of pass 4", "compute: pass 4 math + global stores"};
    const char *dn[6] = {"dma: wait send_rdy (A done)", "dma: wait ready (peers drained EXCH)", "dma: issue 8 pushes", "dma: wait landed",
                         "dma: remote arrives + wait sent_ok", "dma: issue 256 TMA loads"};
    double cs[5] = {0}, ds[6] = {0}, ctot = 0, dtot = 0;
    for (size_t c = 0; c < ncta; ++c) {
        for (int k = 0; k < 5; ++k) cs[k] += (double)(d[c * 16 + k + 1] - d[c * 16 + k]);
        ctot += (double)(d[c * 16 + 5] - d[c * 16]);
        for (int k = 0; k < 6; ++k) ds[k] += (double)(d[c * 16 + 9 + k] - d[c * 16 + 8 + k]);
        dtot += (double)(d[c * 16 + 14] - d[c * 16 + 8]);
    }
    for (int k = 0; k < 5; ++k) printf("  %-50s %8.0f cyc  %5.1f %%\n", cn[k], cs[k] / ncta, 100 * cs[k] / ctot);
    printf("  compute iteration (A(it) + B(it-1)): mean %.0f cyc\n", ctot / ncta);
    for (int k = 0; k < 6; ++k) printf("  %-50s %8.0f cyc  %5.1f %%\n", dn[k], ds[k] / ncta, 100 * ds[k] / dtot);
    printf("  dma iteration: mean %.0f cyc\n", dtot / ncta);
    return 0;
}
