// Phase timeline of the 65536-point cluster FFT kernel: compiles the kernel source with
// CB_FFTC_TIMELINE (clock64 stamps by thread 0 of every CTA) and prints mean phase durations.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/fftc_timeline scripts/fftc_timeline.cu
#define CB_FFTC_TIMELINE
#include <cstdarg>
#include <cmath>
#include <algorithm>
#include <vector>
#include "../comms-rs_b200/csrc/fft_cluster_kernel.cu"

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
    cudaMalloc(&cb::fftc::g_fftc_dbg, nframes * 16 * 12 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        if (cb::launch_fft65536_cluster(in, out, tw, nframes, false, argc > 2 ? atoi(argv[2]) : 3, 0)) return 1;
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("run %d: %.3f ms  %.1f Gsamples/s  resident clusters %d\n", it, ms, nframes * N / ms / 1e6, cb::fftc::g_fftc_resident);
    }
    std::vector<unsigned long long> d(nframes * 16 * 12);
    cudaMemcpy(d.data(), cb::fftc::g_fftc_dbg, d.size() * 8, cudaMemcpyDeviceToHost);
    const char *names[9] = {"issue global loads", "loads land + bfly y0", "bfly y1 + STS + syncthreads", "LDS + startup wait + pass2 math",
                            "wait peers ready", "push (st.async issue) + twiddle LDG", "wait pushes landed", "pass3 LDS/sync/math/STS/sync", "pass4 + STG issue"};
    const size_t cl = (argc > 2 && atoi(argv[2]) == 4) ? 16 : 8, ncta = nframes * cl;
    double sum[9] = {0}, tot = 0;
    std::vector<double> durs;
    for (size_t c = 0; c < ncta; ++c) {
        for (int k = 0; k < 9; ++k) sum[k] += (double)(d[c * 12 + k + 1] - d[c * 12 + k]);
        durs.push_back((double)(d[c * 12 + 9] - d[c * 12]));
        tot += durs.back();
    }
    for (int k = 0; k < 9; ++k) printf("  phase %d %-40s %8.0f cyc  %5.1f %%\n", k, names[k], sum[k] / ncta, 100 * sum[k] / tot);
    std::sort(durs.begin(), durs.end());
    printf("  CTA lifetime (stamp 0 -> 9): mean %.0f  p10 %.0f  p50 %.0f  p90 %.0f cyc\n", tot / ncta, durs[ncta / 10], durs[ncta / 2], durs[ncta * 9 / 10]);
    // concurrency: distinct SMs used, CTAs per SM
    std::vector<int> per(200, 0);
    for (size_t c = 0; c < ncta; ++c) per[d[c * 12 + 10] % 200]++;
    int used = 0, mn = 1 << 30, mx = 0;
    for (int s = 0; s < 200; ++s) if (per[s]) { ++used; mn = std::min(mn, per[s]); mx = std::max(mx, per[s]); }
    printf("  SMs used %d, CTAs per SM min %d max %d (ideal %.1f)\n", used, mn, mx, (double)ncta / 148);
    unsigned long long g0 = ~0ull, g1 = 0;
    for (size_t c = 0; c < ncta; ++c) { g0 = std::min(g0, d[c * 12 + 11]); g1 = std::max(g1, d[c * 12 + 11]); }
    printf("  globaltimer span of CTA ends: %.3f ms\n", (g1 - g0) / 1e6);
    return 0;
}
