// Phase timeline of the tensor-core byte front end (chain_tc_kernel.cu): compiled with CB_CTC_TIMELINE, clock64 stamps of
// local items 20 and 21 (one per converter group) of every CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ctc_timeline scripts/ctc_timeline.cu
#define CB_CTC_TIMELINE
#include <cstdarg>
#include <cmath>
#include <algorithm>
#include <vector>
#include "../comms-rs_b200/csrc/chain_tc_kernel.cu"

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
    const size_t channels = argc > 1 ? atol(argv[1]) : 1024, n_in = argc > 2 ? atol(argv[2]) : 131072, n_out = (n_in + 4) / 5, H = 128;
    unsigned char *x8, *img;
    float *out;
    float2 *hist[2], *prev[2];
    cudaMalloc(&x8, channels * n_in * 2);
    cudaMemset(x8, 0x55, channels * n_in * 2);
    cudaMalloc(&out, channels * n_out * 4);
    for (int i = 0; i < 2; ++i) {
        cudaMalloc(&hist[i], channels * H * 8);
        cudaMemset(hist[i], 0, channels * H * 8);
        cudaMalloc(&prev[i], channels * 8);
        cudaMemset(prev[i], 0, channels * 8);
    }
    std::vector<float> taps(63);
    for (int k = 0; k < 63; ++k) taps[k] = (float)(0.2 * std::cos(0.05 * (k - 31)) / (1 + std::abs(k - 31)));
    std::vector<unsigned char> himg(cb::chain_tc_image_bytes());
    cb::ChainArgs a = {};
    cb::chain_tc_build_image(taps.data(), 63, himg.data(), &a.tc_inv_scale, &a.tc_dc);
    cudaMalloc(&img, himg.size());
    cudaMemcpy(img, himg.data(), himg.size(), cudaMemcpyHostToDevice);
    cb::ChainTaps ct = {};
    for (int k = 0; k < 63; ++k) ct.t[k] = make_float2(taps[k], taps[k]);
    a.x8 = x8; a.out = out; a.hist_in = hist[0]; a.hist_out = hist[1]; a.prev_in = prev[0]; a.prev_out = prev[1];
    a.n_in = n_in; a.n_out = n_out; a.ntaps = 63; a.decim = 5; a.hist_len = (unsigned)H; a.tc_img = img;
    cudaMalloc(&a.tc_seam, cb::chain_tc_seam_entries(n_out, channels) * 8);
    cudaMalloc(&cb::g_ctc_dbg, 148 * 2 * 24 * 8);
    cudaMemset(cb::g_ctc_dbg, 0, 148 * 2 * 24 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        if (cb::launch_chain_tc(a, ct, true, channels, 0)) return 1;
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("run %d: %.3f ms  %.1f Gsamples/s\n", it, ms, channels * n_in / ms / 1e6);
    }
    std::vector<unsigned long long> d(148 * 2 * 24);
    cudaMemcpy(d.data(), cb::g_ctc_dbg, d.size() * 8, cudaMemcpyDeviceToHost);
    struct { int a, b; const char *name; } ph[] = {
        {0, 1, "conv: wait raw_full (TMA landed)"}, {1, 2, "conv: wait a_empty (MMAs of item-2 done)"}, {2, 3, "conv: convert (thread 0: 3 units)"},
        {2, 4, "conv: convert (thread 127: 2 units)"}, {3, 5, "conv: raw_empty arrive"},
        {8, 9, "mma: wait t_empty (epilogue of item-2 drained TMEM)"}, {9, 10, "mma: wait a_full (converted)"}, {10, 11, "mma: issue 20 MMAs + commits"},
        {12, 13, "epi: wait t_full (MMAs done)"}, {13, 14, "epi: tcgen05.ld + t_empty arrive"}, {14, 15, "epi: unscale"}, {15, 16, "epi: barrier (8 warps)"},
        {16, 17, "epi: discriminator + stores"}, {20, 21, "tma: wait raw_empty"},
        {2, 10, "  conv start -> mma sees a_full"}, {11, 13, "  mma commit -> epilogue sees t_full"}, {0, 17, "  item: conv wait start -> epilogue end"}};
    for (auto &p : ph) {
        double sum = 0; size_t n = 0;
        for (size_t c = 0; c < 148 * 2; ++c) {
            const unsigned long long x = d[c * 24 + p.a], y = d[c * 24 + p.b];
            if (x && y && y >= x) { sum += (double)(y - x); ++n; }
        }
        printf("  %-55s %8.0f cyc  (%zu)\n", p.name, n ? sum / n : 0.0, n);
    }
    // spacing between the two consecutive items of a CTA (item period)
    double sp = 0; size_t n = 0;
    for (size_t c = 0; c < 148; ++c) {
        const unsigned long long x = d[(c * 2) * 24 + 17], y = d[(c * 2 + 1) * 24 + 17];
        if (x && y && y > x) { sp += (double)(y - x); ++n; }
    }
    printf("  epilogue end of item 20 -> item 21: %.0f cyc (%zu)\n", n ? sp / n : 0.0, n);
    return 0;
}
