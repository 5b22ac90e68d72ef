#pragma once
#include "common.cuh"

namespace cb {

enum FftKind { FFT_SINGLE = 0, FFT_FOURSTEP = 1, FFT_DIRECT = 2, FFT_BLUESTEIN = 3 };

// Device-side plan: twiddle tables are e^{-/+ 2 pi i k / n} rounded from f64,
// with the direction baked in.
struct FftPlanDev {
    int kind;
    int inverse;
    size_t n;
    int log2n;            // FFT_SINGLE
    int log2n1, log2n2;   // FFT_FOURSTEP: n = n1*n2
    const float2 *tw;     // n entries
    const float2 *tw1;    // n1 entries (four-step)
    const float2 *tw2;    // n2 entries (four-step)
    const float2 *tw16;   // per-pass tables of the radix-16 kernel (fft2_core.cuh), or NULL
    float2 *scratch;      // four-step intermediate, scratch_frames * n
    size_t scratch_frames;
    unsigned *flags;      // n = 65536 fused "rows" form: ticket counter (4 words) + 2 * flags_frames per-frame
                          // dependency counters (or NULL)
    size_t flags_frames;
    const float2 *tw16a, *tw16b;  // radix-16 tables of the n1- and n2-point step transforms (fused two-step form), or NULL
    int big;              // 1: sizes 2^15, 2^17 .. 2^20 run the fused two-step kernel (fft_big_kernel.cu); 0: four-step
    int cluster_tpt;      // n = 65536: 0 = four-step, 1 / 2 / 3 = cluster kernel exchange variants, 4 = 16-CTA clusters,
                          // 5 = 16 x 4096 two-pass, 6 = 256 x 256 fused two-step with an L2 ring, 7 = the same as two launches,
                          // 8 = generic fused two-step, 9 = persistent pipelined clusters (fft_cpipe_kernel.cu),
                          // 10 = 6 with the next item's points prefetched by the TMA engine (K5-R2)
};

size_t fft2_table_len(int log2n);
void fft2_fill_table(int log2n, int inverse, float2 *host_table);
int fft_plan_split(size_t n, int *log2n1, int *log2n2);
int launch_fft65536_cluster(const float2 *in, float2 *out, const float2 *twN, size_t nframes, bool inverse, int tpt,
                            cudaStream_t s);
int launch_fft65536_cpipe(const float2 *in, float2 *out, const float2 *twN, size_t nframes, bool inverse, cudaStream_t s);
int launch_fft65536_rows(const FftPlanDev &p, const float2 *in, float2 *out, size_t nframes, cudaStream_t s);
int launch_fft65536_rows_iq16(const FftPlanDev &p, const uint32_t *in, float in_scale, float2 *out, size_t nframes, cudaStream_t s);
bool fft_fuses_iq16(const FftPlanDev &p, size_t nframes);
int launch_fft_iq16(const FftPlanDev &p, const int16_t *in, float in_scale, float2 *out, size_t nframes, cudaStream_t s);
bool fft_big_applicable(const FftPlanDev &p, size_t nframes);
int launch_fft_big(const FftPlanDev &p, const float2 *in, float2 *out, size_t nframes, cudaStream_t s);
int launch_bluestein_fused(const float2 *x, const float2 *chirp, const float2 *bspec, float2 *spec, float2 *out, uint32_t N,
                           int log2m, const float2 *tw_fwd, const float2 *tw_inv, size_t frames, cudaStream_t s);
int launch_bluestein_pre(const float2 *x, const float2 *chirp, float2 *a, uint32_t N, uint32_t M, size_t frames, cudaStream_t s);
int launch_bluestein_mul(float2 *spec, const float2 *bspec, uint32_t M, size_t frames, cudaStream_t s);
int launch_bluestein_post(const float2 *c, const float2 *chirp, float2 *out, uint32_t N, uint32_t M, size_t frames,
                          cudaStream_t s);
int launch_fft(const FftPlanDev &p, const float2 *in, float2 *out, size_t nframes, cudaStream_t s);

}  // namespace cb
