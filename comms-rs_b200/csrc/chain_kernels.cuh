#pragma once
#include "common.cuh"

namespace cb {

constexpr int CHAIN_MAX_TAP_SLOTS = 256;  // real taps: 1 slot each; complex taps: 2

struct ChainTaps {
    // real taps: t[k] = (h,h);  complex: t[2k] = (hr,hr), t[2k+1] = (hi,hi)
    float2 t[CHAIN_MAX_TAP_SLOTS];
};

struct ChainArgs {
    const float2 *x;        // channels x n_in
    const unsigned char *x8; // or: channels x n_in (u8 I, u8 Q) byte pairs, converted by (b - 127.5) / 127.5 (x unused)
    void *out;              // channels x n_out floats (FM) or complex
    const double *phase_in; // per channel (mixer), wrapped to [0, 2pi)
    double *phase_out;
    const double *dphase;
    const float2 *hist_in;  // channels x hist_len raw input samples, chronological
    float2 *hist_out;
    const float2 *prev_in;  // per channel, last FIR output (FM)
    float2 *prev_out;
    size_t n_in, n_out;
    unsigned ntaps, decim, hist_len;
    unsigned tile_out;      // decimated outputs per CTA
    unsigned span_max;      // shared-memory samples reserved for the input span
    const void *tc_img = nullptr;  // tensor-core byte front end (chain_tc_kernel.cu): prepacked tap image, or NULL
    float tc_inv_scale = 1.f, tc_dc = 0.f;
    float2 *tc_seam = nullptr;     // its scratch: first / last filter output of every tile (chain_tc_seam_entries)
};

// Tensor-core form of the byte front end (u8 IQ -> convert -> <= 64 real taps -> /5 [-> FM], no mixer): chain_tc_kernel.cu
size_t chain_tc_image_bytes();
bool chain_tc_supported(uint32_t ntaps, uint32_t decim, bool mix, bool cplx);
void chain_tc_build_image(const float *taps_re, uint32_t ntaps, unsigned char *img, float *tap_inv_scale, float *dc);
size_t chain_tc_seam_entries(size_t n_out, size_t channels);
bool chain_tc_applicable(const ChainArgs &args, size_t channels);
int launch_chain_tc(const ChainArgs &args, const ChainTaps &taps, bool fm, size_t channels, cudaStream_t s);

// true when launch_chain has a kernel that reads args.x8 itself (otherwise convert first)
bool chain_fuses_u8(const ChainArgs &args, bool cplx);
int launch_chain(const ChainArgs &args, const ChainTaps &taps, bool mix, bool fm, bool cplx, size_t channels,
                 cudaStream_t s);

}  // namespace cb
