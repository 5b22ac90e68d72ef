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
};

// true when launch_chain has a kernel that reads args.x8 itself (otherwise convert first)
bool chain_fuses_u8(const ChainArgs &args, bool cplx);
int launch_chain(const ChainArgs &args, const ChainTaps &taps, bool mix, bool fm, bool cplx, size_t channels,
                 cudaStream_t s);

}  // namespace cb
