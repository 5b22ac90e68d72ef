// Fused bank kernel (K2 / K2b): per channel  [mixer] -> FIR -> decimate(D) -> [FM demod],
// writing only the surviving samples.
//
// Reference node chain being fused (one instance per channel):
//   Mixer::mix           src/mixer.rs:73-84          (per-sample f64 phase)
//   batch_fir            src/filter/fir.rs:87-102    (state carried)
//   DecimateNode         src/util/resample_node.rs:53-65  (indices 0,D,2D.. of each batch)
//   FM::demod            src/modulation/analog.rs:22-34   (prev carried)
//
// Grid (tiles, channels).  A CTA stages the input span of TO decimated outputs in
// shared memory, rotating each sample once on the way in (rotation = per-thread
// base phasor x per-iteration phasor from a small shared table, both derived from
// f64 phases), evaluates only every D-th FIR output with FFMA2 and the taps in the
// kernel-parameter constant bank, then (FM) takes the angle between consecutive
// outputs.  Algorithmic HBM traffic per input sample: 8 B read + (4 or 8)/D B written.
#include "chain_kernels.cuh"
#include "misc_kernels.cuh"

namespace cb {

template <bool MIX, bool FM, bool CPLX>
__global__ void __launch_bounds__(256)
chain_kernel(const __grid_constant__ ChainArgs a, const __grid_constant__ ChainTaps taps)
{
    extern __shared__ __align__(16) float2 csm[];
    const int tid = threadIdx.x;
    const size_t c = blockIdx.y;
    const long long K = a.ntaps, D = a.decim, H = a.hist_len;
    const long long m0 = (long long)blockIdx.x * a.tile_out;
    const long long m_end = (m0 + a.tile_out < (long long)a.n_out) ? m0 + a.tile_out : (long long)a.n_out;
    const int extra = (FM && m0 > 0) ? 1 : 0;  // also compute y[m0-1] for the discriminator
    const long long m_first = m0 - extra;
    const long long g_lo = m_first * D - (K - 1);
    const long long g_hi = (m_end - 1) * D;
    const int span = (int)(g_hi - g_lo + 1);
    const int nout_tile = (int)(m_end - m_first);

    float2 *xs = csm;                                        // span samples
    float2 *ftab = csm + a.span_max;                         // per-iteration phasors
    float2 *ys = ftab + 64;                                  // tile_out + 1 outputs (FM only)

    const float2 *xc = a.x + c * a.n_in;
    const float2 *hc = a.hist_in + c * H;

    // ---- stage (and mix) the input span
    float2 e_t = make_float2(1.f, 0.f);
    if (MIX) {
        const double phi0 = a.phase_in[c], dphi = a.dphase[c];
        e_t = phase_rotation(fma((double)(g_lo + tid), dphi, phi0));
        if (tid < 64) ftab[tid] = phase_rotation((double)(tid * 256) * dphi);
        __syncthreads();
    }
    for (int i = tid, it = 0; i < span; i += 256, ++it) {
        const long long g = g_lo + i;
        float2 v = g >= 0 ? xc[g] : hc[H + g];
        if (MIX) v = cmul(v, cmul(e_t, ftab[it]));
        xs[i] = v;
    }
    __syncthreads();

    // ---- carried state for the next call (last tile of the channel)
    if (blockIdx.x == gridDim.x - 1) {
        float2 *ho = a.hist_out + c * H;
        for (long long i = tid; i < H; i += 256) {
            const long long g = (long long)a.n_in - H + i;
            ho[i] = g >= 0 ? xc[g] : hc[H + g];
        }
        if (MIX && tid == 0) {
            const double twopi = 6.283185307179586232;
            double ph = fma((double)a.n_in, a.dphase[c], a.phase_in[c]);
            ph -= twopi * floor(ph / twopi);
            a.phase_out[c] = ph;
        }
    }

    // ---- decimated FIR
    float *out_f = reinterpret_cast<float *>(a.out) + c * a.n_out;
    float2 *out_c = reinterpret_cast<float2 *>(a.out) + c * a.n_out;
    for (int o = tid; o < nout_tile; o += 256) {
        const long long m = m_first + o;
        const float2 *top = xs + (m * D - g_lo);  // x[m*D]; tap k reads top[-k]
        float2 accA = make_float2(0.f, 0.f), accB = make_float2(0.f, 0.f);
        float2 accC = make_float2(0.f, 0.f), accD = make_float2(0.f, 0.f);
        int k = 0;
        if (CPLX) {
            for (; k < (int)K; ++k) {
                const float2 s = top[-k];
                accA = __ffma2_rn(s, taps.t[2 * k], accA);
                accB = __ffma2_rn(s, taps.t[2 * k + 1], accB);
            }
        } else {
            for (; k + 1 < (int)K; k += 2) {  // two independent chains
                accA = __ffma2_rn(top[-k], taps.t[k], accA);
                accC = __ffma2_rn(top[-k - 1], taps.t[k + 1], accC);
            }
            if (k < (int)K) accA = __ffma2_rn(top[-k], taps.t[k], accA);
            accA = __fadd2_rn(accA, accC);
        }
        (void)accD;
        const float2 y = CPLX ? make_float2(accA.x - accB.y, accA.y + accB.x) : accA;
        if (FM) {
            ys[o] = y;
            if (m == (long long)a.n_out - 1) a.prev_out[c] = y;
        } else if (o >= extra) {
            out_c[m] = y;
        }
    }
    if (FM) {
        __syncthreads();
        for (int o = tid + extra; o < nout_tile; o += 256) {
            const long long m = m_first + o;
            const float2 p = o > 0 ? ys[o - 1] : a.prev_in[c];
            out_f[m] = fm_angle(ys[o], p);
        }
    }
}

int launch_chain(const ChainArgs &args, const ChainTaps &taps, bool mix, bool fm, bool cplx, size_t channels,
                 cudaStream_t s)
{
    if (args.n_in == 0 || channels == 0) return CB_OK;
    const size_t smem = (args.span_max + 64 + args.tile_out + 1) * sizeof(float2);
    const dim3 grid((unsigned)ceil_div(args.n_out, (size_t)args.tile_out), (unsigned)channels);
#define CB_CHAIN_CASE(M, F, C)                                                                         \
    if (mix == M && fm == F && cplx == C) {                                                            \
        auto kern = chain_kernel<M, F, C>;                                                             \
        CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        kern<<<grid, 256, smem, s>>>(args, taps);                                                      \
    }
    CB_CHAIN_CASE(false, false, false)
    CB_CHAIN_CASE(false, false, true)
    CB_CHAIN_CASE(false, true, false)
    CB_CHAIN_CASE(false, true, true)
    CB_CHAIN_CASE(true, false, false)
    CB_CHAIN_CASE(true, false, true)
    CB_CHAIN_CASE(true, true, false)
    CB_CHAIN_CASE(true, true, true)
#undef CB_CHAIN_CASE
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

}  // namespace cb
