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
#include <cstdlib>
#include <cstring>

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
    if (blockIdx.x == gridDim.x - 1 && a.hist_out != nullptr) {
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

// ============================================================================
// v2 (real taps <= KP, decimation D known at compile time): register-tiled decimating FIR.
// A thread owns R consecutive decimated outputs; the staged span lives in shared memory in
// chunks of R*D samples (one chunk per thread, pitch R*D*8 + 16 bytes so that the per-thread
// 128-bit window loads are bank-conflict free).  Every staged sample is read ONCE per thread
// (LDS.128 = 2 samples) and feeds up to R FFMA2s, instead of one LDS.64 per tap per output:
// shared-memory traffic drops from ~58 to ~27 B per input sample, below the HBM-equivalent rate.
// Global loads are 128-bit (2 samples) with the mixer rotation applied on the way in.
// ============================================================================
// complex product a*b with packed FP32: FMUL2 + FFMA2 (scalar-broadcast operands), brr = (-b.y, b.x)
__device__ __forceinline__ float2 rot90(float2 b) { return make_float2(-b.y, b.x); }
__device__ __forceinline__ float2 cmul_p(float2 a, float2 b, float2 brr)
{
    return __ffma2_rn(make_float2(a.x, a.x), b, __fmul2_rn(make_float2(a.y, a.y), brr));
}

// PF = true : persistent CTA, the raw span of item i+1 is prefetched into registers while item i is filtered
// PF = false: the span is loaded in batches of NB 128-bit loads at the start of each item; latency is covered
//             by the other CTAs of the SM (small tiles, several CTAs per SM)
template <bool MIX, bool FM, int D, int KP, int R, int NT, bool PF, int MINB>
__global__ void __launch_bounds__(NT, MINB)
chain2_kernel(const __grid_constant__ ChainArgs a, const __grid_constant__ ChainTaps taps, const unsigned tiles_per_ch,
              const unsigned long long nitems)
{
    constexpr int TO = NT * R;                          // decimated outputs per tile
    constexpr int RD = R * D;                           // input samples per thread chunk
    constexpr int OFF = ((KP + D + RD - 1) / RD) * RD;  // halo samples kept in front of the tile
    constexpr int SPAN = TO * D + OFF;
    constexpr int PITCH = RD * 8 + 16;
    constexpr int NCH = NT + OFF / RD;
    constexpr int NIT = (SPAN / 2 + NT - 1) / NT;
    static_assert(RD % 4 == 0, "chunk pitch must be an odd multiple of 16 bytes");
    constexpr int NB = PF ? NIT : 6;                    // loads in flight per thread
    static_assert(NIT <= 61 && NT >= 64 && KP % 2 == 0 && KP <= 64, "");

    extern __shared__ __align__(16) unsigned char c2sm[];
    unsigned char *xs = c2sm;                                               // NCH chunks
    float2 *ftab = reinterpret_cast<float2 *>(c2sm + NCH * PITCH);          // 2 x 64 phasors
    float2 *ys = ftab + 128;                                                // TO + 1 outputs

    const int tid = threadIdx.x;
    const long long H = a.hist_len;

    // Persistent CTA over work items (channel, tile).  The raw span of item i+1 is loaded into
    // registers while item i is filtered, so HBM loads are always in flight.
    float4 raw[NB];
    auto issue_loads = [&](unsigned long long item, const int b0) {
        const size_t c = (size_t)((unsigned)item / tiles_per_ch);
        const long long g_base = (long long)((unsigned)item % tiles_per_ch) * TO * D - OFF;  // even
        const float2 *xc = a.x + c * a.n_in;
        if (g_base >= 0 && g_base + SPAN <= (long long)a.n_in) {
            // interior tile (all but the first and last of a channel): no bounds checks
            const float4 *src = reinterpret_cast<const float4 *>(xc + g_base) + tid;
#pragma unroll
            for (int ib = 0; ib < NB; ++ib) {
                const int it = b0 + ib;
                if (it < NIT && ((it + 1) * NT <= SPAN / 2 || tid + it * NT < SPAN / 2)) raw[ib] = ldg_stream(src + it * NT);
            }
        } else {
            const float2 *hc = a.hist_in + c * H;
#pragma unroll
            for (int ib = 0; ib < NB; ++ib) {
                const int it = b0 + ib;
                if (it >= NIT) break;
                const int p = tid + it * NT;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (2 * p < SPAN) {
                    const long long g = g_base + 2 * p;
                    if (g >= 0) {
                        if (g + 1 < (long long)a.n_in) v = ldg_stream(reinterpret_cast<const float4 *>(xc + g));
                    } else if (H + g >= 0) {
                        v = *reinterpret_cast<const float4 *>(hc + (H + g));
                    }
                }
                raw[ib] = v;
            }
        }
    };
    // phasor table of an item: slots 0..NIT-1 = e^{j (2 NT it) dphi}, slot 63 = e^{j dphi},
    // slot 62 = e^{j (phi0 + g_base dphi)} (the item's first staged sample)
    auto fill_ftab = [&](unsigned long long item, int buf) {
        if (MIX && (tid < NIT || tid >= 62) && tid < 64) {
            const size_t c = (size_t)((unsigned)item / tiles_per_ch);
            const double dphi = a.dphase[c];
            const long long g_base = (long long)((unsigned)item % tiles_per_ch) * TO * D - OFF;
            const double th = tid == 63 ? dphi : (tid == 62 ? fma((double)g_base, dphi, a.phase_in[c]) : (double)(tid * 2 * NT) * dphi);
            ftab[buf * 64 + tid] = phase_rotation(th);
        }
    };

    // contiguous run of items per CTA: consecutive tiles of the same channel (phasors reused, DRAM pages too)
    const unsigned long long per = (nitems + gridDim.x - 1) / gridDim.x;
    unsigned long long item = (unsigned long long)blockIdx.x * per;
    const unsigned long long item_end = item + per < nitems ? item + per : nitems;
    if (item >= item_end) return;
    if (PF) issue_loads(item, 0);
    fill_ftab(item, 0);
    __syncthreads();
    size_t t_ch = ~(size_t)0;             // channel for which e_thr is valid
    float2 e_thr = make_float2(1.f, 0.f);  // e^{j (2 tid) dphi}
    for (int cur = 0; item < item_end; ++item, cur ^= 1) {
        const size_t c = (size_t)((unsigned)item / tiles_per_ch);
        const unsigned tile = (unsigned)((unsigned)item % tiles_per_ch);
        const long long m0 = (long long)tile * TO;
        const float2 *xc = a.x + c * a.n_in;
        const float2 *hc = a.hist_in + c * H;

        // ---- A: (mix and) store the span; pair p = samples (2p, 2p+1) of the buffer
        {
            float2 e_t = make_float2(1.f, 0.f), e_step = make_float2(1.f, 0.f);
            if (MIX) {
                if (t_ch != c) {
                    e_thr = phase_rotation((double)(2 * tid) * a.dphase[c]);
                    t_ch = c;
                }
                e_t = cmul(ftab[cur * 64 + 62], e_thr);
                e_step = ftab[cur * 64 + 63];
            }
            const float2 e_t_rr = rot90(e_t), e_step_rr = rot90(e_step);
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                if (!PF && it % NB == 0) issue_loads(item, it);
                const int p = tid + it * NT;
                if ((it + 1) * NT <= SPAN / 2 || p < SPAN / 2) {
                    float4 v = raw[it % NB];
                    if (MIX) {
                        const float2 r0 = cmul_p(ftab[cur * 64 + it], e_t, e_t_rr);
                        const float2 r1 = cmul_p(r0, e_step, e_step_rr);
                        const float2 s0 = cmul_p(make_float2(v.x, v.y), r0, rot90(r0));
                        const float2 s1 = cmul_p(make_float2(v.z, v.w), r1, rot90(r1));
                        v = make_float4(s0.x, s0.y, s1.x, s1.y);
                    }
                    const int i = 2 * p;
                    *reinterpret_cast<float4 *>(xs + (i / RD) * PITCH + (i % RD) * 8) = v;
                }
            }
        }
        __syncthreads();  // S1: span staged

        // ---- B: next item's loads go out now and land while this item is filtered
        const unsigned long long nxt = item + 1;
        if (nxt < item_end) {
            if (PF) issue_loads(nxt, 0);
            fill_ftab(nxt, cur ^ 1);
        }

        // ---- carried state for the next call (last tile of the channel)
        if (tile == tiles_per_ch - 1 && a.hist_out != nullptr) {
            float2 *ho = a.hist_out + c * H;
            for (long long i = tid; i < H; i += NT) {
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

        // ---- FM only: y[m0 - 1] for the first discriminator step of the tile (one warp, 2 taps per lane)
        if (FM && tid < 32) {
            float2 y = make_float2(0.f, 0.f);
            if (m0 > 0) {
#pragma unroll
                for (int kk = 0; kk < KP / 32; ++kk) {
                    const int k = tid + 32 * kk;
                    const int i = OFF - D - k;  // >= 0 by construction of OFF
                    const float2 sv = *reinterpret_cast<const float2 *>(xs + (i / RD) * PITCH + (i % RD) * 8);
                    y = __ffma2_rn(sv, taps.t[k], y);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    y.x += __shfl_xor_sync(0xffffffffu, y.x, o);
                    y.y += __shfl_xor_sync(0xffffffffu, y.y, o);
                }
            } else {
                y = a.prev_in[c];
            }
            if (tid == 0) ys[0] = y;
        }

        // ---- C: decimated FIR: thread owns outputs o = R*tid + r; sample at relative index cc = r*D - k
        {
            const unsigned char *base = xs + (tid + OFF / RD) * PITCH;
            float2 acc[R][2];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);
            constexpr int C_LO = -(KP - 1) - ((KP - 1) & 1);  // even start (<= -(KP-1))
            constexpr int C_HI = (R - 1) * D;                  // last sample used
#pragma unroll
            for (int cc = C_LO; cc <= C_HI; cc += 2) {
                const int q = (cc >= 0) ? cc / RD : -((-cc + RD - 1) / RD);  // floor(cc / RD)
                const int off = cc - q * RD;
                const float4 w = *reinterpret_cast<const float4 *>(base + q * PITCH + off * 8);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int ch = cc + h;
                    const float2 sv = h ? make_float2(w.z, w.w) : make_float2(w.x, w.y);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int k = r * D - ch;
                        if (k >= 0 && k < KP) acc[r][k & 1] = __ffma2_rn(sv, taps.t[k], acc[r][k & 1]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float2 y = __fadd2_rn(acc[r][0], acc[r][1]);
                ys[1 + R * tid + r] = y;
                if (FM && m0 + R * tid + r == (long long)a.n_out - 1) a.prev_out[c] = y;
            }
        }
        __syncthreads();  // S2: outputs staged, span reads done, next phasor table visible

        // ---- D: coalesced output
        float *out_f = reinterpret_cast<float *>(a.out) + c * a.n_out;
        float2 *out_c = reinterpret_cast<float2 *>(a.out) + c * a.n_out;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int o = tid + j * NT;
            const long long m = m0 + o;
            if (m < (long long)a.n_out) {
                if (FM) out_f[m] = fm_angle(ys[o + 1], ys[o]);
                else out_c[m] = ys[o + 1];
            }
        }
    }
}

template <bool MIX, bool FM, int D, int KP, int R, int NT, bool PF, int MINB>
static int launch_chain2(const ChainArgs &args, const ChainTaps &taps, size_t channels, cudaStream_t s)
{
    constexpr int TO = NT * R, RD = R * D, OFF = ((KP + D + RD - 1) / RD) * RD;
    constexpr int SMEM = (NT + OFF / RD) * (RD * 8 + 16) + 128 * 8 + (TO + 1) * 8;
    auto kern = chain2_kernel<MIX, FM, D, KP, R, NT, PF, MINB>;
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const unsigned tiles = (unsigned)ceil_div(args.n_out, (size_t)TO);
    const unsigned long long nitems = (unsigned long long)tiles * channels;
    CB_REQUIRE(nitems < (1ull << 32), CB_ERR_UNSUPPORTED, "chain: more than 2^32 tiles in one call");  // 32-bit item arithmetic in the kernel
    int dev = 0, sms = 148, per_sm = 2;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, SMEM);
    if (per_sm < 1) per_sm = 1;
    const unsigned long long cap = (unsigned long long)sms * per_sm;
    const unsigned grid = (unsigned)(nitems < cap ? nitems : cap);
    kern<<<grid, NT, SMEM, s>>>(args, taps, tiles, nitems);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// ============================================================================
// v3 (real taps <= 64, compile-time D): TMA-staged, mixer folded into the taps.
//
//   y[m] = sum_k h[k] x[mD-k] e^{j phi(mD-k)} = e^{j phi(mD)} sum_k (h[k] e^{-j k dphi}) x[mD-k]
//
// so the RAW input span can go from HBM to shared memory by TMA (a few 2 KiB 1-D bulk copies per
// tile, stored linearly) with no register staging, no per-sample rotation pass and no barrier
// between staging and filtering.  (One bulk copy per padded thread-chunk was measured 3x slower:
// small UBLKCPs are issue-bound.)  R outputs per thread with R*D/2 odd makes the linear layout
// bank-conflict free for the 128-bit window loads.  The per-channel rotated
// taps h'[k] live in shared memory (64 complex values, recomputed when the CTA moves to another
// channel) and every output is rotated once by e^{j phi(mD)} = base(item) * e_thr(thread) * step(r).
// A ring of NSTAGE span buffers per CTA keeps (NSTAGE - 1) spans of HBM reads in flight per CTA.
// History chunks (g < 0) are bulk-copied from the channel's history buffer (hist_len >= OFF), so
// the first tile of a channel needs no special path; chunks past n_in are simply not copied.
// Numerics: identical algebra, different rounding points (rel-L2 ~2e-7 vs the sequential f32 form).
// ============================================================================
// U8: the input is the RTL-SDR byte stream of examples/fm_radio.rs (u8 I, u8 Q).  ConvertNode's
// (x - 127.5) / 127.5 (fm_radio.rs:84-87) is split by linearity: the span holds b - 127.5 (exact: a byte permute
// that splices the exponent of 2^23 above the byte, and two packed adds -- no table look-ups, no I2F) and the launch
// scales the taps by 1 / 127.5, so the outputs differ from convert-then-filter only in rounding (~1e-7 relative)
// while HBM carries 2 instead of 8 bytes per input sample.  The carried history stays ConvertNode's exact values.
template <bool MIX, bool FM, bool U8, int D, int R, int NT, int NSTAGE, int MINB>
__global__ void __launch_bounds__(NT + 32, MINB)
chain3_kernel(const __grid_constant__ ChainArgs a, const __grid_constant__ ChainTaps taps, const unsigned tiles_per_ch,
              const unsigned long long nitems)
{
    constexpr int KP = 64;
    constexpr int TO = NT * R;
    constexpr int RD = R * D;
    constexpr int OFF = U8 ? (KP + D + 7) / 8 * 8 : (KP + D + 1) / 2 * 2;  // halo samples in front of the tile
    constexpr int SPAN = TO * D + OFF;
    constexpr int STAGE = (SPAN * 8 + 127) / 128 * 128;  // f32 span
    constexpr int RAWSTAGE = (SPAN * 2 + 127) / 128 * 128;  // U8: raw byte span
    constexpr int PIECE = U8 ? 1024 : 256;              // samples per bulk copy (2 KiB)
    static_assert(!U8 || (TO * D) % 8 == 0, "u8 tiles must start on 16-byte boundaries");
    // the span is stored LINEARLY; thread t reads 128-bit words at a lane stride of RD*8 bytes, which is
    // bank-conflict free exactly when RD/2 is odd (stride = odd multiple of 16 bytes)
    static_assert(RD % 2 == 0 && (RD / 2) % 2 == 1 && NT >= 64 + R + 1, "");

    extern __shared__ __align__(128) unsigned char c3sm[];
    unsigned char *stages = c3sm;                                     // ring of raw spans (f32, or bytes when U8)
    unsigned char *span8 = c3sm + NSTAGE * RAWSTAGE;                  // U8: the one converted f32 span
    constexpr int RING = U8 ? NSTAGE * RAWSTAGE + STAGE : NSTAGE * STAGE;
    float2 *tsm = reinterpret_cast<float2 *>(c3sm + RING);            // tsm[k + 1] = h'[k], tsm[0] = tsm[65] = 0
    float2 *cst = tsm + KP + 2;                                       // e^{j r D dphi} (r < R), then e^{-j D dphi}
    float2 *basep = cst + 8;                                          // per stage: e^{j (phi0 + m0 D dphi)}
    float2 *ys = basep + 8;                                           // 2 x (TO + 1) outputs
    __shared__ __align__(8) uint64_t full[NSTAGE];

    // NT consumer threads (warps 0 .. NT/32-1) + one producer warp (TMA issue, per-item base phasor,
    // and the FM look-back output), so that no filtering warp carries extra work into the barrier
    const int tid = threadIdx.x, lane = tid & 31;
    const bool producer = tid >= NT;
    const long long H = a.hist_len;

    const unsigned long long per = (nitems + gridDim.x - 1) / gridDim.x;
    const unsigned long long item0 = (unsigned long long)blockIdx.x * per;
    const unsigned long long item_end = item0 + per < nitems ? item0 + per : nitems;
    if (item0 >= item_end) return;

    // stage memory starts out finite (chunks past the end of a batch are never copied)
    for (int i = tid; i < (U8 ? NSTAGE * RAWSTAGE + STAGE : NSTAGE * STAGE) / 16; i += NT + 32)
        reinterpret_cast<uint4 *>(stages)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) mbar_init(&full[i], 1);
        fence_mbar_init();
    }
    fence_proxy_async();
    __syncthreads();

    // warp 0: bulk-copy the span of `item` into stage `st`
    auto produce = [&](unsigned long long item, int st) {
        const size_t c = (size_t)((unsigned)item / tiles_per_ch);
        const long long tile = (long long)((unsigned)item % tiles_per_ch);
        const long long g_base = tile * TO * D - OFF;
        long long g_hi = g_base + SPAN;
        if (g_hi > (long long)a.n_in) g_hi = (long long)a.n_in;
        const long long g_lo = g_base < 0 ? 0 : g_base;
        if (lane == 0) {
            if (MIX) basep[st] = phase_rotation(fma((double)(tile * TO * D), a.dphase[c], a.phase_in[c]));
            mbar_arrive_expect_tx(&full[st], (uint32_t)(U8 ? (g_hi - g_lo) * 2 : (g_hi - g_base) * 8));
        }
        __syncwarp();
        if (U8) {  // bytes only; the f32 history of the first tile is fetched by the conversion pass
            const unsigned char *xc = a.x8 + 2 * c * a.n_in;
            unsigned char *dst = stages + st * RAWSTAGE;
            for (long long g = g_lo + (long long)lane * PIECE; g < g_hi; g += 32 * PIECE) {
                const long long n = g_hi - g < PIECE ? g_hi - g : PIECE;
                tma_load_1d(dst + (g - g_base) * 2, xc + 2 * g, (uint32_t)(n * 2), &full[st]);
            }
        } else {
            const float2 *xc = a.x + c * a.n_in;
            const float2 *hc = a.hist_in + c * H + H;
            unsigned char *dst = stages + st * STAGE;
            if (g_base < 0 && lane == 31) tma_load_1d(dst, hc + g_base, (uint32_t)(-g_base * 8), &full[st]);  // history
            for (long long g = g_lo + (long long)lane * PIECE; g < g_hi; g += 32 * PIECE) {
                const long long n = g_hi - g < PIECE ? g_hi - g : PIECE;
                tma_load_1d(dst + (g - g_base) * 8, xc + g, (uint32_t)(n * 8), &full[st]);
            }
        }
    };
    if (producer) {
        for (int i = 0; i < NSTAGE; ++i)
            if (item0 + i < item_end) produce(item0 + i, i);
    }

    size_t cur_c = ~(size_t)0;
    float2 e_thr = make_float2(1.f, 0.f);
    unsigned long long li = 0;
    for (unsigned long long item = item0; item < item_end; ++item, ++li) {
        const int st = (int)(li % NSTAGE);
        const uint32_t parity = (uint32_t)((li / NSTAGE) & 1);
        // output staging: two buffers so that the stores of one tile overlap the filter of the next; with byte input the
        // conversion barrier (every thread arrives after its own output reads of the previous tile) already separates
        // them, so one buffer is enough -- which is what lets a fourth CTA fit per SM
        const int yb = U8 ? 0 : (int)(li & 1) * (TO + 1);
        const size_t c = (size_t)((unsigned)item / tiles_per_ch);
        const unsigned tile = (unsigned)((unsigned)item % tiles_per_ch);
        const long long m0 = (long long)tile * TO;

        if (MIX && c != cur_c) {  // new channel: rotated taps and the per-channel phasors
            const double dphi = a.dphase[c];
            if (tid < KP) {
                const float2 w = phase_rotation(-(double)tid * dphi);
                const float h = taps.t[tid].x;
                tsm[tid + 1] = make_float2(h * w.x, h * w.y);
            } else if (tid < KP + R) {
                cst[tid - KP] = phase_rotation((double)((tid - KP) * D) * dphi);
            } else if (tid == KP + R) {
                cst[R] = phase_rotation(-(double)D * dphi);
                tsm[0] = tsm[KP + 1] = make_float2(0.f, 0.f);
            }
            if (!producer) e_thr = phase_rotation((double)(RD * tid) * dphi);
            cur_c = c;
            __syncthreads();
        }

        mbar_wait_long(&full[st], parity);
        const unsigned char *sbase = U8 ? span8 : stages + st * STAGE;
        if (U8) {
            // raw bytes -> f32 span (pairs of samples: one 32-bit word in, one 128-bit word out)
            if (!producer) {
                const long long g_base = (long long)tile * TO * D - OFF;
                const uint32_t *raw = reinterpret_cast<const uint32_t *>(stages + st * RAWSTAGE);
                const float2 *hc = a.hist_in + c * H + H;
                constexpr int NCV = (SPAN / 2 + NT - 1) / NT;
#pragma unroll 8
                for (int it = 0; it < NCV; ++it) {  // compile-time trip count: the loads of several words overlap
                    const int p = tid + it * NT;
                    const long long g = g_base + 2 * p;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (g < 0) {  // carried history is in the converted domain: back to b - 127.5
                        v = *reinterpret_cast<const float4 *>(hc + g);
                        v = make_float4(v.x * 127.5f, v.y * 127.5f, v.z * 127.5f, v.w * 127.5f);
                    } else if (g < (long long)a.n_in && p < SPAN / 2) {
                        // byte b -> 2^23 + b (exponent byte spliced in by PRMT) -> b -> b - 127.5, every step exact
                        const uint32_t w = raw[p];
                        const float2 big = make_float2(-8388608.f, -8388608.f), half = make_float2(-127.5f, -127.5f);
                        const float2 s0 = make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440)),
                                                      __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7441)));
                        const float2 s1 = make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7442)),
                                                      __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7443)));
                        const float2 c0 = __fadd2_rn(__fadd2_rn(s0, big), half), c1 = __fadd2_rn(__fadd2_rn(s1, big), half);
                        v = make_float4(c0.x, c0.y, c1.x, c1.y);
                    }
                    if (p < SPAN / 2) reinterpret_cast<float4 *>(span8)[p] = v;
                }
            }
            asm volatile("bar.sync 2, %0;" ::"r"(NT + 32) : "memory");  // span converted (consumers + producer warp)
        }

        // ---- FM only: y[m0 - 1] for the first discriminator step of the tile (warp 0, 2 taps per lane)
        if (FM && producer) {
            float2 y = make_float2(0.f, 0.f);
            if (m0 > 0) {
                float2 yA = make_float2(0.f, 0.f), yB = make_float2(0.f, 0.f);
#pragma unroll
                for (int kk = 0; kk < KP / 32; ++kk) {
                    const int k = lane + 32 * kk;
                    const int i = OFF - D - k;  // >= 0 by construction of OFF
                    const float2 sv = *reinterpret_cast<const float2 *>(sbase + i * 8);
                    if (MIX) {
                        const float2 h = tsm[k + 1];
                        yA = __ffma2_rn(sv, make_float2(h.x, h.x), yA);
                        yB = __ffma2_rn(sv, make_float2(h.y, h.y), yB);
                    } else {
                        yA = __ffma2_rn(sv, taps.t[k], yA);
                    }
                }
                y = MIX ? make_float2(yA.x - yB.y, yA.y + yB.x) : yA;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    y.x += __shfl_xor_sync(0xffffffffu, y.x, o);
                    y.y += __shfl_xor_sync(0xffffffffu, y.y, o);
                }
                if (MIX) y = cmul(y, cmul(basep[st], cst[R]));
            } else {
                y = a.prev_in[c];
            }
            if (lane == 0) ys[yb] = y;
        }

        // ---- decimated FIR: thread owns outputs o = R*tid + r; sample at relative index cc = r*D - k
        if (!producer) {
            const unsigned char *base = sbase + (tid * RD + OFF) * 8;
            float2 accA[R][2], accB[R][2];
#pragma unroll
            for (int r = 0; r < R; ++r) accA[r][0] = accA[r][1] = accB[r][0] = accB[r][1] = make_float2(0.f, 0.f);
            constexpr int C_LO = -(KP - 1) - ((KP - 1) & 1);  // even start (<= -(KP-1))
            constexpr int C_HI = (R - 1) * D;                  // last sample used
            float2 tq[KP + 2];                                 // tq[k + 1] = h'[k]; registers, loaded at first use
#pragma unroll
            for (int cc = C_LO; cc <= C_HI; cc += 2) {
                const float4 w = *reinterpret_cast<const float4 *>(base + cc * 8);
                if (MIX && cc <= 0) {  // taps k = -cc-1, -cc are used for the first time (by r = 0) now
                    const float4 t2 = *reinterpret_cast<const float4 *>(tsm + (-cc));
                    tq[-cc] = make_float2(t2.x, t2.y);
                    tq[-cc + 1] = make_float2(t2.z, t2.w);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int ch = cc + h;
                    const float2 sv = h ? make_float2(w.z, w.w) : make_float2(w.x, w.y);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int k = r * D - ch;
                        if (k >= 0 && k < KP) {
                            if (MIX) {
                                accA[r][k & 1] = __ffma2_rn(sv, make_float2(tq[k + 1].x, tq[k + 1].x), accA[r][k & 1]);
                                accB[r][k & 1] = __ffma2_rn(sv, make_float2(tq[k + 1].y, tq[k + 1].y), accB[r][k & 1]);
                            } else {
                                accA[r][k & 1] = __ffma2_rn(sv, taps.t[k], accA[r][k & 1]);
                            }
                        }
                    }
                }
            }
            float2 ph = make_float2(1.f, 0.f);
            if (MIX) ph = cmul(basep[st], e_thr);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float2 y = __fadd2_rn(accA[r][0], accA[r][1]);
                if (MIX) {
                    const float2 b = __fadd2_rn(accB[r][0], accB[r][1]);
                    y = cmul(make_float2(y.x - b.y, y.y + b.x), cmul(ph, cst[r]));
                }
                ys[yb + 1 + R * tid + r] = y;
                if (FM && m0 + R * tid + r == (long long)a.n_out - 1) a.prev_out[c] = y;
            }
        }
        __syncthreads();  // outputs staged; every thread is done with stage st

        if (producer) {
            if (item + NSTAGE < item_end) produce(item + NSTAGE, st);
            continue;
        }

        // ---- carried state for the next call (last tile of the channel)
        if (tile == tiles_per_ch - 1 && a.hist_out != nullptr) {
            const float2 *hc = a.hist_in + c * H;
            float2 *ho = a.hist_out + c * H;
            for (long long i = tid; i < H; i += NT) {
                const long long g = (long long)a.n_in - H + i;
                float2 v;
                if (g < 0) {
                    v = hc[H + g];
                } else if (U8) {
                    const unsigned char *b = a.x8 + 2 * (c * a.n_in + g);  // the state keeps ConvertNode's exact values
                    v = make_float2(__fdiv_rn(__fsub_rn((float)b[0], 127.5f), 127.5f), __fdiv_rn(__fsub_rn((float)b[1], 127.5f), 127.5f));
                } else {
                    v = a.x[c * a.n_in + g];
                }
                ho[i] = v;
            }
            if (MIX && tid == 0) {
                const double twopi = 6.283185307179586232;
                double phs = fma((double)a.n_in, a.dphase[c], a.phase_in[c]);
                phs -= twopi * floor(phs / twopi);
                a.phase_out[c] = phs;
            }
        }

        // ---- coalesced output
        float *out_f = reinterpret_cast<float *>(a.out) + c * a.n_out;
        float2 *out_c = reinterpret_cast<float2 *>(a.out) + c * a.n_out;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int o = tid + j * NT;
            const long long m = m0 + o;
            if (m < (long long)a.n_out) {
                if (FM) out_f[m] = fm_angle_fast(ys[yb + o + 1], ys[yb + o]);
                else out_c[m] = ys[yb + o + 1];
            }
        }
    }
}

template <bool MIX, bool FM, bool U8, int D, int R, int NT, int NSTAGE, int MINB>
static int launch_chain3(const ChainArgs &args, const ChainTaps &taps, size_t channels, cudaStream_t s)
{
    constexpr int TO = NT * R, OFF = U8 ? (64 + D + 7) / 8 * 8 : (64 + D + 1) / 2 * 2;
    constexpr int F32SPAN = ((TO * D + OFF) * 8 + 127) / 128 * 128, RAWSPAN = ((TO * D + OFF) * 2 + 127) / 128 * 128;
    constexpr int SMEM = (U8 ? NSTAGE * RAWSPAN + F32SPAN : NSTAGE * F32SPAN) + (66 + 8 + 8) * 8 + (U8 ? 1 : 2) * (TO + 1) * 8;
    auto kern = chain3_kernel<MIX, FM, U8, D, R, NT, NSTAGE, MINB>;
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    ChainTaps scaled = taps;
    if (U8)  // the span holds b - 127.5: ConvertNode's division moves into the taps
        for (int k = 0; k < CHAIN_MAX_TAP_SLOTS; ++k)
            scaled.t[k] = make_float2((float)((double)taps.t[k].x / 127.5), (float)((double)taps.t[k].y / 127.5));
    const unsigned tiles = (unsigned)ceil_div(args.n_out, (size_t)TO);
    const unsigned long long nitems = (unsigned long long)tiles * channels;
    CB_REQUIRE(nitems < (1ull << 32), CB_ERR_UNSUPPORTED, "chain: more than 2^32 tiles in one call");  // 32-bit item arithmetic in the kernel
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT + 32, SMEM);
    if (per_sm < 1) per_sm = 1;
    const unsigned long long cap = (unsigned long long)sms * per_sm;
    const unsigned grid = (unsigned)(nitems < cap ? nitems : cap);
    kern<<<grid, NT + 32, SMEM, s>>>(args, scaled, tiles, nitems);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

template <int D, int R, int NT, int NSTAGE, int MINB>
static int launch_chain3_shape(const ChainArgs &args, const ChainTaps &taps, bool mix, bool fm, size_t channels, cudaStream_t s)
{
    if (args.x8 != nullptr) {
        if (mix && fm) return launch_chain3<true, true, true, D, R, NT, NSTAGE, MINB>(args, taps, channels, s);
        if (mix) return launch_chain3<true, false, true, D, R, NT, NSTAGE, MINB>(args, taps, channels, s);
        if (fm) return launch_chain3<false, true, true, D, R, NT, NSTAGE, MINB>(args, taps, channels, s);
        return launch_chain3<false, false, true, D, R, NT, NSTAGE, MINB>(args, taps, channels, s);
    }
    if (mix && fm) return launch_chain3<true, true, false, D, R, NT, NSTAGE, MINB>(args, taps, channels, s);
    if (mix) return launch_chain3<true, false, false, D, R, NT, NSTAGE, MINB>(args, taps, channels, s);
    if (fm) return launch_chain3<false, true, false, D, R, NT, NSTAGE, MINB>(args, taps, channels, s);
    return launch_chain3<false, false, false, D, R, NT, NSTAGE, MINB>(args, taps, channels, s);
}

// tile shape of the v2 kernel: R outputs per thread, NT threads, register prefetch or not
template <int D, int R, int NT, bool PF, int MINB>
static int launch_chain2_shape(const ChainArgs &args, const ChainTaps &taps, bool mix, bool fm, size_t channels, cudaStream_t s)
{
    if (mix && fm) return launch_chain2<true, true, D, 64, R, NT, PF, MINB>(args, taps, channels, s);
    if (mix) return launch_chain2<true, false, D, 64, R, NT, PF, MINB>(args, taps, channels, s);
    if (fm) return launch_chain2<false, true, D, 64, R, NT, PF, MINB>(args, taps, channels, s);
    return launch_chain2<false, false, D, 64, R, NT, PF, MINB>(args, taps, channels, s);
}

// Kernel choice for real taps <= 64 and compile-time D.  COMMS_B200_CHAIN_PATH = auto (default) | v2 | v3
// selects the family; v3 (TMA-staged, mixer folded into the taps) needs R*D/2 odd for its linear
// conflict-free layout, so it is instantiated for D = 10 (R = 3) and D = 5 (R = 6).
template <int D, int R>
static int launch_chain2_d(const ChainArgs &args, const ChainTaps &taps, bool mix, bool fm, size_t channels, cudaStream_t s)
{
    static const int path = [] {
        const char *e = getenv("COMMS_B200_CHAIN_PATH");
        return (e && strcmp(e, "v2") == 0) ? 2 : ((e && strcmp(e, "v3") == 0) ? 3 : 0);
    }();
    if (args.x8 != nullptr && args.hist_len >= 128) {  // byte input: one f32 span + a small raw ring -> more CTAs per SM
        if constexpr (D == 10) return launch_chain3_shape<D, 3, 128, 2, 4>(args, taps, mix, fm, channels, s);
        if constexpr (D == 5) {
            // without the mixer (fm_radio as shipped) the kernel fits 4 CTAs per SM; the rotated-tap variants need more
            // than the 102 registers that would leave them and stay at 3
            if (!mix)
                return fm ? launch_chain3<false, true, true, D, 6, 128, 2, 4>(args, taps, channels, s)
                          : launch_chain3<false, false, true, D, 6, 128, 2, 4>(args, taps, channels, s);
            return launch_chain3_shape<D, 6, 128, 2, 3>(args, taps, mix, fm, channels, s);
        }
    }
    if (path != 2 && args.hist_len >= 128) {
        if constexpr (D == 10) return launch_chain3_shape<D, 3, 128, 2, 3>(args, taps, mix, fm, channels, s);
        if constexpr (D == 5) return launch_chain3_shape<D, 6, 128, 2, 2>(args, taps, mix, fm, channels, s);
    }
    if (args.x8 != nullptr) {
        set_error("chain: no fused u8 kernel for this shape");
        return CB_ERR_UNSUPPORTED;
    }
    return launch_chain2_shape<D, R, 256, true, 2>(args, taps, mix, fm, channels, s);
}

bool chain_fuses_u8(const ChainArgs &args, bool cplx)
{
    return !cplx && args.ntaps <= 64 && args.hist_len >= 128 && args.n_in % 8 == 0 && (args.decim == 10 || args.decim == 5) &&
           (reinterpret_cast<uintptr_t>(args.x8) & 15) == 0;
}

int launch_chain(const ChainArgs &args, const ChainTaps &taps, bool mix, bool fm, bool cplx, size_t channels,
                 cudaStream_t s)
{
    if (args.n_in == 0 || channels == 0) return CB_OK;
    if (args.x8 != nullptr) {
        if (!chain_fuses_u8(args, cplx)) {
            set_error("chain: no fused u8 kernel for this shape");
            return CB_ERR_UNSUPPORTED;
        }
        if (!mix && chain_tc_applicable(args, channels)) return launch_chain_tc(args, taps, fm, channels, s);
        return args.decim == 10 ? launch_chain2_d<10, 2>(args, taps, mix, fm, channels, s)
                                : launch_chain2_d<5, 4>(args, taps, mix, fm, channels, s);
    }
    // v2: real taps <= 64, even batch length and 16-byte aligned input (128-bit staging loads)
    if (!cplx && args.ntaps <= 64 && args.hist_len % 2 == 0 && args.n_in % 2 == 0 &&
        (reinterpret_cast<uintptr_t>(args.x) & 15) == 0) {
        switch (args.decim) {
        case 10: return launch_chain2_d<10, 2>(args, taps, mix, fm, channels, s);
        case 5: return launch_chain2_d<5, 4>(args, taps, mix, fm, channels, s);
        case 4: return launch_chain2_d<4, 4>(args, taps, mix, fm, channels, s);
        case 2: return launch_chain2_d<2, 4>(args, taps, mix, fm, channels, s);
        case 8: return launch_chain2_d<8, 2>(args, taps, mix, fm, channels, s);
        default: break;
        }
    }
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
