// K2-TC: the fm_radio front end on raw RTL-SDR bytes -- ConvertNode -> BatchFirNode (<= 64 real taps) -> DecimateNode(5)
// [-> FMDemodNode] (examples/fm_radio.rs:84-97, 144-160) -- as a polyphase Toeplitz GEMM on the 5th-generation tensor
// cores (tcgen05.mma, accumulators in TMEM), sm_100a only.
//
// Why.  On the CUDA cores this chain costs ~24 instructions per input sample (12.6 packed MACs, the discriminator, the
// byte conversion, window loads): chain3_kernel runs it at 0.47 Tsamples/s = 20 % of the 2.8 B/sample HBM roof and
// cannot get past ~1 Tsample/s at perfect issue.  Here the MACs leave the issue slots altogether.
//
// Formulation.  y[m] = sum_k h[k] x[5 m - k] splits over the five input phases x_d[i] = x[5 i + d]:
//     y[m] = sum_j g_0[j] x_0[m - j] + sum_{d=1..4} sum_j g_d[j] x_d[m - j],   g_0[j] = h[5 j],  g_d[j] = h[5 (j - 1) + 5 - d] (j >= 1)
// i.e. five plain FIRs of <= 14 taps on the DECIMATED grid, summed.  Each is the block-Toeplitz GEMM of fir_tc_kernel.cu
// with 16 outputs per row: the phase stream S_d, stored once in shared memory as interleaved (re, im) fp16 with a row
// pitch of 16 samples = 64 bytes (K-major, 64-byte swizzle), IS the Toeplitz operand -- K-block kb of row r is row r + kb
// of the same buffer -- and all five accumulate into one TMEM tile:
//     D[r][n] = sum_d sum_k S~_d[32 r + k] * B_d[n][k],  k in [0, 64),  n = 2 q + c (output q of the row, re / im),
//     B_d[2 q + c][2 kk + cc] = (c == cc) g_d[q + 16 - kk]   (real taps act on re and im alike)
// 20 MMAs of M128 x N64 x K16 per tile of 2048 outputs (10240 input samples).
//
// Precision.  A byte is exact in fp16, so the stream needs no lo term, no block scale and no max reduction: the converter
// stores S = b - 128 (PRMT splices the byte into the mantissa of 1024.0, one packed subtract) and the missing 0.5 of
// ConvertNode's (b - 127.5) / 127.5 is the constant 0.5 sum(h) / 127.5 added in the epilogue.  The taps are h / 127.5,
// block-scaled by a power of two and split into fp16 hi + lo (columns 0-31 / 32-63 of the accumulator): f32-level
// accuracy (rel-L2 ~1e-7 against convert-then-filter).
//
// Roles (20 warps, one persistent CTA per SM; items = (channel, tile) round-robin):
//   warp 11    TMA producer: the tile's raw bytes incl. 80 samples of halo, 2 KiB bulk copies into a ring of 2 stages
//   warps 0-4 / 5-9  two converter groups (even / odd items): bytes -> fp16 phase streams, 10 x STS.128 per 40 samples;
//              warp 0 of a group also evaluates y[first - 1] for the tile's first discriminator step (2 taps per lane)
//   warp 10    one elected lane issues the 20 tcgen05.mma per tile, commits a_empty / t_full
//   warps 12-19 epilogue, two per TMEM sub-partition (lane = row; warps 12-15 take outputs 0-7 of the row, 16-19 outputs
//              8-15): tcgen05.ld -> add halves, unscale, + dc -> FM discriminator (the previous output is in the thread,
//              re-read from TMEM, or the last output of the row below through shared memory / the look-back value)
//              -> 32 contiguous bytes per lane.  The discriminator is ~35 instructions per output: with four epilogue
//              warps it bounded the kernel at 3.1 k cycles per tile
// Algorithmic HBM traffic: 2 B read per input sample + 4 B (FM) or 8 B written per 5: 2.8 B/sample with the FM tail.
#include <cuda_fp16.h>

#include <vector>

#include "chain_kernels.cuh"
#include "misc_kernels.cuh"

namespace cb {

namespace ctc {

constexpr int D = 5;                 // decimation
constexpr int RS = 16;               // outputs per row
constexpr int ROWS = 128;            // MMA M
constexpr int TO = ROWS * RS;        // 2048 outputs per tile
constexpr int TI = TO * D;           // 10240 input samples per tile
constexpr int HALO = RS;             // history per phase stream (decimated grid): one row >= 14 taps
constexpr int HIN = HALO * D;        // = 80 input samples of halo
constexpr int NCW = 5;               // warps per converter group: 258 units of 40 samples per tile = at most 2 per thread
constexpr int NCONV = 32 * NCW;      // converter threads per group
constexpr int W_MMA = 2 * NCW, W_TMA = 2 * NCW + 1;
constexpr int W_EPI = 12;            // first of the 8 epilogue warps (two per TMEM sub-partition: outputs 0-7 / 8-15 of a row)
constexpr int NEPI = 256;            // epilogue threads
constexpr int NTHREADS = 32 * (W_EPI + 8);
static_assert(W_EPI % 4 == 0 && W_EPI > W_TMA, "epilogue warp w owns TMEM sub-partition w % 4");
constexpr int A_PHASE = 17 * 512;    // bytes per phase stream: (128 + 1) rows x 64 B, rounded to the 512-byte swizzle atom
constexpr int A_STAGE = D * A_PHASE;
constexpr int B_PHASE = 2 * 4096;    // per phase: 2 K-blocks x (64 rows x 64 B)
constexpr int B_BYTES = D * B_PHASE;
constexpr int RAWB = ((TI + HIN) * 2 + 127) / 128 * 128;
constexpr int NUNIT = (TI + HIN) / 40;  // converter work units of 40 samples (8 per phase)
constexpr int NRAW = 4;              // raw byte tiles in flight: with two, a stage's cycle (HBM latency + conversion) bounded the kernel
constexpr int OSTAGE = TO * 4;       // FM outputs of one tile, staged when the channel's output row is not 16-byte aligned
constexpr int SMEM = B_BYTES + 2 * A_STAGE + NRAW * RAWB + 2 * OSTAGE + 1024;
static_assert(SMEM <= 227 * 1024, "shared memory budget");
static_assert((TI + HIN) % 40 == 0, "whole units");

struct Args {
    const unsigned char *x8;   // channels x n_in (u8 I, u8 Q)
    void *out;                 // channels x n_out floats (FM) or complex
    const float2 *hist_in;     // channels x hist_len ConvertNode values (f32), chronological
    float2 *hist_out;
    const float2 *prev_in;     // per channel, last filter output (FM)
    float2 *prev_out;
    const uint4 *bimg;         // prepacked tap image, B_BYTES
    unsigned long long n_in, n_out;
    unsigned hist_len, ntaps, tiles_per_ch;
    unsigned long long nitems;
    float tap_inv_scale;       // undoes the taps' block scale
    float dc;                  // 0.5 * sum(h) / 127.5
    float2 *seam;              // FM: per item, its first and its last filter output (2 * nitems entries); the first
                               // discriminator output of every tile is written by chain_tc_seam_kernel afterwards
#ifdef CB_CTC_TIMELINE
    unsigned long long *dbg;   // scripts/ctc_timeline.cu: clock64 stamps of two steady-state items per CTA
#endif
};

#ifdef CB_CTC_TIMELINE
#define CTC_STAMP(cond, k)                                                                              \
    do {                                                                                                \
        if ((cond) && (it == 20 || it == 21)) a.dbg[((size_t)blockIdx.x * 2 + (it & 1)) * 24 + (k)] = clock64(); \
    } while (0)
#else
#define CTC_STAMP(cond, k)
#endif

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
template <bool ACC>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t *r)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t *r)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_ld2(uint32_t taddr, uint32_t *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 64-byte swizzle shared-memory matrix descriptor: 8-row groups 512 bytes apart.  The swizzle XOR acts on
// absolute shared-memory address bits, so a start address shifted by whole 64-byte rows (the Toeplitz trick) or by 32-byte
// K steps needs no base-offset field.
__device__ __forceinline__ uint64_t kdesc64(uint32_t saddr)
{
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;  // SWIZZLE_64B
    return d;
}
__host__ __device__ __forceinline__ uint32_t swz64(uint32_t o) { return o ^ (((o >> 7) & 3u) << 4); }

// half2 (I - 128, Q - 128) of sample `odd` of a 32-bit word holding two (u8 I, u8 Q) samples: PRMT splices each byte
// into the mantissa of fp16 1024.0 (0x6400, ulp 1), one packed subtract of 1152 leaves b - 128, both steps exact
__device__ __forceinline__ uint32_t bytes_to_half2(uint32_t w, int odd)
{
    const uint32_t spliced = __byte_perm(w, 0x64646464u, odd ? 0x4342 : 0x4140);
    const __half2 h = *reinterpret_cast<const __half2 *>(&spliced);
    const __half2 v = __hsub2(h, __floats2half2_rn(1152.f, 1152.f));
    return *reinterpret_cast<const uint32_t *>(&v);
}

// (channel, tile) of the items first, first + stride, ... without a 64-bit division per item (each costs ~100 instructions
// per warp, and every role of the CTA walks the item list)
struct ItemWalk {
    unsigned c, tile, dc, dt, tiles;
    __device__ __forceinline__ ItemWalk(unsigned first, unsigned stride, unsigned tiles_per_ch)
        : c(first / tiles_per_ch), tile(first % tiles_per_ch), dc(stride / tiles_per_ch), dt(stride % tiles_per_ch), tiles(tiles_per_ch)
    {
    }
    __device__ __forceinline__ void next()
    {
        c += dc;
        tile += dt;
        if (tile >= tiles) {
            tile -= tiles;
            ++c;
        }
    }
};

template <bool FM>
__global__ void __launch_bounds__(NTHREADS, 1) chain_tc_kernel(const __grid_constant__ Args a, const __grid_constant__ ChainTaps taps)
{
    constexpr uint32_t IDESC = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);  // f16 x f16 -> f32, M128 N64

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *sB = smem;                     // tap image
    unsigned char *sA = sB + B_BYTES;             // 2 stages x 5 phase streams
    unsigned char *sRaw = sA + 2 * A_STAGE;       // NRAW raw byte tiles (TMA destination)
    float *sOut = reinterpret_cast<float *>(sRaw + NRAW * RAWB);  // 2 x TO floats
    __shared__ __align__(8) uint64_t raw_full[NRAW], raw_empty[NRAW], a_full[2], a_empty[2], t_full[2], t_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float2 ylast[2][ROWS];             // last output of every row of the tile (two tiles deep)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long H = a.hist_len;

    if (tid == 0) {
        for (int i = 0; i < NRAW; ++i) {
            mbar_init(&raw_full[i], 1);
            mbar_init(&raw_empty[i], NCONV);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], NCONV);
            mbar_init(&a_empty[i], 1);
            mbar_init(&t_full[i], 1);
            mbar_init(&t_empty[i], NEPI);
        }
        fence_mbar_init();
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < B_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4 *>(sB)[i] = a.bimg[i];
    // phase streams start out as S = -0.5 (a ConvertNode value of 0): the row past the tile that the last K-block of row
    // 127 reads is then harmless, and stays so (the converters never write beyond TO + HALO samples)
    for (int i = tid; i < 2 * A_STAGE / 4; i += NTHREADS) reinterpret_cast<uint32_t *>(sA)[i] = 0xB800B800u;
    for (int i = tid; i < NRAW * RAWB / 16; i += NTHREADS) reinterpret_cast<uint4 *>(sRaw)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == W_TMA) {
        // ------------------------------------------------------------------ TMA producer
        unsigned long long it = 0;
        ItemWalk w(blockIdx.x, gridDim.x, a.tiles_per_ch);
        for (unsigned long long item = blockIdx.x; item < a.nitems; item += gridDim.x, ++it, w.next()) {
            const int s = (int)(it % NRAW);
            const unsigned long long c = w.c;
            const long long tile = w.tile;
            const long long g0 = tile * TI - HIN;
            const long long g_lo = g0 < 0 ? 0 : g0;
            long long g_hi = g0 + TI + HIN;
            if (g_hi > (long long)a.n_in) g_hi = (long long)a.n_in;
            CTC_STAMP(lane == 0, 20);
            mbar_wait_long(&raw_empty[s], (uint32_t)(((it / NRAW) & 1) ^ 1));
            CTC_STAMP(lane == 0, 21);
            const unsigned char *xc = a.x8 + 2 * c * a.n_in;
            unsigned char *dst = sRaw + s * RAWB;
            if (lane == 0) mbar_arrive_expect_tx(&raw_full[s], (uint32_t)((g_hi - g_lo) * 2));
            __syncwarp();
            for (long long g = g_lo + (long long)lane * 1024; g < g_hi; g += 32 * 1024) {  // 2 KiB pieces
                const long long n = g_hi - g < 1024 ? g_hi - g : 1024;
                tma_load_1d(dst + (g - g0) * 2, xc + 2 * g, (uint32_t)(n * 2), &raw_full[s]);
            }
        }
    } else if (warp < 2 * NCW) {
        // ------------------------------------------------------------------ converters (two groups, even / odd items)
        const int grp = warp / NCW, gt = tid - grp * NCONV;
        unsigned long long it = grp;
        ItemWalk w(blockIdx.x + grp * gridDim.x, 2u * gridDim.x, a.tiles_per_ch);
        for (unsigned long long item = blockIdx.x + (unsigned long long)grp * gridDim.x; item < a.nitems;
             item += 2ull * gridDim.x, it += 2, w.next()) {
            const int s = grp;
            const uint32_t ph = (uint32_t)((it >> 1) & 1);
            const unsigned long long c = w.c;
            const long long tile = w.tile;
            const long long g0 = tile * TI - HIN;

            const int rs = (int)(it % NRAW);
            CTC_STAMP(gt == 0, 0);
            mbar_wait_long(&raw_full[rs], (uint32_t)((it / NRAW) & 1));
            CTC_STAMP(gt == 0, 1);
            mbar_wait_long(&a_empty[s], ph ^ 1);  // the MMAs that read this stage two items ago are done
            CTC_STAMP(gt == 0, 2);
            const unsigned char *raw = sRaw + rs * RAWB;
            unsigned char *stage = sA + s * A_STAGE;
            const float2 *hc = a.hist_in + c * H + H;
            if (tile == 0 && gt < HIN) {
                // the 80 samples in front of the batch: carried ConvertNode values back to b - 128 (exact after rounding),
                // one per thread (S = -0.5 where no history exists); units 0 and 1 are skipped below
                const long long g = (long long)gt - HIN;
                uint32_t r = 0xB800B800u;
                if (H + g >= 0) {
                    const float2 hv = hc[g];
                    const __half2 hh = __floats2half2_rn(fmaf(hv.x, 127.5f, -0.5f), fmaf(hv.y, 127.5f, -0.5f));
                    r = *reinterpret_cast<const uint32_t *>(&hh);
                }
                *reinterpret_cast<uint32_t *>(stage + (gt % D) * A_PHASE + swz64(4u * (uint32_t)(gt / D))) = r;
            }
            for (int u = gt; u < NUNIT; u += NCONV) {
                if (tile == 0 && u < HIN / 40) continue;  // the history units, converted above
                uint32_t v[40];  // sample s of the unit: phase s % 5, slot s / 5
                const long long gu = g0 + 40 * u;
                if (gu >= 0 && gu + 40 <= (long long)a.n_in) {  // the unit lies inside the batch
                    uint32_t w[20];
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const uint4 q = *reinterpret_cast<const uint4 *>(raw + u * 80 + k * 16);
                        w[4 * k] = q.x;
                        w[4 * k + 1] = q.y;
                        w[4 * k + 2] = q.z;
                        w[4 * k + 3] = q.w;
                    }
#pragma unroll
                    for (int sidx = 0; sidx < 40; ++sidx) v[sidx] = bytes_to_half2(w[sidx >> 1], sidx & 1);
                } else if (gu >= (long long)a.n_in) {  // behind the batch: ConvertNode values of 0
#pragma unroll
                    for (int sidx = 0; sidx < 40; ++sidx) v[sidx] = 0xB800B800u;
                } else {  // the unit that straddles the end of the batch
#pragma unroll
                    for (int sidx = 0; sidx < 40; ++sidx) {
                        const long long g = gu + sidx;
                        uint32_t r = 0xB800B800u;  // S = -0.5: a ConvertNode value of 0
                        if (g < (long long)a.n_in) {
                            const uint32_t wd = *reinterpret_cast<const uint32_t *>(raw + (40 * u + (sidx & ~1)) * 2);
                            r = bytes_to_half2(wd, sidx & 1);
                        }
                        v[sidx] = r;
                    }
                }
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    unsigned char *pb = stage + d * A_PHASE;
                    const uint32_t o = 32u * (uint32_t)u;  // slot 8 u of the phase stream, 4 bytes per sample
                    *reinterpret_cast<uint4 *>(pb + swz64(o)) = make_uint4(v[d], v[d + 5], v[d + 10], v[d + 15]);
                    *reinterpret_cast<uint4 *>(pb + swz64(o + 16)) = make_uint4(v[d + 20], v[d + 25], v[d + 30], v[d + 35]);
                }
            }
            fence_proxy_async();
            mbar_arrive(&a_full[s]);
            CTC_STAMP(gt == 0, 3);
            CTC_STAMP(gt == NCONV - 1, 4);
            mbar_arrive(&raw_empty[rs]);  // the raw tile has been consumed: the TMA warp may refill the stage
            CTC_STAMP(gt == 0, 5);
            // ---- carried state for the next call (last tile of the channel): ConvertNode's exact values
            if (tile == (long long)a.tiles_per_ch - 1 && a.hist_out != nullptr) {
                const float2 *hin = a.hist_in + c * H;
                float2 *ho = a.hist_out + c * H;
                for (long long i = gt; i < H; i += NCONV) {
                    const long long g = (long long)a.n_in - H + i;
                    float2 hv;
                    if (g < 0) {
                        hv = hin[H + g];
                    } else {
                        const unsigned char *b = a.x8 + 2 * (c * a.n_in + g);
                        hv = make_float2(__fdiv_rn(__fsub_rn((float)b[0], 127.5f), 127.5f), __fdiv_rn(__fsub_rn((float)b[1], 127.5f), 127.5f));
                    }
                    ho[i] = hv;
                }
            }
        }
    } else if (warp == W_MMA) {
        // ------------------------------------------------------------------ MMA issuer
        // The WHOLE warp runs this loop and one elected lane issues.  Inside `if (lane == 0)` the compiler cannot prove the
        // descriptors warp-uniform and wraps every tcgen05.mma in an ELECT / R2UR.BROADCAST / branch sequence (14
        // instructions, ~100 cycles each on a scheduler shared with four busy warps: 2.0 k cycles per tile, which bounded
        // the kernel); warp-uniform code keeps them in uniform registers.
        unsigned long long it = 0;
        const uint64_t bd0 = kdesc64(smem_u32(sB));
        const uint64_t ad0 = kdesc64(smem_u32(sA));
        for (unsigned long long item = blockIdx.x; item < a.nitems; item += gridDim.x, ++it) {
            const int s = (int)(it & 1);
            const uint32_t ph = (uint32_t)((it >> 1) & 1);
            CTC_STAMP(lane == 0, 8);
            mbar_wait_long(&t_empty[s], ph ^ 1);
            CTC_STAMP(lane == 0, 9);
            mbar_wait_long(&a_full[s], ph);
            CTC_STAMP(lane == 0, 10);
            tc_fence_after();
            const uint64_t ad = ad0 + (uint64_t)((s * A_STAGE) >> 4);
            const uint32_t dcol = tmem_base + (uint32_t)s * 64u;
            if (elect_one()) {
#pragma unroll
                for (int t = 0; t < D * 4; ++t) {
                    const int d = t >> 2, kb = (t >> 1) & 1, ks = t & 1;
                    const uint64_t aoff = (uint64_t)((d * A_PHASE + kb * 64 + ks * 32) >> 4);
                    const uint64_t boff = (uint64_t)((d * B_PHASE + kb * 4096 + ks * 32) >> 4);
                    if (t == 0) tc_mma<false>(dcol, ad + aoff, bd0 + boff, IDESC);
                    else tc_mma<true>(dcol, ad + aoff, bd0 + boff, IDESC);
                }
                tc_commit(&a_empty[s]);
                tc_commit(&t_full[s]);
            }
            __syncwarp();
            CTC_STAMP(lane == 0, 11);
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 8-15)
        const int e = (warp - W_EPI) & 3;    // TMEM sub-partition = warp % 4
        const int half = (warp - W_EPI) >> 2;  // outputs 8 half .. 8 half + 7 of the row
        const int row = 32 * e + lane;
        unsigned long long it = 0;
        ItemWalk w(blockIdx.x, gridDim.x, a.tiles_per_ch);
        for (unsigned long long item = blockIdx.x; item < a.nitems; item += gridDim.x, ++it, w.next()) {
            const int s = (int)(it & 1);
            const uint32_t ph = (uint32_t)((it >> 1) & 1);
            const unsigned long long c = w.c;
            const long long tile = w.tile;
            const long long m0 = tile * TO + (long long)row * RS + 8 * half;  // first output of this thread
            CTC_STAMP(tid == 32 * W_EPI, 12);
            mbar_wait_long(&t_full[s], ph);
            CTC_STAMP(tid == 32 * W_EPI, 13);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * e) << 16) + (uint32_t)s * 64u + 16u * (uint32_t)half;
            uint32_t p[16], r[16], pp[2] = {0u, 0u}, rp[2] = {0u, 0u};
            tc_ld16(taddr, p);
            tc_ld16(taddr + 32, r);
            if (FM && half) {  // output 7 of the row: the predecessor of this thread's first output
                tc_ld2(taddr - 2, pp);
                tc_ld2(taddr + 30, rp);
            }
            tc_wait_ld();
            tc_fence_before();
            mbar_arrive(&t_empty[s]);
            CTC_STAMP(tid == 32 * W_EPI, 14);
            float2 y[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                y[q] = make_float2(fmaf(__uint_as_float(p[2 * q]) + __uint_as_float(r[2 * q]), a.tap_inv_scale, a.dc),
                                   fmaf(__uint_as_float(p[2 * q + 1]) + __uint_as_float(r[2 * q + 1]), a.tap_inv_scale, a.dc));
            if (FM) {
                float2 yp = make_float2(fmaf(__uint_as_float(pp[0]) + __uint_as_float(rp[0]), a.tap_inv_scale, a.dc),
                                        fmaf(__uint_as_float(pp[1]) + __uint_as_float(rp[1]), a.tap_inv_scale, a.dc));
                if (half) ylast[it & 1][row] = y[7];  // output 15 of the row: the predecessor of the next row's first
                CTC_STAMP(tid == 32 * W_EPI, 15);
                asm volatile("bar.sync 5, %0;" ::"n"(NEPI) : "memory");
                CTC_STAMP(tid == 32 * W_EPI, 16);
                if (!half) {
                    if (row > 0) {
                        yp = ylast[it & 1][row - 1];
                    } else {
                        a.seam[2 * item] = y[0];  // the tile's first output: its discriminator step is the seam pass's
                        yp = y[0];
                    }
                }
                if (half && row == ROWS - 1) a.seam[2 * item + 1] = y[7];  // the tile's last output (zeros-only rows: unused)
                float o[8];
#pragma unroll
                for (int q = 0; q < 8; q += 2) {
                    const float2 t = fm_angle_fast2(y[q], q ? y[q - 1] : yp, y[q + 1], y[q]);
                    o[q] = t.x;
                    o[q + 1] = t.y;
                }
                float *chan = reinterpret_cast<float *>(a.out) + c * a.n_out;
                float *dst = chan + m0;
                if ((reinterpret_cast<uintptr_t>(chan) & 15) == 0) {  // (the same for every thread of the CTA)
                    if (m0 + 8 <= (long long)a.n_out) {
                        stg_stream(reinterpret_cast<float4 *>(dst), make_float4(o[0], o[1], o[2], o[3]));
                        stg_stream(reinterpret_cast<float4 *>(dst + 4), make_float4(o[4], o[5], o[6], o[7]));
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            if (m0 + q < (long long)a.n_out) dst[q] = o[q];
                    }
                } else {
                    // A channel's output row starts wherever c * n_out puts it (n_out = 26215 for fm_radio's reads): lanes
                    // that own rows would write 4 bytes each at a 64-byte pitch.  Through shared memory instead: every
                    // store instruction of a warp writes 128 contiguous bytes.  Two buffers, one barrier per tile: a
                    // buffer is rewritten two tiles later, after every thread has passed the next tile's barrier.
                    float *st = sOut + (it & 1) * TO;
                    *reinterpret_cast<float4 *>(st + RS * row + 8 * half) = make_float4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<float4 *>(st + RS * row + 8 * half + 4) = make_float4(o[4], o[5], o[6], o[7]);
                    asm volatile("bar.sync 6, %0;" ::"n"(NEPI) : "memory");
                    const int et = tid - 32 * W_EPI;
                    const long long t0 = tile * TO;
#pragma unroll
                    for (int k = 0; k < TO / NEPI; ++k) {
                        const int idx = et + k * NEPI;
                        if (t0 + idx < (long long)a.n_out) chan[t0 + idx] = st[idx];
                    }
                }
                CTC_STAMP(tid == 32 * W_EPI, 17);
                const long long last = (long long)a.n_out - 1 - m0;  // carried FM state: the call's last filter output
                if (last >= 0 && last < 8) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (q == last) a.prev_out[c] = y[q];
                }
            } else {
                float2 *dst = reinterpret_cast<float2 *>(a.out) + c * a.n_out + m0;
                if (m0 + 8 <= (long long)a.n_out && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                    for (int q = 0; q < 8; q += 2) stg_stream(reinterpret_cast<float4 *>(dst + q), make_float4(y[q].x, y[q].y, y[q + 1].x, y[q + 1].y));
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (m0 + q < (long long)a.n_out) dst[q] = y[q];
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
    }
}

// First discriminator output of every tile: arg(y[first] * conj(y[first - 1])), y[first - 1] = the previous tile's last
// output (another CTA's) or the carried FM state for the first tile of the call (analog.rs:22-34, 43-47).
__global__ void __launch_bounds__(256) chain_tc_seam_kernel(const float2 *__restrict__ seam, const float2 *__restrict__ prev_in,
                                                            float *__restrict__ out, unsigned long long nitems, unsigned tiles_per_ch,
                                                            unsigned long long n_out)
{
    const unsigned long long item = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= nitems) return;
    const unsigned long long c = item / tiles_per_ch, tile = item % tiles_per_ch;
    const float2 yp = tile == 0 ? prev_in[c] : seam[2 * (item - 1) + 1];
    out[c * n_out + tile * TO] = fm_angle_fast(seam[2 * item], yp);
}

}  // namespace ctc

// ------------------------------------------------------------------------------------ host side
size_t chain_tc_seam_entries(size_t n_out, size_t channels) { return 2 * ceil_div(n_out, (size_t)ctc::TO) * channels; }
size_t chain_tc_image_bytes() { return ctc::B_BYTES; }

bool chain_tc_supported(uint32_t ntaps, uint32_t decim, bool mix, bool cplx) { return !mix && !cplx && decim == ctc::D && ntaps >= 1 && ntaps <= 64; }

// Tap image: per phase d and K-block kb, 64 rows (n < 32: hi part of B_d[n][32 kb .. 32 kb + 31], n >= 32: lo part) x 64
// bytes, 64-byte swizzled like the device reads it.  taps: real parts of the filter; the image holds taps / 127.5.
void chain_tc_build_image(const float *taps_re, uint32_t ntaps, unsigned char *img, float *tap_inv_scale, float *dc)
{
    using namespace ctc;
    double g[D][RS + 1];  // g[d][j], j = 0 .. 13 used
    double sum = 0.0, mx = 0.0;
    for (int d = 0; d < D; ++d)
        for (int j = 0; j <= RS; ++j) {
            const int k = d == 0 ? D * j : D * (j - 1) + D - d;
            const double v = (j >= (d == 0 ? 0 : 1) && k >= 0 && k < (int)ntaps) ? (double)taps_re[k] / 127.5 : 0.0;
            g[d][j] = v;
            sum += v;
            if (fabs(v) > mx) mx = fabs(v);
        }
    *dc = (float)(0.5 * sum);
    int eb = 0;
    if (mx > 0.0) frexp(mx, &eb);               // mx = f * 2^eb, f in [0.5, 1)
    const double sc = ldexp(1.0, 15 - eb);      // mx * sc in [2^14, 2^15)
    *tap_inv_scale = (float)ldexp(1.0, eb - 15);
    memset(img, 0, B_BYTES);
    for (int d = 0; d < D; ++d)
        for (int n = 0; n < 32; ++n) {
            const int q = n >> 1, c = n & 1;
            for (int k = 0; k < 64; ++k) {
                const int kk = k >> 1, cc = k & 1;
                const int j = q + HALO - kk;
                double v = 0.0;
                if (c == cc && j >= 0 && j <= RS) v = g[d][j] * sc;
                const __half h = __float2half_rn((float)v);
                const __half l = __float2half_rn((float)(v - (double)__half2float(h)));
                const int kb = k >> 5, kr = k & 31;
                const uint32_t o_hi = swz64((uint32_t)n * 64u + (uint32_t)kr * 2u), o_lo = swz64((uint32_t)(32 + n) * 64u + (uint32_t)kr * 2u);
                unsigned char *base = img + (size_t)d * B_PHASE + (size_t)kb * 4096;
                memcpy(base + o_hi, &h, 2);
                memcpy(base + o_lo, &l, 2);
            }
        }
}

bool chain_tc_applicable(const ChainArgs &args, size_t channels)
{
    if (args.x8 == nullptr || args.tc_img == nullptr || args.tc_seam == nullptr || args.decim != (unsigned)ctc::D || args.ntaps > 64)
        return false;
    if (args.n_in % 8 != 0 || (reinterpret_cast<uintptr_t>(args.x8) & 15) != 0 || args.hist_len < (unsigned)ctc::HIN) return false;
    const char *e = getenv("COMMS_B200_CHAIN_PATH");  // auto (default) | tc (every batch size) | v3 / v2 (CUDA-core kernels only)
    const int path = (e && strcmp(e, "tc") == 0) ? 1 : ((e && (strcmp(e, "v3") == 0 || strcmp(e, "v2") == 0)) ? -1 : 0);
    if (path < 0) return false;
    const size_t tiles = ceil_div(args.n_out, (size_t)ctc::TO);
    if (tiles * channels >= ((size_t)1 << 31)) return false;  // the kernel walks items with 32-bit arithmetic
    return path > 0 || tiles * channels >= 96;  // small calls: the CUDA-core kernel's 768-output tiles fill the SMs better
}

#ifdef CB_CTC_TIMELINE
unsigned long long *g_ctc_dbg = nullptr;
#endif

int launch_chain_tc(const ChainArgs &args, const ChainTaps &taps, bool fm, size_t channels, cudaStream_t s)
{
    ctc::Args a;
    a.x8 = args.x8;
    a.out = args.out;
    a.hist_in = args.hist_in;
    a.hist_out = args.hist_out;
    a.prev_in = args.prev_in;
    a.prev_out = args.prev_out;
    a.bimg = reinterpret_cast<const uint4 *>(args.tc_img);
    a.n_in = args.n_in;
    a.n_out = args.n_out;
    a.hist_len = args.hist_len;
    a.ntaps = args.ntaps;
    a.tiles_per_ch = (unsigned)ceil_div(args.n_out, (size_t)ctc::TO);
    a.nitems = (unsigned long long)a.tiles_per_ch * channels;
    a.tap_inv_scale = args.tc_inv_scale;
    a.dc = args.tc_dc;
    a.seam = args.tc_seam;
#ifdef CB_CTC_TIMELINE
    a.dbg = g_ctc_dbg;
#endif
    ChainTaps scaled = taps;  // look-back value: taps / 127.5 against b - 127.5, as the CUDA-core kernel
    for (int k = 0; k < CHAIN_MAX_TAP_SLOTS; ++k) {
        const float t = k < (int)args.ntaps ? (float)((double)taps.t[k].x / 127.5) : 0.f;
        scaled.t[k] = make_float2(t, t);
    }
    auto kern = fm ? ctc::chain_tc_kernel<true> : ctc::chain_tc_kernel<false>;
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ctc::SMEM));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)(a.nitems < (unsigned long long)sms ? a.nitems : (unsigned long long)sms);
    kern<<<grid, ctc::NTHREADS, ctc::SMEM, s>>>(a, scaled);
    count_launch();
    if (fm) {
        ctc::chain_tc_seam_kernel<<<(unsigned)ceil_div((size_t)a.nitems, (size_t)256), 256, 0, s>>>(
            a.seam, a.prev_in, reinterpret_cast<float *>(a.out), a.nitems, a.tiles_per_ch, a.n_out);
        count_launch();
    }
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

}  // namespace cb
