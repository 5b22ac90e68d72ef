// K5-P: 65536-point FFT / IFFT in ONE HBM pass on persistent, software-pipelined 8-CTA clusters (sm_100a).
//
// Reference semantics (src/fft/mod.rs:73-96): X[k] = sum_n x[n] e^{-/+ j 2 pi k n / N}, unnormalised.
//
// Why another 65536-point kernel.  A frame is 512 KiB, more than one SM's shared memory, so the transform needs an
// exchange between SMs.  Through L2 (K5-R, fft_rows_kernel.cu) every sample crosses the SM<->L2 fabric four times and
// that fabric carries about what HBM does (~7.3 TB/s measured): 55 % of the 16 B/sample roof, whatever is done to the
// kernel.  Through distributed shared memory the crossing disappears, but the first cluster kernel (K5-C,
// fft_cluster_kernel.cu) spent 17 % of a CTA's life issuing remote stores, 19 % waiting for its peers and never had
// the next frame's loads in flight: 42 %.  This kernel keeps the cluster form and takes the data movement off the
// compute warps:
//
//   n = 256 n1 + n2,  k = k1 + 256 k2      (the same 256 x 256 split as the other two)
//   X[k1 + 256 k2] = sum_{n2} W256^{n2 k2} { W65536^{n2 k1} sum_{n1} x[256 n1 + n2] W256^{n1 k1} }
//
//   * one persistent CTA per SM, clusters of 8; cluster c transforms frames c, c + C, c + 2C, ...
//   * a DMA warp moves everything: TMA bulk copies HBM -> IN[b] (256 rows of 256 B: the CTA's 32 columns n2),
//     cp.async.bulk shared::cta -> shared::cluster for the all-to-all (8 blocks of 8 KiB, one per peer, each landing on
//     the receiver's mbarrier), and the refill of IN[b] two frames ahead;
//   * 16 compute warps only do shared-memory <-> register math: step A (32 column FFTs of 256 points, two radix-16
//     passes) in place in IN[b], leaving Y[k1][n2] in exactly the layout the pushes need (block r' = rows k1 in
//     [32 r', 32 r' + 32) is contiguous); step B (twiddle, 32 row FFTs) in place in the exchange buffer, results
//     straight from registers to HBM (lane = k1: 256 contiguous bytes per warp store);
//   * software pipeline: an iteration runs A(i) and then B(i-1).  The pushes of frame i-1 leave when EVERY CTA of the
//     cluster has read its exchange buffer for frame i-2 (`ready`), and land while A(i) runs, so the all-to-all latency
//     and the skew between the CTAs hide behind step A instead of stalling step B; loads have a whole period to land.
//
// Shared memory: IN[2] (2 x 64 KiB) + EXCH (32 rows x 273 float2, padded for the in-place row passes) = 196.3 KiB.
// mbarriers: full[b] (TMA landed), send_rdy (A done: IN[b] holds the send layout), ready (8 arrivals: all CTAs have
// drained EXCH), landed (64 KiB of pushes arrived), sent_ok (8 arrivals: my pushes were received, IN[b] may be refilled).
// Algorithmic HBM traffic: 8 B read + 8 B written per sample; DSMEM: 7 B per sample each way.
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)

#include "fft2_core.cuh"
#include "fft_kernels.cuh"

namespace cb {

namespace fftp {

constexpr int NF = 65536;
constexpr int CL = 8;               // CTAs per cluster
constexpr int COLS = 256 / CL;      // columns n2 (step A), then rows k1 (step B), per CTA
constexpr int NCOMP = 512;          // compute threads: 16 warps x 32 lanes = 16 radix-16 tasks x 32 columns / rows
constexpr int NGRP = 256;           // ... in two independent groups (16 columns / rows each) with their own named barrier
constexpr int NTHREADS = NCOMP + 32;
constexpr int IN_F2 = 256 * COLS;   // float2 per IN buffer
constexpr int RP = 273;             // padded row pitch of the exchange buffer for the row passes (float2)
constexpr int EXCH_F2 = COLS * RP;
constexpr int BLOCK_BYTES = COLS * COLS * 8;  // one peer's block of the all-to-all
constexpr int TWP_F2 = 16 * 16;     // W256^(j m), j, m = 0..15: the inter-pass twiddles of both steps
constexpr int SMEM = (2 * IN_F2 + EXCH_F2 + TWP_F2) * (int)sizeof(float2) + 128;

__device__ __forceinline__ uint32_t cluster_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_rank(uint32_t saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// shared::cta -> shared::cluster bulk copy, completion (in bytes) on the DESTINATION CTA's mbarrier
__device__ __forceinline__ void dsmem_push(uint32_t dst_cluster_addr, const void *src_smem, uint32_t bytes, uint32_t remote_bar)
{
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster_addr),
                 "r"(smem_u32(src_smem)), "r"(bytes), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void grp_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(NGRP) : "memory"); }
// one frame's share of this CTA, HBM -> shared: box of 256 rows (n1) x 256 bytes (the CTA's 32 columns), ONE instruction
__device__ __forceinline__ void tma_load_box(void *dst_smem, const CUtensorMap *map, int x, int y, uint64_t *bar, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
            smem_u32(dst_smem)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void st_cs(float2 *p, float2 v)
{
    asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// scripts/fftp_timeline.cu compiles this file with CB_FFTP_TIMELINE: clock64 stamps of one steady-state iteration
#ifdef CB_FFTP_TIMELINE
#define FFTP_DBG_PARAM , unsigned long long *dbg
#define FFTP_STAMP(cond, k)                                                        \
    do {                                                                           \
        if ((cond) && it == 6) dbg[(size_t)blockIdx.x * 16 + (k)] = clock64();      \
    } while (0)
#else
#define FFTP_DBG_PARAM
#define FFTP_STAMP(cond, k)
#endif

template <bool INV>
__global__ void __launch_bounds__(NTHREADS, 1)
fft65536_cpipe_kernel(const __grid_constant__ CUtensorMap in_map, float2 *__restrict__ out, const float2 *__restrict__ twN,
                      unsigned long long nframes, unsigned nclusters FFTP_DBG_PARAM)
{
    using namespace fft2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2 *sIn = reinterpret_cast<float2 *>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    float2 *sEx = sIn + 2 * IN_F2;
    float2 *sTw = sEx + EXCH_F2;  // sTw[16 j + m] = W256^(j m)
    __shared__ __align__(8) uint64_t full[2], send_rdy, ready, landed, sent_ok;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_rank();
    const unsigned long long cid = cluster_id();
    // frames of this cluster: cid, cid + nclusters, ...
    const unsigned long long nf = cid < nframes ? (nframes - cid + nclusters - 1) / nclusters : 0;

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&send_rdy, 2);       // one arrival per compute group
        mbar_init(&ready, 2 * CL);     // both groups of all CTAs
        mbar_init(&landed, 1);
        mbar_init(&sent_ok, CL);
        fence_mbar_init();
    }
    if (tid < TWP_F2) sTw[tid] = __ldg(twN + 256 * (((tid >> 4) * (tid & 15)) & 255));
    __syncthreads();
    // every CTA's barriers are initialised before any peer arrives on them or pushes into this CTA
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");

    if (warp == NCOMP / 32) {
        // ------------------------------------------------------------------ DMA warp
        const uint64_t pol = l2_evict_first_policy();
        auto issue_load = [&](unsigned long long it) {
            const int b = (int)(it & 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&full[b], IN_F2 * 8);
                tma_load_box(sIn + b * IN_F2, &in_map, 2 * COLS * (int)rank, (int)(256 * (cid + it * nclusters)), &full[b], pol);
            }
            __syncwarp();
        };
        if (nf > 0) issue_load(0);
        if (nf > 1) issue_load(1);
        const uint32_t ex_addr = smem_u32(sEx), landed_addr = smem_u32(&landed), sent_addr = smem_u32(&sent_ok);
        for (unsigned long long it = 0; it < nf; ++it) {
            const int b = (int)(it & 1);
            FFTP_STAMP(lane == 0, 8);
            mbar_wait(&send_rdy, (uint32_t)(it & 1));                       // A(it) done: IN[b] holds Y[k1][n2 local]
            FFTP_STAMP(lane == 0, 9);
            if (it >= 1) mbar_wait_cluster(&ready, (uint32_t)((it - 1) & 1));  // every CTA has drained EXCH of frame it-1
            FFTP_STAMP(lane == 0, 10);
            if (lane == 0) mbar_arrive_expect_tx(&landed, CL * BLOCK_BYTES);
            __syncwarp();
            if (lane < CL)  // block `lane` (rows k1 of CTA `lane`) -> slot `rank` of that CTA's exchange buffer
                dsmem_push(map_rank(ex_addr + rank * BLOCK_BYTES, (uint32_t)lane), sIn + b * IN_F2 + lane * (COLS * COLS), BLOCK_BYTES,
                           map_rank(landed_addr, (uint32_t)lane));
            FFTP_STAMP(lane == 0, 11);
            mbar_wait_cluster(&landed, (uint32_t)(it & 1));                 // all 8 blocks of frame `it` are here ...
            FFTP_STAMP(lane == 0, 12);
            if (lane < CL) mbar_arrive_remote(map_rank(sent_addr, (uint32_t)lane));  // ... tell their senders
            mbar_wait_cluster(&sent_ok, (uint32_t)(it & 1));                // my 8 blocks were received: IN[b] is free
            FFTP_STAMP(lane == 0, 13);
            if (it + 2 < nf) issue_load(it + 2);
            FFTP_STAMP(lane == 0, 14);
        }
    } else {
        // ------------------------------------------------------------------ compute warps
        // two independent groups of 8 warps: group g owns columns (step A) / rows (step B) [16 g, 16 g + 16) of the CTA's 32.
        // Within a group: lane -> (cl = lane & 15 column / row, jl = lane >> 4), warp wg -> radix-16 task j = 2 wg + jl.
        const int g = warp >> 3, wg = warp & 7;
        const int c = 16 * g + (lane & 15), j = 2 * wg + (lane >> 4);
        const uint32_t ready_addr = smem_u32(&ready);
        // frame-invariant twiddles of this thread
        const int k1l3 = 16 * g + 2 * wg + (lane >> 4), lo3 = lane & 15;  // pass 3: row k1l3, n2 digit lo3
        const int k1g = COLS * (int)rank + k1l3;
        const float2 tw3c = __ldg(twN + k1g * lo3), tw3w = __ldg(twN + 16 * k1g);
        for (unsigned long long it = 0; it <= nf; ++it) {
            float2 v[16];
            if (it < nf) {
                // ---- step A(it), in place in IN[b]
                float2 *buf = sIn + (it & 1) * IN_F2;
                FFTP_STAMP(tid == 0, 0);
                mbar_wait(&full[it & 1], (uint32_t)((it >> 1) & 1));
                FFTP_STAMP(tid == 0, 1);
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = buf[(j + 16 * m) * COLS + c];
                bfly16<INV>(v);
                grp_sync(g);  // every thread of the group holds its inputs (the group's 16 columns are its own)
#pragma unroll
                for (int sl = 0; sl < 16; ++sl) buf[(16 * j + q16(sl)) * COLS + c] = v[sl];
                grp_sync(g);
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = buf[(j + 16 * m) * COLS + c];
#pragma unroll
                for (int m = 1; m < 16; ++m) v[m] = cmul(v[m], sTw[16 * j + m]);
                bfly16<INV>(v);
                grp_sync(g);
#pragma unroll
                for (int sl = 0; sl < 16; ++sl) buf[(j + 16 * q16(sl)) * COLS + c] = v[sl];  // Y[k1][c], k1 = j + 16 q
                fence_proxy_async();  // the pushes read this through the async proxy
                grp_sync(g);
                if ((tid & (NGRP - 1)) == 0) mbar_arrive_local(&send_rdy);
                FFTP_STAMP(tid == 0, 2);
            }
            if (it >= 1) {
                // ---- step B(it-1), in place in EXCH; block s of EXCH holds Y[k1 local][n2 = 32 s + i]
                const unsigned long long frame = cid + (it - 1) * nclusters;
                mbar_wait_cluster(&landed, (uint32_t)((it - 1) & 1));
                FFTP_STAMP(tid == 0, 3);
                // The padded row layout of passes 3 -> 4 (row pitch RP) overlaps OTHER rows' landed blocks, so the rewrite
                // needs every row of the CTA read first: one CTA-wide barrier here; the rest of step B is per group.
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = sEx[(m >> 1) * (COLS * COLS) + k1l3 * COLS + lo3 + 16 * (m & 1)];  // n2 = lo + 16 m
                twiddle16c(v, tw3c, tw3w);
                bfly16<INV>(v);
                asm volatile("bar.sync 3, %0;" ::"n"(NCOMP) : "memory");
#pragma unroll
                for (int sl = 0; sl < 16; ++sl) sEx[k1l3 * RP + pad16(16 * lo3 + q16(sl))] = v[sl];
                grp_sync(g);
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = sEx[c * RP + pad16(j + 16 * m)];
                grp_sync(g);  // the group's rows are drained; when both groups of every CTA are, the peers may push again
                if ((tid & (NGRP - 1)) < CL) mbar_arrive_remote(map_rank(ready_addr, (uint32_t)(tid & (NGRP - 1))));
                FFTP_STAMP(tid == 0, 4);
#pragma unroll
                for (int m = 1; m < 16; ++m) v[m] = cmul(v[m], sTw[16 * j + m]);
                bfly16<INV>(v);
                float2 *dst = out + frame * NF + COLS * rank + c;
#pragma unroll
                for (int sl = 0; sl < 16; ++sl) st_cs(dst + 256 * (j + 16 * q16(sl)), v[sl]);  // X[k1 + 256 k2], k2 = j + 16 q
                FFTP_STAMP(tid == 0, 5);
            }
        }
    }
    // no CTA leaves while a peer may still arrive on its barriers
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

#ifdef CB_FFTP_TIMELINE
static unsigned long long *g_fftp_dbg = nullptr;
static int g_fftp_resident = 0;
#endif

// input viewed as a 2-D f32 tensor: 512 floats per row (one n1), 256 * nframes rows; box = 64 floats x 256 rows
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            (void)cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

bool cpipe_available() { return encode_fn() != nullptr; }

template <bool INV>
static int launch(const float2 *in, float2 *out, const float2 *twN, size_t nframes, cudaStream_t s)
{
    auto kern = fft65536_cpipe_kernel<INV>;
    CB_REQUIRE(encode_fn() != nullptr, CB_ERR_UNSUPPORTED, "fft: cuTensorMapEncodeTiled is not available from this driver");
    CB_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0, CB_ERR_INVALID_ARG, "fft: input must be 16-byte aligned");
    CUtensorMap map;
    {
        const cuuint64_t dims[2] = {512, (cuuint64_t)256 * nframes};
        const cuuint64_t strides[1] = {2048};
        const cuuint32_t box[2] = {2 * COLS, 256}, estr[2] = {1, 1};
        const CUresult r = encode_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float2 *>(in), dims, strides, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CB_REQUIRE(r == CUDA_SUCCESS, CB_ERR_CUDA, "fft: cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    static int resident[2] = {0, 0};
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(NTHREADS, 1, 1);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (resident[INV] == 0) {
        int nc = 0;
        cfg.gridDim = dim3(148 * CL, 1, 1);
        if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess || nc <= 0) {
            (void)cudaGetLastError();
            nc = 16;
        }
        resident[INV] = nc;
    }
    const unsigned nclusters = (unsigned)(nframes < (size_t)resident[INV] ? nframes : (size_t)resident[INV]);
    cfg.gridDim = dim3(nclusters * CL, 1, 1);
#ifdef CB_FFTP_TIMELINE
    g_fftp_resident = resident[INV];
    CB_CUDA(cudaLaunchKernelEx(&cfg, kern, map, out, twN, (unsigned long long)nframes, nclusters, g_fftp_dbg));
#else
    CB_CUDA(cudaLaunchKernelEx(&cfg, kern, map, out, twN, (unsigned long long)nframes, nclusters));
#endif
    count_launch();
    return CB_OK;
}

}  // namespace fftp

// twN: 65536 entries e^{-/+ 2 pi i k / 65536} (direction baked in, as FftPlanDev::tw)
int launch_fft65536_cpipe(const float2 *in, float2 *out, const float2 *twN, size_t nframes, bool inverse, cudaStream_t s)
{
    if (nframes == 0) return CB_OK;
    return inverse ? fftp::launch<true>(in, out, twN, nframes, s) : fftp::launch<false>(in, out, twN, nframes, s);
}

}  // namespace cb
