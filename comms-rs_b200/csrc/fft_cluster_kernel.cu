// K5-C: 65536-point FFT / IFFT in ONE HBM pass on a thread-block cluster (sm_100a).
//
// Reference semantics (src/fft/mod.rs:73-96): X[k] = sum_n x[n] e^{-/+ j 2 pi k n / N}, unnormalised.
//
// A 65536-point frame is 512 KiB -- more than one SM's shared memory -- so the single-pass
// kernel of fft_kernels.cu stops at 8192 points and the four-step path costs two HBM round trips
// plus one launch pair per L2-sized group.  Here one frame is owned by a cluster of 8 CTAs
// (8 x 64 KiB of distributed shared memory) and read / written from HBM exactly once.
//
//   n = 256 n1 + n2,  k = k1 + 256 k2
//   X[k1 + 256 k2] = sum_{n2} W256^{n2 k2} { W65536^{n2 k1} sum_{n1} x[256 n1 + n2] W256^{n1 k1} }
//
// CTA `rank` owns the 32 columns n2 in [32 rank, 32 rank + 32).
//   pass 1-2 (step A): 32 column FFTs of length 256 (= 16 x 16, Stockham radix 16 in registers,
//             one shared-memory exchange); lane = column, so every global load is 256 contiguous bytes
//   exchange: once every CTA of the cluster no longer reads its own buffer, every thread pushes its
//             results Y[k1][n2] straight into the buffer of the CTA that owns k1 (remote shared-memory
//             stores, 256 contiguous bytes per half-warp).  Three interchangeable synchronisations
//             (template MODE; all measured within 4 % of each other, the exchange itself -- 57 KiB per
//             CTA over a ~20 B/clk DSMEM port -- is what costs): barrier.cluster arrive.release / wait
//             around the pushes (default), or remote mbarrier arrives for "ready" and st.async with
//             transaction bytes for "landed".  A relaxed arrive is NOT enough for "ready": the reads
//             must have been performed, not just issued (found by the full-size Parseval test).
//   pass 3-4 (step B): CTA `rank` now holds rows k1 in [32 rank, 32 rank + 32), all 256 n2:
//             twiddle W^{k1 n2} on the way in (two table look-ups + a depth-4 power tree),
//             32 row FFTs of length 256; lane = k1, so every global store is 256 contiguous bytes
// Work split inside a CTA (256 threads): thread (cp, j), cp = lane & 15, j = 2 warp + (lane >> 4),
// owns the column PAIR (2 cp, 2 cp + 1) and the radix-16 task j of both columns, so global loads /
// stores and the remote pushes are 128-bit and a half-warp always moves 256 contiguous bytes.
// Buffer layout: element (row r, i) at float2 index 256 r + 2 (r >> 1) + (i ^ ((r >> 4) & 1)):
// the 64-bit strided reads (same i, rows 2 cp) and the 128-bit contiguous writes (rows 2 cp,
// i = 16 j + 2 t, + 1) are both bank-conflict free; the XOR only swaps the halves of a 16-byte chunk.
// The next frame this cluster slot will see is prefetched into L2 (cp.async.bulk.prefetch.L2) while
// the current one is transformed.
// Algorithmic HBM traffic: 8 B read + 8 B written per sample.
#include "fft2_core.cuh"
#include "fft_kernels.cuh"

namespace cb {

namespace fftc {

constexpr int NF = 65536;
// CL CTAs per cluster (8: portable; 16: non-portable opt-in), COLS = 256 / CL columns (then rows) per CTA
__host__ __device__ constexpr int smem_bytes(int cl) { return ((256 / cl) * 256 + 2 * (128 / cl)) * (int)sizeof(float2); }

__device__ __forceinline__ uint32_t cluster_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// start-up rendezvous only (mbarrier inits visible cluster-wide, every peer CTA is running): relaxed
// arrive + fence.mbarrier_init, so no MEMBAR.GPU / L1 invalidate is emitted
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// arrive on a peer CTA's mbarrier (address from mapa)
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
__device__ __forceinline__ uint32_t map_rank(uint32_t saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// 16-byte asynchronous store into a peer CTA's shared memory; completion is counted (in bytes) on
// that CTA's mbarrier, which is what its threads wait on -- no cluster-wide barrier or fence
__device__ __forceinline__ void st_async4(uint32_t addr, float2 a, float2 b, uint32_t remote_bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
                 "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// float2 index of element (row r, i) of the CTA buffer
__device__ __forceinline__ int rowbase(int r) { return 256 * r + 2 * (r >> 1); }

// contiguous run b[16 j + q], q = 0..15, of one row from the bfly16 slot order, as 8 x 128-bit stores
// sw = the row's XOR bit: the two halves of each 16-byte chunk swap
__device__ __forceinline__ void store_run16(float2 *rowp, int j, int sw, const float2 *v)
{
    float4 *p = reinterpret_cast<float4 *>(rowp + 16 * j);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        // q = 2t lives in slot s0 = q16^-1(2t), q = 2t + 1 in s1; q16 is an involution on 0..15
        const float2 a = v[fft2::q16(2 * t)], b = v[fft2::q16(2 * t + 1)];
        p[t] = sw ? make_float4(b.x, b.y, a.x, a.y) : make_float4(a.x, a.y, b.x, b.y);
    }
}

// scripts/fftc_timeline.cu compiles this file with CB_FFTC_TIMELINE to stamp the phases of every CTA
#ifdef CB_FFTC_TIMELINE
#define FFTC_DBG_PARAM , unsigned long long *dbg
#define FFTC_STAMP(k)                                                        \
    do {                                                                     \
        if (threadIdx.x == 0) dbg[(size_t)blockIdx.x * 12 + (k)] = clock64(); \
    } while (0)
#else
#define FFTC_DBG_PARAM
#define FFTC_STAMP(k)
#endif

__device__ __forceinline__ void st_cluster4(uint32_t addr, float2 a, float2 b)
{
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y)
                 : "memory");
}
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }

// MODE 0: remote-arrive mbarrier ("ready") + st.async with transaction bytes ("landed")
// MODE 1: remote-arrive mbarrier ("ready") + plain remote stores and one barrier.cluster ("landed")
// MODE 2: barrier.cluster for both
template <bool INV, int MODE, int CL>
__global__ void __launch_bounds__(2048 / CL, CL == 16 ? 4 : 2)
fft65536_cluster_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, const float2 *__restrict__ twN,
                        size_t nframes, unsigned prefetch_dist FFTC_DBG_PARAM)
{
    using namespace fft2;
    constexpr int COLS = 256 / CL, CP = COLS / 2;  // columns per CTA, column pairs = lanes per task index j
    constexpr int NTH = CP * 16, NWARP = NTH / 32;
    FFTC_STAMP(0);
    extern __shared__ __align__(16) float2 sbuf[];
    // bar_ready: one arrival per warp of every CTA of the cluster once it no longer reads its own buffer
    // bar_full : the 64 KiB that the 8 CTAs push into this CTA's buffer (transaction bytes)
    __shared__ __align__(8) uint64_t bar_ready, bar_full;
    if (threadIdx.x == 0) {
        mbar_init(&bar_ready, CL * NWARP);
        mbar_init(&bar_full, 1);
        fence_mbar_init();
        if (MODE == 0) mbar_arrive_expect_tx(&bar_full, COLS * 256 * (uint32_t)sizeof(float2));
    }
    cluster_arrive_relaxed();
    const uint32_t rank = cluster_rank();
    const size_t frame = blockIdx.x / CL;
    const int lane = threadIdx.x & 31, cp = lane % CP, j = (32 / CP) * (threadIdx.x >> 5) + lane / CP;
    const float2 *src = in + frame * NF + COLS * rank + 2 * cp;
    float2 *dst = out + frame * NF + COLS * rank + 2 * cp;
    const int sw = (cp >> 3) & 1;               // XOR bit of rows 2 cp and 2 cp + 1
    float2 *row0 = sbuf + rowbase(2 * cp), *row1 = sbuf + rowbase(2 * cp + 1);

    float2 y0[16], y1[16];
    // ---- pass 1: x[256 (j + 16 m) + n2] -> b[16 j + q]
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const float4 t = ldg_stream(reinterpret_cast<const float4 *>(src + 256 * (j + 16 * m)));
        y0[m] = make_float2(t.x, t.y);
        y1[m] = make_float2(t.z, t.w);
    }
    if (frame + prefetch_dist < nframes) {  // the row segments of the slice this CTA slot sees next
#pragma unroll
        for (int r = 0; r < 256 / NTH; ++r)
            prefetch_l2(in + (frame + prefetch_dist) * NF + 256 * (threadIdx.x + r * NTH) + COLS * rank, COLS * 8);
    }
    FFTC_STAMP(1);
    bfly16<INV>(y0);
    FFTC_STAMP(2);
    bfly16<INV>(y1);
    store_run16(row0, j, sw, y0);
    store_run16(row1, j, sw, y1);
    __syncthreads();
    FFTC_STAMP(3);

    // ---- pass 2: b[j + 16 m] -> Y[k1 = j + 16 q][n2], kept in registers across the cluster barrier
    const float2 wj = __ldg(twN + 256 * j);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        y0[m] = row0[(j + 16 * m) ^ sw];
        y1[m] = row1[(j + 16 * m) ^ sw];
    }
    cluster_wait();  // start-up rendezvous: peers are running and their mbarriers are initialised
    // this warp is done reading the CTA's buffer: tell every CTA of the cluster (release: the reads
    // above are performed, not just issued, before a peer may overwrite the buffer)
    if (MODE == 2) {
        cluster_arrive_release();
    } else {
        __syncwarp();
        if (lane < CL) mbar_arrive_remote(map_rank(smem_u32(&bar_ready), (uint32_t)lane));
    }
    twiddle16(y0, wj);
    bfly16<INV>(y0);
    twiddle16(y1, wj);
    bfly16<INV>(y1);
    FFTC_STAMP(4);
    if (MODE == 2) cluster_wait();
    else mbar_wait_cluster(&bar_ready, 0);  // every CTA of the cluster is done reading its own buffer
    FFTC_STAMP(5);

    // ---- exchange: Y[k1][n2], Y[k1][n2 + 1] -> CTA k1 / 32, row k1 % 32, i = n2 (one 128-bit remote store)
    {
        // k1 = j + 16 q -> rank k1 / COLS, row rr = k1 % COLS = j + 16 (q % (COLS/16)), whose XOR bit is (rr >> 4) & 1
        const uint32_t local = smem_u32(sbuf + rowbase(j) + COLS * (int)rank + 2 * cp), lbar = smem_u32(&bar_full);
#pragma unroll
        for (int sl = 0; sl < 16; ++sl) {
            const int q = q16(sl);
            const int drank = (16 * q) / COLS, dhi = q % (COLS / 16);  // dhi = 1: rows 16..31 (COLS = 32 only)
            const uint32_t a = map_rank(local, (uint32_t)drank) + (uint32_t)(dhi * (256 * 16 + 16) * (int)sizeof(float2));
            const uint32_t rb = map_rank(lbar, (uint32_t)drank);
            if (MODE == 0) {
                if (dhi) st_async4(a, y1[sl], y0[sl], rb);
                else st_async4(a, y0[sl], y1[sl], rb);
            } else {
                if (dhi) st_cluster4(a, y1[sl], y0[sl]);
                else st_cluster4(a, y0[sl], y1[sl]);
            }
        }
    }
    if (MODE != 0) cluster_arrive_release();
    // ---- pass 3 twiddles while the pushes land: rows k1 = 32 rank + 2 cp (+ 1), W^{k1 (j + 16 m)}
    const int k1 = COLS * (int)rank + 2 * cp;
    const float2 c0 = __ldg(twN + k1 * j), c1 = __ldg(twN + (k1 + 1) * j);
    const float2 w0 = __ldg(twN + 16 * k1), w1 = __ldg(twN + 16 * (k1 + 1));
    FFTC_STAMP(6);
    if (MODE == 0) mbar_wait(&bar_full, 0);   // all 64 KiB pushed into this CTA's buffer have landed
    else cluster_wait();
    FFTC_STAMP(7);

    // ---- pass 3: R[k1][j + 16 m] * W^{k1 (j + 16 m)} -> b[16 j + q]  (in place)
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        y0[m] = row0[(j + 16 * m) ^ sw];
        y1[m] = row1[(j + 16 * m) ^ sw];
    }
    __syncthreads();
    twiddle16c(y0, c0, w0);
    bfly16<INV>(y0);
    store_run16(row0, j, sw, y0);
    twiddle16c(y1, c1, w1);
    bfly16<INV>(y1);
    store_run16(row1, j, sw, y1);
    __syncthreads();
    FFTC_STAMP(8);

    // ---- pass 4: b[j + 16 m] -> X[k1 + 256 (j + 16 q)]
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        y0[m] = row0[(j + 16 * m) ^ sw];
        y1[m] = row1[(j + 16 * m) ^ sw];
    }
    twiddle16(y0, wj);
    bfly16<INV>(y0);
    twiddle16(y1, wj);
    bfly16<INV>(y1);
#pragma unroll
    for (int sl = 0; sl < 16; ++sl)
        stg_stream(reinterpret_cast<float4 *>(dst + 256 * (j + 16 * q16(sl))), make_float4(y0[sl].x, y0[sl].y, y1[sl].x, y1[sl].y));
    FFTC_STAMP(9);
#ifdef CB_FFTC_TIMELINE
    if (threadIdx.x == 0) {
        unsigned smid;
        unsigned long long gt;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        dbg[(size_t)blockIdx.x * 12 + 10] = smid;
        dbg[(size_t)blockIdx.x * 12 + 11] = gt;
    }
#endif
}

#ifdef CB_FFTC_TIMELINE
static unsigned long long *g_fftc_dbg = nullptr;
static int g_fftc_resident = 0;
#endif

template <bool INV, int MODE, int CL>
static int launch(const float2 *in, float2 *out, const float2 *twN, size_t nframes, cudaStream_t s)
{
    auto kern = fft65536_cluster_kernel<INV, MODE, CL>;
    static int resident[2] = {0, 0};  // co-resident clusters on this device (prefetch distance)
    constexpr int SMEM = smem_bytes(CL);
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    if (CL > 8) CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nframes * CL), 1, 1);
    cfg.blockDim = dim3(2048 / CL, 1, 1);
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
        cudaLaunchConfig_t q = cfg;
        q.gridDim = dim3(148 * 4 * CL, 1, 1);
        if (cudaOccupancyMaxActiveClusters(&nc, kern, &q) != cudaSuccess || nc <= 0) {
            cudaGetLastError();
            nc = 32;
        }
        resident[INV] = nc;
    }
#ifdef CB_FFTC_TIMELINE
    g_fftc_resident = resident[INV];
#endif
#ifdef CB_FFTC_TIMELINE
    CB_CUDA(cudaLaunchKernelEx(&cfg, kern, in, out, twN, nframes, (unsigned)resident[INV], g_fftc_dbg));
#else
    CB_CUDA(cudaLaunchKernelEx(&cfg, kern, in, out, twN, nframes, (unsigned)resident[INV]));
#endif
    count_launch();
    return CB_OK;
}

}  // namespace fftc

// twN: 65536 entries e^{-/+ 2 pi i k / 65536} (direction baked in, as FftPlanDev::tw)
int launch_fft65536_cluster(const float2 *in, float2 *out, const float2 *twN, size_t nframes, bool inverse, int tpt,
                            cudaStream_t s)
{
    if (nframes == 0) return CB_OK;
    // grid.x = 8 * frames must stay below 2^31
    const size_t max_frames = (size_t)1 << 26;
    for (size_t done = 0; done < nframes; done += max_frames) {
        const size_t g = nframes - done < max_frames ? nframes - done : max_frames;
        const float2 *gi = in + done * fftc::NF;
        float2 *go = out + done * fftc::NF;
        int rc;
        if (tpt == 1) rc = inverse ? fftc::launch<true, 0, 8>(gi, go, twN, g, s) : fftc::launch<false, 0, 8>(gi, go, twN, g, s);
        else if (tpt == 2) rc = inverse ? fftc::launch<true, 1, 8>(gi, go, twN, g, s) : fftc::launch<false, 1, 8>(gi, go, twN, g, s);
        else if (tpt == 4) rc = inverse ? fftc::launch<true, 2, 16>(gi, go, twN, g, s) : fftc::launch<false, 2, 16>(gi, go, twN, g, s);
        else rc = inverse ? fftc::launch<true, 2, 8>(gi, go, twN, g, s) : fftc::launch<false, 2, 8>(gi, go, twN, g, s);
        if (rc) return rc;
    }
    return CB_OK;
}

}  // namespace cb
