// K5-R: 65536-point FFT / IFFT over a 256 x 256 split: two streaming steps, run either as ONE persistent kernel
// with the intermediate in an L2-resident ring (fft65536_fused_kernel, the default) or as two launches with a
// batch-sized scratch (sm_100a).
//
// Reference semantics (src/fft/mod.rs:73-96): X[k] = sum_n x[n] e^{-/+ j 2 pi k n / N}, unnormalised.
//
//   n = 256 n1 + n2,  k = k1 + 256 k2
//   X[k1 + 256 k2] = sum_{n2} W256^{n2 k2} { W65536^{n2 k1} sum_{n1} x[256 n1 + n2] W256^{n1 k1} }
//
// step A (fft65536_cols_kernel): a CTA takes 16 columns n2 of one frame (16 x 256 points, 35 KiB of shared
//   memory), does the 16 column FFTs of length 256 (two radix-16 passes in registers, one shared-memory
//   exchange) and writes Y[k1][n2] transposed into the scratch: 128 contiguous bytes per half-warp on both sides.
// step B (fft65536_rows_kernel): a CTA takes 16 rows k1 of the scratch (2 KiB contiguous each), applies
//   W^{k1 n2} on the way in, does the 16 row FFTs and scatters X[k1 + 256 k2], again 128 bytes per half-warp.
// Both kernels use 256 threads, 64 registers and one 35 KiB buffer, so four CTAs (32 warps) are resident per SM
// -- the configuration that lets fft2_frames_kernel stream at 7 TB/s; each step runs at ~6.7 TB/s.
// Algorithmic HBM traffic: 8 B read + 8 B written per sample; actual traffic is twice that (the intermediate makes
// one round trip through a scratch as large as the batch), which still beats the one-pass cluster kernel
// (1.29 ms vs 1.55 ms for 2^28 samples) because that one is bound by its DSMEM all-to-all, not by HBM.
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)
#include "fft2_core.cuh"
#include "fft_kernels.cuh"

namespace cb {

namespace fftr {

constexpr int NF = 65536;
constexpr int RP = 273;  // row pitch (float2): pad16(256) = 272, + 1 so that rows fall into different banks
constexpr int SMEM = 16 * RP * (int)sizeof(float2);

__device__ __forceinline__ float2 ld_cs(const float2 *p)
{
    float2 v;
    asm volatile("ld.global.cs.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_cs(float2 *p, float2 v)
{
    asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// thread (c = lane & 15, j = 2 warp + lane / 16): column (or row) c of the CTA's 16, radix-16 task j
template <bool INV>
__global__ void __launch_bounds__(256, 4)
fft65536_cols_kernel(const float2 *__restrict__ in, float2 *__restrict__ mid, const float2 *__restrict__ twN)
{
    using namespace fft2;
    extern __shared__ __align__(16) float2 rsm[];
    const size_t frame = blockIdx.x >> 4;
    const int cb16 = (int)(blockIdx.x & 15) * 16;
    const int lane = threadIdx.x & 31, c = lane & 15, j = 2 * (threadIdx.x >> 5) + (lane >> 4);
    const float2 *src = in + frame * NF + cb16 + c;
    float2 *dst = mid + frame * NF + cb16 + c;
    float2 *row = rsm + c * RP;
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = ld_cs(src + 256 * (j + 16 * m));
    bfly16<INV>(v);
#pragma unroll
    for (int sl = 0; sl < 16; ++sl) row[pad16(16 * j + q16(sl))] = v[sl];
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = row[pad16(j + 16 * m)];
    twiddle16(v, __ldg(twN + 256 * j));
    bfly16<INV>(v);  // slot sl: Y[k1 = j + 16 q16(sl)][n2 = cb16 + c]
#pragma unroll
    for (int sl = 0; sl < 16; ++sl) dst[256 * (j + 16 * q16(sl))] = v[sl];
}

template <bool INV>
__global__ void __launch_bounds__(256, 4)
fft65536_rows_kernel(const float2 *__restrict__ mid, float2 *__restrict__ out, const float2 *__restrict__ twN)
{
    using namespace fft2;
    extern __shared__ __align__(16) float2 rsm[];
    const size_t frame = blockIdx.x >> 4;
    const int rb16 = (int)(blockIdx.x & 15) * 16;
    const int lane = threadIdx.x & 31, lo = lane & 15, hi = 2 * (threadIdx.x >> 5) + (lane >> 4);
    float2 v[16];
    {   // pass 3: thread (n2 digit j = lo, row r = hi): 128 contiguous bytes of one row per half-warp
        const int k1 = rb16 + hi;
        const float2 *src = mid + frame * NF + (size_t)k1 * 256 + lo;
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = src[16 * m];
        twiddle16c(v, __ldg(twN + k1 * lo), __ldg(twN + 16 * k1));  // W^{k1 (lo + 16 m)}
        bfly16<INV>(v);
        float2 *row = rsm + hi * RP;
#pragma unroll
        for (int sl = 0; sl < 16; ++sl) row[pad16(16 * lo + q16(sl))] = v[sl];
    }
    __syncthreads();
    {   // pass 4: thread (row r = lo, task j = hi): the 16 rows of one output index per half-warp
        const float2 *row = rsm + lo * RP;
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = row[pad16(hi + 16 * m)];
        twiddle16(v, __ldg(twN + 256 * hi));
        bfly16<INV>(v);  // slot sl: X[k1 = rb16 + lo + 256 (hi + 16 q16(sl))]
        float2 *dst = out + frame * NF + rb16 + lo;
#pragma unroll
        for (int sl = 0; sl < 16; ++sl) st_cs(dst + 256 * (hi + 16 * q16(sl)), v[sl]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Fused form: ONE persistent kernel runs both steps, step A running `lag` frames ahead of step B, with the
// intermediate in a small ring of scratch frames that stays resident in L2 (written by step A, read by step B a
// few microseconds later, overwritten `ring` frames later while still dirty in L2), so that it never travels to
// HBM.  Work items (16 per frame and step) are taken in a fixed order -- A(0..lag-1), then A(lag+u), B(u)
// alternating -- from a global ticket counter, so an item only ever waits for items with smaller tickets, and a
// ticket is only ever held by a CTA that is running: the earliest unfinished item can always proceed, whatever
// share of the grid is resident (two such kernels on different streams cannot deadlock each other, as a static
// item-to-CTA assignment could).  Per-frame counters in global memory carry the
// dependencies (release: st + __threadfence + atomicAdd; acquire: ld.acquire.gpu polled by warp 0 as a whole, then the
// block barrier; the scratch is read with
// ld.global.cg so that no stale L1 line of an earlier use of the slot is seen).
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// A whole warp waits, converged, until *p >= need (lane 0 reads, the value is broadcast).  A single thread spinning
// ahead of a block barrier while the other 31 lanes of its warp run on was observed to let the barrier release
// without it; keeping the warp together avoids depending on that.
__device__ __forceinline__ void poll_at_least(const unsigned *p, unsigned need, int lane)
{
    for (;;) {
        unsigned f = lane == 0 ? ld_acquire(p) : 0u;
        f = __shfl_sync(0xffffffffu, f, 0);
        if (f >= need) break;
        __nanosleep(64);
    }
}

__device__ __forceinline__ void discard_l2_line(const float2 *p)
{
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}
__device__ __forceinline__ void red_release(unsigned *p)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

struct FusedItem {
    bool valid, is_a;
    unsigned long long frame;
    int part16;
};
__device__ __forceinline__ FusedItem fused_decode(unsigned long long item, unsigned lag, unsigned long long nframes)
{
    FusedItem it;
    const unsigned long long chunk = item >> 4;
    it.part16 = (int)(item & 15) * 16;
    if (chunk < lag) {
        it.is_a = true;
        it.frame = chunk;
    } else {
        const unsigned long long t = chunk - lag;
        it.is_a = (t & 1) == 0;
        it.frame = it.is_a ? lag + (t >> 1) : (t >> 1);
    }
    it.valid = it.frame < nframes;
    return it;
}

// Two of the three L2 round trips that used to sit in series in front of every item are off the critical path: thread
// 0 requests the NEXT ticket at the top of an item (it is only read after the item's closing barrier), and the
// previous item's counter is released by warp 1 once this item's loads are in flight (red.release orders the stores
// of the whole block, which the closing barrier made visible to that thread, before the increment).  A CTA thus holds
// two tickets, so the ring has to cover twice the items in flight: 96 frames (48 MiB) measured best, was 32.
// (Looking at the next item's dependency counter ahead of time as well, or prefetching the next item's points with
// the TMA engine -- K5-R2 below -- measured no better: profiles/r03h_fft65536_variants.txt.)
// IN16: the input is i16 IQ pairs (one 32-bit word per sample, src/io/raw_iq.rs:78-140), widened as in_scale * (i16 as
// f32) in step A's loads
__device__ __forceinline__ float2 ld_cs_iq16(const uint32_t *p, float scale)
{
    uint32_t w;
    asm volatile("ld.global.cs.b32 %0, [%1];" : "=r"(w) : "l"(p));
    return make_float2(__fmul_rn(scale, (float)(int16_t)(w & 0xFFFFu)), __fmul_rn(scale, (float)(int16_t)(w >> 16)));
}

template <bool INV, bool IN16 = false, int MINB = 4>
__global__ void __launch_bounds__(256, MINB)
fft65536_fused_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, float2 *__restrict__ scratch,
                      const float2 *__restrict__ twN, unsigned *ticket, unsigned *flags_a, unsigned *flags_b,
                      unsigned long long nframes, unsigned lag, unsigned ring, float in_scale = 1.f)
{
    using namespace fft2;
    extern __shared__ __align__(16) float2 rsm[];
    __shared__ unsigned s_cur[2];
    const int lane = threadIdx.x & 31, lo = lane & 15, hi = 2 * (threadIdx.x >> 5) + (lane >> 4);
    const unsigned long long total = (lag + 2ull * nframes) * 16;
    if (threadIdx.x == 0) s_cur[0] = atomicAdd(ticket, 1u);
    __syncthreads();
    unsigned *pending = nullptr;  // the counter the previous item still has to bump (block-uniform)
    for (int par = 0;; par ^= 1) {
        const unsigned long long item = s_cur[par];
        if (item >= total) break;
        unsigned nxt = 0;  // thread 0: the next ticket, in flight until the end of the item
        if (threadIdx.x == 0) nxt = atomicAdd(ticket, 1u);
        const FusedItem it = fused_decode(item, lag, nframes);
        const unsigned long long frame = it.frame;
        const int part16 = it.part16;
        float2 *slot = scratch + (frame % ring) * NF;
        float2 v[16];
        if (it.valid && it.is_a) {
            float2 *row = rsm + lo * RP;
            if constexpr (IN16) {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(in) + frame * NF + part16 + lo;
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = ld_cs_iq16(src + 256 * (hi + 16 * m), in_scale);
            } else {
                const float2 *src = in + frame * NF + part16 + lo;
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = ld_cs(src + 256 * (hi + 16 * m));
            }
            if (threadIdx.x == 32 && pending != nullptr) red_release(pending);
            // the slot must have been consumed by its last reader before anything is stored into it: warp 0 polls
            // (all 32 lanes together) while the input loads are in flight; the block barrier below orders the
            // observation before every thread's stores
            if (frame >= ring && threadIdx.x < 32) poll_at_least(flags_b + (frame - ring), 16u, lane);
            bfly16<INV>(v);
#pragma unroll
            for (int sl = 0; sl < 16; ++sl) row[pad16(16 * hi + q16(sl))] = v[sl];
            __syncthreads();
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = row[pad16(hi + 16 * m)];
            twiddle16(v, __ldg(twN + 256 * hi));
            bfly16<INV>(v);
            float2 *dst = slot + part16 + lo;
#pragma unroll
            for (int sl = 0; sl < 16; ++sl) __stcg(dst + 256 * (hi + 16 * q16(sl)), v[sl]);
        } else if (it.valid) {
            if (threadIdx.x == 32 && pending != nullptr) red_release(pending);
            // all 16 column blocks of the frame must be in
            if (threadIdx.x < 32) poll_at_least(flags_a + frame, 16u, lane);
            __syncthreads();
            const int k1 = part16 + hi;
            const float2 *src = slot + (size_t)k1 * 256 + lo;
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = __ldcg(src + 16 * m);
            twiddle16c(v, __ldg(twN + k1 * lo), __ldg(twN + 16 * k1));
            bfly16<INV>(v);
            float2 *rowp = rsm + hi * RP;
#pragma unroll
            for (int sl = 0; sl < 16; ++sl) rowp[pad16(16 * lo + q16(sl))] = v[sl];
            __syncthreads();  // every thread has consumed its scratch reads
            // the 16 rows are dead until step A of frame + ring overwrites them: drop their (dirty) lines from L2 instead
            // of letting them be written back -- one 128-byte line per thread
            discard_l2_line(slot + (size_t)(part16 + (threadIdx.x >> 4)) * 256 + (threadIdx.x & 15) * 16);
            const float2 *row = rsm + lo * RP;
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = row[pad16(hi + 16 * m)];
            twiddle16(v, __ldg(twN + 256 * hi));
            bfly16<INV>(v);
            float2 *dst = out + frame * NF + part16 + lo;
#pragma unroll
            for (int sl = 0; sl < 16; ++sl) st_cs(dst + 256 * (hi + 16 * q16(sl)), v[sl]);
        } else if (threadIdx.x == 32 && pending != nullptr) {
            red_release(pending);
        }
        if (threadIdx.x == 0) s_cur[par ^ 1] = nxt;
        __syncthreads();  // shared-memory reads done, the next ticket published, every thread's stores ordered before the release
        pending = !it.valid ? nullptr : it.is_a ? flags_a + frame : flags_b + frame;
    }
    if (threadIdx.x == 32 && pending != nullptr) red_release(pending);
}

// ---------------------------------------------------------------------------------------------------------
// Prefetching form of the fused kernel (K5-R2).  Same items, tickets, ring and counters; what changes is how an
// item's 32 KiB get on chip: the TMA engine brings them (step A: ONE tensor-map box of 256 rows x 128 bytes of the
// input; step B: 16 bulk copies of one 2 KiB scratch row each), into the buffer the item is then transformed in
// place in, and it does so for the NEXT item while the current one is being computed -- no registers, no load
// instructions and no load latency in the eight computing warps.  Two 33 KiB buffers per CTA, three CTAs per SM.
// A ninth warp does everything that waits on an L2 round trip: the ticket two items ahead, the acquire on the
// next item's dependency counter, the copies, the release of the previous item's counter.  A buffer's mbarrier
// completing means "the item's points are here AND its dependency was met", so the computing warps carry no
// dependency logic at all.  (A first version with per-thread 8-byte cp.async copies lost more in issue time than
// it hid: profiles/r03e_fft65536_prefetch_phases.txt.)
//
// Layouts, chosen so that every thread overwrites exactly the 16 points it has just read (no barrier between the
// read and the write of a pass, no second buffer) and every access is a contiguous 128 bytes per half-warp:
//   step A: [n1][c]  (16 points = 128 B per input row, as the box lands); thread (c = lo, j = hi) reads rows
//           j + 16 m, writes its result k = 16 j + q over row j + 16 q; the second pass reads rows 16 hi + m
//   step B: [k1][n2] with a row pitch of 2064 B; thread (lo, hi) reads n2 = lo + 16 m of row hi, writes result
//           16 lo + q over n2 = lo + 16 q; the second pass reads 16 contiguous points of row lo (8 LDS.128)
// Deadlock freedom as before, with one more rule: the ninth warp only LOOKS at the next item's dependency (if it
// is not met, the item is loaded one round later, after this round's counter has been released, and only then
// with a blocking wait) -- a CTA never waits while it still owes a release.
constexpr int PF_THREADS = 288;
constexpr int BPITCH = 258;                 // float2 per step-B row: 2 KiB + 16 B
constexpr int BUF_BYTES = 16 * BPITCH * 8;  // 33 024 (step A uses the first 32 KiB)
constexpr int SMEM_PF = 2 * BUF_BYTES;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// the input is read once: its lines are marked evict-first so that they do not push the scratch ring out of L2
__device__ __forceinline__ void tma_box_2d(void *dst_smem, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
            smem_u32(dst_smem)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_1d(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#define FFTR_BAR_END() asm volatile("bar.sync 1, 288;" ::: "memory")
#define FFTR_BAR_MID() asm volatile("bar.sync 2, 256;" ::: "memory")

// scripts/fftr_timeline.cu compiles this file with CB_FFTR_STATS: per-CTA clock64 sums of the phases of every item
// (thread 0's view) and counts of prefetched items
#ifdef CB_FFTR_STATS
__device__ unsigned long long *g_fftr_dbg = nullptr;
#define FFTR_T(k)                                    \
    if (threadIdx.x == 0) {                          \
        const long long now_ = clock64();            \
        st_[k] += (unsigned long long)(now_ - t_);   \
        t_ = now_;                                   \
    }
#define FFTR_C(k) if (threadIdx.x == 0) st_[k] += 1ull
#else
#define FFTR_T(k)
#define FFTR_C(k)
#endif

__device__ __forceinline__ const unsigned *pf_dependency(const FusedItem &it, const unsigned *flags_a, const unsigned *flags_b,
                                                         unsigned ring)
{
    if (!it.is_a) return flags_a + it.frame;                      // step B reads what the 16 step-A items of the frame wrote
    return it.frame >= ring ? flags_b + (it.frame - ring) : nullptr;  // step A overwrites the slot of frame - ring
}

// the ninth warp, converged: arm the buffer's mbarrier and start the item's copies (the dependency has been observed
// by lane 0, acquire, and the warp has synchronised since)
__device__ __forceinline__ void pf_issue(const FusedItem &it, const CUtensorMap *in_map, const float2 *scratch, unsigned ring,
                                         unsigned char *buf, uint64_t *bar, int lane)
{
    if (lane == 0) mbar_expect_tx(bar, 32768u);
    __syncwarp();
    if (it.is_a) {
        if (lane == 0) tma_box_2d(buf, in_map, 2 * it.part16, (int)(it.frame * 256), bar);
    } else if (lane < 16) {
        fence_proxy_async_global();  // the scratch was written with ordinary stores; the copies read it through the async proxy
        const float2 *src = scratch + (size_t)((unsigned)it.frame % ring) * NF + (size_t)(it.part16 + lane) * 256;
        tma_bulk_1d(buf + lane * (BPITCH * 8), src, 2048u, bar);
    }
    __syncwarp();
}

template <bool INV>
__global__ void __launch_bounds__(PF_THREADS, 3)
fft65536_pf_kernel(const __grid_constant__ CUtensorMap in_map, float2 *__restrict__ out, float2 *__restrict__ scratch,
                   const float2 *__restrict__ twN, unsigned *ticket, unsigned *flags_a, unsigned *flags_b,
                   unsigned long long nframes, unsigned lag, unsigned ring)
{
    using namespace fft2;
    extern __shared__ __align__(128) unsigned char pf_smem[];
    __shared__ __align__(8) uint64_t full[2];
    // s_ld[b]: the copies of the item that will use buffer b have been started; s_nrdy[b]: the dependency of the item
    // AFTER it was seen complete already
    __shared__ unsigned s_cur[2], s_nx[2], s_ld[2], s_nrdy[2];
    const int lane = threadIdx.x & 31, lo = lane & 15, hi = 2 * (threadIdx.x >> 5) + (lane >> 4);
    const bool scout = threadIdx.x >= 256;
    const unsigned long long total = (lag + 2ull * nframes) * 16;
    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_cur[0] = atomicAdd(ticket, 1u);
        s_nx[0] = atomicAdd(ticket, 1u);
        s_ld[0] = 0u;
        s_nrdy[0] = 0u;
    }
    __syncthreads();
#ifdef CB_FFTR_STATS
    unsigned long long st_[12] = {0};
    long long t_ = clock64();
#endif
    unsigned *pending = nullptr;  // the counter the previous item still has to bump (block-uniform)
    unsigned phbits = 0;          // bit b: the parity the next wait on full[b] uses
    for (int par = 0;; par ^= 1) {
        const unsigned long long item = s_cur[par];
        const unsigned nx = s_nx[par];
        if (item >= total) break;
        const FusedItem it = fused_decode(item, lag, nframes);
        const unsigned long long frame = it.frame;
        unsigned char *buf = pf_smem + par * BUF_BYTES;
        if (scout) {
            // warp-uniform: lane 0 talks to global memory, its findings are broadcast, lanes 0..15 start step-B rows
            const FusedItem ni = fused_decode(nx, lag, nframes);
            const bool nvalid = ni.valid && nx < total;
            if (lane == 0 && pending != nullptr) red_release(pending);
            if (it.valid && !s_ld[par]) {  // not started a round ago: now nothing is owed, so waiting is safe
                const unsigned *dep = pf_dependency(it, flags_a, flags_b, ring);
                if (dep != nullptr) poll_at_least(dep, 16u, lane);
                pf_issue(it, &in_map, scratch, ring, buf, &full[par], lane);
            }
            bool ld = nvalid && s_nrdy[par] != 0u;  // the next item's dependency was seen complete a round ago
            if (ld) pf_issue(ni, &in_map, scratch, ring, pf_smem + (par ^ 1) * BUF_BYTES, &full[par ^ 1], lane);
            unsigned t3 = 0u;
            if (lane == 0) t3 = atomicAdd(ticket, 1u);
            if (nvalid && !ld) {
                const unsigned *dep = pf_dependency(ni, flags_a, flags_b, ring);
                unsigned f = 16u;
                if (dep != nullptr && lane == 0) f = ld_acquire(dep);
                ld = __shfl_sync(0xffffffffu, f, 0) >= 16u;
                if (ld) pf_issue(ni, &in_map, scratch, ring, pf_smem + (par ^ 1) * BUF_BYTES, &full[par ^ 1], lane);
            }
            if (lane == 0) {
                // a look at the dependency of the item after the next one, so that the next round starts with the answer
                const FusedItem n3 = fused_decode(t3, lag, nframes);
                unsigned r3 = 0u;
                if (n3.valid && t3 < total) {
                    const unsigned *dep3 = pf_dependency(n3, flags_a, flags_b, ring);
                    r3 = dep3 == nullptr || ld_acquire(dep3) >= 16u ? 1u : 0u;
                }
                s_cur[par ^ 1] = nx;
                s_nx[par ^ 1] = t3;
                s_ld[par ^ 1] = ld ? 1u : 0u;
                s_nrdy[par ^ 1] = r3;
            }
            __syncwarp();
        } else if (it.valid) {
            FFTR_C(6);
#ifdef CB_FFTR_STATS
            if (threadIdx.x == 0 && s_ld[par]) st_[7] += 1ull;
#endif
            const int part16 = it.part16;
            float2 v[16];
            mbar_wait(&full[par], (phbits >> par) & 1u);
            phbits ^= 1u << par;
            FFTR_T(1);
            if (it.is_a) {
                float2 *sa = reinterpret_cast<float2 *>(buf) + lo;
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = sa[(hi + 16 * m) * 16];
                bfly16<INV>(v);
#pragma unroll
                for (int sl = 0; sl < 16; ++sl) sa[(hi + 16 * q16(sl)) * 16] = v[sl];
                FFTR_T(2);
                FFTR_BAR_MID();
                FFTR_T(3);
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = sa[(16 * hi + m) * 16];
                twiddle16(v, __ldg(twN + 256 * hi));
                bfly16<INV>(v);
                float2 *dst = scratch + (size_t)((unsigned)frame % ring) * NF + part16 + lo;
                FFTR_T(8);
#pragma unroll
                for (int sl = 0; sl < 16; ++sl) __stcg(dst + 256 * (hi + 16 * q16(sl)), v[sl]);
            } else {
                const int k1 = part16 + hi;
                // the rows are on chip and dead in the scratch until frame + ring overwrites them: drop their dirty L2 lines
                discard_l2_line(scratch + (size_t)((unsigned)frame % ring) * NF + (size_t)(part16 + (threadIdx.x >> 4)) * 256 +
                                (threadIdx.x & 15) * 16);
                float2 *sb = reinterpret_cast<float2 *>(buf) + hi * BPITCH + lo;
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = sb[16 * m];
                twiddle16c(v, __ldg(twN + k1 * lo), __ldg(twN + 16 * k1));
                bfly16<INV>(v);
#pragma unroll
                for (int sl = 0; sl < 16; ++sl) sb[16 * q16(sl)] = v[sl];
                FFTR_T(2);
                FFTR_BAR_MID();
                FFTR_T(3);
                const float4 *sr = reinterpret_cast<const float4 *>(reinterpret_cast<const float2 *>(buf) + lo * BPITCH + 16 * hi);
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const float4 t = sr[m];
                    v[2 * m] = make_float2(t.x, t.y);
                    v[2 * m + 1] = make_float2(t.z, t.w);
                }
                twiddle16(v, __ldg(twN + 256 * hi));
                bfly16<INV>(v);
                float2 *dst = out + frame * NF + part16 + lo;
                FFTR_T(8);
#pragma unroll
                for (int sl = 0; sl < 16; ++sl) st_cs(dst + 256 * (hi + 16 * q16(sl)), v[sl]);
            }
            // this thread's shared-memory accesses come before the copies that will overwrite the buffer through the
            // async proxy (the scratch stores of step A are ordered on the reading side: acquire, then a proxy fence)
            fence_proxy_async_smem();
            FFTR_T(4);
        }
        FFTR_BAR_END();  // buffer reads done, the next item's words published, every store ordered before the release
        FFTR_T(5);
        pending = !it.valid ? nullptr : it.is_a ? flags_a + frame : flags_b + frame;
    }
    if (threadIdx.x == 256 && pending != nullptr) red_release(pending);
#ifdef CB_FFTR_STATS
    if (threadIdx.x == 0 && g_fftr_dbg != nullptr)
        for (int k = 0; k < 12; ++k) g_fftr_dbg[blockIdx.x * 12 + k] = st_[k];
#endif
}

// input viewed as a 2-D f32 tensor: 512 floats per row (one n1), 256 * nframes rows; box = 32 floats x 256 rows
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

static bool pf_usable(const float2 *in) { return (reinterpret_cast<uintptr_t>(in) & 15) == 0 && encode_fn() != nullptr; }

template <bool INV>
static int launch_fused_pf(const FftPlanDev &p, const float2 *in, float2 *out, size_t nframes, cudaStream_t s)
{
    auto kf = fft65536_pf_kernel<INV>;
    CUtensorMap map;
    {
        const cuuint64_t dims[2] = {512, (cuuint64_t)256 * nframes};
        const cuuint64_t strides[1] = {2048};
        const cuuint32_t box[2] = {32, 256}, estr[2] = {1, 1};
        const CUresult r = encode_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float2 *>(in), dims, strides, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CB_REQUIRE(r == CUDA_SUCCESS, CB_ERR_CUDA, "fft: cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    CB_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_PF));
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kf, PF_THREADS, SMEM_PF);
    if (per_sm < 1) per_sm = 1;
    const unsigned ring = (unsigned)p.scratch_frames;
    unsigned lag = ring / 2;
    if (lag < 1) lag = 1;
    const unsigned long long items = (lag + 2ull * nframes) * 16;
    const unsigned long long cap = (unsigned long long)sms * per_sm;  // one CTA per resident slot
    const unsigned grid = (unsigned)(items < cap ? items : cap);
    CB_CUDA(cudaMemsetAsync(p.flags, 0, (4 + 2 * nframes) * sizeof(unsigned), s));
    kf<<<grid, PF_THREADS, SMEM_PF, s>>>(map, out, p.scratch, p.tw, p.flags, p.flags + 4, p.flags + 4 + nframes, nframes, lag, ring);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

template <bool INV, bool IN16, int MINB>
static int launch_fused_m(const FftPlanDev &p, const float2 *in, float2 *out, size_t nframes, cudaStream_t s, float in_scale)
{
    auto kf = fft65536_fused_kernel<INV, IN16, MINB>;
    CB_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kf, 256, SMEM);
    if (per_sm < 1) per_sm = 1;
    const unsigned ring = (unsigned)p.scratch_frames;
    unsigned lag = ring / 2;
    if (lag < 1) lag = 1;
    const unsigned long long items = (lag + 2ull * nframes) * 16;
    const unsigned long long cap = (unsigned long long)sms * per_sm;  // one CTA per resident slot
    const unsigned grid = (unsigned)(items < cap ? items : cap);
    CB_CUDA(cudaMemsetAsync(p.flags, 0, (4 + 2 * nframes) * sizeof(unsigned), s));
    kf<<<grid, 256, SMEM, s>>>(in, out, p.scratch, p.tw, p.flags, p.flags + 4, p.flags + 4 + nframes, nframes, lag, ring, in_scale);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// CTAs per SM: the kernel compiles to 48 registers without spills when asked to (64 when not), so five CTAs (40 warps)
// fit beside their 35 KiB buffers -- and are slower (1.128 vs 1.086 ms; six spill: 1.81 ms): more warps in flight do not
// help a kernel that waits for the memory system.  Four it is; COMMS_B200_FFT_CTAS = 4 | 5 | 6 for comparison.
template <bool INV, bool IN16 = false>
static int launch_fused(const FftPlanDev &p, const float2 *in, float2 *out, size_t nframes, cudaStream_t s, float in_scale = 1.f)
{
    static const int ctas = [] { const char *e = getenv("COMMS_B200_FFT_CTAS"); return e ? atoi(e) : 4; }();
    if (ctas == 5) return launch_fused_m<INV, IN16, 5>(p, in, out, nframes, s, in_scale);
    if (ctas == 6) return launch_fused_m<INV, IN16, 6>(p, in, out, nframes, s, in_scale);
    return launch_fused_m<INV, IN16, 4>(p, in, out, nframes, s, in_scale);
}

// Frames are processed in groups of p.scratch_frames (the caller sizes the scratch to the whole batch when it
// can: one group = two launches measured fastest; L2-sized groups, with or without overlapping step B of one
// group with step A of the next on a second stream, or a persisting-L2 window on the scratch, all measured
// slower -- the bubbles between dependent launches cost more than the saved HBM round trip).
template <bool INV>
static int launch(const FftPlanDev &p, const float2 *in, float2 *out, size_t nframes, cudaStream_t s)
{
    auto ka = fft65536_cols_kernel<INV>;
    auto kb = fft65536_rows_kernel<INV>;
    CB_CUDA(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    CB_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const size_t gf = p.scratch_frames ? p.scratch_frames : 1;
    for (size_t done = 0; done < nframes; done += gf) {
        const size_t n = nframes - done < gf ? nframes - done : gf;
        ka<<<(unsigned)(n * 16), 256, SMEM, s>>>(in + done * NF, p.scratch, p.tw);
        count_launch();
        kb<<<(unsigned)(n * 16), 256, SMEM, s>>>(p.scratch, out + done * NF, p.tw);
        count_launch();
    }
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

}  // namespace fftr

int launch_fft65536_rows_iq16(const FftPlanDev &p, const uint32_t *in, float in_scale, float2 *out, size_t nframes, cudaStream_t s)
{
    if (nframes == 0) return CB_OK;
    const float2 *in2 = reinterpret_cast<const float2 *>(in);  // the kernel reads it as 32-bit words
    return p.inverse ? fftr::launch_fused<true, true>(p, in2, out, nframes, s, in_scale)
                     : fftr::launch_fused<false, true>(p, in2, out, nframes, s, in_scale);
}

int launch_fft65536_rows(const FftPlanDev &p, const float2 *in, float2 *out, size_t nframes, cudaStream_t s)
{
    if (nframes == 0) return CB_OK;
    if (p.cluster_tpt == 10 && p.flags != nullptr && p.flags_frames >= nframes && p.scratch_frames >= 4 && nframes < (1ull << 23) &&
        fftr::pf_usable(in))
        return p.inverse ? fftr::launch_fused_pf<true>(p, in, out, nframes, s) : fftr::launch_fused_pf<false>(p, in, out, nframes, s);
    if (p.cluster_tpt == 6 && p.flags != nullptr && p.flags_frames >= nframes && p.scratch_frames >= 4)
        return p.inverse ? fftr::launch_fused<true>(p, in, out, nframes, s) : fftr::launch_fused<false>(p, in, out, nframes, s);
    return p.inverse ? fftr::launch<true>(p, in, out, nframes, s) : fftr::launch<false>(p, in, out, nframes, s);
}

}  // namespace cb
