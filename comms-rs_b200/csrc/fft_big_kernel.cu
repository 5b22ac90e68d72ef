// K5-B: FFT / IFFT of 2^15 and 2^17 .. 2^20 points (and, for cross-checking, 2^16) over an N1 x N2 split: the fused
// two-step form of fft_rows_kernel.cu (K5-R) with the step transforms taken from the radix-16 core (fft2_core.cuh),
// so that any N1, N2 in 128 .. 1024 works (sm_100a).
//
// Reference semantics (src/fft/mod.rs:73-96): X[k] = sum_n x[n] e^{-/+ j 2 pi k n / N}, unnormalised.
//
//   n = N2 n1 + n2,  k = k1 + N1 k2
//   X[k1 + N1 k2] = sum_{n2} W_N2^{n2 k2} { W_N^{n2 k1} sum_{n1} x[N2 n1 + n2] W_N1^{n1 k1} }
//
// step A, one item = GA adjacent columns n2 of one frame: N1-point transforms down the columns (thread = column
//   tid % GA, butterfly index tid / GA: a warp touches GA * 8 contiguous bytes of GA / 16 ... rows), Y[k1][n2]
//   stored in place into a scratch frame;
// step B, one item = GB adjacent rows k1 of the scratch frame: W_N^{k1 n2} applied on the way in (lanes along the
//   row), N2-point transforms along the rows, last pass with lanes across the rows so that X[k1 + N1 k2] leaves in
//   GB * 8 contiguous bytes.
// GA * N1 / 16 = GB * N2 / 16 = threads per CTA (256 up to 2^17, 512 for 2^18 / 2^19, 1024 for 2^20), one in-place
// shared-memory buffer of 35 / 70 / 139 KiB, 64 registers: 32 resident warps per SM.
// One persistent kernel runs both steps: items are handed out by a global ticket counter in the order
// A(0 .. lag-1), then A(lag + u), B(u) alternating, per-frame counters carry the dependencies, and the intermediate
// lives in a ring of scratch frames small enough to stay in L2 (see fft_rows_kernel.cu for why this cannot deadlock).
// Algorithmic HBM traffic: 8 B read + 8 B written per sample; SM <-> L2 traffic is twice that.
#include "fft2_core.cuh"
#include "fft_kernels.cuh"

#include <cstdlib>

namespace cb {

namespace fftb {

template <int L1, int L2, int NT_>
struct Geo {
    using PA = fft2::Plan<L1>;
    using PB = fft2::Plan<L2>;
    static constexpr int N1 = 1 << L1, N2 = 1 << L2;
    static constexpr size_t N = (size_t)N1 * N2;
    static constexpr int NT = NT_;
    static constexpr int GA = NT / PA::T, GB = NT / PB::T;        // columns per step-A item, rows per step-B item
    static constexpr int RPA = PA::PADN + 1, RPB = PB::PADN + 1;  // odd pitches: rows fall into different banks
    static constexpr int WORDS = GA * RPA > GB * RPB ? GA * RPA : GB * RPB;
    static constexpr int SMEM = WORDS * (int)sizeof(float2);
    static constexpr int ITEMS = N2 / GA;  // items per frame and step
    static constexpr int MINB = 1024 / NT;
    static_assert(N1 / GB == ITEMS, "both steps must split a frame into the same number of items");
    static_assert(GA >= 4 && GB >= 4, "at least 32 contiguous bytes per access");
    static_assert(PA::PASSES >= 2 && PB::PASSES >= 2, "step transforms exchange through shared memory");
};

struct Sm {
    float2 *base;
    __device__ __forceinline__ float2 ld(int i) const { return base[fft2::pad16(i)]; }
    __device__ __forceinline__ void st(int i, float2 v) const { base[fft2::pad16(i)] = v; }
};

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
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release(unsigned *p)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
// a whole converged warp waits until *p >= need (see fft_rows_kernel.cu)
__device__ __forceinline__ void poll_at_least(const unsigned *p, unsigned need, int lane)
{
    for (;;) {
        unsigned f = lane == 0 ? ld_acquire(p) : 0u;
        f = __shfl_sync(0xffffffffu, f, 0);
        if (f >= need) break;
        __nanosleep(64);
    }
}

// passes PASS .. last of an L-point transform whose frame sits in `buf` (in place: a barrier between the loads and
// the stores of a middle pass); the last pass hands its results to gst(index, value)
template <int L, bool INV, int PASS, bool SYNC_FIRST, typename GST>
__device__ __forceinline__ void later_passes(int j, const float2 *tw, GST gst, float2 *buf)
{
    using PL = fft2::Plan<L>;
    if constexpr (PASS < PL::PASSES) {
        if constexpr (SYNC_FIRST) __syncthreads();
        Sm sm{buf};
        auto mid = [] { __syncthreads(); };
        auto gld = [](int) { return make_float2(0.f, 0.f); };
        fft2::run_pass<L, INV, PASS>(j, tw, gld, gst, sm, sm, mid);
        later_passes<L, INV, PASS + 1, true>(j, tw, gst, buf);
    }
}

template <int L1, int L2, int NT, bool INV>
__global__ void __launch_bounds__(NT, 1024 / NT)
fft_big_fused_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, float2 *__restrict__ scratch,
                     const float2 *__restrict__ twN, const float2 *__restrict__ twA, const float2 *__restrict__ twB,
                     unsigned *ticket, unsigned *flags_a, unsigned *flags_b, unsigned long long nframes, unsigned lag,
                     unsigned ring)
{
    using namespace fft2;
    using G = Geo<L1, L2, NT>;
    extern __shared__ __align__(16) float2 bsm[];
    __shared__ unsigned s_cur[2];
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned long long total = (lag + 2ull * nframes) * G::ITEMS;
    // as in fft65536_fused_kernel: the next ticket is requested at the top of an item and read after its closing
    // barrier; the previous item's counter is released by warp 1 (red.release, after that barrier) once this item's
    // loads have been issued
    if (tid == 0) s_cur[0] = atomicAdd(ticket, 1u);
    __syncthreads();
    unsigned *pending = nullptr;
    for (int par = 0;; par ^= 1) {
        const unsigned long long item = s_cur[par];
        if (item >= total) break;
        unsigned nxt = 0;
        if (tid == 0) nxt = atomicAdd(ticket, 1u);
        const unsigned long long chunk = item / G::ITEMS;
        const int part = (int)(item % G::ITEMS);
        bool is_a;
        unsigned long long frame;
        if (chunk < lag) {
            is_a = true;
            frame = chunk;
        } else {
            const unsigned long long t = chunk - lag;
            is_a = (t & 1) == 0;
            frame = is_a ? lag + (t >> 1) : (t >> 1);
        }
        const bool valid = frame < nframes;
        float2 *slot = scratch + (frame % ring) * G::N;
        if (valid && is_a) {
            const int c = tid % G::GA, j = tid / G::GA;
            const float2 *src = in + frame * G::N + part * G::GA + c;
            float2 *dst = slot + part * G::GA + c;
            float2 *buf = bsm + c * G::RPA;
            Sm sm{buf};
            auto nomid = [] {};
            auto gld = [&](int i) { return ld_cs(src + (size_t)i * G::N2); };
            auto gst = [&](int i, float2 v) { __stcg(dst + (size_t)i * G::N2, v); };
            run_pass<L1, INV, 0>(j, twA, gld, gst, sm, sm, nomid);
            if (tid == 32 && pending != nullptr) red_release(pending);
            // the slot must have been consumed by its last reader before anything is stored into it; the barrier
            // that opens the next pass orders warp 0's observation before every thread's stores
            if (frame >= ring && tid < 32) poll_at_least(flags_b + (frame - ring), (unsigned)G::ITEMS, lane);
            later_passes<L1, INV, 1, true>(j, twA, gst, buf);
        } else if (valid) {
            if (tid == 32 && pending != nullptr) red_release(pending);
            if (tid < 32) poll_at_least(flags_a + frame, (unsigned)G::ITEMS, lane);  // every column block is in
            __syncthreads();
            constexpr int T2 = G::PB::T;
            {   // pass 0, lanes along the row: thread (row tid / T2, butterfly index tid % T2)
                const int r = tid / T2, j = tid % T2;
                const int k1 = part * G::GB + r;
                const float2 *src = slot + (size_t)k1 * G::N2 + j;
                float2 v[16];
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = __ldcg(src + m * T2);
                twiddle16c(v, __ldg(twN + (size_t)k1 * j), __ldg(twN + (size_t)k1 * T2));  // W_N^{k1 (j + m T2)}
                bfly16<INV>(v);
                float2 *row = bsm + r * G::RPB;
#pragma unroll
                for (int sl = 0; sl < 16; ++sl) row[pad16(16 * j + q16(sl))] = v[sl];
            }
            __syncthreads();  // every thread has consumed its scratch reads
            {   // the GB rows are dead until step A of frame + ring overwrites them: drop their dirty lines from L2 instead
                // of letting them be written back (GB * N2 = 16 NT points = NT lines of 128 bytes, one per thread)
                constexpr int LPR = G::N2 / 16;  // lines per row
                const float2 *line = slot + (size_t)(part * G::GB + tid / LPR) * G::N2 + (tid % LPR) * 16;
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(line) : "memory");
            }
            {   // later passes, lanes across the rows: thread (row tid % GB, butterfly index tid / GB)
                const int r = tid % G::GB, j = tid / G::GB;
                float2 *dst = out + frame * G::N + part * G::GB + r;
                auto gst = [&](int i, float2 v) { st_cs(dst + (size_t)i * G::N1, v); };
                later_passes<L2, INV, 1, false>(j, twB, gst, bsm + r * G::RPB);
            }
        } else if (tid == 32 && pending != nullptr) {
            red_release(pending);
        }
        if (tid == 0) s_cur[par ^ 1] = nxt;
        __syncthreads();  // shared-memory reads done, the next ticket published, every thread's stores ordered before the release
        pending = !valid ? nullptr : is_a ? flags_a + frame : flags_b + frame;
    }
    if (tid == 32 && pending != nullptr) red_release(pending);
}

template <int L1, int L2, int NT, bool INV>
static int launch(const FftPlanDev &p, const float2 *in, float2 *out, size_t nframes, cudaStream_t s)
{
    using G = Geo<L1, L2, NT>;
    auto kf = fft_big_fused_kernel<L1, L2, NT, INV>;
    CB_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kf, G::NT, G::SMEM);
    if (per_sm < 1) per_sm = 1;
    const unsigned ring = (unsigned)p.scratch_frames;
    unsigned lag = ring / 2;
    if (lag < 1) lag = 1;
    const unsigned long long items = (lag + 2ull * nframes) * G::ITEMS;
    const unsigned long long cap = (unsigned long long)sms * per_sm;  // one CTA per resident slot
    const unsigned grid = (unsigned)(items < cap ? items : cap);
    CB_CUDA(cudaMemsetAsync(p.flags, 0, (4 + 2 * nframes) * sizeof(unsigned), s));
    kf<<<grid, G::NT, G::SMEM, s>>>(in, out, p.scratch, p.tw, p.tw16a, p.tw16b, p.flags, p.flags + 4,
                                    p.flags + 4 + nframes, nframes, lag, ring);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

template <bool INV>
static int launch_dir(const FftPlanDev &p, const float2 *in, float2 *out, size_t nframes, cudaStream_t s)
{
    // threads per CTA: wider items (more contiguous bytes per access) against more resident CTAs per SM; measured best
    // per size on 2^28 samples (COMMS_B200_FFT_BIG_NT = 256 | 512 | 1024 overrides, for experiments)
    static const int force = [] { const char *e = getenv("COMMS_B200_FFT_BIG_NT"); return e ? atoi(e) : 0; }();
    const int l = p.log2n1 + p.log2n2;
    const int nt = (force == 256 || force == 512 || force == 1024) && l >= 17 ? force : (l <= 17 ? 256 : (l <= 19 ? 512 : 1024));
#define CB_BIG(L1, L2)                                                                  \
    nt == 256 ? launch<L1, L2, 256, INV>(p, in, out, nframes, s)                        \
              : (nt == 512 ? launch<L1, L2, 512, INV>(p, in, out, nframes, s) : launch<L1, L2, 1024, INV>(p, in, out, nframes, s))
    switch (l) {
    case 15: return launch<7, 8, 256, INV>(p, in, out, nframes, s);
    case 16: return launch<8, 8, 256, INV>(p, in, out, nframes, s);
    case 17: return CB_BIG(8, 9);
    case 18: return CB_BIG(9, 9);
    case 19: return CB_BIG(9, 10);
    case 20: return CB_BIG(10, 10);
#undef CB_BIG
    default: set_error("fft: no fused two-step kernel for 2^%d", p.log2n1 + p.log2n2); return CB_ERR_UNSUPPORTED;
    }
}

}  // namespace fftb

bool fft_big_applicable(const FftPlanDev &p, size_t nframes)
{
    const int l = p.log2n1 + p.log2n2;
    return p.kind == FFT_FOURSTEP && p.big && l >= 15 && l <= 20 && p.log2n1 == l / 2 && p.tw16a && p.tw16b && p.flags &&
           p.flags_frames >= nframes && p.scratch_frames >= 2;
}

int launch_fft_big(const FftPlanDev &p, const float2 *in, float2 *out, size_t nframes, cudaStream_t s)
{
    if (nframes == 0) return CB_OK;
    return p.inverse ? fftb::launch_dir<true>(p, in, out, nframes, s) : fftb::launch_dir<false>(p, in, out, nframes, s);
}

}  // namespace cb
