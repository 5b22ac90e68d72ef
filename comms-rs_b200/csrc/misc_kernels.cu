// Mixer (K6), FM demod, stand-alone resamplers, bit-exact edge stages and the
// synthetic-input generator.  All HBM-bound element-wise kernels: 128-bit
// coalesced accesses, grid-stride over a grid sized in multiples of the SM count.
#include "misc_kernels.cuh"

namespace cb {

static inline unsigned grid_for(size_t work_items, unsigned per_block)
{
    size_t b = ceil_div(work_items ? work_items : 1, (size_t)per_block);
    const size_t cap = 148 * 8;
    return (unsigned)(b < cap ? b : cap);
}

// ---------------------------------------------------------------- mixer
// y[n] = x[n] * exp(j*(phase0 + n*dphase))  (src/mixer.rs:73-84), phase in f64.
__global__ void __launch_bounds__(256)
mixer_kernel(const float2 *__restrict__ x, float2 *__restrict__ y, size_t n, double phase0, double dphase)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 31) == 0;
    size_t tail = 0;
    if (vec) {
        // four samples per thread: one 256-bit load and store; the f64 phase is evaluated for the first sample of
        // the quad, the other three follow by e^{j dphase}, e^{2j dphase}, e^{3j dphase} (each one product away)
        const float2 s1 = phase_rotation(dphase), s2 = phase_rotation(2.0 * dphase), s3 = phase_rotation(3.0 * dphase);
        const size_t nquad = n >> 2;
        for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquad; q += stride) {
            float v[8], o[8];
            ldg_stream8(reinterpret_cast<const float *>(x + 4 * q), v);
            const float2 r0 = phase_rotation(fma((double)(4 * q), dphase, phase0));
            const float2 r1 = cmul(r0, s1), r2 = cmul(r0, s2), r3 = cmul(r0, s3);
            const float2 a = cmul(make_float2(v[0], v[1]), r0), b = cmul(make_float2(v[2], v[3]), r1);
            const float2 c = cmul(make_float2(v[4], v[5]), r2), d = cmul(make_float2(v[6], v[7]), r3);
            o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y; o[6] = d.x; o[7] = d.y;
            stg_stream8(reinterpret_cast<float *>(y + 4 * q), o);
        }
        tail = nquad << 2;
    }
    for (size_t i = tail + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = cmul(x[i], phase_rotation(fma((double)i, dphase, phase0)));
}

int launch_mixer(const float2 *x, float2 *y, size_t n, double phase0, double dphase, cudaStream_t s)
{
    if (n == 0) return CB_OK;
    mixer_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, s>>>(x, y, n, phase0, dphase);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// ---------------------------------------------------------------- FM demod
// out[n] = atan2(Im, Re)(x[n] * conj(x[n-1]))  (src/modulation/analog.rs:22-34).
// The product is formed with individually rounded operations in the
// reference's operand order so that signed zeros (first sample: prev = 0)
// select the same atan2 branch.
// Four samples per thread (one 256-bit load, one 128-bit store: 64 KiB of reads in flight per SM at full occupancy
// and 25 registers); the sample before a thread's quad comes from the neighbouring lane by shuffle (lane 0 re-reads
// it), the angle from the branch-free atan2 of misc_kernels.cuh.
__global__ void __launch_bounds__(256)
fm_kernel(const float2 *__restrict__ x, float *__restrict__ out, size_t n, const float2 *prev_in, float2 *prev_out)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) & 31) | (reinterpret_cast<uintptr_t>(out) & 15)) == 0;
    size_t tail = 0;  // first sample not covered by the vector loop
    if (vec) {
        const size_t nquad = n >> 2;
        const size_t nquad_pad = (nquad + 31) & ~(size_t)31;  // whole warps iterate together (shuffles)
        for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquad_pad; q += stride) {
            const bool live = q < nquad;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (live) ldg_stream8(reinterpret_cast<const float *>(x + 4 * q), v);
            float2 pv = make_float2(__shfl_up_sync(0xffffffffu, v[6], 1), __shfl_up_sync(0xffffffffu, v[7], 1));
            if ((threadIdx.x & 31) == 0 && live) pv = q ? x[4 * q - 1] : *prev_in;
            if (live) {
                const float2 s0 = make_float2(v[0], v[1]), s1 = make_float2(v[2], v[3]);
                const float2 s2 = make_float2(v[4], v[5]), s3 = make_float2(v[6], v[7]);
                stg_stream(reinterpret_cast<float4 *>(out) + q,
                           make_float4(fm_angle_fast(s0, pv), fm_angle_fast(s1, s0), fm_angle_fast(s2, s1), fm_angle_fast(s3, s2)));
            }
        }
        tail = nquad << 2;
    }
    // unaligned buffers, and the up to three samples behind the last quad
    for (size_t i = tail + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = fm_angle_fast(x[i], i ? x[i - 1] : *prev_in);
    if (blockIdx.x == 0 && threadIdx.x == 0) *prev_out = n ? x[n - 1] : *prev_in;
}

int launch_fm(const float2 *x, float *out, size_t n, const float2 *prev_in, float2 *prev_out, cudaStream_t s)
{
    fm_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, s>>>(x, out, n, prev_in, prev_out);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// ---------------------------------------------------------------- resamplers
template <typename T>
__global__ void __launch_bounds__(256) decimate_kernel(const T *__restrict__ in, T *__restrict__ out, size_t n_out, size_t rate)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) out[i] = in[i * rate];
}

template <typename T>
__global__ void __launch_bounds__(256) upsample_kernel(const T *__restrict__ in, T *__restrict__ out, size_t n_out, size_t rate)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    T zero;
    memset(&zero, 0, sizeof(T));
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) {
        const size_t q = i / rate;
        out[i] = (i - q * rate == 0) ? in[q] : zero;
    }
}

int launch_decimate(const void *in, void *out, size_t n_out, size_t elem, size_t rate, cudaStream_t s)
{
    if (n_out == 0) return CB_OK;
    const unsigned g = grid_for(n_out, 256);
    switch (elem) {
    case 4: decimate_kernel<float><<<g, 256, 0, s>>>((const float *)in, (float *)out, n_out, rate); break;
    case 8: decimate_kernel<float2><<<g, 256, 0, s>>>((const float2 *)in, (float2 *)out, n_out, rate); break;
    case 16: decimate_kernel<float4><<<g, 256, 0, s>>>((const float4 *)in, (float4 *)out, n_out, rate); break;
    case 2: decimate_kernel<short><<<g, 256, 0, s>>>((const short *)in, (short *)out, n_out, rate); break;
    case 1: decimate_kernel<char><<<g, 256, 0, s>>>((const char *)in, (char *)out, n_out, rate); break;
    default: set_error("decimate: unsupported element size %zu", elem); return CB_ERR_UNSUPPORTED;
    }
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

int launch_upsample(const void *in, void *out, size_t n_out, size_t elem, size_t rate, cudaStream_t s)
{
    if (n_out == 0) return CB_OK;
    const unsigned g = grid_for(n_out, 256);
    switch (elem) {
    case 4: upsample_kernel<float><<<g, 256, 0, s>>>((const float *)in, (float *)out, n_out, rate); break;
    case 8: upsample_kernel<float2><<<g, 256, 0, s>>>((const float2 *)in, (float2 *)out, n_out, rate); break;
    case 16: upsample_kernel<float4><<<g, 256, 0, s>>>((const float4 *)in, (float4 *)out, n_out, rate); break;
    case 2: upsample_kernel<short><<<g, 256, 0, s>>>((const short *)in, (short *)out, n_out, rate); break;
    case 1: upsample_kernel<char><<<g, 256, 0, s>>>((const char *)in, (char *)out, n_out, rate); break;
    default: set_error("upsample: unsupported element size %zu", elem); return CB_ERR_UNSUPPORTED;
    }
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// ---------------------------------------------------------------- edges
__global__ void __launch_bounds__(256)
bits_to_symbols_kernel(const uint8_t *__restrict__ bits, float2 *__restrict__ sym, size_t nsym, int mode)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nsym; i += stride) {
        if (mode == 0) sym[i] = make_float2((float)bits[i] * 2.0f - 1.0f, 0.0f);
        else sym[i] = make_float2((float)bits[2 * i] * 2.0f - 1.0f, (float)bits[2 * i + 1] * 2.0f - 1.0f);
    }
}

int launch_bits_to_symbols(const uint8_t *bits, float2 *sym, size_t nsym, int mode, cudaStream_t s)
{
    if (nsym == 0) return CB_OK;
    bits_to_symbols_kernel<<<grid_for(nsym, 256), 256, 0, s>>>(bits, sym, nsym, mode);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// Rust `as i16`: truncate toward zero, saturate, NaN -> 0
__device__ __forceinline__ int quant_i16(float v, float scale)
{
    int q = __float2int_rz(__fmul_rn(scale, v));
    return q > 32767 ? 32767 : (q < -32768 ? -32768 : q);
}

// eight floats per thread (256-bit load, 128-bit store) when the buffers allow it
__global__ void __launch_bounds__(256)
quantize_i16_kernel(const float *__restrict__ in, int16_t *__restrict__ out, size_t n, float scale)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t tail = 0;
    if (((reinterpret_cast<uintptr_t>(in) & 31) | (reinterpret_cast<uintptr_t>(out) & 15)) == 0) {
        const size_t n8 = n >> 3;
        for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += stride) {
            float v[8];
            ldg_stream8(in + 8 * g, v);
            uint4 o;
            o.x = (uint32_t)(quant_i16(v[0], scale) & 0xFFFF) | ((uint32_t)quant_i16(v[1], scale) << 16);
            o.y = (uint32_t)(quant_i16(v[2], scale) & 0xFFFF) | ((uint32_t)quant_i16(v[3], scale) << 16);
            o.z = (uint32_t)(quant_i16(v[4], scale) & 0xFFFF) | ((uint32_t)quant_i16(v[5], scale) << 16);
            o.w = (uint32_t)(quant_i16(v[6], scale) & 0xFFFF) | ((uint32_t)quant_i16(v[7], scale) << 16);
            reinterpret_cast<uint4 *>(out)[g] = o;
        }
        tail = n8 << 3;
    }
    for (size_t i = tail + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (int16_t)quant_i16(in[i], scale);
}

int launch_quantize_i16(const float *in, int16_t *out, size_t n, float scale, cudaStream_t s)
{
    if (n == 0) return CB_OK;
    quantize_i16_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, s>>>(in, out, n, scale);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// ---------------------------------------------------------------- IQ edge formats
// u8 offset binary (RTL-SDR; examples/fm_radio.rs:84-87): (x as f32 - 127.5) / 127.5, both operations exactly rounded
__device__ __forceinline__ float conv_u8(uint32_t b) { return __fdiv_rn(__fsub_rn((float)b, 127.5f), 127.5f); }

// eight bytes per thread (64-bit load, 256-bit store) when the buffers allow it
__global__ void __launch_bounds__(256)
convert_u8_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t tail = 0;
    if (((reinterpret_cast<uintptr_t>(in) & 7) | (reinterpret_cast<uintptr_t>(out) & 31)) == 0) {
        const size_t n8 = n >> 3;
        for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += stride) {
            const uint2 w = reinterpret_cast<const uint2 *>(in)[g];
            float v[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[k] = conv_u8((w.x >> (8 * k)) & 255u);
                v[4 + k] = conv_u8((w.y >> (8 * k)) & 255u);
            }
            stg_stream8(out + 8 * g, v);
        }
        tail = n8 << 3;
    }
    for (size_t i = tail + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = conv_u8(in[i]);
}

// interleaved i16 IQ (src/io/raw_iq.rs:20-140): scale * (x as f32); scale = 1 is the plain cast
// eight values per thread (128-bit load, 256-bit store) when the buffers allow it
__global__ void __launch_bounds__(256)
convert_i16_kernel(const int16_t *__restrict__ in, float *__restrict__ out, size_t n, float scale)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t tail = 0;
    if (((reinterpret_cast<uintptr_t>(in) & 15) | (reinterpret_cast<uintptr_t>(out) & 31)) == 0) {
        const size_t n8 = n >> 3;
        for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += stride) {
            const uint4 w = reinterpret_cast<const uint4 *>(in)[g];
            const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
            float v[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[2 * k] = __fmul_rn(scale, (float)(int16_t)(ws[k] & 0xFFFFu));
                v[2 * k + 1] = __fmul_rn(scale, (float)(int16_t)(ws[k] >> 16));
            }
            stg_stream8(out + 8 * g, v);
        }
        tail = n8 << 3;
    }
    for (size_t i = tail + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __fmul_rn(scale, (float)in[i]);
}

int launch_convert_u8(const uint8_t *in, float *out, size_t n, cudaStream_t s)
{
    if (n == 0) return CB_OK;
    convert_u8_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, s>>>(in, out, n);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

int launch_convert_i16(const int16_t *in, float *out, size_t n, float scale, cudaStream_t s)
{
    if (n == 0) return CB_OK;
    convert_i16_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, s>>>(in, out, n, scale);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// ---------------------------------------------------------------- real <-> complex glue of examples/fm_radio.rs
// Convert2Node (fm_radio.rs:98-118): x -> Complex(x, 0);  Convert3Node (:122-142): z -> z.re
__global__ void __launch_bounds__(256) real_to_complex_kernel(const float *__restrict__ in, float2 *__restrict__ out, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = make_float2(in[i], 0.f);
}

__global__ void __launch_bounds__(256) complex_real_kernel(const float2 *__restrict__ in, float *__restrict__ out, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i].x;
}

int launch_real_to_complex(const float *in, float2 *out, size_t n, cudaStream_t s)
{
    if (n == 0) return CB_OK;
    real_to_complex_kernel<<<grid_for(n, 256), 256, 0, s>>>(in, out, n);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

int launch_complex_real(const float2 *in, float *out, size_t n, cudaStream_t s)
{
    if (n == 0) return CB_OK;
    complex_real_kernel<<<grid_for(n, 256), 256, 0, s>>>(in, out, n);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// ---------------------------------------------------------------- synthetic data
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256)
synth_kernel(float *__restrict__ out, size_t nfloats, unsigned long long base)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nfloats; i += stride) {
        const unsigned long long z = splitmix64(base + i);
        out[i] = __fadd_rn(__fmul_rn((float)(z >> 40), 1.0f / 8388608.0f), -1.0f);
    }
}

int launch_synth(float *out, size_t nfloats, unsigned long long base, cudaStream_t s)
{
    if (nfloats == 0) return CB_OK;
    synth_kernel<<<grid_for(nfloats, 256), 256, 0, s>>>(out, nfloats, base);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

}  // namespace cb
