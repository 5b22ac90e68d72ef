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
    const size_t npair = n >> 1;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    if (vec) {
        const float2 step = phase_rotation(dphase);  // e^{j dphase}: second sample of the pair
        for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npair; p += stride) {
            const float4 v = ldg_stream(reinterpret_cast<const float4 *>(x) + p);
            const float2 r0 = phase_rotation(fma((double)(2 * p), dphase, phase0));
            const float2 r1 = cmul(r0, step);
            const float2 a = cmul(make_float2(v.x, v.y), r0);
            const float2 b = cmul(make_float2(v.z, v.w), r1);
            stg_stream(reinterpret_cast<float4 *>(y) + p, make_float4(a.x, a.y, b.x, b.y));
        }
        if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
            const size_t i = n - 1;
            y[i] = cmul(x[i], phase_rotation(fma((double)i, dphase, phase0)));
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            y[i] = cmul(x[i], phase_rotation(fma((double)i, dphase, phase0)));
    }
}

int launch_mixer(const float2 *x, float2 *y, size_t n, double phase0, double dphase, cudaStream_t s)
{
    if (n == 0) return CB_OK;
    mixer_kernel<<<grid_for(n / 2 + 1, 256), 256, 0, s>>>(x, y, n, phase0, dphase);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// ---------------------------------------------------------------- FM demod
// out[n] = atan2(Im, Re)(x[n] * conj(x[n-1]))  (src/modulation/analog.rs:22-34).
// The product is formed with individually rounded operations in the
// reference's operand order so that signed zeros (first sample: prev = 0)
// select the same atan2 branch.
__global__ void __launch_bounds__(256)
fm_kernel(const float2 *__restrict__ x, float *__restrict__ out, size_t n, const float2 *prev_in, float2 *prev_out)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float2 s = x[i];
        const float2 p = i ? x[i - 1] : *prev_in;
        out[i] = fm_angle(s, p);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *prev_out = n ? x[n - 1] : *prev_in;
}

int launch_fm(const float2 *x, float *out, size_t n, const float2 *prev_in, float2 *prev_out, cudaStream_t s)
{
    fm_kernel<<<grid_for(n, 256), 256, 0, s>>>(x, out, n, prev_in, prev_out);
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

__global__ void __launch_bounds__(256)
quantize_i16_kernel(const float *__restrict__ in, int16_t *__restrict__ out, size_t n, float scale)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // Rust `as i16`: truncate toward zero, saturate, NaN -> 0
        int v = __float2int_rz(__fmul_rn(scale, in[i]));
        v = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
        out[i] = (int16_t)v;
    }
}

int launch_quantize_i16(const float *in, int16_t *out, size_t n, float scale, cudaStream_t s)
{
    if (n == 0) return CB_OK;
    quantize_i16_kernel<<<grid_for(n, 256), 256, 0, s>>>(in, out, n, scale);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// ---------------------------------------------------------------- IQ edge formats
// u8 offset binary (RTL-SDR; examples/fm_radio.rs:84-87): (x as f32 - 127.5) / 127.5, both operations exactly rounded
__global__ void __launch_bounds__(256)
convert_u8_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __fdiv_rn(__fsub_rn((float)in[i], 127.5f), 127.5f);
}

// interleaved i16 IQ (src/io/raw_iq.rs:20-140): scale * (x as f32); scale = 1 is the plain cast
__global__ void __launch_bounds__(256)
convert_i16_kernel(const int16_t *__restrict__ in, float *__restrict__ out, size_t n, float scale)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __fmul_rn(scale, (float)in[i]);
}

int launch_convert_u8(const uint8_t *in, float *out, size_t n, cudaStream_t s)
{
    if (n == 0) return CB_OK;
    convert_u8_kernel<<<grid_for(n, 256), 256, 0, s>>>(in, out, n);
    count_launch();
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

int launch_convert_i16(const int16_t *in, float *out, size_t n, float scale, cudaStream_t s)
{
    if (n == 0) return CB_OK;
    convert_i16_kernel<<<grid_for(n, 256), 256, 0, s>>>(in, out, n, scale);
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
