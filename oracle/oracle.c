/*
 * oracle.c -- CPU restatement of the comms-rs FIR / mixer / FFT hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or the
 * timed CPU baseline.  The product (libcomms_b200.so) never links or calls it.
 *
 * Every function cites the reference lines it restates (paths relative to the
 * upstream tree, ostrosco/comms-rs).  The reference is Rust and cannot be
 * built in this image (no cargo/rustc), so this is a restatement, pinned by
 * the reference's own golden vectors (tests/test_oracle_golden.py):
 *   FIR      src/filter/fir_node.rs:259-314      pulse  src/pulse.rs:129-183
 *   resample src/util/resample_node.rs:139-175   mixer  src/mixer.rs:184-223,274-313
 *   FFT N=10 src/fft/fft_node.rs:194-244         taps   src/util/math.rs:359-520
 *   PRN      src/prns.rs:178-220                 maps   src/modulation/digital.rs:52-157
 * PARITY UNPINNED (no reference vector exists): IFFT, every power-of-two FFT
 * size, FM demod.  rustfft 2.1.0 (Cargo.lock:841-850) is not vendored; the FFT
 * here is the published definition X[k] = sum x[n] e^{-+j2 pi kn/N} (no scaling)
 * evaluated in f64 and rounded to the element type, which is what
 * src/fft/mod.rs:73-96 does around rustfft.
 *
 * Build: gcc -O2 -ffp-contract=off (Rust never contracts a*b+c into an FMA;
 * every f32 operation below is individually rounded, like the reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

typedef struct { float re, im; } c32;
typedef struct { double re, im; } c64;
typedef struct { int16_t re, im; } ci16;

/* ------------------------------------------------------------------ */
/* FIR: src/filter/fir.rs:43-54 (fir) and :87-102 (batch_fir).         */
/* state.rotate_right(1); state[0] = x; y = sum_k taps[k]*state[k]     */
/* over zip(taps, state) => k < min(len taps, len state), left fold    */
/* from zero, num-complex Mul: (a.re*b.re - a.im*b.im,                 */
/*                              a.re*b.im + a.im*b.re), a = tap.       */
/* ------------------------------------------------------------------ */
#define DEF_FIR_LITERAL(NAME, CT, RT)                                          \
ORC_API void NAME(const CT *in, size_t n, const CT *taps, size_t ntaps,        \
                  CT *state, size_t nstate, CT *out)                           \
{                                                                              \
    size_t kk = ntaps < nstate ? ntaps : nstate;                               \
    for (size_t i = 0; i < n; ++i) {                                           \
        if (nstate > 0) {                                                      \
            /* rotate_right(1) then overwrite slot 0 (fir.rs:97-98) */         \
            memmove(state + 1, state, (nstate - 1) * sizeof(CT));              \
            state[0] = in[i];                                                  \
        }                                                                      \
        RT sr = 0, si = 0;                                                     \
        for (size_t k = 0; k < kk; ++k) {                                      \
            RT pr = (RT)(taps[k].re * state[k].re) - (RT)(taps[k].im * state[k].im); \
            RT pi = (RT)(taps[k].re * state[k].im) + (RT)(taps[k].im * state[k].re); \
            sr = (RT)(sr + pr);                                                \
            si = (RT)(si + pi);                                                \
        }                                                                      \
        out[i].re = sr;                                                        \
        out[i].im = si;                                                        \
    }                                                                          \
}
DEF_FIR_LITERAL(orc_batch_fir_c32, c32, float)
DEF_FIR_LITERAL(orc_batch_fir_c64, c64, double)
DEF_FIR_LITERAL(orc_batch_fir_ci16, ci16, int16_t)

/*
 * Same arithmetic, same order, without the per-sample memmove: the delay line
 * is kept as a linear buffer [state reversed | input].  Bit-identical to the
 * literal form (tests assert it); used where the literal O(n*K) memmove would
 * make a parity case take minutes.  state is updated exactly as the reference
 * leaves it (newest first).
 */
ORC_API void orc_batch_fir_c32_fast(const c32 *in, size_t n, const c32 *taps,
                                    size_t ntaps, c32 *state, size_t nstate,
                                    c32 *out)
{
    size_t kk = ntaps < nstate ? ntaps : nstate;
    if (nstate == 0) { /* zip over an empty state: every output is 0 */
        for (size_t i = 0; i < n; ++i) { out[i].re = 0; out[i].im = 0; }
        return;
    }
    c32 *line = (c32 *)malloc((nstate + n) * sizeof(c32));
    for (size_t k = 0; k < nstate; ++k) line[nstate - 1 - k] = state[k];
    memcpy(line + nstate, in, n * sizeof(c32));
    for (size_t i = 0; i < n; ++i) {
        const c32 *x = line + nstate + i; /* x[-k] = sample k steps back */
        float sr = 0, si = 0;
        for (size_t k = 0; k < kk; ++k) {
            float pr = taps[k].re * x[-(long)k].re - taps[k].im * x[-(long)k].im;
            float pi = taps[k].re * x[-(long)k].im + taps[k].im * x[-(long)k].re;
            sr = sr + pr;
            si = si + pi;
        }
        out[i].re = sr;
        out[i].im = si;
    }
    for (size_t k = 0; k < nstate; ++k) state[k] = line[nstate + n - 1 - k];
    free(line);
}

/* ------------------------------------------------------------------ */
/* Resampling: src/util/resample_node.rs:53-65 (decimate: indices      */
/* 0,D,2D.. of THIS batch; D in {0,1} copies) and :120-131 (upsample:  */
/* out[i*L] = in[i], zeros elsewhere; L in {0,1} copies).  Element     */
/* size is a parameter because the reference is generic over T.        */
/* ------------------------------------------------------------------ */
ORC_API size_t orc_decimate(const void *in, size_t n, size_t elem, size_t rate,
                            void *out)
{
    if (rate == 0 || rate == 1) { memcpy(out, in, n * elem); return n; }
    size_t m = 0;
    for (size_t ix = 0; ix < n; ix += rate, ++m)
        memcpy((char *)out + m * elem, (const char *)in + ix * elem, elem);
    return m;
}

ORC_API size_t orc_upsample(const void *in, size_t n, size_t elem, size_t rate,
                            void *out)
{
    if (rate == 0 || rate == 1) { memcpy(out, in, n * elem); return n; }
    memset(out, 0, n * rate * elem);
    for (size_t i = 0; i < n; ++i)
        memcpy((char *)out + i * rate * elem, (const char *)in + i * elem, elem);
    return n * rate;
}

/* ------------------------------------------------------------------ */
/* PulseNode::run, src/pulse.rs:82-92: per symbol, fir(symbol) then    */
/* sam_per_sym-1 calls of fir(0).  Batched over n symbols here.        */
/* ------------------------------------------------------------------ */
#define DEF_PULSE(NAME, CT, FIRNAME)                                           \
ORC_API void NAME(const CT *sym, size_t n, const CT *taps, size_t ntaps,       \
                  CT *state, size_t nstate, size_t sam_per_sym, CT *out)       \
{                                                                              \
    CT zero; memset(&zero, 0, sizeof zero);                                    \
    for (size_t i = 0; i < n; ++i) {                                           \
        FIRNAME(sym + i, 1, taps, ntaps, state, nstate, out + i * sam_per_sym);\
        for (size_t p = 1; p < sam_per_sym; ++p)                               \
            FIRNAME(&zero, 1, taps, ntaps, state, nstate,                      \
                    out + i * sam_per_sym + p);                                \
    }                                                                          \
}
DEF_PULSE(orc_pulse_c32, c32, orc_batch_fir_c32)
DEF_PULSE(orc_pulse_ci16, ci16, orc_batch_fir_ci16)

/* ------------------------------------------------------------------ */
/* Mixer: src/mixer.rs:43-51 (new: dphase wrapped into [0,2pi)) and    */
/* :73-84 (mix: f64 multiply by exp(j*phase); phase += dphase; ONE     */
/* conditional wrap when phase > 2pi; result cast back to T).          */
/* Complex::exp(0 + j*phi) = from_polar(exp(0)=1, phi) = (cos, sin).   */
/* ------------------------------------------------------------------ */
ORC_API double orc_mixer_wrap_dphase(double dphase)
{
    while (dphase >= 2.0 * M_PI) dphase -= 2.0 * M_PI;
    while (dphase < 0.0) dphase += 2.0 * M_PI;
    return dphase;
}

#define DEF_MIX(NAME, CT, RT)                                                  \
ORC_API void NAME(const CT *in, size_t n, double *phase, double dphase, CT *out)\
{                                                                              \
    double ph = *phase;                                                        \
    for (size_t i = 0; i < n; ++i) {                                           \
        double xr = (double)in[i].re, xi = (double)in[i].im;                   \
        double c = 1.0 * cos(ph), s = 1.0 * sin(ph);                           \
        double rr = xr * c - xi * s;                                           \
        double ri = xr * s + xi * c;                                           \
        ph += dphase;                                                          \
        if (ph > 2.0 * M_PI) ph -= 2.0 * M_PI;                                 \
        out[i].re = (RT)rr;                                                    \
        out[i].im = (RT)ri;                                                    \
    }                                                                          \
    *phase = ph;                                                               \
}
DEF_MIX(orc_mix_c32, c32, float)
DEF_MIX(orc_mix_c64, c64, double)

/* ------------------------------------------------------------------ */
/* Nco::push: src/demodulation/nco.rs:71-77 (Nco::new :41-49 wraps     */
/* dphase like Mixer::new).  phase += dphase + perr; ONE conditional   */
/* wrap when phase > 2pi; out = exp(j*phase) of the UPDATED phase.     */
/* PARITY UNPINNED: the reference has no test for the NCO.             */
/* ------------------------------------------------------------------ */
ORC_API void orc_nco_push(const double *perr, size_t n, double *phase, double dphase, c64 *out)
{
    double ph = *phase;
    for (size_t i = 0; i < n; ++i) {
        ph += dphase + perr[i];
        if (ph > 2.0 * M_PI) ph -= 2.0 * M_PI;
        out[i].re = 1.0 * cos(ph);
        out[i].im = 1.0 * sin(ph);
    }
    *phase = ph;
}

/* ------------------------------------------------------------------ */
/* FM demod: src/modulation/analog.rs:22-34; prev starts at 0 (:43-47) */
/* theta = samp * conj(prev); out = atan2(theta.im, theta.re).         */
/* ------------------------------------------------------------------ */
ORC_API void orc_fm_demod_c32(const c32 *in, size_t n, c32 *prev, float *out)
{
    c32 p = *prev;
    for (size_t i = 0; i < n; ++i) {
        float br = p.re, bi = -p.im; /* conj(prev) */
        float tr = in[i].re * br - in[i].im * bi;
        float ti = in[i].re * bi + in[i].im * br;
        out[i] = atan2f(ti, tr);
        p = in[i];
    }
    *prev = p;
}

/* ------------------------------------------------------------------ */
/* FFT: src/fft/mod.rs:73-96 -- convert to f64, unnormalised DFT with  */
/* exponent sign - (forward) or + (ifft=true, fft_node.rs:65-67),      */
/* convert back to T.  rustfft itself is absent; this is the defining  */
/* sum (orc_dft_*: O(N^2), long-double accumulation, any N) and an     */
/* f64 radix-2 FFT with exact per-index twiddles for power-of-two N.   */
/* ------------------------------------------------------------------ */
ORC_API void orc_dft_f64(const c64 *in, size_t n, int inverse, c64 *out)
{
    const long double tw = (inverse ? 2.0L : -2.0L) * 3.14159265358979323846264338327950288L / (long double)n;
    for (size_t k = 0; k < n; ++k) {
        long double sr = 0, si = 0;
        for (size_t j = 0; j < n; ++j) {
            size_t m = (k * j) % n;
            long double c = cosl(tw * (long double)m), s = sinl(tw * (long double)m);
            sr += (long double)in[j].re * c - (long double)in[j].im * s;
            si += (long double)in[j].re * s + (long double)in[j].im * c;
        }
        out[k].re = (double)sr;
        out[k].im = (double)si;
    }
}

static void fft_pow2_f64(c64 *a, size_t n, int inverse, const c64 *tw)
{
    /* bit reversal */
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { c64 t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        size_t half = len >> 1, step = n / len;
        for (size_t i = 0; i < n; i += len) {
            for (size_t j = 0; j < half; ++j) {
                c64 w = tw[j * step];
                double wi = inverse ? -w.im : w.im;
                c64 u = a[i + j], v = a[i + j + half];
                double tr = v.re * w.re - v.im * wi;
                double ti = v.re * wi + v.im * w.re;
                a[i + j].re = u.re + tr; a[i + j].im = u.im + ti;
                a[i + j + half].re = u.re - tr; a[i + j + half].im = u.im - ti;
            }
        }
    }
}

/* frames: nframes contiguous frames of n samples; f32 in/out, f64 inside. */
ORC_API int orc_fft_c32(const c32 *in, size_t n, size_t nframes, int inverse,
                        c32 *out)
{
    if (n == 0) return -1;
    int pow2 = (n & (n - 1)) == 0;
    c64 *buf = (c64 *)malloc(n * sizeof(c64));
    c64 *res = (c64 *)malloc(n * sizeof(c64));
    c64 *tw = NULL;
    if (pow2) {
        tw = (c64 *)malloc((n / 2 + 1) * sizeof(c64));
        for (size_t j = 0; j < n / 2 + 1; ++j) { /* forward twiddles e^{-j2pi j/n} */
            long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)n;
            tw[j].re = (double)cosl(a);
            tw[j].im = (double)sinl(a);
        }
    }
    for (size_t f = 0; f < nframes; ++f) {
        for (size_t i = 0; i < n; ++i) { /* mod.rs:78-83 */
            buf[i].re = (double)in[f * n + i].re;
            buf[i].im = (double)in[f * n + i].im;
        }
        if (pow2) { fft_pow2_f64(buf, n, inverse, tw); memcpy(res, buf, n * sizeof(c64)); }
        else orc_dft_f64(buf, n, inverse, res);
        for (size_t i = 0; i < n; ++i) { /* mod.rs:89-94 */
            out[f * n + i].re = (float)res[i].re;
            out[f * n + i].im = (float)res[i].im;
        }
    }
    free(buf); free(res); free(tw);
    return 0;
}

ORC_API int orc_fft_c64(const c64 *in, size_t n, size_t nframes, int inverse,
                        c64 *out)
{
    if (n == 0) return -1;
    int pow2 = (n & (n - 1)) == 0;
    c64 *tw = NULL;
    if (pow2) {
        tw = (c64 *)malloc((n / 2 + 1) * sizeof(c64));
        for (size_t j = 0; j < n / 2 + 1; ++j) {
            long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)n;
            tw[j].re = (double)cosl(a);
            tw[j].im = (double)sinl(a);
        }
    }
    for (size_t f = 0; f < nframes; ++f) {
        if (pow2) {
            memcpy(out + f * n, in + f * n, n * sizeof(c64));
            fft_pow2_f64(out + f * n, n, inverse, tw);
        } else orc_dft_f64(in + f * n, n, inverse, out + f * n);
    }
    free(tw);
    return 0;
}

/* ------------------------------------------------------------------ */
/* Tap generators: src/util/math.rs.  All math in f64, cast at the end.*/
/* Return 0 ok, 1 = MathError::InvalidRolloffError.                    */
/* ------------------------------------------------------------------ */
ORC_API double orc_sinc(double x) /* math.rs:120-126 */
{
    return x != 0.0 ? sin(M_PI * x) / (M_PI * x) : 1.0;
}

ORC_API int orc_rrc_taps_f64(uint32_t n_taps, double sam_per_sym, double beta,
                             double *out) /* math.rs:221-280 */
{
    if (beta < 0.0 || beta > 1.0) return 1;
    const double tsym = 1.0, fs = sam_per_sym / tsym;
    const double fzero = (1.0 / tsym) * (1.0 + beta * (4.0 / M_PI - 1.0));
    const double zero_denom = beta != 0.0 ? tsym / (4.0 * beta) : 0.0;
    for (uint32_t i = 0; i < n_taps; ++i) {
        double t = ((double)i - (double)(n_taps - 1) / 2.0) / fs;
        double v;
        if (fabs(t) < 2.220446049250313e-16) {
            v = fzero;
        } else if (fabs(t - zero_denom) < 2.220446049250313e-16 ||
                   fabs(t + zero_denom) < 2.220446049250313e-16) {
            v = (beta / (tsym * sqrt(2.0))) *
                ((1.0 + 2.0 / M_PI) * sin(M_PI / (4.0 * beta)) +
                 (1.0 - (2.0 / M_PI)) * cos(M_PI / (4.0 * beta)));
        } else {
            double q = 4.0 * beta * (t / tsym);
            v = (1.0 / tsym) *
                (sin(M_PI * (t / tsym) * (1.0 - beta)) +
                 4.0 * beta * (t / tsym) * cos(M_PI * (t / tsym) * (1.0 + beta))) /
                (M_PI * (t / tsym) * (1.0 - q * q));
        }
        out[i] = v;
    }
    return 0;
}

ORC_API int orc_rc_taps_f64(uint32_t n_taps, double sam_per_sym, double beta,
                            double *out) /* math.rs:151-196 */
{
    if (beta < 0.0 || beta > 1.0) return 1;
    const double tsym = 1.0, fs = sam_per_sym / tsym;
    const double zero_denom = beta != 0.0 ? tsym / (2.0 * beta) : 0.0;
    for (uint32_t i = 0; i < n_taps; ++i) {
        double t = ((double)i - (double)(n_taps - 1) / 2.0) / fs;
        double v;
        if (fabs(t - zero_denom) < 2.220446049250313e-16 ||
            fabs(t + zero_denom) < 2.220446049250313e-16) {
            v = (M_PI / (4.0 * tsym)) * orc_sinc(1.0 / (2.0 * beta));
        } else {
            double q = (2.0 * beta * t) / tsym;
            v = (1.0 / tsym) * orc_sinc(t / tsym) * cos((M_PI * beta * t) / tsym) /
                (1.0 - q * q);
        }
        out[i] = v;
    }
    return 0;
}

ORC_API void orc_gaussian_taps_f64(uint32_t n_taps, double sam_per_sym,
                                   double alpha, double *out) /* math.rs:79-102 */
{
    const double fs = sam_per_sym / 1.0;
    for (uint32_t i = 0; i < n_taps; ++i) {
        double t = ((double)i - (double)(n_taps - 1) / 2.0) / fs;
        out[i] = sqrt(alpha / M_PI) * exp(-alpha * (t * t));
    }
}

/* math.rs:307-342; returns the (odd) number of taps written. */
ORC_API int orc_qfilt_taps_f64(uint32_t n_taps, double alpha, uint32_t sam_per_sym,
                               double *out, uint32_t *n_out)
{
    if (alpha < 0.0 || alpha > 1.0) return 1;
    uint32_t real_n = n_taps;
    if (n_taps % 2 == 0) real_n += 1;
    int32_t d = (int32_t)floor((double)real_n / 2.0);
    for (uint32_t x = 0; x < real_n; ++x) {
        double tt = (double)((int32_t)x - d) / (double)sam_per_sym;
        double two_alpha_tt = 2.0 * alpha * tt;
        if (fabs(two_alpha_tt) == 1.0) {
            out[x] = sin(M_PI * alpha * tt) / (8.0 * tt);
        } else {
            double numerator = alpha * cos(M_PI * alpha * tt);
            double denominator = M_PI * (1.0 - (two_alpha_tt * two_alpha_tt));
            out[x] = numerator / denominator;
        }
    }
    *n_out = real_n;
    return 0;
}

/* ------------------------------------------------------------------ */
/* PRN generator: src/prns.rs:64-71.  width = bits of the register     */
/* type (8/16/32/64).  fb = parity(state & mask); out = MSB;           */
/* state = (state << 1) | fb, truncated to the width.                  */
/* ------------------------------------------------------------------ */
ORC_API void orc_prn_bits(uint64_t poly_mask, uint64_t *state, unsigned width,
                          size_t n, uint8_t *out)
{
    uint64_t wm = width >= 64 ? ~0ULL : ((1ULL << width) - 1ULL);
    uint64_t st = *state & wm;
    for (size_t i = 0; i < n; ++i) {
        uint64_t fb = (uint64_t)(__builtin_popcountll(st & poly_mask & wm) & 1);
        out[i] = (uint8_t)(st >> (width - 1));
        st = ((st << 1) & wm) | fb;
    }
    *state = st;
}

/* ------------------------------------------------------------------ */
/* Symbol maps.  Library tables src/modulation/digital.rs:6-44         */
/* (0 -> +1, 1 -> -1; bytes LSB first) and the example maps            */
/* examples/single_thread_bpsk.rs:29-32 (b -> 2b-1 + 0j) and           */
/* examples/single_thread_qpsk.rs:29-36 (even bit -> re, odd -> im).   */
/* ------------------------------------------------------------------ */
ORC_API int orc_bpsk_bit_mod(uint8_t bit, ci16 *out)
{
    if (bit == 0) { out->re = 1; out->im = 0; return 0; }
    if (bit == 1) { out->re = -1; out->im = 0; return 0; }
    return 1; /* None */
}

ORC_API void orc_bpsk_byte_mod(uint8_t byte, ci16 *out8)
{
    for (unsigned i = 0; i < 8; ++i)
        orc_bpsk_bit_mod((uint8_t)(((1u << i) & byte) >> i), out8 + i);
}

ORC_API int orc_qpsk_bit_mod(uint8_t bits, ci16 *out)
{
    switch (bits) {
    case 0: out->re = 1; out->im = 1; return 0;
    case 1: out->re = -1; out->im = 1; return 0;
    case 2: out->re = 1; out->im = -1; return 0;
    case 3: out->re = -1; out->im = -1; return 0;
    default: return 1;
    }
}

ORC_API void orc_qpsk_byte_mod(uint8_t byte, ci16 *out4)
{
    for (unsigned i = 0; i < 8; i += 2)
        orc_qpsk_bit_mod((uint8_t)(((3u << i) & byte) >> i), out4 + i / 2);
}

ORC_API void orc_example_bpsk_map(const uint8_t *bits, size_t n, c32 *out)
{
    for (size_t i = 0; i < n; ++i) {
        out[i].re = (float)bits[i] * 2.0f - 1.0f;
        out[i].im = 0.0f;
    }
}

ORC_API void orc_example_qpsk_map(const uint8_t *bits, size_t nbits, c32 *out)
{
    for (size_t i = 0; i + 1 < nbits; i += 2) {
        out[i / 2].re = (float)bits[i] * 2.0f - 1.0f;
        out[i / 2].im = (float)bits[i + 1] * 2.0f - 1.0f;
    }
}

/* ------------------------------------------------------------------ */
/* Example edge conversions.  (8192.0*x) as i16: Rust `as` truncates   */
/* toward zero and saturates, NaN -> 0 (single_thread_bpsk.rs:40-48).  */
/* u8 IQ -> f32: (u8 - 127.5)/127.5 (examples/fm_radio.rs:84-87).      */
/* ------------------------------------------------------------------ */
ORC_API void orc_quantize_i16(const float *in, size_t n, float scale, int16_t *out)
{
    for (size_t i = 0; i < n; ++i) {
        float v = scale * in[i];
        int16_t q;
        if (v != v) q = 0;
        else if (v >= 32767.0f) q = 32767;
        else if (v <= -32768.0f) q = -32768;
        else q = (int16_t)v; /* C truncates toward zero in range */
        out[i] = q;
    }
}

ORC_API void orc_u8_to_f32(const uint8_t *in, size_t n, float *out)
{
    for (size_t i = 0; i < n; ++i) out[i] = ((float)in[i] - 127.5f) / 127.5f;
}

/* ------------------------------------------------------------------ */
/* Synthetic-input generator shared by oracle, library and bench       */
/* (SURVEY 8(d)): splitmix64(seed + index) -> top 24 bits -> [-1,1).   */
/* Float index i of a c32 stream: re = 2n, im = 2n+1.  Not part of the */
/* reference; specified here so CPU and GPU sides agree bit for bit.   */
/* ------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

ORC_API void orc_synth_uniform_f32(uint64_t seed, uint64_t first, size_t n, float *out)
{
    for (size_t i = 0; i < n; ++i) {
        uint64_t z = splitmix64(seed + first + i);
        out[i] = (float)(z >> 40) * (1.0f / 8388608.0f) - 1.0f;
    }
}

/* ------------------------------------------------------------------ */
/* Whole-chain restatements used as parity cases / CPU baselines.      */
/* ------------------------------------------------------------------ */

/* examples/single_thread_bpsk.rs:16-48 with PrnGen bits (BASELINE cfg 1):
 * per batch: bits -> 2b-1 -> zero-stuff x sps -> batch_fir (state carried). */
ORC_API void orc_bpsk_chain(uint64_t mask, uint64_t *prn_state, unsigned width,
                            size_t nsym, size_t batch, size_t sps,
                            const c32 *taps, size_t ntaps, c32 *state, size_t nstate,
                            uint8_t *bits_out, c32 *shaped_out, int16_t *iq_out)
{
    uint8_t *bits = (uint8_t *)malloc(batch);
    c32 *sym = (c32 *)malloc(batch * sizeof(c32));
    c32 *ups = (c32 *)malloc(batch * sps * sizeof(c32));
    for (size_t done = 0; done < nsym; done += batch) {
        size_t nb = nsym - done < batch ? nsym - done : batch;
        orc_prn_bits(mask, prn_state, width, nb, bits);
        if (bits_out) memcpy(bits_out + done, bits, nb);
        orc_example_bpsk_map(bits, nb, sym);
        orc_upsample(sym, nb, sizeof(c32), sps, ups);
        orc_batch_fir_c32_fast(ups, nb * sps, taps, ntaps, state, nstate,
                               shaped_out + done * sps);
    }
    if (iq_out) orc_quantize_i16((const float *)shaped_out, nsym * sps * 2, 8192.0f, iq_out);
    free(bits); free(sym); free(ups);
}

/* BASELINE cfg 4 restated per channel and per batch with the reference node
 * order: Mixer::mix per sample -> batch_fir -> decimate (phase reset per
 * batch) -> FM::demod (prev carried).  out receives ceil(n/D) floats. */
ORC_API size_t orc_fm_chain_batch(const c32 *in, size_t n, double *phase,
                                  double dphase, const c32 *taps, size_t ntaps,
                                  c32 *state, size_t nstate, size_t decim,
                                  c32 *fm_prev, int do_mix, int do_fm,
                                  c32 *out_c, float *out_f)
{
    c32 *mixed = (c32 *)malloc((n ? n : 1) * sizeof(c32));
    c32 *filt = (c32 *)malloc((n ? n : 1) * sizeof(c32));
    if (do_mix) orc_mix_c32(in, n, phase, dphase, mixed);
    else memcpy(mixed, in, n * sizeof(c32));
    orc_batch_fir_c32_fast(mixed, n, taps, ntaps, state, nstate, filt);
    size_t m = orc_decimate(filt, n, sizeof(c32), decim, mixed);
    if (out_c) memcpy(out_c, mixed, m * sizeof(c32));
    if (do_fm) orc_fm_demod_c32(mixed, m, fm_prev, out_f);
    free(mixed); free(filt);
    return m;
}
