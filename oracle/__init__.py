"""ctypes front-end of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never from the product
package.  Each wrapper keeps the name of the reference item it restates; the
file:line citations live beside the C functions.

Restated here in numpy rather than in oracle.c (f64 glue around the C batch_fir):
frequency_offset_estimate and TimingEstimator (src/demodulation), pinned by the
reference's own recovery tests (tests/test_oracle_golden.py::test_*_reference_kat;
the reference draws its symbols from rand's SmallRng, not reproducible here, so the
pin is the statistic, not a vector).  fft(): lengths that are neither a power of two
nor <= 4200 are evaluated through numpy's f64 transform (same definition, pinned to
the C direct sum at n = 1201).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict = {}


def build(force: bool = False) -> None:
    """Compile liboracle.so / liboracle_native.so with gcc (see Makefile)."""
    if force or not all(
        os.path.exists(os.path.join(_HERE, f))
        and os.path.getmtime(os.path.join(_HERE, f)) >= os.path.getmtime(os.path.join(_HERE, "oracle.c"))
        for f in ("liboracle.so", "liboracle_native.so")
    ):
        subprocess.run(["make", "-C", _HERE, "-s", "all"], check=True)


def lib(native: bool = False) -> C.CDLL:
    name = "liboracle_native.so" if native else "liboracle.so"
    if name not in _LIBS:
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_sinc.restype = C.c_double
        L.orc_sinc.argtypes = [C.c_double]
        L.orc_mixer_wrap_dphase.restype = C.c_double
        L.orc_mixer_wrap_dphase.argtypes = [C.c_double]
        for f in ("orc_decimate", "orc_upsample", "orc_fm_chain_batch"):
            getattr(L, f).restype = C.c_size_t
        _LIBS[name] = L
    return _LIBS[name]


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _c32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.complex64))


_SZ = C.c_size_t

# ---------------------------------------------------------------- FIR


def batch_fir(inp, taps, state, *, literal: bool = False, native: bool = False):
    """batch_fir (src/filter/fir.rs:87-102).  Returns (out, new_state).

    dtype follows `inp`: complex64, complex128, or an (n,2) int16 array.
    literal=True runs the per-sample rotate form; the default linear-buffer
    form is bit-identical (asserted in tests) and fast.
    """
    L = lib(native)
    inp = np.asarray(inp)
    if inp.dtype == np.int16:
        x = np.ascontiguousarray(inp).reshape(-1, 2)
        t = np.ascontiguousarray(np.asarray(taps, dtype=np.int16)).reshape(-1, 2)
        s = np.array(np.asarray(state, dtype=np.int16), copy=True).reshape(-1, 2)
        out = np.zeros_like(x)
        L.orc_batch_fir_ci16(_p(x), _SZ(len(x)), _p(t), _SZ(len(t)), _p(s), _SZ(len(s)), _p(out))
        return out, s
    if inp.dtype == np.complex128:
        x = np.ascontiguousarray(inp)
        t = np.ascontiguousarray(np.asarray(taps, dtype=np.complex128))
        s = np.array(np.asarray(state, dtype=np.complex128), copy=True)
        out = np.zeros_like(x)
        L.orc_batch_fir_c64(_p(x), _SZ(len(x)), _p(t), _SZ(len(t)), _p(s), _SZ(len(s)), _p(out))
        return out, s
    x = _c32(inp)
    t = _c32(taps)
    s = np.array(_c32(state), copy=True)
    out = np.zeros_like(x)
    fn = L.orc_batch_fir_c32 if literal else L.orc_batch_fir_c32_fast
    fn(_p(x), _SZ(len(x)), _p(t), _SZ(len(t)), _p(s), _SZ(len(s)), _p(out))
    return out, s


def fir(sample, taps, state):
    """fir (src/filter/fir.rs:43-54): one sample in, one out."""
    out, st = batch_fir(np.asarray([sample], dtype=np.complex64), taps, state, literal=True)
    return out[0], st


# ---------------------------------------------------------------- resample


def decimate(data, rate: int):
    """DecimateNode::decimate (src/util/resample_node.rs:53-65)."""
    a = np.ascontiguousarray(data)
    out = np.empty_like(a)
    m = lib().orc_decimate(_p(a), _SZ(len(a)), _SZ(a.dtype.itemsize * int(np.prod(a.shape[1:], dtype=np.int64))), _SZ(rate), _p(out))
    return out[:m].copy()


def upsample(data, rate: int):
    """UpsampleNode::upsample (src/util/resample_node.rs:120-131)."""
    a = np.ascontiguousarray(data)
    r = rate if rate > 1 else 1
    out = np.empty((len(a) * r,) + a.shape[1:], dtype=a.dtype)
    m = lib().orc_upsample(_p(a), _SZ(len(a)), _SZ(a.dtype.itemsize * int(np.prod(a.shape[1:], dtype=np.int64))), _SZ(rate), _p(out))
    return out[:m]


def pulse(symbols, taps, state, sam_per_sym: int):
    """PulseNode::run over a batch of symbols (src/pulse.rs:82-92)."""
    L = lib()
    symbols = np.asarray(symbols)
    if symbols.dtype == np.int16:
        x = np.ascontiguousarray(symbols).reshape(-1, 2)
        t = np.ascontiguousarray(np.asarray(taps, dtype=np.int16)).reshape(-1, 2)
        s = np.array(np.asarray(state, dtype=np.int16), copy=True).reshape(-1, 2)
        out = np.zeros((len(x) * sam_per_sym, 2), dtype=np.int16)
        L.orc_pulse_ci16(_p(x), _SZ(len(x)), _p(t), _SZ(len(t)), _p(s), _SZ(len(s)), _SZ(sam_per_sym), _p(out))
        return out, s
    x, t = _c32(symbols), _c32(taps)
    s = np.array(_c32(state), copy=True)
    out = np.zeros(len(x) * sam_per_sym, dtype=np.complex64)
    L.orc_pulse_c32(_p(x), _SZ(len(x)), _p(t), _SZ(len(t)), _p(s), _SZ(len(s)), _SZ(sam_per_sym), _p(out))
    return out, s


# ---------------------------------------------------------------- mixer / FM


class Mixer:
    """Mixer (src/mixer.rs:16-84).  Note MixerNode::new(dphase, phase) order."""

    def __init__(self, phase: float, dphase: float):
        self.phase = float(phase)
        self.dphase = float(lib().orc_mixer_wrap_dphase(C.c_double(dphase)))

    def mix(self, inp, native: bool = False):
        a = np.asarray(inp)
        ph = C.c_double(self.phase)
        if a.dtype == np.complex128:
            x = np.ascontiguousarray(a)
            out = np.empty_like(x)
            lib(native).orc_mix_c64(_p(x), _SZ(len(x)), C.byref(ph), C.c_double(self.dphase), _p(out))
        else:
            x = _c32(a)
            out = np.empty_like(x)
            lib(native).orc_mix_c32(_p(x), _SZ(len(x)), C.byref(ph), C.c_double(self.dphase), _p(out))
        self.phase = ph.value
        return out


class Nco:
    """Nco (src/demodulation/nco.rs:15-78); NcoNode::new(dphase, phase) order is the node's (:112-127)."""

    def __init__(self, phase: float, dphase: float):
        self.phase = float(phase)
        self.dphase = float(lib().orc_mixer_wrap_dphase(C.c_double(dphase)))  # same wrap loop as Mixer::new (nco.rs:41-49)

    def push(self, perr):
        """One sample (float) or a batch (array): the per-sample recurrence applied in order."""
        scalar = np.isscalar(perr)
        e = np.ascontiguousarray(np.atleast_1d(np.asarray(perr, dtype=np.float64)))
        out = np.empty(len(e), dtype=np.complex128)
        ph = C.c_double(self.phase)
        lib().orc_nco_push(_p(e), _SZ(len(e)), C.byref(ph), C.c_double(self.dphase), _p(out))
        self.phase = ph.value
        return out[0] if scalar else out


class FM:
    """FM (src/modulation/analog.rs:7-47); prev starts at 0."""

    def __init__(self):
        self.prev = np.zeros(1, dtype=np.complex64)

    def demod(self, samples):
        x = _c32(samples)
        out = np.empty(len(x), dtype=np.float32)
        lib().orc_fm_demod_c32(_p(x), _SZ(len(x)), _p(self.prev), _p(out))
        return out


# ---------------------------------------------------------------- FFT


_DIRECT_MAX = 4200  # longest non power-of-two length evaluated by the C direct sum


def fft(frames, n: int, inverse: bool = False):
    """BatchFFT::run_fft per frame (src/fft/mod.rs:73-96): f64 inside, T outside."""
    a = np.asarray(frames)
    if a.dtype == np.complex128:
        x = np.ascontiguousarray(a).reshape(-1)
        fn = lib().orc_fft_c64
    else:
        x = _c32(a).reshape(-1)
        fn = lib().orc_fft_c32
    if n == 0 or len(x) % n:
        raise ValueError("input length must be a multiple of fft_size")
    if n > _DIRECT_MAX and n & (n - 1):
        # the C restatement evaluates other lengths as the O(n^2) sum in f64: minutes per frame out here.  Same
        # definition through numpy's f64 transform instead (pinned to the direct sum in tests/test_oracle_golden.py)
        z = x.astype(np.complex128).reshape(-1, n)
        z = np.fft.ifft(z, axis=1) * n if inverse else np.fft.fft(z, axis=1)
        return z.reshape(-1).astype(x.dtype)
    out = np.empty_like(x)
    rc = fn(_p(x), _SZ(n), _SZ(len(x) // n), C.c_int(int(inverse)), _p(out))
    if rc:
        raise ValueError("fft failed")
    return out


# ---------------------------------------------------------------- taps


def _taps(fn, n, *args):
    out = np.empty(n, dtype=np.float64)
    rc = fn(C.c_uint32(n), *args, _p(out))
    if rc:
        raise ValueError("InvalidRolloffError")
    return out


def rrc_taps(n_taps: int, sam_per_sym: float, beta: float, dtype=np.complex64):
    """rrc_taps (src/util/math.rs:221-280): f64 math, cast to T, im = 0."""
    return _taps(lib().orc_rrc_taps_f64, n_taps, C.c_double(sam_per_sym), C.c_double(beta)).astype(dtype)


def rc_taps(n_taps: int, sam_per_sym: float, beta: float, dtype=np.complex128):
    return _taps(lib().orc_rc_taps_f64, n_taps, C.c_double(sam_per_sym), C.c_double(beta)).astype(dtype)


def gaussian_taps(n_taps: int, sam_per_sym: float, alpha: float, dtype=np.complex128):
    out = np.empty(n_taps, dtype=np.float64)
    lib().orc_gaussian_taps_f64(C.c_uint32(n_taps), C.c_double(sam_per_sym), C.c_double(alpha), _p(out))
    return out.astype(dtype)


def rect_taps(n_taps: int, dtype=np.complex64):
    """rect_taps (src/util/math.rs:48-55)."""
    return np.ones(n_taps, dtype=dtype)


def qfilt_taps(n_taps: int, alpha: float, sam_per_sym: int):
    out = np.empty(n_taps + 1, dtype=np.float64)
    m = C.c_uint32(0)
    rc = lib().orc_qfilt_taps_f64(C.c_uint32(n_taps), C.c_double(alpha), C.c_uint32(sam_per_sym), _p(out), C.byref(m))
    if rc:
        raise ValueError("InvalidRolloffError")
    return out[: m.value].copy()


def frequency_offset_estimate(samples) -> float:
    """frequency_offset_estimate (src/demodulation/frequency_estimator.rs:27-42): f64, the sum taken in
    index order as Iterator::sum does."""
    x = np.asarray(samples, dtype=np.complex128)
    if len(x) < 2:
        return 0.0
    mult = x[1:] * np.conj(x[:-1])
    acc = np.cumsum(mult)[-1]  # sequential left fold
    return float(np.arctan2(acc.imag, acc.real))


class TimingEstimator:
    """TimingEstimator::new / push (src/demodulation/timing_estimator.rs:43-58, 85-112)."""

    def __init__(self, n: int, d: int, alpha: float):
        self.n, self.d = int(n), int(d)
        self.qfilt = qfilt_taps(2 * n * d + 1, alpha, n).astype(np.complex128)  # :47-48
        self.delay = np.zeros(n * d + 1, dtype=np.complex128)                   # :49-50
        self.delay[-1] = 1.0

    def push(self, samples) -> float:
        s = np.asarray(samples, dtype=np.complex128)
        i = np.arange(len(s), dtype=np.float64)
        ang = -np.pi * i / float(self.n)                      # :90  (-PI * i) / n
        r = np.cos(ang) + 1j * np.sin(ang)                    # Complex::new(0, ang).exp()
        qin = np.conj(s) * r                                  # :93
        din = s * r                                           # :94
        qout, _ = batch_fir(qin, self.qfilt, np.zeros(2 * self.n * self.d + 1, np.complex128))  # :98-102
        dout, _ = batch_fir(din, self.delay, np.zeros(self.n * self.d + 1, np.complex128))      # :100-103
        if len(s) == 0:
            acc = 0j
        else:
            acc = np.cumsum(qout * dout)[-1]                  # :108-109, sequential sum
        return float(-float(self.n) * np.arctan2(acc.imag, acc.real) / (2.0 * np.pi))  # :111


def sinc(x: float) -> float:
    return float(lib().orc_sinc(C.c_double(x)))


def cast_complex(z: complex, dtype):
    """cast_complex (src/util/math.rs:20-28): NumCast each part; None on overflow."""
    info = np.iinfo(dtype) if np.issubdtype(dtype, np.integer) else None
    parts = []
    for v in (z.real, z.imag):
        if info is not None:
            t = int(v)  # NumCast float->int truncates toward zero
            if t < info.min or t > info.max:
                return None
            parts.append(dtype(t))
        else:
            parts.append(dtype(v))
    return parts[0], parts[1]


# ---------------------------------------------------------------- PRN / maps


class PrnGen:
    """PrnGen<T> (src/prns.rs:38-71); width = bit width of T."""

    def __init__(self, poly_mask: int, state: int, width: int = 8):
        self.mask, self.width = int(poly_mask), int(width)
        self.state = C.c_uint64(int(state))

    def bits(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint8)
        lib().orc_prn_bits(C.c_uint64(self.mask), C.byref(self.state), C.c_uint(self.width), _SZ(n), _p(out))
        return out

    def next_byte(self) -> int:
        return int(self.bits(1)[0])


def bpsk_bit_mod(bit: int):
    o = np.zeros(2, dtype=np.int16)
    return None if lib().orc_bpsk_bit_mod(C.c_uint8(bit), _p(o)) else (int(o[0]), int(o[1]))


def bpsk_byte_mod(byte: int):
    o = np.zeros((8, 2), dtype=np.int16)
    lib().orc_bpsk_byte_mod(C.c_uint8(byte), _p(o))
    return [tuple(int(v) for v in r) for r in o]


def qpsk_bit_mod(bits: int):
    o = np.zeros(2, dtype=np.int16)
    return None if lib().orc_qpsk_bit_mod(C.c_uint8(bits), _p(o)) else (int(o[0]), int(o[1]))


def qpsk_byte_mod(byte: int):
    o = np.zeros((4, 2), dtype=np.int16)
    lib().orc_qpsk_byte_mod(C.c_uint8(byte), _p(o))
    return [tuple(int(v) for v in r) for r in o]


def example_bpsk_map(bits) -> np.ndarray:
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    out = np.empty(len(b), dtype=np.complex64)
    lib().orc_example_bpsk_map(_p(b), _SZ(len(b)), _p(out))
    return out


def example_qpsk_map(bits) -> np.ndarray:
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    out = np.empty(len(b) // 2, dtype=np.complex64)
    lib().orc_example_qpsk_map(_p(b), _SZ(len(b)), _p(out))
    return out


def quantize_i16(x, scale: float = 8192.0) -> np.ndarray:
    f = np.ascontiguousarray(np.asarray(x, dtype=np.complex64)).view(np.float32)
    out = np.empty(len(f), dtype=np.int16)
    lib().orc_quantize_i16(_p(f), _SZ(len(f)), C.c_float(scale), _p(out))
    return out


def u8_to_f32(x) -> np.ndarray:
    b = np.ascontiguousarray(x, dtype=np.uint8)
    out = np.empty(len(b), dtype=np.float32)
    lib().orc_u8_to_f32(_p(b), _SZ(len(b)), _p(out))
    return out


# ---------------------------------------------------------------- synthetic data


def synth_uniform_c32(seed: int, first_sample: int, n: int) -> np.ndarray:
    """n complex samples of the shared splitmix64 generator (floats 2*first..)."""
    out = np.empty(2 * n, dtype=np.float32)
    lib().orc_synth_uniform_f32(C.c_uint64(seed), C.c_uint64(2 * first_sample), _SZ(2 * n), _p(out))
    return out.view(np.complex64)


# ---------------------------------------------------------------- chains


def bpsk_chain(nsym: int, batch: int = 4096, sps: int = 4, mask: int = 0xB8, state0: int = 0x01,
               taps=None, native: bool = False):
    """examples/single_thread_bpsk.rs:16-48 with PrnGen bits (BASELINE cfg 1)."""
    t = _c32(rrc_taps(32, 4.0, 0.25) if taps is None else taps)
    st = np.zeros(len(t), dtype=np.complex64)
    ps = C.c_uint64(state0)
    bits = np.empty(nsym, dtype=np.uint8)
    shaped = np.empty(nsym * sps, dtype=np.complex64)
    iq = np.empty(nsym * sps * 2, dtype=np.int16)
    lib(native).orc_bpsk_chain(C.c_uint64(mask), C.byref(ps), C.c_uint(8), _SZ(nsym), _SZ(batch), _SZ(sps),
                               _p(t), _SZ(len(t)), _p(st), _SZ(len(st)), _p(bits), _p(shaped), _p(iq))
    return bits, shaped, iq, st


class FmChain:
    """mixer -> batch_fir -> decimate -> FM demod, one channel (BASELINE cfg 4)."""

    def __init__(self, dphase, phase, taps, decim, state=None, do_mix=True, do_fm=True, native=False):
        self.m = Mixer(phase, dphase)
        self.taps = _c32(taps)
        self.state = np.zeros(len(self.taps), dtype=np.complex64) if state is None else np.array(_c32(state), copy=True)
        self.decim, self.do_mix, self.do_fm, self.native = int(decim), do_mix, do_fm, native
        self.prev = np.zeros(1, dtype=np.complex64)

    def run(self, inp):
        x = _c32(inp)
        oc = np.empty(max(len(x), 1), dtype=np.complex64)
        of = np.empty(max(len(x), 1), dtype=np.float32)
        ph = C.c_double(self.m.phase)
        m = lib(self.native).orc_fm_chain_batch(
            _p(x), _SZ(len(x)), C.byref(ph), C.c_double(self.m.dphase), _p(self.taps), _SZ(len(self.taps)),
            _p(self.state), _SZ(len(self.state)), _SZ(self.decim), _p(self.prev),
            C.c_int(int(self.do_mix)), C.c_int(int(self.do_fm)), _p(oc), _p(of))
        self.m.phase = ph.value
        return (of[:m].copy() if self.do_fm else oc[:m].copy())
