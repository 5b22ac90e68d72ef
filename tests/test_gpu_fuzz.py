"""Randomised sizes around the tile boundaries of the tensor-core / TMA kernels, against the CPU oracle.
Needs a B200: `pytest -m gpu`.  Seeds are fixed: the cases are reproducible."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIR_TOL = 1e-5


@pytest.fixture(scope="module")
def cb():
    import comms_rs_b200 as m

    m.init(0)
    return m


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    assert a.shape == b.shape, (a.shape, b.shape)
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / d if d > 0 else np.linalg.norm(a - b)


def rnd_c32(rng, n):
    return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)


def sizes_near(rng, unit, count, lo=1):
    """Batch sizes at, just below and just above multiples of `unit`, plus a few arbitrary ones."""
    out = []
    for _ in range(count):
        k = int(rng.integers(0, 4))
        d = int(rng.choice([-3, -2, -1, 0, 1, 2, 3, 17, -64, 65]))
        out.append(max(lo, k * unit + d))
    out += [int(v) for v in rng.integers(lo, 3 * unit, 3)]
    return out


@pytest.mark.parametrize("ntaps,seed", [(64, 1), (63, 2), (33, 3), (100, 4), (128, 5), (7, 6)])
def test_fir_tensor_core_batches_around_tiles(cb, oracle, ntaps, seed, monkeypatch):
    # K1-TC (TMA raw-tile ring, 4096-sample tiles): odd / tiny / tile-straddling batches with carried state
    monkeypatch.setenv("COMMS_B200_FIR_PATH", "tc")
    rng = np.random.default_rng(seed)
    t = rnd_c32(rng, ntaps)
    sizes = sizes_near(rng, 4096, 8)
    x = rnd_c32(rng, sum(sizes))
    want, st = oracle.batch_fir(x, t, np.zeros(ntaps, np.complex64))
    node = cb.BatchFirNode(t)
    pos, outs = 0, []
    for s in sizes:
        outs.append(node.run(x[pos:pos + s]))
        assert len(outs[-1]) == s
        pos += s
    assert rel_l2(np.concatenate(outs), want) <= FIR_TOL, sizes
    assert node.state.tobytes() == st.tobytes()


@pytest.mark.parametrize("L,ntaps,cplx,seed", [(8, 1024, False, 1), (8, 512, True, 2), (4, 32, False, 3), (4, 100, True, 4), (8, 9, False, 5)])
def test_polyphase_tensor_core_batches_around_tiles(cb, oracle, L, ntaps, cplx, seed, monkeypatch):
    # K3-TC: tiles of 1920 (x8) / 4064 (x4) symbols; f32 and fused-i16 outputs on the same stream
    monkeypatch.setenv("COMMS_B200_FIR_PATH", "tc")
    rng = np.random.default_rng(100 + seed)
    t = rnd_c32(rng, ntaps) if cplx else rng.uniform(-1, 1, ntaps).astype(np.complex64)
    sizes = sizes_near(rng, 1920 if L == 8 else 4064, 6)
    sym = rnd_c32(rng, sum(sizes))
    want, st = oracle.batch_fir(oracle.upsample(sym, L), t, np.zeros(ntaps, np.complex64))
    a, b = cb.BatchFirNode(t, None, interp=L), cb.BatchFirNode(t, None, interp=L)
    pos, outs, q = 0, [], []
    for s in sizes:
        outs.append(a.run(sym[pos:pos + s]))
        q.append(b.run_i16(sym[pos:pos + s], 4096.0))
        assert len(outs[-1]) == s * L and len(q[-1]) == s * L
        pos += s
    got = np.concatenate(outs)
    assert rel_l2(got, want) <= FIR_TOL, sizes
    assert a.state.tobytes() == st.tobytes() == b.state.tobytes()
    assert np.array_equal(np.concatenate(q), oracle.quantize_i16(got, 4096.0).reshape(-1, 2))


@pytest.mark.parametrize("D,mix,fm,seed", [(10, True, True, 1), (5, False, True, 2), (10, True, False, 3), (5, True, True, 4)])
def test_chain_batches_around_tiles(cb, oracle, D, mix, fm, seed):
    # K2c (TMA span ring, 3840-sample tiles): batch lengths around tile multiples, f32 and u8 input alternating
    rng = np.random.default_rng(200 + seed)
    C = 3
    k = np.arange(63) - 31
    taps = (np.sinc(k / 5) * np.hamming(63) / 5).astype(np.float32).astype(np.complex64)
    dph = -2 * np.pi * rng.uniform(-0.4, 0.4, C)
    bank = cb.ChainBank(C, taps, D, dphase=dph if mix else None, with_fm=fm)
    refs = [oracle.FmChain(dph[c], 0.0, taps, D, do_mix=mix, do_fm=fm) for c in range(C)]
    for i, n in enumerate(sizes_near(rng, 3840, 7, lo=2)):
        n += n & 1  # the fused kernels want even batch lengths; odd ones take the generic kernel (covered elsewhere)
        if i % 2:
            iq = rng.integers(0, 256, (C, n, 2), dtype=np.uint8)
            x = oracle.u8_to_f32(iq.reshape(-1)).view(np.complex64).reshape(C, n)
            got = bank.run_u8(iq)
        else:
            x = rnd_c32(rng, C * n).reshape(C, n)
            got = bank.run(x)
        assert got.shape == (C, -(-n // D))
        for c in range(C):
            want = refs[c].run(x[c])
            if fm:
                d = np.abs(got[c].astype(np.float64) - want.astype(np.float64))
                d = np.minimum(d, 2 * np.pi - d)
                assert np.median(d) < 2e-6 and np.mean(d > 1e-3) < 5e-3, (c, n)
            else:
                assert rel_l2(got[c], want) <= FIR_TOL, (c, n)


@pytest.mark.parametrize("frames", [1, 2, 15, 16, 17, 33, 100])
def test_fft65536_frame_counts(cb, oracle, frames):
    # fused rows kernel: lag (16 frames) and ring (32 frames) boundaries
    rng = np.random.default_rng(frames)
    n = 65536
    x = rnd_c32(rng, frames * n)
    got = cb.FFTBatchNode(n, False).run(x)
    for f in sorted({0, frames - 1, frames // 2}):
        want = oracle.fft(x[f * n:(f + 1) * n], n, False)
        assert rel_l2(got[f * n:(f + 1) * n], want) <= 1e-4, f


@pytest.mark.parametrize("path", ["default", "fourstep"])
@pytest.mark.parametrize("log2n,frames", [(15, 1), (15, 70), (15, 400), (17, 3), (17, 40), (17, 100), (18, 11), (18, 50), (19, 9),
                                          (19, 26), (20, 2), (20, 10)])
def test_fft_large_sizes(cb, oracle, log2n, frames, path, monkeypatch):
    # K5-B (fused two-step kernel over an N1 x N2 split, intermediate in an L2 ring of 192 / 48 / 24 / 12 / 6 frames) and
    # the four-step fallback: frame counts below, at and beyond lag and ring; forward and inverse
    if path != "default":
        monkeypatch.setenv("COMMS_B200_FFT_PATH", path)
    n = 1 << log2n
    rng = np.random.default_rng(log2n * 100 + frames)
    x = rnd_c32(rng, frames * n)
    x[(frames // 2) * n:(frames // 2 + 1) * n] = 0
    x[(frames // 2) * n + 54321 % n] = 1  # an impulse frame: every output has modulus 1 and a known phase
    for inverse in (False, True):
        got = cb.FFTBatchNode(n, inverse).run(x)
        for f in sorted({0, 1 % frames, frames // 2, frames - 1}):
            want = oracle.fft(x[f * n:(f + 1) * n], n, inverse)
            assert rel_l2(got[f * n:(f + 1) * n], want) <= 1e-4, (f, inverse)
        # every frame: Parseval (unnormalised transform: sum |X|^2 = n sum |x|^2)
        e_in = (np.abs(x.reshape(frames, n).astype(np.complex128)) ** 2).sum(axis=1)
        e_out = (np.abs(got.reshape(frames, n).astype(np.complex128)) ** 2).sum(axis=1)
        assert np.all(np.abs(e_out / (n * e_in) - 1) < 1e-5)


@pytest.mark.parametrize("n,frames", [(1000, 5000), (3000, 1500), (12000, 100), (70000, 9)])
def test_fft_any_length_many_frames(cb, oracle, n, frames):
    # chirp-z path: more frames than one work-buffer group (32 MiB / M), device entry, direct path as a second opinion
    import torch

    rng = np.random.default_rng(n)
    x = rnd_c32(rng, frames * n)
    node = cb.FFTBatchNode(n, False)
    d_x = torch.from_numpy(x).cuda()
    d_y = torch.empty_like(d_x)
    ts = torch.cuda.Stream()
    torch.cuda.synchronize()
    node.run_dev(d_x.data_ptr(), frames * n, d_y.data_ptr(), ts.cuda_stream)
    torch.cuda.synchronize()
    got = d_y.cpu().numpy()
    for f in sorted({0, 1, frames // 2, frames - 2, frames - 1}):
        want = oracle.fft(x[f * n:(f + 1) * n], n, False)
        assert rel_l2(got[f * n:(f + 1) * n], want) <= 1e-4, f
    e_in = (np.abs(x.reshape(frames, n).astype(np.complex128)) ** 2).sum(axis=1)
    e_out = (np.abs(got.reshape(frames, n).astype(np.complex128)) ** 2).sum(axis=1)
    assert np.all(np.abs(e_out / (n * e_in) - 1) < 1e-5)
    back = cb.FFTBatchNode(n, True).run(got[:3 * n]) / n
    assert rel_l2(back, x[:3 * n]) <= 1e-4


def test_fft_direct_and_chirpz_agree(cb, monkeypatch):
    rng = np.random.default_rng(5)
    n = 1000
    x = rnd_c32(rng, 4 * n)
    a = cb.FFTBatchNode(n, False).run(x)
    monkeypatch.setenv("COMMS_B200_FFT_PATH", "direct")
    b = cb.FFTBatchNode(n, False).run(x)
    assert rel_l2(a, b) <= 2e-6


@pytest.mark.parametrize("n", [129, 1000, 4097, 8191])
def test_fft_chirpz_fused_and_split_forms_agree(cb, oracle, n):
    # M <= 16384: chirp / product / scaling folded into the two transforms; the five-launch form is the same algebra.
    # (The choice is read once per process, so the split form runs in a child interpreter.)
    import os
    import subprocess
    import sys
    import tempfile

    rng = np.random.default_rng(n)
    x = rnd_c32(rng, 7 * n)
    a = cb.FFTBatchNode(n, True).run(x)
    assert rel_l2(a, oracle.fft(x, n, True)) <= 1e-4
    with tempfile.TemporaryDirectory() as td:
        np.save(os.path.join(td, "x.npy"), x)
        code = ("import sys, numpy as np; sys.path.insert(0, %r); import comms_rs_b200 as cb;"
                "x = np.load(%r); np.save(%r, cb.FFTBatchNode(%d, True).run(x))"
                % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(td, "x.npy"),
                   os.path.join(td, "y.npy"), n))
        env = dict(os.environ, COMMS_B200_FFT_CHIRPZ="split")
        subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=240)
        b = np.load(os.path.join(td, "y.npy"))
    assert rel_l2(a, b) <= 2e-6


@pytest.mark.timeout(300)
@pytest.mark.parametrize("path", ["default", "rowspf"])
def test_fft65536_two_handles_concurrently(cb, oracle, path, monkeypatch):
    # two fused 65536-point kernels in flight on different streams share the SMs; the ticket-ordered work
    # distribution must let both finish (a static item-to-CTA assignment could deadlock here) with right results.
    # rowspf: the prefetching form, whose ninth warp holds two tickets ahead and waits for a dependency only after it
    # has released its own item
    import threading

    import torch

    if path != "default":
        monkeypatch.setenv("COMMS_B200_FFT_PATH", path)

    n, frames = 65536, 512
    xs = [torch.empty(frames * n, dtype=torch.complex64, device="cuda") for _ in range(2)]
    ys = [torch.empty_like(x) for x in xs]
    streams = [torch.cuda.Stream() for _ in range(2)]
    nodes = [cb.FFTBatchNode(n, False), cb.FFTBatchNode(n, True)]
    for i in range(2):
        cb.synth_uniform_dev(7 + i, 0, frames * n, xs[i].data_ptr(), streams[i].cuda_stream)
    torch.cuda.synchronize()

    def work(i):
        for _ in range(6):
            nodes[i].run_dev(xs[i].data_ptr(), frames * n, ys[i].data_ptr(), streams[i].cuda_stream)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    torch.cuda.synchronize()
    for i in range(2):
        for f in (0, frames - 1, 257):
            xin = oracle.synth_uniform_c32(7 + i, f * n, n)
            assert rel_l2(ys[i][f * n:(f + 1) * n].cpu().numpy(), oracle.fft(xin, n, bool(i))) <= 1e-4, (i, f)


@pytest.mark.parametrize("cplx", [False, True])
def test_fir_tensor_core_dynamic_range(cb, oracle, cplx, monkeypatch):
    # K1-TC scales every 4096-sample tile by its own power of two before the fp16 split: a quiet stretch next to a
    # loud one must keep the 1e-5 tolerance on its own, not only in the global norm
    monkeypatch.setenv("COMMS_B200_FIR_PATH", "tc")
    rng = np.random.default_rng(77 + cplx)
    t = rnd_c32(rng, 64) if cplx else rng.uniform(-1, 1, 64).astype(np.complex64)
    x = rnd_c32(rng, 120_000)
    x[20_000:40_000] *= np.float32(1e-5)
    x[60_000:80_000] *= np.float32(1e4)
    x[100_000:100_003] = 0
    want, _ = oracle.batch_fir(x, t, np.zeros(64, np.complex64))
    got = cb.BatchFirNode(t).run(x)
    assert rel_l2(got, want) <= FIR_TOL
    for lo, hi in ((20_000 + 4096 + 64, 40_000), (60_000, 80_000), (90_000, 120_000)):
        assert rel_l2(got[lo:hi], want[lo:hi]) <= FIR_TOL, (lo, hi)


def _windows_ok(got, want, win, tol):
    """every `win`-sample window on its own (a bad stretch cannot hide in the global norm); all-zero windows must be zero"""
    n = len(want) // win * win
    g = got[:n].astype(np.complex128).reshape(-1, win)
    w = want[:n].astype(np.complex128).reshape(-1, win)
    num = np.linalg.norm(g - w, axis=1)
    den = np.linalg.norm(w, axis=1)
    rel = np.where(den > 0, num / np.where(den > 0, den, 1), num)
    worst = int(np.argmax(rel))
    assert rel[worst] <= tol, (worst * win, float(rel[worst]))


@pytest.mark.parametrize("exp", [6, 8, 10])
@pytest.mark.parametrize("cplx", [False, True])
def test_fir_tensor_core_mixed_tile_windows(cb, oracle, cplx, exp, monkeypatch):
    # The tile that CONTAINS both the loud and the quiet stretch: one power-of-two scale per 4096-sample tile cannot
    # hold f32 accuracy there (the quiet samples' lo terms go denormal), the reference (plain f32, fir.rs:87-102) has no
    # such limit.  Such tiles are flagged by the converter warps and recomputed in f32 direct form (FirFix): every
    # 64-sample window must keep 1e-5, at in-tile dynamic ranges of 1e6, 1e8 and 1e10.
    monkeypatch.setenv("COMMS_B200_FIR_PATH", "tc")
    rng = np.random.default_rng(700 + exp + cplx)
    t = rnd_c32(rng, 64) if cplx else rng.uniform(-1, 1, 64).astype(np.complex64)
    x = rnd_c32(rng, 70_000)
    q = np.float32(10.0 ** -exp)
    x[9_000:11_000] *= q          # inside tile 2
    x[20_400:20_700] *= q         # across the edge of tiles 4 / 5
    x[33_000:41_500] *= q         # two whole tiles and parts of their neighbours
    x[50_000:50_200] = 0          # exact zeros are not a quiet stretch
    want, st = oracle.batch_fir(x, t, np.zeros(64, np.complex64))
    node = cb.BatchFirNode(t)
    got = node.run(x)
    _windows_ok(got, want, 64, FIR_TOL)
    assert node.state.tobytes() == st.tobytes()
    # batch invariance survives the fall-back: two calls, same stream
    node2 = cb.BatchFirNode(t)
    got2 = np.concatenate([node2.run(x[:35_000]), node2.run(x[35_000:])])
    _windows_ok(got2, want, 64, FIR_TOL)


@pytest.mark.parametrize("L,ntaps,cplx", [(8, 1024, False), (4, 32, False), (8, 128, True)])
@pytest.mark.parametrize("exp", [6, 10])
def test_polyphase_tensor_core_mixed_tile_windows(cb, oracle, L, ntaps, cplx, exp, monkeypatch):
    # same for the polyphase bank (fir_ptc_kernel.cu), symbol tiles of ~1900 symbols
    monkeypatch.setenv("COMMS_B200_FIR_PATH", "tc")
    rng = np.random.default_rng(900 + L + ntaps + exp)
    t = rnd_c32(rng, ntaps) if cplx else rng.uniform(-1, 1, ntaps).astype(np.complex64)
    sym = rnd_c32(rng, 12_000)
    q = np.float32(10.0 ** -exp)
    sym[2_500:3_100] *= q
    sym[5_600:9_900] *= q
    want, st = oracle.batch_fir(oracle.upsample(sym, L), t, np.zeros(ntaps, np.complex64))
    node = cb.BatchFirNode(t, None, interp=L)
    got = node.run(sym)
    _windows_ok(got, want, 64 * L, FIR_TOL)
    assert node.state.tobytes() == st.tobytes()
    if not cplx and L == 4:  # the fused quantiser takes the same fall-back
        scale = float(2.0 ** exp)
        i16 = cb.BatchFirNode(t, None, interp=L).run_i16(sym, scale)
        assert np.array_equal(i16, oracle.quantize_i16(got, scale).reshape(-1, 2))


@pytest.mark.parametrize("interp", [1, 8])
def test_tensor_core_nonfinite_samples_stay_local(cb, oracle, interp, monkeypatch):
    # an Inf / NaN sample poisons exactly the outputs whose taps reach it (K outputs, fir.rs:99), as in the reference --
    # not the whole 4096-sample tile that shares its block scale
    monkeypatch.setenv("COMMS_B200_FIR_PATH", "tc")
    rng = np.random.default_rng(1234 + interp)
    ntaps = 64 if interp == 1 else 256
    t = rnd_c32(rng, ntaps) if interp == 1 else rng.uniform(-1, 1, ntaps).astype(np.complex64)
    x = rnd_c32(rng, 40_000)
    x[12_345] = np.complex64(complex(np.inf, 0.5))
    x[20_479] = np.complex64(complex(np.nan, 0.0))    # last sample of a tile: reaches into the next one
    x[30_000] = np.complex64(complex(1.0, -np.inf))
    with np.errstate(all="ignore"):
        want, _ = oracle.batch_fir(oracle.upsample(x, interp) if interp > 1 else x, t, np.zeros(ntaps, np.complex64))
    got = cb.BatchFirNode(t, None, interp=interp).run(x)
    bad_w = ~(np.isfinite(want.real) & np.isfinite(want.imag))
    bad_g = ~(np.isfinite(got.real) & np.isfinite(got.imag))
    assert bad_w.sum() == 3 * ntaps
    assert np.array_equal(bad_g, bad_w)
    ok = ~bad_w
    g, w = np.where(ok, got, 0), np.where(ok, want, 0)
    _windows_ok(g, w, 64 * interp, FIR_TOL)


@pytest.mark.parametrize("ntaps,seed", [(129, 1), (500, 2), (1024, 3), (1025, 4)])
def test_long_fir_overlap_save_batches(cb, oracle, ntaps, seed):
    # 129 .. 1025 taps: fast convolution (4096-point FFT -> x Hf -> IFFT) from 8192 samples per call on, the direct
    # form below that; state carried across both kinds of call
    rng = np.random.default_rng(300 + seed)
    t = rnd_c32(rng, ntaps)
    hop = 4096 - (ntaps - 1)
    sizes = [8192, hop * 3, hop * 3 + 1, 100, 8191, 20_011, 8192 + hop - 1, 3]
    x = rnd_c32(rng, sum(sizes))
    want, st = oracle.batch_fir(x, t, np.zeros(ntaps, np.complex64))
    node = cb.BatchFirNode(t)
    pos, outs = 0, []
    for s in sizes:
        outs.append(node.run(x[pos:pos + s]))
        assert len(outs[-1]) == s
        pos += s
    got = np.concatenate(outs)
    assert rel_l2(got, want) <= FIR_TOL
    pos = 0
    for s in sizes:  # every call on its own, so a bad frame edge cannot hide in the global norm
        assert rel_l2(got[pos:pos + s], want[pos:pos + s]) <= 2 * FIR_TOL, (s, pos)
        pos += s
    assert node.state.tobytes() == st.tobytes()


def test_no_writes_outside_the_output_buffers(cb, monkeypatch):
    # every device entry writes exactly its output range: outputs sit inside guard regions filled with a sentinel
    # (compute-sanitizer is not available on the GPU pool; stray stores past a ragged tail would show up here)
    import torch

    monkeypatch.setenv("COMMS_B200_FIR_PATH", "tc")
    G = 4096  # guard bytes on both sides
    ts = torch.cuda.Stream()
    s = ts.cuda_stream

    def guarded(nbytes):
        buf = torch.full((G + nbytes + G,), 0xA5, dtype=torch.uint8, device="cuda")
        return buf, buf.data_ptr() + G

    def check(buf, nbytes, what):
        torch.cuda.synchronize()
        assert bool((buf[:G] == 0xA5).all()) and bool((buf[G + nbytes:] == 0xA5).all()), what

    rng = np.random.default_rng(5)
    for n in (4097, 12_345, 65_537):
        x = torch.empty(n, dtype=torch.complex64, device="cuda")
        cb.synth_uniform_dev(3, 0, n, x.data_ptr(), s)
        # plain FIR (tensor-core kernel forced), x4 / x8 polyphase in f32 and i16, long-filter fast convolution
        for taps, L in ((rnd_c32(rng, 64), 1), (rnd_c32(rng, 300), 1), (rng.uniform(-1, 1, 32).astype(np.complex64), 4),
                        (rnd_c32(rng, 200), 8)):
            buf, p = guarded(8 * n * L)
            fa, fb = cb.BatchFirNode(taps, None, interp=L), cb.BatchFirNode(taps, None, interp=L)
            fa.run_dev(x.data_ptr(), n, p, n * L, s)
            check(buf, 8 * n * L, ("fir", len(taps), L, n))
            buf, p = guarded(4 * n * L)
            fb.run_dev_i16(x.data_ptr(), n, 8192.0, p, n * L, s)
            check(buf, 4 * n * L, ("fir i16", len(taps), L, n))
        # stand-alone mixer / FM
        buf, p = guarded(8 * n)
        mixer = cb.MixerNode(0.3, 0.1)
        cb._lib.check(cb.load().cb_mixer_run_dev(mixer._h, x.data_ptr(), n, p, s))
        check(buf, 8 * n, ("mixer", n))
        buf, p = guarded(4 * n)
        node = cb.FMDemodNode()
        cb._lib.check(cb.load().cb_fm_run_dev(node._h, x.data_ptr(), n, p, s))
        check(buf, 4 * n, ("fm", n))
        # edge formats
        buf, p = guarded(4 * n)
        cb.quantize_i16_dev(x.data_ptr(), 2 * n, 8192.0, p, s)
        check(buf, 4 * n, ("quantize", n))
        b8 = torch.randint(0, 256, (2 * n,), dtype=torch.uint8, device="cuda")
        buf, p = guarded(8 * n)
        cb.convert_u8_dev(b8.data_ptr(), n, p, s)
        check(buf, 8 * n, ("convert_u8", n))
    # fused bank: f32 and byte input, ragged last tile
    k = np.arange(63) - 31
    taps = (np.sinc(k / 5) * np.hamming(63) / 5).astype(np.float32).astype(np.complex64)
    for D, fm in ((10, True), (5, False)):
        C, n = 3, 10_000
        bank = cb.ChainBank(C, taps, D, dphase=[0.1, -0.2, 0.3], with_fm=fm)
        no = bank.out_len(n)
        x = torch.empty(C * n, dtype=torch.complex64, device="cuda")
        cb.synth_uniform_dev(4, 0, C * n, x.data_ptr(), s)
        esz = 4 if fm else 8
        buf, p = guarded(esz * C * no)
        bank.run_dev(x.data_ptr(), n, p, no, s)
        check(buf, esz * C * no, ("chain", D, fm))
        b8 = torch.randint(0, 256, (C * n * 2,), dtype=torch.uint8, device="cuda")
        buf, p = guarded(esz * C * no)
        bank.run_dev_u8(b8.data_ptr(), n, p, no, s)
        check(buf, esz * C * no, ("chain u8", D, fm))
    # round 2: tensor-core byte front end (ragged last tile, odd n_out -> unaligned rows through the staged store path),
    # NCO scan, i16-IQ edges, the pipelined-cluster FFT
    monkeypatch.setenv("COMMS_B200_CHAIN_PATH", "tc")
    for fm in (True, False):
        C, n = 3, 10_248 * 2 + 40  # n_out = 4107 (odd): channel rows 1 and 2 start unaligned
        bank = cb.ChainBank(C, taps, 5, dphase=None, with_fm=fm)
        no = bank.out_len(n)
        b8 = torch.randint(0, 256, (C * n * 2,), dtype=torch.uint8, device="cuda")
        esz = 4 if fm else 8
        buf, p = guarded(esz * C * no)
        bank.run_dev_u8(b8.data_ptr(), n, p, no, s)
        check(buf, esz * C * no, ("chain tc", fm))
    monkeypatch.delenv("COMMS_B200_CHAIN_PATH")
    for n in (1, 2047, 2049, 70_001):
        e = torch.randn(n, dtype=torch.float64, device="cuda") * 0.01
        buf, p = guarded(16 * n)
        cb.NcoNode(0.3, 0.1).run_dev(e.data_ptr(), n, p, s)
        check(buf, 16 * n, ("nco", n))
    for n, L in ((70_001, 1), (66_000, 4)):
        iq = torch.randint(-3000, 3000, (2 * n,), dtype=torch.int16, device="cuda")
        t = rnd_c32(rng, 64) if L == 1 else rng.uniform(-1, 1, 32).astype(np.complex64)
        buf, p = guarded(4 * n * L)
        cb.BatchFirNode(t, None, interp=L).run_dev_iq16(iq.data_ptr(), n, 1.0 / 4096, 2048.0, p, n * L, s)
        check(buf, 4 * n * L, ("fir iq16", n, L))
    monkeypatch.setenv("COMMS_B200_FFT_PATH", "cpipe")
    for frames in (1, 20):
        x = torch.empty(65536 * frames, dtype=torch.complex64, device="cuda")
        cb.synth_uniform_dev(6, 0, 65536 * frames, x.data_ptr(), s)
        buf, p = guarded(8 * 65536 * frames)
        cb.FFTBatchNode(65536, False).run_dev(x.data_ptr(), 65536 * frames, p, s)
        check(buf, 8 * 65536 * frames, ("fft cpipe", frames))
    monkeypatch.delenv("COMMS_B200_FFT_PATH")
    # FFTs
    for N, frames in ((4096, 5), (16384, 3), (65536, 35)):
        x = torch.empty(N * frames, dtype=torch.complex64, device="cuda")
        cb.synth_uniform_dev(5, 0, N * frames, x.data_ptr(), s)
        buf, p = guarded(8 * N * frames)
        fft = cb.FFTBatchNode(N, False)
        fft.run_dev(x.data_ptr(), N * frames, p, s)
        check(buf, 8 * N * frames, ("fft", N))
