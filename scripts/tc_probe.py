"""Probe of the tensor-core FIR path (fir_tc_kernel.cu) against the CPU oracle.
usage: python scripts/tc_probe.py [mode ...]   (COMMS_B200_TC_DESC_MODE values to try)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

modes = [0]
os.environ["COMMS_B200_FIR_PATH"] = "tc"
import comms_rs_b200 as cb
import oracle

cb.init(0)
rng = np.random.default_rng(1)


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.complex128) - b) / max(np.linalg.norm(b), 1e-30))


for mode in modes:
    os.environ["COMMS_B200_TC_DESC_MODE"] = str(mode)
    for ntaps, n, cplx, scale in [(64, 4096, True, 1.0), (64, 20000, True, 1.0), (64, 300001, True, 1e-4), (33, 9000, False, 37.0),
                                  (17, 5000, True, 1.0), (100, 70000, True, 1.0), (128, 12345, True, 1e3)]:
        x = ((rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)) * scale).astype(np.complex64)
        t = rng.uniform(-1, 1, ntaps) + (1j * rng.uniform(-1, 1, ntaps) if cplx else 0)
        t = t.astype(np.complex64)
        st = (rng.uniform(-1, 1, ntaps) + 1j * rng.uniform(-1, 1, ntaps)).astype(np.complex64) * scale
        want, _ = oracle.batch_fir(x, t, st.copy())
        node = cb.BatchFirNode(t, st.copy())
        t0 = time.time()
        got = node.run(x)
        e = rel(got, want)
        # second call: state carried
        x2 = x[: n // 3]
        want2, _ = oracle.batch_fir(np.concatenate([x, x2]), t, st.copy())
        got2 = node.run(x2)
        e2 = rel(got2, want2[n:])
        print(f"mode={mode} ntaps={ntaps} n={n} cplx={cplx} scale={scale}: rel_l2={e:.3e} second_call={e2:.3e} "
              f"{'OK' if e < 1e-5 and e2 < 1e-5 else 'FAIL'} ({time.time() - t0:.2f}s)", flush=True)
