"""CPU emulation of device code: the FFT v2 core (comms-rs_b200/csrc/fft2_core.cuh) is host/device
code; tests/emul/fft2_emul.cpp runs every pass for every thread of a frame on the CPU, compares
with an O(N^2) DFT in double and checks that the padded shared-memory layout is bank-conflict free.
No GPU and no oracle needed."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fft2_core_emulation(tmp_path):
    exe = tmp_path / "fft2_emul"
    src = os.path.join(ROOT, "tests", "emul", "fft2_emul.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I/usr/local/cuda/include", src, "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("fft2 emul")]
    assert len(lines) >= 10 and all(l.endswith("OK") for l in lines), r.stdout
    # the padded layout must be conflict free (degree 1) for every pass of every size >= 256
    for l in lines:
        n = int(l.split("N=")[1].split()[0])
        if n >= 256:
            assert "worst_bank_conflict=1 " in l, l
