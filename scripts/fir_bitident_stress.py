"""GPU stress: the 64-tap tensor-core FIR over 2^28 samples, alternately in one call and in 256 calls of 2^20 with carried
state, N times; every result must equal the first bit for bit (one-shot == batched, run-to-run determinism).
usage (on a GPU box, from the repo root): python scripts/fir_bitident_stress.py [iterations]"""
import numpy as np, torch, sys
sys.path.insert(0, ".")
import comms_rs_b200 as cb
import oracle
n = 1 << 28
s = torch.cuda.current_stream().cuda_stream
x = torch.empty(n, dtype=torch.complex64, device="cuda")
cb.synth_uniform_dev(1234, 0, n, x.data_ptr(), s)
taps = oracle.rrc_taps(64, 4.0, 0.25)
ref = None
bad = 0
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
y = torch.empty(n, dtype=torch.complex64, device="cuda")
for rep in range(iters):
    node = cb.BatchFirNode(taps)
    y.zero_()
    if rep % 2 == 0:
        node.run_dev(x.data_ptr(), n, y.data_ptr(), n, s)
    else:
        B = 1 << 20
        for b in range(0, n, B):
            node.run_dev(x.data_ptr() + 8 * b, B, y.data_ptr() + 8 * b, B, s)
    torch.cuda.synchronize()
    if ref is None:
        ref = y.clone()
        continue
    d = (ref.view(torch.int32) != y.view(torch.int32)).view(-1, 2).any(dim=1)
    idx = torch.nonzero(d).flatten()
    if idx.numel():
        bad += 1
        t = torch.unique(idx // 4096)
        print("rep", rep, "batched" if rep % 2 else "one-shot", "mismatching samples", idx.numel(), "tiles", t[:10].tolist(), "n tiles", t.numel(),
              "offsets in tile", (idx[:6] % 4096).tolist(), "tile%256", (t[:10] % 256).tolist())
        k = int(idx[0])
        print("   ", ref[k].item(), y[k].item())
print("done", iters, "bad", bad)
