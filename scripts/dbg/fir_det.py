import numpy as np, torch, sys
sys.path.insert(0, ".")
import comms_rs_b200 as cb
n = 1 << 28
s = torch.cuda.current_stream().cuda_stream
x = torch.empty(n, dtype=torch.complex64, device="cuda")
cb.synth_uniform_dev(1234, 0, n, x.data_ptr(), s)
rng = np.random.default_rng(1)
import oracle
taps = oracle.rrc_taps(64, 4.0, 0.25)
ys = []
for rep in range(3):
    node = cb.BatchFirNode(taps)
    y = torch.empty(n, dtype=torch.complex64, device="cuda")
    node.run_dev(x.data_ptr(), n, y.data_ptr(), n, s)
    torch.cuda.synchronize()
    ys.append(y)
node2 = cb.BatchFirNode(taps)
y2 = torch.empty(n, dtype=torch.complex64, device="cuda")
B = 1 << 20
for b in range(0, n, B):
    node2.run_dev(x.data_ptr() + 8 * b, B, y2.data_ptr() + 8 * b, B, s)
torch.cuda.synchronize()
ys.append(y2)
for i in range(1, 4):
    d = (ys[0].view(torch.int32) != ys[i].view(torch.int32)).view(-1, 2).any(dim=1)
    idx = torch.nonzero(d).flatten()
    print("pair 0 vs", i, "mismatching samples:", idx.numel())
    if idx.numel():
        t = torch.unique(idx // 4096)
        print("  tiles:", t[:20].tolist(), "count", t.numel(), " tile %256:", torch.unique(t % 256)[:20].tolist())
        k = idx[:8].tolist()
        print("  first idx:", k, [(i % 4096) for i in k])
        print("  vals", ys[0][k[0]].item(), ys[i][k[0]].item())
