"""2+ GPU check of the segment + halo partitioning with the C-ABI gather (run under torchrun):
every rank filters its own overlap-save segment of one stream (halo = the reference's initial `state`),
cb_gather_segments_dev (ncclAllGather) reassembles the stream, rank 0 compares it with the one-GPU result."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import comms_rs_b200 as cb  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    cb.init(local)
    dist.init_process_group("gloo")  # host channel for the 128-byte id only
    ids = [cb.sharding.SegmentGather.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    g = cb.sharding.SegmentGather(world, rank, ids[0])
    seed, per, ntaps = 4242, 1 << 22, 64
    taps = (cb.rrc_taps(ntaps, 4.0, 0.25) * np.exp(0.1j * np.arange(ntaps))).astype(np.complex64)
    ts = torch.cuda.Stream()
    s = ts.cuda_stream
    start = rank * per
    x = torch.empty(per, dtype=torch.complex64, device="cuda")
    cb.synth_uniform_dev(seed, start, per, x.data_ptr(), s)
    halo = None
    if start:
        h = torch.empty(ntaps, dtype=torch.complex64, device="cuda")
        cb.synth_uniform_dev(seed, start - ntaps, ntaps, h.data_ptr(), s)
        torch.cuda.synchronize()
        halo = h.cpu().numpy()[::-1].copy()
    node = cb.BatchFirNode(taps, halo)
    y = torch.empty(per, dtype=torch.complex64, device="cuda")
    full = torch.empty(world * per, dtype=torch.complex64, device="cuda")
    node.run_dev(x.data_ptr(), per, y.data_ptr(), per, s)
    g.gather_dev(y.data_ptr(), per, full.data_ptr(), s)
    torch.cuda.synchronize()
    if rank == 0:
        xa = torch.empty(world * per, dtype=torch.complex64, device="cuda")
        ya = torch.empty_like(xa)
        cb.synth_uniform_dev(seed, 0, world * per, xa.data_ptr(), s)
        cb.BatchFirNode(taps).run_dev(xa.data_ptr(), world * per, ya.data_ptr(), world * per, s)
        torch.cuda.synchronize()
        err = float((full - ya).abs().double().norm() / ya.abs().double().norm())
        print(f"gathered stream of {world} segments vs one-GPU result: rel-L2 {err:.3e}", flush=True)
        assert err <= 2e-6, err
        print("mg_gather_check ok", flush=True)
    dist.barrier()
    g.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
