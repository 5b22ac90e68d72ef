"""Multi-GPU partitioning of the hot path: only what shards naturally.

  * independent channels / FFT frames: contiguous blocks per rank, no exchange;
  * one long stream: contiguous overlap-save segments; rank r's halo is the K
    samples before its segment, handed to the node as the reference's initial
    `state` (BatchFirNode::new(taps, Some(state)), src/filter/fir_node.rs:193-200),
    its mixer phase is phase0 + start * dphase (MixerNode::new(dphase, Some(phase)),
    src/mixer.rs:128-134).  Segment starts are kept on the decimation grid.
The only collective is the ordered gather of segment outputs (NCCL on GPUs,
gloo in the CPU tests).
"""
from __future__ import annotations

import math


def segment_bounds(total: int, world: int, rank: int, multiple: int = 1) -> tuple[int, int]:
    """[start, end) of rank's segment; every start is a multiple of `multiple`."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    multiple = max(int(multiple), 1)
    per = -(-total // world)
    per = -(-per // multiple) * multiple
    start = min(rank * per, total)
    return start, min(start + per, total)


def block_shard(items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block of independent items (channels, frames) owned by rank."""
    base, extra = divmod(items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def halo_state(prev_samples, k: int):
    """Reference `state` for a segment: the k samples before it, newest first, zero
    padded when fewer exist (start of stream)."""
    import numpy as np

    prev = np.asarray(prev_samples)
    st = np.zeros(k, dtype=prev.dtype if prev.size else np.complex64)
    m = min(k, len(prev))
    if m:
        st[:m] = prev[len(prev) - m:][::-1]
    return st


def segment_phase(phase0: float, dphase: float, start: int) -> float:
    """Mixer phase at sample `start` of the stream, wrapped to [0, 2 pi)."""
    return math.fmod(phase0 + math.fmod(start * dphase, 2 * math.pi), 2 * math.pi) % (2 * math.pi)


def gather_ordered(local, sizes=None, group=None):
    """Concatenate every rank's output segment in rank order on every rank.
    `local` is a 1-D torch tensor; `sizes` the per-rank lengths when they differ."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if sizes is None:
        sizes = [local.numel()] * world
    if len(set(sizes)) == 1:
        out = torch.empty(world * sizes[0], dtype=local.dtype, device=local.device)
        if local.is_complex():  # gloo has no complex all_gather
            dist.all_gather_into_tensor(torch.view_as_real(out).reshape(-1), torch.view_as_real(local).reshape(-1), group=group)
        else:
            dist.all_gather_into_tensor(out, local, group=group)
        return out
    cap = max(sizes)
    pad = torch.zeros(cap, dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    full = gather_ordered(pad, [cap] * world, group)
    return torch.cat([full[r * cap: r * cap + sizes[r]] for r in range(world)])


class SegmentGather:
    """The same ordered gather through the library's own C-ABI entry (cb_comm_* / cb_gather_segments_dev:
    ncclAllGather over NVLink), for hosts that do not use torch.distributed.  `id_bytes` is the 128-byte id made
    by rank 0 with `SegmentGather.unique_id()` and handed to the other ranks over any host channel."""

    def __init__(self, nranks: int, rank: int, id_bytes: bytes):
        import ctypes as C

        from . import _lib

        self._lib, self._h = _lib, C.c_void_p()
        buf = C.create_string_buffer(bytes(id_bytes), 128)
        _lib.check(_lib.load().cb_comm_init(int(nranks), int(rank), buf, C.byref(self._h)))
        self.nranks, self.rank = int(nranks), int(rank)

    @staticmethod
    def unique_id() -> bytes:
        import ctypes as C

        from . import _lib

        buf = C.create_string_buffer(128)
        _lib.check(_lib.load().cb_comm_unique_id(buf))
        return buf.raw

    def gather_dev(self, d_seg: int, n_samples: int, d_all: int, stream: int = 0) -> None:
        """d_all (device, nranks * n_samples complex f32) <- every rank's n_samples-sample segment, in rank order."""
        self._lib.check(self._lib.load().cb_gather_segments_dev(self._h, d_seg, n_samples, d_all, stream))

    def close(self) -> None:
        if self._h:
            self._lib.load().cb_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
