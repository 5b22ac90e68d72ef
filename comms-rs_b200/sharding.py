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


def halo_state(prev_samples, k: int, interp: int = 1):
    """Reference `state` for a segment: the k entries of the delay line before it, newest first, zero padded when
    fewer exist (start of stream).  interp = L > 1: the delay line is in the zero-stuffed domain
    (UpsampleNode -> BatchFirNode, PulseNode): the i-th most recent input symbol sits at index i*L - 1, zeros
    between -- a 1024-tap x8 bank needs the 127 symbols before the segment."""
    import numpy as np

    prev = np.asarray(prev_samples)
    st = np.zeros(k, dtype=prev.dtype if prev.size else np.complex64)
    L = max(int(interp), 1)
    m = min(k // L, len(prev))
    if m:
        st[L - 1: m * L: L] = prev[len(prev) - m:][::-1]
    return st


_TWO_PI_50 = "6.28318530717958647692528676655900576839433879875021"


def segment_phase(phase0: float, dphase: float, start: int) -> float:
    """Mixer phase at sample `start` of the stream, wrapped to [0, 2 pi).  start * dphase is formed exactly
    (both are rationals) and reduced against 2 pi to 50 digits, so the result is the correctly rounded phase for
    any stream position (a plain f64 product is already 2e-6 rad off at 2^31 samples)."""
    from fractions import Fraction

    two_pi = Fraction(_TWO_PI_50)
    p = Fraction(phase0) + Fraction(int(start)) * Fraction(dphase)
    p -= (p // two_pi) * two_pi
    v = float(p)
    return v if v < 2 * math.pi else 0.0


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

    def _counts(self, counts):
        import ctypes as C

        if len(counts) != self.nranks:
            raise ValueError("counts must have one entry per rank")
        return (C.c_size_t * self.nranks)(*[int(c) for c in counts])

    def gather_to_root_dev(self, d_seg: int, counts, elem_bytes: int, root: int, d_all: int, stream: int = 0) -> None:
        """Ordered gather of per-rank segments of any lengths onto `root` only (grouped ncclSend / ncclRecv):
        counts[r] elements of elem_bytes from rank r land at d_all + sum(counts[:r]).  Every rank passes the same
        counts, so empty segments are skipped consistently."""
        self._lib.check(self._lib.load().cb_gather_segments_to_root_dev(
            self._h, d_seg, self._counts(counts), int(elem_bytes), int(root), d_all, stream))

    def allgather_var_dev(self, d_seg: int, counts, elem_bytes: int, d_all: int, stream: int = 0) -> None:
        """The same ordered stream on every rank (segments of any lengths)."""
        self._lib.check(self._lib.load().cb_allgather_segments_var_dev(
            self._h, d_seg, self._counts(counts), int(elem_bytes), d_all, stream))

    def close(self) -> None:
        if self._h:
            self._lib.load().cb_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def peer_export(d_ptr: int) -> bytes:
    """64-byte CUDA IPC handle of a device allocation (its base pointer) for cb_peer_open on another rank."""
    import ctypes as C

    from . import _lib

    buf = C.create_string_buffer(64)
    _lib.check(_lib.load().cb_peer_export(d_ptr, buf))
    return buf.raw


def peer_open(handle: bytes) -> int:
    """Maps another rank's exported allocation; the returned device pointer is valid as d_out of any *_run_dev."""
    import ctypes as C

    from . import _lib

    out = C.c_void_p()
    _lib.check(_lib.load().cb_peer_open(C.create_string_buffer(bytes(handle), 64), C.byref(out)))
    return out.value


def peer_close(d_mapped: int) -> None:
    from . import _lib

    _lib.check(_lib.load().cb_peer_close(d_mapped))
