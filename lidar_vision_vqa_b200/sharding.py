"""Multi-GPU layout of the pillar path: frames are independent units, so a batch is split contiguously over the ranks
(one process per GPU) and each rank runs the whole path on its own frames with NO data-path collective.  The only
exchange is the optional hand-over of BEV tokens to the rank that runs the consumer (the VQA model's VATLiDAR,
src/encoder-decoder/training/models/vat_lidar.py:187-304): the COMPACT form -- pillar features ``[M_r, F]`` and
coordinates ``[M_r, 4]`` -- is gathered (about 3.7 MB per frame instead of the 64 MiB dense canvas) and densified on the
destination with the same scatter kernel.

The reference has no collective on this path (frames go to disk); its gather helper for variable-length payloads
(src/lidar-encoder/pcdet/utils/commu_utils.py:50-111: sizes first, then padded payloads) is the model for
:func:`gather_bev_tokens`, minus the pickling.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_frames: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous split; the first ``n_frames % world_size`` ranks take one extra frame."""
    base, extra = divmod(n_frames, world_size)
    out, lo = [], 0
    for r in range(world_size):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def shard_points(points: torch.Tensor, frame_offsets: torch.Tensor, rank: int, world_size: int):
    """Rows of the packed ``points [N, C]`` owned by ``rank`` plus re-based offsets.  ``frame_offsets`` is a host or
    device ``[B+1]`` int tensor; the result's offsets start at 0 and its frames are numbered from 0 locally."""
    offs = frame_offsets.to("cpu", torch.int64)
    lo, hi = shard_bounds(offs.numel() - 1, world_size)[rank]
    p0, p1 = int(offs[lo]), int(offs[hi])
    local = (offs[lo:hi + 1] - p0).to(torch.int32)
    return points[p0:p1], local, lo


@dataclass
class GatheredTokens:
    """Compact BEV tokens of the whole batch on the destination rank (None elsewhere)."""

    pillar_features: torch.Tensor  # [sum M, F]
    voxel_coords: torch.Tensor     # [sum M, 4] (b,z,y,x), b is the GLOBAL frame index
    pillars_per_rank: List[int]
    n_frames: int


def gather_bev_tokens(pillar_features: torch.Tensor, voxel_coords: torch.Tensor, frame_base: int, n_frames_total: int,
                      dst: int = 0, group: Optional[dist.ProcessGroup] = None) -> Optional[GatheredTokens]:
    """Every rank contributes its ``[M_r, F]`` features and ``[M_r, 4]`` coordinates (local frame numbering); ``dst``
    receives the concatenation in rank order with frame indices shifted to global numbering.  Two collectives: an
    all-gather of the row counts, then one gather of padded payloads (features and coordinates travel in one buffer:
    the int32 coordinates are bit-cast into four extra float columns)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        c = voxel_coords.to(torch.int32).clone()
        c[:, 0] += frame_base
        return GatheredTokens(pillar_features, c, [pillar_features.shape[0]], n_frames_total)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = pillar_features.device
    m, f = pillar_features.shape
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([m], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts_h = counts.cpu().tolist()
    m_max = max(counts_h) if counts_h else 0
    payload = torch.zeros((m_max, f + 4), dtype=torch.float32, device=dev)
    payload[:m, :f] = pillar_features
    coords = voxel_coords.to(torch.int32).clone()
    coords[:, 0] += frame_base
    payload[:m, f:] = coords.view(torch.float32)
    if rank == dst:
        bufs = [torch.empty_like(payload) for _ in range(world)]
        dist.gather(payload, bufs, dst=dst, group=group)
        feats = torch.cat([b[:c, :f] for b, c in zip(bufs, counts_h)], dim=0)
        crd = torch.cat([b[:c, f:].contiguous().view(torch.int32) for b, c in zip(bufs, counts_h)], dim=0)
        return GatheredTokens(feats, crd, counts_h, n_frames_total)
    dist.gather(payload, None, dst=dst, group=group)
    return None


def wire_layout(rows: int, f: int, frames: int, feat_bytes: int = 4):
    """Byte offsets of the compact result of one step as ONE message: features ``[rows, f]`` | coordinates ``[rows, 4]``
    int32 | counts ``[frames + 1]`` int32 (padded to 16 bytes).  The same layout as ``ops.EncodeBuffers(wire_slab=True)``."""
    a = rows * f * feat_bytes
    b = rows * 16
    c = (4 * (frames + 1) + 15) // 16 * 16
    return a, b, c


class TokenGatherer:
    """Steady-state hand-over of compact BEV tokens to the fusion rank: one message per rank and step, no host
    synchronisation, no per-step size negotiation, no concatenation afterwards.

    Every rank's encoder writes ``pillar_features``, ``voxel_coords`` and ``pillar_count`` into ONE contiguous slab
    (``ops.EncodeBuffers(capacity=rows, wire_slab=True)``) that is sent in place; the destination receives the slabs into
    fixed segments of one buffer and exposes the views ``feats [W, rows, F]``, ``coords [W, rows, 4]``, ``counts
    [W, B_r + 1]``.  ``rows`` is a capacity agreed once (e.g. 1.1 x the pillar count of a warm-up batch): the encoder drops
    pillars beyond it but still reports the true count, the counts travel with the payload and stay on the device, where
    :func:`ops.rebase_segments` turns padding rows into frame ``-1`` (skipped by the scatter and by the tokeniser), shifts
    live rows to global frame numbers and raises the ``overflow`` flag when a count did not fit; :meth:`check` reads that flag
    outside the hot loop.

    The transfers of a step are one NCCL group (``batch_isend_irecv``) issued on the gatherer's own high-priority stream,
    ordered after the encoder by an event, so step k's transfer overlaps step k+1's kernels.  This is the shape of the
    reference's variable-length gather helper (src/lidar-encoder/pcdet/utils/commu_utils.py:50-111: sizes first, then
    payloads padded to the maximum, pickled) with its weaknesses removed.  ``dtype=torch.float16`` sends the feature rows as
    float16 (one cast on the sender), halving the bytes that bound the destination's NVLink ingress."""

    def __init__(self, rows: int, f: int, frames_per_rank: int, device, dst: int = 0, dtype: torch.dtype = torch.float32,
                 group: Optional[dist.ProcessGroup] = None, slots: int = 2):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dst, self.rows, self.f, self.frames_per_rank = int(dst), int(rows), int(f), int(frames_per_rank)
        self.device, self.dtype = torch.device(device), dtype
        self.cuda = self.device.type == "cuda"
        self.stream = torch.cuda.Stream(device=self.device, priority=-1) if self.cuda else None
        self.overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.feat_bytes = 2 if dtype == torch.float16 else 4
        self.a32, self.b, self.c = wire_layout(self.rows, self.f, self.frames_per_rank, 4)
        self.a = self.rows * self.f * self.feat_bytes
        self.slab_bytes = self.a + self.b + self.c
        self.slots = []
        for _ in range(max(1, slots)):
            slot = {"sent": torch.cuda.Event() if self.cuda else None, "ready": torch.cuda.Event() if self.cuda else None}
            if dtype != torch.float32:  # the cast needs a staging slab on the sender
                slot["stage"] = torch.zeros(self.slab_bytes, dtype=torch.uint8, device=self.device)
            if self.rank == self.dst:
                w = torch.zeros((self.world, self.slab_bytes), dtype=torch.uint8, device=self.device)
                slot["wire"] = w
                slot["feats"] = w[:, :self.a].view(dtype).view(self.world, self.rows, self.f)
                slot["coords"] = w[:, self.a:self.a + self.b].view(torch.int32).view(self.world, self.rows, 4)
                slot["counts"] = w[:, self.a + self.b:self.a + self.b + 4 * (self.frames_per_rank + 1)].view(torch.int32)
            self.slots.append(slot)
        self._k = 0

    def bytes_per_rank(self) -> int:
        return self.slab_bytes

    def exchange(self, wire: torch.Tensor, after: Optional["torch.cuda.Event"] = None):
        """Enqueues the transfer of one step.  ``wire`` is the encoder's slab (``EncodeBuffers.wire`` with
        ``capacity == rows``); it must stay untouched until the returned slot's ``sent`` event.  On the destination the slot's
        ``feats`` / ``coords`` / ``counts`` views are valid after its ``ready`` event (coords already rebased).  Returns
        the slot."""
        from . import ops

        slot = self.slots[self._k % len(self.slots)]
        self._k += 1
        if wire.dtype != torch.uint8 or wire.numel() != self.a32 + self.b + self.c:
            raise ValueError("wire must be the uint8 slab of EncodeBuffers(capacity=rows, wire_slab=True)")
        ctx = torch.cuda.stream(self.stream) if self.cuda else _null_ctx()
        with ctx:
            if self.cuda:
                if after is not None:
                    self.stream.wait_event(after)
                else:
                    self.stream.wait_stream(torch.cuda.current_stream(self.device))
            send = wire
            if self.dtype != torch.float32:
                st = slot["stage"]
                st[:self.a].view(self.dtype).copy_(wire[:self.a32].view(torch.float32))
                st[self.a:].copy_(wire[self.a32:])
                send = st
            if self.rank == self.dst:
                slot["wire"][self.rank].copy_(send)
                if self.world > 1:
                    p2p = [dist.P2POp(dist.irecv, slot["wire"][r], r, self.group) for r in range(self.world) if r != self.dst]
                    for w in dist.batch_isend_irecv(p2p):
                        w.wait()
            else:
                for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, send, self.dst, self.group)]):
                    w.wait()
            if self.cuda:
                slot["sent"].record(self.stream)
                if self.rank == self.dst:
                    ops.rebase_segments(slot["coords"], slot["counts"], self.frames_per_rank, overflow=self.overflow)
                    slot["ready"].record(self.stream)
        return slot

    def check(self) -> None:
        """Host-side check on the destination (synchronises): raises when some step's pillar count did not fit ``rows``."""
        if self.cuda and int(self.overflow.item()) != 0:
            raise RuntimeError(f"TokenGatherer: a rank produced more than rows={self.rows} pillars; rows were dropped")


class _null_ctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def densify_segments(slot: dict, n_frames_total: int, nx: int, ny: int, variant: str = "auto",
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Dense ``[B, F, ny, nx]`` canvas of a :class:`TokenGatherer` slot on the destination (fp32 feature rows)."""
    from . import ops

    feats = slot["feats"].reshape(-1, slot["feats"].shape[-1])  # (a copy when the segments are strided views)
    if feats.dtype != torch.float32:
        feats = feats.float()
    return ops.scatter_bev(feats, slot["coords"].reshape(-1, 4), n_frames_total, nx, ny, 1, variant=variant, out=out)


def densify(tokens: GatheredTokens, nx: int, ny: int, variant: str = "auto") -> torch.Tensor:
    """Dense ``[B, F, ny, nx]`` canvas of gathered tokens on the destination GPU (the same scatter kernel as the
    single-GPU path)."""
    from . import ops

    return ops.scatter_bev(tokens.pillar_features, tokens.voxel_coords, tokens.n_frames, nx, ny, 1, variant=variant)


def bind_to_gpu_numa_node(device_index: int) -> Optional[List[int]]:
    """Restricts this process to the CPUs NVML reports as local to GPU ``device_index`` (one process per GPU, so pinned
    host buffers allocated afterwards are first-touched on that GPU's NUMA node and the host->device copies of the ranks do
    not all cross the same socket link).  Returns the CPU list, or None when NVML / affinity is unavailable."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[device_index]) if vis else device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:  # noqa: BLE001 - best effort: no NVML, no permission, exotic topology
        return None
