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
        c = voxel_coords.clone()
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
