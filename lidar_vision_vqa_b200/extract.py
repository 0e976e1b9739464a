"""Bulk BEV extraction: the B200 counterpart of the product's caller of the LiDAR encoder,
src/get-data/precompute_bev_features.py:295-411.  There the loop is

    DataLoader(CPU voxelisation) -> load_data_to_gpu (padded voxels) -> model.forward -> FeatureCatcher hook
    -> np.save(<token>.npy, bev.astype(np.float16))          (:350-395)

strictly serial, one ``cuda.synchronize`` per batch.  Here raw sweeps go in and float16 ``[C, H, W]`` maps come out, with the
host->device copy of batch k+1, the kernels of batch k and the device->host copy of batch k-1 in flight at the same time;
the float32 -> float16 conversion happens inside the scatter kernel (the canvas is written once, in its final dtype).

What is tapped follows the reference's hook priority (:254: ``spatial_features_2d`` > ``encoded_spconv_tensor`` >
``spatial_features``): with a ``backbone`` (``ship='features2d'``) the extractor saves ``spatial_features_2d`` -- what the
product stores for a pillar model with a ``BACKBONE_2D`` -- computed by :class:`backbone.BaseBEVBackbone` straight from the
pillar rows (the canvas is never materialised); without one it saves ``spatial_features``.

    ex = BevExtractor(vfe, batch_size=16, max_points_per_frame=40_000)
    for token, bev in ex.run(iter_of_(token, points[N, C])):      # bev: np.float16 [C, H, W]
        ...
    ex.run_to_dir(items, out_dir)                                  # writes <out_dir>/<token>.npy like the reference
"""
from __future__ import annotations

import os
from typing import Iterable, Iterator, List, Sequence, Tuple

import numpy as np
import torch

from .modules import PillarVFEFromPoints
from .pipeline import PillarEncoderPipeline


class BevExtractor:
    def __init__(self, vfe: PillarVFEFromPoints, batch_size: int, max_points_per_frame: int, depth: int = 3,
                 out_dtype: torch.dtype = torch.float16, ship: str = "dense", backbone=None):
        """``ship='dense'``: the float16 canvas crosses the host link (32 MiB per 512^2 frame), which is what the reference
        stores.  ``ship='compact'``: only the occupied cells cross it -- float16 pillar rows ``[M, F]`` + int16 ``(y, x)``
        ``[M, 2]`` per frame (~1.8 MB per frame at cfg2) -- and ``run`` yields ``(token, (rows, yx))``;
        :func:`densify_compact` rebuilds the identical canvas on the consumer's side.  ``ship='features2d'`` (needs
        ``backbone``, an eval-mode ``BaseBEVBackbone`` on the same device): ``run`` yields the float16
        ``spatial_features_2d [C2, H2, W2]`` of each frame (12 MiB per frame for the nuScenes pillar model)."""
        if ship not in ("dense", "compact", "features2d"):
            raise ValueError("ship must be 'dense', 'compact' or 'features2d'")
        if (ship == "features2d") != (backbone is not None):
            raise ValueError("ship='features2d' and backbone= go together")
        self.ship = ship
        self.vfe = vfe
        self.batch_size = int(batch_size)
        self.depth = max(2, int(depth))
        self.c = vfe.num_raw_point_features
        self.max_points = int(max_points_per_frame) * self.batch_size
        self.pipe = PillarEncoderPipeline(vfe, n_frames=self.batch_size, max_points=self.max_points, depth=self.depth,
                                          bev_dtype=out_dtype, with_bev=ship == "dense", backbone=backbone)
        self.d2h_bytes_last = 0
        # pinned staging: collated points in, canvases out (one pair per pipeline slot)
        self._pts = [torch.empty((self.max_points, self.c + 1), dtype=torch.float32).pin_memory()
                     for _ in range(self.depth)]
        nx, ny, nz = vfe.grid.grid_size
        f = vfe.get_output_feature_dim()
        self._bev = self._rows = self._yx = self._f2d = None
        if ship == "features2d":
            c2, h2, w2 = backbone.output_shape(ny, nx)
            self._f2d = [torch.empty((self.batch_size, c2, h2, w2), dtype=torch.float16).pin_memory() for _ in range(self.depth)]
        elif ship == "dense":
            self._bev = [torch.empty((self.batch_size, f * nz, ny, nx), dtype=out_dtype).pin_memory()
                         for _ in range(self.depth)]
        else:
            cap = self.pipe.slots[0].buffers.capacity
            self._rows = [torch.empty((cap, f), dtype=torch.float16).pin_memory() for _ in range(self.depth)]
            self._yx = [torch.empty((cap, 2), dtype=torch.int16).pin_memory() for _ in range(self.depth)]
        self._copied = [torch.cuda.Event() for _ in range(self.depth)]

    def _collate(self, frames: Sequence[np.ndarray], slot: int) -> torch.Tensor:
        """datasets/dataset.py:237-244: frame index prepended as column 0, frames concatenated."""
        buf = self._pts[slot]
        n = 0
        for b, f in enumerate(frames):
            f = np.asarray(f, dtype=np.float32)
            if f.ndim != 2 or f.shape[1] < self.c:
                raise ValueError(f"frame {b}: expected [N, >={self.c}] points")
            k = f.shape[0]
            if n + k > buf.shape[0]:
                raise ValueError("batch exceeds max_points_per_frame * batch_size")
            view = buf[n:n + k].numpy()
            view[:, 0] = b
            view[:, 1:] = f[:, :self.c]
            n += k
        return buf[:n]

    def run(self, items: Iterable[Tuple[str, np.ndarray]]) -> Iterator[Tuple[str, np.ndarray]]:
        """Yields ``(token, bev)`` in input order; ``bev`` is a numpy view into a pinned buffer that stays valid until
        ``depth - 2`` further batches have been yielded (the buffer is refilled as soon as its slot is submitted again: copy
        the array to keep it longer; ``run_to_dir`` writes it out at once)."""
        pending: List[Tuple[int, List[str], int]] = []  # (ticket, tokens, slot)
        batch_tokens: List[str] = []
        batch_frames: List[np.ndarray] = []
        slot = 0

        def collect():
            ticket, tokens, s = pending.pop(0)
            res = self.pipe.result(ticket)
            stream = self.pipe.slots[ticket % len(self.pipe.slots)].stream
            if self.ship in ("features2d", "dense"):  # the device->host copy was enqueued right behind the kernels (flush)
                self._copied[s].synchronize()
                buf = self._f2d[s] if self.ship == "features2d" else self._bev[s]
                self.d2h_bytes_last = buf.numel() * buf.element_size()
                arr = buf.numpy()
                for i, tok in enumerate(tokens):
                    yield tok, arr[i]
                return
            counts = res["pillars_per_frame"].numpy()
            m = int(counts.sum())
            with torch.cuda.stream(stream):  # (the casts are plain dtype conversions of the rows that cross the link)
                self._rows[s][:m].copy_(res["pillar_features"].to(torch.float16), non_blocking=True)
                self._yx[s][:m].copy_(res["voxel_coords"][:, 2:4].to(torch.int16), non_blocking=True)
                self._copied[s].record()
            self._copied[s].synchronize()
            self.d2h_bytes_last = m * (self._rows[s].shape[1] * 2 + 4)
            rows, yx = self._rows[s].numpy(), self._yx[s].numpy()
            lo = 0
            for i, tok in enumerate(tokens):
                hi = lo + int(counts[i])
                yield tok, (rows[lo:hi], yx[lo:hi])
                lo = hi

        def flush():
            nonlocal slot, batch_tokens, batch_frames
            pts = self._collate(batch_frames, slot)
            ticket = self.pipe.submit(pts)
            pslot = self.pipe.slots[ticket % len(self.pipe.slots)]
            if self.ship in ("features2d", "dense"):
                # what crosses the link does not depend on the pillar counts: copy it out on the batch's own stream, so the
                # transfer of batch k runs under the kernels of batch k + 1 (the cast is a plain dtype conversion)
                with torch.cuda.stream(pslot.stream):
                    if self.ship == "features2d":
                        self._f2d[slot].copy_(pslot.features_2d.to(torch.float16), non_blocking=True)
                    else:
                        self._bev[slot].copy_(pslot.buffers.bev, non_blocking=True)
                    self._copied[slot].record()
            pending.append((ticket, batch_tokens, slot))
            slot = (slot + 1) % self.depth
            batch_tokens, batch_frames = [], []

        for tok, pts in items:
            batch_tokens.append(tok)
            batch_frames.append(pts)
            if len(batch_frames) == self.batch_size:
                if len(pending) == self.depth - 1:
                    yield from collect()
                flush()
        if batch_frames:
            if len(pending) == self.depth - 1:
                yield from collect()
            flush()
        while pending:
            yield from collect()

    @staticmethod
    def densify_compact(rows: np.ndarray, yx: np.ndarray, ny: int, nx: int) -> np.ndarray:
        """The ``[C, H, W]`` float16 map of one frame from its compact form (consumer side, numpy)."""
        bev = np.zeros((rows.shape[1], ny, nx), dtype=np.float16)
        bev[:, yx[:, 0].astype(np.int64), yx[:, 1].astype(np.int64)] = rows.T
        return bev

    def run_to_dir(self, items: Iterable[Tuple[str, np.ndarray]], out_dir: str) -> int:
        """``np.save(<out_dir>/<token>.npy, bev)`` per sample, float16 ``[C, H, W]`` (precompute_bev_features.py:391-395)."""
        os.makedirs(out_dir, exist_ok=True)
        n = 0
        nx, ny, _ = self.vfe.grid.grid_size
        for tok, bev in self.run(items):
            if self.ship == "compact":
                bev = self.densify_compact(bev[0], bev[1], ny, nx)
            np.save(os.path.join(out_dir, f"{tok}.npy"), bev)
            n += 1
        return n
