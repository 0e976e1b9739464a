"""BEV tokeniser: the head of the reference's ``VATLiDAR.forward`` on the B200 path.

Reference: src/encoder-decoder/training/models/vat_lidar.py
    ctor    :63-121   refine (depthwise 3x3 + GELU), proj (1x1), norm_tokens, geo_mlp, view_embed
    _grid   :123-185  geometry (x, y, r, sin, cos) and 6-way sector id per cell, cached per (H, W, device)
    forward :206-253  tokens = norm_tokens(proj(refine(bev))) + geo_mlp(geom) + view_embed[sector]   -> [B, H*W, d_model]

:class:`VATLiDARTokenizer` owns exactly those parameters under exactly those names, so
``tokenizer.load_state_dict(vat_lidar.state_dict(), strict=False)`` takes them from a reference checkpoint (queries, blocks
and the output head are not part of this path).  ``forward(bev)`` is the reference's call; ``forward_pillars`` produces the
same tokens from pillar rows without ever materialising the dense canvas.  Inference only; CUDA (sm_100) only: the
arithmetic lives in libpillars_b200.so (csrc/tokens.cu) and there is no fallback.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _native, ops
from ._native import NativeLibraryError, PillarsTokenizer, check

NUM_VIEWS = 6  # vat_lidar.py:39


def grid_tables(h: int, w: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Host-side geometry tables of ``VATLiDAR._grid`` (vat_lidar.py:139-183): geom [h*w,5] fp32, sector [h*w] int32.
    Input independent, a few KB to MB, built once per (h, w) on the CPU with the same elementary functions the reference
    calls; the sector thresholds are applied in fp32 like the reference's tensor comparisons."""
    yv, xv = torch.meshgrid(torch.linspace(-1.0, 1.0, h), torch.linspace(-1.0, 1.0, w), indexing="ij")
    r = torch.clamp((xv * xv + yv * yv).sqrt(), 0.0, 1.0)
    theta = torch.atan2(yv, xv)
    geom = torch.stack((xv, yv, r, torch.sin(theta), torch.cos(theta)), dim=-1).reshape(h * w, 5).contiguous()
    ft = theta.reshape(-1)
    third, two3 = math.pi / 3, 2 * math.pi / 3
    # sector order of the reference (:166-181): 60-degree bins counted from -180 degrees map to ids 5,3,4,1,0,2
    sid = torch.full((h * w,), 2, dtype=torch.int32)  # [120, 180] incl. +pi
    sid[ft < two3] = 0
    sid[ft < third] = 1
    sid[ft < 0.0] = 4
    sid[ft < -third] = 3
    sid[ft < -two3] = 5
    return geom, sid


class VATLiDARTokenizer(nn.Module):
    """Drop-in for the tokenising part of ``VATLiDAR`` (same constructor arguments for that part, same parameter names)."""

    def __init__(self, c_in: int, d_model: int, projection: str = "auto"):
        """``projection`` selects how the 1x1 projection runs (all variants are fp32-accurate and parity-tested):
        ``"fma"``   fp32 FFMA2 on the FMA pipes inside one fused kernel, any supported shape;
        ``"umma"``  tcgen05.mma.kind::tf32 as a 3-term hi/lo split, accumulator in tensor memory, the active cells of the
                    batch compacted into 128-row tiles (c_in 32 or 64, d_model 128 or 256): 1.31 ms against 1.42 ms of
                    ``"fma"`` on 16 x 512^2 at d = 256, 0.79 against 0.96 ms at d = 128;
        ``"auto"``  ``"umma"`` where it exists, else ``"fma"``."""
        super().__init__()
        if projection not in ("auto", "fma", "umma"):
            raise ValueError("projection must be one of auto / fma / umma")
        umma_ok = c_in in (32, 64) and d_model in (128, 256)
        self.projection = ("umma" if umma_ok else "fma") if projection == "auto" else projection
        self._umma: Optional[torch.Tensor] = None
        self.c_in, self.d_model = int(c_in), int(d_model)
        self.refine = nn.Sequential(nn.Conv2d(c_in, c_in, kernel_size=3, padding=1, groups=c_in), nn.GELU())
        self.proj = nn.Conv2d(c_in, d_model, kernel_size=1, bias=True)
        self.norm_tokens = nn.LayerNorm(d_model)
        self.geo_mlp = nn.Sequential(nn.Linear(5, d_model), nn.GELU(), nn.Linear(d_model, d_model))
        self.view_embed = nn.Parameter(torch.zeros(NUM_VIEWS, d_model))
        self._packed: Optional[Dict[str, torch.Tensor]] = None
        self._tables: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor]] = {}

    # ---- parameters in the layout the kernels read, rebuilt when the module moves or loads a checkpoint ---------------
    def _apply(self, fn, *a, **k):
        self._packed, self._tables, self._umma = None, {}, None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed, self._tables, self._umma = None, {}, None
        return super().load_state_dict(*a, **k)

    def _pack(self, dev) -> Dict[str, torch.Tensor]:
        if self._packed is None:
            f32 = dict(device=dev, dtype=torch.float32)
            d, c = self.d_model, self.c_in
            self._packed = {
                "dw_w": self.refine[0].weight.detach().to(**f32).reshape(c, 9).contiguous(),
                "dw_b": self.refine[0].bias.detach().to(**f32).contiguous(),
                "wt": self.proj.weight.detach().to(**f32).reshape(d, c).t().contiguous(),
                "pb": self.proj.bias.detach().to(**f32).contiguous(),
                "gamma": self.norm_tokens.weight.detach().to(**f32).contiguous(),
                "beta": self.norm_tokens.bias.detach().to(**f32).contiguous(),
                "w1": self.geo_mlp[0].weight.detach().to(**f32).contiguous(),
                "b1": self.geo_mlp[0].bias.detach().to(**f32).contiguous(),
                "w2t": self.geo_mlp[2].weight.detach().to(**f32).t().contiguous(),
                "b2": self.geo_mlp[2].bias.detach().to(**f32).contiguous(),
                "view": self.view_embed.detach().to(**f32).contiguous(),
            }
        return self._packed

    def _native_struct(self, pk, pe=None, bg=None) -> PillarsTokenizer:
        t = PillarsTokenizer()
        t.c_in, t.d_model = self.c_in, self.d_model
        t.dw_weight, t.dw_bias = pk["dw_w"].data_ptr(), pk["dw_b"].data_ptr()
        t.proj_weight_t, t.proj_bias = pk["wt"].data_ptr(), pk["pb"].data_ptr()
        t.ln_weight, t.ln_bias, t.ln_eps = pk["gamma"].data_ptr(), pk["beta"].data_ptr(), float(self.norm_tokens.eps)
        t.pe = None if pe is None else pe.data_ptr()
        t.background = None if bg is None else bg.data_ptr()
        t.proj_umma = self._umma.data_ptr() if (self._umma is not None and pe is not None) else None
        return t

    def _device(self) -> torch.device:
        dev = self.view_embed.device
        if dev.type != "cuda":
            raise NativeLibraryError("VATLiDARTokenizer runs on a B200 only: move the module to a CUDA device (no CPU fallback)")
        ops._require_device(self.view_embed)
        return dev

    def tables(self, h: int, w: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(PE [h*w, d], background [d]) on the device, computed once per (h, w) -- the counterpart of the reference's
        ``_cache`` (:119-121), extended to geo_mlp's output because the weights are frozen in eval mode."""
        if self.training:
            raise RuntimeError("VATLiDARTokenizer is inference-only (tables are cached): call .eval()")
        key = (int(h), int(w))
        if key not in self._tables:
            dev = self._device()
            pk = self._pack(dev)
            geom, sid = grid_tables(h, w)
            geom, sid = geom.to(dev), sid.to(dev)
            pe = torch.empty((h * w, self.d_model), dtype=torch.float32, device=dev)
            bg = torch.empty((self.d_model,), dtype=torch.float32, device=dev)
            nat = self._native_struct(pk)
            umma = None
            if self.projection == "umma" and self._umma is None and self.c_in in (32, 64) and self.d_model in (128, 256):
                umma = torch.empty((2 * self.c_in * self.d_model,), dtype=torch.float32, device=dev)
            check(_native.load().pillars_tokens_prepare(ctypes.byref(nat), geom.data_ptr(), sid.data_ptr(), h, w,
                                                        pk["w1"].data_ptr(), pk["b1"].data_ptr(), pk["w2t"].data_ptr(),
                                                        pk["b2"].data_ptr(), pk["view"].data_ptr(), pe.data_ptr(),
                                                        bg.data_ptr(), None if umma is None else umma.data_ptr(),
                                                        ops._stream_ptr()), "pillars_tokens_prepare")
            # computed once, then read by every later call on ANY stream: finish it here (one-time cost)
            torch.cuda.current_stream(dev).synchronize()
            if umma is not None:
                self._umma = umma
            self._tables[key] = (pe, bg)
        return self._tables[key]

    # ---- the reference's call -----------------------------------------------------------------------------------------
    def forward(self, bev: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """bev [B, c_in, H, W] fp32 -> tokens [B, H*W, d_model] (vat_lidar.py:206-253)."""
        if bev.dim() != 4 or bev.shape[1] != self.c_in:
            raise ValueError(f"expected bev [B, {self.c_in}, H, W], got {tuple(bev.shape)}")
        dev = self._device()
        if bev.device != dev or bev.dtype != torch.float32:
            raise ValueError("bev must be a float32 tensor on the module's device")
        bev = bev.contiguous()
        b, _, h, w = bev.shape
        pe, bg = self.tables(h, w)
        lib = _native.load()
        tokens = self._out(out, b, h, w, dev)
        need = lib.pillars_tokens_workspace_bytes(b, self.c_in, h, w, 1)
        ws = ops.workspace(need, dev, slot=7)
        nat = self._native_struct(self._pack(dev), pe, bg)
        check(lib.pillars_bev_tokens_dense(bev.data_ptr(), b, h, w, ctypes.byref(nat), tokens.data_ptr(), ws.data_ptr(),
                                           ws.numel(), ops._stream_ptr()), "pillars_bev_tokens_dense")
        return tokens

    def _out(self, out, b, h, w, dev):
        if out is None:
            return torch.empty((b, h * w, self.d_model), dtype=torch.float32, device=dev)
        if tuple(out.shape) != (b, h * w, self.d_model) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 [B, H*W, d_model] tensor")
        return out

    # ---- canvas-free calls -----------------------------------------------------------------------------------------------
    def forward_pillars(self, pillar_features: torch.Tensor, voxel_coords: torch.Tensor, batch_size: int, grid_hw,
                        pillar_count: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Same tokens from the scatter's INPUTS (``pillar_features`` [M, c_in], ``voxel_coords`` [M,4] (b,z,y,x) int32 or
        fp32) -- what ``forward(PointPillarScatter(...))`` returns, without the dense canvas in between.
        ``pillar_count``: optional device int32 whose LAST element is the live row count (rows beyond it are ignored)."""
        dev = self._device()
        h, w = int(grid_hw[0]), int(grid_hw[1])
        pe, bg = self.tables(h, w)
        feats = pillar_features.contiguous()
        coords = voxel_coords.contiguous()
        if feats.dim() != 2 or feats.shape[1] != self.c_in or feats.dtype != torch.float32:
            raise ValueError(f"pillar_features must be float32 [M, {self.c_in}]")
        if coords.dtype not in (torch.int32, torch.float32) or coords.shape != (feats.shape[0], 4):
            raise ValueError("voxel_coords must be int32 or float32 [M, 4]")
        lib = _native.load()
        tokens = self._out(out, batch_size, h, w, dev)
        need = lib.pillars_tokens_workspace_bytes(batch_size, self.c_in, h, w, 0)
        ws = ops.workspace(need, dev, slot=7)
        nat = self._native_struct(self._pack(dev), pe, bg)
        m_dev = None if pillar_count is None else pillar_count[-1:].data_ptr()
        check(lib.pillars_bev_tokens(feats.data_ptr(), coords.data_ptr(), int(coords.dtype == torch.float32), feats.shape[0],
                                     m_dev, batch_size, h, w, ctypes.byref(nat), tokens.data_ptr(), ws.data_ptr(),
                                     ws.numel(), ops._stream_ptr()), "pillars_bev_tokens")
        return tokens

    def forward_index_map(self, pillar_features: torch.Tensor, cell_row: torch.Tensor,
                          out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Same tokens from pillar rows and an existing BEV index map ``cell_row`` [B, H, W] int32 (-1 = empty), e.g. the
        one :func:`ops.encode_bev` leaves in its workspace (:func:`encode_index_map`)."""
        dev = self._device()
        if cell_row.dtype != torch.int32 or cell_row.dim() != 3 or not cell_row.is_contiguous():
            raise ValueError("cell_row must be a contiguous int32 [B, H, W] tensor")
        b, h, w = cell_row.shape
        pe, bg = self.tables(h, w)
        feats = pillar_features.contiguous()
        tokens = self._out(out, b, h, w, dev)
        nat = self._native_struct(self._pack(dev), pe, bg)
        lib = _native.load()
        ws_ptr, ws_len = None, 0
        if self._umma is not None:  # the tcgen05 variant keeps its (frame, cell) pair list in a workspace
            ws = ops.workspace(lib.pillars_tokens_workspace_bytes(b, self.c_in, h, w, 0), dev, slot=7)
            ws_ptr, ws_len = ws.data_ptr(), ws.numel()
        check(lib.pillars_bev_tokens_map(feats.data_ptr(), cell_row.data_ptr(), b, h, w, ctypes.byref(nat), tokens.data_ptr(),
                                         ws_ptr, ws_len, ops._stream_ptr()), "pillars_bev_tokens_map")
        return tokens


def encode_index_map(buffers: "ops.EncodeBuffers", n_points: int, n_frames: int, grid: "ops.GridSpec") -> torch.Tensor:
    """View of the BEV index map [n_frames, ny, nx] that ``ops.encode_bev(points, ..., buffers=buffers)`` left in its
    workspace (valid until ``buffers`` is reused).  ``n_points`` must be the row count of the ``points`` of THAT call: the
    workspace is carved by it.  ``ops.encode_bev(..., want_index_map=True)`` returns the same view without that pitfall."""
    g = grid.native()
    off = _native.load().pillars_workspace_cell_row_offset(n_points, n_frames, ctypes.byref(g))
    nx, ny, _ = grid.grid_size
    nbytes = 4 * n_frames * ny * nx
    return buffers.ws[off:off + nbytes].view(torch.int32).view(n_frames, ny, nx)
