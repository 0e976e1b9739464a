"""BaseBEVBackbone on the tcgen05 convolution kernel.

Mirrors `pcdet/models/backbones_2d/base_bev_backbone.py:6-112`: same constructor arguments, same sub-module tree -- so the
state-dict keys (`blocks.0.1.weight`, `blocks.0.2.running_mean`, `deblocks.2.0.weight`, ...) and therefore the reference's
checkpoints load unchanged -- same `forward(data_dict) -> data_dict` writing `spatial_features_2d` (NCHW fp32).

What differs is how eval-mode forward computes: every Conv2d/ConvTranspose2d + BatchNorm2d + ReLU triple is ONE launch of
`pillars_conv_forward` (csrc/conv_umma.cu), activations stay NHWC between layers, the three up-sampling branches write their
channel windows of the concatenated output directly (no `torch.cat`), and the first layer gathers its input from the pillar
rows through the BEV index map when the VFE left them in `data_dict` (`pillar_features` + `bev_index_map`): the dense
`spatial_features` canvas is then not read at all.  Given only `spatial_features`, the canvas is first compacted into that
form (`pillars_canvas_to_rows`).  There is no PyTorch fallback: training mode and shapes outside the kernel's set raise."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from . import _native

_EPS, _MOMENTUM = 1e-3, 0.01  # base_bev_backbone.py:36


def _cfg_get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    getter = getattr(cfg, "get", None)
    if callable(getter):
        return getter(key, default)
    return getattr(cfg, key, default)


class _Layer:
    """One fused conv + BN + ReLU launch: geometry plus the prepared weight image."""

    def __init__(self, conv: nn.Module, bn: nn.BatchNorm2d, transposed: bool):
        self.conv, self.bn, self.transposed = conv, bn, transposed
        w = conv.weight
        if transposed:
            c_in, c_out, k = w.shape[0], w.shape[1], w.shape[2]
            stride = conv.stride[0]
            if k != stride or k not in (1, 2, 4):
                raise NotImplementedError(f"ConvTranspose2d kernel {k} stride {stride}: only kernel == stride in (1, 2, 4)")
            self.desc = dict(c_in=c_in, c_out=c_out, k=1, stride=1, pad=0, up=k)
        else:
            c_out, c_in, k = w.shape[0], w.shape[1], w.shape[2]
            self.desc = dict(c_in=c_in, c_out=c_out, k=k, stride=conv.stride[0], pad=1 if k == 3 else 0, up=1)
        self.image: Optional[torch.Tensor] = None
        self.shift: Optional[torch.Tensor] = None
        self._key = None

    def native(self, round_out: bool) -> _native.PillarsConv:
        d = self.desc
        return _native.PillarsConv(d["c_in"], d["c_out"], d["k"], d["stride"], d["pad"], d["up"], 1, 1 if round_out else 0)

    def prepare(self, device: torch.device) -> None:
        params = (self.conv.weight, self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var)
        def version(t):  # inference tensors (a module moved under torch.inference_mode) do not track one
            try:
                return t._version
            except RuntimeError:
                return -1

        key = tuple((t.data_ptr(), version(t)) for t in params) + (str(device),)
        if key == self._key:
            return
        lib = _native.load()
        with torch.no_grad():
            w = self.conv.weight.detach().to(device=device, dtype=torch.float32)
            if self.transposed and self.desc["up"] == 1:  # ConvTranspose2d(k=1, s=1) is a 1x1 convolution with [in, out] weights
                w = w.permute(1, 0, 2, 3)
            w = w.contiguous()
            scale = (self.bn.weight.detach().float() / torch.sqrt(self.bn.running_var.detach().float() + self.bn.eps)).to(device)
            shift = (self.bn.bias.detach().float().to(device) - self.bn.running_mean.detach().float().to(device) * scale)
            scale, shift = scale.contiguous(), shift.contiguous()
        cv = self.native(False)
        nbytes = lib.pillars_conv_weight_bytes(cv)
        if nbytes == 0:
            raise NotImplementedError(f"convolution {self.desc} is outside the set the tcgen05 kernel is instantiated for")
        image = torch.empty(nbytes // 4, dtype=torch.float32, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        _native.check(lib.pillars_conv_prepare(cv, w.data_ptr(), scale.data_ptr(), image.data_ptr(), stream), "pillars_conv_prepare")
        # built once, then read by every later call on ANY stream (a pipeline runs each batch on its own): finish it here
        torch.cuda.current_stream(device).synchronize()
        self.image, self.shift, self._key = image, shift, key


def conv_forward(layer: _Layer, out: torch.Tensor, out_c_total: int, out_c_off: int, out_nchw: bool, n_frames: int, h_in: int,
                 w_in: int, x_nhwc: Optional[torch.Tensor] = None, rows: Optional[torch.Tensor] = None,
                 cell_row: Optional[torch.Tensor] = None, round_out: bool = True, error: Optional[torch.Tensor] = None) -> None:
    lib = _native.load()
    dev = out.device
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = lib.pillars_conv_forward(layer.native(round_out), layer.image.data_ptr(), layer.shift.data_ptr(),
                                  x_nhwc.data_ptr() if x_nhwc is not None else None,
                                  rows.data_ptr() if rows is not None else None,
                                  cell_row.data_ptr() if cell_row is not None else None, n_frames, h_in, w_in, out.data_ptr(),
                                  out_c_total, out_c_off, 1 if out_nchw else 0, error.data_ptr() if error is not None else None,
                                  stream)
    _native.check(rc, "pillars_conv_forward")


class BaseBEVBackbone(nn.Module):
    """Drop-in for the reference class of the same name (registry `backbones_2d.__all__['BaseBEVBackbone']`)."""

    def __init__(self, model_cfg, input_channels):
        super().__init__()
        self.model_cfg = model_cfg
        layer_nums = _cfg_get(model_cfg, "LAYER_NUMS")
        if layer_nums is not None:
            layer_strides, num_filters = _cfg_get(model_cfg, "LAYER_STRIDES"), _cfg_get(model_cfg, "NUM_FILTERS")
            assert len(layer_nums) == len(layer_strides) == len(num_filters)
        else:
            layer_nums = layer_strides = num_filters = []
        upsample_strides = _cfg_get(model_cfg, "UPSAMPLE_STRIDES")
        if upsample_strides is not None:
            num_upsample_filters = _cfg_get(model_cfg, "NUM_UPSAMPLE_FILTERS")
            assert len(upsample_strides) == len(num_upsample_filters)
        else:
            upsample_strides = num_upsample_filters = []
        levels = len(layer_nums)
        c_in_list = [input_channels, *num_filters[:-1]]
        self.blocks, self.deblocks = nn.ModuleList(), nn.ModuleList()

        def bn(c):
            return nn.BatchNorm2d(c, eps=_EPS, momentum=_MOMENTUM)

        for i in range(levels):
            # sub-module positions are part of the checkpoint format: [pad, conv, bn, relu] + [conv, bn, relu] * n
            seq: List[nn.Module] = [nn.ZeroPad2d(1),
                                    nn.Conv2d(c_in_list[i], num_filters[i], 3, stride=layer_strides[i], padding=0, bias=False),
                                    bn(num_filters[i]), nn.ReLU()]
            for _ in range(layer_nums[i]):
                seq += [nn.Conv2d(num_filters[i], num_filters[i], 3, padding=1, bias=False), bn(num_filters[i]), nn.ReLU()]
            self.blocks.append(nn.Sequential(*seq))
            if len(upsample_strides) > 0:
                s = upsample_strides[i]
                if s > 1 or (s == 1 and not _cfg_get(model_cfg, "USE_CONV_FOR_NO_STRIDE", False)):
                    up = nn.ConvTranspose2d(num_filters[i], num_upsample_filters[i], s, stride=s, bias=False)
                else:
                    k = int(round(1 / s))
                    up = nn.Conv2d(num_filters[i], num_upsample_filters[i], k, stride=k, bias=False)
                self.deblocks.append(nn.Sequential(up, bn(num_upsample_filters[i]), nn.ReLU()))
        c_in = sum(num_upsample_filters)
        if len(upsample_strides) > levels:
            s = upsample_strides[-1]
            self.deblocks.append(nn.Sequential(nn.ConvTranspose2d(c_in, c_in, s, stride=s, bias=False), bn(c_in), nn.ReLU()))
        self.num_bev_features = c_in
        self._plan = None

    # ---- the fused plan ---------------------------------------------------------------------------------------------------
    def _build_plan(self):
        blocks = []
        for seq in self.blocks:
            mods = list(seq)
            layers = [_Layer(mods[1], mods[2], False)]
            for j in range(4, len(mods), 3):
                layers.append(_Layer(mods[j], mods[j + 1], False))
            blocks.append(layers)
        de = [_Layer(seq[0], seq[1], isinstance(seq[0], nn.ConvTranspose2d)) for seq in self.deblocks]
        if len(de) > len(blocks):
            raise NotImplementedError("the trailing deblock over the concatenated branches is not instantiated")
        if len(de) == 0 and len(blocks) != 1:
            raise NotImplementedError("multi-level backbone without up-sampling branches")
        self._plan = (blocks, de)

    @staticmethod
    def _out_hw(h, w, d):
        if d["up"] > 1:
            return d["up"] * h, d["up"] * w
        return (h + 2 * d["pad"] - d["k"]) // d["stride"] + 1, (w + 2 * d["pad"] - d["k"]) // d["stride"] + 1

    def output_shape(self, h: int, w: int):
        """(channels, height, width) of ``spatial_features_2d`` for an ``h x w`` input canvas."""
        if self._plan is None:
            self._build_plan()
        blocks, de = self._plan
        shapes = []
        for layers in blocks:
            for l in layers:
                h, w = self._out_hw(h, w, l.desc)
            shapes.append((h, w))
        if de:
            oh, ow = self._out_hw(*shapes[0], de[0].desc)
            return sum(l.desc["c_out"] for l in de), oh, ow
        return blocks[0][-1].desc["c_out"], shapes[0][0], shapes[0][1]

    def forward(self, data_dict):
        if self.training:
            raise NotImplementedError("BaseBEVBackbone (B200) is an inference path: call .eval(); training runs the reference module")
        if self._plan is None:
            self._build_plan()
        blocks, de = self._plan
        rows, cell_row, x = None, None, None
        if data_dict.get("bev_index_map") is not None and data_dict.get("pillar_features") is not None:
            rows = data_dict["pillar_features"]
            cell_row = data_dict["bev_index_map"]
            dev = rows.device
            nb, h, w = cell_row.shape
        else:
            canvas = data_dict["spatial_features"]
            if canvas.dtype != torch.float32 or not canvas.is_cuda:
                raise _native.NativeLibraryError("spatial_features must be a float32 CUDA tensor (no CPU path)")
            canvas = canvas.contiguous()
            dev = canvas.device
            nb, c, h, w = canvas.shape
            lib = _native.load()
            cell_row = torch.empty((nb, h, w), dtype=torch.int32, device=dev)
            rows = torch.empty((max(nb * h * w, 1), c), dtype=torch.float32, device=dev)
            counter = torch.empty(1, dtype=torch.int32, device=dev)
            _native.check(lib.pillars_canvas_to_rows(canvas.data_ptr(), nb, c, h, w, cell_row.data_ptr(), rows.data_ptr(),
                                                     counter.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                          "pillars_canvas_to_rows")
        if rows.dtype != torch.float32 or cell_row.dtype != torch.int32:
            raise _native.NativeLibraryError("pillar rows must be float32 and the index map int32")
        rows, cell_row = rows.contiguous(), cell_row.contiguous()
        for layers in blocks:
            for l in layers:
                l.prepare(dev)
        for l in de:
            l.prepare(dev)
        error = torch.zeros(1, dtype=torch.int32, device=dev)
        h0, w0 = h, w

        # output geometry first: every branch must land on the same image
        shapes = []
        hh, ww = h, w
        for layers in blocks:
            for l in layers:
                hh, ww = self._out_hw(hh, ww, l.desc)
            shapes.append((hh, ww))
        if de:
            outs = [self._out_hw(*shapes[i], de[i].desc) for i in range(len(blocks))]
            if any(o != outs[0] for o in outs):
                raise ValueError(f"up-sampling branches disagree on the output size: {outs}")
            oh, ow = outs[0]
            c_total = sum(l.desc["c_out"] for l in de)
        else:
            oh, ow = shapes[0]
            c_total = blocks[0][-1].desc["c_out"]
        result = torch.empty((nb, c_total, oh, ow), dtype=torch.float32, device=dev)

        c_off = 0
        for i, layers in enumerate(blocks):
            for j, l in enumerate(layers):
                oh_l, ow_l = self._out_hw(h, w, l.desc)
                last_plain = (not de) and i == len(blocks) - 1 and j == len(layers) - 1
                if last_plain:
                    conv_forward(l, result, c_total, 0, True, nb, h, w, x_nhwc=x, rows=rows if x is None else None,
                                 cell_row=cell_row if x is None else None, round_out=False, error=error)
                else:
                    y = torch.empty((nb, oh_l, ow_l, l.desc["c_out"]), dtype=torch.float32, device=dev)
                    conv_forward(l, y, l.desc["c_out"], 0, False, nb, h, w, x_nhwc=x, rows=rows if x is None else None,
                                 cell_row=cell_row if x is None else None, round_out=True, error=error)
                    x = y
                h, w = oh_l, ow_l
            if de:
                conv_forward(de[i], result, c_total, c_off, True, nb, h, w, x_nhwc=x, round_out=False, error=error)
                c_off += de[i].desc["c_out"]
        data_dict["spatial_features_2d"] = result
        data_dict["_conv_error_word"] = error
        return data_dict
