"""BaseBEVBackbone at the cfg2 shape (16 x 64 x 512^2 pillar canvas): this repo's tcgen05 path vs the reference's own module
(oracle/_ref copy) eager on the same GPU (cuDNN, TF32 allowed = PyTorch default, and TF32 off).  Prints per-layer kernel times.
Usage: python profiles/scripts/backbone_times.py [frames] [config]"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lidar_vision_vqa_b200 import _native, backbone as B

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
name = sys.argv[2] if len(sys.argv) > 2 else "nuscenes_multihead"
CFG = {"nuscenes_multihead": dict(LAYER_NUMS=[3, 5, 5], LAYER_STRIDES=[2, 2, 2], NUM_FILTERS=[64, 128, 256],
                                   UPSAMPLE_STRIDES=[0.5, 1, 2], NUM_UPSAMPLE_FILTERS=[128, 128, 128]),
       "kitti_pointpillar": dict(LAYER_NUMS=[3, 5, 5], LAYER_STRIDES=[2, 2, 2], NUM_FILTERS=[64, 128, 256],
                                 UPSAMPLE_STRIDES=[1, 2, 4], NUM_UPSAMPLE_FILTERS=[128, 128, 128])}[name]
dev = torch.device("cuda:0")
h = w = 512
g = torch.Generator().manual_seed(0)
occ = torch.rand((nb, h, w), generator=g) < 0.05
pos = occ.nonzero()
idx = torch.full((nb, h, w), -1, dtype=torch.int32)
idx[pos[:, 0], pos[:, 1], pos[:, 2]] = torch.arange(len(pos), dtype=torch.int32)
rows = torch.rand((len(pos), 64), generator=g)
rows_d, idx_d = rows.to(dev), idx.to(dev)
canvas = torch.zeros((nb, 64, h, w), device=dev)
canvas.permute(0, 2, 3, 1)[pos[:, 0].to(dev), pos[:, 1].to(dev), pos[:, 2].to(dev)] = rows_d

m = B.BaseBEVBackbone(CFG, 64).eval().to(dev)
def timed(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))
res = {"frames": nb, "config": name, "pillars": int(len(pos))}
with torch.inference_mode():
    res["ours_rows_map_ms"] = timed(lambda: m({"pillar_features": rows_d, "bev_index_map": idx_d}))
    res["ours_canvas_ms"] = timed(lambda: m({"spatial_features": canvas}))
    # per layer (each launch timed alone)
    m({"pillar_features": rows_d, "bev_index_map": idx_d})
    blocks, de = m._plan
    per = []
    x, hh, ww = None, h, w
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    for i, layers in enumerate(blocks):
        for l in layers:
            oh, ow = m._out_hw(hh, ww, l.desc)
            y = torch.empty((nb, oh, ow, l.desc["c_out"]), device=dev)
            f = (lambda l=l, y=y, x=x, hh=hh, ww=ww: B.conv_forward(l, y, l.desc["c_out"], 0, False, nb, hh, ww, x_nhwc=x,
                 rows=rows_d if x is None else None, cell_row=idx_d if x is None else None, error=err))
            t = timed(f, n=5, warm=2)
            d = l.desc
            taps = d["k"] * d["k"]
            fl = 2.0 * nb * oh * ow * d["c_out"] * d["c_in"] * taps
            per.append(dict(layer=f"block{i} {d['c_in']}->{d['c_out']} k{d['k']} s{d['stride']} @{oh}x{ow}", ms=round(t, 4),
                            tflops=round(fl / t / 1e9, 1)))
            x, hh, ww = y, oh, ow
        if de:
            l = de[i]; d = l.desc
            oh, ow = m._out_hw(hh, ww, d)
            out = torch.empty((nb, 384, oh, ow), device=dev)
            f = (lambda l=l, out=out, x=x, hh=hh, ww=ww: B.conv_forward(l, out, 384, 0, True, nb, hh, ww, x_nhwc=x, round_out=False, error=err))
            t = timed(f, n=5, warm=2)
            fl = 2.0 * nb * hh * ww * d["c_out"] * d["c_in"] * (d["up"] ** 2 if d["up"] > 1 else d["k"] ** 2) / (1 if d["up"] > 1 else d["stride"] ** 2)
            per.append(dict(layer=f"deblock{i} {d['c_in']}->{d['c_out']} k{d['k']} s{d['stride']} up{d['up']}", ms=round(t, 4), tflops=round(fl / t / 1e9, 1)))
    res["layers"] = per
    res["sum_layers_ms"] = round(sum(p["ms"] for p in per), 4)
    assert int(err.item()) == 0

# the reference's own module, eager on this GPU
try:
    from oracle import ref_loader as R
    ref = R.load_bev_backbone()(R.AttrDict(CFG), 64).eval().to(dev)
    with torch.inference_mode():
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cudnn.benchmark = True
        res["reference_eager_tf32_ms"] = timed(lambda: ref({"spatial_features": canvas}), n=5, warm=3)
        torch.backends.cudnn.allow_tf32 = False
        res["reference_eager_fp32_ms"] = timed(lambda: ref({"spatial_features": canvas}), n=5, warm=3)
        torch.backends.cudnn.allow_tf32 = True
        refc = ref.to(memory_format=torch.channels_last)
        cl = canvas.contiguous(memory_format=torch.channels_last)
        res["reference_eager_tf32_channels_last_ms"] = timed(lambda: refc({"spatial_features": cl}), n=5, warm=3)
except Exception as e:  # noqa: BLE001
    res["reference_error"] = repr(e)
fl_total = sum(2.0 * 1 for _ in [0])
print(json.dumps(res, indent=1))
