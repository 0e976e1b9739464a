"""Nanosecond phase stamps of the grouping / feature kernels of one pillars_encode_bev call (pillars_set_debug_times).
Usage: python profiles/scripts/phase_times.py [workload] [hash|dense]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import lidar_vision_vqa_b200 as L
from lidar_vision_vqa_b200 import _native, ops, synth
from oracle import pillar_oracle as po

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2_nuscenes32_b16_pillar0.2_bev512"
if len(sys.argv) > 2:
    ops.set_grouping(sys.argv[2])
dev = torch.device("cuda:0")
model, gc, nb = synth.WORKLOADS[wl]
grid = L.GridSpec.from_range(gc.point_cloud_range, gc.voxel_size, gc.max_points_per_voxel, gc.max_voxels)
pts, offs = synth.make_batch(nb, model, 5)
sd = po.random_pfn_params(11, [64], True, seed=0)
pfn = ops.fold_pfn(sd["pfn_layers.0.linear.weight"], (sd["pfn_layers.0.norm.weight"], sd["pfn_layers.0.norm.bias"],
                   sd["pfn_layers.0.norm.running_mean"], sd["pfn_layers.0.norm.running_var"], 1e-3), None, c_point=5,
                   use_absolute_xyz=True, with_distance=False, voxel_size=grid.voxel_size,
                   point_cloud_range=grid.point_cloud_range, device=dev)
p, o = torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev)
bufs = ops.EncodeBuffers(len(pts), nb, grid, 64, dev)
lib = _native.load()
for _ in range(5):
    ops.encode_bev(p, o, grid, pfn, buffers=bufs)
torch.cuda.synchronize()
NAMES = {0: "fill first CTA in", 1: "fill last CTA out", 2: "insert first CTA in", 3: "insert last tile quantised",
         4: "insert first past wait (fill done)", 5: "insert last CTA out", 6: "scan first CTA in",
         7: "scan last CTA past wait", 8: "scan first CTA past wait (insert done)", 9: "scan last tile scanned",
         11: "scan last look-back resolved", 13: "scan last CTA out", 14: "place first CTA in", 16: "place first past wait",
         17: "place last CTA out", 18: "walk first CTA in", 20: "walk first past wait", 21: "walk last warp chunks done", 22: "walk FIRST warp chunks done",
         23: "walk last warp out", 24: "walk FIRST warp out", 25: "insert last CTA in", 27: "insert last tile in shared memory"}
rows = []
for rep in range(7):
    buf = torch.zeros(32, dtype=torch.int64, device=dev)
    lib.pillars_set_debug_times(buf.data_ptr())
    ops.encode_bev(p, o, grid, pfn, buffers=bufs)
    torch.cuda.synchronize()
    lib.pillars_set_debug_times(None)
    v = buf.cpu().numpy().view(np.uint64)
    t = {k: int((~v[k]) if k % 2 == 0 else v[k]) for k in NAMES if v[k] != 0}
    t0 = min(t.values())
    rows.append({k: (x - t0) / 1e3 for k, x in t.items()})
keys = sorted(rows[0], key=lambda k: np.median([r[k] for r in rows if k in r]))
print(f"{wl}: microseconds from the first stamp (median of {len(rows)} calls)")
for k in keys:
    print(f"  {NAMES[k]:40s} {np.median([r[k] for r in rows if k in r]):8.2f}")
