"""Phase stamps of the streaming feature kernel on a two-layer [64,64] stack (pillars_encode_stack).
Usage: python profiles/scripts/phase_times_stack.py [workload]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import lidar_vision_vqa_b200 as L
from lidar_vision_vqa_b200 import _native, ops, synth
from oracle import pillar_oracle as po

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2_nuscenes32_b16_pillar0.2_bev512"
dev = torch.device("cuda:0")
model, gc, nb = synth.WORKLOADS[wl]
grid = L.GridSpec.from_range(gc.point_cloud_range, gc.voxel_size, gc.max_points_per_voxel, gc.max_voxels)
pts, offs = synth.make_batch(nb, model, 5)
p, o = torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev)
filters = [64, 64]
sd = po.random_pfn_params(11, filters, True, seed=0)
layers = [(torch.as_tensor(sd[f"pfn_layers.{i}.linear.weight"]),
           tuple(torch.as_tensor(sd[f"pfn_layers.{i}.norm.{k}"]) for k in ("weight", "bias", "running_mean", "running_var")) + (1e-3,), None)
          for i in range(len(filters))]
st = ops.fold_pfn_stack(layers, c_point=5, use_absolute_xyz=True, with_distance=False, voxel_size=grid.voxel_size,
                        point_cloud_range=grid.point_cloud_range, device=dev)
lib = _native.load()
for _ in range(3):
    ops.encode_stack(p, o, grid, st, with_bev=False)
torch.cuda.synchronize()
NAMES = {18: "walk first CTA in", 20: "walk first past wait", 22: "walk FIRST warp chunks done", 21: "walk LAST warp chunks done",
         24: "walk first warp out", 23: "walk last warp out"}
rows = []
for rep in range(7):
    buf = torch.zeros(32, dtype=torch.int64, device=dev)
    lib.pillars_set_debug_times(buf.data_ptr())
    ops.encode_stack(p, o, grid, st, with_bev=False)
    torch.cuda.synchronize()
    lib.pillars_set_debug_times(None)
    v = buf.cpu().numpy().view(np.uint64)
    t = {k: int((~v[k]) if k % 2 == 0 else v[k]) for k in NAMES if v[k] != 0}
    t0 = t[18]
    rows.append({k: (x - t0) / 1e3 for k, x in t.items()})
print(f"{wl} NUM_FILTERS {filters}: microseconds from the walk kernel's first CTA (median of {len(rows)} calls)")
for k in NAMES:
    print(f"  {NAMES[k]:40s} {np.median([r[k] for r in rows if k in r]):8.2f}")
