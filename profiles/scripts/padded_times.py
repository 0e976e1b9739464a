"""PillarVFE.forward on the reference's padded voxels (pillars_pfn_dense): folded kernel k_pfn_padded vs the faithful
11-feature kernel k_pfn_dense, and the scatter from coordinates (pillars_scatter_bev).  Usage: python profiles/scripts/padded_times.py"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import lidar_vision_vqa_b200 as L
from lidar_vision_vqa_b200 import _native, ops, synth
from oracle import pillar_oracle as po
wl = "cfg2_nuscenes32_b16_pillar0.2_bev512"
dev = torch.device("cuda:0")
model, gc, nb = synth.WORKLOADS[wl]
pts, offs = synth.make_batch(nb, model, 5)
v = po.voxelize_batch(pts, offs, gc.point_cloud_range, gc.voxel_size, gc.max_points_per_voxel, gc.max_voxels)
vox, npts, crd = (torch.from_numpy(v[k]).to(dev) for k in ("voxels", "num_points", "coords"))
sd = po.random_pfn_params(11, [64], True, seed=0)
pfn = ops.fold_pfn(sd["pfn_layers.0.linear.weight"], (sd["pfn_layers.0.norm.weight"], sd["pfn_layers.0.norm.bias"],
                   sd["pfn_layers.0.norm.running_mean"], sd["pfn_layers.0.norm.running_var"], 1e-3), None, c_point=5,
                   use_absolute_xyz=True, with_distance=False, voxel_size=gc.voxel_size,
                   point_cloud_range=gc.point_cloud_range, device=dev)
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts)) * 1e3
m = vox.shape[0]
bytes_ = vox.numel() * 4 + m * 20 + m * 256
t_new = timed(lambda: ops.pfn_dense(vox, npts, crd, pfn, gc.voxel_size))
f_new = ops.pfn_dense(vox, npts, crd, pfn, gc.voxel_size)
ops.force_generic_features(True)
t_old = timed(lambda: ops.pfn_dense(vox, npts, crd, pfn, gc.voxel_size))
f_old = ops.pfn_dense(vox, npts, crd, pfn, gc.voxel_size)
ops.force_generic_features(False)
nx, ny, _ = gc.grid_size
bev = torch.empty((nb, 64, ny, nx), device=dev)
t_sc = timed(lambda: ops.scatter_bev(f_new, crd, nb, nx, ny, out=bev))
print(f"{wl}: M = {m} pillars x P = {vox.shape[1]} x C = 5 ({bytes_ / 1e6:.0f} MB in + out)")
print(f"  k_pfn_padded (folded)   {t_new:7.1f} us = {bytes_ / t_new / 1e3:6.0f} GB/s")
print(f"  k_pfn_dense  (faithful) {t_old:7.1f} us = {bytes_ / t_old / 1e3:6.0f} GB/s;  max |new - old| = {(f_new - f_old).abs().max().item():.2e}")
print(f"  pillars_scatter_bev (index map from coordinates + canvas) {t_sc:7.1f} us")
