"""CUDA-event timing of the non-headline variants on cfg2-sized input (16 sweeps): DynPillarVFE [64] / [64,64],
DynamicPillarVFESimple2D [32], hard two-layer fused path, hard PillarVFE on padded voxels (1 and 2 layers).
Usage: python profiles/time_variants.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lidar_vision_vqa_b200 as L  # noqa: E402
from lidar_vision_vqa_b200 import ops, synth  # noqa: E402
from oracle import pillar_oracle as po  # noqa: E402  (weights generator only)


class C(dict):
    __getattr__ = dict.__getitem__


dev = torch.device("cuda:0")
model, gc, nb = synth.WORKLOADS["cfg2_nuscenes32_b16_pillar0.2_bev512"]
grid = L.GridSpec.from_range(gc.point_cloud_range, gc.voxel_size, 32, 30000)
pts, offs = synth.make_batch(nb, model, 5)
pb = torch.from_numpy(synth.to_pcdet_points(pts, offs)).to(dev)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return float(np.median(ts))


def build(cls, filters, c_in, **extra):
    cfg = C(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=filters, **extra)
    vfe = cls(model_cfg=cfg, num_point_features=5, voxel_size=list(gc.voxel_size),
              point_cloud_range=np.asarray(gc.point_cloud_range, np.float32), grid_size=np.asarray(grid.grid_size))
    vfe.load_state_dict(po.random_pfn_params(c_in, filters, True, seed=0))
    return vfe.eval().to(dev)


rows = []
for name, cls, filters, c_in in [("DynPillarVFE [64]", L.DynamicPillarVFE, [64], 11),
                                 ("DynPillarVFE [64,64]", L.DynamicPillarVFE, [64, 64], 11),
                                 ("DynamicPillarVFESimple2D [32]", L.DynamicPillarVFESimple2D, [32], 8)]:
    vfe = build(cls, filters, c_in)
    bd = vfe({"points": pb, "batch_size": nb})
    m = bd["pillar_features"].shape[0]
    rows.append((name, timed(lambda: vfe({"points": pb, "batch_size": nb})), m))
for name, filters in [("PillarVFEFromPoints [64] (streaming kernel) + fused scatter", [64]),
                      ("PillarVFEFromPoints [64,64] (general kernel) + fused scatter", [64, 64])]:
    vfe = build(L.PillarVFEFromPoints, filters, 11, MAX_POINTS_PER_VOXEL=32, MAX_NUMBER_OF_VOXELS=30000, FUSE_SCATTER=True)
    bd = vfe({"points": pb, "batch_size": nb})
    rows.append((name, timed(lambda: vfe({"points": pb, "batch_size": nb})), bd["pillar_features"].shape[0]))
# padded-voxel input (the reference's own format)
v = ops.voxelize(torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev), grid, want_voxels=True)
m = int(v["pillar_count"][-1].item())
vox, npts, crd = v["voxels"][:m], v["voxel_num_points"][:m].float(), v["voxel_coords"][:m].float()
for name, filters in [("PillarVFE [64] on padded voxels [M,32,5]", [64]), ("PillarVFE [64,64] on padded voxels", [64, 64])]:
    vfe = build(L.PillarVFE, filters, 11)
    rows.append((name, timed(lambda: vfe({"voxels": vox, "voxel_num_points": npts, "voxel_coords": crd})), m))
sc = L.PointPillarScatter(model_cfg=C(NUM_BEV_FEATURES=64), grid_size=np.asarray(grid.grid_size))
feats = torch.randn(m, 64, device=dev)
rows.append(("PointPillarScatter module (index map + canvas)",
             timed(lambda: sc({"pillar_features": feats, "voxel_coords": crd, "batch_size": nb})), m))
print(f"cfg2: {len(pts)} points, {nb} frames; module-level call incl. the host sync for the pillar count")
for name, us, m in rows:
    print(f"  {name:62s} {us:8.1f} us   M={m}")
