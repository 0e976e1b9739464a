"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules (imported from /root/reference).

Run in the build container only:   python tests/golden/make_golden.py
The reference tree does not travel to the GPU box, the .npz files do.  Every file stores the exact inputs fed
to the reference module and the outputs it returned, plus the weights (reference state_dict keys).

Reference modules exercised (src/lidar-encoder/pcdet/...):
  models/backbones_3d/vfe/pillar_vfe.py:52-123            PillarVFE(+PFNLayer)      -> pillar_features
  models/backbones_2d/map_to_bev/pointpillar_scatter.py:5-37  PointPillarScatter     -> spatial_features
  models/backbones_3d/vfe/dynamic_pillar_vfe.py:49-142    DynamicPillarVFE          -> voxel_coords, pillar_features
The voxel tensors fed to PillarVFE come from oracle/voxelize_ref.c (spconv itself is absent; see that file).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pillar_oracle as po  # noqa: E402
from oracle import ref_loader  # noqa: E402
from lidar_vision_vqa_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
RANGE = (-12.8, -12.8, -5.0, 12.8, 12.8, 3.0)
VOXEL = (0.4, 0.4, 8.0)


def small_batch(batch, num_features, n_keep, seed0):
    """Near-field crop of synthetic sweeps: few thousand points, dense rings => pillars above the cap."""
    frames = []
    for b in range(batch):
        p = synth.make_sweep(seed0 + b, synth.NUSCENES_32, num_features)
        r = np.hypot(p[:, 0], p[:, 1])
        p = p[r < 19.0][:n_keep]  # some points fall outside RANGE on purpose
        frames.append(p)
    offs = np.zeros(batch + 1, np.int32)
    offs[1:] = np.cumsum([len(f) for f in frames])
    return np.concatenate(frames, 0), offs


def run_case(name, *, batch, c, p_max, max_voxels, n_keep, seed, use_norm=True, with_distance=False,
             use_abs=True, num_filters=(64,), coords_float=True, scatter=True):
    ref = ref_loader.load_reference()
    pts, offs = small_batch(batch, c, n_keep, seed)
    vox = po.voxelize_batch(pts, offs, RANGE, VOXEL, p_max, max_voxels)
    voxels, npts, coords = po.collate_voxels(vox, as_float=coords_float)
    grid = po.grid_size_of(RANGE, VOXEL)

    cfg = ref.AttrDict(USE_NORM=use_norm, WITH_DISTANCE=with_distance, USE_ABSLOTE_XYZ=use_abs,
                       NUM_FILTERS=list(num_filters))
    vfe = ref.PillarVFE(model_cfg=cfg, num_point_features=c, voxel_size=list(VOXEL),
                        point_cloud_range=np.asarray(RANGE, np.float32), grid_size=grid,
                        depth_downsample_factor=None)
    c_in = vfe.pfn_layers[0].linear.in_features
    sd = po.random_pfn_params(c_in, num_filters, use_norm, seed=seed + 100)
    vfe.load_state_dict(sd, strict=True)
    vfe.eval()
    bd = {"voxels": torch.from_numpy(voxels), "voxel_num_points": torch.from_numpy(npts),
          "voxel_coords": torch.from_numpy(coords), "batch_size": batch}
    with torch.inference_mode():
        bd = vfe(bd)
        feats = bd["pillar_features"].clone()
        out = {"pillar_features": feats.numpy()}
        if scatter and feats.dim() == 2:
            sc = ref.PointPillarScatter(model_cfg=ref.AttrDict(NUM_BEV_FEATURES=int(num_filters[-1])),
                                        grid_size=grid)
            bd = sc(bd)
            out["spatial_features"] = bd["spatial_features"].numpy()
    save = {
        "points": pts, "frame_offsets": offs,
        "voxels": voxels, "voxel_num_points": npts, "voxel_coords": coords,
        "point_voxel": vox["point_voxel"], "point_slot": vox["point_slot"],
        "range": np.asarray(RANGE, np.float32), "voxel_size": np.asarray(VOXEL, np.float32),
        "grid_size": grid, "max_points": np.int32(p_max), "max_voxels": np.int32(max_voxels),
        "use_norm": np.bool_(use_norm), "with_distance": np.bool_(with_distance), "use_abs": np.bool_(use_abs),
        "num_filters": np.asarray(num_filters, np.int32),
    }
    for k, v in sd.items():
        save["sd." + k] = v.numpy()
    for k, v in out.items():
        save["out." + k] = v
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **save)
    print(f"{name}: N={len(pts)} M={voxels.shape[0]} max_n={int(npts.max())} feats={tuple(feats.shape)} "
          f"-> {os.path.getsize(path) / 1024:.0f} KiB")


def run_dynamic(name, *, batch, c, n_keep, seed, num_filters=(64,), keep_z_outliers=False, simple2d=False,
                with_distance=False, use_abs=True):
    ref = ref_loader.load_reference()
    pb_full = synth.to_pcdet_points(*small_batch(batch, c, n_keep, seed))
    if keep_z_outliers:
        # DynamicPillarVFE never range-checks z (dynamic_pillar_vfe.py:93-96): push some points outside [zmin, zmax)
        pb = pb_full.copy()
        pb[::7, 3] += 9.0
        pb[::11, 3] -= 9.0
    else:
        # keep z inside the range so the pillar set is comparable with the hard voxeliser
        pb = pb_full[(pb_full[:, 3] >= RANGE[2]) & (pb_full[:, 3] < RANGE[5])]
    grid = po.grid_size_of(RANGE, VOXEL)
    cfg = ref.AttrDict(USE_NORM=True, WITH_DISTANCE=with_distance, USE_ABSLOTE_XYZ=use_abs, NUM_FILTERS=list(num_filters))
    cls = ref.DynamicPillarVFESimple2D if simple2d else ref.DynamicPillarVFE
    with ref_loader.cuda_is_identity():
        vfe = cls(model_cfg=cfg, num_point_features=c, voxel_size=list(VOXEL), grid_size=grid,
                  point_cloud_range=np.asarray(RANGE, np.float32))
    sd = po.random_pfn_params(vfe.pfn_layers[0].linear.in_features, num_filters, True, seed=seed + 100)
    vfe.load_state_dict(sd, strict=True)
    vfe.eval()
    with torch.inference_mode():
        bd = vfe({"points": torch.from_numpy(pb), "batch_size": batch})
    ckey = "pillar_coords" if simple2d else "voxel_coords"
    save = {"points_b": pb, "range": np.asarray(RANGE, np.float32), "voxel_size": np.asarray(VOXEL, np.float32),
            "grid_size": grid, "out.voxel_coords": bd[ckey].numpy(),
            "out.pillar_features": bd["pillar_features"].numpy(),
            "num_filters": np.asarray(num_filters, np.int32), "simple2d": np.bool_(simple2d),
            "with_distance": np.bool_(with_distance), "use_abs": np.bool_(use_abs), "batch": np.int32(batch),
            "c": np.int32(c)}
    for k, v in sd.items():
        save["sd." + k] = v.numpy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **save)
    print(f"{name}: N={len(pb)} M={save['out.voxel_coords'].shape[0]} -> {os.path.getsize(path) / 1024:.0f} KiB")


def run_scatter3d(name, *, batch, nx, ny, nz, c_before, m_per_frame, seed):
    """PointPillarScatter3d on random unique cells (the DSVT-pillar configs use it with nz > 1)."""
    ref = ref_loader.load_reference()
    rng = np.random.default_rng(seed)
    coords, feats = [], []
    for b in range(batch):
        cells = rng.choice(nx * ny * nz, m_per_frame, replace=False)
        z, rem = cells // (nx * ny), cells % (nx * ny)
        coords.append(np.stack([np.full(m_per_frame, b), z, rem // nx, rem % nx], 1))
        feats.append(rng.standard_normal((m_per_frame, c_before)).astype(np.float32))
    coords = np.concatenate(coords).astype(np.float32)  # fp32 like load_data_to_gpu
    feats = np.concatenate(feats)
    sc = ref.PointPillarScatter3d(ref.AttrDict(INPUT_SHAPE=[nx, ny, nz], NUM_BEV_FEATURES=c_before * nz), grid_size=None)
    with torch.inference_mode():
        out = sc({"pillar_features": torch.from_numpy(feats), "voxel_coords": torch.from_numpy(coords)})
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, pillar_features=feats, voxel_coords=coords, input_shape=np.array([nx, ny, nz]),
                        num_bev_features=np.int32(c_before * nz), **{"out.spatial_features": out["spatial_features"].numpy()})
    print(f"{name}: out {tuple(out['spatial_features'].shape)} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(1)
    run_case("vfe_c5_p32_f32coords", batch=2, c=5, p_max=32, max_voxels=4000, n_keep=2000, seed=0)
    run_case("vfe_c4_p20_i32coords", batch=3, c=4, p_max=20, max_voxels=4000, n_keep=1200, seed=10,
             coords_float=False)
    run_case("vfe_c5_p8_capbinds_maxvox200", batch=2, c=5, p_max=8, max_voxels=200, n_keep=3000, seed=20)
    run_case("vfe_c5_dist_noabs", batch=1, c=5, p_max=16, max_voxels=1000, n_keep=1200, seed=30,
             with_distance=True, use_abs=False)
    run_case("vfe_c5_nonorm", batch=1, c=5, p_max=16, max_voxels=1000, n_keep=1200, seed=40, use_norm=False)
    run_case("vfe_c5_m1", batch=1, c=5, p_max=8, max_voxels=1, n_keep=200, seed=50, scatter=False)
    run_case("vfe_c5_2layer", batch=1, c=5, p_max=12, max_voxels=1000, n_keep=1200, seed=60, num_filters=(64, 64))
    run_dynamic("dyn_c5", batch=2, c=5, n_keep=2000, seed=70)
    run_dynamic("dyn_c5_2layer_zout", batch=2, c=5, n_keep=1500, seed=71, num_filters=(64, 64), keep_z_outliers=True)
    run_dynamic("dyn2d_c5_f32", batch=2, c=5, n_keep=1500, seed=72, num_filters=(32,), simple2d=True,
                keep_z_outliers=True)
    run_dynamic("dyn_c4_dist_noabs", batch=1, c=4, n_keep=1200, seed=73, num_filters=(64,), with_distance=True,
                use_abs=False)
    run_scatter3d("scatter3d_nz2", batch=2, nx=40, ny=36, nz=2, c_before=32, m_per_frame=300, seed=80)
