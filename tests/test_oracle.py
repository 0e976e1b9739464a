"""CPU tests: the oracle restatement against the golden vectors produced by the real reference modules
(tests/golden/make_golden.py), against an independent pure-Python transcription, and -- when the reference tree
is present (build container) -- against the live reference."""
import numpy as np
import pytest
import torch

from conftest import load_golden, vfe_golden_names
from lidar_vision_vqa_b200 import synth


def _sd_t(g):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in g["state_dict"].items()}


@pytest.mark.parametrize("name", vfe_golden_names())
def test_pillar_vfe_restatement_matches_reference_golden(oracle, name):
    g = load_golden(name)
    out = oracle.pillar_vfe(g["voxels"], g["voxel_num_points"], g["voxel_coords"], _sd_t(g), g["voxel_size"],
                            g["range"], use_norm=bool(g["use_norm"]), with_distance=bool(g["with_distance"]),
                            use_absolute_xyz=bool(g["use_abs"])).numpy()
    ref = g["out.pillar_features"]
    if ref.ndim == 1:  # the reference squeeze()s M == 1 (pillar_vfe.py:121)
        assert out.shape[0] == 1
        out = out[0]
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", [n for n in vfe_golden_names() if n != "vfe_c5_m1"])
def test_scatter_restatement_matches_reference_golden(oracle, name):
    g = load_golden(name)
    nx, ny = int(g["grid_size"][0]), int(g["grid_size"][1])
    bev = oracle.scatter_bev(g["out.pillar_features"], g["voxel_coords"], nx, ny)
    np.testing.assert_array_equal(bev, g["out.spatial_features"])


def test_scatter3d_restatement_matches_reference_golden(oracle):
    """PointPillarScatter3d (pointpillar_scatter.py:40-73), nz = 2, fp32 coords as after load_data_to_gpu."""
    g = load_golden("scatter3d_nz2")
    nx, ny, nz = (int(v) for v in g["input_shape"])
    bev = oracle.scatter_bev(g["pillar_features"], g["voxel_coords"], nx, ny, nz=nz)
    assert bev.shape == g["out.spatial_features"].shape == (2, int(g["num_bev_features"]), ny, nx)
    np.testing.assert_array_equal(bev, g["out.spatial_features"])


@pytest.mark.parametrize("name", vfe_golden_names())
def test_golden_voxel_inputs_are_reproducible(oracle, name):
    """The voxel tensors fed to the reference came from the C voxeliser; regenerate them from the stored points."""
    g = load_golden(name)
    v = oracle.voxelize_batch(g["points"], g["frame_offsets"], g["range"], g["voxel_size"], int(g["max_points"]),
                              int(g["max_voxels"]))
    np.testing.assert_array_equal(v["voxels"], g["voxels"])
    np.testing.assert_array_equal(v["coords"], g["voxel_coords"].astype(np.int32))
    np.testing.assert_array_equal(v["num_points"], g["voxel_num_points"].astype(np.int32))
    np.testing.assert_array_equal(v["point_voxel"], g["point_voxel"])
    np.testing.assert_array_equal(v["point_slot"], g["point_slot"])


@pytest.mark.parametrize("max_points,max_voxels", [(32, 100000), (4, 100000), (8, 150), (1, 1)])
def test_c_voxeliser_equals_python_transcription(oracle, max_points, max_voxels):
    pts = synth.make_sweep(3, synth.NUSCENES_32, 5)[:3000]
    rng, vs = (-20.0, -20.0, -5.0, 20.0, 20.0, 3.0), (0.5, 0.5, 8.0)
    a = oracle.voxelize_hard(pts, rng, vs, max_points, max_voxels)
    b = oracle.voxelize_hard_py(pts, rng, vs, max_points, max_voxels)
    for k in ("voxels", "coords", "num_points", "point_voxel", "point_slot"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


def test_voxeliser_boundaries(oracle):
    rng, vs = (0.0, 0.0, 0.0, 4.0, 4.0, 2.0), (1.0, 1.0, 2.0)
    pts = np.array([
        [0.0, 0.0, 0.0, 1, 0],        # exactly on the lower corner -> cell (0,0,0)
        [4.0, 1.0, 1.0, 2, 0],        # x == range max -> rejected (mask_points_by_range would keep it)
        [3.9999998, 1.0, 1.0, 3, 0],  # just inside
        [1.0, 1.0, 2.0, 4, 0],        # z == range max -> rejected
        [-1e-7, 1.0, 1.0, 5, 0],      # tiny negative -> floor = -1 -> rejected
        [np.nan, 1.0, 1.0, 6, 0],     # non-finite -> rejected
        [1.0, np.inf, 1.0, 7, 0],
        [0.5, 0.5, 0.5, 8, 0],        # second point of cell (0,0,0)
    ], np.float32)
    o = oracle.voxelize_hard(pts, rng, vs, 4, 10)
    np.testing.assert_array_equal(o["point_voxel"], [0, -1, 1, -1, -1, -1, -1, 0])
    np.testing.assert_array_equal(o["point_slot"], [0, -1, 0, -1, -1, -1, -1, 1])
    np.testing.assert_array_equal(o["coords"], [[0, 0, 0], [0, 1, 3]])
    np.testing.assert_array_equal(o["num_points"], [2, 1])


def test_voxeliser_empty_and_all_rejected(oracle):
    rng, vs = (0.0, 0.0, 0.0, 4.0, 4.0, 2.0), (1.0, 1.0, 2.0)
    o = oracle.voxelize_hard(np.zeros((0, 5), np.float32), rng, vs, 4, 10)
    assert o["coords"].shape == (0, 3) and o["voxels"].shape == (0, 4, 5)
    o = oracle.voxelize_hard(np.full((7, 5), 100.0, np.float32), rng, vs, 4, 10)
    assert o["coords"].shape[0] == 0 and (o["point_voxel"] == -1).all()


def test_pillar_set_matches_reference_dynamic_vfe_golden(oracle):
    """Set-level pin of the quantisation: the hard voxeliser's pillar set and (capped) counts must equal what the
    reference's DynamicPillarVFE produced (z pre-filtered to the range, max_voxels not binding)."""
    g = load_golden("dyn_c5")
    pb = g["points_b"]
    ref_coords = g["out.voxel_coords"]  # (b, 0, iy, ix), sorted by (b, ix, iy)
    # restated dynamic quantisation reproduces the reference's coordinate list exactly
    dc, dcnt = oracle.dynamic_pillar_sets(pb, g["range"], g["voxel_size"])
    np.testing.assert_array_equal(dc, ref_coords)
    # hard voxeliser per frame
    nb = int(pb[:, 0].max()) + 1
    offs = np.searchsorted(pb[:, 0], np.arange(nb + 1)).astype(np.int32)
    v = oracle.voxelize_batch(np.ascontiguousarray(pb[:, 1:]), offs, g["range"], g["voxel_size"], 1000, 100000,
                              want_voxels=False)
    nx, ny = int(g["grid_size"][0]), int(g["grid_size"][1])
    key_h = v["coords"][:, 0].astype(np.int64) * nx * ny + v["coords"][:, 3] * ny + v["coords"][:, 2]
    key_d = dc[:, 0].astype(np.int64) * nx * ny + dc[:, 3] * ny + dc[:, 2]
    order = np.argsort(key_h)
    np.testing.assert_array_equal(key_h[order], key_d)
    np.testing.assert_array_equal(v["num_points"][order], dcnt)


DYN_GOLDENS = ["dyn_c5", "dyn_c5_2layer_zout", "dyn2d_c5_f32", "dyn_c4_dist_noabs"]


@pytest.mark.parametrize("name", DYN_GOLDENS)
def test_dynamic_vfe_restatement_matches_reference_golden(name, oracle):
    """oracle.dynamic_pillar_vfe vs the reference's DynamicPillarVFE / DynamicPillarVFESimple2D outputs: coordinates and
    row order bit-exact (z outliers included: the dynamic variants never range-check z), features to 1e-5."""
    import torch

    g = load_golden(name)
    sd = {k: torch.from_numpy(v) for k, v in g["state_dict"].items()}
    feats, coords, _ = oracle.dynamic_pillar_vfe(g["points_b"], sd, g["voxel_size"], g["range"],
                                                 with_distance=bool(g["with_distance"]),
                                                 use_absolute_xyz=bool(g["use_abs"]), simple2d=bool(g["simple2d"]))
    np.testing.assert_array_equal(coords, g["out.voxel_coords"])
    assert feats.shape == g["out.pillar_features"].shape
    np.testing.assert_allclose(feats.numpy(), g["out.pillar_features"], rtol=1e-5, atol=1e-5)


def test_live_reference_if_present(oracle):
    """In the build container re-run the reference module itself on a fresh random case (not a stored golden)."""
    from oracle import ref_loader

    if not ref_loader.reference_available():
        pytest.skip("reference tree not present (expected on the GPU box)")
    ref = ref_loader.load_reference()
    rng, vs = (-10.0, -10.0, -5.0, 10.0, 10.0, 3.0), (0.25, 0.25, 8.0)
    pts, offs = synth.make_batch(2, synth.NUSCENES_32, 5, seed0=123)
    sel = np.hypot(pts[:, 0], pts[:, 1]) < 14
    pts = pts[sel][:6000]
    offs = np.array([0, 3000, len(pts)], np.int32)
    v = oracle.voxelize_batch(pts, offs, rng, vs, 10, 5000)
    voxels, npts, coords = oracle.collate_voxels(v)
    grid = oracle.grid_size_of(rng, vs)
    cfg = ref.AttrDict(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64])
    m = ref.PillarVFE(model_cfg=cfg, num_point_features=5, voxel_size=list(vs),
                      point_cloud_range=np.asarray(rng, np.float32), grid_size=grid)
    sd = oracle.random_pfn_params(11, [64], True, seed=5)
    m.load_state_dict(sd)
    m.eval()
    with torch.inference_mode():
        bd = m({"voxels": torch.from_numpy(voxels), "voxel_num_points": torch.from_numpy(npts),
                "voxel_coords": torch.from_numpy(coords)})
        bd = ref.PointPillarScatter(ref.AttrDict(NUM_BEV_FEATURES=64), grid)(bd)
    mine = oracle.pillar_vfe(voxels, npts, coords, sd, vs, rng).numpy()
    np.testing.assert_allclose(mine, bd["pillar_features"].numpy(), rtol=1e-5, atol=1e-5)
    bev = oracle.scatter_bev(bd["pillar_features"].numpy(), coords, int(grid[0]), int(grid[1]))
    np.testing.assert_array_equal(bev, bd["spatial_features"].numpy())
