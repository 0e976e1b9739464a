"""GPU parity tests (run with ``-m gpu`` on a B200): the CUDA path, called through the C ABI, against the CPU oracle
on identical seeded inputs and against the golden vectors produced by the reference modules.

Bars (BASELINE.json north_star): voxel coordinates, pillar ids and point->pillar membership BIT-EXACT;
pillar features and BEV tokens within rtol 1e-3 (fp32).  ``FEAT_RTOL``/``FEAT_ATOL`` below are that tolerance; the
observed error is ~1e-6, so the tests also assert a much tighter "expected" bound to catch regressions early.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, vfe_golden_names
from lidar_vision_vqa_b200 import synth

pytestmark = pytest.mark.gpu


class C(dict):
    __getattr__ = dict.__getitem__

FEAT_RTOL = 1e-3   # the north_star tolerance
FEAT_ATOL = 1e-5   # exact zeros in empty cells are compared exactly elsewhere
TIGHT_RTOL, TIGHT_ATOL = 2e-5, 2e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def L():
    import lidar_vision_vqa_b200 as pkg
    from lidar_vision_vqa_b200 import ops

    pkg.ops = ops
    return pkg


def _sd_t(g):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in g["state_dict"].items()}


def _trim(res):
    m = int(res["pillar_count"][-1].item())
    out = {}
    for k, v in res.items():
        if k in ("voxel_coords", "voxel_num_points", "voxels", "pillar_features"):
            out[k] = v[:m].cpu().numpy()
        elif k != "bev":
            out[k] = v.cpu().numpy()
    out["m"] = m
    return out


GROUP_CASES = {
    # name: (frames, sweep model, C, range, voxel, P, max_voxels, points kept per frame)
    "small_p32": (2, synth.NUSCENES_32, 5, (-12.8, -12.8, -5, 12.8, 12.8, 3), (0.4, 0.4, 8), 32, 4000, 3000),
    "cap_binds_p4": (3, synth.NUSCENES_32, 5, (-12.8, -12.8, -5, 12.8, 12.8, 3), (0.4, 0.4, 8), 4, 4000, 4000),
    "maxvox_binds": (3, synth.NUSCENES_32, 4, (-12.8, -12.8, -5, 12.8, 12.8, 3), (0.4, 0.4, 8), 8, 150, 4000),
    "p1_maxvox1": (2, synth.NUSCENES_32, 5, (-12.8, -12.8, -5, 12.8, 12.8, 3), (0.4, 0.4, 8), 1, 1, 500),
    "cfg1_full_sweep": (1, synth.NUSCENES_32, 5, (-51.2, -51.2, -5, 51.2, 51.2, 3), (0.2, 0.2, 8), 32, 30000, None),
    "cfg1_p20": (2, synth.NUSCENES_32, 5, (-51.2, -51.2, -5, 51.2, 51.2, 3), (0.2, 0.2, 8), 20, 30000, None),
    "tenSweep_maxvox30000_binds": (1, synth.NUSCENES_10SWEEP, 5, (-51.2, -51.2, -5, 51.2, 51.2, 3), (0.2, 0.2, 8), 32,
                                   30000, None),
    "waymo_0.1m": (1, synth.WAYMO_64, 5, (-51.2, -51.2, -2, 51.2, 51.2, 4), (0.1, 0.1, 6), 32, 200000, None),
    "3d_voxels_nz8": (2, synth.NUSCENES_32, 5, (-12.8, -12.8, -5, 12.8, 12.8, 3), (0.4, 0.4, 1.0), 5, 4000, 3000),
}


def _make_case(name):
    nb, model, c, rng, vs, p, mv, keep = GROUP_CASES[name]
    frames = []
    for b in range(nb):
        f = synth.make_sweep(1000 + 17 * b + len(name), model, c)
        if keep is not None:
            f = f[np.hypot(f[:, 0], f[:, 1]) < 19.0][:keep]
        frames.append(f)
    offs = np.zeros(nb + 1, np.int32)
    offs[1:] = np.cumsum([len(f) for f in frames])
    return np.concatenate(frames, 0), offs, rng, vs, p, mv


@pytest.fixture(params=["hash", "dense"])
def grouping(request, L):
    """Both grouping implementations: the open-addressing hash table and the direct-mapped cell table."""
    L.ops.set_grouping(request.param)
    yield request.param
    L.ops.set_grouping("auto")


@pytest.mark.parametrize("name", list(GROUP_CASES))
def test_grouping_bit_exact(name, grouping, dev, L, oracle):
    pts, offs, rng, vs, p, mv = _make_case(name)
    grid = L.GridSpec.from_range(rng, vs, p, mv)
    ref = oracle.voxelize_batch(pts, offs, rng, vs, p, mv)
    res = L.ops.voxelize(torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev), grid, want_voxels=True,
                         want_membership=True)
    got = _trim(res)
    np.testing.assert_array_equal(got["pillar_count"][:-1], ref["pillars_per_frame"])
    assert got["m"] == ref["coords"].shape[0]
    np.testing.assert_array_equal(got["voxel_coords"], ref["coords"])
    np.testing.assert_array_equal(got["voxel_num_points"], ref["num_points"])
    np.testing.assert_array_equal(got["point_pillar"], ref["point_voxel"])
    np.testing.assert_array_equal(got["point_slot"], ref["point_slot"])
    np.testing.assert_array_equal(got["voxels"].view(np.uint32), ref["voxels"].view(np.uint32))


def test_grouping_pcdet_layout_and_frame_offsets(grouping, dev, L, oracle):
    """The collated [N, 1+C] layout with the frame index in column 0, including an empty middle frame."""
    pts, offs, rng, vs, p, mv = _make_case("small_p32")
    offs3 = np.array([offs[0], offs[1], offs[1], offs[2]], np.int32)  # frame 1 is empty
    pb = np.empty((len(pts), 6), np.float32)
    pb[:, 1:] = pts
    pb[:offs[1], 0] = 0
    pb[offs[1]:, 0] = 2
    t = torch.from_numpy(pb).to(dev)
    got_offs = L.ops.frame_offsets_from_points(t, 3)
    np.testing.assert_array_equal(got_offs.cpu().numpy(), offs3)
    grid = L.GridSpec.from_range(rng, vs, p, mv)
    res = _trim(L.ops.voxelize(t, got_offs, grid, col0=1, want_voxels=True, want_membership=True))
    ref = oracle.voxelize_batch(pts, offs3, rng, vs, p, mv)
    np.testing.assert_array_equal(res["voxel_coords"], ref["coords"])
    np.testing.assert_array_equal(res["point_slot"], ref["point_slot"])
    np.testing.assert_array_equal(res["voxels"], ref["voxels"])
    np.testing.assert_array_equal(res["pillar_count"][:-1], ref["pillars_per_frame"])


def test_grouping_edge_cases(grouping, dev, L, oracle):
    rng, vs = (0.0, 0.0, 0.0, 4.0, 4.0, 2.0), (1.0, 1.0, 2.0)
    grid = L.GridSpec.from_range(rng, vs, 4, 10)
    # boundaries, non-finite values, duplicates
    pts = np.array([[0, 0, 0, 1, 0], [4.0, 1, 1, 2, 0], [3.9999998, 1, 1, 3, 0], [1, 1, 2.0, 4, 0],
                    [-1e-7, 1, 1, 5, 0], [np.nan, 1, 1, 6, 0], [1, np.inf, 1, 7, 0], [0.5, 0.5, 0.5, 8, 0],
                    [0.5, 0.5, 0.5, 8, 0]], np.float32)
    offs = np.array([0, len(pts)], np.int32)
    got = _trim(L.ops.voxelize(torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev), grid,
                               want_membership=True))
    ref = oracle.voxelize_batch(pts, offs, rng, vs, 4, 10)
    for k, rk in (("voxel_coords", "coords"), ("voxel_num_points", "num_points"), ("point_pillar", "point_voxel"),
                  ("point_slot", "point_slot"), ("voxels", "voxels")):
        np.testing.assert_array_equal(got[k], ref[rk], err_msg=k)
    # no points at all / nothing in range / several empty frames
    for arr, o in ((np.zeros((0, 5), np.float32), [0, 0, 0]), (np.full((7, 5), 100.0, np.float32), [0, 3, 7]),
                   (pts, [0, 0, len(pts), len(pts)])):
        o = np.asarray(o, np.int32)
        got = _trim(L.ops.voxelize(torch.from_numpy(arr).to(dev), torch.from_numpy(o).to(dev), grid,
                                   want_membership=True))
        ref = oracle.voxelize_batch(arr, o, rng, vs, 4, 10)
        assert got["m"] == ref["coords"].shape[0]
        np.testing.assert_array_equal(got["pillar_count"][:-1], ref["pillars_per_frame"])
        np.testing.assert_array_equal(got["voxel_coords"], ref["coords"])
        np.testing.assert_array_equal(got["point_slot"], ref["point_slot"])


def test_grouping_one_huge_pillar_and_many_frames(grouping, dev, L, oracle):
    """20 000 points in one cell (cap 32 -> radix select over a long list) and 40 tiny frames (> one warp of frames)."""
    rng, vs = (0.0, 0.0, 0.0, 8.0, 8.0, 2.0), (1.0, 1.0, 2.0)
    r = np.random.default_rng(5)
    big = np.concatenate([r.uniform(3.0, 4.0, (20000, 2)), r.uniform(0, 2, (20000, 1)), r.uniform(0, 1, (20000, 2))],
                         1).astype(np.float32)
    rest = np.concatenate([r.uniform(-1, 9, (6000, 2)), r.uniform(-0.5, 2.5, (6000, 1)), r.uniform(0, 1, (6000, 2))],
                          1).astype(np.float32)
    pts = np.concatenate([big, rest])[r.permutation(26000)]
    cuts = np.sort(r.integers(0, 26000, 39))
    offs = np.concatenate([[0], cuts, [26000]]).astype(np.int32)
    grid = L.GridSpec.from_range(rng, vs, 32, 50)
    got = _trim(L.ops.voxelize(torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev), grid,
                               want_membership=True))
    ref = oracle.voxelize_batch(pts, offs, rng, vs, 32, 50)
    np.testing.assert_array_equal(got["voxel_coords"], ref["coords"])
    np.testing.assert_array_equal(got["voxel_num_points"], ref["num_points"])
    np.testing.assert_array_equal(got["point_pillar"], ref["point_voxel"])
    np.testing.assert_array_equal(got["point_slot"], ref["point_slot"])
    np.testing.assert_array_equal(got["voxels"], ref["voxels"])


# ------------------------------------------------------------------------------------------------
# golden vectors from the reference modules
# ------------------------------------------------------------------------------------------------
def _cfg(g):
    from lidar_vision_vqa_b200.synth import GridConfig  # noqa: F401

    class C(dict):
        __getattr__ = dict.__getitem__

    return C(USE_NORM=bool(g["use_norm"]), WITH_DISTANCE=bool(g["with_distance"]), USE_ABSLOTE_XYZ=bool(g["use_abs"]),
             NUM_FILTERS=[int(v) for v in g["num_filters"]])


SINGLE_LAYER = [n for n in vfe_golden_names() if "2layer" not in n]


@pytest.mark.parametrize("name", SINGLE_LAYER)
def test_pillar_vfe_module_vs_reference_golden(name, dev, L):
    g = load_golden(name)
    c = g["voxels"].shape[2]
    vfe = L.PillarVFE(model_cfg=_cfg(g), num_point_features=c, voxel_size=list(g["voxel_size"]),
                      point_cloud_range=g["range"], grid_size=g["grid_size"], depth_downsample_factor=None)
    missing = vfe.load_state_dict(_sd_t(g), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    vfe.eval().to(dev)
    bd = {"voxels": torch.from_numpy(g["voxels"]).to(dev),
          "voxel_num_points": torch.from_numpy(g["voxel_num_points"]).to(dev),
          "voxel_coords": torch.from_numpy(g["voxel_coords"]).to(dev), "batch_size": len(g["frame_offsets"]) - 1}
    out = vfe(bd)["pillar_features"].cpu().numpy()
    ref = g["out.pillar_features"]
    assert out.shape == ref.shape  # includes the squeeze() of M == 1
    np.testing.assert_allclose(out, ref, rtol=FEAT_RTOL, atol=FEAT_ATOL)
    np.testing.assert_allclose(out, ref, rtol=TIGHT_RTOL, atol=TIGHT_ATOL)


@pytest.mark.parametrize("variant", ["plain", "wide", "auto"])
@pytest.mark.parametrize("name", [n for n in SINGLE_LAYER if n != "vfe_c5_m1"])
def test_scatter_module_vs_reference_golden(name, variant, dev, L):
    g = load_golden(name)

    class C(dict):
        __getattr__ = dict.__getitem__

    sc = L.PointPillarScatter(model_cfg=C(NUM_BEV_FEATURES=64, SCATTER_VARIANT=variant), grid_size=g["grid_size"])
    bd = {"pillar_features": torch.from_numpy(g["out.pillar_features"]).to(dev),
          "voxel_coords": torch.from_numpy(g["voxel_coords"]).to(dev)}  # no batch_size: reference rule (coords max + 1)
    bev = sc(bd)["spatial_features"].cpu().numpy()
    ref = g["out.spatial_features"]
    assert bev.shape == ref.shape
    np.testing.assert_array_equal(bev.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("name", [n for n in SINGLE_LAYER if n != "vfe_c5_m1"])
def test_fused_points_to_bev_vs_reference_golden(name, dev, L):
    """points -> (grouping + PFN + scatter) in one call, against the reference modules' outputs."""
    g = load_golden(name)
    c = g["voxels"].shape[2]

    class C(dict):
        __getattr__ = dict.__getitem__

    cfg = _cfg(g)
    cfg.update(MAX_POINTS_PER_VOXEL=int(g["max_points"]), MAX_NUMBER_OF_VOXELS={"train": 1, "test": int(g["max_voxels"])},
               FUSE_SCATTER=True)
    vfe = L.PillarVFEFromPoints(model_cfg=cfg, num_point_features=c, voxel_size=list(g["voxel_size"]),
                                point_cloud_range=g["range"], grid_size=g["grid_size"])
    vfe.load_state_dict(_sd_t(g), strict=True)
    vfe.eval().to(dev)
    sc = L.PointPillarScatter(model_cfg=C(NUM_BEV_FEATURES=64), grid_size=g["grid_size"])
    pb = synth.to_pcdet_points(g["points"], g["frame_offsets"])
    bd = {"points": torch.from_numpy(pb).to(dev), "batch_size": len(g["frame_offsets"]) - 1}
    bd = sc(vfe(bd))
    np.testing.assert_array_equal(bd["voxel_coords"].cpu().numpy(), g["voxel_coords"].astype(np.int32))
    np.testing.assert_array_equal(bd["voxel_num_points"].cpu().numpy(), g["voxel_num_points"].astype(np.int32))
    np.testing.assert_allclose(bd["pillar_features"].cpu().numpy(), g["out.pillar_features"], rtol=FEAT_RTOL,
                               atol=FEAT_ATOL)
    bev = bd["spatial_features"].cpu().numpy()
    ref = g["out.spatial_features"]
    # trailing empty frames: the reference sizes the batch from coords (pointpillar_scatter.py:17)
    np.testing.assert_allclose(bev[:ref.shape[0]], ref, rtol=FEAT_RTOL, atol=FEAT_ATOL)
    assert (bev[ref.shape[0]:] == 0).all()
    assert ((bev == 0) == (np.pad(ref, [(0, bev.shape[0] - ref.shape[0])] + [(0, 0)] * 3) == 0)).all()


# ------------------------------------------------------------------------------------------------
# CUDA path vs the oracle on fresh seeded inputs (sizes the oracle finishes in seconds)
# ------------------------------------------------------------------------------------------------
def _run_fused(L, dev, pts, offs, grid, sd, c, variant="auto", **kw):
    pfn = L.ops.fold_pfn(sd["pfn_layers.0.linear.weight"],
                         (sd["pfn_layers.0.norm.weight"], sd["pfn_layers.0.norm.bias"],
                          sd["pfn_layers.0.norm.running_mean"], sd["pfn_layers.0.norm.running_var"], 1e-3), None,
                         c_point=c, use_absolute_xyz=True, with_distance=False, voxel_size=grid.voxel_size,
                         point_cloud_range=grid.point_cloud_range, device=dev)
    return L.ops.encode_bev(torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev), grid, pfn,
                            scatter_variant=variant, **kw)


@pytest.fixture(params=["fast", "generic"])
def feature_kernel(request, L):
    """Both feature kernels: the constant-bank fast one (default when eligible) and the generic 16-lane one."""
    L.ops.force_generic_features(request.param == "generic")
    yield request.param
    L.ops.force_generic_features(False)


@pytest.mark.parametrize("name", ["cfg1_full_sweep", "cfg1_p20", "tenSweep_maxvox30000_binds", "waymo_0.1m",
                                  "cap_binds_p4", "maxvox_binds", "p1_maxvox1"])
def test_fused_path_vs_oracle(name, feature_kernel, grouping, dev, L, oracle):
    pts, offs, rng, vs, p, mv = _make_case(name)
    c = pts.shape[1]
    grid = L.GridSpec.from_range(rng, vs, p, mv)
    sd = oracle.random_pfn_params(c + 6, [64], True, seed=3)
    res = _run_fused(L, dev, pts, offs, grid, sd, c)
    m = int(res["pillar_count"][-1].item())
    ref_v = oracle.voxelize_batch(pts, offs, rng, vs, p, mv)
    assert m == ref_v["coords"].shape[0]
    np.testing.assert_array_equal(res["voxel_coords"][:m].cpu().numpy(), ref_v["coords"])
    np.testing.assert_array_equal(res["voxel_num_points"][:m].cpu().numpy(), ref_v["num_points"])
    ref_f = oracle.pillar_vfe(ref_v["voxels"], ref_v["num_points"], ref_v["coords"], sd, vs, rng).numpy()
    got_f = res["pillar_features"][:m].cpu().numpy()
    np.testing.assert_allclose(got_f, ref_f, rtol=FEAT_RTOL, atol=FEAT_ATOL)
    np.testing.assert_allclose(got_f, ref_f, rtol=1e-4, atol=1e-4)
    nx, ny, _ = grid.grid_size
    nb = len(offs) - 1
    bev = res["bev"].cpu().numpy()
    # the scatter itself moves bits: exact against the oracle scatter of OUR features ...
    np.testing.assert_array_equal(bev, oracle.scatter_bev(got_f, ref_v["coords"], nx, ny, batch_size=nb))
    # ... and within tolerance of the all-oracle canvas, with exactly the same empty cells
    ref_bev = oracle.scatter_bev(ref_f, ref_v["coords"], nx, ny, batch_size=nb)
    np.testing.assert_allclose(bev, ref_bev, rtol=FEAT_RTOL, atol=FEAT_ATOL)


def test_fused_huge_pillar_spanning_chunks(feature_kernel, grouping, dev, L, oracle):
    """One cell holding 20 000 points (its list spans ~80 chunks of the fast kernel) next to ordinary pillars."""
    rng, vs = (0.0, 0.0, 0.0, 8.0, 8.0, 2.0), (1.0, 1.0, 2.0)
    r = np.random.default_rng(11)
    big = np.concatenate([r.uniform(3.0, 4.0, (20000, 2)), r.uniform(0, 2, (20000, 1)), r.uniform(0, 255, (20000, 1)),
                          r.uniform(0, 0.5, (20000, 1))], 1).astype(np.float32)
    rest = np.concatenate([r.uniform(-1, 9, (6000, 2)), r.uniform(-0.5, 2.5, (6000, 1)), r.uniform(0, 255, (6000, 1)),
                           r.uniform(0, 0.5, (6000, 1))], 1).astype(np.float32)
    pts = np.concatenate([big, rest])[r.permutation(26000)]
    offs = np.array([0, 9000, 9000, 26000], np.int32)
    for p_max in (32, 5, 100):  # 100: more kept points than the long-pillar path's 32-record scratch holds at once
        grid = L.GridSpec.from_range(rng, vs, p_max, 60)
        sd = oracle.random_pfn_params(11, [64], True, seed=13)
        res = _run_fused(L, dev, pts, offs, grid, sd, 5)
        m = int(res["pillar_count"][-1].item())
        ref_v = oracle.voxelize_batch(pts, offs, rng, vs, p_max, 60)
        assert m == ref_v["coords"].shape[0]
        np.testing.assert_array_equal(res["voxel_coords"][:m].cpu().numpy(), ref_v["coords"])
        np.testing.assert_array_equal(res["voxel_num_points"][:m].cpu().numpy(), ref_v["num_points"])
        ref_f = oracle.pillar_vfe(ref_v["voxels"], ref_v["num_points"], ref_v["coords"], sd, vs, rng).numpy()
        np.testing.assert_allclose(res["pillar_features"][:m].cpu().numpy(), ref_f, rtol=FEAT_RTOL, atol=FEAT_ATOL)


def test_fused_path_c4_and_scatter_variants_agree(dev, L, oracle):
    pts, offs, rng, vs, p, mv = _make_case("maxvox_binds")  # C = 4 (the product's nuScenes yaml)
    grid = L.GridSpec.from_range(rng, vs, p, mv)
    sd = oracle.random_pfn_params(10, [64], True, seed=4)
    outs = {v: _run_fused(L, dev, pts, offs, grid, sd, 4, variant=v)["bev"].clone() for v in
            ("plain", "wide", "auto")}
    assert torch.equal(outs["plain"], outs["wide"])
    assert torch.equal(outs["plain"], outs["auto"])
    ref_v = oracle.voxelize_batch(pts, offs, rng, vs, p, mv)
    ref_f = oracle.pillar_vfe(ref_v["voxels"], ref_v["num_points"], ref_v["coords"], sd, vs, rng).numpy()
    ref_bev = oracle.scatter_bev(ref_f, ref_v["coords"], grid.grid_size[0], grid.grid_size[1], batch_size=len(offs) - 1)
    np.testing.assert_allclose(outs["auto"].cpu().numpy(), ref_bev, rtol=FEAT_RTOL, atol=FEAT_ATOL)


@pytest.mark.parametrize("variant", ["plain", "wide", "auto"])
def test_scatter3d_module_vs_reference_golden(variant, dev, L):
    """PointPillarScatter3d (pointpillar_scatter.py:40-73): nz = 2, 32 channels per pillar -> [B, 64, ny, nx]."""
    g = load_golden("scatter3d_nz2")

    class C(dict):
        __getattr__ = dict.__getitem__

    sc = L.PointPillarScatter3d(model_cfg=C(INPUT_SHAPE=[int(v) for v in g["input_shape"]],
                                            NUM_BEV_FEATURES=int(g["num_bev_features"]), SCATTER_VARIANT=variant),
                                grid_size=None)
    for coords in (g["voxel_coords"], g["voxel_coords"].astype(np.int32)):
        bd = {"pillar_features": torch.from_numpy(g["pillar_features"]).to(dev),
              "voxel_coords": torch.from_numpy(coords).to(dev)}
        bev = sc(bd)["spatial_features"].cpu().numpy()
        ref = g["out.spatial_features"]
        assert bev.shape == ref.shape
        np.testing.assert_array_equal(bev.view(np.uint32), ref.view(np.uint32))


def test_scatter_odd_shapes(dev, L, oracle):
    """Grids that are not multiples of the 256-cell tile, of 4, and a channel count other than 64."""
    r = np.random.default_rng(0)
    for (nx, ny, f, nb) in ((432, 496, 64, 2), (100, 36, 32, 3), (37, 21, 64, 2), (64, 64, 128, 1)):
        m = min(700, nx * ny // 3)
        cells = np.stack([r.choice(nx * ny, m, replace=False) for _ in range(nb)])
        coords = np.concatenate([np.stack([np.full(m, b), np.zeros(m, int), cells[b] // nx, cells[b] % nx], 1)
                                 for b in range(nb)]).astype(np.int32)
        feats = r.standard_normal((m * nb, f)).astype(np.float32)
        ref = oracle.scatter_bev(feats, coords, nx, ny, batch_size=nb)
        for variant in ("plain", "wide", "auto"):
            for cd in (coords, coords.astype(np.float32)):
                bev = L.ops.scatter_bev(torch.from_numpy(feats).to(dev), torch.from_numpy(cd).to(dev), nb, nx, ny,
                                        variant=variant)
                np.testing.assert_array_equal(bev.cpu().numpy(), ref, err_msg=f"{nx}x{ny} f={f} {variant}")


# ------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties + full oracle comparison where it takes seconds
# ------------------------------------------------------------------------------------------------
def test_cfg2_full_size_properties_and_oracle(feature_kernel, dev, L, oracle):
    model, gc, nb = synth.WORKLOADS["cfg2_nuscenes32_b16_pillar0.2_bev512"]
    pts, offs = synth.make_batch(nb, model, 5)
    grid = L.GridSpec.from_range(gc.point_cloud_range, gc.voxel_size, gc.max_points_per_voxel, gc.max_voxels)
    sd = oracle.random_pfn_params(11, [64], True, seed=0)
    r1 = _run_fused(L, dev, pts, offs, grid, sd, 5, want_membership=True)
    snap = {k: v.clone() for k, v in r1.items()}
    r2 = _run_fused(L, dev, pts, offs, grid, sd, 5, want_membership=True)
    m = int(snap["pillar_count"][-1].item())
    for k in snap:  # idempotence / run-to-run determinism, bit for bit, although atomics land in any order
        rows = m if k in ("pillar_features", "voxel_coords", "voxel_num_points") else None  # rest is capacity padding
        assert torch.equal(snap[k][:rows], r2[k][:rows]), k
    counts = snap["pillar_count"][:-1].cpu().numpy()
    coords = snap["voxel_coords"][:m].cpu().numpy()
    npts = snap["voxel_num_points"][:m].cpu().numpy()
    slot = snap["point_slot"].cpu().numpy()
    pil = snap["point_pillar"].cpu().numpy()
    # conservation: every stored point is counted once; slots inside a pillar are 0..n-1
    assert (slot >= 0).sum() == npts.sum()
    assert np.array_equal(np.bincount(pil[slot >= 0], minlength=m), npts)
    assert (np.diff(coords[:, 0]) >= 0).all() and np.array_equal(np.bincount(coords[:, 0], minlength=nb), counts)
    # one canvas cell per pillar, zero elsewhere; channel sums of the canvas equal the feature sums (linearity)
    bev = snap["bev"]
    assert int((bev != 0).any(dim=1).sum().item()) <= m
    occ = torch.zeros((nb, 512, 512), dtype=torch.bool, device=dev)
    ct = snap["voxel_coords"][:m].long()
    occ[ct[:, 0], ct[:, 2], ct[:, 3]] = True
    assert int(occ.sum().item()) == m
    assert bool(((bev != 0).any(dim=1) <= occ).all())
    fsum = torch.zeros((nb, 64), dtype=torch.float64, device=dev).index_add_(0, ct[:, 0], snap["pillar_features"][:m].double())
    assert torch.allclose(bev.double().sum(dim=(2, 3)), fsum, rtol=1e-9, atol=1e-9)
    # full oracle comparison (a few seconds of CPU)
    ref_v = oracle.voxelize_batch(pts, offs, gc.point_cloud_range, gc.voxel_size, gc.max_points_per_voxel, gc.max_voxels)
    np.testing.assert_array_equal(coords, ref_v["coords"])
    np.testing.assert_array_equal(npts, ref_v["num_points"])
    np.testing.assert_array_equal(slot, ref_v["point_slot"])
    np.testing.assert_array_equal(pil, ref_v["point_voxel"])
    ref_f = oracle.pillar_vfe(ref_v["voxels"], ref_v["num_points"], ref_v["coords"], sd, gc.voxel_size,
                              gc.point_cloud_range).numpy()
    np.testing.assert_allclose(snap["pillar_features"][:m].cpu().numpy(), ref_f, rtol=FEAT_RTOL, atol=FEAT_ATOL)


def test_hard_variant_equals_fused_variant(dev, L, oracle):
    """PillarVFE('voxels' from the CPU voxeliser) and PillarVFEFromPoints('points') are two routes to the same answer."""
    pts, offs, rng, vs, p, mv = _make_case("cfg1_p20")
    grid = L.GridSpec.from_range(rng, vs, p, mv)
    sd = oracle.random_pfn_params(11, [64], True, seed=9)
    res = _run_fused(L, dev, pts, offs, grid, sd, 5)
    m = int(res["pillar_count"][-1].item())
    v = oracle.voxelize_batch(pts, offs, rng, vs, p, mv)

    class C(dict):
        __getattr__ = dict.__getitem__

    vfe = L.PillarVFE(model_cfg=C(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64]),
                      num_point_features=5, voxel_size=list(vs), point_cloud_range=np.asarray(rng, np.float32),
                      grid_size=grid.grid_size)
    vfe.load_state_dict(sd)
    vfe.eval().to(dev)
    voxels, npts, coords = oracle.collate_voxels(v)
    out = vfe({"voxels": torch.from_numpy(voxels).to(dev), "voxel_num_points": torch.from_numpy(npts).to(dev),
               "voxel_coords": torch.from_numpy(coords).to(dev)})["pillar_features"]
    torch.testing.assert_close(out, res["pillar_features"][:m], rtol=1e-5, atol=1e-5)


def test_cpu_tensors_and_training_mode_fail_loudly(dev, L):
    class C(dict):
        __getattr__ = dict.__getitem__

    vfe = L.PillarVFE(model_cfg=C(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64]),
                      num_point_features=5, voxel_size=[0.2, 0.2, 8], point_cloud_range=[-1, -1, -1, 1, 1, 1])
    bd = {"voxels": torch.zeros(3, 4, 5), "voxel_num_points": torch.ones(3), "voxel_coords": torch.zeros(3, 4)}
    with pytest.raises(L.NativeLibraryError):
        vfe.eval()(dict(bd))
    with pytest.raises(RuntimeError):
        vfe.train().to(dev)({k: v.to(dev) for k, v in bd.items()})


# ---------------------------------------------------------------------------------------------------------------------
# general feature stack (csrc/pfn_multi.cu): two-layer PFN, DynamicPillarVFE, DynamicPillarVFESimple2D
# ---------------------------------------------------------------------------------------------------------------------
DYN_GOLDENS = ["dyn_c5", "dyn_c5_2layer_zout", "dyn2d_c5_f32", "dyn_c4_dist_noabs"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", DYN_GOLDENS)
def test_dynamic_vfe_modules_vs_reference_golden(name, dev, L):
    """DynPillarVFE / DynamicPillarVFESimple2D drop-ins on the reference's own inputs: voxel_coords / pillar_coords and the
    row order bit-exact, pillar_features within rtol 1e-3 (fp32)."""
    g = load_golden(name)
    simple = bool(g["simple2d"])
    cls = L.DynamicPillarVFESimple2D if simple else L.DynamicPillarVFE
    cfg = C(USE_NORM=True, WITH_DISTANCE=bool(g["with_distance"]), USE_ABSLOTE_XYZ=bool(g["use_abs"]),
            NUM_FILTERS=[int(v) for v in g["num_filters"]])
    vfe = cls(model_cfg=cfg, num_point_features=int(g["c"]), voxel_size=[float(v) for v in g["voxel_size"]],
              grid_size=g["grid_size"], point_cloud_range=g["range"])
    vfe.load_state_dict(_sd_t(g), strict=True)
    vfe.eval().to(dev)
    bd = vfe({"points": torch.from_numpy(g["points_b"]).to(dev), "batch_size": int(g["batch"])})
    key = "pillar_coords" if simple else "voxel_coords"
    got_c = bd[key].cpu().numpy()
    np.testing.assert_array_equal(got_c, g["out.voxel_coords"])
    assert bd[key].dtype == torch.int32
    np.testing.assert_allclose(bd["pillar_features"].cpu().numpy(), g["out.pillar_features"], rtol=1e-3, atol=1e-5)
    if not simple:
        assert bd["voxel_features"] is bd["pillar_features"]
    # the batch size can also be derived from the points (the reference never reads batch_size here)
    bd2 = vfe({"points": torch.from_numpy(g["points_b"]).to(dev)})
    np.testing.assert_array_equal(bd2[key].cpu().numpy(), g["out.voxel_coords"])


@pytest.mark.gpu
def test_dynamic_vfe_counts_and_scatter_roundtrip(dev, L, oracle):
    """Uncapped per-pillar counts equal torch.unique's, and the dynamic coords feed PointPillarScatter unchanged."""
    g = load_golden("dyn_c5")
    cfg = C(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64])
    vfe = L.DynamicPillarVFE(model_cfg=cfg, num_point_features=5, voxel_size=[float(v) for v in g["voxel_size"]],
                             grid_size=g["grid_size"], point_cloud_range=g["range"])
    vfe.load_state_dict(_sd_t(g), strict=True)
    vfe.eval().to(dev)
    bd = vfe({"points": torch.from_numpy(g["points_b"]).to(dev), "batch_size": 2})
    _, cnt = oracle.dynamic_pillar_sets(g["points_b"], g["range"], g["voxel_size"])
    np.testing.assert_array_equal(bd["voxel_num_points"].cpu().numpy(), cnt)
    sc = L.PointPillarScatter(model_cfg=C(NUM_BEV_FEATURES=64), grid_size=g["grid_size"])
    bd = sc(bd)
    ref = oracle.scatter_bev(bd["pillar_features"].cpu().numpy(), bd["voxel_coords"].cpu().numpy(),
                             int(g["grid_size"][0]), int(g["grid_size"][1]), batch_size=2)
    np.testing.assert_array_equal(bd["spatial_features"].cpu().numpy(), ref)


@pytest.mark.gpu
def test_two_layer_pillar_vfe_vs_reference_golden(dev, L):
    """NUM_FILTERS [64, 64] (waymo_models/pointpillar_1x.yaml:34) on the reference's padded voxels: the padded-slot row is
    NOT re-masked between the layers (pillar_vfe.py:44-49,119-120)."""
    g = load_golden("vfe_c5_2layer")
    vfe = L.PillarVFE(model_cfg=_cfg(g), num_point_features=5, voxel_size=[float(v) for v in g["voxel_size"]],
                      point_cloud_range=g["range"], grid_size=g["grid_size"])
    vfe.load_state_dict(_sd_t(g), strict=True)
    vfe.eval().to(dev)
    for as_int in (False, True):
        npts = torch.from_numpy(g["voxel_num_points"]).to(dev)
        crd = torch.from_numpy(g["voxel_coords"]).to(dev)
        if as_int:
            npts, crd = npts.int(), crd.int()
        bd = vfe({"voxels": torch.from_numpy(g["voxels"]).to(dev), "voxel_num_points": npts, "voxel_coords": crd})
        np.testing.assert_allclose(bd["pillar_features"].cpu().numpy(), g["out.pillar_features"], rtol=1e-3, atol=1e-5)


@pytest.mark.gpu
def test_two_layer_fused_points_to_bev_vs_reference_golden(dev, L):
    """points -> grouping -> two-layer stack -> BEV in one call equals the reference's PillarVFE + PointPillarScatter on the
    voxels the hard voxeliser makes of the same points."""
    g = load_golden("vfe_c5_2layer")
    cfg = _cfg(g)
    cfg["MAX_POINTS_PER_VOXEL"] = int(g["max_points"])
    cfg["MAX_NUMBER_OF_VOXELS"] = int(g["max_voxels"])
    cfg["FUSE_SCATTER"] = True
    vfe = L.PillarVFEFromPoints(model_cfg=cfg, num_point_features=5, voxel_size=[float(v) for v in g["voxel_size"]],
                                point_cloud_range=g["range"], grid_size=g["grid_size"])
    vfe.load_state_dict(_sd_t(g), strict=True)
    vfe.eval().to(dev)
    offs = g["frame_offsets"]
    from lidar_vision_vqa_b200 import synth

    pb = torch.from_numpy(synth.to_pcdet_points(g["points"], offs)).to(dev)
    bd = vfe({"points": pb, "batch_size": len(offs) - 1})
    np.testing.assert_array_equal(bd["voxel_coords"].cpu().numpy(), g["voxel_coords"].astype(np.int32))
    np.testing.assert_array_equal(bd["voxel_num_points"].cpu().numpy(), g["voxel_num_points"].astype(np.int32))
    np.testing.assert_allclose(bd["pillar_features"].cpu().numpy(), g["out.pillar_features"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(bd["spatial_features"].cpu().numpy(), g["out.spatial_features"], rtol=1e-3, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["vfe_c5_p32_f32coords", "vfe_c5_p8_capbinds_maxvox200", "vfe_c5_dist_noabs"])
def test_general_kernel_agrees_with_single_layer_goldens(name, dev, L):
    """The general kernel run on single-layer configurations (hard semantics, caps binding) against the reference."""
    from lidar_vision_vqa_b200 import ops, synth

    g = load_golden(name)
    sd = _sd_t(g)
    grid = L.GridSpec(tuple(float(v) for v in g["range"]), tuple(float(v) for v in g["voxel_size"]),
                      tuple(int(v) for v in g["grid_size"]), int(g["max_points"]), int(g["max_voxels"]))
    bn = (sd["pfn_layers.0.norm.weight"], sd["pfn_layers.0.norm.bias"], sd["pfn_layers.0.norm.running_mean"],
          sd["pfn_layers.0.norm.running_var"], 1e-3)
    stack = ops.fold_pfn_stack([(sd["pfn_layers.0.linear.weight"], bn, None)], c_point=g["points"].shape[1],
                               use_absolute_xyz=bool(g["use_abs"]), with_distance=bool(g["with_distance"]),
                               voxel_size=grid.voxel_size, point_cloud_range=grid.point_cloud_range, device=dev)
    res = ops.encode_stack(torch.from_numpy(g["points"]).to(dev), torch.from_numpy(g["frame_offsets"]).to(dev), grid,
                           stack, with_bev=True)
    m = int(res["pillar_count"][-1].item())
    assert m == g["voxel_coords"].shape[0]
    np.testing.assert_array_equal(res["voxel_coords"][:m].cpu().numpy(), g["voxel_coords"].astype(np.int32))
    np.testing.assert_allclose(res["pillar_features"][:m].cpu().numpy(), g["out.pillar_features"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(res["bev"].cpu().numpy(), g["out.spatial_features"], rtol=1e-3, atol=1e-5)


# ---------------------------------------------------------------------------------------------------------------------
# float16 canvas + the bulk extractor (src/get-data/precompute_bev_features.py)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["vfe_c5_p32_f32coords", "vfe_c4_p20_i32coords"])
def test_half_canvas_equals_reference_astype_float16(name, dev, L):
    """PointPillarScatter(OUT_DTYPE float16) on the reference's own pillar features == the reference canvas cast the way
    the extractor casts it (precompute_bev_features.py:394 astype(np.float16)): bit-exact."""
    g = load_golden(name)
    sc = L.PointPillarScatter(model_cfg=C(NUM_BEV_FEATURES=64, OUT_DTYPE="float16"), grid_size=g["grid_size"])
    bd = sc({"pillar_features": torch.from_numpy(g["out.pillar_features"]).to(dev),
             "voxel_coords": torch.from_numpy(g["voxel_coords"]).to(dev), "batch_size": len(g["frame_offsets"]) - 1})
    got = bd["spatial_features"]
    assert got.dtype == torch.float16
    want = g["out.spatial_features"].astype(np.float16)
    np.testing.assert_array_equal(got.cpu().numpy().view(np.uint16), want.view(np.uint16))


@pytest.mark.gpu
def test_bev_extractor_matches_oracle_and_reference_file_format(dev, L, oracle, tmp_path):
    """BevExtractor over 5 frames in batches of 2 (ragged last batch): per-token float16 [C,H,W] .npy files whose content
    equals voxelise -> PillarVFE -> PointPillarScatter -> astype(float16) of the CPU oracle (fp16 of values within 1e-3)."""
    from lidar_vision_vqa_b200 import synth
    from lidar_vision_vqa_b200.extract import BevExtractor

    rng, vs, p, mv = (-25.6, -25.6, -5.0, 25.6, 25.6, 3.0), (0.4, 0.4, 8.0), 16, 5000
    frames = [synth.make_sweep(300 + i, synth.NUSCENES_32, 5)[:6000 + 500 * i] for i in range(5)]
    sd = oracle.random_pfn_params(11, [64], True, seed=9)
    cfg = C(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64], MAX_POINTS_PER_VOXEL=p,
            MAX_NUMBER_OF_VOXELS=mv)
    grid_size = oracle.grid_size_of(rng, vs)
    vfe = L.PillarVFEFromPoints(model_cfg=cfg, num_point_features=5, voxel_size=list(vs),
                                point_cloud_range=np.asarray(rng, np.float32), grid_size=grid_size)
    vfe.load_state_dict(sd)
    vfe.eval().to(dev)
    ex = BevExtractor(vfe, batch_size=2, max_points_per_frame=9000, depth=2)
    items = [(f"tok{i:02d}", f) for i, f in enumerate(frames)]
    n = ex.run_to_dir(items, str(tmp_path))
    assert n == 5
    nx, ny = int(grid_size[0]), int(grid_size[1])
    for tok, f in items:
        got = np.load(tmp_path / f"{tok}.npy")
        assert got.dtype == np.float16 and got.shape == (64, ny, nx)
        v = oracle.voxelize_hard(f, rng, vs, p, mv)
        c4 = np.concatenate([np.zeros((len(v["coords"]), 1), np.int32), v["coords"]], 1)
        feats = oracle.pillar_vfe(v["voxels"], v["num_points"], c4, sd, vs, rng).numpy()
        ref = oracle.scatter_bev(feats, c4, nx, ny, batch_size=1)[0]
        # occupancy pattern exact, values fp16(within 1e-3 of the reference)
        np.testing.assert_array_equal(got != 0, ref.astype(np.float16) != 0)
        np.testing.assert_allclose(got.astype(np.float32), ref, rtol=2e-3, atol=1e-3)
    # generator form returns the same arrays in input order
    toks = [t for t, _ in ex.run(items)]
    assert toks == [t for t, _ in items]


@pytest.mark.gpu
def test_fused_half_canvas_is_the_cast_of_the_float_canvas(dev, L, oracle):
    """encode_bev with a float16 canvas == float32 canvas of the same call cast with round-to-nearest-even."""
    from lidar_vision_vqa_b200 import ops, synth

    rng, vs = (-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), (0.2, 0.2, 8.0)
    pts, offs = synth.make_batch(2, synth.NUSCENES_32, 5, seed0=21)
    grid = L.GridSpec.from_range(rng, vs, 32, 30000)
    sd = oracle.random_pfn_params(11, [64], True, seed=2)
    pfn = ops.fold_pfn(sd["pfn_layers.0.linear.weight"],
                       (sd["pfn_layers.0.norm.weight"], sd["pfn_layers.0.norm.bias"],
                        sd["pfn_layers.0.norm.running_mean"], sd["pfn_layers.0.norm.running_var"], 1e-3), None,
                       c_point=5, use_absolute_xyz=True, with_distance=False, voxel_size=vs, point_cloud_range=rng,
                       device=dev)
    p_d, o_d = torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev)
    b32 = ops.EncodeBuffers(len(pts), 2, grid, 64, dev)
    b16 = ops.EncodeBuffers(len(pts), 2, grid, 64, dev, bev_dtype=torch.float16)
    r32 = ops.encode_bev(p_d, o_d, grid, pfn, buffers=b32)
    r16 = ops.encode_bev(p_d, o_d, grid, pfn, buffers=b16)
    torch.cuda.synchronize()
    assert r16["bev"].dtype == torch.float16
    np.testing.assert_array_equal(r16["bev"].cpu().numpy().view(np.uint16),
                                  r32["bev"].cpu().numpy().astype(np.float16).view(np.uint16))


@pytest.mark.gpu
@pytest.mark.parametrize("filters", [[64, 64], [32], [48, 40]])
def test_general_kernel_vs_oracle_caps_binding(filters, dev, L, oracle):
    """Two-layer / narrow stacks through the fused hard path where BOTH caps bind (P = 4 points per pillar, 150 pillars
    per frame) and pillars far longer than the cap exist: grouping bit-exact, features rtol 1e-3 vs the CPU oracle
    (pillar_vfe.py:44-49: the padded row is not re-masked between layers)."""
    from lidar_vision_vqa_b200 import synth

    rng, vs, p, mv = (-12.8, -12.8, -5.0, 12.8, 12.8, 3.0), (0.8, 0.8, 8.0), 4, 150
    frames = [synth.make_sweep(500 + i, synth.NUSCENES_32, 5)[:5000] for i in range(3)]
    frames[1] = frames[1][:0]  # an empty frame in the middle
    offs = np.zeros(4, np.int32)
    offs[1:] = np.cumsum([len(f) for f in frames])
    pts = np.concatenate(frames, 0)
    sd = oracle.random_pfn_params(11, filters, True, seed=3)
    cfg = C(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=filters, MAX_POINTS_PER_VOXEL=p,
            MAX_NUMBER_OF_VOXELS=mv, FUSE_SCATTER=True)
    grid_size = oracle.grid_size_of(rng, vs)
    vfe = L.PillarVFEFromPoints(model_cfg=cfg, num_point_features=5, voxel_size=list(vs),
                                point_cloud_range=np.asarray(rng, np.float32), grid_size=grid_size)
    vfe.load_state_dict(sd)
    vfe.eval().to(dev)
    bd = vfe({"points": torch.from_numpy(synth.to_pcdet_points(pts, offs)).to(dev), "batch_size": 3})
    ref = oracle.voxelize_batch(pts, offs, rng, vs, p, mv)
    assert ref["num_points"].max() == p and (ref["pillars_per_frame"] == [mv, 0, mv]).all()
    np.testing.assert_array_equal(bd["voxel_coords"].cpu().numpy(), ref["coords"])
    np.testing.assert_array_equal(bd["voxel_num_points"].cpu().numpy(), ref["num_points"])
    ref_f = oracle.pillar_vfe(ref["voxels"], ref["num_points"], ref["coords"], sd, vs, rng).numpy()
    np.testing.assert_allclose(bd["pillar_features"].cpu().numpy(), ref_f, rtol=1e-3, atol=1e-5)
    ref_bev = oracle.scatter_bev(bd["pillar_features"].cpu().numpy(), ref["coords"], int(grid_size[0]), int(grid_size[1]),
                                 batch_size=3)
    np.testing.assert_array_equal(bd["spatial_features"].cpu().numpy(), ref_bev)
    # the same stack on the reference's padded voxels (hard module) gives the same rows
    hard = L.PillarVFE(model_cfg=C(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=filters),
                       num_point_features=5, voxel_size=list(vs), point_cloud_range=np.asarray(rng, np.float32),
                       grid_size=grid_size)
    hard.load_state_dict(sd)
    hard.eval().to(dev)
    bd2 = hard({"voxels": torch.from_numpy(ref["voxels"]).to(dev),
                "voxel_num_points": torch.from_numpy(ref["num_points"]).to(dev).float(),
                "voxel_coords": torch.from_numpy(ref["coords"]).to(dev).float()})
    np.testing.assert_allclose(bd2["pillar_features"].cpu().numpy(), ref_f, rtol=1e-3, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("simple", [False, True])
def test_dynamic_vfe_vs_oracle_edge_cases(simple, dev, L, oracle):
    """Dynamic variants against the CPU restatement on a fresh case: an empty frame, a frame whose points all lie outside
    x/y, one pillar holding thousands of points (no cap), points far outside z, and an empty batch."""
    from lidar_vision_vqa_b200 import synth

    rng, vs = (-20.0, -20.0, -5.0, 20.0, 20.0, 3.0), (0.5, 0.5, 8.0)
    r = np.random.default_rng(11)
    f0 = synth.make_sweep(700, synth.NUSCENES_32, 5)[:4000]
    f0[::5, 2] += 30.0
    f1 = f0[:0]
    f2 = f0[:300].copy()
    f2[:, :2] += 500.0  # nothing in range
    f3 = np.concatenate([np.tile(np.array([[1.23, -4.56, 0.1, 7.0, 0.0]], np.float32), (3000, 1)) +
                         r.normal(0, 0.01, (3000, 5)).astype(np.float32), f0[:500]], 0)
    frames = [f0, f1, f2, f3]
    offs = np.zeros(5, np.int32)
    offs[1:] = np.cumsum([len(f) for f in frames])
    pb = synth.to_pcdet_points(np.concatenate(frames, 0), offs)
    filters = [32] if simple else [64, 64]
    c_in = 8 if simple else 11
    sd = oracle.random_pfn_params(c_in, filters, True, seed=5)
    cls = L.DynamicPillarVFESimple2D if simple else L.DynamicPillarVFE
    vfe = cls(model_cfg=C(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=filters),
              num_point_features=5, voxel_size=list(vs), grid_size=oracle.grid_size_of(rng, vs),
              point_cloud_range=np.asarray(rng, np.float32))
    vfe.load_state_dict(sd)
    vfe.eval().to(dev)
    bd = vfe({"points": torch.from_numpy(pb).to(dev), "batch_size": 4})
    ref_f, ref_c, ref_n = oracle.dynamic_pillar_vfe(pb, sd, vs, rng, simple2d=simple)
    key = "pillar_coords" if simple else "voxel_coords"
    np.testing.assert_array_equal(bd[key].cpu().numpy(), ref_c)
    np.testing.assert_array_equal(bd["voxel_num_points"].cpu().numpy(), ref_n)
    assert ref_n.max() >= 3000
    np.testing.assert_allclose(bd["pillar_features"].cpu().numpy(), ref_f.numpy(), rtol=1e-3, atol=1e-5)
    empty = vfe({"points": torch.zeros((0, 6), device=dev), "batch_size": 2})
    assert empty["pillar_features"].shape == (0, filters[-1]) and empty[key].shape[0] == 0


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[2] and configs[3] at their full batch sizes
# ---------------------------------------------------------------------------------------------------------------------
def _sampled_feature_check(oracle, ref_v, got_f, sd, vs, rng, max_points, step=11):
    """Pillars are independent, so the CPU oracle (pillar_vfe.py:94-123) is evaluated on every `step`-th pillar plus every
    pillar that reached the per-pillar cap; keeps the full-size tests at seconds of CPU."""
    m = ref_v["coords"].shape[0]
    sel = np.zeros(m, bool)
    sel[::step] = True
    sel |= ref_v["num_points"] >= max_points
    sel &= np.cumsum(sel) <= 120000
    idx = np.nonzero(sel)[0]
    ref_f = oracle.pillar_vfe(ref_v["voxels"][idx], ref_v["num_points"][idx], ref_v["coords"][idx], sd, vs, rng).numpy()
    np.testing.assert_allclose(got_f[idx], ref_f, rtol=FEAT_RTOL, atol=FEAT_ATOL)
    return idx.size


FULL_CASES = {
    # name: (workload, max_voxels override or None, NUM_FILTERS)
    "cfg3_b8_maxvox200000": ("cfg3_10sweep_p32_b8", None, [64]),
    "cfg3_b8_maxvox30000_binds": ("cfg3_10sweep_p32_b8", 30000, [64]),
    "cfg4_b8_1024": ("cfg4_waymo64_pillar0.1_bev1024", None, [64]),
    # Waymo's own PFN (tools/cfgs/waymo_models/pointpillar_1x.yaml:34)
    "cfg4_b8_1024_filters64_64": ("cfg4_waymo64_pillar0.1_bev1024", None, [64, 64]),
    "cfg3_b8_filters64_64": ("cfg3_10sweep_p32_b8", None, [64, 64]),
}


@pytest.mark.parametrize("name", list(FULL_CASES))
def test_full_size_configs_vs_oracle(name, dev, L, oracle):
    """cfg3 (8 x ~315k points, 10-sweep clouds, P = 32; max_voxels 200 000 and the binding 30 000) and cfg4 (8 x ~160k
    points, 0.1 m pillars, 1024^2 canvas) at BASELINE.json's batch sizes: grouping bit-exact against the CPU voxeliser,
    pillar features within rtol 1e-3 of the CPU PillarVFE, canvas = exact scatter of those features."""
    wl, mv_override, filters = FULL_CASES[name]
    model, gc, nb = synth.WORKLOADS[wl]
    mv = gc.max_voxels if mv_override is None else mv_override
    p = gc.max_points_per_voxel
    pts, offs = synth.make_batch(nb, model, 5)
    rng, vs = gc.point_cloud_range, gc.voxel_size
    nx, ny, _ = gc.grid_size
    sd = oracle.random_pfn_params(11, filters, True, seed=21)
    cfg = C(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=filters, MAX_POINTS_PER_VOXEL=p,
            MAX_NUMBER_OF_VOXELS=mv, FUSE_SCATTER=True)
    vfe = L.PillarVFEFromPoints(model_cfg=cfg, num_point_features=5, voxel_size=list(vs),
                                point_cloud_range=np.asarray(rng, np.float32), grid_size=gc.grid_size)
    vfe.load_state_dict(sd)
    vfe.eval().to(dev)
    bd = vfe({"points": torch.from_numpy(synth.to_pcdet_points(pts, offs)).to(dev), "batch_size": nb})
    ref_v = oracle.voxelize_batch(pts, offs, rng, vs, p, mv)
    m = ref_v["coords"].shape[0]
    if mv_override is not None:
        assert (ref_v["pillars_per_frame"] == mv).all(), "the cap is meant to bind in this case"
    np.testing.assert_array_equal(bd["pillars_per_frame"].numpy(), ref_v["pillars_per_frame"])
    np.testing.assert_array_equal(bd["voxel_coords"].cpu().numpy(), ref_v["coords"])
    np.testing.assert_array_equal(bd["voxel_num_points"].cpu().numpy(), ref_v["num_points"])
    got_f = bd["pillar_features"].cpu().numpy()
    assert got_f.shape == (m, filters[-1])
    assert _sampled_feature_check(oracle, ref_v, got_f, sd, vs, rng, p) > 10000
    # the canvas holds exactly those rows: compare a hashed projection of every frame instead of moving 2 GiB to the host
    bev = bd["spatial_features"]
    assert bev.shape == (nb, filters[-1], ny, nx)
    ct = bd["voxel_coords"].long()
    gathered = bev[ct[:, 0], :, ct[:, 2], ct[:, 3]]
    assert torch.equal(gathered, bd["pillar_features"])
    assert int((bev != 0).any(dim=1).sum().item()) <= m
    fsum = torch.zeros((nb, filters[-1]), dtype=torch.float64, device=dev).index_add_(0, ct[:, 0],
                                                                                     bd["pillar_features"].double())
    assert torch.allclose(bev.double().sum(dim=(2, 3)), fsum, rtol=1e-9, atol=1e-6)


def test_full_size_cfg3_membership_bit_exact(grouping, dev, L, oracle):
    """Point -> pillar membership and slots at cfg3's full size (2.5 M points, 3 % of the pillars over the cap), through
    both grouping implementations."""
    model, gc, nb = synth.WORKLOADS["cfg3_10sweep_p32_b8"]
    pts, offs = synth.make_batch(nb, model, 5)
    grid = L.GridSpec.from_range(gc.point_cloud_range, gc.voxel_size, gc.max_points_per_voxel, gc.max_voxels)
    res = L.ops.voxelize(torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev), grid, want_voxels=False,
                         want_membership=True)
    ref = oracle.voxelize_batch(pts, offs, gc.point_cloud_range, gc.voxel_size, gc.max_points_per_voxel, gc.max_voxels,
                                want_voxels=False)
    m = int(res["pillar_count"][-1].item())
    assert m == ref["coords"].shape[0]
    np.testing.assert_array_equal(res["voxel_coords"][:m].cpu().numpy(), ref["coords"])
    np.testing.assert_array_equal(res["voxel_num_points"][:m].cpu().numpy(), ref["num_points"])
    np.testing.assert_array_equal(res["point_pillar"].cpu().numpy(), ref["point_voxel"])
    np.testing.assert_array_equal(res["point_slot"].cpu().numpy(), ref["point_slot"])


# ---------------------------------------------------------------------------------------------------------------------
# the reference's pre-filter is subsumed by the voxeliser's own bounds check
# ---------------------------------------------------------------------------------------------------------------------
def test_range_mask_then_voxelise_equals_voxelise(grouping, dev, L, oracle):
    """data_processor.py:79-93 filters with common_utils.mask_points_by_range (utils/common_utils.py:78-81: x,y in
    [min, max] INCLUSIVE) before voxelising.  The voxeliser rejects a superset (x == max quantises to cell nx, which is
    out of the grid), and filtering keeps the relative order of the survivors, so voxelise(mask(p)) == voxelise(p):
    same pillars in the same order, same points in the same slots."""
    rng, vs, p, mv = (-12.8, -12.8, -5.0, 12.8, 12.8, 3.0), (0.4, 0.4, 8.0), 8, 600
    frames = []
    for b in range(3):
        f = synth.make_sweep(900 + b, synth.NUSCENES_32, 5)[:9000]
        # points exactly on every face of the range, just inside and just outside
        edge = np.array([[12.8, 0.3, 0, 1, 0], [-12.8, 0.3, 0, 2, 0], [0.3, 12.8, 0, 3, 0], [0.3, -12.8, 0, 4, 0],
                         [np.nextafter(np.float32(12.8), np.float32(0)), 1.1, 0, 5, 0],
                         [np.nextafter(np.float32(12.8), np.float32(20)), 1.1, 0, 6, 0], [12.8, 12.8, 0, 7, 0],
                         [-12.8, -12.8, 0, 8, 0]], np.float32)
        f[100:100 + len(edge)] = edge
        frames.append(f)
    masked = []
    for f in frames:  # mask_points_by_range, restated (common_utils.py:78-81)
        keep = (f[:, 0] >= rng[0]) & (f[:, 0] <= rng[3]) & (f[:, 1] >= rng[1]) & (f[:, 1] <= rng[4])
        assert 0 < keep.sum() < len(f)
        masked.append(f[keep])
    grid = L.GridSpec.from_range(rng, vs, p, mv)

    def run(fr):
        offs = np.zeros(len(fr) + 1, np.int32)
        offs[1:] = np.cumsum([len(f) for f in fr])
        pts = np.concatenate(fr, 0)
        out = _trim(L.ops.voxelize(torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev), grid, want_voxels=True,
                                   want_membership=True))
        return pts, offs, out

    pts_a, offs_a, a = run(frames)
    pts_b, offs_b, b = run(masked)
    for k in ("pillar_count", "voxel_coords", "voxel_num_points", "voxels"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    # membership of the surviving points is unchanged; the dropped ones were rejected by the voxeliser anyway
    keep_all = np.concatenate([(f[:, 0] >= rng[0]) & (f[:, 0] <= rng[3]) & (f[:, 1] >= rng[1]) & (f[:, 1] <= rng[4])
                               for f in frames])
    np.testing.assert_array_equal(a["point_pillar"][keep_all], b["point_pillar"])
    np.testing.assert_array_equal(a["point_slot"][keep_all], b["point_slot"])
    assert (a["point_pillar"][~keep_all] == -1).all()
    # x == range_max passes the reference's inclusive mask but belongs to no cell: both routes drop it
    on_max = np.nonzero((pts_b[:, 0] == np.float32(rng[3])) | (pts_b[:, 1] == np.float32(rng[4])))[0]
    assert on_max.size >= 6 and (b["point_pillar"][on_max] == -1).all()
    ref = oracle.voxelize_batch(pts_b, offs_b, rng, vs, p, mv)
    np.testing.assert_array_equal(b["voxel_coords"], ref["coords"])
    np.testing.assert_array_equal(b["point_slot"], ref["point_slot"])


def test_grouping_1000_frames(grouping, dev, L, oracle):
    """1000 small frames in one call (the API allows 1024): frame tables, per-frame caps and tiles that straddle many
    frames."""
    rng, vs, p, mv = (0.0, 0.0, 0.0, 6.4, 6.4, 2.0), (0.4, 0.4, 2.0), 3, 40
    r = np.random.default_rng(77)
    sizes = r.integers(0, 120, 1000)
    sizes[[0, 17, 500, 999]] = 0
    offs = np.zeros(1001, np.int32)
    offs[1:] = np.cumsum(sizes)
    n = int(offs[-1])
    pts = np.concatenate([r.uniform(-0.3, 6.7, (n, 2)), r.uniform(-0.2, 2.2, (n, 1)), r.uniform(0, 1, (n, 2))],
                         1).astype(np.float32)
    grid = L.GridSpec.from_range(rng, vs, p, mv)
    got = _trim(L.ops.voxelize(torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev), grid, want_voxels=True,
                               want_membership=True))
    ref = oracle.voxelize_batch(pts, offs, rng, vs, p, mv)
    assert (ref["pillars_per_frame"] == mv).any() and (ref["pillars_per_frame"] == 0).any()
    np.testing.assert_array_equal(got["pillar_count"][:-1], ref["pillars_per_frame"])
    np.testing.assert_array_equal(got["voxel_coords"], ref["coords"])
    np.testing.assert_array_equal(got["voxel_num_points"], ref["num_points"])
    np.testing.assert_array_equal(got["point_pillar"], ref["point_voxel"])
    np.testing.assert_array_equal(got["point_slot"], ref["point_slot"])
    np.testing.assert_array_equal(got["voxels"], ref["voxels"])


def test_fused_path_on_empty_batches(grouping, dev, L, oracle):
    """encode_bev / the module on a batch with no points at all, and on frames whose points all fall outside the range:
    zero pillars, an all-zero canvas, no error."""
    rng, vs = (-12.8, -12.8, -5.0, 12.8, 12.8, 3.0), (0.4, 0.4, 8.0)
    grid = L.GridSpec.from_range(rng, vs, 8, 100)
    sd = oracle.random_pfn_params(11, [64], True, seed=2)
    for pts, offs in ((np.zeros((0, 5), np.float32), [0, 0, 0]), (np.full((50, 5), 400.0, np.float32), [0, 20, 50])):
        res = _run_fused(L, dev, pts, np.asarray(offs, np.int32), grid, sd, 5)
        assert res["pillar_count"].cpu().tolist() == [0, 0, 0]
        assert res["bev"].shape == (2, 64, 64, 64) and not bool(res["bev"].any())
    cfg = C(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64], MAX_POINTS_PER_VOXEL=8,
            MAX_NUMBER_OF_VOXELS=100, FUSE_SCATTER=True)
    vfe = L.PillarVFEFromPoints(model_cfg=cfg, num_point_features=5, voxel_size=list(vs),
                                point_cloud_range=np.asarray(rng, np.float32), grid_size=grid.grid_size)
    vfe.load_state_dict(sd)
    vfe.eval().to(dev)
    bd = vfe({"points": torch.zeros((0, 6), device=dev), "batch_size": 2})
    assert bd["pillar_features"].shape == (0, 64) and bd["voxel_coords"].shape == (0, 4)
    assert bd["spatial_features"].shape == (2, 64, 64, 64) and not bool(bd["spatial_features"].any())


def test_grouping_unshuffled_sweeps(grouping, dev, L, oracle):
    """Sweeps in firing order (no shuffle_points): consecutive points fall into the same pillar, which is the case the
    in-warp merging of the insert kernels exists for (one atomic per run of equal cells)."""
    rng, vs, p, mv = (-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), (0.8, 0.8, 8.0), 6, 4000
    frames = [synth.make_sweep(40 + b, synth.NUSCENES_32, 5, shuffle=False) for b in range(3)]
    offs = np.zeros(4, np.int32)
    offs[1:] = np.cumsum([len(f) for f in frames])
    pts = np.concatenate(frames, 0)
    grid = L.GridSpec.from_range(rng, vs, p, mv)
    got = _trim(L.ops.voxelize(torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev), grid, want_voxels=True,
                               want_membership=True))
    ref = oracle.voxelize_batch(pts, offs, rng, vs, p, mv)
    assert ref["num_points"].max() == p
    np.testing.assert_array_equal(got["voxel_coords"], ref["coords"])
    np.testing.assert_array_equal(got["voxel_num_points"], ref["num_points"])
    np.testing.assert_array_equal(got["point_pillar"], ref["point_voxel"])
    np.testing.assert_array_equal(got["point_slot"], ref["point_slot"])
    np.testing.assert_array_equal(got["voxels"], ref["voxels"])


def test_rebase_segments_and_gatherer_single_rank(dev, L, oracle):
    """The device side of the multi-GPU hand-over on one GPU: segments of gathered coordinates (strided views of one wire
    buffer) are rebased to global frame numbers, padding rows get frame -1, an oversized count raises the overflow flag;
    the scatter of a rebased segment buffer equals the scatter of the concatenated live rows."""
    from lidar_vision_vqa_b200 import sharding

    s_n, rows, f, frames = 3, 50, 64, 2
    a, b, c = sharding.wire_layout(rows, f, frames)
    wire = torch.zeros((s_n, a + b + c), dtype=torch.uint8, device=dev)
    feats = wire[:, :a].view(torch.float32).view(s_n, rows, f)
    coords = wire[:, a:a + b].view(torch.int32).view(s_n, rows, 4)
    counts = wire[:, a + b:a + b + 4 * (frames + 1)].view(torch.int32)
    g = torch.Generator().manual_seed(0)
    live = [37, 0, 50]
    ref_rows, ref_coords = [], []
    for s in range(s_n):
        feats[s] = torch.randn(rows, f, generator=g).to(dev)
        cells = torch.randperm(32 * 32, generator=g)[:rows]
        cc = torch.stack([torch.randint(0, frames, (rows,), generator=g), torch.zeros(rows, dtype=torch.long), cells // 32,
                          cells % 32], 1).int()
        coords[s] = cc.to(dev)
        counts[s] = torch.tensor([0, 0, live[s]], dtype=torch.int32, device=dev)
        ref_rows.append(feats[s, :live[s]].cpu().numpy())
        rc = cc[:live[s]].numpy().copy()
        rc[:, 0] += s * frames
        ref_coords.append(rc)
    over = torch.zeros(1, dtype=torch.int32, device=dev)
    L.ops.rebase_segments(coords, counts, frames, overflow=over)
    got = coords.cpu().numpy()
    for s in range(s_n):
        np.testing.assert_array_equal(got[s, :live[s]], ref_coords[s])
        assert (got[s, live[s]:, 0] == -1).all()
    assert int(over.item()) == 0
    bev = sharding.densify_segments({"feats": feats, "coords": coords}, s_n * frames, 32, 32)
    ref = oracle.scatter_bev(np.concatenate(ref_rows), np.concatenate(ref_coords), 32, 32, batch_size=s_n * frames)
    np.testing.assert_array_equal(bev.cpu().numpy(), ref)
    counts[1] = torch.tensor([0, 0, rows + 1], dtype=torch.int32, device=dev)
    L.ops.rebase_segments(coords, counts, frames, overflow=over)
    assert int(over.item()) == 1
    # world size 1: the gatherer copies the slab into its own segment and rebases it
    pts, offs, rng, vs, p, mv = _make_case("small_p32")
    grid = L.GridSpec.from_range(rng, vs, 8, 500)
    bufs = L.ops.EncodeBuffers(len(pts), 2, grid, 64, dev, capacity=1100, wire_slab=True)
    gat = sharding.TokenGatherer(1100, 64, 2, dev, slots=1)
    sd = oracle.random_pfn_params(11, [64], True, seed=3)
    pfn = L.ops.fold_pfn(sd["pfn_layers.0.linear.weight"],
                         (sd["pfn_layers.0.norm.weight"], sd["pfn_layers.0.norm.bias"], sd["pfn_layers.0.norm.running_mean"],
                          sd["pfn_layers.0.norm.running_var"], 1e-3), None, c_point=5, use_absolute_xyz=True,
                         with_distance=False, voxel_size=vs, point_cloud_range=rng, device=dev)
    res = L.ops.encode_bev(torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev), grid, pfn, buffers=bufs)
    slot = gat.exchange(bufs.wire)
    torch.cuda.current_stream().wait_stream(gat.stream)
    torch.cuda.synchronize()
    gat.check()
    m = int(res["pillar_count"][-1].item())
    assert 0 < m <= 1100
    np.testing.assert_array_equal(slot["counts"][0].cpu().numpy(), res["pillar_count"].cpu().numpy())
    canvas = sharding.densify_segments(slot, 2, grid.grid_size[0], grid.grid_size[1])
    assert torch.equal(canvas, res["bev"])


@pytest.mark.gpu
def test_two_layer_stack_with_buffers_ring_and_reproducibility(dev, L):
    """NUM_FILTERS [64, 64] through pre-allocated EncodeBuffers (ops.encode_stack(buffers=...), the module's OUTPUT_RING)
    gives bit-identical rows, coordinates and canvas as the allocating call, call after call (the streaming two-layer
    kernel is order independent: fixed-point mean sums, max over points, per-point chains that do not depend on the point's
    place in the pillar's list)."""
    from lidar_vision_vqa_b200 import ops, synth

    g = load_golden("vfe_c5_2layer")
    sd = _sd_t(g)
    grid = L.GridSpec(tuple(float(v) for v in g["range"]), tuple(float(v) for v in g["voxel_size"]),
                      tuple(int(v) for v in g["grid_size"]), int(g["max_points"]), int(g["max_voxels"]))
    layers = []
    for i in range(2):
        bn = tuple(sd[f"pfn_layers.{i}.norm.{k}"] for k in ("weight", "bias", "running_mean", "running_var")) + (1e-3,)
        layers.append((sd[f"pfn_layers.{i}.linear.weight"], bn, None))
    stack = ops.fold_pfn_stack(layers, c_point=5, use_absolute_xyz=True, with_distance=False, voxel_size=grid.voxel_size,
                               point_cloud_range=grid.point_cloud_range, device=dev)
    p = torch.from_numpy(g["points"]).to(dev)
    o = torch.from_numpy(g["frame_offsets"]).to(dev)
    nb = o.numel() - 1
    ref = ops.encode_stack(p, o, grid, stack, with_bev=True)
    m = int(ref["pillar_count"][-1].item())
    np.testing.assert_allclose(ref["pillar_features"][:m].cpu().numpy(), g["out.pillar_features"], rtol=1e-3, atol=1e-5)
    bufs = ops.EncodeBuffers(p.shape[0], nb, grid, 64, dev)
    for _ in range(3):
        res = ops.encode_stack(p, o, grid, stack, with_bev=True, buffers=bufs)
        torch.cuda.synchronize()
        assert res["bev"] is bufs.bev
        for k in ("pillar_count", "bev"):
            assert torch.equal(res[k], ref[k]), k
        for k in ("pillar_features", "voxel_coords", "voxel_num_points"):
            assert torch.equal(res[k][:m], ref[k][:m]), k
    with pytest.raises(ValueError):
        ops.encode_stack(p, o, grid, stack, dynamic=True, buffers=bufs)
    # the module with OUTPUT_RING and no host synchronisation returns the same rows
    cfg = _cfg(g)
    cfg.update(MAX_POINTS_PER_VOXEL=int(g["max_points"]), MAX_NUMBER_OF_VOXELS=int(g["max_voxels"]), FUSE_SCATTER=True,
               OUTPUT_RING=2, SYNC_COUNTS=False)
    vfe = L.PillarVFEFromPoints(model_cfg=cfg, num_point_features=5, voxel_size=[float(v) for v in g["voxel_size"]],
                                point_cloud_range=g["range"], grid_size=g["grid_size"])
    vfe.load_state_dict(sd, strict=True)
    vfe.eval().to(dev)
    pb = torch.from_numpy(synth.to_pcdet_points(g["points"], g["frame_offsets"])).to(dev)
    for _ in range(3):
        bd = vfe({"points": pb, "batch_size": nb})
        torch.cuda.synchronize()
        assert torch.equal(bd["pillar_features"][:m], ref["pillar_features"][:m])
        assert torch.equal(bd["spatial_features"], ref["bev"])


@pytest.mark.gpu
@pytest.mark.parametrize("c_point", [5, 4])
def test_two_layer_huge_pillar_and_caps(c_point, dev, L, oracle):
    """NUM_FILTERS [64, 64] on the streaming kernel where its long-pillar path matters: one cell with 20 000 points next to
    ordinary pillars, P = 5 (cap binds inside the window), 32 and 100 (more kept points than one scratch batch: the
    compaction flushes), 5- and 4-channel points; grouping bit-exact, rows rtol 1e-3 vs the CPU oracle."""
    from lidar_vision_vqa_b200 import ops

    rng, vs = (0.0, 0.0, 0.0, 8.0, 8.0, 2.0), (1.0, 1.0, 2.0)
    r = np.random.default_rng(12)
    big = np.concatenate([r.uniform(3.0, 4.0, (20000, 2)), r.uniform(0, 2, (20000, 1)), r.uniform(0, 255, (20000, 1)),
                          r.uniform(0, 0.5, (20000, 1))], 1).astype(np.float32)
    mid = np.concatenate([r.uniform(5.0, 6.0, (70, 2)), r.uniform(0, 2, (70, 1)), r.uniform(0, 255, (70, 1)),
                          r.uniform(0, 0.5, (70, 1))], 1).astype(np.float32)  # a 70-point pillar: indices stay in registers
    rest = np.concatenate([r.uniform(-1, 9, (6000, 2)), r.uniform(-0.5, 2.5, (6000, 1)), r.uniform(0, 255, (6000, 1)),
                           r.uniform(0, 0.5, (6000, 1))], 1).astype(np.float32)
    pts = np.ascontiguousarray(np.concatenate([big, mid, rest])[r.permutation(26070)][:, :c_point])
    offs = np.array([0, 9000, 9000, 26070], np.int32)
    sd = oracle.random_pfn_params(c_point + 6, [64, 64], True, seed=14)
    layers = []
    for i in range(2):
        bn = tuple(torch.as_tensor(sd[f"pfn_layers.{i}.norm.{k}"]) for k in ("weight", "bias", "running_mean", "running_var")) + (1e-3,)
        layers.append((torch.as_tensor(sd[f"pfn_layers.{i}.linear.weight"]), bn, None))
    for p_max in (5, 32, 100):
        grid = L.GridSpec.from_range(rng, vs, p_max, 60)
        stack = ops.fold_pfn_stack(layers, c_point=c_point, use_absolute_xyz=True, with_distance=False,
                                   voxel_size=grid.voxel_size, point_cloud_range=grid.point_cloud_range, device=dev)
        res = ops.encode_stack(torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev), grid, stack, with_bev=True)
        m = int(res["pillar_count"][-1].item())
        ref_v = oracle.voxelize_batch(pts, offs, rng, vs, p_max, 60)
        assert m == ref_v["coords"].shape[0]
        np.testing.assert_array_equal(res["voxel_coords"][:m].cpu().numpy(), ref_v["coords"])
        np.testing.assert_array_equal(res["voxel_num_points"][:m].cpu().numpy(), ref_v["num_points"])
        ref_f = oracle.pillar_vfe(ref_v["voxels"], ref_v["num_points"], ref_v["coords"], sd, vs, rng).numpy()
        got_f = res["pillar_features"][:m].cpu().numpy()
        np.testing.assert_allclose(got_f, ref_f, rtol=FEAT_RTOL, atol=FEAT_ATOL)
        nx, ny, _ = grid.grid_size
        np.testing.assert_array_equal(res["bev"].cpu().numpy(), oracle.scatter_bev(got_f, ref_v["coords"], nx, ny, batch_size=3))


@pytest.mark.gpu
@pytest.mark.parametrize("wl,filters", [("cfg2_nuscenes32_b16_pillar0.2_bev512", [64]),
                                        ("cfg2_nuscenes32_b16_pillar0.2_bev512", [64, 64]),
                                        ("cfg3_10sweep_p32_b8", [64, 64])])
def test_dynamic_vfe_full_size_vs_oracle(wl, filters, dev, L, oracle):
    """DynamicPillarVFE (dynamic_pillar_vfe.py:14-142) at BASELINE's batch sizes through the streaming kernels' dynamic
    variant: rows in sorted-key order and uncapped counts bit-exact, features rtol 1e-3 vs the CPU restatement; cfg3 has
    ~18 k pillars of more than 32 points (none is capped in this variant), and a few points are pushed far outside z."""
    from lidar_vision_vqa_b200 import synth

    model, gc, nb = synth.WORKLOADS[wl]
    pts, offs = synth.make_batch(nb, model, 5)
    pts = pts.copy()
    pts[::997, 2] += 40.0  # z is neither range checked nor part of the cell (dynamic_pillar_vfe.py:93-96)
    pb = synth.to_pcdet_points(pts, offs)
    rng, vs = gc.point_cloud_range, gc.voxel_size
    sd = oracle.random_pfn_params(11, filters, True, seed=8)
    vfe = L.DynamicPillarVFE(model_cfg=C(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=filters),
                             num_point_features=5, voxel_size=list(vs), grid_size=gc.grid_size,
                             point_cloud_range=np.asarray(rng, np.float32))
    vfe.load_state_dict(sd)
    vfe.eval().to(dev)
    bd = vfe({"points": torch.from_numpy(pb).to(dev), "batch_size": nb})
    ref_f, ref_c, ref_n = oracle.dynamic_pillar_vfe(pb, sd, vs, rng)
    np.testing.assert_array_equal(bd["voxel_coords"].cpu().numpy(), ref_c)
    np.testing.assert_array_equal(bd["voxel_num_points"].cpu().numpy(), ref_n)
    # rtol 1e-3 plus an absolute term of 2e-6 of the row's own scale (at least the usual 1e-5): with intensities up to 255
    # and the shifted points a row holds outputs of 30-40 next to outputs near zero, and fp32 summation order alone moves
    # the latter by ~1e-5 (observed: one element of 16 M off by 1.35e-5 at a value of 0.0034 in a row reaching 26)
    got, ref = bd["pillar_features"].cpu().numpy(), ref_f.numpy()
    tol = 1e-3 * np.abs(ref) + np.maximum(1e-5, 2e-6 * np.abs(ref).max(axis=1, keepdims=True))
    bad = np.abs(got - ref) > tol
    assert not bad.any(), (int(bad.sum()), float(np.abs(got - ref)[bad].max()))
    # twice the same rows, bit for bit
    bd2 = vfe({"points": torch.from_numpy(pb).to(dev), "batch_size": nb})
    assert torch.equal(bd2["pillar_features"], bd["pillar_features"])


@pytest.mark.gpu
@pytest.mark.parametrize("filters", [[64, 64], [48, 40]])
def test_two_layer_stack_emits_the_index_map(filters, dev, L, oracle):
    """EMIT_INDEX_MAP with a two-layer PFN (streaming two-layer kernel and the general kernel): ``bev_index_map[b, y, x]`` is
    the row of the pillar in that cell and -1 elsewhere, also where max_voxels dropped pillars -- the canvas-free input of
    the tokeniser and of the backbone's first layer."""
    from lidar_vision_vqa_b200 import synth

    rng, vs, p, mv = (-12.8, -12.8, -5.0, 12.8, 12.8, 3.0), (0.4, 0.4, 8.0), 8, 900
    frames = [synth.make_sweep(900 + i, synth.NUSCENES_32, 5)[:6000] for i in range(2)]
    offs = np.zeros(3, np.int32)
    offs[1:] = np.cumsum([len(f) for f in frames])
    pts = np.concatenate(frames, 0)
    sd = oracle.random_pfn_params(11, filters, True, seed=31)
    grid_size = oracle.grid_size_of(rng, vs)
    cfg = C(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=filters, MAX_POINTS_PER_VOXEL=p,
            MAX_NUMBER_OF_VOXELS=mv, FUSE_SCATTER=False, EMIT_INDEX_MAP=True)
    vfe = L.PillarVFEFromPoints(model_cfg=cfg, num_point_features=5, voxel_size=list(vs),
                                point_cloud_range=np.asarray(rng, np.float32), grid_size=grid_size)
    vfe.load_state_dict(sd)
    vfe.eval().to(dev)
    bd = vfe({"points": torch.from_numpy(synth.to_pcdet_points(pts, offs)).to(dev), "batch_size": 2})
    ref = oracle.voxelize_batch(pts, offs, rng, vs, p, mv)
    assert (ref["pillars_per_frame"] == mv).all(), "max_voxels is meant to bind"
    np.testing.assert_array_equal(bd["voxel_coords"].cpu().numpy(), ref["coords"])
    nx, ny = int(grid_size[0]), int(grid_size[1])
    want = np.full((2, ny, nx), -1, np.int32)
    c = ref["coords"]
    want[c[:, 0], c[:, 2], c[:, 3]] = np.arange(c.shape[0], dtype=np.int32)
    np.testing.assert_array_equal(bd["bev_index_map"].cpu().numpy(), want)
    ref_f = oracle.pillar_vfe(ref["voxels"], ref["num_points"], ref["coords"], sd, vs, rng).numpy()
    np.testing.assert_allclose(bd["pillar_features"].cpu().numpy(), ref_f, rtol=1e-3, atol=1e-5)


@pytest.mark.gpu
def test_padded_voxels_with_nonzero_padding_and_stray_points(dev, L, oracle):
    """PillarVFE.forward on padded voxels sums ALL P slots for the mean, whatever the padding holds (pillar_vfe.py:97), and
    takes whatever points it is given.  The folded kernel's exact integer sums apply to clean voxeliser output only; pillars
    with non-zero padding, with a "valid" point far outside their cell, or with a float num_points tensor must take its
    float path and still match the reference arithmetic (CPU oracle), next to clean pillars in the same launch."""
    from lidar_vision_vqa_b200 import ops, synth

    rng, vs, p, mv = (-12.8, -12.8, -5.0, 12.8, 12.8, 3.0), (0.4, 0.4, 8.0), 12, 5000
    pts = synth.make_sweep(77, synth.NUSCENES_32, 5)[:9000]
    offs = np.array([0, len(pts)], np.int32)
    ref = oracle.voxelize_batch(pts, offs, rng, vs, p, mv)
    voxels, npts, coords = ref["voxels"].copy(), ref["num_points"].copy(), ref["coords"].copy()
    m = voxels.shape[0]
    r = np.random.default_rng(3)
    dirty = r.choice(m, m // 7, replace=False)           # garbage in the padding slots
    for g in dirty:
        if npts[g] < p:
            voxels[g, npts[g]:, :] = r.normal(0, 3.0, (p - npts[g], 5)).astype(np.float32)
    stray = r.choice(m, m // 11, replace=False)           # a valid point 7 m away from its pillar
    voxels[stray, 0, :3] += np.float32(7.0)
    sd = oracle.random_pfn_params(11, [64], True, seed=41)
    want = oracle.pillar_vfe(voxels, npts, coords, sd, vs, rng).numpy()
    pfn = ops.fold_pfn(sd["pfn_layers.0.linear.weight"], (sd["pfn_layers.0.norm.weight"], sd["pfn_layers.0.norm.bias"],
                       sd["pfn_layers.0.norm.running_mean"], sd["pfn_layers.0.norm.running_var"], 1e-3), None, c_point=5,
                       use_absolute_xyz=True, with_distance=False, voxel_size=vs, point_cloud_range=rng, device=dev)
    assert pfn.folded is not None
    v_d, c_d = torch.from_numpy(voxels).to(dev), torch.from_numpy(coords).to(dev)
    for n_t in (torch.from_numpy(npts).to(dev), torch.from_numpy(npts.astype(np.float32)).to(dev)):
        got = ops.pfn_dense(v_d, n_t, c_d, pfn, vs).cpu().numpy()
        scale = np.maximum(1.0, np.abs(want).max(axis=1, keepdims=True))
        assert (np.abs(got - want) <= 1e-3 * np.abs(want) + 1e-5 * scale).all()
    ops.force_generic_features(True)
    try:
        faithful = ops.pfn_dense(v_d, torch.from_numpy(npts).to(dev), c_d, pfn, vs).cpu().numpy()
    finally:
        ops.force_generic_features(False)
    assert (np.abs(faithful - want) <= 1e-3 * np.abs(want) + 1e-5 * np.maximum(1.0, np.abs(want).max(axis=1, keepdims=True))).all()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", list(range(8)))
def test_streaming_variants_randomised(seed, dev, L, oracle):
    """Random grids, caps, frame layouts and clustered clouds through the three late additions to the streaming kernel --
    hard [64, 64], dynamic [64], dynamic [64, 64] -- against the CPU oracle: pillars of exactly 32 / 33 points, pillars
    that straddle chunk boundaries, caps of 1 ... 40 points, max_voxels binding or not, empty frames."""
    from lidar_vision_vqa_b200 import synth

    r = np.random.default_rng(1000 + seed)
    vs_xy = float(r.choice([0.2, 0.4, 0.8, 1.6]))
    half = vs_xy * int(r.choice([16, 32, 64]))
    rng, vs = (-half, -half, -5.0, half, half, 3.0), (vs_xy, vs_xy, 8.0)
    p = int(r.choice([1, 2, 5, 20, 32, 40]))
    n_frames = int(r.integers(1, 5))
    frames = []
    for f in range(n_frames):
        if n_frames > 1 and r.random() < 0.2:
            frames.append(np.zeros((0, 5), np.float32))
            continue
        base = synth.make_sweep(3000 + 10 * seed + f, synth.NUSCENES_32, 5)[: int(r.integers(2000, 12000))]
        base[:, :2] *= np.float32(half / 51.2)
        blobs = []
        for cnt in (32, 33, 31, 64, 65, int(r.integers(100, 700))):  # pillars of chosen sizes inside one cell each
            cx, cy = (np.floor(r.uniform(-half, half - vs_xy, 2) / vs_xy) + 0.5) * vs_xy
            b = np.concatenate([r.uniform(-0.45, 0.45, (cnt, 2)) * vs_xy + [cx, cy], r.uniform(-4.5, 2.5, (cnt, 1)),
                                r.uniform(0, 255, (cnt, 1)), r.uniform(0, 0.5, (cnt, 1))], 1)
            blobs.append(b.astype(np.float32))
        fr = np.concatenate([base] + blobs, 0)
        frames.append(fr[r.permutation(len(fr))])
    offs = np.zeros(n_frames + 1, np.int32)
    offs[1:] = np.cumsum([len(f) for f in frames])
    pts = np.concatenate(frames, 0)
    pb = synth.to_pcdet_points(pts, offs)
    grid_size = oracle.grid_size_of(rng, vs)
    ref0 = oracle.voxelize_batch(pts, offs, rng, vs, p, 10 ** 6)
    per_frame = int(max(1, ref0["pillars_per_frame"].max()))
    mv = per_frame + 5 if r.random() < 0.5 else max(1, per_frame // 2)  # not binding / binding

    def tol_ok(got, ref):
        scale = np.maximum(1.0, np.abs(ref).max(axis=1, keepdims=True)) if ref.size else 1.0
        return (np.abs(got - ref) <= 1e-3 * np.abs(ref) + 1e-5 * scale).all()

    # hard semantics, two layers
    sd2 = oracle.random_pfn_params(11, [64, 64], True, seed=seed)
    cfg = C(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64, 64], MAX_POINTS_PER_VOXEL=p,
            MAX_NUMBER_OF_VOXELS=mv, FUSE_SCATTER=True)
    vfe = L.PillarVFEFromPoints(model_cfg=cfg, num_point_features=5, voxel_size=list(vs),
                                point_cloud_range=np.asarray(rng, np.float32), grid_size=grid_size)
    vfe.load_state_dict(sd2)
    vfe.eval().to(dev)
    bd = vfe({"points": torch.from_numpy(pb).to(dev), "batch_size": n_frames})
    ref = oracle.voxelize_batch(pts, offs, rng, vs, p, mv)
    np.testing.assert_array_equal(bd["voxel_coords"].cpu().numpy(), ref["coords"])
    np.testing.assert_array_equal(bd["voxel_num_points"].cpu().numpy(), ref["num_points"])
    ref_f = oracle.pillar_vfe(ref["voxels"], ref["num_points"], ref["coords"], sd2, vs, rng).numpy().reshape(-1, 64)
    got_f = bd["pillar_features"].cpu().numpy().reshape(-1, 64)
    assert tol_ok(got_f, ref_f)
    np.testing.assert_array_equal(bd["spatial_features"].cpu().numpy(),
                                  oracle.scatter_bev(got_f, ref["coords"], int(grid_size[0]), int(grid_size[1]), batch_size=n_frames))
    # dynamic semantics, one and two layers
    for filters in ([64], [64, 64]):
        sd = oracle.random_pfn_params(11, filters, True, seed=seed + 50)
        dyn = L.DynamicPillarVFE(model_cfg=C(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=filters),
                                 num_point_features=5, voxel_size=list(vs), grid_size=grid_size,
                                 point_cloud_range=np.asarray(rng, np.float32))
        dyn.load_state_dict(sd)
        dyn.eval().to(dev)
        bd = dyn({"points": torch.from_numpy(pb).to(dev), "batch_size": n_frames})
        ref_f, ref_c, ref_n = oracle.dynamic_pillar_vfe(pb, sd, vs, rng)
        np.testing.assert_array_equal(bd["voxel_coords"].cpu().numpy(), ref_c)
        np.testing.assert_array_equal(bd["voxel_num_points"].cpu().numpy(), ref_n)
        assert tol_ok(bd["pillar_features"].cpu().numpy().reshape(-1, 64), ref_f.numpy().reshape(-1, 64))
