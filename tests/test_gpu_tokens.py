"""GPU parity tests (``-m gpu``, B200) of the BEV tokeniser (csrc/tokens.cu through the C ABI) against the golden vectors
made by the reference's own ``VATLiDAR.forward`` and against the CPU oracle (oracle/tokens_oracle.py).

Bar (BASELINE.json north_star): BEV tokens within rtol 1e-3 in fp32; observed ~1e-6, asserted much tighter below.
"""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, load_golden
from lidar_vision_vqa_b200 import synth
from oracle import tokens_oracle as to

pytestmark = pytest.mark.gpu

TOK_RTOL, TOK_ATOL = 1e-3, 1e-4      # north_star tolerance
TIGHT_RTOL, TIGHT_ATOL = 2e-5, 3e-5  # what the fp32 FMA path actually achieves


def token_golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "tok_*.npz")))


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def make_tokenizer(sd, dev, projection="auto"):
    from lidar_vision_vqa_b200 import tokens as T

    c, d = sd["refine.0.bias"].shape[0], sd["proj.bias"].shape[0]
    tk = T.VATLiDARTokenizer(c_in=c, d_model=d, projection=projection)
    tk.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return tk.eval().to(dev)


def rows_of(bev):
    """(pillar_features, voxel_coords, pillar_count) holding exactly the non-zero cells of a canvas, frame by frame."""
    b, c, h, w = bev.shape
    occ = np.abs(bev).max(1) > 0
    bi, yi, xi = np.nonzero(occ)
    feats = bev[bi, :, yi, xi].astype(np.float32)
    coords = np.stack([bi, np.zeros_like(bi), yi, xi], 1).astype(np.int32)
    return feats, coords


@pytest.mark.parametrize("projection", ["umma", "fma"])
@pytest.mark.parametrize("name", token_golden_names())
def test_dense_forward_matches_reference_golden(dev, name, projection):
    """Both projection variants: 3-term TF32 split on the tensor cores (where instantiated) and fp32 on the FMA pipes."""
    g = load_golden(name)
    tk = make_tokenizer(g["state_dict"], dev, projection)
    out = tk(torch.from_numpy(g["bev"]).to(dev)).cpu().numpy()
    assert out.shape == g["out.tokens"].shape
    np.testing.assert_allclose(out, g["out.tokens"], rtol=TOK_RTOL, atol=TOK_ATOL)
    np.testing.assert_allclose(out, g["out.tokens"], rtol=TIGHT_RTOL, atol=TIGHT_ATOL)


@pytest.mark.parametrize("name", token_golden_names())
def test_tables_match_oracle(dev, name):
    g = load_golden(name)
    sd = g["state_dict"]
    _, _, h, w = g["bev"].shape
    tk = make_tokenizer(sd, dev)
    pe, bg = tk.tables(h, w)
    np.testing.assert_allclose(pe.cpu().numpy(), to.positional_table(sd, h, w), rtol=TIGHT_RTOL, atol=TIGHT_ATOL)
    np.testing.assert_allclose(bg.cpu().numpy(), to.background_token(sd), rtol=TIGHT_RTOL, atol=TIGHT_ATOL)


@pytest.mark.parametrize("name", token_golden_names())
@pytest.mark.parametrize("coords_float", [False, True])
def test_pillar_rows_give_the_same_tokens_as_the_canvas(dev, name, coords_float):
    """forward_pillars (no canvas) == forward(canvas) bit for bit, and both match the reference."""
    g = load_golden(name)
    tk = make_tokenizer(g["state_dict"], dev)
    bev = g["bev"]
    b, _, h, w = bev.shape
    feats, coords = rows_of(bev)
    perm = np.random.default_rng(0).permutation(len(feats))  # row order is storage only
    ct = torch.from_numpy(coords[perm]).to(dev)
    dense = tk(torch.from_numpy(bev).to(dev))
    rows = tk.forward_pillars(torch.from_numpy(feats[perm]).to(dev), ct.float() if coords_float else ct, b, (h, w))
    assert torch.equal(dense, rows)
    np.testing.assert_allclose(rows.cpu().numpy(), g["out.tokens"], rtol=TIGHT_RTOL, atol=TIGHT_ATOL)


def test_live_row_count_and_empty_and_full_canvases(dev):
    sd = to.random_token_params(16, 128, seed=5)
    tk = make_tokenizer(sd, dev)
    rng = np.random.default_rng(1)
    h, w, b = 10, 37, 3
    # empty canvas: every token is background + PE
    out = tk(torch.zeros(b, 16, h, w, device=dev)).cpu().numpy()
    expect = to.background_token(sd)[None] + to.positional_table(sd, h, w)
    for bi in range(b):
        np.testing.assert_allclose(out[bi], expect, rtol=TIGHT_RTOL, atol=TIGHT_ATOL)
    # no zero anywhere: every cell takes the arithmetic path
    bev = (rng.standard_normal((b, 16, h, w)) + 3.0).astype(np.float32)
    out = tk(torch.from_numpy(bev).to(dev)).cpu().numpy()
    np.testing.assert_allclose(out, to.bev_tokens(bev, sd), rtol=TIGHT_RTOL, atol=TIGHT_ATOL)
    # rows beyond the live count (device scalar) are ignored
    occ = rng.random((b, 1, h, w)) < 0.2
    bev = np.where(occ, np.maximum(rng.standard_normal((b, 16, h, w)), 0), 0).astype(np.float32)
    feats, coords = rows_of(bev)
    m = len(feats)
    junk_f = rng.standard_normal((50, 16)).astype(np.float32)
    junk_c = np.stack([np.zeros(50), np.zeros(50), rng.integers(0, h, 50), rng.integers(0, w, 50)], 1).astype(np.int32)
    count = torch.tensor([0, 0, 0, m], dtype=torch.int32, device=dev)
    out = tk.forward_pillars(torch.from_numpy(np.concatenate([feats, junk_f])).to(dev),
                             torch.from_numpy(np.concatenate([coords, junk_c])).to(dev), b, (h, w), pillar_count=count)
    np.testing.assert_allclose(out.cpu().numpy(), to.bev_tokens(bev, sd), rtol=TIGHT_RTOL, atol=TIGHT_ATOL)
    # zero frames / zero rows
    assert tk(torch.zeros(0, 16, h, w, device=dev)).shape == (0, h * w, 128)
    out = tk.forward_pillars(torch.zeros(0, 16, device=dev), torch.zeros(0, 4, dtype=torch.int32, device=dev), 2, (h, w))
    np.testing.assert_allclose(out[1].cpu().numpy(), expect, rtol=TIGHT_RTOL, atol=TIGHT_ATOL)


@pytest.mark.parametrize("c,d,h,w", [(4, 128, 5, 70), (128, 512, 9, 33), (256, 768, 6, 40), (64, 896, 12, 12),
                                     (512, 1024, 4, 34), (64, 640, 3, 65)])
def test_shapes_of_the_reference_test_matrix(dev, c, d, h, w):
    """(c_in, d_model) pairs of the reference's own test sweep (training-test/models/test_vat_lidar.py:249-253: 256/512,
    128/256, 512/768) plus the product's d_model = 896 and the instantiation limits."""
    sd = to.random_token_params(c, d, seed=c + d)
    tk = make_tokenizer(sd, dev)
    rng = np.random.default_rng(c)
    occ = rng.random((2, 1, h, w)) < 0.25
    bev = np.where(occ, rng.standard_normal((2, c, h, w)), 0).astype(np.float32)
    out = tk(torch.from_numpy(bev).to(dev)).cpu().numpy()
    np.testing.assert_allclose(out, to.bev_tokens(bev, sd), rtol=TIGHT_RTOL, atol=1e-4)


@pytest.mark.parametrize("projection", ["umma", "fma"])
def test_frame_chunks_empty_and_full_canvases_on_both_projections(dev, projection):
    """Edge cases of the tile machinery: more than 16 frames (index-map windows are staged 16 frames at a time), a pair count
    that is not a multiple of the 128-row tile, no active pair at all, every pair active; c_in = 32 is a shape the tcgen05
    variant is instantiated for."""
    sd = to.random_token_params(32, 128, seed=9)
    tk = make_tokenizer(sd, dev, projection)
    assert tk.projection == projection
    rng = np.random.default_rng(4)
    h, w, b = 12, 44, 19
    occ = rng.random((b, 1, h, w)) < 0.1
    bev = np.where(occ, np.maximum(rng.standard_normal((b, 32, h, w)), 0), 0).astype(np.float32)
    ref = to.bev_tokens(bev, sd)
    out = tk(torch.from_numpy(bev).to(dev))
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=TIGHT_RTOL, atol=TIGHT_ATOL)
    feats, coords = rows_of(bev)
    rows = tk.forward_pillars(torch.from_numpy(feats).to(dev), torch.from_numpy(coords).to(dev), b, (h, w))
    assert torch.equal(out, rows)
    # nothing active: only the streamed background
    expect = to.background_token(sd)[None] + to.positional_table(sd, h, w)
    out0 = tk(torch.zeros(2, 32, h, w, device=dev)).cpu().numpy()
    np.testing.assert_allclose(out0[1], expect, rtol=TIGHT_RTOL, atol=TIGHT_ATOL)
    # everything active: 2 * 12 * 44 = 1056 pairs = 8 full tiles + 32 rows
    dense = (rng.standard_normal((2, 32, h, w)) + 2.5).astype(np.float32)
    np.testing.assert_allclose(tk(torch.from_numpy(dense).to(dev)).cpu().numpy(), to.bev_tokens(dense, sd),
                               rtol=TIGHT_RTOL, atol=TIGHT_ATOL)


def test_unsupported_shapes_fail_loudly(dev):
    from lidar_vision_vqa_b200 import NativeLibraryError
    from lidar_vision_vqa_b200 import tokens as T

    tk = T.VATLiDARTokenizer(c_in=8, d_model=96).eval().to(dev)
    with pytest.raises(NativeLibraryError):
        tk(torch.zeros(1, 8, 4, 4, device=dev))
    tk = T.VATLiDARTokenizer(c_in=8, d_model=128).to(dev)  # training mode
    with pytest.raises(RuntimeError):
        tk(torch.zeros(1, 8, 4, 4, device=dev))


def test_cfg2_canvas_tokens_from_the_fused_encoder(dev):
    """Full-size path: synthetic nuScenes sweeps -> pillars_encode_bev -> tokens from the index map the encoder left behind.
    Checked against (a) the dense entry point on the encoder's own canvas, bit for bit, (b) the oracle on frame 0,
    (c) the background property on every cell whose window is empty, (d) a second run (determinism)."""
    import lidar_vision_vqa_b200 as L
    from lidar_vision_vqa_b200 import ops
    from lidar_vision_vqa_b200 import tokens as T
    from oracle import pillar_oracle as po

    model, gc, _ = synth.WORKLOADS["cfg2_nuscenes32_b16_pillar0.2_bev512"]
    grid = L.GridSpec.from_range(gc.point_cloud_range, gc.voxel_size, gc.max_points_per_voxel, gc.max_voxels)
    nb = 2
    pts, offs = synth.make_batch(nb, model, 5)
    sdp = po.random_pfn_params(11, [64], True, seed=0)
    pfn = ops.fold_pfn(sdp["pfn_layers.0.linear.weight"], (sdp["pfn_layers.0.norm.weight"], sdp["pfn_layers.0.norm.bias"],
                       sdp["pfn_layers.0.norm.running_mean"], sdp["pfn_layers.0.norm.running_var"], 1e-3), None,
                       c_point=5, use_absolute_xyz=True, with_distance=False, voxel_size=grid.voxel_size,
                       point_cloud_range=grid.point_cloud_range, device=dev)
    p, o = torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev)
    bufs = ops.EncodeBuffers(len(pts), nb, grid, 64, dev)
    res = ops.encode_bev(p, o, grid, pfn, buffers=bufs, want_index_map=True)
    sd = to.random_token_params(64, 128, seed=11)
    tk = make_tokenizer(sd, dev)
    cell_row = res["cell_row"]
    assert torch.equal(cell_row, T.encode_index_map(bufs, len(pts), nb, grid))
    m = int(res["pillar_count"][-1].item())
    assert int((cell_row >= 0).sum().item()) == m
    tok_map = tk.forward_index_map(res["pillar_features"], cell_row)
    for other in ("fma",):  # tcgen05 split (default for this shape) vs fp32 FFMA2
        tok_o = make_tokenizer(sd, dev, projection=other).forward_index_map(res["pillar_features"], cell_row)
        assert float((tok_map - tok_o).abs().max()) < 3e-5, other
    tok_dense = tk(res["bev"])
    tok_rows = tk.forward_pillars(res["pillar_features"], res["voxel_coords"], nb, (512, 512), pillar_count=res["pillar_count"])
    assert torch.equal(tok_map, tok_dense) and torch.equal(tok_map, tok_rows)
    assert torch.equal(tok_map, tk.forward_index_map(res["pillar_features"], cell_row))
    bev0 = res["bev"][:1].cpu().numpy()
    ref0 = to.bev_tokens(bev0, sd)
    np.testing.assert_allclose(tok_map[:1].cpu().numpy(), ref0, rtol=TIGHT_RTOL, atol=1e-4)
    occ = (cell_row >= 0).float()[:, None]
    near = torch.nn.functional.max_pool2d(occ, 3, stride=1, padding=1)[:, 0] > 0
    pe, bg = tk.tables(512, 512)
    expect = (pe + bg[None]).view(512, 512, -1)
    for b in range(nb):
        far = ~near[b]
        assert 0.5 < float(far.float().mean()) < 0.99
        assert torch.equal(tok_map[b].view(512, 512, -1)[far], expect[far])
