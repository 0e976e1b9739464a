"""CPU tests of the BEV tokeniser restatement (oracle/tokens_oracle.py) against golden vectors produced by the reference's
own ``VATLiDAR.forward`` (tests/golden/make_golden_tokens.py) and, in the build container, against the live class."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, load_golden
from oracle import ref_loader
from oracle import tokens_oracle as to


def token_golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "tok_*.npz")))


@pytest.mark.parametrize("name", token_golden_names())
def test_tokens_restatement_matches_reference_golden(name):
    g = load_golden(name)
    out = to.bev_tokens(g["bev"], g["state_dict"])
    assert out.shape == g["out.tokens"].shape
    np.testing.assert_allclose(out, g["out.tokens"], rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("name", token_golden_names())
def test_geometry_restatement_matches_reference_golden(name):
    g = load_golden(name)
    _, _, h, w = g["bev"].shape
    geom, sid = to.grid_geometry(h, w)
    np.testing.assert_array_equal(sid, g["sid"])  # sector ids are integers: exact
    assert (sid >= 0).all()
    np.testing.assert_allclose(geom, g["geom"], rtol=0, atol=1e-6)
    np.testing.assert_array_equal(geom[:, :2], g["geom"][:, :2])  # the linspace coordinates bit for bit


@pytest.mark.parametrize("name", ["tok_c64_d256_24x20_sparse", "tok_c32_d128_13x9_odd"])
def test_empty_window_cells_hold_background_plus_pe(name):
    """The property the sparse-aware kernel relies on: a cell whose zero-padded 3x3 window is all zeros yields
    LN(proj(GELU(refine.bias))) + PE(cell), whatever the rest of the canvas holds -- checked on the REFERENCE tokens."""
    g = load_golden(name)
    sd = g["state_dict"]
    b, c, h, w = g["bev"].shape
    occ = np.abs(g["bev"]).max(1) > 0
    pad = np.zeros((b, h + 2, w + 2), bool)
    pad[:, 1:-1, 1:-1] = occ
    near = np.zeros_like(occ)
    for dy in range(3):
        for dx in range(3):
            near |= pad[:, dy:dy + h, dx:dx + w]
    assert 0 < near.sum() < near.size
    expect = to.background_token(sd)[None, :] + to.positional_table(sd, h, w)
    tok = g["out.tokens"].reshape(b, h, w, -1)
    for bi in range(b):
        far = ~near[bi]
        np.testing.assert_allclose(tok[bi][far], expect.reshape(h, w, -1)[far], rtol=1e-5, atol=2e-5)


@pytest.mark.skipif(not ref_loader.vat_lidar_available(), reason="reference tree absent (GPU box)")
@pytest.mark.parametrize("c,d,h,w", [(16, 128, 50, 50), (8, 128, 33, 64), (24, 256, 7, 96)])
def test_restatement_matches_live_reference(c, d, h, w):
    import torch

    vat = ref_loader.load_vat_lidar()
    model = vat(c_in=c, d_model=d, n_queries=6, n_layers=1, n_heads=2).eval()
    sd = to.random_token_params(c, d, seed=c + h)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    rng = np.random.default_rng(h * w)
    bev = rng.standard_normal((2, c, h, w)).astype(np.float32)
    bev[:, :, rng.random((h, w)) > 0.2] = 0.0
    ref = ref_loader.vat_lidar_kv_tokens(model, torch.from_numpy(bev)).numpy()
    np.testing.assert_allclose(to.bev_tokens(bev, sd), ref, rtol=1e-5, atol=2e-5)
    geom, sid = to.grid_geometry(h, w)
    g_ref, s_ref = model._grid(h, w, torch.device("cpu"))
    np.testing.assert_array_equal(sid, s_ref.numpy())
    np.testing.assert_allclose(geom, g_ref.numpy(), rtol=0, atol=1e-6)
