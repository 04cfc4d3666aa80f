"""CPU tests of the oracle itself: golden fixtures, structural properties, finite-difference gradients (fp64).
PARITY UNPINNED — see oracle/splat_oracle.c: the fixtures are self-made (tests/golden/make_golden.py)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import GOLDEN_FILES, load_golden, split14, tan_half
from lgm_b200.cameras import orbit_views
from lgm_b200.synthetic import make_gaussians, make_cameras


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", GOLDEN_FILES)
def test_golden_forward_bit_exact(oracle32, name):
    c = load_golden(name)
    W, H = int(c["W"]), int(c["H"])
    pre, b, f = oracle32.rasterize(c["means"], c["scales"], c["rots"], c["opac"], c["cols"], c["view"], c["proj"],
                                   c["bg"], W, H, float(c["tanfovx"]), float(c["tanfovy"]))
    assert np.array_equal(pre["radii"], c["radii"])
    assert np.array_equal(pre["tiles"], c["tiles"])
    assert np.array_equal(pre["rects"], c["rects"])
    assert np.array_equal(_bits(pre["xy"]), _bits(c["xy"]))
    assert np.array_equal(_bits(pre["depth"]), _bits(c["depth"]))
    assert np.array_equal(_bits(pre["conic_opacity"]), _bits(c["conic_opacity"]))
    assert np.array_equal(b["keys"], c["keys"]) and np.array_equal(b["vals"], c["vals"])
    assert np.array_equal(b["ranges"], c["ranges"])
    assert np.array_equal(f["n_contrib"], c["n_contrib"])
    # same binary, same libm: bit-exact; tolerance only guards a different glibc expf on another box
    np.testing.assert_allclose(f["image"], c["image"], atol=2e-6)
    np.testing.assert_allclose(f["alpha"], c["alpha"], atol=2e-6)
    np.testing.assert_allclose(f["depth"], c["depth_img"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("name", GOLDEN_FILES)
def test_golden_backward(oracle32, name):
    c = load_golden(name)
    W, H = int(c["W"]), int(c["H"])
    args = (c["means"], c["scales"], c["rots"], c["opac"], c["cols"], c["view"], c["proj"], c["bg"], W, H,
            float(c["tanfovx"]), float(c["tanfovy"]))
    pre, b, f = oracle32.rasterize(*args)
    g = oracle32.rasterize_backward(*args, pre, b, f, c["d_img"], c["d_alpha"], c["d_depth"])
    for k in ("dL_dmeans", "dL_dscales", "dL_drots", "dL_dopacity", "dL_dcolor", "dL_dmean2D", "dL_dconic"):
        scale = np.abs(c[k]).max() + 1e-20
        assert np.abs(g[k] - c[k]).max() <= 1e-5 * scale, k


def test_hand_placed_semantics():
    c = load_golden("hand_placed_7.npz")
    assert c["radii"][0] == 0 and c["tiles"][0] == 0            # behind the camera: culled
    assert np.all(c["radii"][1:] > 0)
    assert _bits(c["depth"][3]) == _bits(c["depth"][4])          # identical-depth pair ...
    k, v = c["keys"], c["vals"]
    same = np.where((k[1:] == k[:-1]))[0]
    assert len(same) > 0 and np.all(v[same] < v[same + 1])       # ... ties resolved by ascending Gaussian index
    assert c["alpha"].max() > 0.98                               # the opaque splat saturates pixels
    assert np.all(c["alpha"] <= 1.0 + 1e-6)


def _check_structure(pre, b, W, H):
    L = b["L"]
    assert int(pre["tiles"].sum()) == L
    keys, vals, ranges = b["keys"], b["vals"], b["ranges"]
    assert np.all(keys[1:] >= keys[:-1])
    # stable: equal keys keep ascending Gaussian index
    eq = np.where(keys[1:] == keys[:-1])[0]
    assert np.all(vals[eq] < vals[eq + 1])
    # ranges partition [0, L) in tile order; empty tiles are (0,0)
    ntiles = ((W + 15) // 16) * ((H + 15) // 16)
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    cnt = np.bincount(tiles, minlength=ntiles)
    nz = np.where(cnt > 0)[0]
    starts = np.concatenate([[0], np.cumsum(cnt[nz])[:-1]]) if len(nz) else np.zeros(0, np.int64)
    assert np.array_equal(ranges[nz, 0], starts) and np.array_equal(ranges[nz, 1] - ranges[nz, 0], cnt[nz])
    assert np.all(ranges[cnt == 0] == 0)
    # same multiset before / after the sort
    assert np.array_equal(np.sort(b["unsorted_keys"]), keys)


@pytest.mark.parametrize("kind,W,H", [("trained", 256, 256), ("init", 96, 80), ("trained", 40, 200)])
def test_binning_structure(oracle32, kind, W, H):
    g = make_gaussians(1, 3000, kind, seed=3)[0].numpy()
    cv, cvp, _ = make_cameras(1, 1, seed=9)
    means, opac, scales, rots, cols = split14(g)
    t = tan_half(49.1)
    pre = oracle32.preprocess(means, scales, rots, opac, cv[0, 0].numpy(), cvp[0, 0].numpy(), W, H, t * W / H, t)
    b = oracle32.bin(pre, W, H)
    _check_structure(pre, b, W, H)
    f = oracle32.composite_fwd(pre, b, cols, np.ones(3), W, H)
    assert np.all(f["alpha"] <= 1.0 + 1e-5) and np.all(f["alpha"] >= 0)
    assert np.all(f["n_contrib"].reshape(-1) <= (b["ranges"][:, 1] - b["ranges"][:, 0]).max())


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), P=st.integers(0, 64), W=st.integers(1, 70), H=st.integers(1, 70),
       big=st.booleans())
def test_property_random_small(seed, P, W, H, big):
    from oracle.oracle import Oracle
    o = Oracle("f32")
    rng = np.random.RandomState(seed)
    cv, cvp, _ = orbit_views(1, 1.5, 49.1, 0.5, 2.5, seed=seed)
    means = rng.uniform(-0.8, 0.8, (P, 3)).astype(np.float32)
    scales = np.exp(rng.uniform(-5, -1 if big else -3, (P, 3))).astype(np.float32)
    rots = rng.randn(P, 4).astype(np.float32)
    opac = rng.uniform(0, 1, P).astype(np.float32)
    cols = rng.uniform(0, 1, (P, 3)).astype(np.float32)
    t = tan_half(49.1)
    pre, b, f = o.rasterize(means, scales, rots, opac, cols, cv[0].numpy(), cvp[0].numpy(), [0.5, 0.5, 0.5], W, H,
                            t * W / H, t)
    _check_structure(pre, b, W, H)
    assert np.isfinite(f["image"]).all() and np.all(f["alpha"] <= 1 + 1e-5)
    # empty scene / fully culled -> pure background, alpha 0
    if b["L"] == 0:
        assert np.allclose(f["image"], 0.5) and np.all(f["alpha"] == 0)


def test_empty_and_behind(oracle32):
    cv, cvp, _ = orbit_views(1, 1.5, 49.1, 0.5, 2.5, seed=0)
    t = tan_half(49.1)
    z = np.zeros
    pre, b, f = oracle32.rasterize(z((0, 3)), z((0, 3)), z((0, 4)), z(0), z((0, 3)), cv[0].numpy(), cvp[0].numpy(),
                                   [1, 0, 0], 20, 20, t, t)
    assert b["L"] == 0 and np.all(f["image"][0] == 1) and np.all(f["image"][1:] == 0)
    # everything far behind the camera
    means = np.tile(cv[0].numpy()[3, :3] * 0 + 100.0, (5, 1)).astype(np.float32)
    vis = oracle32.mark_visible(means, cv[0].numpy())
    pre = oracle32.preprocess(means, np.full((5, 3), 0.01), np.tile([1, 0, 0, 0], (5, 1)), np.ones(5), cv[0].numpy(),
                              cvp[0].numpy(), 20, 20, t, t)
    assert np.array_equal(pre["radii"] > 0, vis & (pre["tiles"] > 0))


def test_f32_matches_f64(oracle32, oracle64):
    g = make_gaussians(1, 2000, "trained", seed=5)[0].numpy()
    g[:, 4:7] *= 3
    cv, cvp, _ = make_cameras(1, 1, seed=2)
    means, opac, scales, rots, cols = split14(g)
    t = tan_half(49.1)
    a = oracle32.rasterize(means, scales, rots, opac, cols, cv[0, 0].numpy(), cvp[0, 0].numpy(), [1, 1, 1], 128, 128, t, t)
    b = oracle64.rasterize(means, scales, rots, opac, cols, cv[0, 0].numpy(), cvp[0, 0].numpy(), [1, 1, 1], 128, 128, t, t)
    assert (a[0]["radii"] != b[0]["radii"]).mean() < 1e-3
    # threshold-chaotic (alpha < 1/255 skip): a few pixels may differ by one dropped contribution
    err = np.abs(a[2]["image"] - b[2]["image"])
    assert np.quantile(err, 0.999) < 1e-4 and err.max() < 2.0 / 255.0


def _fd_check(o, arrs, name, key, loss, analytic, eps=1e-6, tol=2e-4):
    base = arrs[name]
    bad = 0
    for idx in np.ndindex(base.shape):
        ap, am = dict(arrs), dict(arrs)
        ap[name] = base.copy(); ap[name][idx] += eps
        am[name] = base.copy(); am[name][idx] -= eps
        num = (loss(**ap) - loss(**am)) / (2 * eps)
        ana = analytic[key][idx]
        if abs(num - ana) > tol * max(1e-4, abs(num), abs(ana)):
            bad += 1
    return bad


def test_backward_matches_finite_differences(oracle64):
    """The fp64 oracle's backward (A.5 + A.6) is the gradient of its forward (A.1 + A.4), incl. depth & alpha."""
    o = oracle64
    rng = np.random.RandomState(0)
    P, W, H = 16, 40, 24
    t = tan_half(49.1)
    tanx = t * W / H
    cv, cvp, _ = orbit_views(1, 1.5, 49.1, 0.5, 2.5, seed=3)
    view, proj = cv[0].numpy().astype(np.float64).ravel(), cvp[0].numpy().astype(np.float64).ravel()
    rots = rng.randn(P, 4)
    rots = rots / np.linalg.norm(rots, axis=1, keepdims=True) * rng.uniform(0.8, 1.2, (P, 1))  # unnormalised
    arrs = dict(means=rng.uniform(-0.4, 0.4, (P, 3)), scales=np.exp(rng.uniform(-3.5, -2.0, (P, 3))), rots=rots,
                opac=rng.uniform(0.2, 0.9, P), cols=rng.uniform(0, 1, (P, 3)))
    bg = np.array([0.3, 0.6, 0.9])
    wi, wa, wd = rng.randn(3, H, W), rng.randn(H, W), rng.randn(H, W)

    def run(means, scales, rots, opac, cols):
        return o.rasterize(means, scales, rots, opac, cols, view, proj, bg, W, H, tanx, t)

    def loss(**a):
        _, _, f = run(**a)
        return (f["image"] * wi).sum() + (f["alpha"][0] * wa).sum() + (f["depth"][0] * wd).sum()

    pre, b, f = run(**arrs)
    assert b["L"] > 20 and f["n_contrib"].max() >= 3
    g = o.rasterize_backward(arrs["means"], arrs["scales"], arrs["rots"], arrs["opac"], arrs["cols"], view, proj, bg,
                             W, H, tanx, t, pre, b, f, wi, wa, wd)
    assert np.abs(g["dL_dmeans"]).max() > 0 and np.abs(g["dL_drots"]).max() > 0
    total_bad = 0
    for name, key in (("means", "dL_dmeans"), ("scales", "dL_dscales"), ("rots", "dL_drots"),
                      ("opac", "dL_dopacity"), ("cols", "dL_dcolor")):
        total_bad += _fd_check(o, arrs, name, key, loss, g)
    n = sum(v.size for v in arrs.values())
    assert total_bad <= 0.02 * n  # a finite difference may straddle a skip threshold


def test_render_step_matches_single_view_loop(oracle32):
    """orc_render_step (OpenMP over views, grads summed per scene) == loop of single-view calls."""
    B, V, N, S = 2, 3, 400, 48
    g = make_gaussians(B, N, "trained", seed=1).numpy()
    g[:, :, 4:7] *= 5
    cv, cvp, _ = make_cameras(B, V, seed=4)
    t = tan_half(49.1)
    rng = np.random.RandomState(1)
    dimg = rng.randn(B, V, 3, S, S).astype(np.float32)
    dal = rng.randn(B, V, 1, S, S).astype(np.float32)
    ddp = rng.randn(B, V, 1, S, S).astype(np.float32)
    bg = np.array([0.1, 0.2, 0.3], np.float32)
    r = oracle32.render_step(g, cv.numpy(), cvp.numpy(), bg, S, S, t, t, 1.0, dimg, dal, ddp)
    dg = np.zeros_like(g)
    for b in range(B):
        means, opac, scales, rots, cols = split14(g[b])
        for v in range(V):
            a = (means, scales, rots, opac, cols, cv[b, v].numpy(), cvp[b, v].numpy(), bg, S, S, t, t)
            pre, bn, f = oracle32.rasterize(*a)
            assert np.array_equal(f["image"], r["image"][b, v]) and np.array_equal(pre["radii"], r["radii"][b, v])
            gr = oracle32.rasterize_backward(*a, pre, bn, f, dimg[b, v], dal[b, v, 0], ddp[b, v, 0])
            dg[b, :, 0:3] += gr["dL_dmeans"]; dg[b, :, 3] += gr["dL_dopacity"]; dg[b, :, 4:7] += gr["dL_dscales"]
            dg[b, :, 7:11] += gr["dL_drots"]; dg[b, :, 11:14] += gr["dL_dcolor"]
    np.testing.assert_allclose(r["dgaussians"], dg, rtol=1e-4, atol=1e-5 * np.abs(dg).max())


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_sh_backward_matches_finite_differences(oracle64, deg):
    """orc_sh_backward (basis-table form) against central differences of orc_sh_forward in fp64."""
    rng = np.random.RandomState(deg)
    P, M = 5, 16
    means, campos, shs = rng.randn(P, 3), np.array([0.3, -0.2, 2.0]), rng.randn(P, M, 3) * 0.5
    w = rng.randn(P, 3)
    col, cl = oracle64.sh_forward(means, campos, shs, deg)
    dsh, dm = oracle64.sh_backward(means, campos, shs, deg, cl, w)
    assert (col >= 0).all() and ((col == 0) == (cl == 1)).all()
    assert (dsh[:, (deg + 1) ** 2:] == 0).all()

    def loss(m, s):
        return float((oracle64.sh_forward(m, campos, s, deg)[0] * w).sum())

    eps = 1e-6
    for idx in np.ndindex(means.shape):
        a, b = means.copy(), means.copy()
        a[idx] += eps
        b[idx] -= eps
        num = (loss(a, shs) - loss(b, shs)) / (2 * eps)
        assert abs(num - dm[idx]) <= 1e-6 * max(1.0, abs(num))
    for idx in list(np.ndindex(shs.shape))[::5]:
        a, b = shs.copy(), shs.copy()
        a[idx] += eps
        b[idx] -= eps
        num = (loss(means, a) - loss(means, b)) / (2 * eps)
        assert abs(num - dsh[idx]) <= 1e-6 * max(1.0, abs(num))


def test_order_claim_behind_direct_binning(oracle32):
    """What the CUDA direct binning path relies on (direct_bin.cu): the stable sort of the emitted (tile | depth) keys
    leaves, inside every tile, the instances in ascending (depth bits, value) order — so ANY sort of the unique pairs
    (depth bits << 32 | value) per tile reproduces upstream's list, in whatever order a tile's instances were collected.
    Checked on the oracle's own emit + sort, with every Gaussian duplicated (exact depth ties inside every tile)."""
    rng = np.random.RandomState(3)
    W, H, P = 112, 80, 1500
    g = make_gaussians(1, P, "trained", seed=13)[0].numpy()
    g[:, 4:7] *= 6.0
    g = np.concatenate([g, g, g[: P // 2]], 0)  # duplicates: identical depth bits
    cv, cvp, _ = make_cameras(1, 1, seed=13)
    means, opac, scales, rots, _ = split14(g)
    t = tan_half(49.1)
    pre = oracle32.preprocess(means, scales, rots, opac, cv[0, 0].numpy(), cvp[0, 0].numpy(), W, H, t * W / H, t, 1.0)
    b = oracle32.bin(pre, W, H)
    assert b["L"] > 5000
    tile = (b["unsorted_keys"] >> np.uint64(32)).astype(np.int64)
    dbits = (b["unsorted_keys"] & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    val = b["unsorted_vals"].astype(np.uint64)
    # collect per tile in a scrambled order (what the atomics of the scatter kernel do), then sort the unique pairs
    perm = rng.permutation(b["L"])
    tile, pair = tile[perm], (dbits[perm] << np.uint64(32)) | val[perm]
    order = np.lexsort((pair, tile))
    assert np.array_equal(b["keys"], (tile[order].astype(np.uint64) << np.uint64(32)) | (pair[order] >> np.uint64(32)))
    assert np.array_equal(b["vals"], (pair[order] & np.uint64(0xFFFFFFFF)).astype(np.uint32))
    k = b["keys"]
    assert (k[1:] == k[:-1]).mean() > 0.3  # the ties are real
    # and the ranges are the exclusive scan of the per-tile counts
    counts = np.bincount(tile, minlength=len(b["ranges"]))
    start = np.concatenate([[0], np.cumsum(counts)[:-1]])
    ne = counts > 0
    assert np.array_equal(b["ranges"][ne, 0], start[ne]) and np.array_equal(b["ranges"][ne, 1], (start + counts)[ne])
    assert (b["ranges"][~ne] == 0).all()
