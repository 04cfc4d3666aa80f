"""Generates tests/golden/*.npz from the CPU oracle (oracle/splat_oracle.c, fp32 bit-pinned build).

PARITY UNPINNED: the reference's rasterizer (ashawkey/diff-gaussian-rasterization) cannot run here and the
reference holds no golden vectors for this path, so these are self-made pins: they freeze the
oracle's behaviour (guarding it against regressions) and give the GPU tests fixed known-answer cases.

    python tests/golden/make_golden.py
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.oracle import Oracle  # noqa: E402
from lgm_b200.cameras import orbit_views  # noqa: E402
from lgm_b200.synthetic import make_gaussians  # noqa: E402


def hand_placed_scene():
    """7 Gaussians on a 32x32 image: one behind the camera, one on a tile edge, one nearly
    degenerate, two with identical depth, one saturating a pixel, one ordinary."""
    cv, cvp, _ = orbit_views(1, 1.5, 49.1, 0.5, 2.5, seed=11, elevation_range=(0.0, 0.0))
    view, proj = cv[0].numpy().ravel(), cvp[0].numpy().ravel()
    cam = -view.reshape(4, 4)[3, :3] @ np.linalg.inv(view.reshape(4, 4)[:3, :3])  # camera centre (row-vector conv.)
    fwd = view.reshape(4, 4)[:3, 2]  # world direction of +z_view
    right, up = view.reshape(4, 4)[:3, 0], view.reshape(4, 4)[:3, 1]
    at = lambda d, x=0.0, y=0.0: cam + fwd * d + right * x + up * y
    means = np.stack([
        at(-0.5),             # behind the camera -> culled
        at(1.5, 0.0, 0.0),    # centre: projects onto the tile corner (15.5, 15.5)
        at(1.4, 0.2, 0.1),    # nearly degenerate (tiny, one long axis)
        at(1.6, -0.3, 0.2),   # identical depth pair
        at(1.6, -0.25, 0.25), # identical depth pair
        at(1.2, 0.3, -0.3),   # opaque and large: saturates pixels
        at(1.8, -0.2, -0.2),  # ordinary
    ]).astype(np.float32)
    scales = np.array([[0.05] * 3, [0.05, 0.04, 0.03], [1e-4, 0.2, 1e-4], [0.06] * 3, [0.06] * 3, [0.25] * 3,
                       [0.08, 0.03, 0.05]], np.float32)
    rots = np.array([[1, 0, 0, 0], [0.9, 0.1, 0.3, 0.2], [0.7, 0.7, 0, 0], [1, 0, 0, 0], [1, 0, 0, 0], [1, 0, 0, 0],
                     [0.5, 0.5, 0.5, 0.5]], np.float32)
    opac = np.array([0.9, 0.8, 0.7, 0.6, 0.6, 0.999, 0.5], np.float32)
    cols = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [0, 1, 1], [1, 0, 1], [0.5, 0.5, 0.5]], np.float32)
    return dict(means=means, scales=scales, rots=rots, opac=opac, cols=cols, view=view, proj=proj,
                bg=np.array([0.2, 0.4, 0.6], np.float32), W=32, H=32, fovy=49.1)


def run_case(o, c, grads_seed=0):
    tan = math.tan(0.5 * math.radians(c["fovy"]))
    W, H = c["W"], c["H"]
    tanx = tan * W / H
    pre, binned, fwd = o.rasterize(c["means"], c["scales"], c["rots"], c["opac"], c["cols"], c["view"], c["proj"],
                                   c["bg"], W, H, tanx, tan)
    rng = np.random.RandomState(grads_seed)
    d_img = rng.randn(3, H, W).astype(np.float32)
    d_alpha = rng.randn(H, W).astype(np.float32)
    d_depth = rng.randn(H, W).astype(np.float32)
    bwd = o.rasterize_backward(c["means"], c["scales"], c["rots"], c["opac"], c["cols"], c["view"], c["proj"], c["bg"],
                               W, H, tanx, tan, pre, binned, fwd, d_img, d_alpha, d_depth)
    out = dict(c)
    out.update(tanfovx=np.float32(tanx), tanfovy=np.float32(tan), d_img=d_img, d_alpha=d_alpha, d_depth=d_depth,
               depth=pre["depth"], radii=pre["radii"], xy=pre["xy"], conic_opacity=pre["conic_opacity"],
               tiles=pre["tiles"], rects=pre["rects"], keys=binned["keys"], vals=binned["vals"],
               ranges=binned["ranges"], unsorted_keys=binned["unsorted_keys"], image=fwd["image"], alpha=fwd["alpha"],
               depth_img=fwd["depth"], n_contrib=fwd["n_contrib"])
    out.update({k: v for k, v in bwd.items()})
    return out


def main():
    o = Oracle("f32")
    np.savez_compressed(os.path.join(HERE, "hand_placed_7.npz"), **run_case(o, hand_placed_scene()))
    # 2: a small random "trained-like" scene, non-square image with a ragged last tile row/column
    g = make_gaussians(1, 300, "trained", seed=77)[0].numpy()
    g[:, 4:7] *= 6.0  # larger footprints so tiles hold tens of Gaussians
    cv, cvp, _ = orbit_views(1, 1.5, 49.1, 0.5, 2.5, seed=5)
    c = dict(means=g[:, 0:3].copy(), scales=g[:, 4:7].copy(), rots=g[:, 7:11].copy(), opac=g[:, 3].copy(),
             cols=g[:, 11:14].copy(), view=cv[0].numpy().ravel(), proj=cvp[0].numpy().ravel(),
             bg=np.array([1, 1, 1], np.float32), W=72, H=40, fovy=49.1)
    np.savez_compressed(os.path.join(HERE, "random_300_72x40.npz"), **run_case(o, c, 1))
    # 3: "init-like" (large, saturating splats), 64x64
    g = make_gaussians(1, 500, "init", seed=78)[0].numpy()
    cv, cvp, _ = orbit_views(1, 1.5, 60.0, 0.5, 2.5, seed=6)
    c = dict(means=g[:, 0:3].copy(), scales=g[:, 4:7].copy(), rots=g[:, 7:11].copy(), opac=g[:, 3].copy(),
             cols=g[:, 11:14].copy(), view=cv[0].numpy().ravel(), proj=cvp[0].numpy().ravel(),
             bg=np.array([0.1, 0.9, 0.3], np.float32), W=64, H=64, fovy=60.0)
    np.savez_compressed(os.path.join(HERE, "init_500_64x64.npz"), **run_case(o, c, 2))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            z = np.load(os.path.join(HERE, f))
            print(f, "P", len(z["radii"]), "visible", int((z["radii"] > 0).sum()), "L", len(z["keys"]),
                  "max n_contrib", int(z["n_contrib"].max()), os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
