"""Regenerates the golden vectors from the REAL rasterizer, when it exists, and diffs the oracle against them.

The reference's arithmetic for this path lives in ashawkey/diff-gaussian-rasterization (imported at
/root/reference/core/gs.py:7-10, install recipe /root/reference/readme.md:13-15): CUDA-only, absent from /root/reference
and from the offline wheelhouse, so `oracle/` restates SURVEY.md Appendix A and PARITY IS UNPINNED.  This script is the
ready hook for the day the package is importable on a GPU box (pip-installed, or placed under baseline/_ref/):

    python tests/golden/make_golden_from_ref.py            # needs a CUDA device and the package

It (1) renders the three golden scenes of make_golden.py through the package's `_C.rasterize_gaussians` /
`rasterize_gaussians_backward`, reading sorted keys / values / ranges / n_contrib out of its arenas with the layout of
SURVEY.md Appendix A.7 (baseline/ref_rasterizer.arena_fields), (2) writes them as tests/golden/ref_<name>.npz, and
(3) diffs the fp32 oracle field by field: bitwise on radii / xy / depth / conic / keys / values / ranges / n_contrib,
max-abs on images and gradients.  Where the package disagrees with Appendix A the package wins: fix the oracle and
splat_math.cuh, regenerate tests/golden/*.npz with make_golden.py, and drop "PARITY UNPINNED" from the headers.
Without the package it says so and exits 0 (tests/test_gpu_baseline.py::test_product_vs_real_package_if_installed skips).

The call shape follows the package's own Python wrapper as /root/reference/core/gs.py:58-85 drives it [EXT, recalled]:
  num_rendered, color, depth, alpha, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(
      bg, means3D, colors_precomp, opacities, scales, rotations, scale_modifier, cov3Ds_precomp, viewmatrix, projmatrix,
      tanfovx, tanfovy, image_height, image_width, sh, sh_degree, campos, prefiltered, debug)            # 19 arguments
  grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
  grad_rotations = _C.rasterize_gaussians_backward(
      bg, means3D, radii, colors_precomp, scales, rotations, scale_modifier, cov3Ds_precomp, viewmatrix, projmatrix,
      tanfovx, tanfovy, grad_out_color, grad_out_depth, grad_out_alpha, sh, sh_degree, campos, geomBuffer, num_rendered,
      binningBuffer, imgBuffer, alpha, debug)                                                            # 24 arguments
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)


def real_rasterize_view(real, means3D, colors_precomp, opacities, scales, rotations, rs):
    """One forward through the real package; same dict as baseline.ref_rasterizer.rasterize_view."""
    import torch
    empty = torch.Tensor([]).to(means3D.device)
    n, color, depth, alpha, radii, geom, binning, img = real._C.rasterize_gaussians(
        rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, empty, rs.viewmatrix, rs.projmatrix,
        rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, empty, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
    return dict(color=color, depth=depth, alpha=alpha, radii=radii, num_rendered=int(n), geom=geom, binning=binning, img=img)


def real_backward_view(real, means3D, colors_precomp, scales, rotations, rs, fwd, d_color, d_depth, d_alpha):
    import torch
    empty = torch.Tensor([]).to(means3D.device)
    g = real._C.rasterize_gaussians_backward(
        rs.bg, means3D, fwd["radii"], colors_precomp, scales, rotations, rs.scale_modifier, empty, rs.viewmatrix, rs.projmatrix,
        rs.tanfovx, rs.tanfovy, d_color, d_depth, d_alpha, empty, rs.sh_degree, rs.campos, fwd["geom"], fwd["num_rendered"],
        fwd["binning"], fwd["img"], fwd["alpha"], rs.debug)
    names = ("dL_dmean2D", "dL_dcolor", "dL_dopacity", "dL_dmeans", "dL_dcov3d", "dL_dsh", "dL_dscales", "dL_drots")
    return dict(zip(names, g))


def main():
    import torch
    from baseline import ref_rasterizer
    real = ref_rasterizer.real_package()
    if real is None:
        print("diff_gaussian_rasterization (with its _C extension) is not importable: nothing regenerated, parity stays unpinned")
        return 0
    if not torch.cuda.is_available():
        print("the real rasterizer is CUDA-only: run this on a GPU box")
        return 0
    from lgm_b200 import GaussianRasterizationSettings
    dev = torch.device("cuda:0")
    worst = {}
    for name in ("hand_placed_7.npz", "random_300_72x40.npz", "init_500_64x64.npz"):
        z = np.load(os.path.join(HERE, name))
        c = {k: z[k] for k in z.files}
        W, H, P = int(c["W"]), int(c["H"]), len(c["radii"])
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
        rs = GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=float(c["tanfovx"]), tanfovy=float(c["tanfovy"]),
                                           bg=t(c["bg"]), scale_modifier=1.0, viewmatrix=t(c["view"]).reshape(4, 4),
                                           projmatrix=t(c["proj"]).reshape(4, 4), sh_degree=0, campos=torch.zeros(3, device=dev),
                                           prefiltered=False, debug=False)
        m3, col, op, sc, ro = t(c["means"]), t(c["cols"]), t(c["opac"]).reshape(P, 1), t(c["scales"]), t(c["rots"])
        fwd = real_rasterize_view(real, m3, col, op, sc, ro, rs)
        ntiles = ((W + 15) // 16) * ((H + 15) // 16)
        fld = ref_rasterizer.arena_fields(P, fwd["num_rendered"], W * H, fwd["geom"], fwd["binning"], fwd["img"], ntiles)
        bwd = real_backward_view(real, m3, col, sc, ro, rs, fwd, t(c["d_img"]), t(c["d_depth"]).reshape(1, H, W), t(c["d_alpha"]).reshape(1, H, W))
        ref = dict(radii=fwd["radii"].cpu().numpy(), xy=fld["means2D"].cpu().numpy(), conic_opacity=fld["conic_opacity"].cpu().numpy(),
                   keys=fld["keys"].cpu().numpy().view(np.uint64), vals=fld["vals"].cpu().numpy().view(np.uint32),
                   ranges=fld["ranges"].cpu().numpy().view(np.uint32), n_contrib=fld["n_contrib"].cpu().numpy().view(np.uint32).reshape(H, W),
                   image=fwd["color"].cpu().numpy(), alpha=fwd["alpha"].cpu().numpy(), depth_img=fwd["depth"].cpu().numpy(),
                   dL_dmeans=bwd["dL_dmeans"].cpu().numpy(), dL_dscales=bwd["dL_dscales"].cpu().numpy(),
                   dL_drots=bwd["dL_drots"].cpu().numpy(), dL_dopacity=bwd["dL_dopacity"].cpu().numpy().reshape(-1),
                   dL_dcolor=bwd["dL_dcolor"].cpu().numpy(), dL_dmean2D=bwd["dL_dmean2D"].cpu().numpy()[:, :2])
        np.savez_compressed(os.path.join(HERE, "ref_" + name), **{**c, **ref})
        print(f"== {name}: real package vs oracle-made golden")
        vis = ref["radii"] > 0
        for k in ("radii", "keys", "vals", "ranges", "n_contrib"):
            same = np.array_equal(np.asarray(ref[k]).reshape(-1), np.asarray(c[k]).reshape(-1))
            print(f"   {k:14s} {'bit-identical' if same else 'DIFFERS'}")
            worst[k] = worst.get(k, True) and same
        for k in ("xy", "conic_opacity"):  # the package leaves culled rows uninitialised: compare the visible ones
            same = np.array_equal(ref[k][vis].view(np.uint32), np.ascontiguousarray(c[k], np.float32)[vis].view(np.uint32))
            print(f"   {k:14s} {'bit-identical' if same else 'DIFFERS'} (visible rows)")
            worst[k] = worst.get(k, True) and same
        for k in ("image", "alpha", "depth_img", "dL_dmeans", "dL_dscales", "dL_drots", "dL_dopacity", "dL_dcolor", "dL_dmean2D"):
            a, b = np.asarray(ref[k], np.float64).reshape(-1), np.asarray(c[k], np.float64).reshape(-1)
            print(f"   {k:14s} max-abs diff {np.abs(a - b).max():.3e} (scale {np.abs(b).max():.3e})")
    print("bitwise fields all identical:", all(worst.values()))
    return 0


if __name__ == "__main__":
    sys.exit(main())
