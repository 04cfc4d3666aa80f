"""Python front end of baseline/ref_rasterizer.cu — the REFERENCE-SHAPED CUDA rasterizer.  TEST / MEASUREMENT
INFRASTRUCTURE ONLY: nothing under lgm_b200/ imports this module.

It restates, in the external package's own host shape, what `diff_gaussian_rasterization` does around its `_C`
extension (SURVEY.md §3 call stack A, §8a rows a4 / a11):

  * `_RasterizeGaussians.forward`: one `rasterize_gaussians` call per VIEW; the three arenas are byte tensors resized
    through callbacks; eleven tensors saved for the backward;
  * `_RasterizeGaussians.backward`: one `rasterize_gaussians_backward` call per view, ten freshly zero-filled
    per-view gradient tensors (dL_dmeans3D, dL_dmeans2D, dL_dcolors, dL_ddepths, dL_dconic, dL_dopacity, dL_dcov3D,
    dL_dsh, dL_dscales, dL_drotations); autograd then sums the V per-view gradients of each scene;
  * `render_loop`: the B x V Python loop of /root/reference/core/gs.py:42-93 (5 slice + contiguous().float() per scene,
    settings NamedTuple + nn.Module + zeros_like(means2D) per view, clamp, stack, view).

If the REAL package is importable (`import diff_gaussian_rasterization` resolving to something with a `_C` extension,
e.g. installed under baseline/_ref/), `real_package()` returns it and tests/test_gpu_baseline.py compares against it
as well; offline it never is.
"""
import ctypes
import os
import subprocess
import sys

import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libref_rasterizer.so")
SRC = os.path.join(_HERE, "ref_rasterizer.cu")
_lib = None
_RESIZE = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)


def build(force=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a (cross-compiles without a GPU); CUB ships with the toolkit."""
    deps = [SRC, os.path.join(ROOT, "lgm_b200", "csrc", "splat_math.cuh")]
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(d) for d in deps):
        nvcc = next((c for c in (os.environ.get("CUDA_HOME", "") + "/bin/nvcc", "/usr/local/cuda/bin/nvcc") if os.path.exists(c)), "nvcc")
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
                               "-Xcompiler", "-fPIC", "-o", LIB_PATH, SRC])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = ctypes.CDLL(LIB_PATH)
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        l.ref_forward.restype = ci
        l.ref_forward.argtypes = [_RESIZE, vp, _RESIZE, vp, _RESIZE, vp, ci, vp, ci, ci, vp, vp, vp, vp, cf, vp, vp, vp, cf, cf,
                                  vp, vp, vp, vp]
        l.ref_backward.restype = ci
        l.ref_backward.argtypes = [ci, ci, vp, ci, ci, vp, vp, vp, vp, cf, vp, vp, vp, cf, cf, vp, vp, vp, vp, vp, vp, vp] + [vp] * 8
        l.ref_last_error.restype = ctypes.c_char_p
        l.ref_arena_offsets.restype = None
        l.ref_arena_offsets.argtypes = [ci, ci, ci, ctypes.POINTER(ctypes.c_size_t)]
        _lib = l
    return _lib


def real_package():
    """The real `diff_gaussian_rasterization` (with its compiled `_C`) when one is installed — e.g. under baseline/_ref/ —
    else None.  The in-repo import-name shim (diff_gaussian_rasterization/ at the repo root) does not count."""
    ref_dir = os.path.join(_HERE, "_ref")
    added = False
    if os.path.isdir(ref_dir) and ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
        added = True
    try:
        import importlib
        for name in list(sys.modules):
            if name == "diff_gaussian_rasterization" and not hasattr(sys.modules[name], "_C"):
                del sys.modules[name]
        mod = importlib.import_module("diff_gaussian_rasterization")
        return mod if hasattr(mod, "_C") else None
    except Exception:
        return None
    finally:
        if added:
            sys.path.remove(ref_dir)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _resizer(t):
    """The package's resizeFunctional: grow the byte tensor, hand back its data pointer."""
    def fn(_ctx, n):
        t.resize_(int(n))
        return t.data_ptr()
    return _RESIZE(fn)


def rasterize_view(means3D, colors_precomp, opacities, scales, rotations, rs):
    """One `rasterize_gaussians` call: outputs + the three arenas (geomBuffer, binningBuffer, imgBuffer)."""
    L = lib()
    dev = means3D.device
    P, H, W = means3D.shape[0], int(rs.image_height), int(rs.image_width)
    # the four outputs, as RasterizeGaussiansCUDA allocates them (torch::full)
    out_color = torch.full((3, H, W), 0.0, dtype=torch.float32, device=dev)
    out_depth = torch.full((1, H, W), 0.0, dtype=torch.float32, device=dev)
    out_alpha = torch.full((1, H, W), 0.0, dtype=torch.float32, device=dev)
    radii = torch.full((P,), 0, dtype=torch.int32, device=dev)
    geom, binning, img = (torch.empty(0, dtype=torch.uint8, device=dev) for _ in range(3))
    cbs = [_resizer(geom), _resizer(binning), _resizer(img)]
    bg, vm, pm = rs.bg.contiguous(), rs.viewmatrix.contiguous(), rs.projmatrix.contiguous()
    n = L.ref_forward(cbs[0], None, cbs[1], None, cbs[2], None, P, _p(bg), W, H, _p(means3D), _p(colors_precomp),
                      _p(opacities), _p(scales), float(rs.scale_modifier), _p(rotations), _p(vm), _p(pm), float(rs.tanfovx),
                      float(rs.tanfovy), _p(out_color), _p(out_depth), _p(out_alpha), _p(radii))
    if n < 0:
        raise RuntimeError(f"ref_forward failed: {L.ref_last_error().decode()}")
    return dict(color=out_color, depth=out_depth, alpha=out_alpha, radii=radii, num_rendered=n, geom=geom, binning=binning,
                img=img, bg=bg, vm=vm, pm=pm)


class _RefRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, colors_precomp, opacities, scales, rotations, rs):
        o = rasterize_view(means3D, colors_precomp, opacities, scales, rotations, rs)
        ctx.rs, ctx.num_rendered = rs, o["num_rendered"]
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, o["radii"], o["geom"], o["binning"], o["img"], o["alpha"],
                              o["bg"], o["vm"], o["pm"])
        ctx.mark_non_differentiable(o["radii"])
        return o["color"], o["radii"], o["depth"], o["alpha"]

    @staticmethod
    def backward(ctx, grad_color, _grad_radii, grad_depth, grad_alpha):
        L = lib()
        colors_precomp, means3D, scales, rotations, radii, geom, binning, img, out_alpha, bg, vm, pm = ctx.saved_tensors
        rs = ctx.rs
        dev = means3D.device
        P, H, W = means3D.shape[0], int(rs.image_height), int(rs.image_width)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        # the ten zero-filled per-view gradient tensors of RasterizeGaussiansBackwardCUDA
        dL_dmeans3D, dL_dmeans2D, dL_dcolors, dL_ddepths = z(P, 3), z(P, 3), z(P, 3), z(P, 1)
        dL_dconic, dL_dopacity, dL_dcov3D, dL_dsh = z(P, 2, 2), z(P, 1), z(P, 6), z(P, 0, 3)
        dL_dscales, dL_drotations = z(P, 3), z(P, 4)
        gc = grad_color.contiguous()
        gd = (z(1, H, W) if grad_depth is None else grad_depth.contiguous())
        ga = (z(1, H, W) if grad_alpha is None else grad_alpha.contiguous())
        rc = L.ref_backward(P, ctx.num_rendered, _p(bg), W, H, _p(means3D), _p(colors_precomp), _p(out_alpha), _p(scales),
                            float(rs.scale_modifier), _p(rotations), _p(vm), _p(pm), float(rs.tanfovx), float(rs.tanfovy), _p(radii),
                            _p(geom), _p(binning), _p(img), _p(gc), _p(gd), _p(ga), _p(dL_dmeans2D), _p(dL_dconic), _p(dL_dopacity),
                            _p(dL_dcolors), _p(dL_ddepths), _p(dL_dmeans3D), _p(dL_dscales), _p(dL_drotations))
        if rc < 0:
            raise RuntimeError(f"ref_backward failed: {L.ref_last_error().decode()}")
        return dL_dmeans3D, dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dscales, dL_drotations, None


class RefGaussianRasterizer(nn.Module):
    """GaussianRasterizer of the external package (colors_precomp + scales / rotations path, the one LGM uses)."""

    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None):
        if shs is not None or cov3D_precomp is not None or colors_precomp is None or scales is None or rotations is None:
            raise Exception("the reference-shaped baseline serves LGM's call: colors_precomp + scales + rotations")
        return _RefRasterize.apply(means3D, means2D, colors_precomp, opacities, scales, rotations, self.raster_settings)


def render_loop(gaussians, cam_view, cam_view_proj, cam_pos, bg_color, output_size, tan_half_fov, rasterizer_cls=None,
                settings_cls=None, scale_modifier=1.0):
    """/root/reference/core/gs.py:42-93 restated: the B x V double loop, one rasterizer call per view."""
    from lgm_b200.rasterizer import GaussianRasterizationSettings
    rasterizer_cls = rasterizer_cls or RefGaussianRasterizer
    settings_cls = settings_cls or GaussianRasterizationSettings
    B, V = cam_view.shape[:2]
    images, alphas = [], []
    for b in range(B):
        means3D = gaussians[b, :, 0:3].contiguous().float()
        opacity = gaussians[b, :, 3:4].contiguous().float()
        scales = gaussians[b, :, 4:7].contiguous().float()
        rotations = gaussians[b, :, 7:11].contiguous().float()
        rgbs = gaussians[b, :, 11:].contiguous().float()
        for v in range(V):
            view_matrix = cam_view[b, v].float()
            view_proj_matrix = cam_view_proj[b, v].float()
            campos = cam_pos[b, v].float()
            raster_settings = settings_cls(
                image_height=output_size, image_width=output_size, tanfovx=tan_half_fov, tanfovy=tan_half_fov, bg=bg_color,
                scale_modifier=scale_modifier, viewmatrix=view_matrix, projmatrix=view_proj_matrix, sh_degree=0, campos=campos,
                prefiltered=False, debug=False)
            rasterizer = rasterizer_cls(raster_settings=raster_settings)
            rendered_image, radii, rendered_depth, rendered_alpha = rasterizer(
                means3D=means3D, means2D=torch.zeros_like(means3D, dtype=torch.float32, device=means3D.device), shs=None,
                colors_precomp=rgbs, opacities=opacity, scales=scales, rotations=rotations, cov3D_precomp=None)
            images.append(rendered_image.clamp(0, 1))
            alphas.append(rendered_alpha)
    images = torch.stack(images, dim=0).view(B, V, 3, output_size, output_size)
    alphas = torch.stack(alphas, dim=0).view(B, V, 1, output_size, output_size)
    return {"image": images, "alpha": alphas}


def bench_leg(g_dev, cv_dev, cvp_dev, cp_dev, bg, d_img, d_alpha, S, tan_half_fov, native_value, unit, steps=2, warmup=1):
    """bench.py's `gpu_baseline`: the whole step of the headline workload through render_loop + one backward, on the
    legacy default stream, timed with CUDA events; returns a dict for the JSON line."""
    build()
    lib()
    B, V = cv_dev.shape[:2]

    def step(cls=None):
        g = g_dev.detach().requires_grad_(True)
        out = render_loop(g, cv_dev, cvp_dev, cp_dev, bg, S, tan_half_fov, rasterizer_cls=cls)
        torch.autograd.backward([out["image"].view(B * V, 3, S, S), out["alpha"].view(B * V, 1, S, S)], [d_img, d_alpha])
        return g.grad

    def time_it(cls=None):
        for _ in range(warmup):
            step(cls)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(cls)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    ms = time_it()
    value = B * V / (ms * 1e-3)
    res = {"value": value, "unit": unit, "ms_per_step": ms, "steps": steps, "kind": "reference-shaped restatement",
           "same_config": True, "native_over_baseline": native_value / value,
           "what": "baseline/ref_rasterizer.cu (per-view launches on the default stream, CUB scan + blocking 4-byte D2H + CUB "
                   "radix sort, 256-thread tiles with 256-Gaussian rounds, expf, 10 atomicAdd per contributing pair) driven by "
                   "the B x V Python loop of /root/reference/core/gs.py:42-93 with one autograd backward per view; the real "
                   "package (ashawkey/diff-gaussian-rasterization) is CUDA-only and not installable offline"}
    real = real_package()
    if real is not None:  # the actual reference, when someone has installed it under baseline/_ref
        ms_r = time_it(real.GaussianRasterizer)
        res["real_package"] = {"value": B * V / (ms_r * 1e-3), "ms_per_step": ms_r, "native_over_real": native_value / (B * V / (ms_r * 1e-3))}
    return res


def arena_fields(P, L, n_pixels, geom, binning, img, n_tiles):
    """Read sorted keys / values / ranges / n_contrib / means2D / conic_opacity out of the arenas (Appendix A.7)."""
    off = (ctypes.c_size_t * 6)()
    lib().ref_arena_offsets(P, max(L, 0), n_pixels, off)
    return dict(
        keys=_slice(binning, off[0], torch.int64, L), vals=_slice(binning, off[1], torch.int32, L),
        ranges=_slice(img, off[2], torch.int32, 2 * n_tiles).view(n_tiles, 2), n_contrib=_slice(img, off[3], torch.int32, n_pixels),
        means2D=_slice(geom, off[4], torch.float32, 2 * P).view(P, 2), conic_opacity=_slice(geom, off[5], torch.float32, 4 * P).view(P, 4))


def _slice(buf, off, dtype, count):
    """Field at byte offset `off` of an arena whose base is the (aligned) data pointer of `buf`.  ref_arena_offsets
    computed the offsets from a NULL base; torch allocations are at least 256-byte aligned, so they carry over."""
    assert buf.data_ptr() % 128 == 0
    nbytes = count * torch.empty(0, dtype=dtype).element_size()
    return buf[off:off + nbytes].view(dtype)
