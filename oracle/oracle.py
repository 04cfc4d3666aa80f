"""ctypes front end of the CPU oracle (oracle/splat_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
module; nothing under lgm_b200/ does.  PARITY UNPINNED (see the C file's header): the reference's rasterizer
(ashawkey/diff-gaussian-rasterization, CUDA-only) is not available, so this restates SURVEY.md Appendix A.

Two builds: Oracle("f32") is bit-pinned fp32 (radii / keys / ranges compare bit-exactly with the CUDA path),
Oracle("f64") is the double-precision arbiter for gradients.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False):
    """Compile liboracle_f32.so / liboracle_f64.so with oracle/Makefile (gcc, OpenMP)."""
    outs = [os.path.join(_HERE, f"liboracle_{p}.so") for p in ("f32", "f64")]
    src = os.path.join(_HERE, "splat_oracle.c")
    if force or any((not os.path.exists(o)) or os.path.getmtime(o) < os.path.getmtime(src) for o in outs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return outs


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class Oracle:
    def __init__(self, precision="f32"):
        assert precision in ("f32", "f64")
        path = os.path.join(_HERE, f"liboracle_{precision}.so")
        if not os.path.exists(path):
            build()
        self.lib = ctypes.CDLL(path)
        self.dt = np.float32 if precision == "f32" else np.float64
        self.creal = ctypes.c_float if precision == "f32" else ctypes.c_double
        self.lib.orc_count_instances.restype = ctypes.c_int64
        self.lib.orc_render_step.restype = ctypes.c_int64
        assert self.lib.orc_real_bytes() == np.dtype(self.dt).itemsize

    def num_threads(self):
        return int(self.lib.orc_num_threads())

    def set_num_threads(self, n):
        """Override OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1 to every rank)."""
        self.lib.orc_set_num_threads(ctypes.c_int(int(n)))

    def _a(self, x, shape=None):
        a = np.ascontiguousarray(np.asarray(x, dtype=self.dt))
        if shape is not None:
            a = a.reshape(shape)
        return a

    # ---- A.1 ------------------------------------------------------------------------------------
    def preprocess(self, means, scales, rots, opac, view, proj, W, H, tanfovx, tanfovy, scale_modifier=1.0):
        P = len(means)
        means, scales, rots = self._a(means, (P, 3)), self._a(scales, (P, 3)), self._a(rots, (P, 4))
        opac, view, proj = self._a(opac, (P,)), self._a(view, (16,)), self._a(proj, (16,))
        out = dict(
            depth=np.zeros(P, self.dt), radii=np.zeros(P, np.int32), xy=np.zeros((P, 2), self.dt),
            conic_opacity=np.zeros((P, 4), self.dt), tiles=np.zeros(P, np.uint32),
            cov3d=np.zeros((P, 6), self.dt), rects=np.zeros((P, 4), np.int32))
        self.lib.orc_preprocess(
            ctypes.c_int(P), _p(means), _p(scales), _p(rots), _p(opac), self.creal(scale_modifier), _p(view),
            _p(proj), ctypes.c_int(W), ctypes.c_int(H), self.creal(tanfovx), self.creal(tanfovy), _p(out["depth"]),
            _p(out["radii"]), _p(out["xy"]), _p(out["conic_opacity"]), _p(out["tiles"]), _p(out["cov3d"]),
            _p(out["rects"]))
        return out

    def mark_visible(self, means, view):
        P = len(means)
        vis = np.zeros(P, np.uint8)
        self.lib.orc_mark_visible(ctypes.c_int(P), _p(self._a(means, (P, 3))), _p(self._a(view, (16,))), _p(vis))
        return vis.astype(bool)

    # ---- A.2 / A.3 ------------------------------------------------------------------------------
    def bin(self, pre, W, H):
        P = len(pre["radii"])
        L = int(self.lib.orc_count_instances(ctypes.c_int(P), _p(pre["tiles"])))
        ntiles = ((W + 15) // 16) * ((H + 15) // 16)
        n = max(L, 1)
        out = dict(L=L, keys=np.zeros(n, np.uint64), vals=np.zeros(n, np.uint32),
                   ranges=np.zeros((ntiles, 2), np.uint32), unsorted_keys=np.zeros(n, np.uint64),
                   unsorted_vals=np.zeros(n, np.uint32))
        self.lib.orc_bin(ctypes.c_int(P), _p(pre["radii"]), _p(pre["xy"]), _p(pre["depth"]), ctypes.c_int(W),
                         ctypes.c_int(H), ctypes.c_int64(L), _p(out["keys"]), _p(out["vals"]), _p(out["ranges"]),
                         _p(out["unsorted_keys"]), _p(out["unsorted_vals"]))
        for k in ("keys", "vals", "unsorted_keys", "unsorted_vals"):
            out[k] = out[k][:L]
        return out

    # ---- A.4 ------------------------------------------------------------------------------------
    def composite_fwd(self, pre, binned, colors, bg, W, H):
        P = len(pre["radii"])
        colors, bg = self._a(colors, (P, 3)), self._a(bg, (3,))
        out = dict(image=np.zeros((3, H, W), self.dt), alpha=np.zeros((1, H, W), self.dt),
                   depth=np.zeros((1, H, W), self.dt), n_contrib=np.zeros((H, W), np.uint32))
        vals = binned["vals"] if binned["L"] > 0 else np.zeros(1, np.uint32)
        self.lib.orc_composite_fwd(ctypes.c_int(W), ctypes.c_int(H), _p(binned["ranges"]), _p(vals), _p(pre["xy"]),
                                   _p(pre["conic_opacity"]), _p(colors), _p(pre["depth"]), _p(bg), _p(out["image"]),
                                   _p(out["alpha"]), _p(out["depth"]), _p(out["n_contrib"]))
        return out

    # ---- A.5 ------------------------------------------------------------------------------------
    def composite_bwd(self, pre, binned, colors, bg, fwd, dL_dimage, dL_dalpha, dL_ddepth, W, H):
        P = len(pre["radii"])
        colors, bg = self._a(colors, (P, 3)), self._a(bg, (3,))
        dimg, dal, ddp = self._a(dL_dimage, (3, H, W)), self._a(dL_dalpha, (H, W)), self._a(dL_ddepth, (H, W))
        out = dict(dL_dmean2D=np.zeros((P, 2), self.dt), dL_dconic=np.zeros((P, 3), self.dt),
                   dL_dopacity=np.zeros(P, self.dt), dL_dcolor=np.zeros((P, 3), self.dt),
                   dL_ddepth=np.zeros(P, self.dt))
        vals = binned["vals"] if binned["L"] > 0 else np.zeros(1, np.uint32)
        self.lib.orc_composite_bwd(
            ctypes.c_int(W), ctypes.c_int(H), _p(binned["ranges"]), _p(vals), _p(pre["xy"]), _p(pre["conic_opacity"]),
            _p(colors), _p(pre["depth"]), _p(bg), _p(self._a(fwd["alpha"])), _p(fwd["n_contrib"]), _p(dimg), _p(dal),
            _p(ddp), _p(out["dL_dmean2D"]), _p(out["dL_dconic"]), _p(out["dL_dopacity"]), _p(out["dL_dcolor"]),
            _p(out["dL_ddepth"]))
        return out

    # ---- A.6 ------------------------------------------------------------------------------------
    def preprocess_bwd(self, means, scales, rots, view, proj, W, H, tanfovx, tanfovy, radii, dL_dmean2D, dL_dconic,
                       dL_ddepth, scale_modifier=1.0):
        P = len(means)
        means, scales, rots = self._a(means, (P, 3)), self._a(scales, (P, 3)), self._a(rots, (P, 4))
        view, proj = self._a(view, (16,)), self._a(proj, (16,))
        out = dict(dL_dmeans=np.zeros((P, 3), self.dt), dL_dscales=np.zeros((P, 3), self.dt),
                   dL_drots=np.zeros((P, 4), self.dt), dL_dcov3d=np.zeros((P, 6), self.dt))
        self.lib.orc_preprocess_bwd(
            ctypes.c_int(P), _p(means), _p(scales), _p(rots), self.creal(scale_modifier), _p(view), _p(proj),
            ctypes.c_int(W), ctypes.c_int(H), self.creal(tanfovx), self.creal(tanfovy),
            _p(np.ascontiguousarray(radii, np.int32)), _p(self._a(dL_dmean2D, (P, 2))), _p(self._a(dL_dconic, (P, 3))),
            _p(self._a(dL_ddepth, (P,))), _p(out["dL_dmeans"]), _p(out["dL_dscales"]), _p(out["dL_drots"]),
            _p(out["dL_dcov3d"]))
        return out

    # ---- SH colours (the `shs` input; not on LGM's path) -------------------------------------------
    def sh_forward(self, means, campos, shs, deg):
        P, M = shs.shape[0], shs.shape[1]
        means, campos, shs = self._a(means, (P, 3)), self._a(campos, (3,)), self._a(shs, (P, M, 3))
        colors, clamped = np.zeros((P, 3), self.dt), np.zeros((P, 3), np.uint8)
        self.lib.orc_sh_forward(ctypes.c_int(P), ctypes.c_int(deg), ctypes.c_int(M), _p(means), _p(campos), _p(shs),
                                _p(colors), _p(clamped))
        return colors, clamped

    def sh_backward(self, means, campos, shs, deg, clamped, dL_dcolor):
        P, M = shs.shape[0], shs.shape[1]
        means, campos, shs = self._a(means, (P, 3)), self._a(campos, (3,)), self._a(shs, (P, M, 3))
        dsh, dmeans = np.zeros((P, M, 3), self.dt), np.zeros((P, 3), self.dt)
        self.lib.orc_sh_backward(ctypes.c_int(P), ctypes.c_int(deg), ctypes.c_int(M), _p(means), _p(campos), _p(shs),
                                 _p(np.ascontiguousarray(clamped, np.uint8)), _p(self._a(dL_dcolor, (P, 3))), _p(dsh),
                                 _p(dmeans))
        return dsh, dmeans

    # ---- single view, the GaussianRasterizer call of core/gs.py:76-85 ------------------------------
    def rasterize(self, means, scales, rots, opac, colors, view, proj, bg, W, H, tanfovx, tanfovy,
                  scale_modifier=1.0):
        pre = self.preprocess(means, scales, rots, opac, view, proj, W, H, tanfovx, tanfovy, scale_modifier)
        binned = self.bin(pre, W, H)
        fwd = self.composite_fwd(pre, binned, colors, bg, W, H)
        return pre, binned, fwd

    def rasterize_backward(self, means, scales, rots, opac, colors, view, proj, bg, W, H, tanfovx, tanfovy, pre,
                           binned, fwd, dL_dimage, dL_dalpha, dL_ddepth, scale_modifier=1.0):
        cb = self.composite_bwd(pre, binned, colors, bg, fwd, dL_dimage, dL_dalpha, dL_ddepth, W, H)
        pb = self.preprocess_bwd(means, scales, rots, view, proj, W, H, tanfovx, tanfovy, pre["radii"],
                                 cb["dL_dmean2D"], cb["dL_dconic"], cb["dL_ddepth"], scale_modifier)
        return dict(dL_dmeans=pb["dL_dmeans"], dL_dscales=pb["dL_dscales"], dL_drots=pb["dL_drots"],
                    dL_dopacity=cb["dL_dopacity"], dL_dcolor=cb["dL_dcolor"], dL_dmean2D=cb["dL_dmean2D"],
                    dL_dconic=cb["dL_dconic"], dL_ddepth=cb["dL_ddepth"], dL_dcov3d=pb["dL_dcov3d"])

    # ---- whole step, the B x V loop of core/gs.py:42-93 (OpenMP over views) ---------------------------
    def render_step(self, gaussians, view_mats, proj_mats, bg, W, H, tanfovx, tanfovy, scale_modifier=1.0,
                    dL_dimage=None, dL_dalpha=None, dL_ddepth=None):
        gaussians = self._a(gaussians)
        B, N, C = gaussians.shape
        assert C == 14
        view_mats = self._a(view_mats).reshape(B, -1, 16)
        V = view_mats.shape[1]
        proj_mats = self._a(proj_mats, (B, V, 16))
        bg = self._a(bg, (3,))
        images = np.zeros((B, V, 3, H, W), self.dt)
        alphas = np.zeros((B, V, 1, H, W), self.dt)
        depths = np.zeros((B, V, 1, H, W), self.dt)
        radii = np.zeros((B, V, N), np.int32)
        bwd = dL_dimage is not None
        dg = np.zeros((B, N, 14), self.dt) if bwd else None
        if bwd:
            dL_dimage = self._a(dL_dimage, (B, V, 3, H, W))
            dL_dalpha = self._a(dL_dalpha if dL_dalpha is not None else np.zeros((B, V, 1, H, W)), (B, V, 1, H, W))
            dL_ddepth = self._a(dL_ddepth if dL_ddepth is not None else np.zeros((B, V, 1, H, W)), (B, V, 1, H, W))
        L = self.lib.orc_render_step(
            ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(V), _p(gaussians), _p(view_mats), _p(proj_mats),
            ctypes.c_int(W), ctypes.c_int(H), self.creal(tanfovx), self.creal(tanfovy), self.creal(scale_modifier),
            _p(bg), _p(images), _p(alphas), _p(depths), _p(radii), _p(dL_dimage) if bwd else None,
            _p(dL_dalpha) if bwd else None, _p(dL_ddepth) if bwd else None, _p(dg))
        return dict(image=images, alpha=alphas, depth=depths, radii=radii, num_rendered=int(L), dgaussians=dg)
