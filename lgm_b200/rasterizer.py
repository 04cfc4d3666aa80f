"""Drop-in for `diff_gaussian_rasterization` (the external package imported at /root/reference/core/gs.py:7-10):
`GaussianRasterizationSettings` and `GaussianRasterizer`, same field order, argument names, return order
(color, radii, depth, alpha) and error behaviour (SURVEY.md §8b level 1) — backed by the batched sm_100a kernels
with n_views = 1.

Not on LGM's path (core/gs.py:79-84 passes colors_precomp, scales and rotations) but served for API completeness:
`shs` (spherical-harmonics colours) by the sh.cu kernels, whose output enters the renderer exactly like
colors_precomp, and `cov3D_precomp` [P,6] by the covariance inputs of the preprocess kernels (the gradient then
stops at the covariance, as upstream's dL_dcov3D).
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _lib, ops


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class _RasterizeGaussians(torch.autograd.Function):
    """One view.  Gradients in upstream's order: means3D, means2D, sh, colors_precomp, opacities, scales,
    rotations, cov3Ds_precomp, raster_settings."""

    @staticmethod
    def forward(ctx, means3D, means2D, colors_precomp, opacities, scales, rotations, cov3D_precomp, raster_settings):
        rs = raster_settings
        dev = means3D.device
        P = means3D.shape[0]
        if cov3D_precomp is not None:  # the scale / rotation columns are ignored by the kernels
            scales, rotations = means3D.new_zeros(P, 3), means3D.new_zeros(P, 4)
            ctx.cov3d = cov3D_precomp.float().reshape(1, P, 6).contiguous()
        else:
            ctx.cov3d = None
        g = torch.cat([means3D.float(), opacities.float().reshape(P, 1), scales.float(), rotations.float(),
                       colors_precomp.float()], dim=-1).reshape(1, P, 14).contiguous()
        cfg = ops.ViewConfig(int(rs.image_height), int(rs.image_width), float(rs.tanfovx), float(rs.tanfovy),
                             float(rs.scale_modifier))
        vm = rs.viewmatrix.to(dev).float().reshape(1, 16).contiguous()
        pm = rs.projmatrix.to(dev).float().reshape(1, 16).contiguous()
        bg = rs.bg.to(dev).float().reshape(3).contiguous()
        view_scene = torch.zeros(1, dtype=torch.int32, device=dev)
        offsets = torch.tensor([0, 1], dtype=torch.int32, device=dev)
        image, alpha, depth, st = ops.forward_views(g, vm, pm, view_scene, offsets, bg, cfg, cov3d=ctx.cov3d)
        ctx.st = st
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(g, vm, pm, bg, alpha)
        radii = st.radii.view(P)
        ctx.mark_non_differentiable(radii)
        return image[0], radii, depth[0], alpha[0]

    @staticmethod
    def backward(ctx, grad_color, _grad_radii, grad_depth, grad_alpha):
        g, vm, pm, bg, alpha = ctx.saved_tensors
        st = ctx.st
        H, W = st.cfg.image_height, st.cfg.image_width
        z = lambda t, c: (torch.zeros(1, c, H, W, device=g.device) if t is None else t.reshape(1, c, H, W).contiguous().float())
        res = ops.backward_views(g, vm, pm, bg, st, alpha, z(grad_color, 3), z(grad_alpha, 1),
                                 None if grad_depth is None else z(grad_depth, 1), cov3d=ctx.cov3d)
        d_gauss, rows = res[0], res[1]
        d = d_gauss[0]
        P = d.shape[0]
        d_means2D = torch.zeros(P, 3, device=g.device)
        d_means2D[:, :2] = ops.screen_gradients(st, rows)[:P, 0:2]
        # (means3D, means2D, colors_precomp, opacities, scales, rotations, cov3D_precomp, raster_settings)
        if ctx.cov3d is not None:
            return d[:, 0:3], d_means2D, d[:, 11:14], d[:, 3:4], None, None, res[2][0], None
        return d[:, 0:3], d_means2D, d[:, 11:14], d[:, 3:4], d[:, 4:7], d[:, 7:11], None, None


class _SHToColor(torch.autograd.Function):
    """colors [P,3] = max(0.5 + SH(normalize(means3D - campos)) . shs, 0) with the active degree of the raster
    settings (upstream computeColorFromSH); gradients to means3D (through the view direction) and shs."""

    @staticmethod
    def forward(ctx, means3D, shs, campos, degree):
        P = means3D.shape[0]
        if shs.dim() != 3 or shs.shape[0] != P or shs.shape[2] != 3:
            raise _lib.LgmError("shs must have dimensions (num_points, num_coeffs, 3)")
        means, sh, cam = means3D.float().contiguous(), shs.float().contiguous(), campos.float().reshape(3).contiguous()
        colors = torch.empty(P, 3, device=means.device)
        clamped = torch.empty(P, 3, dtype=torch.uint8, device=means.device)
        _lib.check(_lib.lib().lgm_sh_forward(ops._stream(), P, int(degree), sh.shape[1], _lib.ptr(means), _lib.ptr(cam),
                                             _lib.ptr(sh), _lib.ptr(colors), _lib.ptr(clamped)), "lgm_sh_forward")
        ops.launch_counter["kernels"] += 1 if P else 0
        ctx.degree = int(degree)
        ctx.save_for_backward(means, sh, cam, clamped)
        return colors

    @staticmethod
    def backward(ctx, grad_colors):
        means, sh, cam, clamped = ctx.saved_tensors
        P = means.shape[0]
        g = grad_colors.float().contiguous()
        d_sh, d_means = torch.empty_like(sh), torch.empty_like(means)
        _lib.check(_lib.lib().lgm_sh_backward(ops._stream(), P, ctx.degree, sh.shape[1], _lib.ptr(means), _lib.ptr(cam),
                                              _lib.ptr(sh), _lib.ptr(clamped), _lib.ptr(g), _lib.ptr(d_sh),
                                              _lib.ptr(d_means)), "lgm_sh_backward")
        ops.launch_counter["kernels"] += 1 if P else 0
        return d_means, d_sh, None, None


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # upstream: rasterizer.markVisible -> _C.mark_visible(positions, viewmatrix, projmatrix), under no_grad
        with torch.no_grad():
            rs = self.raster_settings
            pos = positions.float().contiguous()
            if pos.dim() != 2 or pos.shape[1] != 3:
                raise _lib.LgmError("means3D must have dimensions (num_points, 3)")
            vis = torch.empty(pos.shape[0], dtype=torch.uint8, device=pos.device)
            vm = rs.viewmatrix.to(pos.device).float().reshape(16).contiguous()
            L = _lib.lib()
            _lib.check(L.lgm_mark_visible(ops._stream(), pos.shape[0], _lib.ptr(pos), _lib.ptr(vm), _lib.ptr(vis)),
                       "lgm_mark_visible")
            return vis.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if cov3D_precomp is not None and (cov3D_precomp.dim() != 2 or cov3D_precomp.shape != (means3D.shape[0], 6)):
            raise _lib.LgmError("cov3D_precomp must have dimensions (num_points, 6)")
        if means3D.dim() != 2 or means3D.shape[1] != 3:
            raise _lib.LgmError("means3D must have dimensions (num_points, 3)")
        if not means3D.is_cuda:
            raise _lib.LgmError("lgm_b200 has no CPU path: tensors must be on a CUDA device")
        if shs is not None:
            rs = self.raster_settings
            colors_precomp = _SHToColor.apply(means3D, shs, rs.campos.to(means3D.device), int(rs.sh_degree))
        return _RasterizeGaussians.apply(means3D, means2D, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                         self.raster_settings)
