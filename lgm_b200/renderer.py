"""Drop-in for /root/reference/core/gs.py::GaussianRenderer (lines 16-98): same constructor argument, same `render`
signature and dict keys — plus "depth", which the reference computes and discards (core/gs.py:76).

The reference loops over B and V in Python and calls the rasterizer once per view (core/gs.py:42-85); here all
B x V views go through ONE batched set of CUDA launches with a single host readback per step.
"""
import math
from types import SimpleNamespace

import torch

from . import ops, ply


def default_options(output_size=256, fovy=49.1, znear=0.5, zfar=2.5):
    """The four fields of /root/reference/core/options.py the renderer reads (core/gs.py:23-29,59-60)."""
    return SimpleNamespace(output_size=output_size, fovy=fovy, znear=znear, zfar=zfar)


class GaussianRenderer:
    def __init__(self, opt, device="cuda"):
        self.opt = opt
        self.device = torch.device(device)
        self.bg_color = torch.tensor([1, 1, 1], dtype=torch.float32, device=self.device)  # core/gs.py:20
        # intrinsics, core/gs.py:23-29
        self.tan_half_fov = math.tan(0.5 * math.radians(self.opt.fovy))
        self.proj_matrix = torch.zeros(4, 4, dtype=torch.float32)
        self.proj_matrix[0, 0] = 1 / self.tan_half_fov
        self.proj_matrix[1, 1] = 1 / self.tan_half_fov
        self.proj_matrix[2, 2] = (opt.zfar + opt.znear) / (opt.zfar - opt.znear)
        self.proj_matrix[3, 2] = -(opt.zfar * opt.znear) / (opt.zfar - opt.znear)
        self.proj_matrix[2, 3] = 1
        self._view_scene_cache = {}

    def _view_scene(self, B, V):
        key = (B, V)
        if key not in self._view_scene_cache:
            self._view_scene_cache[key] = torch.arange(B, dtype=torch.int32).repeat_interleave(V)
        return self._view_scene_cache[key]

    def render(self, gaussians, cam_view, cam_view_proj, cam_pos, bg_color=None, scale_modifier=1,
               max_views_per_call=None, return_depth=True):
        # gaussians: [B, N, 14]; cam_view, cam_view_proj: [B, V, 4, 4]; cam_pos: [B, V, 3] (only used by SH
        # evaluation upstream, which LGM never triggers — kept for signature parity)
        B, V = cam_view.shape[:2]
        S = int(self.opt.output_size)
        g = gaussians.contiguous().float()                     # core/gs.py:45-49 (.contiguous().float())
        vm = cam_view.reshape(B * V, 16).contiguous().float()  # core/gs.py:54-55
        pm = cam_view_proj.reshape(B * V, 16).contiguous().float()
        bg = (self.bg_color if bg_color is None else bg_color).to(g.device).float().reshape(3).contiguous()
        # core/gs.py:87 clamps the image (alpha is not clamped): fused into the compositing kernels (clamp_image)
        cfg = ops.ViewConfig(S, S, float(self.tan_half_fov), float(self.tan_half_fov), float(scale_modifier),
                             clamp_image=True, want_depth=bool(return_depth))
        image, alpha, depth, _radii = ops.render_views(g, vm, pm, self._view_scene(B, V), bg, cfg, max_views_per_call)
        out = {
            "image": image.view(B, V, 3, S, S),   # [B, V, 3, H, W]
            "alpha": alpha.view(B, V, 1, S, S),   # [B, V, 1, H, W]
        }
        # a superset of the reference's dict (which computes the depth and drops it, core/gs.py:76); return_depth=False
        # gives exactly the reference's keys and skips the depth image in the kernels
        if return_depth:
            out["depth"] = depth.view(B, V, 1, S, S)
        return out

    # on-disk format of the path's input (/root/reference/core/gs.py:101-190)
    def save_ply(self, gaussians, path, compatible=True):
        return ply.save_ply(gaussians, path, compatible)

    def load_ply(self, path, compatible=True):
        return ply.load_ply(path, compatible)
