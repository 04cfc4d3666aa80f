"""Camera conventions the splat renderer accepts, restated because kiui is not installed.

Follows the reference's callers: orbit camera as used at /root/reference/core/provider_lvis.py:135 and
/root/reference/infer.py:118 (kiui.cam.orbit_camera, OpenGL look-at), the COLMAP flip and matrix products of
/root/reference/core/provider_lvis.py:201-209, and the projection of /root/reference/core/gs.py:23-29.
Matrices are row-vector convention: cam_view = inverse(c2w)^T, cam_view_proj = cam_view @ P; the kernels read
the flat 16 floats as m[i + 4 k].
"""
import math

import numpy as np
import torch


def projection_matrix(fovy_deg, znear, zfar):
    """/root/reference/core/gs.py:23-29."""
    t = math.tan(0.5 * math.radians(fovy_deg))
    P = torch.zeros(4, 4, dtype=torch.float32)
    P[0, 0] = 1.0 / t
    P[1, 1] = 1.0 / t
    P[2, 2] = (zfar + znear) / (zfar - znear)
    P[3, 2] = -(zfar * znear) / (zfar - znear)
    P[2, 3] = 1.0
    return P


def _normalize(v):
    return v / (np.linalg.norm(v) + 1e-20)


def orbit_camera(elevation_deg, azimuth_deg, radius=1.0, target=None):
    """Camera-to-world 4x4 (OpenGL convention), numpy float32.  SURVEY.md Appendix B."""
    e, a = math.radians(elevation_deg), math.radians(azimuth_deg)
    pos = np.array([radius * math.cos(e) * math.sin(a), -radius * math.sin(e), radius * math.cos(e) * math.cos(a)],
                   dtype=np.float64)
    tgt = np.zeros(3) if target is None else np.asarray(target, np.float64)
    pos = pos + tgt
    fwd = _normalize(pos - tgt)
    right = _normalize(np.cross(np.array([0.0, 1.0, 0.0]), fwd))
    up = _normalize(np.cross(fwd, right))
    T = np.eye(4, dtype=np.float64)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = right, up, fwd, pos
    return T.astype(np.float32)


def camera_matrices(c2w_opengl, proj):
    """(cam_view, cam_view_proj, cam_pos) as /root/reference/core/provider_lvis.py:201-209."""
    c2w = torch.as_tensor(c2w_opengl, dtype=torch.float32).clone()
    c2w[..., :3, 1:3] *= -1  # OpenGL -> COLMAP
    cam_view = torch.inverse(c2w).transpose(-1, -2)
    cam_view_proj = cam_view @ proj
    cam_pos = -c2w[..., :3, 3]
    return cam_view.contiguous(), cam_view_proj.contiguous(), cam_pos.contiguous()


def orbit_views(n_views, radius, fovy_deg, znear, zfar, seed=0, elevation_range=(-30.0, 30.0)):
    """V orbit cameras: azimuth = 360 v / V + U[0, 360/V), elevation ~ U[range]  (SURVEY.md §8d)."""
    rng = np.random.RandomState(seed)
    proj = projection_matrix(fovy_deg, znear, zfar)
    c2ws = []
    for v in range(n_views):
        el = rng.uniform(*elevation_range)
        az = 360.0 * v / n_views + rng.uniform(0.0, 360.0 / n_views)
        c2ws.append(orbit_camera(el, az, radius))
    return camera_matrices(np.stack(c2ws), proj)
