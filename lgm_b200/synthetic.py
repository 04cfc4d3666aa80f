"""Seeded synthetic inputs for the splat-render path.  CPU generators, fp32.

trained-like: positions uniform in a ball r = 0.5, scale = exp(U[-6.5,-4]), unit quaternions,
              opacity = sigmoid(N(0, 2^2)), rgb = U[0,1].
init-like   : LGM's output activations (/root/reference/core/models.py:40-44) applied to N(0,1) noise:
              pos = clamp(x,-1,1), opacity = sigmoid, scale = 0.1 softplus, rot = normalize, rgb = 0.5 tanh + 0.5.
"""
import torch
import torch.nn.functional as F

from .cameras import orbit_views


def make_gaussians(B, N, kind="trained", seed=1234):
    out = []
    for b in range(B):
        g = torch.Generator().manual_seed(seed + b)
        if kind == "trained":
            d = torch.randn(N, 3, generator=g)
            d = d / d.norm(dim=-1, keepdim=True).clamp_min(1e-12)
            r = 0.5 * torch.rand(N, 1, generator=g) ** (1.0 / 3.0)
            pos = d * r
            scale = torch.exp(-6.5 + 2.5 * torch.rand(N, 3, generator=g))
            rot = F.normalize(torch.randn(N, 4, generator=g), dim=-1)
            opacity = torch.sigmoid(2.0 * torch.randn(N, 1, generator=g))
            rgb = torch.rand(N, 3, generator=g)
        elif kind == "init":
            x = torch.randn(N, 14, generator=g)
            pos = x[:, 0:3].clamp(-1, 1)
            opacity = torch.sigmoid(x[:, 3:4])
            scale = 0.1 * F.softplus(x[:, 4:7])
            rot = F.normalize(x[:, 7:11], dim=-1)
            rgb = 0.5 * torch.tanh(x[:, 11:14]) + 0.5
        else:
            raise ValueError(kind)
        out.append(torch.cat([pos, opacity, scale, rot, rgb], dim=-1))
    return torch.stack(out).float().contiguous()  # [B,N,14]


def make_cameras(B, V, fovy=49.1, znear=0.5, zfar=2.5, radius=1.5, seed=1234):
    cv, cvp, cp = [], [], []
    for b in range(B):
        a, p, c = orbit_views(V, radius, fovy, znear, zfar, seed=seed + 7919 * (b + 1))
        cv.append(a), cvp.append(p), cp.append(c)
    return torch.stack(cv), torch.stack(cvp), torch.stack(cp)  # [B,V,4,4], [B,V,4,4], [B,V,3]


def make_upstream_grads(B, V, H, W, seed=1234, with_depth=False):
    """Loss-shaped grads: 2 (x - G) / numel with x, G ~ U[0,1]  (/root/reference/core/models.py:153)."""
    g = torch.Generator().manual_seed(seed + 99)
    n_img, n_a = B * V * 3 * H * W, B * V * H * W
    d_img = 2.0 * (torch.rand(B, V, 3, H, W, generator=g) - torch.rand(B, V, 3, H, W, generator=g)) / n_img
    d_alpha = 2.0 * (torch.rand(B, V, 1, H, W, generator=g) - torch.rand(B, V, 1, H, W, generator=g)) / n_a
    d_depth = (2.0 * (torch.rand(B, V, 1, H, W, generator=g) - 0.5) / n_a) if with_depth else torch.zeros(B, V, 1, H, W)
    return d_img.float(), d_alpha.float(), d_depth.float()


def make_bg(seed=1234):
    g = torch.Generator().manual_seed(seed + 5)
    return torch.rand(3, generator=g).float()
