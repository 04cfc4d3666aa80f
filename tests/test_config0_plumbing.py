"""BASELINE.json configs[0] — the plumbing case: reference UNet output -> 1x1 conv -> activations -> render one 256^2 view.

tests/golden/config0_unet_splatter.npz was produced IN THE BUILD CONTAINER by the reference's own network code
(/root/reference/core/unet.py, core/options.py `tiny`, the reshape / activations of core/models.py:95-115; generator:
tests/golden/make_unet_fixture.py) — it holds the raw splatter image x [1, 16384, 14] and the Gaussians the reference's
activations make of it.  Here:
  * CPU: a numpy restatement of models.py:40-44 reproduces the stored Gaussians (this pins `F.normalize`'s default dim=1:
    every quaternion component is normalised over the 16,384 Gaussians), and the oracle renders them;
  * GPU: lgm_b200.activate_gaussians(x) -> GaussianRenderer.render -> loss -> backward, against the oracle's render of the
    reference's Gaussians and the oracle's gradient chained through torch's autograd of the reference activations.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, tan_half
from lgm_b200.synthetic import make_bg, make_cameras

FIXTURE = os.path.join(GOLDEN_DIR, "config0_unet_splatter.npz")
S, FOVY = 256, 49.1


def _load():
    z = np.load(FIXTURE)
    return {k: z[k] for k in z.files}


def _reference_activations(x):
    """/root/reference/core/models.py:40-44,109-115 in torch (double precision capable)."""
    import torch.nn.functional as F
    pos = x[..., 0:3].clamp(-1, 1)
    opacity = torch.sigmoid(x[..., 3:4])
    scale = 0.1 * F.softplus(x[..., 4:7])
    rotation = F.normalize(x[..., 7:11])          # no dim: the reference's call
    rgbs = 0.5 * torch.tanh(x[..., 11:]) + 0.5
    return torch.cat([pos, opacity, scale, rotation, rgbs], dim=-1)


def test_fixture_and_activation_semantics(oracle32):
    c = _load()
    x, g = c["x"], c["gaussians"]
    assert x.shape == (1, 16384, 14) and g.shape == x.shape and str(c["preset"]) == "tiny" and int(c["unet_params"]) > 50e6
    x64 = x.astype(np.float64)
    rot = x64[..., 7:11]
    # dim=1 of a [B, N, 4] tensor is N: per (scene, component) norm over the Gaussians — NOT unit quaternions
    ref_rot = rot / np.maximum(np.sqrt((rot ** 2).sum(axis=1, keepdims=True)), 1e-12)
    assert np.abs(g[..., 7:11] - ref_rot).max() <= 1e-6
    assert np.abs(np.linalg.norm(g[..., 7:11], axis=-1) - 1).min() > 0.9      # quaternion norms are ~1/sqrt(N), far from 1
    assert np.abs(g[..., 0:3] - np.clip(x64[..., 0:3], -1, 1)).max() <= 1e-6
    assert np.abs(g[..., 3] - 1 / (1 + np.exp(-x64[..., 3]))).max() <= 1e-6
    assert np.abs(g[..., 4:7] - 0.1 * np.log1p(np.exp(x64[..., 4:7]))).max() <= 1e-6
    assert np.abs(g[..., 11:] - (0.5 * np.tanh(x64[..., 11:]) + 0.5)).max() <= 1e-6
    assert np.abs(g - _reference_activations(torch.from_numpy(x)).numpy()).max() <= 1e-6
    # the oracle renders the reference's Gaussians: one 256^2 view (configs[0])
    cv, cvp, _ = make_cameras(1, 1, fovy=FOVY, seed=3)
    t = tan_half(FOVY)
    r = oracle32.render_step(g, cv.numpy(), cvp.numpy(), make_bg(3).numpy(), S, S, t, t)
    assert np.isfinite(r["image"]).all() and r["alpha"].max() > 0.5 and r["num_rendered"] > 10000


@pytest.mark.gpu
def test_config0_pipeline_on_gpu(oracle32, oracle64):
    from lgm_b200 import GaussianRenderer, activate_gaussians, default_options
    dev = "cuda:0"
    c = _load()
    x = torch.from_numpy(c["x"])
    cv, cvp, cp = make_cameras(1, 1, fovy=FOVY, seed=3)
    bg = make_bg(3)
    xg = x.to(dev).requires_grad_(True)
    g = activate_gaussians(xg)                                   # models.py:109-115 (rot_axis="reference")
    assert (g.detach().cpu() - torch.from_numpy(c["gaussians"])).abs().max().item() <= 2e-6
    r = GaussianRenderer(default_options(output_size=S, fovy=FOVY), device=dev)
    out = r.render(g, cv.to(dev), cvp.to(dev), cp.to(dev), bg_color=bg.to(dev))     # models.py:141
    rng = np.random.RandomState(0)
    wi, wa = rng.randn(1, 1, 3, S, S).astype(np.float32), rng.randn(1, 1, 1, S, S).astype(np.float32)
    ((out["image"] * torch.tensor(wi, device=dev)).sum() + (out["alpha"] * torch.tensor(wa, device=dev)).sum()).backward()
    torch.cuda.synchronize()
    t = tan_half(FOVY)
    a = (cv.numpy(), cvp.numpy(), bg.numpy(), S, S, t, t)
    fw = oracle32.render_step(c["gaussians"], *a)
    assert np.abs(out["image"].detach().cpu().numpy() - np.clip(fw["image"], 0, 1)).max() <= 1e-4
    assert np.abs(out["alpha"].detach().cpu().numpy() - fw["alpha"]).max() <= 1e-4
    mask = ((fw["image"] >= 0) & (fw["image"] <= 1)).astype(np.float32)        # the renderer's clamp (core/gs.py:87)
    for o, tol in ((oracle32, 1e-3), (oracle64, 5e-3)):
        bw = o.render_step(c["gaussians"], *a, 1.0, wi * mask, wa, np.zeros((1, 1, 1, S, S), np.float32))
        xr = x.double().requires_grad_(True)
        (_reference_activations(xr) * torch.from_numpy(bw["dgaussians"]).double()).sum().backward()
        ref = xr.grad.numpy()
        got = xg.grad.cpu().numpy()
        for sl in (slice(0, 3), slice(3, 4), slice(4, 7), slice(7, 11), slice(11, 14)):
            scale = np.abs(ref[..., sl]).max() + 1e-30
            assert np.abs(got[..., sl] - ref[..., sl]).max() <= tol * scale, (o.dt, sl)
