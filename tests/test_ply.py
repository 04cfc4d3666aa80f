"""PLY on-disk format (/root/reference/core/gs.py:101-190): header layout, pruning, activation
inversion and save -> load round trip.  Host-side only (no GPU)."""
import numpy as np
import torch

from lgm_b200.ply import PROPS, SH_C0, load_ply, save_ply
from lgm_b200.synthetic import make_gaussians


def test_header_and_layout(tmp_path):
    g = make_gaussians(1, 50, "trained", seed=1)
    g[0, :, 3] = g[0, :, 3].clamp(0.01, 0.99)
    g[0, :5, 3] = 0.001                                     # below the 0.005 pruning threshold
    p = str(tmp_path / "a.ply")
    kept = save_ply(g, p, compatible=True)
    assert kept == 45
    raw = open(p, "rb").read()
    head, body = raw.split(b"end_header\n")
    lines = head.decode().strip().split("\n")
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 45"]
    assert [l.split()[2] for l in lines[3:]] == PROPS and all(l.split()[1] == "float" for l in lines[3:])
    arr = np.frombuffer(body, "<f4").reshape(45, 14)
    src = g[0, 5:]
    assert np.allclose(arr[:, 0:3], src[:, 0:3].numpy())
    assert np.allclose(arr[:, 3:6], ((src[:, 11:14] - 0.5) / SH_C0).numpy(), atol=1e-6)       # f_dc
    assert np.allclose(arr[:, 6], torch.log(src[:, 3] / (1 - src[:, 3])).numpy(), atol=1e-5)  # inverse sigmoid
    assert np.allclose(arr[:, 7:10], torch.log(src[:, 4:7] + 1e-8).numpy(), atol=1e-6)
    assert np.allclose(arr[:, 10:14], src[:, 7:11].numpy())


def test_round_trip(tmp_path):
    g = make_gaussians(1, 300, "init", seed=3)
    g[0, :, 3] = g[0, :, 3].clamp(0.01, 0.99)
    for compatible in (True, False):
        p = str(tmp_path / f"b{int(compatible)}.ply")
        save_ply(g, p, compatible=compatible)
        back = load_ply(p, compatible=compatible)
        assert back.shape == (300, 14) and back.dtype == torch.float32
        assert torch.allclose(back, g[0], rtol=1e-4, atol=1e-5)


def test_renderer_methods(tmp_path):
    import lgm_b200.renderer as R
    assert R.GaussianRenderer.save_ply is not None and R.GaussianRenderer.load_ply is not None
    g = make_gaussians(1, 20, "trained", seed=2)
    p = str(tmp_path / "c.ply")
    R.GaussianRenderer.save_ply(None, g, p)
    assert R.GaussianRenderer.load_ply(None, p).shape == (20, 14)
