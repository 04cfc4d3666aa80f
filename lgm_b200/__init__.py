"""lgm_b200 — sm_100a (B200) implementation of LGM's Gaussian-splat render path (forward + backward).

Public surface mirrors the reference's own interface for this path:
  GaussianRasterizationSettings, GaussianRasterizer   (the `diff_gaussian_rasterization` package, core/gs.py:7-10)
  GaussianRenderer                                   (/root/reference/core/gs.py:16-98)
The CUDA library (lgm_b200/liblgm_b200.so, built by `python -m lgm_b200.build`) is loaded on first use; there is no
CPU or PyTorch fallback — a missing library or a non-CUDA tensor raises.
"""
from ._lib import LgmError
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer
from .activations import activate_gaussians
from .losses import lpips_input, mse_image_alpha_loss
from .renderer import GaussianRenderer, default_options

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "GaussianRenderer", "default_options", "LgmError",
           "mse_image_alpha_loss", "lpips_input", "activate_gaussians"]
