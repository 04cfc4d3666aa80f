"""Import-name shim: `from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer`
(/root/reference/core/gs.py:7-10) resolves to the sm_100a implementation in lgm_b200 when this repository is on
sys.path ahead of (or instead of) the external package."""
from lgm_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer"]
