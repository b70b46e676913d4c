"""topo4d_b200 -- B200-native (sm_100a) hot path of Topo4D: the differentiable Gaussian-splatting
rasterizer behind ``diff_gaussian_rasterization`` and face3d's ``render_colors`` texture bake.

Importing the package is cheap and GPU-free; the CUDA library is loaded (and must exist or be
buildable) on first use of an operator.  There is no CPU fallback on the product path.
"""
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, render_views  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "render_views"]
