"""Drop-in module name the reference imports (train.py:19, helpers.py:18-19):

    from diff_gaussian_rasterization import GaussianRasterizer as Renderer
    from diff_gaussian_rasterization import GaussianRasterizationSettings as Camera

Everything is implemented in :mod:`topo4d_b200` (hand-written sm_100a kernels behind a C ABI).
"""
from topo4d_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer, render_views  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "render_views"]
