"""Drop-in for the reference's ``diff_gaussian_rasterization_panoptic`` package
(submodules/diff-gaussian-rasterization-panoptic; NUM_CHANNELS = 19, cuda_rasterizer/config.h:15),
backed by skelsplat_b200's sm_100a library.  Imported by gaussian_renderer/__init__.py:15-22."""
from skelsplat_b200.rasterizer import GaussianRasterizationSettings, rasterize_gaussians  # noqa: F401
from skelsplat_b200.rasterizer import GaussianRasterizer as _Base

NUM_CHANNELS = 19


class GaussianRasterizer(_Base):
    NUM_CHANNELS = NUM_CHANNELS
