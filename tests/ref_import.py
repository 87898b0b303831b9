"""TEST INFRASTRUCTURE ONLY -- imports the reference's OWN hot-path Python, unmodified.

The reference's modules (gaussian_renderer, scene.gaussian_model, utils.loss_utils, utils.general_utils,
fused_ssim) fail to import in this image only because of third-party packages that are absent
(matplotlib, plyfile, tensordict, cupy / cupyx, simple_knn) -- none of which the hot path computes with,
except cupyx.scipy.ndimage.gaussian_filter, for which scipy.ndimage.gaussian_filter stands in (same
algorithm and defaults: truncate=4, mode='reflect').  This module installs stub modules for those names
and imports the reference files from

    /root/reference                 (this container), or
    oracle/_ref/pyref               (the unmodified files staged by `make -C oracle ref_py`; git-ignored, travels
                                     to the GPU box like the compiled reference kernels next to it)

with the rasteriser packages ``diff_gaussian_rasterization_{h36m,panoptic,op}`` bound to either

    backend="ours"   this repo's drop-in packages (the C-ABI library): proves the boundary through the
                     reference's own caller, or
    backend="ref"    oracle/ref_rasterizer.py = the UNMODIFIED reference kernels (oracle/_ref/libref_rast_*.so):
                     the reference pipeline itself, which pins oracle/pipeline.py and generates the goldens.

Nothing in the product imports this file.
"""
import importlib
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ["/root/reference", os.path.join(ROOT, "oracle", "_ref", "pyref")]
_VARIANTS = {"h36m": "diff_gaussian_rasterization_h36m", "panoptic": "diff_gaussian_rasterization_panoptic", "op": "diff_gaussian_rasterization_op"}
_cache = {}


def ref_root():
    for c in CANDIDATES:
        if os.path.exists(os.path.join(c, "gaussian_renderer", "__init__.py")) and os.path.exists(os.path.join(c, "utils", "loss_utils.py")):
            return c
    return None


def available():
    return ref_root() is not None


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__stub__ = True
    sys.modules[name] = m
    return m


def _install_stubs():
    """Stand-ins for the absent third-party imports (only names the reference files touch at import time or on the hot path)."""
    import numpy as np
    import scipy.ndimage
    import torch
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = _stub("matplotlib")
            mpl.pyplot = _stub("matplotlib.pyplot")
    if "plyfile" not in sys.modules:
        _stub("plyfile", PlyData=type("PlyData", (), {}), PlyElement=type("PlyElement", (), {}))
    if "tensordict" not in sys.modules:
        _stub("tensordict", TensorDict=type("TensorDict", (dict,), {"__init__": lambda self, d=None, *a, **k: dict.__init__(self, d or {})}))
    if "cupy" not in sys.modules:
        def asarray(x):        # cp.asarray(torch_cuda_tensor): the stand-in filter runs on the host
            return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)
        _stub("cupy", asarray=asarray)
        cx = _stub("cupyx"); cs = _stub("cupyx.scipy"); cn = _stub("cupyx.scipy.ndimage", gaussian_filter=scipy.ndimage.gaussian_filter)
        cx.scipy = cs; cs.ndimage = cn
    if "simple_knn" not in sys.modules:
        sk = _stub("simple_knn"); sk._C = _stub("simple_knn._C", distCUDA2=None)


def _rasterizer_modules(backend):
    """name -> module for the three diff_gaussian_rasterization_* packages."""
    if backend == "ours":
        return {pkg: importlib.import_module(pkg) for pkg in _VARIANTS.values()}
    from oracle import ref_rasterizer as refr
    mods = {}
    for variant, pkg in _VARIANTS.items():
        def make(variant=variant):
            class GaussianRasterizer(refr.GaussianRasterizer):
                def __init__(self, raster_settings):
                    super().__init__(raster_settings, variant)
            return GaussianRasterizer
        m = types.ModuleType(pkg)
        m.GaussianRasterizationSettings = refr.GaussianRasterizationSettings
        m.GaussianRasterizer = make()
        mods[pkg] = m
    return mods


def load(backend="ours"):
    """Namespace of the reference's modules: .gaussian_renderer (bound to `backend`), .gaussian_model, .loss_utils,
    .general_utils, .utils (the registries `losses`, `consistency_losses`), .root."""
    if backend in _cache:
        return _cache[backend]
    root = ref_root()
    if root is None:
        raise RuntimeError("reference Python not available (neither /root/reference nor oracle/_ref/pyref)")
    _install_stubs()
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    if root not in sys.path:
        sys.path.append(root)
    # `scene/__init__.py` pulls the dataset readers / argument parser (out of scope, more absent packages): register the
    # package WITHOUT running it, so that `scene.gaussian_model` resolves to the reference file
    if "scene" not in sys.modules:
        pkg = types.ModuleType("scene"); pkg.__path__ = [os.path.join(root, "scene")]
        sys.modules["scene"] = pkg
    utils = importlib.import_module("utils")                    # utils/__init__.py: loss + early-stopping registries
    if not os.path.abspath(getattr(utils, "__file__", "")).startswith(os.path.abspath(root)):
        raise RuntimeError(f"a foreign top-level `utils` package shadows the reference's: {utils.__file__}")
    gm = importlib.import_module("scene.gaussian_model")
    saved = {pkg: sys.modules.get(pkg) for pkg in _VARIANTS.values()}
    try:
        sys.modules.update(_rasterizer_modules(backend))
        spec = importlib.util.spec_from_file_location(f"refpy_gaussian_renderer_{backend}", os.path.join(root, "gaussian_renderer", "__init__.py"))
        gr = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(gr)
    finally:
        for pkg, m in saved.items():
            if m is None:
                sys.modules.pop(pkg, None)
            else:
                sys.modules[pkg] = m
    ns = types.SimpleNamespace(root=root, backend=backend, gaussian_renderer=gr, gaussian_model=gm, utils=utils,
                               loss_utils=importlib.import_module("utils.loss_utils"),
                               general_utils=importlib.import_module("utils.general_utils"),
                               graphics_utils=importlib.import_module("utils.graphics_utils"))
    _cache[backend] = ns
    return ns


def load_fused_ssim(kernels="ref"):
    """The reference's fused_ssim/__init__.py (FusedSSIMMap + fused_ssim) on top of
    kernels="ref": ITS OWN kernels, oracle/_ref/fused_ssim_cuda.so (submodules/fused-ssim/ssim.cu compiled unmodified by
                   `make -C oracle ref_ssim`), or
    kernels="ours": this repo's fused_ssim package's extension surface (fusedssim / fusedssim_backward)."""
    from oracle import ref_ssim
    if kernels == "ref":
        return ref_ssim.load()
    import fused_ssim as ours
    ext = types.ModuleType("fused_ssim_cuda")
    ext.fusedssim, ext.fusedssim_backward = ours.fusedssim, ours.fusedssim_backward
    return ref_ssim.load(ext, tag="ours")


def ref_ssim_available():
    from oracle import ref_ssim
    return ref_ssim.available()
