"""TEST INFRASTRUCTURE ONLY -- fused-ssim's OWN kernels and wrapper.

oracle/_ref/fused_ssim_cuda.so is submodules/fused-ssim/ssim.cu + ext.cpp compiled unmodified for sm_100a
(`make -C oracle ref_ssim`); the wrapper is the reference's fused_ssim/__init__.py, taken from /root/reference or from the
staged copy oracle/_ref/pyref (`make -C oracle ref_py`).  Only tests/ and bench.py's reference arm load this."""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "fused_ssim_cuda.so")
_WRAPPERS = ["/root/reference/submodules/fused-ssim/fused_ssim/__init__.py",
             os.path.join(_HERE, "_ref", "pyref", "submodules", "fused-ssim", "fused_ssim", "__init__.py")]
_cache = {}


def wrapper_path():
    for w in _WRAPPERS:
        if os.path.exists(w):
            return w
    return None


def available():
    return os.path.exists(SO) and wrapper_path() is not None


def extension():
    """The compiled reference extension module (fusedssim, fusedssim_backward)."""
    if "ext" not in _cache:
        import torch  # noqa: F401  (libtorch must be loaded first)
        spec = importlib.util.spec_from_file_location("fused_ssim_cuda", SO)
        ext = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ext)
        _cache["ext"] = ext
    return _cache["ext"]


def load(ext=None, tag="ref"):
    """The reference's fused_ssim package (FusedSSIMMap, fused_ssim) bound to `ext` (default: its own kernels)."""
    if tag in _cache:
        return _cache[tag]
    ext = extension() if ext is None else ext
    saved = sys.modules.get("fused_ssim_cuda")
    sys.modules["fused_ssim_cuda"] = ext
    try:
        spec = importlib.util.spec_from_file_location(f"refpy_fused_ssim_{tag}", wrapper_path())
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            sys.modules.pop("fused_ssim_cuda", None)
        else:
            sys.modules["fused_ssim_cuda"] = saved
    _cache[tag] = mod
    return mod
