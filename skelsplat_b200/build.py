"""In-tree build of libskelsplat_b200.so: nvcc, sm_100a only, no torch headers.

``python -m skelsplat_b200.build`` or ``__graft_entry__.build()``.  The shared library is a plain
C-ABI library (include/skelsplat_b200.h); it is git-ignored but travels to the GPU box.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
OUT = os.path.join(OUT_DIR, "libskelsplat_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# No --use_fast_math: parity needs IEEE div/sqrt and the full-precision expf.
FLAGS = ["-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def dependencies():
    return sorted(sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
                  glob.glob(os.path.join(HERE, "..", "include", "*.h")))


def source_hash():
    """sha256 over every kernel source / header, in name order: compiled into the library (ssb_source_hash) and compared by
    lib.lib() on load, so a git-ignored .so built from older sources is refused instead of running silently."""
    import hashlib
    h = hashlib.sha256()
    for d in dependencies():
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def source_manifest():
    """"file:hash,file:hash,..." per kernel source / header: compiled into the library (ssb_source_manifest) so that evidence tied
    to ONE kernel (the ncu constants of profiles/*_traffic.json) can be checked against exactly the files that define it."""
    import hashlib
    return ",".join(f"{os.path.basename(d)}:{hashlib.sha256(open(d, 'rb').read()).hexdigest()[:12]}" for d in dependencies())


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in dependencies())


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: tuning variants only (scripts/gpu_tune_opt.py); the product build takes neither."""
    if out is None and not force and not needs_build():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    out = OUT if out is None else out
    cmd = [NVCC] + FLAGS + [f'-DSSB_SOURCE_HASH="{source_hash()}"', f'-DSSB_SOURCE_MANIFEST="{source_manifest()}"'] + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
