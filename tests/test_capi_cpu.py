"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/lcgs_b200.h
declares, its host-side camera helpers match the oracle, and it refuses to run without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from luisacomputegaussiansplatting_b200 import _capi
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "lcgs_b200.h")).read()
    declared = set(re.findall(r"LCGS_B200_API\s+[A-Za-z_ \*]*?\b(lcgs_b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = C.CDLL(_capi.library_path())
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(_capi.EXPORTED_SYMBOLS)  # the Python binding covers the whole header
    assert _capi.load().lcgs_b200_version() == 200


def test_every_entry_point_cites_the_reference():
    hdr = open(os.path.join(ROOT, "include", "lcgs_b200.h")).read()
    for block in re.findall(r"/\*(?:(?!\*/).)*\*/\s*LCGS_B200_API[^;]*lcgs_b200_(?:sh_process|project|allocate_tiles|"
                            r"scan_inclusive_u32|duplicate_keys|sort_pairs_u64_u32|tile_ranges|blend|splat_forward|"
                            r"render)\b", hdr, flags=re.S):
        assert re.search(r"\.(?:cpp|h|hpp):\d+", block), block[:80]


def test_camera_helpers_match_oracle_bit_for_bit():
    lib = _capi.load()
    for pos, tgt, up, W, H in [((-3.0, -0.5, 3.3), (0.0, 3.0, 0.5), (0.0, -1.0, -1.0), 1920, 1080),
                               ((-3.0, -0.5, 3.3), (0.0, 3.0, 0.5), (0.0, 0.0, 1.0), 800, 800),
                               ((1.0, 2.0, 3.0), (0.5, -4.0, 0.0), (0.0, -1.0, 0.0), 1237, 822)]:
        cam = _capi.Camera()
        assert lib.lcgs_b200_get_lookat_cam(_capi.fvec(pos), _capi.fvec(tgt), _capi.fvec(up), C.byref(cam)) == 0
        cam.aspect_ratio = float(np.float32(W) / np.float32(H))
        cam.width, cam.height = W, H
        vp = _capi.ViewParams()
        assert lib.lcgs_b200_view_params_from_camera(C.byref(cam), C.byref(vp)) == 0
        ovp = orc.view_params(orc.make_camera(pos, tgt, up, W, H))
        assert bytes(vp) == bytes(ovp)


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="box has a GPU")
def test_no_cpu_fallback():
    ctx = C.c_void_p()
    assert _capi.load().lcgs_b200_ctx_create(0, C.byref(ctx)) == _capi.ERR_NO_DEVICE
    assert not ctx.value
    assert b"no CPU fallback" in _capi.load().lcgs_b200_status_string(_capi.ERR_NO_DEVICE)


def test_library_staleness_is_content_based():
    """The shipped .so is matched to the sources by a digest stamp, not by file times (which do not survive the copy
    to the GPU box: ranks of one job once raced to rebuild the library in place)."""
    from luisacomputegaussiansplatting_b200 import build as b

    assert os.path.exists(b.LIB) and os.path.exists(b.STAMP)
    assert not b.is_stale()
    src = os.path.join(b.CSRC, b.SOURCES[0])
    st = os.stat(src)
    try:
        os.utime(src, (st.st_atime, st.st_mtime + 10_000))  # a newer file time alone must not trigger a rebuild
        assert not b.is_stale()
    finally:
        os.utime(src, (st.st_atime, st.st_mtime))
    stamp = open(b.STAMP).read()
    try:
        open(b.STAMP, "w").write("0" * 64 + "\n")              # a different digest must
        assert b.is_stale()
    finally:
        open(b.STAMP, "w").write(stamp)
    assert not b.is_stale()


def test_production_library_has_no_tuning_hooks():
    """Ablation switches and environment-selected kernel geometries exist only in -DLCGS_TUNING builds."""
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", _capi.library_path()], capture_output=True, text=True).stdout
    assert "lcgs_b200_debug_ablate" not in syms and "tuning_env_int" not in syms
    data = open(_capi.library_path(), "rb").read()
    for env in (b"LCGS_SORT_VARIANT", b"LCGS_BLEND_OCC", b"LCGS_REFERENCE_FLOW", b"LCGS_SORT_DEBUG"):
        assert env not in data, env
