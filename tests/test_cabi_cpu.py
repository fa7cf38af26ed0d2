"""CPU-side checks: the C-ABI library loads, exports every symbol include/vils_cabi.h declares, and refuses to compute
without a GPU (there is no CPU fallback).  No compute calls are made here."""
import ctypes as C
import os
import re

import pytest

from mvil_fusion_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vils_cabi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vils_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mvil_fusion_b200 import build
    so = build.build()
    L = C.CDLL(so)
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert L.vils_abi_version() == 2


def test_host_mirror_library_builds_and_exports_the_reference_api():
    from mvil_fusion_b200 import build
    build.build()
    H = C.CDLL(build.HOST_OUT)
    for sym in ("vh_estimator_create", "vh_process_imu", "vh_process_image", "vh_tracker_read", "vh_transform_to_end"):
        assert hasattr(H, sym), sym
    hdr = open(os.path.join(ROOT, "mvil_fusion_b200", "csrc", "host", "vils_host.h")).read()
    for name in ("processIMU", "processImage", "optimization", "slideWindow", "readImage", "TransformToEnd"):   # the reference's method names
        assert name in hdr


def test_struct_layouts_match_header_sizes():
    assert C.sizeof(cabi.VilsPreint) == 467 * 8
    assert C.sizeof(cabi.VilsSummary) == 32
    assert C.sizeof(cabi.VilsSolveOpts) == 56
    assert C.sizeof(cabi.VilsIcp) == 10 * 8 + 16
    assert C.sizeof(cabi.VilsLps) == 7 * 8 + 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mvil_fusion_b200 import lib
    with pytest.raises(lib.VilsError) as e:
        lib.BA()
    assert e.value.code == cabi.VILS_ERR_NO_DEVICE
    import numpy as np
    with pytest.raises(lib.VilsError):
        lib.deskew(np.zeros((4, 8), np.float32), 8, [0, 0, 0, 1], [0, 0, 0], 10.0, 0.5, 70.0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mvil_fusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "vils_oracle" not in src and "oracle_lib" not in src and "oracle/" not in src, os.path.join(dirpath, f)
