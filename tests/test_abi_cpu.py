"""CPU: the C-ABI library loads, exports every symbol include/lbm_b200.h declares, and refuses to
run without a GPU (there is no CPU fallback)."""
import re
import subprocess
from pathlib import Path

import pytest

import lattice_boltzmann_method_gpu_b200 as L

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    txt = (ROOT / "include" / "lbm_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(L.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    out = subprocess.run(["nm", "-D", "--defined-only", str(L.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (lbm_[a-z0-9_]+)", out))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing


def test_library_loads_and_defaults_match_reference_constants():
    lib = L.load_library()
    for s in L.ABI_SYMBOLS:
        assert hasattr(lib, s)
    d = L.case_defaults(L.CASE_GEO_Y_INOUT)
    assert (d.nx, d.ny, d.nz) == (64, 83, 32)          # bifurcation.cu:19
    assert abs(d.tau - 0.55) < 1e-7 and d.out_name == b"bif"
    d = L.case_defaults(L.CASE_LDC)
    assert (d.nx, d.ny, d.nz) == (64, 64, 64) and abs(d.u_max - 0.15 / 2.4705) < 1e-8   # ldc.cu:48-52
    d = L.case_defaults(L.CASE_POISEUILLE)
    assert abs(d.tau - 0.58) < 1e-7 and abs(d.bc[0].value - 0.09714700668) < 1e-8          # pos.cu:39,590
    d = L.case_defaults(L.CASE_GEO_OPENINGS)
    assert (d.nx, d.ny, d.nz) == (291, 291, 372) and d.n_openings == 5 and d.geo_yfast == 1  # cor.cu:19,45-141


def test_sm100a_code_is_embedded():
    out = subprocess.run(["cuobjdump", "-lelf", str(L.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    d = L.case_defaults(L.CASE_LDC)
    with pytest.raises(L.LbmError) as e:
        L.Case(d)
    assert e.value.status == -6  # LBM_ERR_NO_DEVICE


def test_bad_descriptor_is_rejected():
    import ctypes as C

    lib = L.load_library()
    d = L.case_defaults(L.CASE_LDC)
    d.struct_size = 12
    h = C.c_void_p()
    assert lib.lbm_create(C.byref(d), C.byref(h)) == -1
    assert b"size mismatch" in lib.lbm_last_error(None)


def test_voxeliser_argument_errors_need_no_gpu(tmp_path):
    """the front end checks its arguments and reads the surface file before it touches the device, and
    -- like the solver -- has no CPU fallback for the voxelisation itself"""
    import numpy as np
    import torch

    grid = ((0.0, 0.0, 0.0), 0.5, (8, 8, 8))
    with pytest.raises(L.LbmError) as e:
        L.voxelize(tmp_path / "missing.stl", *grid)
    assert "missing.stl" in str(e.value)
    (tmp_path / "junk.stl").write_bytes(b"solid nothing\nendsolid\n")
    with pytest.raises(L.LbmError):
        L.voxelize(tmp_path / "junk.stl", *grid)
    tri = np.zeros((1, 3, 3), np.float32)
    for bad in (dict(spacing=0.0), dict(dims=(0, 8, 8)), dict(z_range=(4, 9)), dict(z_range=(5, 5))):
        args = dict(origin=(0.0, 0.0, 0.0), spacing=0.5, dims=(8, 8, 8))
        args.update(bad)
        with pytest.raises(L.LbmError) as e:
            L.voxelize(tri, **args)
        assert e.value.status == -1  # LBM_ERR_ARG
    if not torch.cuda.is_available():
        with pytest.raises(L.LbmError) as e:
            L.voxelize(tri, *grid)
        assert e.value.status in (-4, -6)  # LBM_ERR_CUDA / LBM_ERR_NO_DEVICE: no silent CPU path


def test_handle_calls_reject_null():
    lib = L.load_library()
    assert lib.lbm_set_output_format(None, 0) != 0
    assert lib.lbm_output_save(None, 0) != 0
    assert lib.lbm_step(None, 1) != 0
