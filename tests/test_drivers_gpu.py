"""GPU: the C drivers (drivers/*.c -> C ABI) are drop-ins for the reference programs: same input
files in the working directory, same files under ./out."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import helpers as H
from test_reference_outputs import GOLD, compare

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def write_inputs(wd):
    flag = H.bif_flag()
    (wd / "geo.txt").write_text("".join("%d " % v for v in flag.ravel()))
    bc = np.load(H.GOLDEN / "bif_bc.npy")
    with open(wd / "bc.txt", "w") as f:
        for p in (1, 2, 0):  # profile plane first (the parity fixture)
            f.write("".join("%.6f " % v for v in bc[p].ravel()))
    (wd / "out").mkdir()


def test_bifurcation_driver_matches_reference_files(tmp_path):
    exe = ROOT / "drivers" / "bifurcation"
    if not exe.exists():
        pytest.skip("drivers not built")
    write_inputs(tmp_path)
    r = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ITERATION # 4400" in r.stdout and "#LATTICE65820" in r.stdout  # bifurcation.cu:1272,1282
    for t in (0, 4400):
        assert (tmp_path / "out" / f"bif_{t}.vtk").exists()
    lines = (tmp_path / "out" / "bif_4400.vtk").read_text().split("\n")
    assert lines[:9] == [str(s) for s in GOLD["bif_header"]]
    body = np.array(lines[9].split(), dtype=np.float32).reshape(30, 79, 62, 3)
    compare("bif", body, 2e-5)
    # write_once (bif:1055-1075): u_y, then u_x, of the whole plane z = NZ/2 in lattice units
    meas = np.array((tmp_path / "meas1.txt").read_text().split(), dtype=np.float64).reshape(2, 83, 64)
    c_u = 0.24159041  # bif:20
    assert np.allclose(meas[0][2:81, 1:63] * c_u, body[16 - 1, :, :, 1], rtol=2e-5, atol=1e-9)
    assert np.allclose(meas[1][2:81, 1:63] * c_u, body[16 - 1, :, :, 0], rtol=2e-5, atol=1e-9)
    assert np.abs(meas[0]).max() > 0.05


def test_driver_reports_missing_geometry(tmp_path):
    exe = ROOT / "drivers" / "bifurcation"
    if not exe.exists():
        pytest.skip("drivers not built")
    r = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "geo.txt" in r.stderr


def test_ldc_driver_small(tmp_path):
    exe = ROOT / "drivers" / "ldc"
    if not exe.exists():
        pytest.skip("drivers not built")
    (tmp_path / "out").mkdir()
    r = subprocess.run([str(exe), "--n", "24", "--steps", "300", "--save", "100"], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "out" / "lid_0.vtk").exists() and (tmp_path / "out" / "lid_300.vtk").exists()
    head = (tmp_path / "out" / "lid_300.vtk").read_text().split("\n")[:9]
    assert head[4] == "DIMENSIONS 20 20 20" and head[7] == "POINT_DATA  8000"  # ldc.cu:592-595
    log = (tmp_path / "out" / "CONVERGENCE.log").read_text().split("\n")
    assert log[0] == "1" and log[-2].startswith("TOTAL RUNNING TIME")
