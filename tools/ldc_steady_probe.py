"""How far is the LDC 64^3 field at the iteration where the reference stops (k=5119) from the true
steady state, and how far is the reference's own output from both?  (development probe)"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import helpers as H
import lattice_boltzmann_method_gpu_b200 as L
from test_reference_outputs import GOLD, vtk_velocity

def planes(V):
    nz, ny, nx = V.shape[:3]
    return V[nz // 2], V[:, ny // 2], V[:, :, nx // 2]

for prec in (L.F32, L.F64):
    c = H.gpu_case("ldc", 64, prec, L.MATH_FAST)
    H.gpu_setup(c, "ldc")
    geo, idx = c.get_geo(), c.get_index()
    out = {}
    done = 0
    for k in (5119, 5121, 10000, 20000, 40000):
        c.step(k - done); done = k
        rho, ux, uy, uz = c.get_fields()
        out[k] = planes(vtk_velocity("ldc", geo.shape, idx, ux, uy, uz))
    ref = (GOLD["ldc_plane_z"], GOLD["ldc_plane_y"], GOLD["ldc_plane_x"])
    sc = float(GOLD["ldc_max_abs"])
    d = lambda a, b: max(float(np.abs(x - y).max()) for x, y in zip(a, b)) / sc
    print("precision", "f32" if prec == L.F32 else "f64")
    for k in out:
        print(f"  |ours({k}) - reference(5119)| = {d(out[k], ref):.3e}   |ours({k}) - ours(40000)| = {d(out[k], out[40000]):.3e}")
