"""Generates tests/golden/ldc_order_model.npz: what the oracle's model of the execution order inside
ldc.cu's `update` launch (oracle/lbm_oracle.c, orc_set_ldc_order) gives after the 5119 iterations the
real program ran on a B200 (tests/golden/reference_gpu_outputs.npz, ldc_last_iter), in the same three
mid planes of the VTK velocity block that fixture holds.

  python tests/golden/make_ldc_order_golden.py        (about 3 minutes on 8 cores)

mode 1 / mode 2 bracket the one thing the code does not fix (links whose wall node is bounced in the
same koff iteration by ANOTHER warp): fresh / stale.  Also stored: mode 1 after 300 steps, which the
CPU test recomputes to make sure the committed planes still come from the committed oracle.
TEST INFRASTRUCTURE; runs the oracle only."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)), sys.path.insert(0, str(ROOT / "tests"))
import helpers as H  # noqa: E402
from test_reference_outputs import GOLD, vtk_velocity  # noqa: E402


def planes(V):
    nz, ny, nx = V.shape[:3]
    return V[nz // 2], V[:, ny // 2], V[:, :, nx // 2]


def run(mode, steps):
    o, geo, idx, _ = H.oracle_case("ldc", 64, np.float32)
    o.set_ldc_order(mode)
    o.step(steps)
    _, ux, uy, uz = o.fields()
    return vtk_velocity("ldc", geo.shape, idx, ux, uy, uz)


if __name__ == "__main__":
    k = int(GOLD["ldc_last_iter"])
    out = {"iterations": k, "guard_steps": 300}
    for mode in (0, 1, 2):
        V = run(mode, k)
        for nm, p in zip(("plane_z", "plane_y", "plane_x"), planes(V)):
            out[f"mode{mode}_{nm}"] = p.astype(np.float32)
        print("mode", mode, "done", flush=True)
    for nm, p in zip(("plane_z", "plane_y", "plane_x"), planes(run(1, 300))):
        out[f"guard_{nm}"] = p.astype(np.float32)
    np.savez_compressed(ROOT / "tests" / "golden" / "ldc_order_model.npz", **out)
