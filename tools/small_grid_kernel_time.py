"""20 steps of the 64^3 cavity and the bifurcation (for an ncu launch list: how long is the step kernel itself).
MEASUREMENT / TEST INFRASTRUCTURE (like tests/): it may run the compiled reference in oracle/_ref or use the
test helpers; nothing here is part of, or imported by, the product package.
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import helpers as H  # noqa: E402
import lattice_boltzmann_method_gpu_b200 as L  # noqa: E402

for name, n in (("ldc", 64), ("bif", None)):
    for st in (L.STORE_DENSE_AB, L.STORE_SPARSE_AB):
        c = H.gpu_case(name, n, L.F32, L.MATH_FAST, storage=st)
        H.gpu_setup(c, name)
        c.step(20)
        print(name, st, c.step_timed(200) / 200 * 1e3, "us/step")
        c.close()
