"""The step kernels under a memory / race checker: every storage, both phases of the in-place
storage, the speculative pull (which deliberately reads past rows into guard cells), boundary slow
paths of all four case rules, and virtual z-slabs whose face launches store into another handle's
buffers.  No torch, no oracle: only liblbm_b200.so through ctypes, so every kernel the tool sees is ours.

  python tools/selfcheck.py          (the library's own self-checking build; what the GPU pool allows)
  compute-sanitizer --tool memcheck  --error-exitcode 7 python tools/sanitize_cases.py
  LBM_SPECULATIVE=1 compute-sanitizer --tool memcheck ... (forces the speculative pull, which small grids do not pick)
  compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_cases.py quick
(tools/run_sanitizer.sh runs both and keeps the logs under gpurun_out/sanitizer/)
MEASUREMENT / TEST INFRASTRUCTURE."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)), sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402

import lattice_boltzmann_method_gpu_b200 as L  # noqa: E402
from lattice_boltzmann_method_gpu_b200 import slab  # noqa: E402

TOTAL = [0, 0, 0]  # out-of-range accesses, doubly touched elements, launches checked (self-checking build only)


def tally(c, what):
    try:
        oob, races, launches = c.selfcheck()
    except L.LbmError:
        return
    TOTAL[0] += oob
    TOTAL[1] += races
    TOTAL[2] += launches
    print(f"   selfcheck {what}: out_of_range={oob} touched_by_two_threads={races} launches_checked={launches}", flush=True)



def cases(quick):
    import helpers as H  # builds cases only; the oracle functions it also holds are not called here

    storages = [("dense_ab", L.STORE_DENSE_AB), ("dense_aa", L.STORE_DENSE_AA), ("sparse_ab", L.STORE_SPARSE_AB)]
    extra = [(n, getattr(L, n)) for n in ("STORE_SPARSE_AA",) if hasattr(L, n)]
    storages += [(n.lower(), v) for n, v in extra]
    names = [("ldc", 24), ("bif", None)] if quick else [("ldc", 24), ("ldc", 33), ("pos", 24), ("bif", None), ("corstep", None)]
    for sname, st in storages:
        for name, n in names:
            for prec in ((L.F32,) if quick else (L.F32, L.F64)):
                yield sname, st, name, n, prec, H


def single(quick):
    for sname, st, name, n, prec, H in cases(quick):
        c = H.gpu_case(name, n, prec, L.MATH_FAST, storage=st, pulse=(0.3, 40.0) if name == "bif" else None)
        H.gpu_setup(c, name)
        c.step(3)   # even, odd, even
        c.step(4)
        f = c.get_fields()
        assert np.isfinite(f[0]).all() and float(np.abs(f[2]).max() + np.abs(f[3]).max()) > 0
        c.residual(L.RES_VELSUM)
        c.get_populations()
        tally(c, f"{sname} {name}")
        c.close()
        print("ok single", sname, name, n, "f64" if prec == L.F64 else "f32", flush=True)


def slabs(quick):
    import helpers as H

    storages = [L.STORE_DENSE_AB, L.STORE_DENSE_AA, L.STORE_SPARSE_AB] + ([L.STORE_SPARSE_AA] if hasattr(L, "STORE_SPARSE_AA") else [])
    for st in storages:
        for name, n, P in ((("ldc", 20, 3),) if quick else (("ldc", 20, 3), ("bif", None, 4))):
            nz = n if name == "ldc" else 32
            cs = [H.gpu_case(name, n, L.F64, L.MATH_FAST, z_range=r, storage=st) for r in slab.slab_ranges(nz, P)]
            for c in cs:
                c.geo_pre()
            offs, total = slab.compact_offsets([c.local_stored_count() for c in cs])
            for c, o in zip(cs, offs):
                c.set_compact_offset(o, total)
                c.index_transform()
                if name == "bif":
                    c.set_bc_planes(*H.bif_bc_planes())
                c.initialize()
            H.attach_virtual_slabs(cs)
            H.step_virtual_slabs(cs, 5)
            for c in cs:
                f = c.get_fields()
                assert np.isfinite(f[0]).all()
            for r, c in enumerate(cs):
                tally(c, f"slab {r} of {P} storage {st} {name}")
                c.close()
            print("ok slabs", st, name, P, flush=True)


if __name__ == "__main__":
    q = len(sys.argv) > 1 and sys.argv[1] == "quick"
    single(q)
    slabs(q)
    print(f"sanitize_cases done: out_of_range={TOTAL[0]} touched_by_two_threads={TOTAL[1]} launches_checked={TOTAL[2]}")
    sys.exit(1 if TOTAL[0] or TOTAL[1] else 0)
