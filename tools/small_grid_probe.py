"""Where the time of the small (64^3) convergence loops goes: per-iteration cost of the fixed loop,
of the residual-every-step loop (lbm_run_converge, no files) and of the VTK dumps.

  python tools/small_grid_probe.py [--n 64]
"""
import argparse
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import lattice_boltzmann_method_gpu_b200 as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    a = ap.parse_args()
    configs = [(L.STORE_DENSE_AB, "dense_ab", -1), (L.STORE_SPARSE_AA, "sparse_aa one launch per step", 0),
               (L.STORE_SPARSE_AA, "sparse_aa persistent", 1)]
    for rule, name in ((L.CASE_LDC, "ldc"), (L.CASE_POISEUILLE, "pos")):
      for storage, sname, persistent in configs:
        for prec in (L.F32, L.F64):
            d = L.case_defaults(rule)
            d.nx = d.ny = d.nz = a.n
            d.z_begin, d.z_end = 0, a.n
            d.precision, d.storage = prec, storage
            out = tempfile.mkdtemp()
            d.out_dir = out.encode()
            c = L.Case(d)
            c.geo_pre()
            c.index_transform()
            c.initialize()
            if persistent >= 0:
                c.set_option("persistent", persistent)
            c.step(50)
            its = 2000
            ms_plain = c.step_timed(its)
            t0 = time.perf_counter()
            c.run_fixed(its, 10 ** 9, False)
            t_fixed = time.perf_counter() - t0
            t0 = time.perf_counter()
            k, _ = c.run_converge(max_it=its, tol=0.0, stag_max=50, time_save=10 ** 9, write_files=False)
            t_conv = time.perf_counter() - t0
            t0 = time.perf_counter()
            for t in range(5):
                c.outputSave(t)
            t_out = (time.perf_counter() - t0) / 5
            print(f"{name} {a.n}^3 {'f32' if prec == L.F32 else 'f64'} {sname}: {c.num_fluid / (ms_plain / its * 1e-3) / 1e6:.0f} MLUPS, step kernel {ms_plain / its * 1e3:.1f} us/step (events), "
                  f"run_fixed {t_fixed / its * 1e6:.1f} us/step, run_converge {t_conv / max(k, 1) * 1e6:.1f} us/iteration ({k} its), "
                  f"one VTK dump {t_out * 1e3:.1f} ms", flush=True)
            c.close()


if __name__ == "__main__":
    main()
