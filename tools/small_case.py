"""One small case, a few launches: the thing to put under ncu when a 64^3 step is slower than it should be.

  python tools/small_case.py [--case ldc|pos|bif] [--n 64] [--precision f32] [--storage sparse_aa] [--persistent -1|0|1] [--overlap 0|1]
                             [--steps 20] [--calls 3]
MEASUREMENT INFRASTRUCTURE."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)), sys.path.insert(0, str(ROOT / "tests"))
import lattice_boltzmann_method_gpu_b200 as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="ldc")
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--precision", default="f32")
    ap.add_argument("--storage", default="sparse_aa")
    ap.add_argument("--persistent", type=int, default=-1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--calls", type=int, default=3)
    ap.add_argument("--overlap", type=int, default=-1, help="lbm_set_option('overlap_launches'): 0 / 1")
    ap.add_argument("--converge", type=int, default=0, help="also time run_converge over this many iterations (no files, never stops early)")
    ap.add_argument("--stag", type=int, default=50)
    ap.add_argument("--tol", type=float, default=0.0)
    ap.add_argument("--no-warm", action="store_true")
    a = ap.parse_args()
    import helpers as H  # case builders only

    st = {"ab": L.STORE_DENSE_AB, "aa": L.STORE_DENSE_AA, "sparse": L.STORE_SPARSE_AB, "sparse_aa": L.STORE_SPARSE_AA}[a.storage]
    c = H.gpu_case(a.case, a.n if a.case in ("ldc", "pos") else None, L.F32 if a.precision == "f32" else L.F64, L.MATH_FAST, storage=st)
    H.gpu_setup(c, a.case)
    if a.persistent >= 0:
        c.set_option("persistent", a.persistent)
    if a.overlap >= 0:
        c.set_option("overlap_launches", a.overlap)
    if not a.no_warm:
        c.step(a.steps)
    for _ in range(0 if a.no_warm else a.calls):
        ms = c.step_timed(a.steps)
        print(f"{a.case} {a.precision} {a.storage} persistent={a.persistent} overlap={a.overlap}: {ms / a.steps * 1e3:.2f} us/step, "
              f"{c.num_fluid * a.steps / (ms * 1e-3) / 1e6:.0f} MLUPS", flush=True)
    if a.converge:
        import time

        for _ in range(2):
            t0 = time.perf_counter()
            its, res = c.run_converge(a.converge - 1, a.tol, a.stag, 10 ** 9, False)
            dt = time.perf_counter() - t0
            print(f"  run_converge stag_max={a.stag} tol={a.tol}: {its} iterations, {dt / its * 1e6:.2f} us/iteration", flush=True)


if __name__ == "__main__":
    main()
