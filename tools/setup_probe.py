"""Where does the set-up time of a 512^3 case go?  Times create / geo_pre / index_transform / initialize / first steps /
close of a single-domain case, several times in one process.  MEASUREMENT INFRASTRUCTURE.

  python tools/setup_probe.py [--n 512] [--storage aa] [--precision f64] [--overlap -1|0|1] [--rounds 3]"""
import argparse
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import lattice_boltzmann_method_gpu_b200 as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--storage", default="aa")
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--overlap", type=int, default=-1)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--workload", default="cavity", choices=["cavity", "vessel"])
    a = ap.parse_args()
    st = {"ab": L.STORE_DENSE_AB, "aa": L.STORE_DENSE_AA, "sparse": L.STORE_SPARSE_AB, "sparse_aa": L.STORE_SPARSE_AA}[a.storage]
    torch.cuda.set_device(0)
    prec = L.F64 if a.precision == "f64" else L.F32
    if a.workload == "vessel":
        import bench

        flag, inlet = bench.vessel_inputs(a.n, 0, a.n)
    for r in range(a.rounds):
        if a.workload == "vessel":
            d = bench.vessel_desc(L, a.n, (0, a.n), prec, st, L.MATH_FAST, 0)
        else:
            d = L.case_defaults(L.CASE_LDC)
            d.nx = d.ny = d.nz = a.n
            d.z_begin, d.z_end = 0, a.n
            d.precision, d.storage, d.math, d.device = prec, st, L.MATH_FAST, 0
        t = [time.perf_counter()]
        c = L.Case(d)
        if a.overlap >= 0:
            c.set_option("overlap_launches", a.overlap)
        if a.workload == "vessel":
            t0 = time.perf_counter()
            c.set_flag_slab(flag, 0)
            print(f"  set_flag_slab {1e3 * (time.perf_counter() - t0):.1f} ms (inside 'create' below)")
        t.append(time.perf_counter())
        c.geo_pre(); torch.cuda.synchronize(); t.append(time.perf_counter())
        c.index_transform(); torch.cuda.synchronize(); t.append(time.perf_counter())
        if a.workload == "vessel":
            c.set_bc_planes(inlet, inlet * 0)
        c.initialize(); torch.cuda.synchronize(); t.append(time.perf_counter())
        c.step(1); t.append(time.perf_counter())
        c.step(20); t.append(time.perf_counter())
        c.close(); torch.cuda.synchronize(); t.append(time.perf_counter())
        names = ["create", "geo_pre", "index_transform", "initialize", "step(1)", "step(20)", "close"]
        print(f"round {r} {a.storage} {a.precision} overlap={a.overlap}: " + ", ".join(f"{n} {1e3 * (t[i + 1] - t[i]):.1f} ms" for i, n in enumerate(names)), flush=True)


if __name__ == "__main__":
    main()
