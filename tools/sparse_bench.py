"""MLUPS of the sparse storage (reference compact order + run segments) against the dense-box storage
on a synthetic vessel bundle: K x K bent tubes along y in an N^3 box, inlet at y=1 (parabolic profile
per tube, optionally pulsatile), pressure outlet at y=NY-2 -- the GEO_Y_INOUT (bifurcation.cu) rule.

  python tools/sparse_bench.py [--n 512] [--k 4] [--radius-frac 0.38] [--steps 50] [--precision f64]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import lattice_boltzmann_method_gpu_b200 as L  # noqa: E402


def tube_bundle(n, k, rfrac, z0=0, z1=None):
    """binary mask [z][y][x] of planes z0..z1: k*k tubes, centre lines bent sinusoidally in x"""
    z1 = n if z1 is None else z1
    pitch = n / k
    r = rfrac * pitch
    z, y, x = np.meshgrid(np.arange(z0, z1, dtype=np.float32), np.arange(n, dtype=np.float32),
                          np.arange(n, dtype=np.float32), indexing="ij", sparse=True)
    bend = 0.08 * pitch * np.sin(2 * np.pi * y / n)
    dx = (x - bend) % pitch - pitch / 2
    dz = z % pitch - pitch / 2
    return ((dx * dx + dz * dz) <= r * r).astype(np.int32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--k", type=int, default=4)
    ap.add_argument("--radius-frac", type=float, default=0.38)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    n = a.n
    flag = tube_bundle(n, a.k, a.radius_frac)
    pitch = n / a.k
    zz, xx = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    rr2 = ((xx % pitch - pitch / 2) ** 2 + (zz % pitch - pitch / 2) ** 2) / (a.radius_frac * pitch) ** 2
    inlet = (0.05 * np.clip(1 - rr2, 0, None)).astype(np.float32)
    out = {}
    for name, storage in (("dense_ab", L.STORE_DENSE_AB), ("dense_aa", L.STORE_DENSE_AA), ("sparse_ab", L.STORE_SPARSE_AB),
                          ("sparse_aa", L.STORE_SPARSE_AA)):
        if a.only and name not in a.only.split(","):
            continue
        d = L.case_defaults(L.CASE_GEO_Y_INOUT)
        d.nx = d.ny = d.nz = n
        d.z_begin, d.z_end = 0, n
        d.precision = L.F64 if a.precision == "f64" else L.F32
        d.storage = storage
        d.pulse_amp, d.pulse_period = 0.3, 200.0
        d.bc[0].pulsatile = 1
        c = L.Case(d)
        t0 = time.time()
        c.set_flag(flag)
        c.geo_pre()
        nlat = c.index_transform()
        c.set_bc_planes(inlet, np.zeros_like(inlet))
        c.initialize()
        setup = time.time() - t0
        c.step(5)
        ms = c.step_timed(a.steps)
        bpl = 304 if a.precision == "f64" else 152
        out[name] = {"mlups": c.num_fluid * a.steps / (ms * 1e-3) / 1e6, "ms_per_step": ms / a.steps,
                     "algorithmic_GBps": c.num_fluid * bpl / (ms / a.steps * 1e-3) / 1e9,
                     "device_GB": c.device_bytes / 1e9, "setup_s": setup, "nlattice": nlat, "fluid": c.num_fluid,
                     "fill": c.num_fluid / n ** 3}
        fields = c.get_fields()
        out[name]["checksum_uy"] = float(np.abs(fields[2].astype(np.float64)).sum())
        c.close()
    out["same_fields"] = len({v["checksum_uy"] for v in out.values()}) == 1
    peak = 6549.1
    try:
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass
    for v in list(out.values()):
        if isinstance(v, dict):
            v["frac_of_measured_peak"] = v["algorithmic_GBps"] / peak
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
