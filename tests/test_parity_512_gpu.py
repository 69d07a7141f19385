"""GPU: parity on the BENCHMARKED configuration -- dense lid-driven cavity 512^3, the workload bench.py
times (BASELINE.json configs[2]).  At this size the population offsets q*qstride + c exceed 2^31
elements (19 * 512^3 = 2.55e9; the reference's own `NLATTICE*q+ind` int arithmetic overflows here,
ldc.cu:80), lbm_get_fields stages plane groups, and the launch covers 1 M CTAs -- none of which the
<= 64^3 parity tests reach.  The CPU oracle runs the same case with all host threads
(~60 GB of host memory in fp64, ~35 GB in fp32; the test picks what fits and says so).

Tolerances are north_star's: FAST arithmetic 1e-12 (fp64) / 1e-5 (fp32) of max|u| after N steps;
STRICT arithmetic bit-exact."""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
N = 512


def _avail_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 2 ** 20
    except OSError:
        pass
    return 0.0


def _pick_precision():
    gb = _avail_gb()
    if gb >= 80:
        return "f64"
    if gb >= 48:
        return "f32"
    pytest.skip(f"host has {gb:.0f} GB available; the 512^3 oracle needs 48 GB (fp32) / 80 GB (fp64)")


def test_dense_512_matches_oracle():
    import lattice_boltzmann_method_gpu_b200 as L
    from bench import oracle_threads

    prec = _pick_precision()
    dt, lp, tol = (np.float64, L.F64, 1e-12) if prec == "f64" else (np.float32, L.F32, 1e-5)
    oracle_threads(os.cpu_count() or 1)
    o, geo, idx, nlat = H.oracle_case("ldc", N, dt)
    assert nlat == N ** 3
    ref = {}
    done = 0
    for upto in (3, 4):  # odd and even totals: both phases of the in-place storage
        o.step(upto - done)
        done = upto
        ref[upto] = [a.copy() for a in o.fields()]
    o.close()
    scale = max(float(np.abs(a).max()) for a in ref[4][1:])
    assert scale > 1e-3  # the lid has set the flow in motion
    first = True
    for storage, math in ((L.STORE_DENSE_AA, L.MATH_FAST), (L.STORE_DENSE_AB, L.MATH_FAST), (L.STORE_DENSE_AA, L.MATH_STRICT)):
        c = H.gpu_case("ldc", N, lp, math, storage=storage)
        assert H.gpu_setup(c, "ldc") == nlat
        if first:  # labels and index table of the full box, bit-exact
            assert np.array_equal(c.get_geo(), geo)
            assert np.array_equal(c.get_index(), idx)
            assert c.num_fluid == (N - 4) ** 3
            first = False
        done = 0
        for upto in (3, 4):
            c.step(upto - done)
            done = upto
            got = c.get_fields()
            for nm, g, r in zip(("rho", "ux", "uy", "uz"), got, ref[upto]):
                if math == L.MATH_STRICT:
                    assert np.array_equal(g, r), f"{nm} step {upto} storage {storage}: {np.abs(g - r).max()}"
                else:
                    s = 1.0 if nm == "rho" else scale
                    err = float(np.abs(g - r).max()) / s
                    assert err <= tol, f"{nm} step {upto} storage {storage}: {err:.3e} > {tol}"
        c.close()
