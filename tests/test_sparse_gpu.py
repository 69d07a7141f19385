"""GPU: sparse storage (LBM_STORE_SPARSE_AB) -- populations, moments and node words in the reference's
own compact order (NLATTICE entries), addressed through run-segment records -- against the oracle
(STRICT: bit-exact), against the dense storage, and sharded into z-slabs."""
import numpy as np
import pytest

import helpers as H
from test_slab_gpu import run_slabs

pytestmark = pytest.mark.gpu
CASES = [("ldc", 24), ("ldc", 37), ("pos", 24), ("pos", 40), ("bif", None), ("cor", None)]


def S():
    import lattice_boltzmann_method_gpu_b200 as L

    return L


@pytest.mark.parametrize("name,n", CASES)
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_sparse_strict_fields_bit_exact(name, n, prec):
    L = S()
    dt = np.float32 if prec == "f32" else np.float64
    o, geo, idx, nlat = H.oracle_case(name, n, dt)
    c = H.gpu_case(name, n, L.F32 if prec == "f32" else L.F64, L.MATH_STRICT, storage=L.STORE_SPARSE_AB)
    assert H.gpu_setup(c, name) == nlat
    for nsteps in (1, 2, 40):
        o.step(nsteps)
        c.step(nsteps)
        for r, g, nm in zip(o.fields(), c.get_fields(), ("rho", "ux", "uy", "uz")):
            assert np.array_equal(r, g), f"{name} {prec} {nm} after {c.step_count} steps: {np.abs(r - g).max()}"


@pytest.mark.parametrize("name,n", [("bif", None), ("cor", None), ("pos", 24)])
def test_sparse_populations_are_the_references_d_scr(name, n):
    """in this storage lbm_debug_get_populations IS d_scr: q-major, NLATTICE entries"""
    L = S()
    from oracle import oracle as O

    o, geo, idx, nlat = H.oracle_case(name, n, np.float64)
    c = H.gpu_case(name, n, L.F64, L.MATH_STRICT, storage=L.STORE_SPARSE_AB)
    H.gpu_setup(c, name)
    o.step(11), c.step(11)
    fo, fg = o.populations(), c.get_populations()
    fl = geo.ravel()[geo.ravel() != 0] == 4
    assert np.array_equal(fo[:, fl], fg[:, fl])  # every population of every fluid node
    zz, yy, xx = np.nonzero(geo == 4)
    for q in range(1, 19):  # and every slot of a solid node that a fluid node pulls
        src = idx[zz - O.CZ[q], yy - O.CY[q], xx - O.CX[q]]
        assert np.array_equal(fo[q, src], fg[q, src])


@pytest.mark.parametrize("name,n", [("bif", None), ("ldc", 40)])
def test_sparse_fast_equals_dense_fast(name, n):
    L = S()
    a = H.gpu_case(name, n, L.F64, L.MATH_FAST, storage=L.STORE_SPARSE_AB)
    b = H.gpu_case(name, n, L.F64, L.MATH_FAST, storage=L.STORE_DENSE_AB)
    H.gpu_setup(a, name), H.gpu_setup(b, name)
    a.step(120), b.step(120)
    for x, y in zip(a.get_fields(), b.get_fields()):
        assert np.array_equal(x, y)
    assert abs(a.calc_res() - b.calc_res()) <= 1e-12 * b.calc_res()
    if name == "bif":  # 65820 stored nodes instead of a 64x83x32 box of populations
        assert a.device_bytes < 0.6 * b.device_bytes


@pytest.mark.parametrize("name,n,P", [("bif", None, 4), ("cor", None, 3), ("pos", 24, 2), ("ldc", 20, 5)])
def test_sparse_slabs_equal_single_domain_bitwise(name, n, P):
    L = S()
    steps = 25
    one = H.gpu_case(name, n, L.F64, L.MATH_FAST, storage=L.STORE_SPARSE_AB)
    nlat = H.gpu_setup(one, name)
    one.step(steps)
    ref = one.get_fields()
    cs, total = run_slabs(name, n, P, steps, L.F64, L.MATH_FAST, storage=L.STORE_SPARSE_AB)
    assert total == nlat
    for k in range(4):
        got = np.concatenate([c.get_fields()[k] for c in cs])
        assert np.array_equal(got, ref[k]), f"field {k}"


def test_sparse_pulsatile_and_run_fixed(tmp_path):
    L = S()
    pulse = (0.3, 40.0)
    o, *_ = H.oracle_case("bif", None, np.float32, pulse=pulse)
    c = H.gpu_case("bif", None, L.F32, L.MATH_STRICT, pulse=pulse, storage=L.STORE_SPARSE_AB)
    c.desc.out_dir = str(tmp_path).encode()
    d = c.desc
    c.close()
    c = L.Case(d)
    c.set_flag(H.bif_flag())
    H.gpu_setup(c, "bif")
    c.run_fixed(60, 30, True)  # iterations 0..60, saves at 0, 30, 60
    o.step(61)
    for r, g in zip(o.fields(), c.get_fields()):
        assert np.array_equal(r, g)
    assert sorted(p.name for p in tmp_path.glob("bif_*.vtk")) == ["bif_0.vtk", "bif_30.vtk", "bif_60.vtk"]
