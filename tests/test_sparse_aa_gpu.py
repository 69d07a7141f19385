"""GPU: sparse IN-PLACE storage (LBM_STORE_SPARSE_AA) -- one population buffer in the reference's compact
order, AA-pattern streaming, every boundary link in the fluid node's own slot (csrc/step_sparse_aa.cuh) --
against the oracle (STRICT: bit-exact at odd and even step counts), against the two-buffer sparse storage,
sharded into z-slabs with fused peer stores, through checkpoints, and with the reference's run loops."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
CASES = [("ldc", 24), ("ldc", 37), ("pos", 24), ("pos", 40), ("bif", None), ("cor", None), ("corstep", None)]


def S():
    import lattice_boltzmann_method_gpu_b200 as L

    return L


@pytest.mark.parametrize("name,n", CASES)
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_sparse_aa_strict_fields_bit_exact(name, n, prec):
    L = S()
    dt = np.float32 if prec == "f32" else np.float64
    o, geo, idx, nlat = H.oracle_case(name, n, dt)
    c = H.gpu_case(name, n, L.F32 if prec == "f32" else L.F64, L.MATH_STRICT, storage=L.STORE_SPARSE_AA)
    assert H.gpu_setup(c, name) == nlat
    for nsteps in (1, 1, 1, 2, 5, 30):  # totals 1, 2, 3, 5, 10, 40: both phases
        o.step(nsteps)
        c.step(nsteps)
        for r, g, nm in zip(o.fields(), c.get_fields(), ("rho", "ux", "uy", "uz")):
            assert np.array_equal(r, g), f"{name} {prec} {nm} after {c.step_count} steps: {np.abs(r - g).max()}"


@pytest.mark.parametrize("name,n", [("bif", None), ("cor", None), ("pos", 24), ("ldc", 24)])
@pytest.mark.parametrize("steps", [8, 9])
def test_sparse_aa_populations_match_oracle(name, n, steps):
    """every population a fluid node is about to pull, as the reference's d_scr holds it"""
    L = S()
    from oracle import oracle as O

    o, geo, idx, nlat = H.oracle_case(name, n, np.float64)
    c = H.gpu_case(name, n, L.F64, L.MATH_STRICT, storage=L.STORE_SPARSE_AA)
    H.gpu_setup(c, name)
    o.step(steps), c.step(steps)
    fo, fg = o.populations(), c.get_populations()
    fluid = 3 if name == "ldc" else 4
    zz, yy, xx = np.nonzero(geo == fluid)
    for q in range(19):
        lab = geo[zz - O.CZ[q], yy - O.CY[q], xx - O.CX[q]]
        src = idx[zz - O.CZ[q], yy - O.CY[q], xx - O.CX[q]]
        if name == "ldc":
            src = src[lab != 1]  # ldc's wall slots lag one step (see test_parity_gpu)
        assert np.array_equal(fo[q, src], fg[q, src]), f"direction {q}"


@pytest.mark.parametrize("name,n", [("bif", None), ("ldc", 40), ("corstep", None)])
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_sparse_aa_fast_equals_two_buffer_fast(name, n, prec):
    L = S()
    lp = L.F32 if prec == "f32" else L.F64
    a = H.gpu_case(name, n, lp, L.MATH_FAST, storage=L.STORE_SPARSE_AA)
    b = H.gpu_case(name, n, lp, L.MATH_FAST, storage=L.STORE_SPARSE_AB)
    H.gpu_setup(a, name), H.gpu_setup(b, name)
    a.step(121), b.step(121)
    for x, y in zip(a.get_fields(), b.get_fields()):
        assert np.array_equal(x, y)
    assert a.device_bytes < 0.72 * b.device_bytes  # one population buffer instead of two (+ the shared index arrays)
    assert abs(a.residual(L.RES_VELSUM) - b.residual(L.RES_VELSUM)) <= 1e-6 * abs(b.residual(L.RES_VELSUM))


@pytest.mark.parametrize("name,n,P", [("bif", None, 4), ("cor", None, 3), ("pos", 24, 2), ("ldc", 20, 5), ("ldc", 12, 12), ("corstep", None, 7)])
def test_sparse_aa_slabs_equal_single_domain_bitwise(name, n, P):
    """fused peer stores between slab handles (the only transport of an in-place storage)"""
    L = S()
    from lattice_boltzmann_method_gpu_b200 import slab

    steps = 25
    one = H.gpu_case(name, n, L.F64, L.MATH_FAST, storage=L.STORE_SPARSE_AA)
    nlat = H.gpu_setup(one, name)
    one.step(steps)
    ref = one.get_fields()
    nz = {"ldc": n, "pos": n, "bif": 32, "cor": 44, "corstep": 28}[name]
    cs = [H.gpu_case(name, n, L.F64, L.MATH_FAST, z_range=r, storage=L.STORE_SPARSE_AA) for r in slab.slab_ranges(nz, P)]
    for c in cs:
        c.geo_pre()
    offs, total = slab.compact_offsets([c.local_stored_count() for c in cs])
    assert total == nlat
    for c, o in zip(cs, offs):
        c.set_compact_offset(o, total)
        c.index_transform()
        if name == "bif":
            c.set_bc_planes(*H.bif_bc_planes())
        c.initialize()
    with pytest.raises(L.LbmError):
        cs[0].step_begin(0)  # in place: no transport without peers
    H.attach_virtual_slabs(cs)
    H.step_virtual_slabs(cs, steps)
    for k in range(4):
        got = np.concatenate([c.get_fields()[k] for c in cs])
        assert np.array_equal(got, ref[k]), f"field {k}"


def test_sparse_aa_pulsatile_run_fixed_and_checkpoint(tmp_path):
    L = S()
    pulse = (0.3, 40.0)
    o, *_ = H.oracle_case("bif", None, np.float32, pulse=pulse)
    c = H.gpu_case("bif", None, L.F32, L.MATH_STRICT, pulse=pulse, storage=L.STORE_SPARSE_AA, out_dir=tmp_path)
    H.gpu_setup(c, "bif")
    c.run_fixed(60, 30, True)  # iterations 0..60, saves at 0, 30, 60
    o.step(61)
    for r, g in zip(o.fields(), c.get_fields()):
        assert np.array_equal(r, g)
    assert sorted(p.name for p in tmp_path.glob("bif_*.vtk")) == ["bif_0.vtk", "bif_30.vtk", "bif_60.vtk"]
    c.checkpoint_save(tmp_path / "ck.bin")  # 61 steps done: odd phase
    c.step(14)
    ref = c.get_fields()
    b = H.gpu_case("bif", None, L.F32, L.MATH_STRICT, pulse=pulse, storage=L.STORE_SPARSE_AA)
    H.gpu_setup(b, "bif")
    b.checkpoint_load(tmp_path / "ck.bin")
    b.step(14)
    for x, y in zip(ref, b.get_fields()):
        assert np.array_equal(x, y)


def test_sparse_aa_convergence_loop():
    """ldc.cu:653-685 with the residual fused into the in-place kernels"""
    L = S()
    res = []
    for storage in (L.STORE_SPARSE_AA, L.STORE_DENSE_AB):
        d = L.case_defaults(L.CASE_POISEUILLE)
        d.nx = d.ny = d.nz = 24
        d.z_begin, d.z_end = 0, 24
        d.precision, d.storage = L.F64, storage
        c = L.Case(d)
        c.geo_pre(), c.index_transform(), c.initialize()
        res.append(c.run_converge(4000, 1e-5, 20, 500, False))
    assert abs(res[0][0] - res[1][0]) <= 2 and res[0][0] > 50


def pitted_duct_flag(nx=40, ny=30, nz=22):
    """a rectangular duct along y with isolated one- and two-voxel pits: the rows next to a pit see
    "fluid, solid, fluid" and "fluid, solid, solid, fluid" in a neighbouring row -- the cases in which the
    sources of a run of fluid nodes are NOT one consecutive id range of the storage's own numbering"""
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    flag = ((x >= 3) & (x <= nx - 4) & (z >= 3) & (z <= nz - 4)).astype(np.int32)
    rng = np.random.default_rng(7)
    for _ in range(60):
        x0, y0, z0 = int(rng.integers(6, nx - 8)), int(rng.integers(4, ny - 4)), int(rng.integers(6, nz - 6))
        flag[z0, y0, x0:x0 + int(rng.integers(1, 4))] = 0
    return flag


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_sparse_aa_irregular_rows(prec):
    L = S()
    from oracle import oracle as O

    flag = pitted_duct_flag()
    nz, ny, nx = flag.shape
    geo = O.geo_pre_bif(flag)
    idx, nlat = O.index_transform(geo)
    zz, yy, xx = np.nonzero(geo == 4)
    for q in range(1, 19):  # the reference's guarantee holds: every source of a fluid node is stored
        assert (geo[zz - O.CZ[q], yy - O.CY[q], xx - O.CX[q]] != 0).all()
    row = geo == 4  # the patterns are really there
    assert (row[:, :, :-2] & ~row[:, :, 1:-1] & row[:, :, 2:]).any() and (row[:, :, :-3] & ~row[:, :, 1:-2] & ~row[:, :, 2:-1] & row[:, :, 3:]).any()
    dt = np.float32 if prec == "f32" else np.float64
    inl = np.full((nz, nx), 0.04, np.float32)
    o = O.Oracle(O.CASE_BIF, geo, idx, nlat, H.TAU_LDC, 0.0, dtype=dt)
    o.set_bc_planes(np.where(geo[:, 1, :] == 2, inl, 0).astype(np.float32), np.zeros_like(inl))
    o.initialize()
    cs = []
    for storage in (L.STORE_SPARSE_AA, L.STORE_SPARSE_AB, L.STORE_DENSE_AA):
        d = L.case_defaults(L.CASE_GEO_Y_INOUT)
        d.nz, d.ny, d.nx = flag.shape
        d.z_begin, d.z_end = 0, nz
        d.storage, d.precision, d.math = storage, L.F32 if prec == "f32" else L.F64, L.MATH_STRICT
        c = L.Case(d)
        c.set_flag(flag)
        c.geo_pre()
        assert c.index_transform() == nlat
        c.set_bc_planes(inl, np.zeros_like(inl))
        c.initialize()
        cs.append(c)
    for nsteps in (1, 1, 1, 2, 20):
        o.step(nsteps)
        for c in cs:
            c.step(nsteps)
            for r, g in zip(o.fields(), c.get_fields()):
                assert np.array_equal(r, g), (c.desc.storage, c.step_count)


@pytest.mark.parametrize("name,n", [("ldc", 24), ("bif", None), ("corstep", None)])
def test_persistent_launches_equal_single_launches(name, n):
    """small grids run a whole batch of steps in ONE cooperative launch (grid barrier between steps,
    csrc/step_sparse_aa.cuh k_sparse_aa_persist); same bits as one launch per step, pulsatile table,
    per-step residual slots and odd batch boundaries included"""
    L = S()
    pulse = (0.3, 40.0) if name == "bif" else None
    runs = []
    for persistent in (1, 0):
        c = H.gpu_case(name, n, L.F64, L.MATH_STRICT, storage=L.STORE_SPARSE_AA, pulse=pulse)
        H.gpu_setup(c, name)
        c.set_option("persistent", persistent)
        l0 = c.launch_count
        for nsteps in (1, 2, 7, 64, 131):
            c.step(nsteps)
        nl = c.launch_count - l0  # before get_fields, which gathers with a kernel of its own
        runs.append(([a.copy() for a in c.get_fields()], nl, c.step_count))
    (fa, la, sa), (fb, lb, sb) = runs
    assert sa == sb == 205
    for x, y in zip(fa, fb):
        assert np.array_equal(x, y)
    assert la == 5 and lb == 205  # one launch per lbm_step call against one per step
    o, *_ = H.oracle_case(name, n, np.float64, pulse=pulse)
    o.step(205)
    for r, g in zip(o.fields(), fa):
        assert np.array_equal(r, g)
    with pytest.raises(L.LbmError):
        c.set_option("no_such_option", 1)


def test_persistent_convergence_loop_stops_on_the_same_iteration():
    L = S()
    res = []
    for persistent in (1, 0):
        d = L.case_defaults(L.CASE_LDC)
        d.nx = d.ny = d.nz = 24
        d.z_begin, d.z_end = 0, 24
        d.precision, d.storage = L.F64, L.STORE_SPARSE_AA
        c = L.Case(d)
        c.geo_pre(), c.index_transform(), c.initialize()
        c.set_option("persistent", persistent)
        res.append((c.run_converge(3000, 1e-5, 20, 500, False), c.get_fields()))
    # S is an atomic double sum whose order differs between the two forms: the stopping iteration may move by one
    assert abs(res[0][0][0] - res[1][0][0]) <= 2 and res[0][0][0] > 50
    if res[0][0][0] == res[1][0][0]:
        for x, y in zip(res[0][1], res[1][1]):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("storage_name", ["sparse_aa", "dense_aa"])
@pytest.mark.parametrize("name,n", [("ldc", 40), ("bif", None), ("pos", 32)])
def test_overlapped_launches_equal_serialised_launches(storage_name, name, n):
    """in-place storages launch their step kernels with programmatic stream serialization (the next kernel starts
    while the previous one drains and waits at griddepcontrol.wait before its first population load,
    csrc/step_dense.cuh grid_dep_launch / grid_dep_wait): same bits as ordinary launches and as the oracle, FAST
    arithmetic against itself, STRICT against the oracle; many short launches in a row so that several are in flight"""
    L = S()
    storage = L.STORE_SPARSE_AA if storage_name == "sparse_aa" else L.STORE_DENSE_AA
    for math, prec, dt in ((L.MATH_FAST, L.F32, np.float32), (L.MATH_STRICT, L.F64, np.float64)):
        runs = []
        for overlap in (1, 0):
            c = H.gpu_case(name, n, prec, math, storage=storage)
            H.gpu_setup(c, name)
            c.set_option("overlap_launches", overlap)
            for nsteps in (1, 2, 50, 151):
                c.step(nsteps)
            runs.append([a.copy() for a in c.get_fields()])
            c.close()
        for x, y in zip(*runs):
            assert np.array_equal(x, y)
        if math == L.MATH_STRICT:
            o, *_ = H.oracle_case(name, n, dt)
            o.step(204)
            for r, g in zip(o.fields(), runs[0]):
                assert np.array_equal(r, g)


@pytest.mark.parametrize("name,n,tol,stag", [("ldc", 24, 1e-5, 20), ("pos", 24, 1e-5, 7)])
def test_convergence_loop_leaves_the_state_of_its_last_iteration(name, n, tol, stag):
    """run_converge launches its steps in batches sized so that the stopping rule (ldc.cu:653-685) can never be met
    inside one; the state afterwards IS the one after exactly `its` steps, stepping goes on from it correctly, and the
    in-place sparse storage stops where the two-buffer storage stops (S is an atomic sum in another order: +-2)"""
    L = S()
    out = {}
    for storage in (L.STORE_SPARSE_AA, L.STORE_DENSE_AB):
        c = H.gpu_case(name, n, L.F64, L.MATH_STRICT, storage=storage)
        H.gpu_setup(c, name)
        its, res = c.run_converge(4000, tol, stag, 500, False)
        assert c.step_count == its
        f_stop = [a.copy() for a in c.get_fields()]
        c.step(5)
        out[storage] = (its, res, f_stop, [a.copy() for a in c.get_fields()])
        c.close()
    (ia, ra, fa, fa5), (ib, rb, fb, fb5) = out[L.STORE_SPARSE_AA], out[L.STORE_DENSE_AB]
    assert 60 < ia < 4000 and abs(ia - ib) <= 2
    c = H.gpu_case(name, n, L.F64, L.MATH_STRICT, storage=L.STORE_SPARSE_AA)
    H.gpu_setup(c, name)
    c.step(ia)
    for x, y in zip(c.get_fields(), fa):
        assert np.array_equal(x, y)
    c.step(5)
    for x, y in zip(c.get_fields(), fa5):
        assert np.array_equal(x, y)
    c.close()
    if ia == ib:
        assert ra == rb
        for x, y in zip(fa, fb):
            assert np.array_equal(x, y)
