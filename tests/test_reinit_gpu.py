"""GPU: a handle may be taken through set_flag .. initialize again with ANOTHER geometry, and every
storage must then size its buffers, host mirrors and checkpoints for the new one; static links of the
in-place storage carry their source's initial population even when that source moves."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _tube_flag(nx, ny, nz, r):
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    return (((x - (nx - 1) / 2) ** 2 + (z - (nz - 1) / 2) ** 2) <= r * r).astype(np.int32)


def _bif_like(storage, flag, L, math=None):
    """GEO_Y_INOUT rule (bifurcation.cu:36-427) on a straight tube along y of radius r"""
    nz, ny, nx = flag.shape
    d = L.case_defaults(L.CASE_GEO_Y_INOUT)
    d.nx, d.ny, d.nz = nx, ny, nz
    d.z_begin, d.z_end = 0, nz
    d.precision, d.math, d.storage = L.F64, L.MATH_STRICT if math is None else math, storage
    c = L.Case(d)
    c.set_flag(flag)
    return c


def _run(c, flag, steps):
    from oracle import oracle as O

    c.set_flag(flag)
    c.geo_pre()
    nlat = c.index_transform()
    nz, ny, nx = flag.shape
    inl = np.full((nz, nx), 0.03, np.float32)
    c.set_bc_planes(inl, np.zeros_like(inl))
    c.initialize()
    c.step(steps)
    geo = O.geo_pre_bif(flag)
    idx, n2 = O.index_transform(geo)
    assert n2 == nlat and np.array_equal(c.get_geo(), geo) and np.array_equal(c.get_index(), idx)
    o = O.Oracle(O.CASE_BIF, geo, idx, nlat, H.TAU_LDC, 0.0, dtype=np.float64)
    o.set_bc_planes(np.where(geo[:, 1, :] == 2, inl, 0).astype(np.float32), np.zeros_like(inl))
    o.initialize()
    o.step(steps)
    for r, g in zip(o.fields(), c.get_fields()):
        assert np.array_equal(r, g)


@pytest.mark.parametrize("storage_name", ["sparse_ab", "sparse_aa", "dense_ab", "dense_aa"])
def test_reinitialise_with_a_larger_mask(storage_name, tmp_path):
    """sparse buffers were sized for the first geometry only"""
    import lattice_boltzmann_method_gpu_b200 as L

    storage = {"dense_ab": L.STORE_DENSE_AB, "dense_aa": L.STORE_DENSE_AA, "sparse_ab": L.STORE_SPARSE_AB,
               "sparse_aa": L.STORE_SPARSE_AA}[storage_name]
    small, big = _tube_flag(40, 30, 40, 6.2), _tube_flag(40, 30, 40, 17.3)
    c = _bif_like(storage, small, L)
    c.desc.out_dir = str(tmp_path).encode()
    _run(c, small, 12)
    b0 = c.device_bytes
    _run(c, big, 12)   # about 8x the stored nodes
    if storage_name.startswith("sparse"):
        assert c.device_bytes > 2 * b0
    _run(c, small, 9)  # and back: fewer nodes than the host mirrors / records of the previous round
    if storage_name.startswith("sparse"):
        assert c.device_bytes <= b0 + 1024


def test_writer_follows_a_new_geometry(tmp_path):
    """the host copy of the index table used by outputSave was never invalidated"""
    import lattice_boltzmann_method_gpu_b200 as L

    small, big = _tube_flag(40, 30, 40, 6.2), _tube_flag(40, 30, 40, 15.1)
    outs = []
    for first in (big, None):  # a handle that saw `big` before, and a fresh one
        c = _bif_like(L.STORE_DENSE_AB, small, L)
        d = tmp_path / ("a" if first is not None else "b")
        d.mkdir()
        c.close()
        c.desc.out_dir = str(d).encode()
        c = L.Case(c.desc)
        if first is not None:
            _run(c, first, 3)
            c.outputSave(3)
        _run(c, small, 5)
        c.outputSave(5)
        outs.append((d / "bif_5.vtk").read_bytes())
    assert outs[0] == outs[1]


@pytest.mark.parametrize("storage_name", ["dense_aa", "dense_ab", "sparse_ab", "sparse_aa"])
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_static_link_from_a_moving_boundary_node(storage_name, prec):
    """a fluid node pulls a direction OUTSIDE the direction set of a label-5 node that was
    initialised with a nonzero velocity (cor.cu:302-306): that slot is a constant, feq_q(1, u0(s))"""
    import lattice_boltzmann_method_gpu_b200 as L

    storage = {"dense_ab": L.STORE_DENSE_AB, "dense_aa": L.STORE_DENSE_AA, "sparse_ab": L.STORE_SPARSE_AB,
               "sparse_aa": L.STORE_SPARSE_AA}[storage_name]
    dt = np.float32 if prec == "f32" else np.float64
    o, geo, idx, nlat = H.oracle_case("corstep", None, dt)
    for math in (L.MATH_STRICT, L.MATH_FAST):
        o, *_ = H.oracle_case("corstep", None, dt)
        c = H.gpu_case("corstep", None, L.F32 if prec == "f32" else L.F64, math, storage=storage)
        assert H.gpu_setup(c, "corstep") == nlat
        for nsteps in (1, 1, 1, 8, 30):
            o.step(nsteps), c.step(nsteps)
            if math == L.MATH_STRICT:
                for r, g in zip(o.fields(), c.get_fields()):
                    assert np.array_equal(r, g), (storage_name, c.step_count)
            else:
                assert H.rel_err(c.get_fields(), o.fields()) < (1e-12 if prec == "f64" else 1e-5)


def test_checkpoint_of_another_geometry_or_tau_is_rejected(tmp_path):
    """same box, same storage, different mask / tau used to load silently"""
    import lattice_boltzmann_method_gpu_b200 as L

    small, big = _tube_flag(40, 30, 40, 6.2), _tube_flag(40, 30, 40, 9.4)
    a = _bif_like(L.STORE_DENSE_AB, small, L)
    _run(a, small, 4)
    a.checkpoint_save(tmp_path / "a.bin")
    b = _bif_like(L.STORE_DENSE_AB, big, L)
    _run(b, big, 4)
    with pytest.raises(L.LbmError, match="geometry"):
        b.checkpoint_load(tmp_path / "a.bin")
    c = _bif_like(L.STORE_DENSE_AB, small, L)
    c.close()
    c.desc.tau = 0.6
    c = L.Case(c.desc)
    c.set_flag(small)
    c.geo_pre(), c.index_transform(), c.initialize()
    with pytest.raises(L.LbmError, match="geometry"):
        c.checkpoint_load(tmp_path / "a.bin")
    a2 = _bif_like(L.STORE_DENSE_AB, small, L)
    _run(a2, small, 1)
    a2.checkpoint_load(tmp_path / "a.bin")  # the matching case still loads
    assert a2.step_count == 4
