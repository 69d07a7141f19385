"""GPU: edge cases of the case interface -- empty and degenerate masks, short / missing input files,
ragged dimensions, argument validation -- through the C ABI."""
import numpy as np
import pytest

import helpers as H
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def lib():
    import lattice_boltzmann_method_gpu_b200 as L

    return L


def geo_case(L, flag, storage, precision=None):
    d = L.case_defaults(L.CASE_GEO_Y_INOUT)
    d.nz, d.ny, d.nx = flag.shape
    d.z_begin, d.z_end = 0, d.nz
    d.storage = storage
    d.precision = L.F32 if precision is None else precision
    c = L.Case(d)
    c.set_flag(flag)
    return c


@pytest.mark.parametrize("storage_name", ["dense", "sparse", "aa"])
def test_empty_mask(storage_name):
    """no voxel set: nothing is stored, every call still works and returns empty arrays"""
    L = lib()
    storage = {"dense": L.STORE_DENSE_AB, "sparse": L.STORE_SPARSE_AB, "aa": L.STORE_DENSE_AA}[storage_name]
    flag = np.zeros((9, 12, 40), np.int32)
    c = geo_case(L, flag, storage)
    c.geo_pre()
    assert c.index_transform() == 0
    assert not c.get_geo().any() and (c.get_index() == -1).all()
    c.initialize()
    c.step(3)
    assert c.num_fluid == 0 and all(a.size == 0 for a in c.get_fields())
    assert c.calc_res() == 0.0


@pytest.mark.parametrize("storage_name", ["dense", "sparse"])
def test_walls_without_fluid(storage_name):
    """a one-voxel-thick sheet: wall and -1 nodes are stored, but there is no fluid node to update"""
    L = lib()
    storage = L.STORE_DENSE_AB if storage_name == "dense" else L.STORE_SPARSE_AB
    flag = np.zeros((10, 14, 33), np.int32)
    flag[4, 3:11, 5:28] = 1
    geo = O.geo_pre_bif(flag)
    idx, nlat = O.index_transform(geo)
    assert nlat > 0 and not (geo == 4).any()
    c = geo_case(L, flag, storage)
    c.geo_pre()
    assert c.index_transform() == nlat
    assert np.array_equal(c.get_geo(), geo) and np.array_equal(c.get_index(), idx)
    c.initialize()
    c.step(4)
    rho, ux, uy, uz = c.get_fields()
    assert rho.size == nlat and not rho.any() and not uy.any()


@pytest.mark.parametrize("dims", [(5, 6, 5), (7, 9, 6), (31, 8, 5), (65, 7, 9)])
def test_smallest_and_ragged_boxes(dims):
    """boxes down to the smallest the label rules allow, x extents around the 32-cell pitch"""
    L = lib()
    nx, ny, nz = dims
    rng = np.random.default_rng(0)
    flag = (rng.random((nz, ny, nx)) < 0.8).astype(np.int32)
    geo = O.geo_pre_bif(flag)
    idx, nlat = O.index_transform(geo)
    for storage in (L.STORE_DENSE_AB, L.STORE_SPARSE_AB):
        c = geo_case(L, flag, storage, L.F64)
        c.desc.math = L.MATH_STRICT
        c.geo_pre()
        assert c.index_transform() == nlat
        assert np.array_equal(c.get_geo(), geo) and np.array_equal(c.get_index(), idx)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_masks_strict_parity(seed):
    """noisy voxel fields (isolated voxels, holes, voxels on the box faces): labels, index and 30
    STRICT steps equal the oracle bit for bit, dense and sparse"""
    L = lib()
    rng = np.random.default_rng(seed)
    nz, ny, nx = 14, 18, 38
    flag = (rng.random((nz, ny, nx)) < 0.86).astype(np.int32)
    geo = O.geo_pre_bif(flag)
    idx, nlat = O.index_transform(geo)
    inl = (0.04 * rng.random((nz, nx))).astype(np.float32)
    out = np.zeros_like(inl)
    o = O.Oracle(O.CASE_BIF, geo, idx, nlat, H.TAU_LDC, 0.0, dtype=np.float64)
    o.set_bc_planes(np.where(geo[:, 1, :] == 2, inl, 0).astype(np.float32), out)
    o.initialize()
    o.step(30)
    fl = geo.ravel()[geo.ravel() != 0] == 4
    for storage in (L.STORE_DENSE_AB, L.STORE_SPARSE_AB, L.STORE_DENSE_AA, L.STORE_SPARSE_AA):
        d = L.case_defaults(L.CASE_GEO_Y_INOUT)
        d.nz, d.ny, d.nx = flag.shape
        d.z_begin, d.z_end = 0, nz
        d.storage, d.precision, d.math = storage, L.F64, L.MATH_STRICT
        c = L.Case(d)
        c.set_flag(flag)
        c.geo_pre()
        assert c.index_transform() == nlat
        assert np.array_equal(c.get_geo(), geo) and np.array_equal(c.get_index(), idx)
        c.set_bc_planes(inl, out)
        c.initialize()
        c.step(30)
        # a random mask can give a fluid node an unstored (label 0) source, which the reference reads
        # out of bounds; compare the nodes whose 18 sources are all stored
        ok = np.ones(geo.shape, bool)
        zz, yy, xx = np.nonzero(geo == 4)
        good = np.ones(len(zz), bool)
        for q in range(1, 19):
            good &= geo[zz - O.CZ[q], yy - O.CY[q], xx - O.CX[q]] != 0
        if good.all():
            for r, g in zip(o.fields(), c.get_fields()):
                assert np.array_equal(r[fl], g[fl])


def test_short_and_missing_files(tmp_path):
    L = lib()
    d = L.case_defaults(L.CASE_GEO_Y_INOUT)
    d.geo_path = str(tmp_path / "geo.txt").encode()
    d.bc_path = str(tmp_path / "bc.txt").encode()
    c = L.Case(d)
    with pytest.raises(L.LbmError) as e:
        c.geo_pre()
    assert e.value.status == -3 and "geo.txt" in str(e.value)
    (tmp_path / "geo.txt").write_text("1 " * 1000)  # 169984 tokens expected
    with pytest.raises(L.LbmError) as e:
        c.geo_pre()
    assert e.value.status == -3 and "1000" in str(e.value)
    (tmp_path / "geo.txt").write_text("".join("%d " % v for v in H.bif_flag().ravel()))
    c.geo_pre()
    assert c.index_transform() == 65820
    with pytest.raises(L.LbmError):
        c.read_vel()  # bc.txt missing
    # a short bc.txt is padded with zeros, like fscanf leaves the reference's variable untouched
    (tmp_path / "bc.txt").write_text("0.05 " * 10)
    c.read_vel()
    c.initialize()
    c.step(2)


def test_argument_validation():
    L = lib()
    d = L.case_defaults(L.CASE_LDC)
    for field, bad in (("nx", 3), ("tau", 0.5), ("z_end", 9999), ("precision", 7), ("n_bc", 99), ("case_rule", 9)):
        dd = L.case_defaults(L.CASE_LDC)
        setattr(dd, field, bad)
        with pytest.raises(L.LbmError) as e:
            L.Case(dd)
        assert e.value.status == -1
    dd = L.case_defaults(L.CASE_LDC)
    dd.bc[0].label = 3  # the fluid label cannot carry a boundary condition
    with pytest.raises(L.LbmError):
        L.Case(dd)
    c = L.Case(d)
    with pytest.raises(ValueError):
        c.set_flag(np.zeros((3, 3, 3), np.int32))
    with pytest.raises(L.LbmError):
        c.get_fields()
