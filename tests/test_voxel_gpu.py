"""GPU voxeliser (csrc/lbm_voxel.cu, through the C ABI) against its CPU restatement, bit for bit, and
as the front end of a run: surface -> mask -> geo_pre -> flow."""
import sys
from pathlib import Path

import numpy as np
import pytest

import helpers as H

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


MESHES = {
    "box": lambda: H.mesh_box((1.0, 1.0, 1.0), (5.0, 5.0, 5.0)),
    "sphere": lambda: H.mesh_sphere((8.1, 7.9, 8.3), 6.0, nu=96, nv=48),
    "tube": lambda: H.mesh_tube(20.0, 3.0, bend=1.5, x0=6.0, z0=5.0),
    "ragged": lambda: np.delete(H.mesh_tube(20.0, 3.0, x0=6.0, z0=5.0), np.arange(10240 - 40, 10240 - 10), axis=0),
}


@pytest.mark.parametrize("mesh,origin,h,dims", [
    ("box", (0.0, 0.0, 0.0), 0.5, (12, 12, 12)),          # rays exactly through shared edges
    ("box", (0.25, 0.25, 0.25), 0.5, (12, 12, 12)),       # centres exactly on faces
    ("sphere", (0.0, 0.0, 0.0), 0.31, (53, 53, 53)),
    ("sphere", (-3.0, 2.0, 5.0), 0.173, (120, 40, 33)),   # surface leaves the grid on several sides
    ("tube", (0.0, 0.0, 0.0), 0.25, (48, 80, 40)),
    ("ragged", (0.0, 0.0, 0.0), 0.25, (48, 80, 40)),      # rows with an odd number of crossings are cleared
    ("tube", (0.0, 0.0, 0.0), 0.01, (1100, 6, 5)),        # rows longer than 1024 voxels (several word groups)
])
def test_gpu_equals_oracle(mesh, origin, h, dims):
    import lattice_boltzmann_method_gpu_b200 as L

    tri = MESHES[mesh]()
    if dims[0] == 1100:
        origin = (2.0, 9.0, 4.9)
    ref = O.voxelize(tri, origin, h, dims)
    got = L.voxelize(tri, origin, h, dims)
    assert got.shape == ref.shape and got.dtype == np.uint8
    assert np.array_equal(got, ref)
    assert ref.sum() > 0
    # a z-range on its own (what one rank of a slab run computes)
    z0, z1 = dims[2] // 3, dims[2] // 3 + max(1, dims[2] // 4)
    assert np.array_equal(L.voxelize(tri, origin, h, dims, z_range=(z0, z1)), ref[z0:z1])


def test_stl_files_and_errors(tmp_path):
    import lattice_boltzmann_method_gpu_b200 as L

    tri = MESHES["tube"]()
    grid = ((0.0, 0.0, 0.0), 0.25, (48, 80, 40))
    ref = O.voxelize(tri, *grid)
    H.write_binary_stl(tmp_path / "t.stl", tri)
    H.write_ascii_stl(tmp_path / "t_ascii.stl", tri)
    assert np.array_equal(L.voxelize(tmp_path / "t.stl", *grid), ref)
    assert np.array_equal(L.voxelize(tmp_path / "t_ascii.stl", *grid), ref)
    with pytest.raises(L.LbmError):
        L.voxelize(tmp_path / "missing.stl", *grid)
    (tmp_path / "junk.stl").write_bytes(b"not a surface")
    with pytest.raises(L.LbmError):
        L.voxelize(tmp_path / "junk.stl", *grid)
    with pytest.raises(L.LbmError):
        L.voxelize(tri, (0, 0, 0), 0.25, (48, 80, 40), z_range=(10, 50))
    assert L.voxelize(np.zeros((0, 3, 3), np.float32), *grid).sum() == 0


def test_surface_to_flow():
    """front end to back end: a tube surface becomes the voxel field, geo_pre labels it by the
    bifurcation rules (inlet at y=1, outlet at y=NY-2), and a parabolic inlet drives a flow through it"""
    import lattice_boltzmann_method_gpu_b200 as L

    nx, ny, nz = 48, 64, 40
    h = 0.25
    tri = H.mesh_tube(ny * h, 3.0, bend=1.0, x0=6.0, z0=5.0)
    flag = L.voxelize(tri, (0.0, 0.0, 0.0), h, (nx, ny, nz))
    d = L.case_defaults(L.CASE_GEO_Y_INOUT)
    d.nx, d.ny, d.nz = nx, ny, nz
    d.z_begin, d.z_end = 0, nz
    d.precision, d.storage = L.F64, L.STORE_SPARSE_AB
    c = L.Case(d)
    c.set_flag(flag.astype(np.int32))
    c.geo_pre()
    nlat = c.index_transform()
    geo = c.get_geo()
    assert (geo == 2).sum() > 0 and (geo == 3).sum() > 0 and (geo == 4).sum() == c.num_fluid
    assert nlat == int((geo != 0).sum())
    zz, xx = np.meshgrid(np.arange(nz), np.arange(nx), indexing="ij")
    r2 = ((xx + 0.5) * h - 6.0) ** 2 + ((zz + 0.5) * h - 5.0) ** 2
    inlet = (0.04 * np.clip(1 - r2 / 9.0, 0, None)).astype(np.float32)
    c.set_bc_planes(inlet, np.zeros_like(inlet))
    c.initialize()
    c.step(400)
    rho, ux, uy, uz = c.get_fields()
    idx = c.get_index()
    mid = idx[:, ny // 2, :]
    v = uy[mid[mid >= 0]]
    assert np.isfinite(rho).all() and abs(rho[rho > 0].mean() - 1) < 0.05
    assert v.max() > 0.005  # the flow has reached the middle of the tube
