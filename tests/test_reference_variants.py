"""Pins against PATCHED-CONSTANT builds of the reference programs (tools/make_reference_variants.py: grid size,
iteration counts and save intervals are compile-time constants in the reference; the kernels are untouched),
run on a B200 by tools/capture_reference.py --variants and reduced to tests/golden/reference_variant_outputs.npz
(tests/golden/make_variant_golden.py):

  * coronary.cu (coronary_cfd/, REPEAT 300000 -> 1000) on a generated 291 x 291 x 372 vessel that satisfies
    the five openings the program hard-codes (cor:77-141; helpers.coronary_like_flag): the GEO_OPENINGS rule,
    its three boundary kinds and the density / pressure / velocity writer are pinned to the REAL program;
  * ldc.cu run to a true steady state (32^3: 40 000 iterations, 64^3: 120 000): there the in-place wall
    bounce of ldc.cu (see tests/test_ldc_order_cpu.py) has nothing left to act on, and what remains between
    the real binary and the library is fp32 rounding of two differently contracted evaluations."""
from pathlib import Path

import numpy as np
import pytest

import helpers as H
from oracle import oracle as O

G = np.load(H.GOLDEN / "reference_variant_outputs.npz")
C_U = np.float32(2.74909090909091)
C_RHO = np.float32(1060.0)


def cor_blocks(geo, idx, rho, ux, uy, uz):
    """the arrays coronary.cu's outputSave prints (cor:948-1011): trimmed box z[1,NZ-2] y[2,NY-3] x[1,NX-2]"""
    def full(a):
        f = np.zeros(geo.shape, np.float32)
        m = idx >= 0
        f[m] = a.astype(np.float32)[idx[m]]
        return f

    V = np.stack([full(ux) * C_U, full(uy) * C_U, full(uz) * C_U], -1)[1:-1, 2:-2, 1:-1]
    D = (full(rho) * C_RHO)[1:-1, 2:-2, 1:-1]
    return V, D


def compare_cor(V, D, tol):
    scale = float(G["cor_max_abs"])
    assert list(V.shape[:3][::-1]) == list(G["cor_dims"])
    n = 0
    for k in G.files:
        if not k.startswith("cor_vel_"):
            continue
        ax, c = k[8], int(k[9:])
        sl = {"z": (c - 1, slice(None), slice(None)), "y": (slice(None), c - 2, slice(None)),
              "x": (slice(None), slice(None), c - 1)}[ax]
        assert float(np.abs(V[sl] - G[k]).max()) / scale < tol, k
        assert float(np.abs(D[sl] - G["cor_rho_" + k[8:]]).max()) / float(C_RHO) < tol, k
        n += 1
    assert n == 12
    s = float(np.sqrt((V.astype(np.float64) ** 2).sum(-1)).sum())
    assert abs(s - float(G["cor_sum_abs"])) / float(G["cor_sum_abs"]) < tol
    assert int((D != 0).sum()) == int(G["cor_rho_nonzero"])
    assert abs(float(D.astype(np.float64).sum()) - float(G["cor_rho_sum"])) / float(G["cor_rho_sum"]) < tol


def test_oracle_reproduces_the_coronary_program():
    """CPU: labels 2,3,5,6,7 all occur, NLATTICE is the number the real program printed, and 1001 steps of the
    oracle land on its fields (2e-5: the program prints 6 digits and is compiled with FMA contraction)"""
    flag = H.coronary_like_flag()
    geo = O.geo_pre_cor(flag, O.cor_reference_rules(291, 291, 372))
    idx, nlat = O.index_transform(geo)
    assert nlat == 340783  # "#LATTICE340783" in the real program's log
    assert set(np.unique(geo)) == {-1, 0, 1, 2, 3, 4, 5, 6, 7}
    o = O.Oracle(O.CASE_COR, geo, idx, nlat, H.TAU_LDC, 0.0, dtype=np.float32)
    o.set_cor_speeds(H.COR_SPEEDS["uin"], H.COR_SPEEDS["uout"], H.COR_SPEEDS["usub"])
    o.initialize()
    o.step(int(G["cor_last_iter"]) + 1)  # loop index 0..1000 inclusive (cor:1100)
    V, D = cor_blocks(geo, idx, *o.fields())
    compare_cor(V, D, 2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("storage_name,from_file", [("sparse_aa", True), ("dense_ab", False), ("sparse_ab", False), ("dense_aa", False)])
def test_gpu_reproduces_the_coronary_program(storage_name, from_file, tmp_path):
    """the library with the reference's own constants (lbm_case_defaults(GEO_OPENINGS) = coronary.cu's), driven
    like its main(): geo.txt in the program's y-fastest order, 1001 steps, one dump"""
    import lattice_boltzmann_method_gpu_b200 as L

    storage = {"dense_ab": L.STORE_DENSE_AB, "dense_aa": L.STORE_DENSE_AA, "sparse_ab": L.STORE_SPARSE_AB,
               "sparse_aa": L.STORE_SPARSE_AA}[storage_name]
    flag = H.coronary_like_flag()
    d = L.case_defaults(L.CASE_GEO_OPENINGS)
    d.precision, d.math, d.storage = L.F32, L.MATH_FAST, storage
    d.out_dir = str(tmp_path).encode()
    if from_file:
        H.write_geo_txt(tmp_path / "geo.txt", flag, yfast=True)
        d.geo_path = str(tmp_path / "geo.txt").encode()
    c = L.Case(d)
    if not from_file:
        c.set_flag(flag)
    c.geo_pre()
    assert c.index_transform() == 340783
    c.initialize()
    c.run_fixed(1000, 1000, from_file)
    geo, idx = c.get_geo(), c.get_index()
    V, D = cor_blocks(geo, idx, *c.get_fields())
    compare_cor(V, D, 2e-5)
    if from_file:  # the writer: same header lines, three sections, values parse back to the same blocks
        lines = (tmp_path / "coronary_1000.vtk").open().read(600).split("\n")[:8]
        assert lines == [str(s) for s in G["cor_header"]]
        sections = {}
        with open(tmp_path / "coronary_1000.vtk") as f:  # three sections of one (long) line each, cor:960-1008
            while True:
                line = f.readline()
                if not line:
                    break
                if line.startswith("SCALARS"):
                    f.readline()
                    sections[line.split()[1]] = np.fromstring(f.readline(), dtype=np.float32, sep=" ")
                elif line.startswith("VECTORS"):
                    sections[line.split()[1]] = np.fromstring(f.readline(), dtype=np.float32, sep=" ")
        assert sorted(sections) == ["DENSITY", "PRESSURE", "VELOCITY"]
        assert float(np.abs(sections["VELOCITY"].reshape(V.shape) - V).max()) <= 1e-5 * float(G["cor_max_abs"])
        assert float(np.abs(sections["DENSITY"].reshape(D.shape) - D).max()) <= 1e-5 * float(C_RHO)
        C_pre = C_RHO * C_U * C_U
        assert np.allclose(sections["PRESSURE"].reshape(D.shape), D / C_RHO * C_pre / 3.0, rtol=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("n,key", [(32, "ldc32"), (64, "ldc64")])
@pytest.mark.parametrize("storage_name", ["dense_ab", "dense_aa", "sparse_aa"])
def test_gpu_matches_the_cavity_program_at_steady_state(n, key, storage_name):
    """ldc.cu after 40 000 / 120 000 iterations.  Tolerance 5e-5 of max|v|, not 1e-5: two correctly rounded fp32
    evaluations that differ only in FMA contraction settle on steady states this far apart -- the CPU oracle
    (gcc, no contraction, the reference's own expression order) is 2.1e-5 from the real binary's field at 32^3."""
    import lattice_boltzmann_method_gpu_b200 as L

    storage = {"dense_ab": L.STORE_DENSE_AB, "dense_aa": L.STORE_DENSE_AA, "sparse_aa": L.STORE_SPARSE_AA}[storage_name]
    c = H.gpu_case("ldc", n, L.F32, L.MATH_FAST, storage=storage)
    H.gpu_setup(c, "ldc")
    c.step(int(G[f"{key}_last_iter"]))
    from test_reference_outputs import vtk_velocity

    geo, idx = c.get_geo(), c.get_index()
    _, ux, uy, uz = c.get_fields()
    V = vtk_velocity("ldc", geo.shape, idx, ux, uy, uz)
    nz, ny, nx = V.shape[:3]
    scale = float(G[f"{key}_max_abs"])
    for nm, got in (("plane_z", V[nz // 2]), ("plane_y", V[:, ny // 2]), ("plane_x", V[:, :, nx // 2])):
        assert float(np.abs(got - G[f"{key}_{nm}"]).max()) / scale < 5e-5, nm
    if key == "ldc32":
        assert float(np.abs(V - G["ldc32_velocity"]).max()) / scale < 5e-5
    s = float(np.sqrt((V.astype(np.float64) ** 2).sum(-1)).sum())
    assert abs(s - float(G[f"{key}_sum_abs"])) / float(G[f"{key}_sum_abs"]) < 5e-5
