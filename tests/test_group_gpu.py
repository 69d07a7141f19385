"""GPU: the multi-GPU run loops INSIDE the C ABI (lbm_create_distributed / lbm_group_*): several z-slabs
driven from one process, crossing populations by fused peer stores, neighbours ordered by events, residuals
summed over the slabs.  Oracle: "P slabs == one domain, bit for bit" (the reference is single-GPU).  On a
1-GPU box every slab lives on device 0 -- the same code path as one slab per device except for the
cudaDeviceEnablePeerAccess call -- so the driver's GPU test run exercises it; with more devices visible the
slabs are spread over them."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _devices(P):
    import torch

    n = torch.cuda.device_count()
    return [r % n for r in range(P)]


def _group(name, n, P, L, storage, math=None, prec=None, out_dir=None):
    c = H.gpu_case(name, n, L.F64 if prec is None else prec, L.MATH_FAST if math is None else math, storage=storage, out_dir=out_dir)
    d = c.desc
    c.close()
    g = L.Group(d, P, _devices(P))
    flag = H.bif_flag() if name == "bif" else (H.openings_mask(name)[0] if name in ("cor", "corstep") else None)
    g.setup(flag=flag, bc_planes=H.bif_bc_planes() if name == "bif" else None)
    return g


@pytest.mark.parametrize("name,n,P", [("ldc", 24, 3), ("ldc", 21, 5), ("pos", 24, 2), ("bif", None, 4), ("cor", None, 3)])
@pytest.mark.parametrize("storage_name", ["dense_ab", "dense_aa", "sparse_ab", "sparse_aa"])
def test_group_equals_single_domain_bitwise(name, n, P, storage_name):
    import lattice_boltzmann_method_gpu_b200 as L

    storage = {"dense_ab": L.STORE_DENSE_AB, "dense_aa": L.STORE_DENSE_AA, "sparse_ab": L.STORE_SPARSE_AB,
               "sparse_aa": L.STORE_SPARSE_AA}[storage_name]
    one = H.gpu_case(name, n, L.F64, L.MATH_FAST, storage=storage)
    nlat = H.gpu_setup(one, name)
    g = _group(name, n, P, L, storage)
    assert g.nlattice == nlat and g.size == P and g.num_fluid == one.num_fluid
    assert np.array_equal(g.get_index(), one.get_index())
    for steps in (1, 2, 20):  # odd and even totals
        one.step(steps)
        g.step(steps)
        for a, b in zip(one.get_fields(), g.get_fields()):
            assert np.array_equal(a, b), (name, storage_name, steps)
    assert abs(g.residual(L.RES_U2SUM) - one.calc_res()) <= 1e-12 * one.calc_res()
    g.close()


@pytest.mark.parametrize("name,rule", [("ldc", 0), ("pos", 1)])
def test_group_convergence_loop_matches_single_domain(name, rule, tmp_path):
    """ldc.cu:653-685 on slabs: same stopping iteration (S is a float sum; its order differs across
    slabs, so +-2), same files, and -- when the iteration agrees -- byte-identical VTK"""
    import lattice_boltzmann_method_gpu_b200 as L

    res = {}
    for kind in ("one", "group"):
        d = L.case_defaults(rule)
        d.nx = d.ny = d.nz = 32
        d.z_begin, d.z_end = 0, 32
        d.precision, d.math, d.storage = L.F32, L.MATH_FAST, L.STORE_DENSE_AA
        out = tmp_path / kind
        out.mkdir()
        d.out_dir = str(out).encode()
        if kind == "one":
            c = L.Case(d)
            c.geo_pre(), c.index_transform(), c.initialize()
            res[kind] = c.run_converge(3000, 1e-5, 50, 500, True)
        else:
            g = L.Group(d, 3, _devices(3))
            g.setup()
            res[kind] = g.run_converge(3000, 1e-5, 50, 500, True)
    (i1, r1), (i2, r2) = res["one"], res["group"]
    assert abs(i1 - i2) <= 2 and i1 > 100, (i1, i2)
    prefix = "lid" if name == "ldc" else "pos"
    for t in [t for t in (0, 500) if t < min(i1, i2)]:
        assert (tmp_path / "one" / f"{prefix}_{t}.vtk").read_bytes() == (tmp_path / "group" / f"{prefix}_{t}.vtk").read_bytes()
    if i1 == i2:
        assert (tmp_path / "one" / f"{prefix}_{i1}.vtk").read_bytes() == (tmp_path / "group" / f"{prefix}_{i2}.vtk").read_bytes()
    l1 = (tmp_path / "one" / "CONVERGENCE.log").read_text().split("\n")
    l2 = (tmp_path / "group" / "CONVERGENCE.log").read_text().split("\n")
    assert np.allclose([float(v) for v in l1[:2]], [float(v) for v in l2[:2]], rtol=1e-3)


def test_group_fixed_loop_writes_the_single_domain_files(tmp_path):
    """bifurcation.cu:1246-1274 on 4 slabs: bif_0.vtk / bif_300.vtk byte-identical to the single-domain run"""
    import lattice_boltzmann_method_gpu_b200 as L

    (tmp_path / "one").mkdir(), (tmp_path / "group").mkdir()
    c = H.gpu_case("bif", None, L.F32, L.MATH_FAST, out_dir=tmp_path / "one")
    H.gpu_setup(c, "bif")
    c.run_fixed(300, 300, True)
    g = _group("bif", None, 4, L, L.STORE_DENSE_AB, prec=L.F32, out_dir=tmp_path / "group")
    g.run_fixed(300, 300, True)
    for t in (0, 300):
        assert (tmp_path / "one" / f"bif_{t}.vtk").read_bytes() == (tmp_path / "group" / f"bif_{t}.vtk").read_bytes()
    a = [float(v) for v in (tmp_path / "one" / "CONVERGENCE.log").read_text().split("\n")[:2]]
    b = [float(v) for v in (tmp_path / "group" / "CONVERGENCE.log").read_text().split("\n")[:2]]
    assert np.allclose(a, b, rtol=1e-4)
    # binary pieces: one per slab
    g.set_output_format(L.OUT_BINARY_VTK)
    g.outputSave(301)
    assert len(list((tmp_path / "group").glob("bif_301_bin.z*.vtk"))) == 4


def test_group_argument_checks():
    import lattice_boltzmann_method_gpu_b200 as L

    d = L.case_defaults(L.CASE_LDC)
    d.nx = d.ny = d.nz = 16
    d.z_begin, d.z_end = 0, 16
    with pytest.raises(L.LbmError):
        L.Group(d, 17)          # more slabs than planes
    g = L.Group(d, 2, [0, 0])
    with pytest.raises(L.LbmError):
        g.step(1)               # before setup
    g.setup()
    g.step(3)
    g.close()


def test_mgpu_c_driver(tmp_path):
    """drivers/ldc_mgpu.c: the reference's main() on slabs, from C, no Python / NCCL"""
    exe, one = ROOT / "drivers" / "ldc_mgpu", ROOT / "drivers" / "ldc"
    if not exe.exists() or not one.exists():
        pytest.skip("drivers not built")
    for wd in ("a", "b"):
        (tmp_path / wd / "out").mkdir(parents=True)
    args = ["--n", "32", "--steps", "600", "--save", "300"]
    r1 = subprocess.run([str(one)] + args, cwd=tmp_path / "a", capture_output=True, text=True, timeout=300)
    dev = ",".join(str(v) for v in _devices(3))
    r2 = subprocess.run([str(exe), "--slabs", "3", "--devices", dev] + args, cwd=tmp_path / "b", capture_output=True, text=True, timeout=300)
    assert r1.returncode == 0 and r2.returncode == 0, r1.stderr + r2.stderr
    assert "#LATTICE32768" in r2.stdout
    # the single-GPU driver uses two-buffer storage, this one streams in place: FAST fp32 agrees bitwise (tests/test_aa_gpu.py)
    for t in (0, 300, 600):
        assert (tmp_path / "a" / "out" / f"lid_{t}.vtk").read_bytes() == (tmp_path / "b" / "out" / f"lid_{t}.vtk").read_bytes()


def test_bc_csv_lists_every_opening_node(tmp_path):
    """write_once of coronary.cu:1033-1051"""
    import lattice_boltzmann_method_gpu_b200 as L

    c = H.gpu_case("cor", None, L.F32, L.MATH_FAST)
    H.gpu_setup(c, "cor")
    c.step(5)
    c.write_bc_csv(tmp_path / "vel.csv")
    rows = [l.split(",") for l in (tmp_path / "vel.csv").read_text().split("\n") if l]
    geo = c.get_geo()
    assert len(rows) == int(np.isin(geo, (2, 3, 5, 6, 7)).sum()) and len(rows[0]) == 6
    x, y, z = (int(v) for v in rows[0][:3])
    assert geo[z, y, x] in (2, 3, 5, 6, 7)
