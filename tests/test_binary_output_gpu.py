"""Binary legacy-VTK output: the same points and values as the reference's ASCII files,
and slab pieces that tile the single-domain file."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def read_binary_vtk(path):
    """-> (header dict, {section name: float32 array})"""
    raw = path.read_bytes()
    pos, head = 0, {}
    sections = {}

    def line():
        nonlocal pos
        end = raw.index(b"\n", pos)
        s = raw[pos:end].decode()
        pos = end + 1
        return s

    assert line() == "# vtk DataFile Version 2.0"
    line()
    assert line() == "BINARY"
    assert line() == "DATASET STRUCTURED_POINTS"
    for _ in range(4):
        k, *v = line().split()
        head[k] = v
    n = int(head["POINT_DATA"][0])
    assert n == int(np.prod([int(t) for t in head["DIMENSIONS"]]))
    while pos < len(raw):
        s = line()
        if not s:
            continue
        if s.startswith("SCALARS"):
            assert line() == "LOOKUP_TABLE default"
            sections[s.split()[1]] = np.frombuffer(raw, dtype=">f4", count=n, offset=pos).astype(np.float32)
            pos += 4 * n
        elif s.startswith("VECTORS"):
            sections[s.split()[1]] = np.frombuffer(raw, dtype=">f4", count=3 * n, offset=pos).astype(np.float32).reshape(n, 3)
            pos += 12 * n
        else:
            raise AssertionError(f"unexpected line {s!r}")
    return head, sections


def read_ascii_vtk(path):
    lines = path.read_text().split("\n")
    head = {ln.split()[0]: ln.split()[1:] for ln in lines[4:8]}
    sections = {}
    i = 8
    while i < len(lines):
        s = lines[i]
        if s.startswith("SCALARS"):
            sections[s.split()[1]] = np.array(lines[i + 2].split(), dtype=np.float32)
            i += 3
        elif s.startswith("VECTORS"):
            sections[s.split()[1]] = np.array(lines[i + 1].split(), dtype=np.float32).reshape(-1, 3)
            i += 2
        else:
            i += 1
    return head, sections


@pytest.mark.parametrize("name,n", [("ldc", 20), ("pos", 16), ("bif", None), ("cor", None)])
def test_binary_file_holds_what_the_ascii_file_prints(name, n, tmp_path):
    import lattice_boltzmann_method_gpu_b200 as L

    c = H.gpu_case(name, n, L.F32, L.MATH_FAST, out_dir=str(tmp_path))
    H.gpu_setup(c, name)
    c.step(30)
    c.outputSave(30)
    c.set_output_format(L.OUT_BINARY_VTK)
    c.outputSave(30)
    prefix = c.desc.out_name.decode()
    ha, sa = read_ascii_vtk(tmp_path / f"{prefix}_30.vtk")
    hb, sb = read_binary_vtk(tmp_path / f"{prefix}_30_bin.vtk")
    assert ha == hb
    assert sa.keys() == sb.keys() and "VELOCITY" in sb
    for k in sa:
        assert sa[k].shape == sb[k].shape
        # the ASCII file prints 6 significant digits (relative error <= 5e-6) of the float the binary file stores exactly
        assert np.allclose(sa[k], sb[k], rtol=1e-5, atol=0.0), k
        assert np.abs(sb[k]).max() > 0
    with pytest.raises(L.LbmError):
        c.set_output_format(7)


def test_slab_pieces_tile_the_single_domain_file(tmp_path):
    """each slab handle writes its own planes; stacked in z they are the single-domain file"""
    import lattice_boltzmann_method_gpu_b200 as L
    from lattice_boltzmann_method_gpu_b200 import slab

    one = H.gpu_case("bif", None, L.F64, L.MATH_FAST, out_dir=str(tmp_path))
    H.gpu_setup(one, "bif")
    one.step(12)
    one.set_output_format(L.OUT_BINARY_VTK)
    one.outputSave(12)
    head, whole = read_binary_vtk(tmp_path / "bif_12_bin.vtk")
    ranges = slab.slab_ranges(32, 3)
    cs = [H.gpu_case("bif", None, L.F64, L.MATH_FAST, z_range=r, out_dir=str(tmp_path)) for r in ranges]
    for c in cs:
        c.geo_pre()
    offs, total = slab.compact_offsets([c.local_stored_count() for c in cs])
    for c, o in zip(cs, offs):
        c.set_compact_offset(o, total)
        c.index_transform()
        c.set_bc_planes(*H.bif_bc_planes())
        c.initialize()
    H.attach_virtual_slabs(cs)
    H.step_virtual_slabs(cs, 12)
    parts, z_seen = [], 0
    for c, (z0, z1) in zip(cs, ranges):
        with pytest.raises(L.LbmError):
            c.outputSave(12)  # the ASCII writer is single-domain only
        c.set_output_format(L.OUT_BINARY_VTK)
        c.outputSave(12)
        hp, sp = read_binary_vtk(tmp_path / f"bif_12_bin.z{z0}.vtk")
        assert hp["DIMENSIONS"][:2] == head["DIMENSIONS"][:2] and hp["SPACING"] == head["SPACING"]
        assert float(hp["ORIGIN"][2]) == pytest.approx(z_seen * float(head["SPACING"][2]))
        z_seen += int(hp["DIMENSIONS"][2])
        parts.append(sp["VELOCITY"])
    assert z_seen == int(head["DIMENSIONS"][2])
    assert np.array_equal(np.concatenate(parts), whole["VELOCITY"])
