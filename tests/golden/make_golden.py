"""Generates the committed fixtures in tests/golden/ from the read-only reference tree.

Run in the authoring container (needs /root/reference):  python tests/golden/make_golden.py

  bif_flag_bits.npy   the shipped bifurcation/geo.txt (64x83x32 binary voxels, order z,y,x)
                      bit-packed with numpy.packbits  (input data, 21 KB)
  bif_bc.npy          the shipped bifurcation/bc.txt as float32[3,32,64] (its three NZ*NX planes)
  counts.json         known answers: thesis section 4.8 lattice count 65820 for that geometry, and
                      the label / link census the CPU oracle produces for it and for the
                      64^3 Poiseuille and LDC masks
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def census(geo, fluid):
    lab = {int(k): int(v) for k, v in zip(*np.unique(geo, return_counts=True))}
    nz, ny, nx = geo.shape
    links = {}
    zz, yy, xx = np.nonzero(geo == fluid)
    for q in range(1, 19):
        sx, sy, sz = xx - O.CX[q], yy - O.CY[q], zz - O.CZ[q]
        src = geo[sz, sy, sx]
        for k, v in zip(*np.unique(src, return_counts=True)):
            links[int(k)] = links.get(int(k), 0) + int(v)
    return lab, links


def main():
    nx, ny, nz = 64, 83, 32
    flag, ntok = O.read_geo_file(REF / "bifurcation/geo.txt", nx, ny, nz)
    assert ntok == nx * ny * nz
    np.save(OUT / "bif_flag_bits.npy", np.packbits(flag.astype(np.uint8).ravel()))
    bc = np.loadtxt(REF / "bifurcation/bc.txt", dtype=np.float32).reshape(3, nz, nx)
    np.save(OUT / "bif_bc.npy", bc)

    out = {"thesis_bif_nlattice": 65820}
    geo = O.geo_pre_bif(flag)
    _, nlat = O.index_transform(geo)
    lab, links = census(geo, 4)
    out["bif"] = {"nlattice": nlat, "labels": lab, "fluid_links_by_source": links}
    geo = O.geo_pre_pos(64, 64, 64)
    _, nlat = O.index_transform(geo)
    lab, links = census(geo, 4)
    out["pos64"] = {"nlattice": nlat, "labels": lab, "fluid_links_by_source": links}
    geo = O.geo_pre_ldc(64, 64, 64)
    lab, links = census(geo, 3)
    out["ldc64"] = {"nlattice": 64 ** 3, "labels": lab, "fluid_links_by_source": links}
    (OUT / "counts.json").write_text(json.dumps(out, indent=1, sort_keys=True))
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
