"""Pin of the voxeliser against the reference's own data: bifurcation/bif.stl voxelised on the
64 x 83 x 32 grid of bifurcation/geo.txt at the reference's own spacing CH (bifurcation.cu:20), grid
origin fitted for the best overlap (the MATLAB step that made geo.txt is not shipped).  Writes tests/golden/bif_voxel_fit.json.

  python tests/golden/make_voxel_fit.py            # CPU oracle (this container: needs /root/reference)
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

REF = Path("/root/reference/bifurcation")


def main():
    tri = O.read_stl(REF / "bif.stl")
    geo = np.array((REF / "geo.txt").read_text().split(), dtype=np.int32).reshape(32, 83, 64).astype(np.uint8)
    lo = tri.reshape(-1, 3).min(0).astype(np.float64)

    def score(h, off):
        m = O.voxelize(tri, lo + np.asarray(off) * h, h, (64, 83, 32))
        return int((m & geo).sum()) / int((m | geo).sum()), m

    H_REF = 0.248925  # the reference's lattice spacing in mm: CH = 0.000248925f m (bifurcation.cu:20)
    best = (0.0, None)
    for dx in np.arange(-2.5, -1.49, 0.25):  # coarse: where the grid sits relative to the surface's bounding box
        for dz in np.arange(-2.0, -0.99, 0.25):
            for dy in np.arange(-2.5, -0.49, 0.5):
                s, _ = score(H_REF, (dx, dy, dz))
                if s > best[0]:
                    best = (s, (H_REF, dx, dy, dz))
    _, dx0, dy0, dz0 = best[1]
    for dx in dx0 + np.arange(-0.1875, 0.19, 0.0625):  # fine
        for dz in dz0 + np.arange(-0.1875, 0.19, 0.0625):
            for dy in dy0 + np.arange(-0.25, 0.26, 0.125):
                s, _ = score(H_REF, (dx, dy, dz))
                if s > best[0]:
                    best = (s, (H_REF, dx, dy, dz))
    h, dx, dy, dz = best[1]
    s, m = score(h, (dx, dy, dz))
    diff = m != geo
    g = geo.astype(bool)
    pad = np.pad(g, 1)
    dil = np.zeros_like(g)
    ero = np.ones_like(g)
    for sz, sy, sx in ((0, 0, 1), (0, 0, -1), (0, 1, 0), (0, -1, 0), (1, 0, 0), (-1, 0, 0)):
        nb = pad[1 + sz:1 + sz + g.shape[0], 1 + sy:1 + sy + g.shape[1], 1 + sx:1 + sx + g.shape[2]]
        dil |= nb
        ero &= nb
    shell = (dil | g) & ~(ero & g)
    out = {"spacing": h, "origin": [float(v) for v in lo + np.array([dx, dy, dz]) * h],
           "offset_in_voxels_from_stl_bbox_min": [float(dx), float(dy), float(dz)], "iou": s,
           "inside_voxels": int(m.sum()), "geo_inside_voxels": int(geo.sum()), "mismatches": int(diff.sum()),
           "mismatches_in_surface_shell_of_geo": int((diff & shell).sum()),
           "packed_checksum": int(np.packbits(m).astype(np.uint64).sum()), "triangles": int(len(tri))}
    (ROOT / "tests" / "golden" / "bif_voxel_fit.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
