"""Turns the captures of the REAL reference programs (oracle/_ref/*_ref run on a B200 by
tools/capture_reference.py, brought back in gpurun_out/ref_capture/) into the small committed
fixture tests/golden/reference_gpu_outputs.npz:

  per case (ldc, pos, bif, bif_shipped): VTK dims, the iteration number of the last VTK the
  program wrote, the residuals it logged, three orthogonal mid-planes of its final velocity field
  (float32, physical units as written = u * C_U) and full-field checksums (sum |v| and the three
  component sums in float64).

  python tests/golden/make_reference_golden.py [gpurun_out/ref_capture]
"""
import sys
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent


def main(src):
    src = Path(src)
    out = {}
    for name in ("ldc", "pos", "bif", "bif_shipped"):
        d = np.load(src / f"{name}.npz", allow_pickle=False)
        nx, ny, nz = (int(v) for v in d["dims"])
        v = d["velocity"].reshape(nz, ny, nx, 3)
        out[f"{name}_dims"] = d["dims"]
        out[f"{name}_last_iter"] = d["last_iter"]
        out[f"{name}_n_vtk"] = d["n_vtk"]
        res = [float(l) for l in d["log"] if l and not l.startswith("TOTAL")]
        out[f"{name}_residuals"] = np.array(res, dtype=np.float64)
        out[f"{name}_plane_z"] = v[nz // 2].copy()
        out[f"{name}_plane_y"] = v[:, ny // 2].copy()
        out[f"{name}_plane_x"] = v[:, :, nx // 2].copy()
        v64 = v.astype(np.float64)
        out[f"{name}_sum_abs"] = np.array(np.sqrt((v64 ** 2).sum(-1)).sum())
        out[f"{name}_sum_comp"] = v64.sum(axis=(0, 1, 2))
        out[f"{name}_max_abs"] = np.array(np.abs(v64).max())
        out[f"{name}_header"] = d["header"]
        print(name, (nx, ny, nz), "last_iter", int(d["last_iter"]), "residuals", len(res), "sum|v|", float(out[f"{name}_sum_abs"]))
    np.savez_compressed(OUT / "reference_gpu_outputs.npz", **out)
    print("wrote", OUT / "reference_gpu_outputs.npz", (OUT / "reference_gpu_outputs.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_capture")
