"""Turns the captures of the PATCHED-CONSTANT reference variants (tools/make_reference_variants.py builds
them from /root/reference, tools/capture_reference.py --variants ran them on a B200 and reduced their
output; brought back in gpurun_out/ref_variants/) into the committed fixture
tests/golden/reference_variant_outputs.npz:

  ldc32 / ldc64 : final velocity block of ldc.cu run to a TRUE steady state (40 000 / 120 000 iterations,
                  kernels untouched, residual machinery off): whole block for 32^3, three mid planes +
                  checksums for 64^3
  cor           : coronary.cu (REPEAT 300000 -> 1000, one dump at the end) on the generated vessel
                  tests/helpers.coronary_like_flag(): velocity and density on 12 planes, checksums, header
  timing        : the reference's own cudaEvent spans at 64^3 / 480^3, with and without its per-kernel
                  syncs + thrust::reduce -- also written to profiles/ by the caller

  python tests/golden/make_variant_golden.py [gpurun_out/ref_variants]"""
import json
import sys
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent


def main(src):
    src = Path(src)
    out = {}
    for name, key in (("ldc_32_steady", "ldc32"), ("ldc_64_steady", "ldc64")):
        d = np.load(src / f"{name}.npz")
        nx, ny, nz = (int(v) for v in d["dims"])
        v = d["velocity"].reshape(nz, ny, nx, 3)
        out[f"{key}_dims"], out[f"{key}_last_iter"] = d["dims"], d["last_iter"]
        if key == "ldc32":
            out[f"{key}_velocity"] = v
        out[f"{key}_plane_z"], out[f"{key}_plane_y"], out[f"{key}_plane_x"] = v[nz // 2].copy(), v[:, ny // 2].copy(), v[:, :, nx // 2].copy()
        v64 = v.astype(np.float64)
        out[f"{key}_sum_abs"] = np.array(np.sqrt((v64 ** 2).sum(-1)).sum())
        out[f"{key}_sum_comp"] = v64.sum(axis=(0, 1, 2))
        out[f"{key}_max_abs"] = np.array(np.abs(v64).max())
    d = np.load(src / "cor_1000.npz")
    for k in d.files:
        if k not in ("stdout", "log"):
            out[f"cor_{k}"] = d[k]
    out["timing_json"] = np.array((src / "timing.json").read_text())
    np.savez_compressed(OUT / "reference_variant_outputs.npz", **out)
    print("wrote", OUT / "reference_variant_outputs.npz", (OUT / "reference_variant_outputs.npz").stat().st_size, "bytes")
    print(json.dumps(json.loads((src / "timing.json").read_text()), indent=1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_variants")
