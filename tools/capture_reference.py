"""Run the reference programs themselves (oracle/_ref/*_ref, compiled unmodified from
/root/reference by oracle/Makefile) on a GPU box and capture what they produce, so that the
oracle and the CUDA path can be pinned to the REAL reference.

  python tools/capture_reference.py gpurun_out/ref_capture
  python tools/capture_reference.py --variants gpurun_out/ref_variants     (baseline/_ref/variants, see
                                                                            tools/make_reference_variants.py)

Per case it stores <case>.npz with the final velocity field parsed from the last VTK the
program wrote (float32, VTK order), the iteration number in that file's name, the lines of
out/CONVERGENCE.log and the program's stdout.  tests/golden/make_reference_golden.py turns
these captures into the small committed fixtures.
MEASUREMENT / TEST INFRASTRUCTURE (like tests/): it may run the compiled reference in oracle/_ref or use the
test helpers; nothing here is part of, or imported by, the product package.
"""
import re
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "oracle" / "_ref"
GOLDEN = ROOT / "tests" / "golden"


def parse_vtk(path):
    txt = Path(path).read_text().split("\n")
    dims = [int(v) for v in txt[4].split()[1:]]
    body = {}
    i = 8
    while i < len(txt):
        line = txt[i]
        if line.startswith("SCALARS"):
            body[line.split()[1]] = np.array(txt[i + 2].split(), dtype=np.float32)
            i += 3
        elif line.startswith("VECTORS"):
            body[line.split()[1]] = np.array(txt[i + 1].split(), dtype=np.float32).reshape(-1, 3)
            i += 2
        else:
            i += 1
    return dims, body


def run_case(name, exe, workdir, prefix, inputs=()):
    wd = Path(workdir) / name
    (wd / "out").mkdir(parents=True)
    for src, dst in inputs:
        shutil.copy(src, wd / dst)
    t0 = time.time()
    r = subprocess.run([str(exe)], cwd=wd, capture_output=True, text=True, timeout=900)
    dt = time.time() - t0
    vtks = sorted((wd / "out").glob(f"{prefix}_*.vtk"), key=lambda p: int(re.findall(r"_(\d+)\.vtk", p.name)[0]))
    if not vtks:
        raise RuntimeError(f"{name}: no VTK written; stdout={r.stdout[-500:]} stderr={r.stderr[-500:]}")
    last = vtks[-1]
    k = int(re.findall(r"_(\d+)\.vtk", last.name)[0])
    dims, body = parse_vtk(last)
    header = last.read_text().split("\n")[:9]
    log = (wd / "out" / "CONVERGENCE.log").read_text().split("\n")
    return dict(dims=np.array(dims), velocity=body["VELOCITY"], last_iter=k, n_vtk=len(vtks), log=np.array(log),
                stdout=np.array(r.stdout.split("\n")), header=np.array(header), wall_s=dt, returncode=r.returncode)


def write_bc_fixture(path, order):
    bc = np.load(GOLDEN / "bif_bc.npy")
    with open(path, "w") as f:
        for p in order:
            f.write("".join("%.6f " % v for v in bc[p].ravel()))


VAR = ROOT / "baseline" / "_ref" / "variants"
# planes kept from coronary.cu's 289 x 287 x 370 output block (global coordinates): the axis plane of the main
# tube, the planes through the three branches, the planes next to every opening
COR_PLANES = {"z": (100, 184, 190, 203), "y": (150, 128, 170), "x": (4, 100, 182, 226, 271)}


def parse_vtk_big(path):
    """coronary.cu's VTK (cor:948-1011): DENSITY, PRESSURE, VELOCITY sections of one line each, ~150 M tokens"""
    out, dims = {}, None
    with open(path) as f:
        while True:
            line = f.readline()
            if not line:
                break
            if line.startswith("DIMENSIONS"):
                dims = [int(v) for v in line.split()[1:]]
            elif line.startswith("SCALARS"):
                name = line.split()[1]
                f.readline()  # LOOKUP_TABLE
                out[name] = np.fromstring(f.readline(), dtype=np.float32, sep=" ")
            elif line.startswith("VECTORS"):
                out[line.split()[1]] = np.fromstring(f.readline(), dtype=np.float32, sep=" ").reshape(-1, 3)
    return dims, out


def run_variant(name, wd, inputs=()):
    wd = Path(wd) / name
    (wd / "out").mkdir(parents=True)
    for src, dst in inputs:
        shutil.copy(src, wd / dst)
    t0 = time.time()
    r = subprocess.run([str(VAR / name)], cwd=wd, capture_output=True, text=True, timeout=1500)
    m = re.search(r"TOTAL RUNNING TIME: ([0-9.eE+-]+) MILLI", r.stdout)
    return wd, (float(m.group(1)) if m else float("nan")), time.time() - t0, r


def variants(outdir):
    """the patched-constant variants built by tools/make_reference_variants.py"""
    sys.path.insert(0, str(ROOT)), sys.path.insert(0, str(ROOT / "tests"))
    import helpers as H

    out = Path(outdir)
    out.mkdir(parents=True, exist_ok=True)
    rows = []
    with tempfile.TemporaryDirectory() as wd:
        # --- timing columns of SURVEY 8(d): the reference's own event span / iterations
        for name, n, its in (("ldc_64_nodump", 64, 10000), ("ldc_64_stripped", 64, 10000), ("ldc_480", 480, 30),
                             ("ldc_480_stripped", 480, 30)):
            if not (VAR / name).exists():
                continue
            _, ms, wall, r = run_variant(name, wd)
            fluid = (n - 4) ** 3
            rows.append(dict(variant=name, grid=f"{n}^3 fp32", iterations=its, total_ms=ms, ms_per_step=ms / its,
                             mlups_fluid=fluid * its / ms / 1e3, mlups_all_nodes=n ** 3 * its / ms / 1e3, wall_s=wall,
                             gbs_at_152B=fluid * its / ms / 1e3 * 152 / 1e3))
            print(rows[-1], flush=True)
        (out / "timing.json").write_text(__import__("json").dumps(rows, indent=1))
        # --- true steady states of the cavity
        for name, n in (("ldc_32_steady", 32), ("ldc_64_steady", 64)):
            if not (VAR / name).exists():
                continue
            w, ms, wall, r = run_variant(name, wd)
            last = max((w / "out").glob("lid_*.vtk"), key=lambda p: int(re.findall(r"_(\d+)\.vtk", p.name)[0]))
            k = int(re.findall(r"_(\d+)\.vtk", last.name)[0])
            dims, body = parse_vtk(last)
            np.savez_compressed(out / f"{name}.npz", dims=np.array(dims), velocity=body["VELOCITY"], last_iter=k, total_ms=ms)
            print(name, "dims", dims, "last_iter", k, "max|u|", float(np.abs(body["VELOCITY"]).max()), "wall_s", round(wall, 1), flush=True)
        # --- coronary.cu on a generated vessel that satisfies its hard-coded openings
        if (VAR / "cor_1000").exists():
            geo = Path(wd) / "geo_cor.txt"
            H.write_geo_txt(geo, H.coronary_like_flag(), yfast=True)
            w, ms, wall, r = run_variant("cor_1000", wd, [(geo, "geo.txt")])
            vtk = w / "out" / "coronary_1000.vtk"
            dims, body = parse_vtk_big(vtk)
            nx, ny, nz = dims
            res = dict(dims=np.array(dims), last_iter=1000, total_ms=ms, header=np.array(vtk.open().read(600).split("\n")[:8]),
                       log=np.array((w / "out" / "CONVERGENCE.log").read_text().split("\n")),
                       stdout=np.array(r.stdout.split("\n")[-6:]))
            V = body["VELOCITY"].reshape(nz, ny, nx, 3)
            D = body["DENSITY"].reshape(nz, ny, nx)
            P = body["PRESSURE"].reshape(nz, ny, nx)
            for ax, coords in COR_PLANES.items():
                for c in coords:
                    sl = {"z": (c - 1, slice(None), slice(None)), "y": (slice(None), c - 2, slice(None)),
                          "x": (slice(None), slice(None), c - 1)}[ax]
                    res[f"vel_{ax}{c}"] = V[sl].copy()
                    res[f"rho_{ax}{c}"] = D[sl].copy()
            v64 = V.astype(np.float64)
            res["sum_abs"] = np.array(np.sqrt((v64 ** 2).sum(-1)).sum())
            res["sum_comp"] = v64.sum(axis=(0, 1, 2))
            res["max_abs"] = np.array(np.abs(v64).max())
            res["rho_sum"] = np.array(D.astype(np.float64).sum())
            res["rho_nonzero"] = np.array(int((D != 0).sum()))
            res["pre_sum"] = np.array(P.astype(np.float64).sum())
            np.savez_compressed(out / "cor_1000.npz", **res)
            print("cor_1000 dims", dims, "max|u|", float(res["max_abs"]), "rho nonzero", int(res["rho_nonzero"]), "total_ms", ms,
                  "wall_s", round(wall, 1), flush=True)
            print("  stdout tail:", [l for l in r.stdout.split("\n") if l][-3:])


def main(outdir):
    out = Path(outdir)
    out.mkdir(parents=True, exist_ok=True)
    with tempfile.TemporaryDirectory() as wd:
        res = {}
        res["ldc"] = run_case("ldc", REF / "ldc_ref", wd, "lid")
        res["pos"] = run_case("pos", REF / "pos_ref", wd, "pos")
        # bifurcation as shipped (inlet plane of bc.txt is all zero -> rest state) ...
        res["bif_shipped"] = run_case("bif_shipped", REF / "bif_ref", wd, "bif",
                                      [(REF / "geo.txt", "geo.txt"), (REF / "bc.txt", "bc.txt")])
        # ... and with the shipped profile plane moved first (the fixture the parity tests use)
        fx = Path(wd) / "bc_fixture.txt"
        write_bc_fixture(fx, (1, 2, 0))
        res["bif"] = run_case("bif", REF / "bif_ref", wd, "bif", [(REF / "geo.txt", "geo.txt"), (fx, "bc.txt")])
        for name, r in res.items():
            np.savez_compressed(out / f"{name}.npz", **r)
            v = r["velocity"]
            print(name, "dims", r["dims"], "last_iter", r["last_iter"], "n_vtk", r["n_vtk"], "max|u|", float(np.abs(v).max()),
                  "wall_s", round(r["wall_s"], 2), "rc", r["returncode"])
            print("  log tail:", [l for l in r["log"] if l][-2:])
            print("  stdout tail:", [l for l in r["stdout"] if l][-3:])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--variants":
        variants(sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/ref_variants")
    else:
        main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_capture")
