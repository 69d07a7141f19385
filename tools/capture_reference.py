"""Run the reference programs themselves (oracle/_ref/*_ref, compiled unmodified from
/root/reference by oracle/Makefile) on a GPU box and capture what they produce, so that the
oracle and the CUDA path can be pinned to the REAL reference.

  python tools/capture_reference.py gpurun_out/ref_capture

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
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_capture")
