"""The reference programs themselves (oracle/_ref, compiled unmodified) and the drop-in drivers
(drivers/, on liblbm_b200.so) run back to back on the same GPU, each on the reference's own three
shipped configurations, each timed by its own "TOTAL RUNNING TIME" line (the reference's cudaEvent
span around its main loop, VTK dumps and per-step residual included; ours: the same loop).

  python tools/compare_reference_runs.py > profiles/r01_reference_vs_ours_64.txt

MEASUREMENT / TEST INFRASTRUCTURE (like tests/): it may run the compiled reference in oracle/_ref or use the
test helpers; nothing here is part of, or imported by, the product package.
"""
import re
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path


ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
from capture_reference import write_bc_fixture  # noqa: E402

REF = ROOT / "oracle" / "_ref"
FLUID = {"ldc": 60 ** 3, "pos": 175200, "bif": 45307}


def run(exe, wd, inputs=(), args=()):
    wd.mkdir(parents=True)
    (wd / "out").mkdir()
    for s, d in inputs:
        shutil.copy(s, wd / d)
    r = subprocess.run([str(exe), *args], cwd=wd, capture_output=True, text=True, timeout=900)
    ms = float(re.search(r"TOTAL RUNNING TIME: ([0-9.eE+-]+) MILLI", r.stdout).group(1))
    its = max(int(m) for m in re.findall(r"_(\d+)\.vtk", " ".join(p.name for p in (wd / "out").glob("*.vtk"))))
    return ms, its


def main():
    rows = []
    with tempfile.TemporaryDirectory() as t:
        t = Path(t)
        fx = t / "bc_fixture.txt"
        write_bc_fixture(fx, (1, 2, 0))
        bif_in = [(REF / "geo.txt", "geo.txt"), (fx, "bc.txt")]
        cases = [("ldc", REF / "ldc_ref", ROOT / "drivers" / "ldc", []),
                 ("pos", REF / "pos_ref", ROOT / "drivers" / "poiseuille", []),
                 ("bif", REF / "bif_ref", ROOT / "drivers" / "bifurcation", bif_in)]
        for name, ref_exe, our_exe, inputs in cases:
            ms_r, it_r = run(ref_exe, t / f"{name}_ref", inputs)
            ms_o, it_o = run(our_exe, t / f"{name}_ours", inputs)
            ms_o64, _ = run(our_exe, t / f"{name}_ours64", inputs, ["--f64"])
            steps_r = it_r + (1 if name == "bif" else 0)
            steps_o = it_o + (1 if name == "bif" else 0)
            rows.append((name, it_r, ms_r, FLUID[name] * steps_r / ms_r / 1e3, it_o, ms_o, FLUID[name] * steps_o / ms_o / 1e3,
                         ms_o64, ms_r / ms_o))
    print("case  | reference program (fp32)            | drop-in driver (fp32)               | driver --f64 | speed-up")
    print("      | last iter   total ms   MLUPS        | last iter   total ms   MLUPS        | total ms     | (fp32)")
    for r in rows:
        print(f"{r[0]:5s} | {r[1]:9d} {r[2]:10.1f} {r[3]:8.1f}        | {r[4]:9d} {r[5]:10.1f} {r[6]:8.1f}        | {r[7]:10.1f}   | {r[8]:6.1f}x")
    print("\nMLUPS = fluid nodes x iterations / the program's own TOTAL RUNNING TIME (whole main loop: residual")
    print("reduction every step for ldc/pos, VTK dumps every 500 iterations, D2H copies).  Grids: 64^3, 64^3, 64x83x32.")


if __name__ == "__main__":
    main()
