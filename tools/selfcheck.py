"""Memory- and race-check of the step kernels WITHOUT compute-sanitizer (the GPU pool refuses to run it).

Builds a second copy of the library with -DLBM_SELFCHECK into variants/selfcheck/ -- in that build every
population load and store of the step kernels (csrc/step_dense.cuh, step_sparse.cuh, step_sparse_aa.cuh:
the LBM_CHK lines) first
  * checks that the address lies inside one of the handle's population buffers, guard cells included
    (memcheck's job; the speculative pull deliberately reads past rows into those guards), and
  * exchanges a (launch id, thread id) tag into a shadow word of that element and counts the element when
    ANOTHER thread of the same launch was there before (the in-place storage and the fluid-side boundary
    slots rely on every element being touched by one thread per launch; racecheck only sees shared memory,
    of which these kernels use none) --
and runs tools/sanitize_cases.py against it: all storages, both phases of the in-place ones, the four case
rules' boundary paths, forced speculative pull, virtual z-slabs.  Peer stores into ANOTHER handle's buffers
are bounds-checked by that design's tests (bitwise equality with the single domain), not here.

  python tools/selfcheck.py [--build-only]      -> gpurun_out/selfcheck/*.log, exit 0 iff all counters are 0
MEASUREMENT / TEST INFRASTRUCTURE."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "lattice_boltzmann_method_gpu_b200" / "csrc"
OUT = ROOT / "variants" / "selfcheck"
LIB = OUT / "liblbm_b200_selfcheck.so"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC", "-DLBM_SELFCHECK"]


def build():
    OUT.mkdir(parents=True, exist_ok=True)
    objs = []
    for name, fmad in (("lbm_geo", False), ("lbm_voxel", False), ("lbm_step_strict", False), ("lbm_step_fast", False), ("lbm_api", True)):
        o = OUT / f"{name}.o"
        cmd = ["nvcc", *FLAGS] + ([] if fmad else ["-fmad=false"]) + ["-c", str(SRC / f"{name}.cu"), "-o", str(o)]
        newest = max(p.stat().st_mtime for p in list(SRC.glob("*.cu*")) + list(SRC.glob("*.h")))
        if not o.exists() or o.stat().st_mtime < newest:
            subprocess.run(cmd, check=True)
        objs.append(str(o))
    subprocess.run(["nvcc", *ARCH, "-shared", "-o", str(LIB), *objs, "-lcudart"], check=True)
    return LIB


def main():
    if not LIB.exists() or "--build-only" in sys.argv or Path("/root/reference").is_dir():
        try:
            build()
        except (subprocess.CalledProcessError, FileNotFoundError) as e:
            if not LIB.exists():
                raise SystemExit(f"cannot build the self-checking library: {e}")
    if "--build-only" in sys.argv:
        print("built", LIB)
        return 0
    logdir = ROOT / "gpurun_out" / "selfcheck"
    logdir.mkdir(parents=True, exist_ok=True)
    rc = 0
    for tag, env_extra, args in (("default", {}, []), ("speculative_pull_forced", {"LBM_SPECULATIVE": "1"}, ["quick"])):
        env = dict(os.environ, LBM_B200_LIB=str(LIB), **env_extra)
        r = subprocess.run([sys.executable, str(ROOT / "tools" / "sanitize_cases.py"), *args], env=env, capture_output=True, text=True)
        (logdir / f"{tag}.log").write_text(r.stdout + r.stderr)
        last = [l for l in r.stdout.split("\n") if l.startswith("sanitize_cases done")]
        print(tag, "exit", r.returncode, last[-1] if last else r.stderr[-400:])
        rc |= r.returncode
    return rc


if __name__ == "__main__":
    sys.exit(main())
