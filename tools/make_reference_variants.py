"""Builds PATCHED-CONSTANT variants of the reference programs into baseline/_ref/variants/ (git-ignored,
travels to the GPU box): the reference compiles its grid size, iteration counts and save intervals in
(ldc.cu:48,615; coronary.cu:18), so the "reference kernel on the same B200" columns of SURVEY 8(d) --
480^3 (the largest cube its int indexing allows, ldc.cu:80), with and without the per-kernel
cudaDeviceSynchronize + thrust::reduce of its loop (ldc.cu:655-662) -- and runs to a true steady state
need edited copies.  The sources are read from /root/reference, edited IN MEMORY by the substitutions
listed below (every one must match, or the build stops), written only under baseline/_ref/variants/ and
compiled there with the same shim headers as oracle/_ref.  The kernels (`update`, `boundary_stream`)
are never touched.  Nothing is copied into the tracked tree.

  python tools/make_reference_variants.py            (authoring container: needs /root/reference + nvcc)

MEASUREMENT / TEST INFRASTRUCTURE."""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
OUT = ROOT / "baseline" / "_ref" / "variants"
SHIM = ROOT / "oracle" / "shim"

LDC = REF / "Lid_driven_cavity" / "ldc.cu"
COR = REF / "coronary_cfd" / "coronary.cu"


def sub(src, pattern, repl, count=1):
    new, n = re.subn(pattern, repl, src, count=count)
    if n == 0:
        raise SystemExit(f"pattern not found: {pattern}")
    return new


def ldc_dims(s, n):
    return sub(s, r"NX = 64, NY = 64, NZ = 64", f"NX = {n}, NY = {n}, NZ = {n}")


def ldc_loop(s, max_it, stop_by_tol, save_every=None, dump=True):
    s = sub(s, r"max_it=10000", f"max_it={max_it}")
    if not stop_by_tol:
        s = sub(s, r"stag_max=50", "stag_max=2000000000")
    if save_every is not None:
        s = sub(s, r"time_save=500", f"time_save={save_every}")
    if not dump:  # no periodic D2H + ASCII VTK (3 GB per dump at 480^3), no final dump
        s = sub(s, r"if\(k%time_save==0\)\{", "if(false){")
        s = sub(s, r"\toutputSave\(output_direc,k\);\n\t\n\t//free memory", "\t\n\t//free memory")
    return s


def ldc_strip(s):
    """kernel-only loop: update + boundary_stream back to back, no syncs, no residual reduction"""
    body = re.search(r"while\(k<=max_it&&tol_count<=stag_max\)\{.*?\n\t\td_tmp=d_scr;", s, re.S).group(0)
    new = body.replace("\t\tcudaDeviceSynchronize();\n", "")
    new = re.sub(r"\t\tcalc_vel_square<<<.*?\n", "", new)
    new = re.sub(r"\t\tsum_next=thrust::reduce.*?\n", "\t\tsum_next=2.f+k;\n", new)
    if new.count("cudaDeviceSynchronize") or "thrust::reduce" in new:
        raise SystemExit("strip failed")
    return s.replace(body, new)


VARIANTS = {
    # name: (source, edits, needs texture shim)
    "ldc_480": (LDC, lambda s: ldc_loop(ldc_dims(s, 480), 29, False, dump=False), False),
    "ldc_480_stripped": (LDC, lambda s: ldc_strip(ldc_loop(ldc_dims(s, 480), 29, False, dump=False)), False),
    "ldc_64_stripped": (LDC, lambda s: ldc_strip(ldc_loop(s, 9999, False, dump=False)), False),
    "ldc_64_nodump": (LDC, lambda s: ldc_loop(s, 9999, False, dump=False), False),
    # true steady states (fields written once, at the end): same kernels, residual machinery off
    "ldc_32_steady": (LDC, lambda s: ldc_strip(ldc_loop(ldc_dims(s, 32), 39999, False, save_every=2000000000)), False),
    "ldc_64_steady": (LDC, lambda s: ldc_strip(ldc_loop(s, 119999, False, save_every=2000000000)), False),
    # coronary.cu with a run length that ends: REPEAT 300000 -> 1000, one dump at the end
    "cor_1000": (COR, lambda s: sub(sub(s, r"REPEAT=300000,time_save=5000", "REPEAT=1000,time_save=1000"),
                                    r"if\(i%time_save==0\)\{", "if(i==REPEAT){"), True),
}


def main():
    if not REF.is_dir():
        print("no /root/reference here: keeping prebuilt baseline/_ref/variants (if any)")
        return
    OUT.mkdir(parents=True, exist_ok=True)
    import tempfile

    # the patched texts exist only in a temporary directory while nvcc runs: nothing but binaries is left in the
    # repository tree (no copy of reference source, edited or not)
    with tempfile.TemporaryDirectory() as tmp:
        for name, (src, edit, tex) in VARIANTS.items():
            text = edit(src.read_text(errors="replace"))
            cu = Path(tmp) / f"{name}.cu"
            cu.write_text(text)
            cmd = ["nvcc", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", f"-I{SHIM}", "-w"]
            if tex:
                cmd += ["-include", "texref_shim.h"]
            cmd += ["-o", str(OUT / name), str(cu)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode:
                sys.stderr.write(r.stderr[-3000:])
                raise SystemExit(f"nvcc failed for {name}")
            print("built", OUT / name)
    for stale in OUT.glob("*.cu"):
        stale.unlink()


if __name__ == "__main__":
    main()
