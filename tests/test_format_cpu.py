"""CPU: the ASCII-VTK float formatter of the library (csrc/vtk_format.h: vtk_write, a %g routine of its own
that hands rounding ties to std::to_chars general/6; all 2^32 float patterns against to_chars:
profiles/r02_format_exhaustive.txt) writes exactly the characters the reference's `ofs << value << ' '` writes (ldc.cu:603-607,
bifurcation.cu:1140-1150) -- checked by a small C++ program on pseudo-random bit patterns of float and
double, values of the solver's magnitudes and the special values."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_to_chars_equals_ostream(tmp_path):
    exe = tmp_path / "fmt_check"
    subprocess.run(["g++", "-O2", "-std=c++17", f"-I{ROOT / 'lattice_boltzmann_method_gpu_b200' / 'csrc'}",
                    str(ROOT / "tests" / "cpp" / "fmt_check.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe), "500000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "0 mismatches" in r.stdout
