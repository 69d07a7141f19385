"""CPU: the geo.txt reader of the library (csrc/geo_text.h: the file mapped, cut at whitespace and parsed by
several threads) returns exactly what the reference's reader returns -- `fscanf(f, "%d ", &tmp)` token by token
(bifurcation.cu:50-61, coronary.cu:45-56) -- on well-formed, short, surplus and malformed files, with 1, 3 and
16 threads.  Checked by a small C++ program (tests/cpp/geo_parse_check.cpp)."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_mapped_parser_equals_fscanf(tmp_path):
    exe = tmp_path / "geo_parse_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", f"-I{ROOT / 'lattice_boltzmann_method_gpu_b200' / 'csrc'}",
                    str(ROOT / "tests" / "cpp" / "geo_parse_check.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "0 mismatches" in r.stdout
