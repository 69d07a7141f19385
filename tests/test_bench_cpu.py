"""CPU: the parts of bench.py that need no GPU -- the reference arm (the CPU oracle timed on a bounded sample, the
JSON line the driver parses) and the host-side helpers that define the workloads (dims per GPU count, the vessel
bundle generator of BASELINE config 5)."""
import json
import subprocess
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3", "--ref-n", "32"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "MLUPS" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 2 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert "32^3" in line["cpu_baseline"]["sample"]  # the sample that was really run is named
    assert line["e2e"] == {"value": line["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "512x512x512" in line["config"]["workload"]  # the arm's config is the GPU arm's


def test_workload_dims_keep_the_nodes_per_gpu():
    a = SimpleNamespace(n=512, dims=None)
    assert bench.global_dims(a, 1) == (512, 512, 512)
    for w in (2, 4, 8):
        gx, gy, gz = bench.global_dims(a, w)
        assert (gx, gy) == (1024, 1024) and gx * gy * gz == w * 512 ** 3  # 1024^2 faces, 2^27 nodes per GPU
    assert bench.vessel_edge(512, 1) == 512 and bench.vessel_edge(512, 8) == 1024
    for w in (2, 4):
        e = bench.vessel_edge(512, w)
        assert e % 32 == 0 and abs(e ** 3 / w / 512 ** 3 - 1) < 0.06


def test_vessel_inputs_are_the_bundle_of_config_5():
    n = 96
    flag, inlet = bench.vessel_inputs(n, 0, n)
    assert flag.shape == (n, n, n) and flag.dtype == np.uint8 and set(np.unique(flag)) == {0, 1}
    assert 0.40 < flag.mean() < 0.47  # 4 x 4 tubes of radius 0.38 pitch: pi 0.38^2 = 45 % of the box
    part, _ = bench.vessel_inputs(n, 10, 20)
    assert np.array_equal(part, flag[10:20])  # a slab is the same field
    assert inlet.shape == (n, n) and inlet.dtype == np.float32 and 0.045 < inlet.max() <= 0.05 and inlet.min() == 0.0
    # the inlet is non-zero exactly where the tubes open onto the y = 0 face (up to the rim cells)
    assert (inlet > 0).mean() <= flag[:, 0, :].mean() + 0.02
