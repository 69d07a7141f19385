"""CPU: why the lid-driven cavity of the real ldc.cu binary differs from "walls first, then fluid".

ldc.cu's `update` bounces its wall nodes in place on d_scr in the same launch in which fluid nodes pull
from them (ldc.cu:75-313).  The oracle can replay the order that launch executes in (oracle/lbm_oracle.c,
orc_set_ldc_order): a thread walks its z-column top-down (koff = 7..0, ldc.cu:66), so e.g. the fluid layer
z = 2 always reads the z = 1 wall BEFORE it is bounced -- a value two iterations old.  Only links whose wall
node is bounced in the same koff iteration by another warp are a true race; mode 1 / mode 2 take them
fresh / stale.

Committed fixture (tests/golden/make_ldc_order_golden.py): the three mid planes of the VTK block after
the 5119 iterations the real program ran on a B200 (reference_gpu_outputs.npz), for modes 0, 1, 2.
Claims checked here, all relative to max|v| of the reference's field:
  * the defined semantics (mode 0, what the product implements bit for bit) is 7.6e-4 away: not rounding;
  * both order models are >= 15x closer (3.4e-5 / 4.7e-5);
  * the reference lies INSIDE the envelope of the two models: |ref - m1| and |ref - m2| are both
    <= |m1 - m2| + 1e-5, i.e. what is left is the race the code really has, and it is < 1e-4;
  * the fixture still comes from the committed oracle (300-step recomputation of mode 1, bit-exact)."""
import numpy as np

import helpers as H
from test_reference_outputs import GOLD, vtk_velocity

FIX = np.load(H.GOLDEN / "ldc_order_model.npz")
PLANES = ("plane_z", "plane_y", "plane_x")


def _err(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) / float(GOLD["ldc_max_abs"])


def test_reference_run_lies_inside_the_order_envelope():
    assert int(FIX["iterations"]) == int(GOLD["ldc_last_iter"])
    e0 = max(_err(GOLD[f"ldc_{p}"], FIX[f"mode0_{p}"]) for p in PLANES)
    e1 = max(_err(GOLD[f"ldc_{p}"], FIX[f"mode1_{p}"]) for p in PLANES)
    e2 = max(_err(GOLD[f"ldc_{p}"], FIX[f"mode2_{p}"]) for p in PLANES)
    width = max(_err(FIX[f"mode1_{p}"], FIX[f"mode2_{p}"]) for p in PLANES)
    assert 5e-4 < e0 < 1e-3, e0          # the gap the GPU test has to allow for (tests/test_reference_outputs.py)
    assert e1 < 5e-5 and e2 < 6e-5, (e1, e2)
    assert e0 > 15 * max(e1, e2)
    assert width < 1e-4, width
    for p in PLANES:                      # pointwise: inside the envelope, plane by plane
        ref, m1, m2 = (a.astype(np.float64) for a in (GOLD[f"ldc_{p}"], FIX[f"mode1_{p}"], FIX[f"mode2_{p}"]))
        w = np.abs(m1 - m2).max() + 1e-5 * float(GOLD["ldc_max_abs"])
        assert np.abs(ref - m1).max() <= w and np.abs(ref - m2).max() <= w


def test_fixture_comes_from_the_committed_oracle():
    o, geo, idx, _ = H.oracle_case("ldc", 64, np.float32)
    o.set_ldc_order(1)
    o.step(int(FIX["guard_steps"]))
    _, ux, uy, uz = o.fields()
    V = vtk_velocity("ldc", geo.shape, idx, ux, uy, uz)
    nz, ny, nx = V.shape[:3]
    for p, got in zip(PLANES, (V[nz // 2], V[:, ny // 2], V[:, :, nx // 2])):
        assert np.array_equal(got, FIX[f"guard_{p}"])


def test_order_modes_agree_where_nothing_moves():
    """with a lid at rest every mode keeps the rest state exactly; and mode 0 == the default"""
    from oracle import oracle as O

    geo = O.geo_pre_ldc(16, 16, 16)
    idx, nlat = O.index_dense(geo.shape)
    outs = []
    for mode in (None, 0, 1, 2):
        o = O.Oracle(O.CASE_LDC, geo, idx, nlat, H.TAU_LDC, H.LDC_UMAX, dtype=np.float64)
        if mode is not None:
            o.set_ldc_order(mode)
        o.initialize()
        o.step(40)
        outs.append(o.fields())
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
    # the models differ from the defined semantics only through the timing of wall values
    assert not np.array_equal(outs[1][3], outs[2][3])
    assert H.rel_err(outs[2], outs[1]) < 5e-2
