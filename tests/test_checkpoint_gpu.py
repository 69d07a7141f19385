"""GPU: checkpoint / restart -- a run restored from a dump continues bit-identically."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("storage_name", ["dense_ab", "dense_aa", "sparse_ab"])
@pytest.mark.parametrize("at", [7, 10])  # odd and even step counts (buffer parity / AA phase)
def test_restart_continues_bitwise(storage_name, at, tmp_path):
    import lattice_boltzmann_method_gpu_b200 as L

    storage = {"dense_ab": L.STORE_DENSE_AB, "dense_aa": L.STORE_DENSE_AA, "sparse_ab": L.STORE_SPARSE_AB}[storage_name]
    a = H.gpu_case("bif", None, L.F64, L.MATH_FAST, pulse=(0.3, 40.0), storage=storage)
    H.gpu_setup(a, "bif")
    a.step(at)
    a.checkpoint_save(tmp_path / "ck.bin")
    a.step(13)
    ref = a.get_fields()
    b = H.gpu_case("bif", None, L.F64, L.MATH_FAST, pulse=(0.3, 40.0), storage=storage)
    H.gpu_setup(b, "bif")
    b.step(3)  # any state: the load overwrites it
    b.checkpoint_load(tmp_path / "ck.bin")
    assert b.step_count == at
    b.step(13)
    for x, y in zip(ref, b.get_fields()):
        assert np.array_equal(x, y)


def test_checkpoint_of_another_case_is_rejected(tmp_path):
    import lattice_boltzmann_method_gpu_b200 as L

    a = H.gpu_case("ldc", 16, L.F32, L.MATH_FAST)
    H.gpu_setup(a, "ldc")
    a.step(2)
    a.checkpoint_save(tmp_path / "ck.bin")
    b = H.gpu_case("ldc", 16, L.F64, L.MATH_FAST)
    H.gpu_setup(b, "ldc")
    with pytest.raises(L.LbmError):
        b.checkpoint_load(tmp_path / "ck.bin")
    with pytest.raises(L.LbmError):
        b.checkpoint_load(tmp_path / "missing.bin")
