"""GPU: z-slab decomposition.  The reference is single-GPU, so the oracle of the multi-GPU path is
"P slabs == one domain, bit for bit".  Here the P slabs live on ONE device ("virtual slabs"): the
same halo pack / unpack / face-first code path as the multi-GPU run, with the transfer done by a
device-to-device copy instead of NCCL."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def run_slabs(name, n, P, steps, precision, math_mode, pulse=None, storage=None, staged=False):
    import torch

    import lattice_boltzmann_method_gpu_b200 as L
    from lattice_boltzmann_method_gpu_b200 import slab

    nz = {"ldc": n, "pos": n, "bif": 32, "cor": 44}[name]
    ranges = slab.slab_ranges(nz, P)
    cs = [H.gpu_case(name, n, precision, math_mode, pulse=pulse, z_range=r, storage=storage) for r in ranges]
    for c in cs:
        c.geo_pre()
    offs, total = slab.compact_offsets([c.local_stored_count() for c in cs])
    for c, o in zip(cs, offs):
        c.set_compact_offset(o, total)
        assert c.index_transform() == total
        if name == "bif":
            c.set_bc_planes(*H.bif_bc_planes())
        c.initialize()
    if staged:  # dense in-place storage, mailbox exchange through local staging buffers (lbm_mail_stage)
        for r, c in enumerate(cs):
            if r > 0:
                c.mail_stage(0)
            if r < P - 1:
                c.mail_stage(1)

    def view(c, side):
        s, r, ns, nr = c.halo_buffers(side)
        if not s:
            return None, None
        it = c.dtype.itemsize
        mk = lambda p, nb: torch.as_tensor(slab._DevBuf(p, max(nb, it), c.dtype), device="cuda")[: nb // it]
        return mk(s, ns), mk(r, nr)

    bufs = [(view(c, 0), view(c, 1)) for c in cs]
    for it in range(steps):
        flags = L.STEP_MOMENTS if it == steps - 1 else 0
        for c in cs:
            c.step_begin(flags)
        if staged:  # the send / receive parts alternate with the step's parity
            bufs = [(view(c, 0), view(c, 1)) for c in cs]
        for c in cs:
            c.step_interior()
            c.sync()
        for r in range(P - 1):
            (_, _), (s_hi, r_hi) = bufs[r]
            (s_lo, r_lo), _ = bufs[r + 1]
            r_lo.copy_(s_hi)  # upward-moving populations of slab r's top plane
            r_hi.copy_(s_lo)  # downward-moving populations of slab r+1's bottom plane
        torch.cuda.synchronize()
        for c in cs:
            c.step_end()
    for c in cs:
        c.sync()
    return cs, total


@pytest.mark.parametrize("name,n,P", [("ldc", 24, 3), ("ldc", 20, 5), ("pos", 24, 2), ("bif", None, 4), ("cor", None, 3)])
@pytest.mark.parametrize("math_name", ["strict", "fast"])
def test_slabs_equal_single_domain_bitwise(name, n, P, math_name):
    import lattice_boltzmann_method_gpu_b200 as L

    mm = L.MATH_STRICT if math_name == "strict" else L.MATH_FAST
    steps = 23
    one = H.gpu_case(name, n, L.F64, mm)
    nlat = H.gpu_setup(one, name)
    one.step(steps)
    ref = one.get_fields()
    cs, total = run_slabs(name, n, P, steps, L.F64, mm)
    assert total == nlat
    assert np.array_equal(np.concatenate([c.get_geo() for c in cs]), one.get_geo())
    assert np.array_equal(np.concatenate([c.get_index() for c in cs]), one.get_index())
    assert sum(c.num_fluid for c in cs) == one.num_fluid
    parts = [c.get_fields() for c in cs]
    firsts = [c.compact_first for c in cs]
    assert firsts == sorted(firsts) and firsts[0] == 0
    for k in range(4):
        got = np.concatenate([p[k] for p in parts])
        assert np.array_equal(got, ref[k]), f"field {k}: max diff {np.abs(got - ref[k]).max()}"


def test_one_plane_slabs():
    """degenerate decomposition: every interior slab owns a single plane (top face == bottom face)"""
    import lattice_boltzmann_method_gpu_b200 as L

    n, steps = 12, 9
    one = H.gpu_case("ldc", n, L.F32, L.MATH_FAST)
    H.gpu_setup(one, "ldc")
    one.step(steps)
    ref = one.get_fields()
    cs, _ = run_slabs("ldc", n, n, steps, L.F32, L.MATH_FAST)
    for k in range(4):
        assert np.array_equal(np.concatenate([c.get_fields()[k] for c in cs]), ref[k])


def test_pulsatile_slabs():
    import lattice_boltzmann_method_gpu_b200 as L

    pulse = (0.25, 30.0)
    one = H.gpu_case("bif", None, L.F64, L.MATH_FAST, pulse=pulse)
    H.gpu_setup(one, "bif")
    one.step(40)
    ref = one.get_fields()
    cs, _ = run_slabs("bif", None, 2, 40, L.F64, L.MATH_FAST, pulse=pulse)
    for k in range(4):
        assert np.array_equal(np.concatenate([c.get_fields()[k] for c in cs]), ref[k])


def test_per_slab_masks():
    """each slab is handed only the planes of the voxel field it needs (lbm_set_flag_slab)"""
    import lattice_boltzmann_method_gpu_b200 as L
    from lattice_boltzmann_method_gpu_b200 import slab

    flag = H.bif_flag()
    one = H.gpu_case("bif", None, L.F64, L.MATH_FAST, storage=L.STORE_SPARSE_AB)
    nlat = H.gpu_setup(one, "bif")
    one.step(15)
    ref = one.get_fields()
    cs = []
    for r in slab.slab_ranges(32, 3):
        c = H.gpu_case("bif", None, L.F64, L.MATH_FAST, z_range=r, storage=L.STORE_SPARSE_AB)
        z0, z1 = c.needed_flag_planes()
        c.set_flag_slab(flag[z0:z1].astype(np.uint8), z0)
        with pytest.raises(L.LbmError):
            c.set_flag_slab(flag[z0 + 1:z1].astype(np.uint8), z0 + 1)  # does not cover the needed planes
        c.geo_pre()
        cs.append(c)
    offs, total = slab.compact_offsets([c.local_stored_count() for c in cs])
    assert total == nlat
    assert np.array_equal(np.concatenate([c.get_geo() for c in cs]), one.get_geo())


@pytest.mark.parametrize("name,n,P,storage_name", [("ldc", 24, 3, "dense"), ("bif", None, 4, "dense"), ("bif", None, 4, "sparse"),
                                                   ("cor", None, 3, "sparse"), ("ldc", 12, 12, "dense"), ("ldc", 24, 3, "aa"),
                                                   ("bif", None, 4, "aa"), ("cor", None, 5, "aa"), ("pos", 24, 2, "aa"),
                                                   ("ldc", 12, 12, "aa")])
def test_fused_peer_store_halo_exchange(name, n, P, storage_name):
    """lbm_p2p_attach: the step kernel stores the crossing populations straight into the neighbouring
    handle's halo plane (here: another handle on the same GPU); no pack / unpack / copy at all"""
    import lattice_boltzmann_method_gpu_b200 as L
    from lattice_boltzmann_method_gpu_b200 import slab

    storage = {"sparse": L.STORE_SPARSE_AB, "dense": L.STORE_DENSE_AB, "aa": L.STORE_DENSE_AA}[storage_name]
    steps = 21
    one = H.gpu_case(name, n, L.F64, L.MATH_FAST, storage=storage)
    H.gpu_setup(one, name)
    one.step(steps)
    ref = one.get_fields()
    nz = {"ldc": n, "pos": n, "bif": 32, "cor": 44}[name]
    cs = [H.gpu_case(name, n, L.F64, L.MATH_FAST, z_range=r, storage=storage) for r in slab.slab_ranges(nz, P)]
    for c in cs:
        c.geo_pre()
    offs, total = slab.compact_offsets([c.local_stored_count() for c in cs])
    for c, o in zip(cs, offs):
        c.set_compact_offset(o, total)
        c.index_transform()
        if name == "bif":
            c.set_bc_planes(*H.bif_bc_planes())
        c.initialize()
    exp = [c.p2p_export() for c in cs]
    for r, c in enumerate(cs):
        for side, nb in ((0, r - 1), (1, r + 1)):
            if 0 <= nb < P:
                e = exp[nb]
                c.p2p_attach(side, e["ptrs"][0], e["ptrs"][1], e["qs"], e["halo_c0"][1 - side], e["face_c0"][1 - side])
    launches0 = sum(c.launch_count for c in cs)
    for it in range(steps):
        for c in cs:
            c.step_begin(L.STEP_MOMENTS if it == steps - 1 else 0)
            c.step_interior()
            c.step_end()
        for c in cs:
            c.sync()  # lock-step
    for k in range(4):
        assert np.array_equal(np.concatenate([c.get_fields()[k] for c in cs]), ref[k]), f"field {k}"
    # only step kernels ran: at most face(s) + interior per slab and step
    assert sum(c.launch_count for c in cs) - launches0 <= 3 * P * steps + 4 * P


@pytest.mark.parametrize("name,n,P", [("ldc", 24, 3), ("ldc", 20, 5), ("pos", 24, 2), ("bif", None, 4), ("cor", None, 3)])
@pytest.mark.parametrize("math_name,prec", [("strict", "f64"), ("fast", "f32")])
def test_staged_mailboxes_equal_single_domain_bitwise(name, n, P, math_name, prec):
    """the in-place dense storage over a transport WITHOUT peer mapping (lbm_mail_stage: the face launches store
    into local staging buffers, a copy -- NCCL send/recv between processes, a device copy here -- moves them, and
    lbm_step_end merges exactly the elements the neighbour wrote into the mailbox): P slabs == one domain, bit
    for bit, both parities, every case rule"""
    import lattice_boltzmann_method_gpu_b200 as L

    mm = L.MATH_STRICT if math_name == "strict" else L.MATH_FAST
    pr = L.F64 if prec == "f64" else L.F32
    steps = 23
    one = H.gpu_case(name, n, pr, mm, storage=L.STORE_DENSE_AA)
    H.gpu_setup(one, name)
    one.step(steps)
    ref = one.get_fields()
    cs, _ = run_slabs(name, n, P, steps, pr, mm, storage=L.STORE_DENSE_AA, staged=True)
    for k in range(4):
        got = np.concatenate([c.get_fields()[k] for c in cs])
        assert np.array_equal(got, ref[k]), f"field {k}: max diff {np.abs(got - ref[k]).max()}"
    # the population buffers are whole again when they are asked for (mailboxes drained), and stepping goes on
    for c in cs:
        c.get_populations()
    one.step(4)
    ref = one.get_fields()
    import torch

    from lattice_boltzmann_method_gpu_b200 import slab

    def view(c, side):
        s_, r_, ns, nr = c.halo_buffers(side)
        it = c.dtype.itemsize
        mk = lambda p, nb: torch.as_tensor(slab._DevBuf(p, max(nb, it), c.dtype), device="cuda")[: nb // it]
        return (mk(s_, ns), mk(r_, nr)) if s_ else (None, None)

    for it in range(4):
        for c in cs:
            c.step_begin(L.STEP_MOMENTS)
        bufs = [(view(c, 0), view(c, 1)) for c in cs]
        for c in cs:
            c.step_interior()
            c.sync()
        for r in range(P - 1):
            (_, _), (s_hi, r_hi) = bufs[r]
            (s_lo, r_lo), _ = bufs[r + 1]
            r_lo.copy_(s_hi)
            r_hi.copy_(s_lo)
        torch.cuda.synchronize()
        for c in cs:
            c.step_end()
    for k in range(4):
        got = np.concatenate([c.get_fields()[k] for c in cs])
        assert np.array_equal(got, ref[k])
