"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path -- slab ranges, compact
offsets, and the halo exchange choreography of lattice_boltzmann_method_gpu_b200.slab -- driven by a
small numpy stand-in for the slab kernels (TEST code; the product never computes on the CPU).

The stand-in follows the library's slab protocol literally: compute owned planes from a buffer with
one halo plane per interior face, "send" the 5 populations crossing each face, receive the
neighbour's into the halo planes.  Two gloo ranks must reproduce the single-domain result bitwise.

The second half does the same for the in-place (AA) storage, whose slabs exchange by peer stores on the
GPU: even steps leave the crossing populations in the neighbour's halo plane, odd steps push them into
its outermost owned plane (shifted in-plane) -- here carried by the same gloo send/recv, which is also how the
library moves them between GPUs that cannot map each other's memory (lbm_mail_stage + SlabCase.enable_staged:
staging planes sent with torch.distributed, merged on arrival; tests/test_slab_gpu.py checks that path on the GPU)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from lattice_boltzmann_method_gpu_b200 import slab  # noqa: E402
from oracle import oracle as O  # noqa: E402

UP = [5, 11, 13, 15, 16]     # c_z = +1
DOWN = [6, 12, 14, 17, 18]   # c_z = -1
W = np.array([1 / 3] + [1 / 18] * 6 + [1 / 36] * 12)


def feq(rho, u):
    cu = O.CX[:, None, None, None] * u[0] + O.CY[:, None, None, None] * u[1] + O.CZ[:, None, None, None] * u[2]
    usq = (u ** 2).sum(0)
    return W[:, None, None, None] * rho * (1 + 3 * cu + 4.5 * cu ** 2 - 1.5 * usq)


def step_planes(f, zlo, zhi, tau):
    """pull-stream + BGK on planes [zlo,zhi) of f[q,z,y,x]; periodic in x,y, explicit in z"""
    out = np.empty((19, zhi - zlo) + f.shape[2:])
    for q in range(19):
        src = f[q, zlo - O.CZ[q]: zhi - O.CZ[q]]
        out[q] = np.roll(np.roll(src, O.CY[q], axis=1), O.CX[q], axis=2)
    rho = out.sum(0)
    u = np.stack([(O.CX[:, None, None, None] * out).sum(0), (O.CY[:, None, None, None] * out).sum(0),
                  (O.CZ[:, None, None, None] * out).sum(0)]) / rho
    return out - (out - feq(rho, u)) / tau


def initial(nz, ny, nx):
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    rho = 1 + 0.01 * np.sin(2 * np.pi * z / nz) * np.cos(2 * np.pi * x / nx)
    u = np.stack([0.02 * np.sin(2 * np.pi * y / ny), 0.01 * np.cos(2 * np.pi * z / nz), 0.015 * np.sin(2 * np.pi * x / nx)])
    return feq(rho, u)


def single_domain(nz, ny, nx, steps, tau):
    """reference: z is closed by frozen planes 0 and nz-1 (never updated), like solid faces"""
    f = initial(nz, ny, nx)
    for _ in range(steps):
        new = f.copy()
        new[:, 1:nz - 1] = step_planes(f, 1, nz - 1, tau)
        f = new
    return f


def worker(rank, world, port, nz, ny, nx, steps, tau, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z0, z1 = slab.slab_ranges(nz, world)[rank]
    lo, hi = rank > 0, rank < world - 1
    zs0, zs1 = z0 - (1 if lo else 0), z1 + (1 if hi else 0)
    f = initial(nz, ny, nx)[:, zs0:zs1].copy()
    # compact offsets: every plane "stores" ny*nx nodes here
    mine = torch.tensor([(z1 - z0) * ny * nx])
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine)
    offs, total = slab.compact_offsets([int(t) for t in allc])
    assert total == nz * ny * nx and offs[rank] == z0 * ny * nx
    own_lo = max(z0, 1) - zs0          # frozen outer planes are not updated
    own_hi = min(z1, nz - 1) - zs0
    for _ in range(steps):
        new = f.copy()
        new[:, own_lo:own_hi] = step_planes(f, own_lo, own_hi, tau)
        s_lo = torch.from_numpy(np.ascontiguousarray(new[DOWN, 1])) if lo else None   # bottom owned plane
        s_hi = torch.from_numpy(np.ascontiguousarray(new[UP, -2])) if hi else None    # top owned plane
        r_lo = torch.empty_like(s_lo) if lo else None
        r_hi = torch.empty_like(s_hi) if hi else None
        for w in slab.exchange_halos(s_lo, r_lo, s_hi, r_hi, rank, world):
            w.wait()
        if lo:
            new[UP, 0] = r_lo.numpy()     # the lower neighbour's upward-moving populations
        if hi:
            new[DOWN, -1] = r_hi.numpy()
        f = new
    ret[rank] = (z0, z1, f[:, (z0 - zs0):(z1 - zs0)].copy())
    dist.destroy_process_group()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_two_gloo_ranks_equal_single_domain(world):
    nz, ny, nx, steps, tau = 13, 6, 8, 7, 0.6
    ref = single_domain(nz, ny, nx, steps, tau)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(worker, args=(world, free_port(), nz, ny, nx, steps, tau, ret), nprocs=world, join=True)
    got = np.concatenate([ret[r][2] for r in range(world)], axis=1)
    # only the populations a neighbour needs are exchanged, so compare what is defined everywhere:
    # all owned planes, all directions
    assert np.array_equal(got, ref)


# ---------------------------------------------------------------------------------------------
# In-place (AA) storage across slabs: the choreography of the fused peer-store exchange
# (csrc/step_dense.cuh, push_to_peers), with gloo send/recv standing in for the peer stores.
#   even step: local; a face-plane node also leaves g_q (q crossing the face) in slot opp(q) of the
#              neighbour's HALO plane, same (y,x)                       -> what its odd step pulls
#   odd step : pulls a[opp q][x - c_q], pushes g_q into a[q][x + c_q]; a target beyond the face is a
#              cell of the neighbour's outermost OWNED plane, shifted in-plane by (c_y, c_x), and the
#              local store is suppressed                                 -> what its even step reads
# Walls (planes 0 and nz-1): half-way bounce-back through the wall cell's slot, as in the kernel.
OPP = O.OPP


def collide(fin, tau):
    rho = fin.sum(0)
    u = np.stack([(O.CX[:, None, None, None] * fin).sum(0), (O.CY[:, None, None, None] * fin).sum(0),
                  (O.CZ[:, None, None, None] * fin).sum(0)]) / rho
    return fin - (fin - feq(rho, u)) / tau


def shift(plane, dy, dx):
    """value at (y, x) moves to (y + dy, x + dx), periodic"""
    return np.roll(np.roll(plane, dy, axis=-2), dx, axis=-1)


def gather(f, zlo, zhi):
    """pull: out[q](z,y,x) = f[q](z - cz, y - cy, x - cx) for planes [zlo, zhi)"""
    out = np.empty((19, zhi - zlo) + f.shape[2:])
    for q in range(19):
        out[q] = shift(f[q, zlo - O.CZ[q]: zhi - O.CZ[q]], O.CY[q], O.CX[q])
    return out


def walls_single_domain(nz, ny, nx, steps, tau):
    """two-buffer reference with bounce-back walls at z = 0 and nz-1 (wall cells hold the link slots)"""
    f = initial(nz, ny, nx)
    for _ in range(steps):
        g = collide(gather(f, 1, nz - 1), tau)
        new = f.copy()
        new[:, 1:nz - 1] = g
        for q in range(19):
            if O.CZ[q] == 1:    # source of link q of plane 1 is the wall plane 0: slot (q, x - c_q) <- g_opp(q)(x)
                new[q, 0] = shift(g[OPP[q], 0], -O.CY[q], -O.CX[q])
            if O.CZ[q] == -1:
                new[q, nz - 1] = shift(g[OPP[q], -1], -O.CY[q], -O.CX[q])
        f = new
    return f


def aa_worker(rank, world, port, nz, ny, nx, steps, tau, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z0, z1 = slab.slab_ranges(nz, world)[rank]
    lo, hi = rank > 0, rank < world - 1
    zs0, zs1 = z0 - (1 if lo else 0), z1 + (1 if hi else 0)
    f0 = initial(nz, ny, nx)
    fl0, fl1 = max(z0, 1), min(z1, nz - 1)          # fluid planes this rank updates (global)
    L0, L1 = fl0 - zs0, fl1 - zs0                   # ... in local plane numbers
    a = f0[:, zs0:zs1].copy()
    a[:, L0:L1] = gather(f0, fl0, fl1)              # pre-streamed initial state of the fluid planes
    wall_lo, wall_hi = z0 == 0, z1 == nz            # this rank holds the wall plane 0 / nz-1

    def exchange(s_lo, s_hi):
        t = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x))
        s_lo, s_hi = t(s_lo), t(s_hi)
        r_lo = torch.empty_like(s_lo) if lo else None
        r_hi = torch.empty_like(s_hi) if hi else None
        for w in slab.exchange_halos(s_lo, r_lo, s_hi, r_hi, rank, world):
            w.wait()
        return (r_lo.numpy() if lo else None), (r_hi.numpy() if hi else None)

    for it in range(steps):
        if it % 2 == 0:  # ---- even: purely local
            g = collide(a[:, L0:L1].copy(), tau)
            for q in range(19):
                a[OPP[q], L0:L1] = g[q]
            if wall_lo:
                for q in np.nonzero(O.CZ == 1)[0]:    # boundary slot of link q: a[opp q][x - c_q] (the wall cell)
                    a[OPP[q], 0] = shift(g[OPP[q], 0], -O.CY[q], -O.CX[q])
            if wall_hi:
                for q in np.nonzero(O.CZ == -1)[0]:
                    a[OPP[q], -1] = shift(g[OPP[q], -1], -O.CY[q], -O.CX[q])
            # peer stores: g_q of the face plane -> slot opp(q) of the neighbour's halo plane, same (y,x)
            r_lo, r_hi = exchange(g[DOWN, 0] if lo else None, g[UP, -1] if hi else None)
            if lo:
                a[OPP[UP], 0] = r_lo      # from the lower neighbour's top plane (its q in UP)
            if hi:
                a[OPP[DOWN], -1] = r_hi
        else:            # ---- odd: pull a[opp q][x - c_q], push to a[q][x + c_q]
            fin = np.empty((19, L1 - L0, ny, nx))
            for q in range(19):
                fin[q] = shift(a[OPP[q], L0 - O.CZ[q]: L1 - O.CZ[q]], O.CY[q], O.CX[q])
            g = collide(fin, tau)
            new = a.copy()
            send_lo, send_hi = [], []
            for q in range(19):
                cz = O.CZ[q]
                moved = shift(g[q], O.CY[q], O.CX[q])                # lands at (y + cy, x + cx)
                t0, t1 = L0 + cz, L1 + cz                             # target planes (local)
                src = slice(0, L1 - L0)
                if cz == 1:
                    if hi:                                            # top plane's target is the neighbour's
                        send_hi.append(moved[-1])
                        src, t1 = slice(0, L1 - L0 - 1), t1 - 1       # local store suppressed
                    elif wall_hi:                                     # target is the wall: link opp(q) bounces
                        new[OPP[q], L1 - 1] = g[q, -1]                # boundary value of link opp(q) -> a[opp q][x]
                        src, t1 = slice(0, L1 - L0 - 1), t1 - 1
                if cz == -1:
                    if lo:
                        send_lo.append(moved[0])
                        src, t0 = slice(1, L1 - L0), t0 + 1
                    elif wall_lo:
                        new[OPP[q], L0] = g[q, 0]
                        src, t0 = slice(1, L1 - L0), t0 + 1
                new[q, t0:t1] = moved[src]
            a = new
            r_lo, r_hi = exchange(np.stack(send_lo) if lo else None, np.stack(send_hi) if hi else None)
            if lo:
                a[UP, L0] = r_lo          # g_q of the lower neighbour's top plane, already shifted in-plane
            if hi:
                a[DOWN, L1 - 1] = r_hi
    ret[rank] = (fl0, fl1, a[:, L0:L1].copy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world,steps", [(2, 6), (2, 7), (3, 8), (3, 5)])
def test_in_place_storage_across_slabs(world, steps):
    nz, ny, nx, tau = 14, 6, 8, 0.6
    f = walls_single_domain(nz, ny, nx, steps, tau)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(aa_worker, args=(world, free_port(), nz, ny, nx, steps, tau, ret), nprocs=world, join=True)
    got = np.concatenate([ret[r][2] for r in range(world)], axis=1)
    if steps % 2 == 0:   # after an odd step the buffer is "pre-streamed": a[q](x) = f[q](x - c_q)
        expect = gather(f, 1, nz - 1)
    else:                # after an even step: a[opp q](x) = f[q](x)
        expect = f[OPP, 1:nz - 1]
    assert np.array_equal(got, expect)


def test_slab_ranges_and_offsets():
    assert slab.slab_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert slab.slab_ranges(8, 8) == [(i, i + 1) for i in range(8)]
    with pytest.raises(ValueError):
        slab.slab_ranges(3, 4)
    assert slab.compact_offsets([5, 0, 7]) == ([0, 5, 5], 12)
