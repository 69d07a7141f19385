"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path -- slab ranges, compact
offsets, and the halo exchange choreography of lattice_boltzmann_method_gpu_b200.slab -- driven by a
small numpy stand-in for the slab kernels (TEST code; the product never computes on the CPU).

The stand-in follows the library's slab protocol literally: compute owned planes from a buffer with
one halo plane per interior face, "send" the 5 populations crossing each face, receive the
neighbour's into the halo planes.  Two gloo ranks must reproduce the single-domain result bitwise."""
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


def test_slab_ranges_and_offsets():
    assert slab.slab_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert slab.slab_ranges(8, 8) == [(i, i + 1) for i in range(8)]
    with pytest.raises(ValueError):
        slab.slab_ranges(3, 4)
    assert slab.compact_offsets([5, 0, 7]) == ([0, 5, 5], 12)
