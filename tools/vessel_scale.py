"""BASELINE.json config 5: a synthetic vessel bundle with a pulsatile inlet, z-slab sharded over the
GPUs of one box (one rank per GPU, NCCL halo exchange), sparse or dense storage.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/vessel_scale.py [--size 1024] [--k 4] [--storage sparse|dense|aa] [--halo p2p|nccl] [--precision f64] [--steps 50] [--verify]

Each rank builds only the planes of the voxel mask it needs (lbm_set_flag_slab).  --verify recomputes
the run as a single domain on rank 0 (small n only) and checks the gathered fields bit for bit.
The pulsatile inlet has no reference implementation (curved vessel/README.md:1): parity unpinned.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import lattice_boltzmann_method_gpu_b200 as L  # noqa: E402
from lattice_boltzmann_method_gpu_b200 import slab  # noqa: E402
from sparse_bench import tube_bundle  # noqa: E402


def inlet_plane(n, k, rfrac):
    pitch = n / k
    zz, xx = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    rr2 = ((xx % pitch - pitch / 2) ** 2 + (zz % pitch - pitch / 2) ** 2) / (rfrac * pitch) ** 2
    return (0.05 * np.clip(1 - rr2, 0, None)).astype(np.float32)


def desc_for(a, z_range, device):
    d = L.case_defaults(L.CASE_GEO_Y_INOUT)
    d.nx = d.ny = d.nz = a.n
    d.z_begin, d.z_end = z_range
    d.precision = L.F64 if a.precision == "f64" else L.F32
    d.storage = {"sparse": L.STORE_SPARSE_AB, "dense": L.STORE_DENSE_AB, "aa": L.STORE_DENSE_AA, "sparse_aa": L.STORE_SPARSE_AA}[a.storage]
    d.pulse_amp, d.pulse_period = 0.3, 200.0
    d.bc[0].pulsatile = 1
    d.device = device
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=1024)
    ap.add_argument("--k", type=int, default=4)
    ap.add_argument("--radius-frac", type=float, default=0.38)
    ap.add_argument("--storage", default="sparse")
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--verify", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    z_range = slab.slab_ranges(a.n, world)[rank]
    c = slab.SlabCase(desc_for(a, z_range, local))
    z0, z1 = c.needed_flag_planes()
    inlet = inlet_plane(a.n, a.k, a.radius_frac)
    c.setup(flag_slab=(tube_bundle(a.n, a.k, a.radius_frac, z0, z1).astype(np.uint8), z0),
            bc_planes=(inlet, np.zeros_like(inlet)))
    halo = "nccl"
    if a.halo == "p2p" and world > 1:
        halo = "p2p" if c.enable_p2p() else "nccl (peer mapping unavailable)"
    c.step(5)
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([c.step_timed(a.steps)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    nf = torch.tensor([c.num_fluid, c.device_bytes], dtype=torch.int64, device="cuda")
    per_rank = [torch.zeros_like(nf) for _ in range(world)]
    dist.all_gather(per_rank, nf)
    fluid = sum(int(t[0]) for t in per_rank)
    out = {"config": f"vessel bundle {a.n}^3, {a.k}x{a.k} bent tubes, pulsatile inlet, {a.storage} storage, {a.precision}",
           "n_gpus": world, "halo_exchange": halo, "fluid_nodes": fluid, "fill": fluid / a.n ** 3, "nlattice": c.nlattice,
           "mlups": fluid * a.steps / (float(ms) * 1e-3) / 1e6, "ms_per_step": float(ms) / a.steps,
           "device_GB_per_rank": [round(int(t[1]) / 1e9, 2) for t in per_rank],
           "fluid_per_rank": [int(t[0]) for t in per_rank]}
    if a.verify:
        total_steps = 5 + a.steps
        mine = [torch.from_numpy(x).cuda() for x in c.get_fields()]
        counts = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(counts, torch.tensor([mine[0].numel()], dtype=torch.int64, device="cuda"))
        counts = [int(t) for t in counts]
        same = True
        gathered = []
        for k in range(4):
            parts = []
            for r in range(world):
                buf = mine[k] if r == rank else torch.zeros(counts[r], dtype=mine[k].dtype, device="cuda")
                dist.broadcast(buf, r)
                parts.append(buf)
            gathered.append(torch.cat(parts).cpu().numpy())
        if rank == 0:
            one = L.Case(desc_for(a, (0, a.n), local))
            one.set_flag_slab(tube_bundle(a.n, a.k, a.radius_frac).astype(np.uint8), 0)
            one.geo_pre(), one.index_transform(), one.set_bc_planes(inlet, np.zeros_like(inlet)), one.initialize()
            one.step(5), one.step(a.steps)
            ref = one.get_fields()
            same = all(np.array_equal(g, r) for g, r in zip(gathered, ref))
            out["bitwise_equal_to_single_domain"] = bool(same)
            out["steps_verified"] = total_steps
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
