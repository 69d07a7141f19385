"""Multi-GPU parity: every rank runs one z-slab (SlabCase, NCCL halo exchange); the gathered fields
must equal the single-domain run computed on rank 0, bit for bit.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/mgpu_check.py

MEASUREMENT / TEST INFRASTRUCTURE (like tests/): it may run the compiled reference in oracle/_ref or use the
test helpers; nothing here is part of, or imported by, the product package.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import helpers as H  # noqa: E402
import lattice_boltzmann_method_gpu_b200 as L  # noqa: E402
from lattice_boltzmann_method_gpu_b200 import slab  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name, n, steps in (("ldc", 40, 60), ("bif", None, 80), ("cor", None, 50), ("pos", 32, 40)):
        nz = {"ldc": n, "pos": n, "bif": 32, "cor": 44}[name]
        z0, z1 = slab.slab_ranges(nz, world)[rank]
        storage = L.STORE_SPARSE_AB if os.environ.get("LBM_SPARSE") == "1" else L.STORE_DENSE_AB
        if os.environ.get("LBM_AA") == "1":  # in-place storage: peer stores are its only transport
            storage = L.STORE_DENSE_AA
        if os.environ.get("LBM_SPARSE_AA") == "1":
            storage = L.STORE_SPARSE_AA
        base = H.gpu_case(name, n, L.F64, L.MATH_FAST, z_range=(z0, z1), storage=storage)
        d = base.desc
        d.device = local
        base.close()
        c = slab.SlabCase(d)
        flag = H.bif_flag() if name == "bif" else (H.synthetic_openings_mask()[0] if name == "cor" else None)
        c.setup(flag=flag, bc_planes=H.bif_bc_planes() if name == "bif" else None)
        if os.environ.get("LBM_STAGED") == "1" and storage == L.STORE_DENSE_AA:
            c.enable_staged()  # in-place storage over NCCL send/recv: no peer mapping (lbm_mail_stage)
        elif os.environ.get("LBM_P2P") == "1" or storage in (L.STORE_DENSE_AA, L.STORE_SPARSE_AA):
            assert c.enable_p2p(), f"peer mapping unavailable: {c.p2p_error}"
        c.step(steps)
        if c._p2p and name == "ldc":  # the multi-process convergence loop: S all-reduced per batch (ldc.cu:653-685)
            its, res = c.run_converge(max_it=steps + 39, tol=0.0, stag_max=10 ** 9, time_save=10 ** 9)
            assert its == steps + 40, its
            steps += its
        mine = [torch.from_numpy(a).cuda() for a in c.get_fields()]
        counts = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(counts, torch.tensor([mine[0].numel()], dtype=torch.int64, device="cuda"))
        counts = [int(t.item()) for t in counts]
        gathered = []
        for k in range(4):
            bufs = [torch.zeros(cn, dtype=mine[k].dtype, device="cuda") for cn in counts]
            dist.all_gather(bufs, mine[k]) if len(set(counts)) == 1 else [dist.broadcast(bufs[r] if r != rank else mine[k], r) for r in range(world)]
            if len(set(counts)) != 1:
                bufs[rank] = mine[k]
            gathered.append(torch.cat(bufs).cpu().numpy())
        if rank == 0:
            one = H.gpu_case(name, n, L.F64, L.MATH_FAST, storage=storage)
            H.gpu_setup(one, name)
            one.step(steps)
            ref = one.get_fields()
            same = all(np.array_equal(g, r) for g, r in zip(gathered, ref))
            print(f"[mgpu] {name}: world={world} nlattice={sum(counts)} bitwise_equal={same}", flush=True)
            ok = ok and same and sum(counts) == len(ref[0])
        c.close()
        dist.barrier()
    flag_t = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag_t, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("[mgpu] ALL OK" if ok else "[mgpu] FAILED", flush=True)
    sys.exit(0 if int(flag_t.item()) else 1)


if __name__ == "__main__":
    main()
