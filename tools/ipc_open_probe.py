"""What does mapping a neighbour's population buffer cost (cudaIpcOpenMemHandle through lbm_p2p_open)?
2+ ranks under torchrun.  Times: a small buffer first, a large one, both again after closing, and the large one
while a small "keep-alive" mapping stays open.  MEASUREMENT INFRASTRUCTURE."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import lattice_boltzmann_method_gpu_b200 as L  # noqa: E402
from lattice_boltzmann_method_gpu_b200 import api  # noqa: E402


def make(n, local):
    d = L.case_defaults(L.CASE_LDC)
    d.nx = d.ny = d.nz = n
    d.z_begin, d.z_end = 0, n
    d.precision, d.storage, d.device = L.F64, L.STORE_DENSE_AA, local
    c = L.Case(d)
    c.geo_pre(), c.index_transform(), c.initialize()
    return c


def exchange(c, world):
    e = c.p2p_export()
    rec = np.frombuffer(e["handles"][0], dtype=np.uint8).copy()
    t = torch.from_numpy(rec).cuda()
    allt = torch.empty(world * 64, dtype=torch.uint8, device="cuda")
    dist.all_gather_into_tensor(allt, t)
    return allt.cpu().numpy().reshape(world, 64)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nb = (rank + 1) % world
    small, big = make(32, local), make(448, local)  # 5 MB and 13.7 GB of populations
    hs, hb = exchange(small, world), exchange(big, world)
    dist.barrier()
    out = {}
    for tag, seq in (("first", ("small", "big")), ("second", ("small", "big"))):
        for which in seq:
            h = (hs if which == "small" else hb)[nb].tobytes()
            t0 = time.perf_counter()
            api.p2p_open(h)
            out[f"{tag}_{which}_open_ms"] = (time.perf_counter() - t0) * 1e3
        dist.barrier()
        for which in seq:
            h = (hs if which == "small" else hb)[nb].tobytes()
            t0 = time.perf_counter()
            api.p2p_release(h)
            out[f"{tag}_{which}_close_ms"] = (time.perf_counter() - t0) * 1e3
        dist.barrier()
    # keep the small mapping open, reopen the big one twice
    api.p2p_open(hs[nb].tobytes())
    for k in range(2):
        t0 = time.perf_counter()
        api.p2p_open(hb[nb].tobytes())
        out[f"keepalive_big_open_{k}_ms"] = (time.perf_counter() - t0) * 1e3
        dist.barrier()
        api.p2p_release(hb[nb].tobytes())
        dist.barrier()
    api.p2p_release(hs[nb].tobytes())
    dist.barrier()
    if rank == 0:
        print({k: round(v, 2) for k, v in out.items()}, flush=True)
    small.close(), big.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
