#!/usr/bin/env python3
"""GPU probe (torchrun, not part of the product): where the set-up time of a DistMatrix goes (BK_DIST_TIMING=1)."""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

os.environ["BK_DIST_TIMING"] = "1"
from pytorch_sparse_solver import distributed as bkd, problems  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 256
    npl, ppg = 2 * n, n // 4
    rows = n ** 3
    offsets = [q * rows for q in range(world + 1)]
    crow, col, val = problems.stencil3d_rows(npl, world * ppg, rank * ppg, (rank + 1) * ppg, device=dev)
    for rep in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        D = bkd.DistMatrix(crow, col, val, offsets, rank, world)
        torch.cuda.synchronize()
        total = 1e3 * (time.perf_counter() - t0)
        if rank in (0, world // 2):
            print(json.dumps(dict(what="dist_setup", rank=rank, world=world, rep=rep, total_ms=round(total, 1), **D.setup_ms)),
                  flush=True)
        D.close()
    # the e2e sequence of bench.py, phase by phase (syncs between the phases)
    from pytorch_sparse_solver import module_a
    from pytorch_sparse_solver.module_a import krylov
    crow_h, col_h, val_h = (t.cpu().pin_memory() for t in (crow, col, val))
    b_h = torch.ones(rows, dtype=torch.float64).pin_memory()
    del crow, col, val
    for rep in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        ph = {}
        t0 = time.perf_counter()

        def lap(name):
            torch.cuda.synchronize()
            nonlocal_t = time.perf_counter()
            ph[name] = round(1e3 * (nonlocal_t - lap.t), 1)
            lap.t = nonlocal_t
        lap.t = t0
        c1, c2, c3 = crow_h.to(dev, non_blocking=True), col_h.to(dev, non_blocking=True), val_h.to(dev, non_blocking=True)
        bd = b_h.to(dev, non_blocking=True)
        lap("h2d")
        D2 = bkd.DistMatrix(c1, c2, c3, offsets, rank, world)
        lap("DistMatrix")
        xe, info = module_a.cg(D2, bd, tol=1e-8)
        lap("module_a.cg")
        xh = xe.cpu()
        lap("d2h")
        dist.barrier()
        lap("barrier")
        if rank == 0:
            print(json.dumps(dict(what="dist_e2e", world=world, rep=rep, total_ms=round(1e3 * (time.perf_counter() - t0), 1),
                                  iterations=int(krylov.last_result["iterations"]),
                                  device_ms=round(krylov.last_result["device_ms"], 1), **ph, setup=D2.setup_ms)), flush=True)
        D2.close()
        del c1, c2, c3, D2
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
