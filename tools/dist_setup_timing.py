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
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
