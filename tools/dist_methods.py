#!/usr/bin/env python3
"""Weak-scaling timing of the distributed BiCGStab and GMRES(30) drivers (torchrun, one rank per GPU).

Geometry = bench.py's multi-GPU CG run (BASELINE config 5): planes of 2n x 2n, n/4 planes per GPU, i.e. n^3 rows per
GPU; convection-diffusion coefficients (SURVEY config 3), manufactured right-hand side.  Fixed windows (tol = 0).
Prints one JSON line per (method, path); run it with --nproc-per-node 1 for the single-GPU denominators."""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from pytorch_sparse_solver import problems
    from pytorch_sparse_solver import distributed as bkd
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    npl, ppg = 2 * n, max(n // 4, 1)
    rows = ppg * npl * npl
    offsets = [q * rows for q in range(world + 1)]
    g3 = (1.0, 0.5, 0.25)
    cd = dict(lower=(-(1 + g3[0]), -(1 + g3[1]), -(1 + g3[2])), diag=6 + sum(g3))
    crow, col, val = problems.stencil3d_rows(npl, world * ppg, rank * ppg, (rank + 1) * ppg, device=dev, **cd)
    nnz_local = val.numel()
    D = bkd.DistMatrix(crow, col, val, offsets, rank, world)
    del crow, col
    xt = torch.randn(rows, dtype=torch.float64, device=dev, generator=torch.Generator(dev).manual_seed(rank))
    b = D.spmv(xt)
    mat_bytes = nnz_local * 12 + 4 * (rows + 1)
    m = 30
    cases = {
        "bicgstab": (lambda: D.bicgstab(b, None, 0.0, 0.0, 100), 2 * mat_bytes + 19 * rows * 8, "iteration"),
        "gmres30": (lambda: D.gmres(b, None, 0.0, 0.0, m, 3, "batched"),
                    (m + 1) * mat_bytes + (m * (m + 1) + 8 * m + 7) * rows * 8, "cycle"),
    }
    for p2p in ((1, 0) if D.p2p else (0,)):
        D.handle.set_option("dist_p2p", p2p)
        for name, (fn, nbytes, unit) in cases.items():
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            units = 0
            e0.record()
            for _ in range(reps):
                _, r = fn()
                units += r["iterations"]
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            per = float(ms) / units
            if rank == 0:
                print(json.dumps({"method": name, "n_gpus": world, "path": "peer-memory" if p2p else "nccl",
                                  "rows_per_gpu": rows, "unit": unit, "ms_per_unit": per,
                                  "units_per_s": 1e3 / per, "bytes_per_unit_per_gpu": nbytes,
                                  "achieved_gbs_per_gpu": nbytes / per / 1e6,
                                  "frac_of_8tbs": nbytes / per / 1e6 / 8000.0, "info": r["info"],
                                  "status": r["status"]}), flush=True)
    D.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
