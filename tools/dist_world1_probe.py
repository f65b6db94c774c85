#!/usr/bin/env python3
"""GPU probe: the distributed code path on ONE rank (world 1) — which SpMV kernel does the folded matvec launch?
Run under `ncu --metrics gpu__time_duration.sum` for the launch list."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29577")
from pytorch_sparse_solver import distributed as bkd, problems  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
n = 256
npl, ppg = 2 * n, n // 4
rows = n ** 3
crow, col, val = problems.stencil3d_rows(npl, ppg, 0, ppg, device=dev)
D = bkd.DistMatrix(crow, col, val, [0, rows], 0, 1)
print("local_info", D.local_info(), flush=True)
b = torch.ones(rows, dtype=torch.float64, device=dev)
for _ in range(3):
    x, r = D.cg(b, None, 0.0, 0.0, 40)
torch.cuda.synchronize()
print("iterations", r["iterations"], "device_ms", r["device_ms"], "us/iter", 1e3 * r["device_ms"] / r["iterations"], flush=True)
D.close()
dist.destroy_process_group()
