#!/usr/bin/env python3
"""GPU probe (not part of the product): phases of the host-buffer solve (BK_HOST_TIMING=1 -> stderr) and the wall time
of module_a.cg(A_cpu_pinned, b_cpu_pinned) around it."""
import os
import sys
import time
from pathlib import Path

os.environ["BK_HOST_TIMING"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
import torch  # noqa: E402

from pytorch_sparse_solver import module_a, problems  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
idx = torch.int32 if (len(sys.argv) > 2 and sys.argv[2] == "int32") else torch.int64
A = problems.poisson3d_csr(n, index_dtype=idx)
A = torch.sparse_csr_tensor(A.crow_indices().pin_memory(), A.col_indices().pin_memory(), A.values().pin_memory(),
                            size=A.shape)
b = torch.ones(A.shape[0], dtype=torch.float64).pin_memory()
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, info = module_a.cg(A, b, tol=1e-8)
    torch.cuda.synchronize()
    print(f"rep {rep}: module_a.cg(A_cpu, b_cpu) wall {1e3 * (time.perf_counter() - t0):.2f} ms, info {info}, "
          f"indices {idx}", file=sys.stderr, flush=True)
