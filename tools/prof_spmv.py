#!/usr/bin/env python3
"""Short driver for ncu captures of the SpMV kernels alone: python tools/prof_spmv.py [n] [use_compress] [reps]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
import torch  # noqa: E402

from pytorch_sparse_solver import _native, problems  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
uc = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
dev = torch.device("cuda", 0)
h = _native.Handle.get(dev)
h.set_option("use_compress", uc)
A = problems.poisson3d_csr(n, device=dev)
m = _native.register_matrix(A)
x = torch.randn(A.shape[0], dtype=torch.float64, device=dev)
for _ in range(reps):
    m.spmv_dot(x, x)
torch.cuda.synchronize()
print("ok kernel", m.info()["kernel"])
