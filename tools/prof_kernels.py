#!/usr/bin/env python3
"""Short driver for ncu captures: a few launches of each hot kernel on the 256^3 Poisson system."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
import torch  # noqa: E402

from pytorch_sparse_solver import _native, problems  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
which = sys.argv[2] if len(sys.argv) > 2 else "cg"
dev = torch.device("cuda", 0)
if which == "cg":
    A = problems.poisson3d_csr(n, device=dev)
else:
    A = problems.convdiff3d_csr(n, device=dev)
m = _native.register_matrix(A)
N = A.shape[0]
x = torch.randn(N, dtype=torch.float64, device=dev)
b = torch.ones(N, dtype=torch.float64, device=dev)
for _ in range(3):
    m.spmv_dot(x, x)
torch.cuda.synchronize()
if which == "cg":
    m.cg(b, None, 0.0, 0.0, 4)
elif which == "bicgstab":
    m.bicgstab(b, None, 0.0, 0.0, 2)
else:
    m.gmres(b, None, 0.0, 0.0, 8, 1, 0)
torch.cuda.synchronize()
print("ok")
