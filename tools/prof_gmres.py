#!/usr/bin/env python3
"""Short driver for ncu captures of the GMRES kernels at full size: one GMRES(30) cycle on CD3D-256 (repeated)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
import torch  # noqa: E402

from pytorch_sparse_solver import _native, problems  # noqa: E402

dev = torch.device("cuda", 0)
A = problems.convdiff3d_csr(256, device=dev)
m = _native.register_matrix(A)
b = torch.ones(A.shape[0], dtype=torch.float64, device=dev)
for _ in range(2):
    x, r = m.gmres(b, None, 1e-30, 0.0, 30, 1, 0)
torch.cuda.synchronize()
print("cycles", r["iterations"], "matvecs", r["matvecs"], "device_ms", r["device_ms"])
