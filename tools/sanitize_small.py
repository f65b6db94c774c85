#!/usr/bin/env python3
"""Small cases for compute-sanitizer (memcheck / racecheck): kernel 7, kernel 6, the persistent solvers, lagged-x CG."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
import torch  # noqa: E402

from pytorch_sparse_solver import _native, module_a, problems  # noqa: E402

dev = torch.device("cuda", 0)
h = _native.Handle.get(dev)
for gen in (lambda: problems.poisson3d_csr(24, device=dev), lambda: problems.poisson2d_csr(70, 64, device=dev),
            lambda: problems.ldc_pressure_csr(40, device=dev)):
    A = gen()
    b = torch.ones(A.shape[0], dtype=torch.float64, device=dev)
    for persistent in (1, 0):
        h.set_option("persistent", persistent)
        h.set_option("fuse_xpay", -1 if persistent else 0)
        for fn, kw in ((module_a.cg, {}), (module_a.bicgstab, {}), (module_a.gmres, dict(restart=10, maxiter=3))):
            x, info = fn(A, b, tol=1e-6, **kw)
    m = _native.register_matrix(A)
    print("kernel", m.info()["kernel"], "n", A.shape[0], flush=True)
torch.cuda.synchronize()
print("sanitize_small done", flush=True)
