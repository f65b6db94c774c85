#!/usr/bin/env python3
"""Small driver for compute-sanitizer (memcheck / racecheck): every kernel family once on tiny systems."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
import torch  # noqa: E402

from pytorch_sparse_solver import _native, module_a, problems  # noqa: E402

dev = torch.device("cuda", 0)
h = _native.Handle.get(dev)
g = torch.Generator().manual_seed(0)
for make in (lambda: problems.poisson3d_csr(9), lambda: problems.convdiff3d_csr(7), lambda: problems.poisson2d_csr(33, 17),
             lambda: torch.randn(70, 70, dtype=torch.float64, generator=g).add_(torch.eye(70, dtype=torch.float64) * 20).to_sparse_csr()):
    A = make().cuda()
    n = A.shape[0]
    m = _native.register_matrix(A)
    x = torch.randn(n, dtype=torch.float64, generator=g).cuda()
    y, d = m.spmv_dot(x, x)
    t = m.transpose()
    t.spmv(x)
    b = m.spmv(x)
    for mode in (0, 1):
        h.set_option("loop_mode", mode)
        module_a.cg(A, b, tol=1e-8, maxiter=50)
        module_a.bicgstab(A, b, tol=1e-8, maxiter=50)
        module_a.gmres(A, b, tol=1e-8, restart=7, maxiter=5)
        module_a.gmres(A, b, tol=1e-8, restart=7, maxiter=5, solve_method="incremental")
    h.set_option("loop_mode", 0)
    h.set_option("use_tma", 0)
    _native.clear_cache()
    module_a.cg(A, b, tol=1e-8, maxiter=20)
    h.set_option("fuse_xpay", 1)
    module_a.cg(A, b, tol=1e-8, maxiter=20)
    h.set_option("fuse_xpay", 0)
    h.set_option("use_tma", 1)
    _native.clear_cache()
    bb = b.clone().requires_grad_(True)
    xs, _ = module_a.cg(A, bb, tol=1e-8, maxiter=50)
    xs.sum().backward()
A32 = problems.poisson3d_csr(8)
A32 = torch.sparse_csr_tensor(A32.crow_indices().cuda(), A32.col_indices().cuda(), A32.values().float().cuda(), size=A32.shape)
module_a.cg(A32, torch.ones(512, dtype=torch.float32, device=dev), tol=1e-5)
Ac = problems.poisson3d_csr(8)
module_a.cg(Ac, torch.ones(512, dtype=torch.float64), tol=1e-8)     # host route
torch.cuda.synchronize()
print("sanitize driver ok")
