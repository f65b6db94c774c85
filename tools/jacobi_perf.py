#!/usr/bin/env python3
"""Built-in Jacobi preconditioner at scale: badly scaled Poisson / convection-diffusion n^3 systems (values vary, so
the SpMV runs on the coded-column kernel 3), fixed-iteration windows for the time per iteration and full solves for
the iteration counts with and without M.  One JSON line per measurement."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
import torch  # noqa: E402

from pytorch_sparse_solver import module_a as ma, problems  # noqa: E402
from pytorch_sparse_solver.module_a import krylov  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


for name, A, solver in (("cg", problems.scaled_poisson3d_csr(n, device=dev), ma.cg),
                        ("bicgstab", problems.scaled_convdiff3d_csr(n, device=dev), ma.bicgstab)):
    N = A.shape[0]
    xt = torch.randn(N, dtype=torch.float64, device=dev, generator=torch.Generator(dev).manual_seed(0))
    b = torch.sparse.mm(A, xt[:, None])[:, 0]
    M = ma.JacobiPreconditioner(A)
    for label, Mx in (("jacobi", M), ("plain", None)):
        dt, _ = timed(lambda: solver(A, b, tol=0.0, atol=0.0, maxiter=100, M=Mx))
        print(json.dumps({"solver": name, "M": label, "n": N, "window_iterations": 100,
                          "us_per_iteration": 1e6 * dt / 100, "route": krylov.last_result["route"],
                          "kernel": krylov.last_result.get("kernel")}), flush=True)
    t0 = time.perf_counter()
    x, info = solver(A, b, tol=1e-8, M=M)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rel = float(torch.linalg.norm(x - xt) / torch.linalg.norm(xt))
    print(json.dumps({"solver": name, "M": "jacobi", "n": N, "tol": 1e-8, "info": info,
                      "iterations": krylov.last_result["iterations"], "seconds": dt, "x_rel_err": rel}), flush=True)
    t0 = time.perf_counter()
    x, info = solver(A, b, tol=1e-8, maxiter=20000)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"solver": name, "M": "plain", "n": N, "tol": 1e-8, "info": info,
                      "iterations": krylov.last_result["iterations"], "seconds": dt, "maxiter": 20000}), flush=True)
