"""Mixed-precision iterative refinement: fp32 inner Krylov solves, fp64 residuals and solution.

The reference computes everything in fp64 (:979-980) and cannot run fp32 tensors at all.  The Krylov loop is
HBM-bound, so an fp32 inner solve moves half the vector bytes per iteration; classical iterative refinement
    r = b - A x (fp64) ;  solve A d = r in fp32 to a loose tolerance ;  x += d (fp64)
recovers an fp64-accurate solution whenever cond(A) * eps_fp32 < 1.  Every vector operation runs in the library
(bk_spmv, bk_axpby, bk_nrm2 and the native fp32 solvers); the outer loop costs one host sync per refinement.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import _native


def refined_solve(A: torch.Tensor, b: torch.Tensor, x0: Optional[torch.Tensor] = None, *, method: str = "cg",
                  tol: float = 1e-10, atol: float = 0.0, inner_tol: float = 1e-4, max_refinements: int = 40,
                  maxiter: Optional[int] = None, restart: int = 20) -> Tuple[torch.Tensor, int]:
    """Solve A x = b to the fp64 tolerance `tol` (same stop rule as the reference: ||b - A x|| <= max(tol ||b||, atol),
    thresholds rounded through fp32 like torch.tensor(tol), :1008-1016) with fp32 inner `method` solves
    ('cg' | 'bicgstab' | 'gmres').  Returns (x fp64, info) with info 0 when the fp64 residual met the tolerance."""
    from . import krylov
    if method not in ("cg", "bicgstab", "gmres"):
        raise ValueError(f"unknown method {method}")
    if not isinstance(A, torch.Tensor) or A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise ValueError("refined_solve needs A as a square 2-D tensor")
    if not A.is_cuda or not b.is_cuda:
        raise _native.NativeLibraryError("refined_solve runs on CUDA tensors (no CPU fallback)")
    with torch.no_grad():
        m64 = _native.register_matrix(A, torch.float64)
        m32 = _native.register_matrix(A, torch.float32)
        b64 = b.detach().to(torch.float64).contiguous()
        bn = float(_native.nrm2(b64))
        thr = max(float(torch.tensor(tol)) * bn, float(torch.tensor(atol)))
        x = torch.zeros_like(b64) if x0 is None else x0.detach().to(torch.float64).contiguous().clone()
        r = _native.axpby(1.0, b64, -1.0, m64.spmv(x)) if x0 is not None else b64.clone()
        inner_its, k, rn = 0, 0, float("inf")
        while True:
            rn = float(_native.nrm2(r))
            if rn <= thr or k >= max_refinements or rn != rn:
                break
            # the inner system is solved for the UNIT residual: fp32 range, and the solvers' absolute safeguards
            # (BiCGStab's |t.t| < eps, GMRES's atol floor) are written for O(1) data
            r32 = _native.div_scalar(r, rn).to(torch.float32)
            if method == "cg":
                d32, res = m32.cg(r32, None, inner_tol, 0.0, maxiter)
            elif method == "bicgstab":
                d32, res = m32.bicgstab(r32, None, inner_tol, 0.0, maxiter)
            else:
                te, ae = krylov._gmres_effective_tolerances(inner_tol, 0.0, r32.numel(), "cuda")
                d32, res = m32.gmres(r32, None, te, ae, restart, maxiter, _native.BK_GMRES_BATCHED)
            inner_its += int(res["iterations"])
            x = _native.axpby(1.0, x, rn, d32.to(torch.float64))
            r = _native.axpby(1.0, b64, -1.0, m64.spmv(x))
            k += 1
        info = 0 if rn <= thr else -1
    krylov._publish(dict(iterations=inner_its, refinements=k, info=info, final_residual=rn, threshold=thr, b_norm=bn,
                         x_norm=float(_native.nrm2(x)), solver=f"refined_{method}", route="native"), None)
    return x.reshape(b.shape), info
