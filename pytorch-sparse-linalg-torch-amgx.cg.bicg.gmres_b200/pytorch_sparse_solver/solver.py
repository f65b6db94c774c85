"""Unified front door with the reference's names and semantics for backend='module_a'
(reference src/pytorch_sparse_solver/solver.py: SparseSolver :84, _select_backend :194, solve :256,
_solve_module_a :320, module-level solve/cg/bicgstab/gmres :524-576).

Only Module A exists in this build: AMGX (module_b) and cuDSS (module_c) are outside the hot-path scope,
report as unavailable and raise the same ValueError the reference raises for a missing backend.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Callable, Dict, List, Optional, Tuple, Union

import torch

from .utils.availability import get_available_backends


class SolverMethod(Enum):
    CG = "cg"
    BICGSTAB = "bicgstab"
    GMRES = "gmres"
    AMG = "amg"
    DIRECT = "direct"


class SolverBackend(Enum):
    MODULE_A = "module_a"
    MODULE_B = "module_b"
    MODULE_C = "module_c"
    AUTO = "auto"


@dataclass
class SolverResult:
    x: torch.Tensor
    converged: bool
    iterations: Optional[int]
    residual: Optional[float]
    backend: str
    method: str


class SparseSolver:
    def __init__(self, default_backend: str = "auto", default_method: str = "cg", verbose: bool = False):
        self.verbose = verbose
        self.default_backend = default_backend
        self.default_method = default_method
        self._available: Optional[Dict[str, bool]] = None
        self._module_a = None

    @property
    def available_backends(self) -> List[str]:
        if self._available is None:
            self._available = get_available_backends()
            if self.verbose:
                for name, ok in self._available.items():
                    print(f"  {'OK ' if ok else '-- '}{name}")
        return [k for k, v in self._available.items() if v]

    def _load_module_a(self):
        if self._module_a is None:
            try:
                from .module_a import cg, bicgstab, gmres
            except ImportError as e:  # pragma: no cover
                raise RuntimeError(f"Failed to load Module A: {e}")
            self._module_a = {'cg': cg, 'bicgstab': bicgstab, 'gmres': gmres}
        return self._module_a

    def _select_backend(self, backend: str, method: str, A) -> Tuple[str, str]:
        available = self.available_backends
        if not available:
            raise RuntimeError("No sparse solver backends are available!")
        if backend != "auto":
            if backend not in available:
                raise ValueError(f"Backend '{backend}' is not available. Available backends: {available}")
            return backend, method
        if method == "direct":
            raise ValueError("Direct solver requires Module C (cuDSS), which is not available. "
                             "Use an iterative method (cg, bicgstab, gmres) instead.")
        if method == "amg":
            raise ValueError("AMG solver requires Module B (AMGX), which is not available.")
        return "module_a", method

    def solve(self, A: Union[torch.Tensor, Callable], b: torch.Tensor, x0: Optional[torch.Tensor] = None,
              method: Optional[str] = None, backend: Optional[str] = None, tol: float = 1e-5, atol: float = 0.0,
              maxiter: Optional[int] = None, M: Optional[Callable] = None, **kwargs
              ) -> Tuple[torch.Tensor, SolverResult]:
        method = self.default_method if method is None else method
        backend = self.default_backend if backend is None else backend
        selected_backend, selected_method = self._select_backend(backend, method, A)
        if self.verbose:
            print(f"Using backend: {selected_backend}, method: {selected_method}")
        if selected_backend != "module_a":  # pragma: no cover - unreachable: only module_a is ever available
            raise ValueError(f"Unknown backend: {selected_backend}")
        return self._solve_module_a(A, b, x0, selected_method, tol, atol, maxiter, M, **kwargs)

    def _solve_module_a(self, A, b, x0, method, tol, atol, maxiter, M, **kwargs):
        module = self._load_module_a()
        if method not in module:
            raise ValueError(f"Method '{method}' not available in Module A. Use: {list(module.keys())}")
        solve_kwargs = {'tol': tol, 'atol': atol}
        if maxiter is not None:
            solve_kwargs['maxiter'] = maxiter
        if M is not None:
            solve_kwargs['M'] = M
        if x0 is not None:
            solve_kwargs['x0'] = x0
        if method == 'gmres':
            for k in ('restart', 'solve_method'):
                if k in kwargs:
                    solve_kwargs[k] = kwargs[k]
        res = {}
        x, info = module[method](A, b, _result=res, **solve_kwargs)   # the solve's own record, not a module global
        iterations = None
        if res.get("route") in ("native", "host") and M is None:
            # ||b - A x|| / ||b|| from the library's own final true-residual pass (reference recomputes it
            # with torch.mv, solver.py:362-368 — same quantity, no cuSPARSE on our path)
            bn = res["b_norm"]
            residual = res["final_residual"] / bn if bn != 0.0 else float('nan')
            iterations = int(res["iterations"])
        elif res.get("route") == "native":
            # built-in preconditioner: the library's final check is on ||M (b - A x)|| (as the reference's, :1008),
            # the router reports the UNpreconditioned relative residual (solver.py:362-368) — one more SpMV of ours
            from . import _native
            with torch.no_grad():
                xw = x.detach()
                r = b.detach().to(xw.dtype) - _native.register_matrix(A, xw.dtype).spmv(xw.contiguous())
                bn = float(_native.nrm2(b.detach().to(xw.dtype).contiguous()))
                residual = float(_native.nrm2(r)) / bn if bn != 0.0 else float('nan')
            iterations = int(res["iterations"])
        else:
            with torch.no_grad():
                Ax = A(x) if callable(A) and not isinstance(A, torch.Tensor) else torch.mv(A, x)
                residual = torch.norm(b - Ax).item() / torch.norm(b).item()
        result = SolverResult(x=x, converged=(info == 0), iterations=iterations, residual=residual,
                              backend="module_a", method=method)
        return x, result

    def cg(self, A, b, **kwargs):
        return self.solve(A, b, method='cg', **kwargs)

    def bicgstab(self, A, b, **kwargs):
        return self.solve(A, b, method='bicgstab', **kwargs)

    def gmres(self, A, b, **kwargs):
        return self.solve(A, b, method='gmres', **kwargs)

    def amg(self, A, b, **kwargs):
        return self.solve(A, b, method='amg', backend='module_b', **kwargs)

    def direct(self, A, b, **kwargs):
        return self.solve(A, b, method='direct', backend='module_c', **kwargs)

    def __repr__(self) -> str:
        return (f"SparseSolver(\n  available_backends={self.available_backends},\n"
                f"  default_backend='{self.default_backend}',\n  default_method='{self.default_method}'\n)")


_default_solver: Optional[SparseSolver] = None


def _get_default_solver() -> SparseSolver:
    global _default_solver
    if _default_solver is None:
        _default_solver = SparseSolver()
    return _default_solver


def solve(A, b, method: str = "cg", backend: str = "auto", **kwargs) -> Tuple[torch.Tensor, SolverResult]:
    return _get_default_solver().solve(A, b, method=method, backend=backend, **kwargs)


def cg(A, b, **kwargs):
    return solve(A, b, method='cg', **kwargs)


def bicgstab(A, b, **kwargs):
    return solve(A, b, method='bicgstab', **kwargs)


def gmres(A, b, **kwargs):
    return solve(A, b, method='gmres', **kwargs)


def amg(A, b, **kwargs):
    return solve(A, b, method='amg', backend='module_b', **kwargs)


def direct_solve(A, b, **kwargs):
    return solve(A, b, method='direct', backend='module_c', **kwargs)
