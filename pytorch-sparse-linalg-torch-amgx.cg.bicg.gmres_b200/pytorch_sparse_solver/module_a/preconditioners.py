"""Built-in preconditioners for Module A.

The reference takes `M` as an arbitrary Python callable (torch_sparse_linalg.py:821, :849), which forces every
application through the interpreter.  `JacobiPreconditioner(A)` is such a callable — `M(r) = r / diag(A)`, exactly
what users of the reference write as `M = lambda r: r / d` — that `cg()`, `bicgstab()` and `gmres()` additionally
recognise: for a CUDA matrix the whole preconditioned iteration then stays on the device (`bk_cg_jacobi`,
`bk_bicgstab_jacobi`, `bk_gmres_jacobi`), with results identical to passing the lambda to the reference (golden
fixtures `*_jacobi*`).  With a callable `A` or on CPU tensors it behaves like any other callable `M`.
"""
from __future__ import annotations

import torch

from .. import _native


class JacobiPreconditioner:
    """M = diag(A)^-1, applied as r / d."""

    def __init__(self, A: torch.Tensor):
        if type(A).__name__ == "DistMatrix" and hasattr(A, "diagonal"):   # row-partitioned: this rank's slice
            d = A.diagonal()
            if bool((d == 0).any()):
                raise ValueError("JacobiPreconditioner: the matrix has a zero on its diagonal")
            self.d = d
            self.shape = (A.n_global, A.n_global)
            return
        if not isinstance(A, torch.Tensor) or A.ndim != 2 or A.shape[0] != A.shape[1]:
            raise ValueError("JacobiPreconditioner needs a square 2-D tensor (dense, COO or CSR)")
        with torch.no_grad():
            Ad = A.detach()
            if Ad.layout == torch.strided:
                d = torch.diagonal(Ad).clone()
            elif Ad.is_cuda and not Ad.is_complex():
                wdt = torch.float32 if Ad.dtype == torch.float32 else torch.float64
                d = _native.register_matrix(Ad, wdt).diagonal()
            else:
                C = Ad.coalesce() if Ad.layout == torch.sparse_coo else Ad.to_sparse_coo().coalesce()
                i, v = C.indices(), C.values()
                m = i[0] == i[1]
                d = torch.zeros(A.shape[0], dtype=v.dtype, device=v.device).index_add_(0, i[0][m], v[m])
        if bool((d == 0).any()):
            raise ValueError("JacobiPreconditioner: the matrix has a zero on its diagonal")
        self.d = d
        self.shape = tuple(A.shape)

    def diagonal(self, dtype: torch.dtype, device) -> torch.Tensor:
        return self.d.to(device=device, dtype=dtype)

    def __call__(self, r):
        if isinstance(r, torch.Tensor):
            return r / self.d.to(device=r.device, dtype=r.dtype)
        raise TypeError("JacobiPreconditioner applies to a single tensor")
