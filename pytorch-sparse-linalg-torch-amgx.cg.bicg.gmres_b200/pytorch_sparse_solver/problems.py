"""Vectorised generators of the synthetic systems named in BASELINE.json / SURVEY.md §8d.

All deterministic.  3-D row index r = (i*n + j)*n + k, columns ascending per row, Dirichlet = out-of-range
neighbours dropped (consistent with the reference's 2-D convention idx = i*ny + j, matrix_utils.py:213-214).
The reference has no 3-D generators; its 2-D builder and the LDC pressure matrix loop in Python
(matrix_utils.py:193-257, FVM_example/LDC_by_torchsp/ldc_solver_common.py:90-135) — these produce the same
matrices without the loops.  CSR index dtype defaults to int64 like torch's own CSR tensors.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch


def _csr_from_masked_stencil(N: int, offsets: Sequence[int], values: Sequence[float], masks: torch.Tensor,
                             dtype, index_dtype, device) -> torch.Tensor:
    r = torch.arange(N, device=device)
    offs = torch.tensor(list(offsets), device=device)
    cols = (r[:, None] + offs[None, :])[masks]
    vals = torch.tensor(list(values), dtype=dtype, device=device)[None, :].expand(N, len(offsets))[masks]
    crow = torch.zeros(N + 1, dtype=torch.int64, device=device)
    crow[1:] = masks.sum(1).cumsum(0)
    return torch.sparse_csr_tensor(crow.to(index_dtype), cols.to(index_dtype), vals, size=(N, N))


def stencil3d_csr(n: int, lower=(-1.0, -1.0, -1.0), upper=(-1.0, -1.0, -1.0), diag: float = 6.0,
                  dtype=torch.float64, index_dtype=torch.int64, device="cpu", nz: int = None) -> torch.Tensor:
    """7-point stencil on an (nz or n) x n x n grid (slab along i when nz is given)."""
    ni = n if nz is None else nz
    N = ni * n * n
    r = torch.arange(N, device=device)
    k = r % n
    j = (r // n) % n
    i = r // (n * n)
    offsets = [-n * n, -n, -1, 0, 1, n, n * n]
    values = [lower[0], lower[1], lower[2], diag, upper[2], upper[1], upper[0]]
    masks = torch.stack([i > 0, j > 0, k > 0, torch.ones(N, dtype=torch.bool, device=device), k < n - 1, j < n - 1,
                         i < ni - 1], 1)
    return _csr_from_masked_stencil(N, offsets, values, masks, dtype, index_dtype, device)


def stencil3d_rows(n: int, ni_total: int, i_begin: int, i_end: int, lower=(-1.0, -1.0, -1.0),
                   upper=(-1.0, -1.0, -1.0), diag: float = 6.0, dtype=torch.float64, index_dtype=torch.int64,
                   device="cpu") -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Rows of planes i in [i_begin, i_end) of the 7-point stencil on an ni_total x n x n grid, as CSR arrays
    (crow, col, val) with GLOBAL column indices — the slab one rank owns in a 1-D row partition."""
    rows = (i_end - i_begin) * n * n
    r = torch.arange(i_begin * n * n, i_end * n * n, device=device)
    k = r % n
    j = (r // n) % n
    i = r // (n * n)
    offs = torch.tensor([-n * n, -n, -1, 0, 1, n, n * n], device=device)
    vals = torch.tensor([lower[0], lower[1], lower[2], diag, upper[2], upper[1], upper[0]], dtype=dtype, device=device)
    masks = torch.stack([i > 0, j > 0, k > 0, torch.ones(rows, dtype=torch.bool, device=device), k < n - 1, j < n - 1,
                         i < ni_total - 1], 1)
    cols = (r[:, None] + offs[None, :])[masks]
    v = vals[None, :].expand(rows, 7)[masks]
    crow = torch.zeros(rows + 1, dtype=torch.int64, device=device)
    crow[1:] = masks.sum(1).cumsum(0)
    return crow.to(index_dtype), cols.to(index_dtype), v


def poisson3d_csr(n: int, **kw) -> torch.Tensor:
    """P3D-n: diag 6, six off-diagonals -1 (SPD)."""
    return stencil3d_csr(n, **kw)


def scaled_convdiff3d_csr(n: int, seed: int = 7, spread: float = 3.0, **kw) -> torch.Tensor:
    """S A S with A = CD3D-n (non-symmetric) and the same diagonal scaling as scaled_poisson3d_csr."""
    return _diag_scaled(convdiff3d_csr(n, **kw), seed, spread)


def scaled_poisson3d_csr(n: int, seed: int = 7, spread: float = 3.0, **kw) -> torch.Tensor:
    """S A S with A = P3D-n and S = diag(10^(spread * u)), u ~ U(-0.5, 0.5) (CPU generator, seeded): SPD, same
    sparsity, strongly varying diagonal — the textbook case for a Jacobi preconditioner."""
    return _diag_scaled(stencil3d_csr(n, **kw), seed, spread)


def _diag_scaled(A: torch.Tensor, seed: int, spread: float) -> torch.Tensor:
    N = A.shape[0]
    u = torch.rand(N, dtype=torch.float64, generator=torch.Generator().manual_seed(seed)) - 0.5
    sc = torch.pow(torch.tensor(10.0, dtype=torch.float64), spread * u).to(A.device)
    crow, col = A.crow_indices(), A.col_indices()
    rows = torch.repeat_interleave(torch.arange(N, device=A.device), crow[1:] - crow[:-1])
    vals = (A.values().double() * sc[rows] * sc[col.long()]).to(A.values().dtype)
    return torch.sparse_csr_tensor(crow, col, vals, size=(N, N))


def csr_diagonal(A: torch.Tensor) -> torch.Tensor:
    """diag(A) of a CSR tensor as a dense vector (set-up helper for tests / fixtures)."""
    N = A.shape[0]
    crow, col = A.crow_indices(), A.col_indices()
    rows = torch.repeat_interleave(torch.arange(N, device=A.device), crow[1:] - crow[:-1])
    d = torch.zeros(N, dtype=A.values().dtype, device=A.device)
    m = rows == col
    d.index_add_(0, rows[m], A.values()[m])
    return d


def convdiff3d_csr(n: int, gamma=(1.0, 0.5, 0.25), **kw) -> torch.Tensor:
    """CD3D-n: first-order upwind convection-diffusion, non-symmetric M-matrix."""
    g = gamma
    return stencil3d_csr(n, lower=(-(1 + g[0]), -(1 + g[1]), -(1 + g[2])), diag=6 + sum(g), **kw)


def poisson2d_csr(nx: int, ny: int, dtype=torch.float64, index_dtype=torch.int64, device="cpu") -> torch.Tensor:
    """Same matrix as create_poisson_2d_sparse_coo(nx, ny).to_sparse_csr() (idx = i*ny + j, diag 4, off -1)."""
    N = nx * ny
    r = torch.arange(N, device=device)
    i, j = r // ny, r % ny
    offsets = [-ny, -1, 0, 1, ny]
    values = [-1.0, -1.0, 4.0, -1.0, -1.0]
    masks = torch.stack([i > 0, j > 0, torch.ones(N, dtype=torch.bool, device=device), j < ny - 1, i < nx - 1], 1)
    return _csr_from_masked_stencil(N, offsets, values, masks, dtype, index_dtype, device)


def ldc_pressure_csr(nx: int, ny: int = None, lx: float = 1.0, ly: float = 1.0, dtype=torch.float64,
                     index_dtype=torch.int64, device="cpu") -> torch.Tensor:
    """Pressure Poisson matrix of the lid-driven-cavity example (ldc_solver_common.py:90-135): 5-point,
    all-Neumann (singular, symmetric negative semidefinite), row index i = row*nx + col, coefficients 1/dx^2,
    1/dy^2, Ap = -(Aw + Ae + An + As)."""
    ny = nx if ny is None else ny
    dx, dy = lx / nx, ly / nx  # the reference uses lx/nx for both (:50-51)
    dx2, dy2 = dx ** 2, dy ** 2
    N = nx * ny
    r = torch.arange(N, device=device)
    row, col = r // nx, r % nx
    aw = torch.where(col > 0, 1.0 / dx2, 0.0).to(dtype)
    ae = torch.where(col < nx - 1, 1.0 / dx2, 0.0).to(dtype)
    a_n = torch.where(row < ny - 1, 1.0 / dy2, 0.0).to(dtype)
    a_s = torch.where(row > 0, 1.0 / dy2, 0.0).to(dtype)
    ap = -(aw + ae + a_n + a_s)
    masks = torch.stack([row > 0, col > 0, torch.ones(N, dtype=torch.bool, device=device), col < nx - 1,
                         row < ny - 1], 1)
    vals = torch.stack([a_s, aw, ap, ae, a_n], 1)[masks]
    offs = torch.tensor([-nx, -1, 0, 1, nx], device=device)
    cols = (r[:, None] + offs[None, :])[masks]
    crow = torch.zeros(N + 1, dtype=torch.int64, device=device)
    crow[1:] = masks.sum(1).cumsum(0)
    return torch.sparse_csr_tensor(crow.to(index_dtype), cols.to(index_dtype), vals, size=(N, N))


def manufactured_rhs(A: torch.Tensor, seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """b = A x_true with x_true = randn(N, fp64, CPU generator seeded `seed`)."""
    n = A.shape[0]
    xt = torch.randn(n, dtype=torch.float64, generator=torch.Generator().manual_seed(seed))
    xt = xt.to(A.device)
    with torch.no_grad():
        b = torch.matmul(A, xt) if not A.is_cuda else None
    if b is None:
        from . import _native
        b = _native.register_matrix(A, torch.float64).spmv(xt.to(torch.float64))
    return b, xt


def cg_bytes_per_iteration(n: int, nnz: int, value_bytes: int = 8, index_bytes: int = 4) -> int:
    """Algorithmic HBM bytes of one CG iteration: nnz*(sv+si) + (n+1)*si + 11*n*sv."""
    return nnz * (value_bytes + index_bytes) + (n + 1) * index_bytes + 11 * n * value_bytes


def bicgstab_bytes_per_iteration(n: int, nnz: int, value_bytes: int = 8, index_bytes: int = 4) -> int:
    return 2 * (nnz * (value_bytes + index_bytes) + (n + 1) * index_bytes) + 19 * n * value_bytes


def gmres_bytes_per_cycle(n: int, nnz: int, m: int, value_bytes: int = 8, index_bytes: int = 4) -> int:
    return (m + 1) * (nnz * (value_bytes + index_bytes) + (n + 1) * index_bytes) + (m * (m + 1) + 8 * m + 7) * n * value_bytes
