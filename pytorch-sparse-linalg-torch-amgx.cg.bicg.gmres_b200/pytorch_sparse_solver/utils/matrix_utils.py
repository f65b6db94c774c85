"""Matrix helpers with the reference's names (utils/matrix_utils.py).  The builders are vectorised (the
reference loops in Python, :143-257) but produce the same coalesced COO tensors: 2-D Poisson uses
idx = i*ny + j, diagonal 4, neighbours -1 (:193-257); tridiagonal diag/off-diag (:143-190)."""
from typing import Callable, Optional, Tuple, Union

import torch


def dense_to_sparse_csr(A: torch.Tensor, device: Optional[str] = None) -> torch.Tensor:
    if A.ndim != 2:
        raise ValueError(f"Expected 2D tensor, got {A.ndim}D")
    coo = A.to_sparse_coo()
    if device is not None and torch.device(device) != A.device:
        coo = coo.to(device)
    return coo.to_sparse_csr()


def sparse_coo_to_csr(sparse_coo: torch.Tensor) -> torch.Tensor:
    if not sparse_coo.is_sparse:
        raise ValueError("Input must be a sparse tensor")
    return sparse_coo.coalesce().to_sparse_csr()


def ensure_sparse_format(A: torch.Tensor, format: str = 'csr') -> torch.Tensor:
    if A.layout == torch.sparse_csr:
        if format == 'csr':
            return A
        A = A.to_sparse_coo()
    if not A.is_sparse:
        A = A.to_sparse_coo()
    if format == 'csr':
        return A.coalesce().to_sparse_csr()
    if format == 'coo':
        return A.coalesce()
    if format == 'csc':
        return A.coalesce().to_sparse_csc()
    raise ValueError(f"Unknown format: {format}. Use 'csr', 'coo', or 'csc'")


def get_csr_components(A: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(values, col_indices, row_ptr) — the reference's order (:86-105)."""
    if A.layout != torch.sparse_csr:
        A = ensure_sparse_format(A, 'csr')
    return A.values(), A.col_indices(), A.crow_indices()


def create_sparse_csr_from_components(values, col_indices, row_ptr, shape, device=None, dtype=None) -> torch.Tensor:
    device = values.device if device is None else device
    dtype = values.dtype if dtype is None else dtype
    return torch.sparse_csr_tensor(row_ptr.to(device), col_indices.to(device),
                                   values.to(device=device, dtype=dtype), size=shape)


def create_tridiagonal_sparse_coo(n: int, diag_val: float = 2.0, off_diag_val: float = -1.0, device: str = 'cpu',
                                  dtype: torch.dtype = torch.float64) -> torch.Tensor:
    i = torch.arange(n, device=device)
    rows = [i, i[:-1], i[1:]]
    cols = [i, i[1:], i[:-1]]
    vals = [torch.full((n,), diag_val, device=device, dtype=dtype),
            torch.full((max(n - 1, 0),), off_diag_val, device=device, dtype=dtype),
            torch.full((max(n - 1, 0),), off_diag_val, device=device, dtype=dtype)]
    idx = torch.stack([torch.cat(rows), torch.cat(cols)])
    return torch.sparse_coo_tensor(idx, torch.cat(vals), (n, n), device=device, dtype=dtype).coalesce()


def create_poisson_2d_sparse_coo(nx: int, ny: int, device: str = 'cpu',
                                 dtype: torch.dtype = torch.float64) -> torch.Tensor:
    n = nx * ny
    k = torch.arange(n, device=device)
    i, j = k // ny, k % ny
    rows, cols, vals = [k], [k], [torch.full((n,), 4.0, device=device, dtype=dtype)]
    for mask, off in ((i > 0, -ny), (i < nx - 1, ny), (j > 0, -1), (j < ny - 1, 1)):
        r = k[mask]
        rows.append(r)
        cols.append(r + off)
        vals.append(torch.full((r.numel(),), -1.0, device=device, dtype=dtype))
    idx = torch.stack([torch.cat(rows), torch.cat(cols)])
    return torch.sparse_coo_tensor(idx, torch.cat(vals), (n, n), device=device, dtype=dtype).coalesce()


def compute_residual(A: Union[torch.Tensor, Callable], x: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    if callable(A) and not isinstance(A, torch.Tensor):
        Ax = A(x)
    elif A.is_sparse:
        Ax = torch.sparse.mm(A, x.unsqueeze(-1)).squeeze(-1)
    else:
        Ax = torch.mv(A, x)
    return b - Ax


def compute_relative_residual(A, x: torch.Tensor, b: torch.Tensor) -> float:
    return (torch.norm(compute_residual(A, x, b)) / torch.norm(b)).item()
