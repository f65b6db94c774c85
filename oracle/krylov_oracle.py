"""CPU oracle for Module A's Krylov path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain restatement (torch-CPU ops, single-tensor `b`, optional preconditioner callable) of the reference algorithms in
src/pytorch_sparse_solver/module_a/torch_sparse_linalg.py.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` leg may import this file; the product package never does.

Why a restatement and not the reference itself: /root/reference does not exist on the GPU box, and the
reference is pure Python on torch primitives, so stating the same recurrences with the same primitives in the
same order (torch.matmul(A, v) :191, torch.vdot :91, elementwise +,-,* as separate ops :165-173) reproduces it
exactly.  PINNED: oracle/pin_reference.py runs the unmodified reference (imported from /root/reference/src in
the build container) next to this file on every fixture case and requires bit-identical x / info / matvec
counts before writing tests/golden/*.npz; tests/test_oracle_golden.py re-checks the oracle against those
vectors everywhere.  The primitive arithmetic lives in PyTorch (reference pins torch>=2.0, pyproject.toml:44;
fixtures generated with torch 2.11.0+cu128 CPU/MKL).

Each function cites the reference lines it restates.  Extra (non-reference) outputs: iteration / matvec counts,
which the reference never reports (solver.py:373) — counted the way SURVEY.md §0 does (wrapping the matvec).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

_INV_SQRT2 = 0.7071067811865476  # :63


class _CountingMatvec:
    """_normalize_matvec (:176-208) for a 2-D tensor, single-leaf input, plus a call counter."""

    def __init__(self, A: torch.Tensor):
        if A.ndim != 2 or A.shape[0] != A.shape[1]:
            raise ValueError(f'linear operator must be a square matrix, but has shape: {A.shape}')
        self.A = A
        self.calls = 0

    def __call__(self, v: torch.Tensor) -> torch.Tensor:
        self.calls += 1
        v_flat = torch.cat([v.flatten()])                       # :188  (a copy, like the reference)
        out = torch.matmul(self.A, v_flat)                      # :191
        return out[0:v.numel()].reshape(v.shape)                # :194-203


def _vdot(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:   # :86-91 (real inputs)
    return torch.vdot(a.to(torch.float64).flatten(), b.to(torch.float64).flatten())


def _norm(x: torch.Tensor) -> torch.Tensor:                    # :154-162
    return torch.sqrt(torch.clamp(_vdot(x, x), min=0.0))


def _safe_normalize(x: torch.Tensor, thresh=None) -> Tuple[torch.Tensor, torch.Tensor]:   # :217-273 (strict_jax)
    norm = _norm(x)
    if thresh is None:
        thresh = torch.finfo(x.dtype).eps
    if not isinstance(thresh, torch.Tensor):
        thresh = torch.tensor(thresh, dtype=x.dtype)
    use_norm = norm > thresh
    normalized = torch.where(use_norm, x / norm.to(x.dtype), torch.zeros_like(x))
    norm = torch.where(use_norm, norm, torch.tensor(0.0, dtype=norm.dtype))
    return normalized, norm


# --------------------------------------------------------------------------------------------------
def _cg_solve(A, b, x0, maxiter, tol, atol, M=None):           # :806-856
    bs = _vdot(b, b)
    atol2 = torch.maximum(torch.square(torch.tensor(tol)) * bs, torch.square(torch.tensor(atol)))   # :815-817
    r = b - A(x0)                                              # :820
    p = z = r if M is None else M(r)                           # :821
    gamma = _vdot(r, z).to(r.dtype)                            # :826
    x, k = x0, 0
    while True:
        rs = gamma if M is None else _vdot(r, r)               # :835-838
        if k >= maxiter or rs <= atol2:                        # :841
            break
        Ap = A(p)                                              # :844
        alpha = gamma / _vdot(p, Ap).to(r.dtype)               # :845
        x = x + alpha * p                                      # :846
        r = r - alpha * Ap                                     # :847
        z = r if M is None else M(r)                           # :848
        gamma_new = _vdot(r, z).to(r.dtype)                    # :849-850
        beta = gamma_new / gamma                               # :851
        p = z + beta * p                                       # :852
        gamma = gamma_new
        k += 1
    return x, k


def _bicgstab_solve(A, b, x0, maxiter, tol, atol, M=None):     # :859-964
    bs = _vdot(b, b)
    atol2 = torch.maximum(torch.square(torch.tensor(tol)) * bs, torch.square(torch.tensor(atol)))   # :870-872
    r0 = b - A(x0)                                             # :875
    rhat = r0
    dtype = r0.dtype
    eps = torch.finfo(dtype).eps
    alpha = torch.tensor(1.0, dtype=dtype)
    omega = torch.tensor(1.0, dtype=dtype)
    rho = torch.tensor(1.0, dtype=dtype)
    x, r, p, q, k = x0, r0, r0, r0, 0                          # :890
    iters = 0
    while k < maxiter and k >= 0:                              # :892
        rs = _vdot(r, r)
        if rs <= atol2:                                        # :895
            break
        rho_new = _vdot(rhat, r)                               # :899
        if torch.abs(rho_new) < eps * torch.abs(rho):          # :902
            k = -10
            break
        beta = rho_new / rho * alpha / omega                   # :906
        p_ = r + beta * (p - omega * q)                        # :907
        phat = p_ if M is None else M(p_)                      # :908
        q_ = A(phat)                                           # :909
        alpha_new = rho_new / _vdot(rhat, q_)                  # :910
        if torch.abs(alpha_new) < eps:                         # :913
            k = -11
            break
        s = r - alpha_new * q_                                 # :917
        exit_early = _vdot(s, s) < atol2                       # :920
        shat = s if M is None else M(s)                        # :922
        t = A(shat)                                            # :923
        t_norm_sq = _vdot(t, t)
        if torch.abs(t_norm_sq) < eps:                         # :927
            omega_new = torch.tensor(0.0, dtype=dtype)
        else:
            omega_new = _vdot(t, s) / t_norm_sq                # :930
        if torch.abs(omega_new) < eps and not exit_early:      # :934
            k = -11
            break
        x_early = x + alpha_new * phat                         # :942
        x_full = x + (alpha_new * phat + omega_new * shat)     # :943
        x = torch.where(exit_early, x_early, x_full)
        r = torch.where(exit_early, s, s - omega_new * t)      # :948-950
        p, q, rho, alpha, omega = p_, q_, rho_new, alpha_new, omega_new
        k += 1
        iters = k
        if exit_early:                                         # :961
            break
    return x, (k if k >= 0 else iters), k


def _isolve(kind: str, A_t: torch.Tensor, b: torch.Tensor, x0, tol, atol, maxiter, M=None):   # :967-1016
    if x0 is None:
        x0 = torch.zeros_like(b)
    b = b.to(torch.float64)
    x0 = x0.to(torch.float64)
    if maxiter is None:
        maxiter = 10 * b.numel()
    if b.shape != x0.shape:
        raise ValueError(f'arrays in x0 and b must have matching shapes: {x0.shape} vs {b.shape}')
    A = _CountingMatvec(A_t)
    status = 0
    if kind == 'cg':
        x, iters = _cg_solve(A, b, x0, maxiter, tol, atol, M)
    else:
        x, iters, status = _bicgstab_solve(A, b, x0, maxiter, tol, atol, M)
    matvecs = A.calls
    final_residual = _norm(b - A(x)) if M is None else _norm(M(b - A(x)))   # :1008
    b_norm = _norm(b)
    atol_tensor = torch.maximum(torch.tensor(tol) * b_norm, torch.tensor(atol))   # :1010-1011
    failed = bool(torch.isnan(_norm(x))) or bool(final_residual > atol_tensor)
    info = -1 if failed else 0
    stats = dict(iterations=int(iters), matvecs=int(matvecs), final_residual=float(final_residual),
                 b_norm=float(b_norm), threshold=float(atol_tensor), status=int(status))
    return x, info, stats


def cg(A: torch.Tensor, b: torch.Tensor, x0: Optional[torch.Tensor] = None, *, tol=1e-5, atol=0.0, maxiter=None,
       M=None):
    """reference cg (:1019-1088) without the autograd wrapper.  Returns (x, info, stats).  M: optional preconditioner
    callable (the fixtures use Jacobi, `lambda r: r / d`)."""
    return _isolve('cg', A, b, x0, tol, atol, maxiter, M)


def bicgstab(A: torch.Tensor, b: torch.Tensor, x0: Optional[torch.Tensor] = None, *, tol=1e-5, atol=0.0,
             maxiter=None, M=None):
    """reference bicgstab (:1091-1158) without the autograd wrapper."""
    return _isolve('bicgstab', A, b, x0, tol, atol, maxiter, M)


# --------------------------------------------------------------------------------------------------
def _kth_arnoldi_iteration(k, A, V, H, M=None):                # :331-388
    eps = torch.finfo(V.dtype).eps
    v = A(V[..., k])                                           # :349-351
    if M is not None:
        v = M(v)                                               # :351  v = M(A(v))
    _, v_norm_0 = _safe_normalize(v)                           # :352
    # one classical Gram-Schmidt pass over ALL columns (:284-328; the second pass can never run)
    h = torch.einsum("...n,...->n", V, v)                      # :279
    v = v - torch.mv(V, h)                                     # :303-304
    r = torch.zeros(V.shape[-1], dtype=V.dtype) + h            # :305
    tol = eps * v_norm_0                                       # :358
    unit_v, v_norm_1 = _safe_normalize(v, thresh=tol)          # :359
    V = V.clone()
    V[..., k + 1] = unit_v                                     # :363-368
    H = H.clone()
    H[:k + 1, k] = r[:k + 1]                                   # :384
    H[k + 1, k] = v_norm_1                                     # :385
    return V, H, bool(v_norm_1 == 0.)                          # :387


def _lstsq(a, b):                                              # :391-428 ('normal_equations')
    b = b.unsqueeze(-1)
    a_t = a.conj().T
    a2 = torch.matmul(a_t, a)
    b2 = torch.matmul(a_t, b)
    try:
        L = torch.linalg.cholesky(a2)
        sol = torch.cholesky_solve(b2, L)
    except RuntimeError:
        sol = torch.linalg.solve(a2, b2)
    return sol.squeeze(-1)


def _givens_rotation(a, b):                                    # :508-518
    b_zero = torch.abs(b) == 0
    a_lt_b = torch.abs(a) < torch.abs(b)
    t = -torch.where(a_lt_b, a, b) / torch.where(a_lt_b, b, a)
    r = torch.rsqrt(1 + torch.abs(t) ** 2).to(t.dtype)
    cs = torch.where(b_zero, torch.tensor(1.0, dtype=t.dtype), torch.where(a_lt_b, r * t, r))
    sn = torch.where(b_zero, torch.tensor(0.0, dtype=t.dtype), torch.where(a_lt_b, r, r * t))
    return cs, sn


def _gmres_batched(A, b, x0, unit_residual, residual_norm, ptol, restart, M=None):     # :431-493
    dtype = b.dtype
    V = torch.cat([unit_residual.unsqueeze(-1), torch.zeros(unit_residual.shape + (restart,), dtype=dtype)], dim=-1)
    H = torch.zeros(restart + 1, restart, dtype=dtype)
    k, breakdown = 0, False
    while k < restart and not breakdown:                       # :463
        V, H, breakdown = _kth_arnoldi_iteration(k, A, V, H, M)
        k += 1
    beta_vec = torch.zeros(restart + 1, dtype=dtype)
    beta_vec[0] = residual_norm.to(dtype)
    y = _lstsq(H[:k + 1, :k], beta_vec[:k + 1]) if k > 0 else torch.zeros(0, dtype=dtype)   # :475-484
    x = x0 + torch.matmul(V[..., :k], y)                       # :488-490
    residual = b - A(x) if M is None else M(b - A(x))          # :491
    unit_residual, residual_norm = _safe_normalize(residual)
    return x, unit_residual, residual_norm


def _gmres_incremental(A, b, x0, unit_residual, residual_norm, ptol, restart, M=None):  # :557-638
    dtype = b.dtype
    V = torch.cat([unit_residual.unsqueeze(-1), torch.zeros(unit_residual.shape + (restart,), dtype=dtype)], dim=-1)
    H = torch.zeros(restart + 1, restart, dtype=dtype)
    R = torch.eye(restart, restart, dtype=dtype)
    givens = torch.zeros((restart, 2), dtype=dtype)
    beta_vec = torch.zeros(restart + 1, dtype=dtype)
    beta_vec[0] = residual_norm.to(dtype)
    k, err = 0, residual_norm
    while k < restart and err > ptol:                          # :591
        V, H, breakdown = _kth_arnoldi_iteration(k, A, V, H, M)
        H_col = H[:k + 2, k].clone()
        for i in range(k):                                     # :599-603
            cs, sn = givens[i, 0], givens[i, 1]
            temp = cs * H_col[i] - sn * H_col[i + 1]
            H_col[i + 1] = sn * H_col[i] + cs * H_col[i + 1]
            H_col[i] = temp
        cs_new, sn_new = _givens_rotation(H_col[k], H_col[k + 1])   # :606
        givens[k, 0], givens[k, 1] = cs_new, sn_new
        H_col[k] = cs_new * H_col[k] - sn_new * H_col[k + 1]   # :611
        H_col[k + 1] = 0.0
        R[:k + 1, k] = H_col[:k + 1]                           # :615
        temp = cs_new * beta_vec[k] - sn_new * beta_vec[k + 1]  # :618-620
        beta_vec[k + 1] = sn_new * beta_vec[k] + cs_new * beta_vec[k + 1]
        beta_vec[k] = temp
        err = torch.abs(beta_vec[k + 1])
        k += 1
        if breakdown:
            break
    if k > 0:
        y = torch.linalg.solve_triangular(R[:k, :k], beta_vec[:k].unsqueeze(-1), upper=True).squeeze(-1)   # :630
        dx = torch.matmul(V[..., :k], y)
    else:
        dx = torch.zeros_like(x0)
    x = x0 + dx
    residual = b - A(x) if M is None else M(b - A(x))          # :636
    unit_residual, residual_norm = _safe_normalize(residual)
    return x, unit_residual, residual_norm


def gmres_tolerances(tol: float, atol: float, size: int, b_norm: torch.Tensor, device_type: str = 'cpu'):
    """:733-753 — returns (atol_tensor, ptol) with the reference's fp32 roundings and device-dependent constants."""
    dtype = torch.float64
    if device_type == 'cuda':
        adaptive_tol = max(tol, 1e-12 * torch.sqrt(torch.tensor(size, dtype=torch.float64)))
        base_atol = torch.finfo(dtype).eps * 1000 * size
    else:
        adaptive_tol = max(tol, 1e-14 * torch.sqrt(torch.tensor(size, dtype=torch.float64)))
        base_atol = torch.finfo(dtype).eps * 100 * size
    atol_tensor = torch.maximum(torch.tensor(adaptive_tol) * b_norm,
                                torch.maximum(torch.tensor(atol), torch.tensor(base_atol)))
    ptol = b_norm * torch.minimum(torch.tensor(1.0), atol_tensor / b_norm)   # M = identity => ||Mb|| = ||b||
    return atol_tensor, ptol


def gmres(A_t: torch.Tensor, b: torch.Tensor, x0: Optional[torch.Tensor] = None, *, tol=1e-5, atol=0.0, restart=20,
          maxiter=None, solve_method='batched', device_type='cpu', M=None):
    """reference gmres (:641-784) + _gmres_solve_with_method (:788-803) without the autograd wrapper."""
    if x0 is None:
        x0 = torch.zeros_like(b)
    b = b.to(torch.float64)
    x0 = x0.to(torch.float64)
    if maxiter is None:
        maxiter = 10 * b.numel()
    A = _CountingMatvec(A_t)
    b_norm = _norm(b)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        atol_tensor, ptol = gmres_tolerances(tol, atol, b.numel(), b_norm, device_type)
    if M is not None:                                          # :750-753  ptol uses ||M b||
        ptol = _norm(M(b)) * torch.minimum(torch.tensor(1.0), atol_tensor / b_norm)
    if solve_method == 'incremental':
        cycle = _gmres_incremental
    elif solve_method == 'batched':
        cycle = _gmres_batched
    else:
        raise ValueError(f"Unsupported solve_method: {solve_method}")
    residual = b - A(x0) if M is None else M(b - A(x0))        # :791
    unit_residual, residual_norm = _safe_normalize(residual)
    k, x = 0, x0
    while k < maxiter and residual_norm > atol_tensor:         # :798
        x, unit_residual, residual_norm = cycle(A, b, x, unit_residual, residual_norm, ptol, restart, M)
        k += 1
    matvecs = A.calls
    final_residual = _norm(b - A(x)) if M is None else _norm(M(b - A(x)))   # :766
    failed = bool(torch.isnan(_norm(x))) or bool(final_residual > atol_tensor * 10)   # :769-770
    info = -1 if failed else 0
    stats = dict(iterations=int(k), matvecs=int(matvecs), final_residual=float(final_residual),
                 b_norm=float(b_norm), threshold=float(atol_tensor * 10), status=0)
    return x, info, stats


# --------------------------------------------------------------------------------------------------
def adjoint_grad_b(kind: str, A: torch.Tensor, grad_x: torch.Tensor, x0, **kw) -> torch.Tensor:
    """ImplicitAdjointFunction.backward (:1237-1248): grad_b = solve(A^T, grad_x, x0, same tolerances).
    A^T is materialised densely/COO here (the reference's `A.T` raises for CSR on torch 2.11)."""
    if A.layout == torch.sparse_csr:
        At = A.to_sparse_coo().t().coalesce().to_sparse_csr()
    elif A.layout == torch.sparse_coo:
        At = A.t().coalesce()
    else:
        At = A.T
    fn = {'cg': cg, 'bicgstab': bicgstab, 'gmres': gmres}[kind]
    g, _info, _stats = fn(At, grad_x, x0, **kw)
    return g
