"""Complex systems (reference torch_sparse_linalg.py:86-127 `_vdot` / `_vdot_real_part`, :1220 the
conjugate transpose of the adjoint solve).

The reference computes in complex128 with torch ops.  Here a complex matrix is registered ONCE as its real-equivalent
2n x 2n CSR matrix — every entry a + ib becomes the block [[a, -b], [b, a]] acting on interleaved (re, im) vectors, which
is exactly the memory layout of `torch.view_as_real` — so the native SpMV kernels serve complex matvecs unchanged, and
complex vectors are handled through their fp64 views by the library's bk_cdot / bk_caxpby kernels:

  cg        for hermitian A every scalar of the reference's recurrence is REAL (`_vdot_real_part`, :100-127): complex CG
            IS real CG on the real-equivalent system, so the solve runs on the native device-resident loop (bk_cg).
  bicgstab  rho, alpha, omega are complex (`_vdot_tree`, :898, :910, :926-930): the reference's recurrence driven from
  gmres     Python with complex scalars on the host, vector work in bk_spmv / bk_cdot / bk_caxpby; GMRES's small
            least-squares problem (<= (restart+1) x restart complex) is solved on the host with numpy.

Only M=None and tensor A are supported (dense / COO / CSR, CPU or CUDA); results are complex128 on b's device.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch

from .. import _native

_EPS = float(torch.finfo(torch.float64).eps)


def real_equivalent_csr(A: torch.Tensor, conj_transpose: bool = False) -> torch.Tensor:
    """Real 2n x 2n CSR tensor of the complex n x n tensor A (dense / COO / CSR), interleaved (re, im) ordering."""
    A = A.detach()
    if conj_transpose:
        A = (A.to_dense() if A.layout != torch.strided else A).conj().T.contiguous()
    if A.layout == torch.sparse_coo:
        A = A.coalesce().to_sparse_csr()
    elif A.layout != torch.sparse_csr:
        A = A.to(torch.complex128).to_sparse_csr()
    crow, col, val = A.crow_indices().long(), A.col_indices().long(), A.values().to(torch.complex128)
    dev = val.device
    n = A.shape[0]
    lens = crow[1:] - crow[:-1]
    nnz = val.numel()
    row = torch.repeat_interleave(torch.arange(n, device=dev), lens)
    pos = torch.arange(nnz, device=dev) - crow[row]                 # position of the entry inside its row
    base = 4 * crow[row]                                            # first real entry of complex row `row`
    vr = torch.view_as_real(val.clone())                            # (.real on a CSR values view raises in torch 2.11)
    ar, ai = vr[:, 0].contiguous(), vr[:, 1].contiguous()
    rcol = torch.empty(4 * nnz, dtype=torch.int64, device=dev)
    rval = torch.empty(4 * nnz, dtype=torch.float64, device=dev)
    top = base + 2 * pos                                            # real row 2i:   [ re  -im ]
    bot = base + 2 * lens[row] + 2 * pos                            # real row 2i+1: [ im   re ]
    rcol[top], rcol[top + 1] = 2 * col, 2 * col + 1
    rval[top], rval[top + 1] = ar, -ai
    rcol[bot], rcol[bot + 1] = 2 * col, 2 * col + 1
    rval[bot], rval[bot + 1] = ai, ar
    rcrow = torch.zeros(2 * n + 1, dtype=torch.int64, device=dev)
    rcrow[1:] = torch.repeat_interleave(2 * lens, 2).cumsum(0)
    return torch.sparse_csr_tensor(rcrow, rcol, rval, size=(2 * n, 2 * n))


def _rview(z: torch.Tensor) -> torch.Tensor:
    """complex128 vector -> its interleaved fp64 view (no copy for contiguous input)."""
    return torch.view_as_real(z.contiguous()).reshape(-1)


def _cview(r: torch.Tensor) -> torch.Tensor:
    return torch.view_as_complex(r.reshape(-1, 2))


_REG = {}


def _registered(A: torch.Tensor, device, conj_transpose: bool):
    """Real-equivalent registration of A on `device`, cached per source storage (validated like every registration)."""
    src = A.detach()
    key = (src.data_ptr() if src.layout == torch.strided else src.values().data_ptr() if src.layout == torch.sparse_csr
           else src._values().data_ptr(), tuple(src.shape), str(src.layout), conj_transpose, str(device))
    hit = _REG.get(key)
    Ar = None
    if hit is None or hit[0]() is None:
        Ar = real_equivalent_csr(src, conj_transpose).to(device)
        import weakref
        _REG[key] = (weakref.ref(A), Ar)
        while len(_REG) > 4:
            _REG.pop(next(iter(_REG)))
    else:
        Ar = hit[1]
    return Ar, _native.register_matrix(Ar, torch.float64)


def _f32(v: float) -> float:
    return float(torch.tensor(v))


def complex_solve(name: str, A: torch.Tensor, b: torch.Tensor, x0: Optional[torch.Tensor] = None, *, tol=1e-5,
                  atol=0.0, maxiter=None, M=None, restart=20, solve_method='batched', _result=None,
                  conj_transpose: bool = False) -> Tuple[torch.Tensor, int]:
    from . import krylov
    if M is not None:
        raise NotImplementedError("complex systems: preconditioners are not supported by this build")
    if not isinstance(b, torch.Tensor) or b.ndim != 1 or b.shape[0] != A.shape[0]:
        raise ValueError(f"b must be a vector of length {A.shape[0]}")
    if x0 is not None and tuple(x0.shape) != tuple(b.shape):
        raise ValueError(f'arrays in x0 and b must have matching shapes: {x0.shape} vs {b.shape}')
    on_cpu = not b.is_cuda
    if on_cpu and not torch.cuda.is_available():
        _native.load_library()
        raise _native.NativeLibraryError("a CUDA device is required: module_a has no CPU fallback in this build")
    dev = torch.device("cuda", torch.cuda.current_device()) if on_cpu else b.device
    n = b.shape[0]
    maxiter = 10 * n if maxiter is None else maxiter
    with torch.no_grad():
        Ar, reg = _registered(A, dev, conj_transpose)
        bz = b.detach().to(device=dev, dtype=torch.complex128)
        xz = None if x0 is None else x0.detach().to(device=dev, dtype=torch.complex128)
        br = _rview(bz)
        if name == "cg":
            # hermitian A: every scalar of the recurrence is real => native real CG on the real-equivalent system
            xr, res = reg.cg(br, None if xz is None else _rview(xz), tol, atol, maxiter)
            x, info = _cview(xr), int(res["info"])
            rec = dict(res, solver="cg", route="complex")
        elif name == "bicgstab":
            x, info, rec = _bicgstab(reg, br, None if xz is None else _rview(xz), tol, atol, maxiter)
        else:
            if restart < 1:
                raise ValueError("restart must be >= 1")
            x, info, rec = _gmres(reg, br, None if xz is None else _rview(xz), tol, atol, restart, maxiter,
                                  solve_method, 'cpu' if on_cpu else 'cuda', n)
    krylov._publish(rec, _result)
    x = x.clone()
    if on_cpu:
        x = x.cpu()
    if b.requires_grad:
        x = _ComplexAdjoint.apply(b, x, A, name, x0, tol, atol, maxiter, restart, solve_method)
    return x, info


class _ComplexAdjoint(torch.autograd.Function):
    """grad_b = solve(A^H, grad_x) — the reference's adjoint for complex A (:1220 `A.T.conj()`)."""

    @staticmethod
    def forward(ctx, b, x, A, name, x0, tol, atol, maxiter, restart, solve_method):
        ctx.A = A
        ctx.meta = (name, x0, tol, atol, maxiter, restart, solve_method)
        return x.clone()

    @staticmethod
    def backward(ctx, grad_output):
        name, x0, tol, atol, maxiter, restart, solve_method = ctx.meta
        from . import krylov
        saved = krylov.last_result
        try:
            g, _ = complex_solve(name, ctx.A, grad_output.detach().contiguous(), x0, tol=tol, atol=atol,
                                 maxiter=maxiter, restart=restart, solve_method=solve_method, conj_transpose=True)
        finally:
            krylov.last_result = saved
        return (g.to(grad_output.dtype),) + (None,) * 9


# ---- vector helpers on interleaved views -------------------------------------------------------------------------------
def _cd(x, y) -> complex:
    v = _native.cdot(x, y).cpu()
    return complex(float(v[0]), float(v[1]))


def _nrm(x) -> float:
    return math.sqrt(max(float(_native.dot(x, x)), 0.0))


def _bicgstab(reg, b, x0, tol, atol, maxiter):
    """reference _bicgstab_solve (:859-964) with complex rho / alpha / omega; _isolve's final check (:1008-1016)."""
    x = torch.zeros_like(b) if x0 is None else x0.clone()
    bs = float(_native.dot(b, b))
    atol2 = max(float(torch.square(torch.tensor(tol))) * bs, float(torch.square(torch.tensor(atol))))
    r = _native.caxpby(1.0, b, -1.0, reg.spmv(x))
    rhat = r.clone()
    alpha = omega = rho = complex(1.0)
    p, q = r.clone(), r.clone()
    k, its, matvecs = 0, 0, 1
    while k < maxiter and k >= 0:
        rs = float(_native.dot(r, r))
        if rs <= atol2:
            break
        rho_new = _cd(rhat, r)
        if abs(rho_new) < _EPS * abs(rho):
            k = -10
            break
        beta = rho_new / rho * alpha / omega
        p_ = _native.caxpby(1.0, r, beta, _native.caxpby(1.0, p, -omega, q))
        q_ = reg.spmv(p_)
        alpha_new = rho_new / _cd(rhat, q_)
        if abs(alpha_new) < _EPS:
            k = -11
            break
        s = _native.caxpby(1.0, r, -alpha_new, q_)
        exit_early = float(_native.dot(s, s)) < atol2
        t = reg.spmv(s)
        matvecs += 2
        tt = _cd(t, t)
        omega_new = complex(0.0) if abs(tt) < _EPS else _cd(t, s) / tt
        if abs(omega_new) < _EPS and not exit_early:
            k = -11
            break
        if exit_early:
            x = _native.caxpby(1.0, x, alpha_new, p_)
            r = s
        else:
            x = _native.caxpby(1.0, x, 1.0, _native.caxpby(alpha_new, p_, omega_new, s))
            r = _native.caxpby(1.0, s, -omega_new, t)
        p, q, rho, alpha, omega = p_, q_, rho_new, alpha_new, omega_new
        k += 1
        its = k
        if exit_early:
            break
    final = _nrm(_native.caxpby(1.0, b, -1.0, reg.spmv(x)))
    bn, xn = math.sqrt(max(bs, 0.0)), _nrm(x)
    thr = max(_f32(tol) * bn, _f32(atol))
    failed = math.isnan(xn) or final > thr
    rec = dict(iterations=its, matvecs=matvecs, info=-1 if failed else 0, final_residual=final, threshold=thr, b_norm=bn,
               x_norm=xn, solver="bicgstab", route="complex")
    return _cview(x), (-1 if failed else 0), rec


def _gmres(reg, b, x0, tol, atol, restart, maxiter, solve_method, devtype, n):
    """reference gmres (:641-784): Arnoldi with complex projection coefficients (`_project_on_columns` conjugates the
    basis, :276-281), one classical Gram-Schmidt pass, least squares of the (k+1) x k Hessenberg matrix on the host."""
    from . import krylov
    x = torch.zeros_like(b) if x0 is None else x0.clone()
    bn = _nrm(b)
    tol_eff, atol_eff = krylov._gmres_effective_tolerances(tol, atol, n, devtype)
    atol_t = max(tol_eff * bn, atol_eff)
    ptol = bn * min(1.0, atol_t / bn) if bn > 0 else 0.0
    incremental = solve_method == 'incremental'

    def normalize(v, thresh=_EPS):
        nv = _nrm(v)
        if nv > thresh:
            return _native.div_scalar(v, nv), nv
        return torch.zeros_like(v), 0.0

    res = _native.caxpby(1.0, b, -1.0, reg.spmv(x))
    v0, beta = normalize(res)
    cycles, matvecs = 0, 1
    while cycles < maxiter and beta > atol_t:
        V = [v0]
        H = np.zeros((restart + 1, restart), dtype=np.complex128)
        k, err = 0, beta
        while k < restart and (not incremental or err > ptol):
            w = reg.spmv(V[k])
            matvecs += 1
            vn0 = _nrm(w)
            vn0 = vn0 if vn0 > _EPS else 0.0
            h = [_cd(V[j], w) for j in range(k + 1)]            # conj(v_j) . w
            qh = torch.zeros_like(w)
            for j in range(k + 1):
                qh = _native.caxpby(1.0, qh, h[j], V[j])
            w = _native.caxpby(1.0, w, -1.0, qh)
            n1 = _nrm(w)
            use = n1 > _EPS * vn0
            vn1 = n1 if use else 0.0
            V.append(_native.div_scalar(w, n1) if use else torch.zeros_like(w))
            H[:k + 1, k] = h
            H[k + 1, k] = vn1
            k += 1
            if incremental:
                rhs = np.zeros(k + 1, dtype=np.complex128)
                rhs[0] = beta
                y_, *_ = np.linalg.lstsq(H[:k + 1, :k], rhs, rcond=None)
                err = float(np.linalg.norm(H[:k + 1, :k] @ y_ - rhs))
            if vn1 == 0.0:
                break
        if k > 0:
            rhs = np.zeros(k + 1, dtype=np.complex128)
            rhs[0] = beta
            y, *_ = np.linalg.lstsq(H[:k + 1, :k], rhs, rcond=None)
            dx = torch.zeros_like(x)
            for j in range(k):
                dx = _native.caxpby(1.0, dx, complex(y[j]), V[j])
            x = _native.caxpby(1.0, x, 1.0, dx)
        res = _native.caxpby(1.0, b, -1.0, reg.spmv(x))
        matvecs += 1
        v0, beta = normalize(res)
        cycles += 1
    final = _nrm(_native.caxpby(1.0, b, -1.0, reg.spmv(x)))
    xn = _nrm(x)
    failed = math.isnan(xn) or final > 10.0 * atol_t
    rec = dict(iterations=cycles, matvecs=matvecs, info=-1 if failed else 0, final_residual=final,
               threshold=10.0 * atol_t, b_norm=bn, x_norm=xn, solver="gmres", route="complex")
    return _cview(x), (-1 if failed else 0), rec
