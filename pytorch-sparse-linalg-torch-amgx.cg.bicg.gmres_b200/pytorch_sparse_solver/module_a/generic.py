"""Generic route of module_a: callable `A`, preconditioner `M`, pytree `b`.

The reference's recurrences (torch_sparse_linalg.py: _cg_solve :806-856, _bicgstab_solve :859-964,
_gmres_batched :431-493, _gmres_incremental :557-638, _kth_arnoldi_iteration :331-388) driven from Python, but every
dot / norm / axpy / division on the vectors is done by the library's deterministic CUDA kernels (bk_dot, bk_axpby,
bk_axpby_dev, bk_div_scalar) and a matrix given as a TENSOR (A with a callable M, or M itself) is applied by the native
SpMV of its registration — never by torch.matmul / cuSPARSE and never by a CPU loop.  The user's callables are called
as they are.  CPU pytrees are staged through the current CUDA device (callables are then fed CPU tensors).

Host synchronisations: CG keeps alpha / beta on the device (bk_axpby_dev reads them from device memory) and syncs
once per iteration for the stop test — exactly the reference's one `bool()` per iteration (:841); BiCGStab syncs four
times per iteration (reference: five, :895-936), GMRES three times per Arnoldi step (reference: at least three).

`transpose=True` solves with A^T (tensor A only): the adjoint solve of the implicit-differentiation backward
(reference :1237-1248), served by the cached device transpose.
"""
from __future__ import annotations

import math
from typing import Any, Callable, List, Optional, Tuple

import torch

from .. import _native
from .torch_tree_util import tree_flatten, tree_leaves, tree_unflatten

_EPS = torch.finfo(torch.float64).eps


class _Space:
    """Flatten/unflatten between the user's pytree and a list of contiguous fp64 CUDA leaves."""

    def __init__(self, b: Any):
        leaves, self.treedef = tree_flatten(b)
        if not leaves:
            raise ValueError("b must contain at least one tensor")
        self.shapes = [tuple(t.shape) for t in leaves]
        self.on_cpu = not leaves[0].is_cuda
        if self.on_cpu and not torch.cuda.is_available():
            _native.load_library()
            raise _native.NativeLibraryError("a CUDA device is required: module_a has no CPU fallback in this build")
        self.device = torch.device("cuda", torch.cuda.current_device()) if self.on_cpu else leaves[0].device
        self.size = sum(t.numel() for t in leaves)

    def to_vec(self, tree: Any) -> List[torch.Tensor]:
        return [t.detach().to(device=self.device, dtype=torch.float64).reshape(-1).contiguous()
                for t in tree_leaves(tree)]

    def to_tree(self, vec: List[torch.Tensor]) -> Any:
        out = [v.reshape(s) for v, s in zip(vec, self.shapes)]
        if self.on_cpu:
            out = [t.cpu() for t in out]
        return tree_unflatten(self.treedef, out)

    def wrap(self, fn: Optional[Callable], transpose: bool = False) -> Callable[[List[torch.Tensor]], List[torch.Tensor]]:
        """User callable pytree -> pytree  ==>  leaf-list -> leaf-list on the work device."""
        if fn is None:
            return lambda v: v
        if isinstance(fn, torch.Tensor):  # a matrix (A next to a callable M, or M given as a matrix): native SpMV
            mat = fn.detach()
            if mat.is_complex():
                raise NotImplementedError("complex operators take the complex route (module_a/complex_route.py)")
            if not mat.is_cuda:
                mat = mat.to(self.device)
            reg = _native.register_matrix(mat, torch.float64)
            if transpose:
                reg = reg.transpose()

            def mv(v):
                flat = v[0] if len(v) == 1 else torch.cat(v)
                y = reg.spmv(flat)
                if len(v) == 1:
                    return [y]
                outs, o = [], 0
                for t in v:
                    outs.append(y[o:o + t.numel()].contiguous())
                    o += t.numel()
                return outs
            return mv
        if transpose:
            raise ValueError("the transposed solve needs A as a tensor")

        def call(v):
            y = fn(self.to_tree(v))
            return self.to_vec(y)
        return call


def _dotd(x: List[torch.Tensor], y: List[torch.Tensor]) -> torch.Tensor:
    """x . y as a 0-dim fp64 DEVICE tensor (no host sync)."""
    acc = None
    for a, b in zip(x, y):
        d = _native.dot(a, b)
        acc = d if acc is None else acc + d
    return acc


def _dot(x: List[torch.Tensor], y: List[torch.Tensor]) -> float:
    return float(_dotd(x, y))


def _dots(pairs) -> List[float]:
    """Several dots, one host sync."""
    vals = [_dotd(x, y) for x, y in pairs]
    return [float(v) for v in torch.stack(vals).cpu()]


def _axpby(a: float, x, b: float, y):
    return [_native.axpby(a, xi, b, yi) for xi, yi in zip(x, y)]


def _axpby_dev(sa: float, a, x, sb: float, b, y):
    """(sa * a) x + (sb * b) y with a, b 0-dim device tensors or None (= 1)."""
    return [_native.axpby_dev(sa, a, xi, sb, b, yi) for xi, yi in zip(x, y)]


def _div(x, d: float):
    return [_native.div_scalar(t, d) for t in x]


def _norm(x) -> float:
    return math.sqrt(max(_dot(x, x), 0.0))


def _f32(v: float) -> float:
    return float(torch.tensor(v))  # the reference's torch.tensor(tol) is fp32 (:816)


def _prepare(A, b, x0, M, transpose: bool = False):
    sp = _Space(b)
    bv = sp.to_vec(b)
    if x0 is None:
        xv = [torch.zeros_like(t) for t in bv]
    else:
        b_leaves, x_leaves = tree_leaves(b), tree_leaves(x0)
        if len(b_leaves) != len(x_leaves):
            raise ValueError('x0 and b must have matching tree structure')
        for bl, xl in zip(b_leaves, x_leaves):
            if bl.shape != xl.shape:
                raise ValueError(f'arrays in x0 and b must have matching shapes: {xl.shape} vs {bl.shape}')
        xv = sp.to_vec(x0)
    Aop = sp.wrap(A, transpose)
    Mop = sp.wrap(M)
    return sp, bv, xv, Aop, Mop


def _finish_isolve(sp, Aop, Mop, bv, xv, tol, atol, its, name):
    from . import krylov
    res = Mop(_axpby(1.0, bv, -1.0, Aop(xv)))
    final, bn, xn = _norm(res), _norm(bv), _norm(xv)
    thr = max(_f32(tol) * bn, _f32(atol))
    failed = math.isnan(xn) or final > thr
    krylov.last_result = dict(iterations=its, matvecs=None, info=-1 if failed else 0, final_residual=final,
                              threshold=thr, b_norm=bn, x_norm=xn, solver=name, route="generic")
    return sp.to_tree(xv), (-1 if failed else 0)


def generic_cg(A, b, x0=None, *, tol=1e-5, atol=0.0, maxiter=None, M=None, transpose=False):
    sp, bv, x, Aop, Mop = _prepare(A, b, x0, M, transpose)
    maxiter = 10 * sp.size if maxiter is None else maxiter
    bs = _dot(bv, bv)
    t32, a32 = torch.tensor(tol), torch.tensor(atol)
    atol2 = max(float(torch.square(t32)) * bs, float(torch.square(a32)))
    r = _axpby(1.0, bv, -1.0, Aop(x))
    z = Mop(r)
    p = z
    gamma = _dotd(r, z)                      # device scalar
    k = 0
    while True:
        rs = gamma if M is None else _dotd(r, r)
        if k >= maxiter or float(rs) <= atol2:   # the iteration's one host sync (reference :841)
            break
        Ap = Aop(p)
        alpha = gamma / _dotd(p, Ap)             # 0-dim device tensors: IEEE fp64 division, as the reference's
        x = _axpby_dev(1.0, None, x, 1.0, alpha, p)
        r = _axpby_dev(1.0, None, r, -1.0, alpha, Ap)
        z = Mop(r)
        gamma_new = _dotd(r, z)
        beta = gamma_new / gamma
        p = _axpby_dev(1.0, None, z, 1.0, beta, p)
        gamma = gamma_new
        k += 1
    return _finish_isolve(sp, Aop, Mop, bv, x, tol, atol, k, "cg")


def generic_bicgstab(A, b, x0=None, *, tol=1e-5, atol=0.0, maxiter=None, M=None, transpose=False):
    sp, bv, x, Aop, Mop = _prepare(A, b, x0, M, transpose)
    maxiter = 10 * sp.size if maxiter is None else maxiter
    bs = _dot(bv, bv)
    atol2 = max(float(torch.square(torch.tensor(tol))) * bs, float(torch.square(torch.tensor(atol))))
    r = _axpby(1.0, bv, -1.0, Aop(x))
    rhat = [t.clone() for t in r]
    alpha = omega = rho = 1.0
    p = [t.clone() for t in r]
    q = [t.clone() for t in r]
    k = 0
    its = 0
    while k < maxiter and k >= 0:
        rs, rho_new = _dots([(r, r), (rhat, r)])
        if rs <= atol2:
            break
        if abs(rho_new) < _EPS * abs(rho):
            k = -10
            break
        beta = rho_new / rho * alpha / omega
        p_ = _axpby(1.0, r, beta, _axpby(1.0, p, -omega, q))
        phat = Mop(p_)
        q_ = Aop(phat)
        alpha_new = rho_new / _dot(rhat, q_)
        if abs(alpha_new) < _EPS:
            k = -11
            break
        s = _axpby(1.0, r, -alpha_new, q_)
        exit_early = _dot(s, s) < atol2
        shat = Mop(s)
        t = Aop(shat)
        ts, tt = _dots([(t, s), (t, t)])
        omega_new = 0.0 if abs(tt) < _EPS else ts / tt
        if abs(omega_new) < _EPS and not exit_early:
            k = -11
            break
        if exit_early:
            x = _axpby(1.0, x, alpha_new, phat)
            r = s
        else:
            x = _axpby(1.0, x, 1.0, _axpby(alpha_new, phat, omega_new, shat))
            r = _axpby(1.0, s, -omega_new, t)
        p, q, rho, alpha, omega = p_, q_, rho_new, alpha_new, omega_new
        k += 1
        its = k
        if exit_early:
            break
    return _finish_isolve(sp, Aop, Mop, bv, x, tol, atol, its, "bicgstab")


def _safe_normalize(v, thresh: float = _EPS):
    n = _norm(v)
    if n > thresh:
        return _div(v, n), n                 # true division, as the reference's y / norm (:268)
    return [torch.zeros_like(t) for t in v], 0.0


def _givens(a: float, b: float) -> Tuple[float, float]:
    if abs(b) == 0.0:
        return 1.0, 0.0
    if abs(a) < abs(b):
        t = -a / b
        r = 1.0 / math.sqrt(1.0 + t * t)
        return r * t, r
    t = -b / a
    r = 1.0 / math.sqrt(1.0 + t * t)
    return r, r * t


def _gmres_cycle(Aop, Mop, bv, x, v0, beta, ptol, restart, incremental):
    V = [v0]
    R = [[0.0] * restart for _ in range(restart + 1)]
    cs, sn = [0.0] * restart, [0.0] * restart
    g = [0.0] * (restart + 1)
    g[0] = beta
    k, err = 0, beta
    while k < restart and (not incremental or err > ptol):
        w = Mop(Aop(V[k]))
        vnorm0 = _norm(w)
        vnorm0 = vnorm0 if vnorm0 > _EPS else 0.0
        h = _dots([(V[j], w) for j in range(k + 1)])
        # w -= V h   (sum first, then subtract — reference :303-304)
        qh = [torch.zeros_like(t) for t in w]
        for j in range(k + 1):
            qh = _axpby(1.0, qh, h[j], V[j])
        w = _axpby(1.0, w, -1.0, qh)
        norm1 = _norm(w)
        use = norm1 > _EPS * vnorm0
        vnorm1 = norm1 if use else 0.0
        V.append(_div(w, norm1) if use else [torch.zeros_like(t) for t in w])
        col = h + [vnorm1]
        for i in range(k):
            tmp = cs[i] * col[i] - sn[i] * col[i + 1]
            col[i + 1] = sn[i] * col[i] + cs[i] * col[i + 1]
            col[i] = tmp
        cs[k], sn[k] = _givens(col[k], col[k + 1])
        col[k] = cs[k] * col[k] - sn[k] * col[k + 1]
        col[k + 1] = 0.0
        for i in range(k + 1):
            R[i][k] = col[i]
        tmp = cs[k] * g[k] - sn[k] * g[k + 1]
        g[k + 1] = sn[k] * g[k] + cs[k] * g[k + 1]
        g[k] = tmp
        err = abs(g[k + 1])
        k += 1
        if vnorm1 == 0.0:
            break
    y = [0.0] * k
    for i in range(k - 1, -1, -1):
        s = g[i] - sum(R[i][c] * y[c] for c in range(i + 1, k))
        y[i] = s / R[i][i]
    dx = [torch.zeros_like(t) for t in x]
    for j in range(k):
        dx = _axpby(1.0, dx, y[j], V[j])
    x = _axpby(1.0, x, 1.0, dx)
    res = Mop(_axpby(1.0, bv, -1.0, Aop(x)))
    v0, beta = _safe_normalize(res)
    return x, v0, beta


def generic_gmres(A, b, x0=None, *, tol=1e-5, atol=0.0, restart=20, maxiter=None, M=None, solve_method='batched',
                  transpose=False):
    from . import krylov
    sp, bv, x, Aop, Mop = _prepare(A, b, x0, M, transpose)
    maxiter = 10 * sp.size if maxiter is None else maxiter
    bn = _norm(bv)
    dev = 'cpu' if sp.on_cpu else 'cuda'
    tol_eff, atol_eff = krylov._gmres_effective_tolerances(tol, atol, sp.size, dev)
    atol_t = max(tol_eff * bn, atol_eff)
    mbn = _norm(Mop(bv))
    ptol = mbn * min(1.0, atol_t / bn) if bn > 0 else 0.0
    res = Mop(_axpby(1.0, bv, -1.0, Aop(x)))
    v0, beta = _safe_normalize(res)
    k = 0
    while k < maxiter and beta > atol_t:
        x, v0, beta = _gmres_cycle(Aop, Mop, bv, x, v0, beta, ptol, restart, solve_method == 'incremental')
        k += 1
    final = _norm(Mop(_axpby(1.0, bv, -1.0, Aop(x))))
    xn = _norm(x)
    failed = math.isnan(xn) or final > 10.0 * atol_t
    krylov.last_result = dict(iterations=k, matvecs=None, info=-1 if failed else 0, final_residual=final,
                              threshold=10.0 * atol_t, b_norm=bn, x_norm=xn, solver="gmres", route="generic")
    return sp.to_tree(x), (-1 if failed else 0)
