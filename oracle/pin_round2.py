#!/usr/bin/env python3
"""Round-2 additions to the golden fixtures (test infrastructure; build container only, needs /root/reference).

    python oracle/pin_round2.py [--full]

Runs the UNMODIFIED reference (pytorch_sparse_solver.module_a imported from /root/reference/src, CPU) and the oracle
next to it, and adds to tests/golden/:
  * autograd_<kind>_jacobi.npz   grad_b with a callable M = lambda r: r / d (reference: ImplicitAdjointFunction is
                                 attached whenever A is a 2-D tensor, also with M — torch_sparse_linalg.py:1079-1086
                                 cg, :1145-1152 bicgstab, :775-782 gmres; the adjoint solve reuses M)
  * autograd_gmres_ldc100.npz    BASELINE configs[3] at the LDC default size nx=100 (ldc_solver_common.py:35), step 1
                                 RHS of the reference driver, GMRES(30) tol 1e-10 + backward of sum(x^2)
  * complex_*.npz                complex128 systems (reference :100-127 _vdot_real_part, :1220 conj transpose)
  * --full: digest_bicgstab_cd3d256_tol1e-10.npz: BiCGStab on CD3D-256 at tol 1e-10 (SURVEY §8c parity protocol:
    x gate at tol 1e-10, reference self-noise 4e-13) — iterations, ||x||, 4096 strided samples of x (~10 CPU-minutes)
and merges the entries into tests/golden/manifest.json under "round2".
"""
import json
import os
import sys
import time
import warnings
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT / "oracle"))
warnings.filterwarnings("ignore")

from pin_reference import load_reference, PKG_DIR  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count())
    ref, _ref_mods = load_reference()
    sys.path.insert(0, str(PKG_DIR))
    sys.path.insert(0, str(ROOT))
    from pytorch_sparse_solver import problems
    from oracle import krylov_oracle as orc
    with open(GOLD / "manifest.json") as f:
        manifest = json.load(f)
    r2 = manifest.setdefault("round2", {})

    # ---- implicit-diff backward with a callable preconditioner -----------------------------------------------
    for kind, A, kw in (("cg", problems.scaled_poisson3d_csr(6), dict(tol=1e-12)),
                        ("bicgstab", problems.scaled_convdiff3d_csr(6), dict(tol=1e-12)),
                        ("gmres", problems.scaled_convdiff3d_csr(6), dict(tol=1e-12, restart=30))):
        Ad = A.to_dense()
        d = problems.csr_diagonal(A)
        b0, _ = problems.manufactured_rhs(A, 5)
        b1 = b0.clone().requires_grad_(True)
        x, info = getattr(ref, kind)(Ad, b1, M=lambda r: r / d, **kw)
        (x ** 2).sum().backward()
        g_ref = b1.grad.clone()
        g_exact = torch.linalg.solve(Ad.T, 2.0 * torch.linalg.solve(Ad, b0))
        err = float((g_ref - g_exact).abs().max() / g_exact.abs().max())
        name = f"autograd_{kind}_jacobi"
        np.savez_compressed(GOLD / f"{name}.npz", b=b0.numpy(), grad_b=g_ref.numpy(), x=x.detach().numpy(), d=d.numpy())
        r2[name] = dict(kind=kind, info=int(info), kwargs=kw, n=int(b0.numel()),
                        gen=dict(matrix="scaled_poisson3d" if kind == "cg" else "scaled_convdiff3d", n=6, seed=7),
                        grad_vs_exact=err)
        print(f"  pinned {name}: info {info}, grad rel err vs analytic {err:.2e}")

    # ---- config 4 at the default LDC size with backward ----------------------------------------------------------
    A = problems.ldc_pressure_csr(100)
    with np.load(GOLD / "gmres_ldc100_step1_batched.npz") as z:
        b0 = torch.from_numpy(z["b"].copy())
    kw = dict(tol=1e-10, maxiter=1000, restart=30)
    b1 = b0.clone().requires_grad_(True)
    t0 = time.time()
    x, info = ref.gmres(A.to_sparse_coo(), b1, **kw)
    (x ** 2).sum().backward()
    xo, _, _ = orc.gmres(A, b0, None, **kw)
    g_orc = orc.adjoint_grad_b("gmres", A, 2.0 * xo, None, **kw)
    assert torch.equal(x.detach(), xo), "oracle forward differs from the reference on LDC-100"
    assert torch.equal(b1.grad, g_orc), "oracle adjoint differs from the reference on LDC-100"
    np.savez_compressed(GOLD / "autograd_gmres_ldc100.npz", b=b0.numpy(), grad_b=b1.grad.numpy(), x=x.detach().numpy())
    r2["autograd_gmres_ldc100"] = dict(kind="gmres", info=int(info), kwargs=kw, n=int(b0.numel()),
                                       gen=dict(matrix="ldc", nx=100), seconds=round(time.time() - t0, 1))
    print(f"  pinned autograd_gmres_ldc100: info {info} ({time.time() - t0:.1f} s)")

    # ---- complex systems (reference :100-127, :1220): dense complex128 A, the reference's own conventions ----------
    g = torch.Generator().manual_seed(21)
    n = 48
    Br = torch.randn(n, n, dtype=torch.float64, generator=g)
    Bi = torch.randn(n, n, dtype=torch.float64, generator=g)
    B = torch.complex(Br, Bi)
    herm = B @ B.conj().T + 40.0 * torch.eye(n, dtype=torch.complex128)          # hermitian positive definite
    gen = B + 12.0 * torch.eye(n, dtype=torch.complex128)                        # general
    bc = torch.complex(torch.randn(n, dtype=torch.float64, generator=g), torch.randn(n, dtype=torch.float64, generator=g))
    for name, kind, Ac, kw in (("complex_cg_herm48", "cg", herm, dict(tol=1e-10)),
                               ("complex_bicgstab_gen48", "bicgstab", gen, dict(tol=1e-10)),
                               ("complex_gmres_gen48", "gmres", gen, dict(tol=1e-10, restart=30))):
        try:
            x, info = getattr(ref, kind)(Ac, bc, **kw)
            res = float(torch.linalg.norm(bc - Ac @ x) / torch.linalg.norm(bc))
            np.savez_compressed(GOLD / f"{name}.npz", A=Ac.numpy(), b=bc.numpy(), x=x.numpy())
            r2[name] = dict(kind=kind, info=int(info), kwargs=kw, n=n, relres=res, dtype=str(x.dtype))
            print(f"  pinned {name}: info {info}, relres {res:.2e}, x dtype {x.dtype}")
        except Exception as e:  # the reference itself may not support the combination: record that fact
            r2[name] = dict(kind=kind, kwargs=kw, n=n, reference_raises=f"{type(e).__name__}: {e}"[:300])
            print(f"  {name}: reference raises {type(e).__name__}: {str(e)[:120]}")

    # ---- block-Jacobi preconditioner written the way users of the reference would (torch.bmm with the block inverses) ---
    from pytorch_sparse_solver.module_a.preconditioners import BlockJacobiPreconditioner
    for kind, A, kw, bs in (("cg", problems.scaled_poisson3d_csr(8), dict(tol=1e-10), 4),
                            ("bicgstab", problems.scaled_convdiff3d_csr(8), dict(tol=1e-10), 8),
                            ("gmres", problems.scaled_convdiff3d_csr(8), dict(tol=1e-10, restart=20), 3)):
        Binv = BlockJacobiPreconditioner.diagonal_block_inverses(A, bs)
        n = A.shape[0]
        nb = Binv.shape[0]

        def Mb(r, Binv=Binv, bs=bs, n=n, nb=nb):
            rp = torch.zeros(nb * bs, dtype=r.dtype)
            rp[:n] = r
            return torch.bmm(Binv, rp.view(nb, bs, 1)).view(-1)[:n]
        b0, _ = problems.manufactured_rhs(A, 6)
        calls = {"n": 0}

        def Aop(v, A=A, calls=calls):
            calls["n"] += 1
            return torch.matmul(A, v)
        x, info = getattr(ref, kind)(Aop, b0, M=Mb, **kw)
        name = f"{kind}_blockjacobi_bs{bs}"
        np.savez_compressed(GOLD / f"{name}.npz", b=b0.numpy(), x=x.numpy())
        r2[name] = dict(kind=kind, info=int(info), kwargs=kw, n=int(n), block_size=bs, matvecs_ref=calls["n"],
                        gen=dict(matrix="scaled_poisson3d" if kind == "cg" else "scaled_convdiff3d", n=8, seed=7))
        print(f"  pinned {name}: info {info}, matvecs {calls['n']}")

    if "--full" in sys.argv:
        t0 = time.time()
        A = problems.convdiff3d_csr(256)
        b, _ = problems.manufactured_rhs(A, 0)
        x, info = ref.bicgstab(A, b, tol=1e-10)
        calls = {"n": 0}

        def Aop(v):
            calls["n"] += 1
            return torch.matmul(A, v)
        # matvec count on a second, counting run would double the cost: derive iterations from the oracle's digest
        idx = torch.linspace(0, b.numel() - 1, 4096).long()
        relres = float(torch.linalg.norm(b - torch.matmul(A, x)) / torch.linalg.norm(b))
        np.savez_compressed(GOLD / "digest_bicgstab_cd3d256_tol1e-10.npz", x_sample_idx=idx.numpy(),
                            x_sample=x[idx].numpy())
        r2["digest_bicgstab_cd3d256_tol1e-10"] = dict(info=int(info), x_norm=float(torch.linalg.norm(x)),
                                                       x_sum=float(x.sum()), relres=relres, threads=os.cpu_count(),
                                                       seconds=round(time.time() - t0, 1))
        print(f"  pinned digest_bicgstab_cd3d256_tol1e-10: info {info} ||x|| {float(torch.linalg.norm(x))!r} "
              f"relres {relres:.3e} ({time.time() - t0:.0f} s)")

    with open(GOLD / "manifest.json", "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("manifest updated")


if __name__ == "__main__":
    main()
