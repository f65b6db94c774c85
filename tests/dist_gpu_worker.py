#!/usr/bin/env python3
"""Multi-GPU parity worker (launched by torchrun, one rank per GPU): distributed SpMV / CG / BiCGStab / GMRES on a row-partitioned
Poisson slab grid against the single-GPU library solve of the same global system.  Exits non-zero on failure."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from pytorch_sparse_solver import _native, module_a, problems
    from pytorch_sparse_solver import distributed as bkd
    from pytorch_sparse_solver.module_a import krylov
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rows = n ** 3
    offsets = [q * rows for q in range(world + 1)]
    crow, col, val = problems.stencil3d_rows(n, world * n, rank * n, (rank + 1) * n, device=dev)
    D = bkd.DistMatrix(crow, col, val, offsets, rank, world)
    assert D.plan.peers == [q for q in (rank - 1, rank + 1) if 0 <= q < world], D.plan.peers
    A = problems.stencil3d_csr(n, nz=world * n, device=dev)     # every rank also holds the global system (checker)
    m = _native.register_matrix(A)
    sl = slice(rank * rows, (rank + 1) * rows)

    def rel(a, b):
        return float(torch.linalg.norm(a - b) / torch.linalg.norm(b))

    xg = torch.randn(world * rows, dtype=torch.float64, device=dev, generator=torch.Generator(dev).manual_seed(5))
    dist.broadcast(xg, 0)
    y = D.spmv(xg[sl].contiguous())
    assert rel(y, m.spmv(xg)[sl]) <= 1e-14, "dist spmv"

    bg = torch.ones(world * rows, dtype=torch.float64, device=dev)
    x_ref, r_ref = m.cg(bg, None, 1e-8, 0.0, None)
    assert D.p2p or os.environ.get("BK_DIST_P2P") == "0", "peer-memory path should connect on an NVLink box"
    if D.p2p and os.environ.get("BK_DIST_FOLD", "1") != "0":
        assert D.folded, "a stencil slab must take the single-kernel (folded) SpMV on the peer path"
    # plain launches / CUDA graph  x  NCCL / peer-memory path  x  folded single-kernel SpMV / local + boundary rows;
    # every solve goes through the reference API: module_a.cg(A=DistMatrix, b=local slab)
    for mode, p2p, fold in ((1, 0, 0), (2, 0, 0), (1, 1, 1), (2, 1, 1), (2, 1, 0)):
        if p2p and not D.p2p:
            continue
        D.handle.set_option("loop_mode", mode)
        D.handle.set_option("dist_p2p", p2p)
        D.handle.set_option("dist_fold", fold)
        x, info = module_a.cg(D, bg[sl].contiguous(), tol=1e-8)
        r = dict(krylov.last_result)
        assert info == r_ref["info"] == 0 and r["route"] == "dist", (r, r_ref)
        assert abs(r["iterations"] - r_ref["iterations"]) <= 2, (r["iterations"], r_ref["iterations"])
        assert rel(x, x_ref[sl]) <= 1e-10, ("dist cg", mode, p2p, fold, rel(x, x_ref[sl]))
        x2, _ = module_a.cg(D, bg[sl].contiguous(), tol=1e-8)
        assert torch.equal(x, x2), "dist cg must be bitwise reproducible"
    D.handle.set_option("loop_mode", 0)
    D.handle.set_option("dist_p2p", 1)
    D.handle.set_option("dist_fold", 1)
    # implicit-diff backward through the row-partitioned transpose, against the single-GPU adjoint of the same system
    b1 = bg[sl].clone().requires_grad_(True)
    x1, _ = module_a.cg(D, b1, tol=1e-10)
    (x1 ** 2).sum().backward()
    bgr = bg.clone().requires_grad_(True)
    xg1, _ = module_a.cg(A, bgr, tol=1e-10)
    (xg1 ** 2).sum().backward()
    assert rel(b1.grad, bgr.grad[sl]) <= 1e-8, ("dist grad_b", rel(b1.grad, bgr.grad[sl]))
    # built-in Jacobi preconditioner on the partitioned matrix
    Mj = module_a.JacobiPreconditioner(D)
    xj, infoj = module_a.cg(D, bg[sl].contiguous(), tol=1e-10, M=Mj)
    xjr, _ = module_a.cg(A, bg, tol=1e-10, M=module_a.JacobiPreconditioner(A))
    assert infoj == 0 and rel(xj, xjr[sl]) <= 1e-9, ("dist jacobi cg", rel(xj, xjr[sl]))
    if D.p2p:   # halo push folded into the p-update kernel (default) vs the separate push kernel: same bits
        xs = {}
        for fuse in (1, 0, 1):
            D.handle.set_option("dist_fuse_push", fuse)
            xs[fuse], rf = D.cg(bg[sl].contiguous(), None, 1e-8, 0.0, None)
            assert rf["info"] == 0 and rf["iterations"] == r["iterations"], (fuse, rf, r)
        assert torch.equal(xs[0], xs[1]), "fused and separate halo push must give identical results"
        D.handle.set_option("dist_fuse_push", 1)
        # lagged-x cut (x updated every second iteration, p ping-pongs): identical bits, whichever parity the loop stops in
        for maxit in (1, 2, 3, 4, 5, 8, 9, None):
            xl = {}
            for lag in (0, 1):
                D.handle.set_option("cg_lag_x", lag)
                xl[lag], rl = D.cg(bg[sl].contiguous(), None, 1e-8, 0.0, maxit)
            assert torch.equal(xl[0], xl[1]), ("dist cg lagged x", maxit)
        D.handle.set_option("cg_lag_x", 1)
        # solves that end before / at the first iteration must leave the push/wait counters paired
        xz, rz = D.cg(torch.zeros_like(bg[sl]), None, 1e-8, 0.0, None)
        assert rz["info"] == 0 and rz["iterations"] == 0 and float(xz.abs().max()) == 0.0
        x1, r1 = D.cg(bg[sl].contiguous(), None, 0.0, 0.0, 1)
        assert r1["iterations"] == 1
        xa, ra = D.cg(bg[sl].contiguous(), None, 1e-8, 0.0, None)
        assert torch.equal(xa, xs[1]), "state after short solves"
    # fixed window + warm start
    x0 = xg * 0.01
    x_ref, r_ref = m.cg(bg, x0, 0.0, 0.0, 7)
    x, r = D.cg(bg[sl].contiguous(), x0[sl].contiguous(), 0.0, 0.0, 7)
    assert r["iterations"] == 7 and rel(x, x_ref[sl]) <= 1e-12, ("window", rel(x, x_ref[sl]))
    D.close()
    dist.barrier()

    # ---- BiCGStab / GMRES on a non-symmetric slab system (convection-diffusion, SURVEY config 3 at test size)
    from pytorch_sparse_solver.module_a.krylov import _gmres_effective_tolerances
    g3 = (1.0, 0.5, 0.25)
    cd = dict(lower=(-(1 + g3[0]), -(1 + g3[1]), -(1 + g3[2])), diag=6 + sum(g3))
    crow, col, val = problems.stencil3d_rows(n, world * n, rank * n, (rank + 1) * n, device=dev, **cd)
    Dc = bkd.DistMatrix(crow, col, val, offsets, rank, world)
    Ac = problems.stencil3d_csr(n, nz=world * n, device=dev, **cd)
    mc = _native.register_matrix(Ac)
    bc = mc.spmv(xg)                                            # manufactured right-hand side
    Ng = world * rows
    xb_ref, rb_ref = mc.bicgstab(bc, None, 1e-10, 0.0, None)
    te, ae = _gmres_effective_tolerances(1e-10, 0.0, Ng, 'cuda')
    xg_ref, rg_ref = mc.gmres(bc, None, te, ae, 30, 1000, _native.BK_GMRES_BATCHED)
    xi_ref, ri_ref = mc.gmres(bc, None, te, ae, 30, 1000, _native.BK_GMRES_INCREMENTAL)
    x0c = xg * 0.5
    xw_ref, rw_ref = mc.bicgstab(bc, x0c, 0.0, 0.0, 6)
    xgw_ref, rgw_ref = mc.gmres(bc, x0c, *_gmres_effective_tolerances(0.0, 0.0, Ng, 'cuda'), 12, 2,
                                _native.BK_GMRES_BATCHED)
    for mode, p2p, fold in ((1, 0, 0), (2, 0, 0), (1, 1, 1), (2, 1, 1), (2, 1, 0)):
        if p2p and not Dc.p2p:
            continue
        Dc.handle.set_option("loop_mode", mode)
        Dc.handle.set_option("dist_p2p", p2p)
        Dc.handle.set_option("dist_fold", fold)
        tag = ("mode", mode, "p2p", p2p, "fold", fold)
        x, r = Dc.bicgstab(bc[sl].contiguous(), None, 1e-10, 0.0, None)
        assert r["info"] == rb_ref["info"] == 0, (tag, r, rb_ref)
        assert abs(r["iterations"] - rb_ref["iterations"]) <= 2, (tag, r["iterations"], rb_ref["iterations"])
        assert rel(x, xb_ref[sl]) <= 1e-9, ("dist bicgstab", tag, rel(x, xb_ref[sl]))
        x2, _ = Dc.bicgstab(bc[sl].contiguous(), None, 1e-10, 0.0, None)
        assert torch.equal(x, x2), ("dist bicgstab must be bitwise reproducible", tag)
        x, r = Dc.bicgstab(bc[sl].contiguous(), x0c[sl].contiguous(), 0.0, 0.0, 6)
        assert r["iterations"] == 6 and rel(x, xw_ref[sl]) <= 1e-11, ("bicgstab window", tag, rel(x, xw_ref[sl]))
        for sm_name, (xr_, rr_) in (("batched", (xg_ref, rg_ref)), ("incremental", (xi_ref, ri_ref))):
            x, r = Dc.gmres(bc[sl].contiguous(), None, 1e-10, 0.0, 30, 1000, sm_name)
            assert r["info"] == rr_["info"] == 0, (tag, sm_name, r, rr_)
            assert abs(r["iterations"] - rr_["iterations"]) <= 1, (tag, sm_name, r["iterations"], rr_["iterations"])
            assert abs(r["matvecs"] - rr_["matvecs"]) <= 2, (tag, sm_name, r["matvecs"], rr_["matvecs"])
            assert rel(x, xr_[sl]) <= 1e-10, ("dist gmres", tag, sm_name, rel(x, xr_[sl]))
        x2, _ = Dc.gmres(bc[sl].contiguous(), None, 1e-10, 0.0, 30, 1000, "incremental")
        assert torch.equal(x, x2), ("dist gmres must be bitwise reproducible", tag)
        x, r = Dc.gmres(bc[sl].contiguous(), x0c[sl].contiguous(), 0.0, 0.0, 12, 2, "batched")
        assert r["iterations"] == 2 and rel(x, xgw_ref[sl]) <= 1e-11, ("gmres window", tag, rel(x, xgw_ref[sl]))
    Dc.handle.set_option("loop_mode", 0)
    Dc.handle.set_option("dist_p2p", 1)
    Dc.handle.set_option("dist_fold", 1)
    # reference API + Jacobi + backward on the non-symmetric system (BiCGStab, GMRES)
    for kind, kw in (("bicgstab", dict(tol=1e-10)), ("gmres", dict(tol=1e-10, restart=30))):
        b1 = bc[sl].clone().requires_grad_(True)
        x1, i1 = getattr(module_a, kind)(Dc, b1, **kw)
        (x1 ** 2).sum().backward()
        bgr = bc.clone().requires_grad_(True)
        xg1, ig = getattr(module_a, kind)(Ac, bgr, **kw)
        (xg1 ** 2).sum().backward()
        assert i1 == ig == 0 and rel(x1.detach(), xg1.detach()[sl]) <= 1e-9, (kind, rel(x1.detach(), xg1.detach()[sl]))
        assert rel(b1.grad, bgr.grad[sl]) <= 1e-7, ("dist grad_b", kind, rel(b1.grad, bgr.grad[sl]))
        xj, ij = getattr(module_a, kind)(Dc, bc[sl].contiguous(), M=module_a.JacobiPreconditioner(Dc), **kw)
        xjr, _ = getattr(module_a, kind)(Ac, bc, M=module_a.JacobiPreconditioner(Ac), **kw)
        assert ij == 0 and rel(xj, xjr[sl]) <= 1e-8, ("dist jacobi", kind, rel(xj, xjr[sl]))
    Dc.close()
    dist.barrier()

    # ---- all-to-all coupling: every rank exchanges halos with every other rank (SPD: diagonally dominant, symmetric)
    N = 6000 * world
    i = torch.arange(N, device=dev)
    rows_, cols_, vals_ = [i], [i], [torch.full((N,), 8.0, dtype=torch.float64, device=dev)]
    for shift, v in ((1, -1.0), (N // world + 7, -0.7), (N // 2 + 3, -0.5), (3 * (N // world) + 11, -0.3)):
        j = (i + shift) % N
        rows_ += [i, j]
        cols_ += [j, i]
        vals_ += [torch.full((N,), v, dtype=torch.float64, device=dev)] * 2
    G = torch.sparse_coo_tensor(torch.stack([torch.cat(rows_), torch.cat(cols_)]), torch.cat(vals_), (N, N)).coalesce()
    G = G.to_sparse_csr()
    offs = bkd.partition_rows(N, world)
    rb, re_ = offs[rank], offs[rank + 1]
    gc = G.crow_indices()
    lo, hi = int(gc[rb]), int(gc[re_])
    D2 = bkd.DistMatrix((gc[rb:re_ + 1] - gc[rb]).contiguous(), G.col_indices()[lo:hi].contiguous(),
                        G.values()[lo:hi].contiguous(), offs, rank, world)
    if world > 2:
        assert len(D2.plan.peers) == world - 1, D2.plan.peers
    mg = _native.register_matrix(G)
    bg2 = torch.randn(N, dtype=torch.float64, device=dev, generator=torch.Generator(dev).manual_seed(9))
    dist.broadcast(bg2, 0)
    assert rel(D2.spmv(bg2[rb:re_].contiguous()), mg.spmv(bg2)[rb:re_]) <= 1e-14, "dist spmv (all-to-all)"
    xr2, rr2 = mg.cg(bg2, None, 1e-10, 0.0, None)
    assert not D2.folded, "irregular couplings do not fit the row-bitmask plan: two-kernel path"
    for p2p in ((0, 1) if D2.p2p else (0,)):
        D2.handle.set_option("dist_p2p", p2p)
        x2, r2 = D2.cg(bg2[rb:re_].contiguous(), None, 1e-10, 0.0, None)
        assert r2["info"] == rr2["info"] == 0 and abs(r2["iterations"] - rr2["iterations"]) <= 2, (r2, rr2)
        assert rel(x2, xr2[rb:re_]) <= 1e-10, ("dist cg all-to-all", p2p, rel(x2, xr2[rb:re_]))
        # the same couplings drive BiCGStab and GMRES (all-to-all halos, fp64)
        xb2, rb2 = mg.bicgstab(bg2, None, 1e-10, 0.0, None)
        x2, r2 = D2.bicgstab(bg2[rb:re_].contiguous(), None, 1e-10, 0.0, None)
        assert r2["info"] == rb2["info"] == 0 and abs(r2["iterations"] - rb2["iterations"]) <= 2, (r2, rb2)
        assert rel(x2, xb2[rb:re_]) <= 1e-9, ("dist bicgstab all-to-all", p2p, rel(x2, xb2[rb:re_]))
        te2, ae2 = _gmres_effective_tolerances(1e-10, 0.0, N, 'cuda')
        xq2, rq2 = mg.gmres(bg2, None, te2, ae2, 20, 1000, _native.BK_GMRES_BATCHED)
        x2, r2 = D2.gmres(bg2[rb:re_].contiguous(), None, 1e-10, 0.0, 20, 1000)
        assert r2["info"] == rq2["info"] == 0 and abs(r2["iterations"] - rq2["iterations"]) <= 1, (r2, rq2)
        assert rel(x2, xq2[rb:re_]) <= 1e-10, ("dist gmres all-to-all", p2p, rel(x2, xq2[rb:re_]))
    D2.handle.set_option("dist_p2p", 1)
    D2.close()
    dist.barrier()
    if rank == 0:
        print(f"dist worker OK: world={world} n={n} iterations={r_ref['iterations']} all-to-all iterations={rr2['iterations']}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
