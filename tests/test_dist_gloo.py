"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): row partition, local/ghost split and halo plan of
pytorch_sparse_solver.distributed — the exact arrays the C library consumes — validated by emulating the
distributed SpMV with torch CPU ops and gloo send/recv and comparing with the global product."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
PKG_DIR = ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _emulated_dist_spmv(sp, plan, x_local, rank):
    """What bk_dist_spmv does, with torch CPU ops: pack, exchange, local block, ghost rows."""
    sendbuf = x_local[plan.send_idx.long()]
    ghost = torch.zeros(sp.ghost_ids.numel(), dtype=x_local.dtype)
    reqs, so, ro = [], 0, 0
    for q, sc, rc in zip(plan.peers, plan.send_counts, plan.recv_counts):
        if sc:
            reqs.append(dist.isend(sendbuf[so:so + sc].contiguous(), dst=q))
        if rc:
            reqs.append(dist.irecv(ghost[ro:ro + rc], src=q))
        so += sc
        ro += rc
    for r in reqs:
        r.wait()
    A_loc = torch.sparse_csr_tensor(sp.loc_rowptr.long(), sp.loc_col.long(), sp.loc_val, size=(sp.n_local, sp.n_local))
    y = torch.matmul(A_loc, x_local)
    if sp.brow_ids.numel():
        A_gh = torch.sparse_csr_tensor(sp.gh_rowptr.long(), sp.gh_col.long(), sp.gh_val,
                                       size=(sp.brow_ids.numel(), max(sp.ghost_ids.numel(), 1)))
        gpad = ghost if ghost.numel() else torch.zeros(1, dtype=x_local.dtype)
        y[sp.brow_ids.long()] += torch.matmul(A_gh, gpad)
    return y, ghost


def _worker(rank, world, port, kind):
    sys.path.insert(0, str(PKG_DIR))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pytorch_sparse_solver import distributed as bkd
        from pytorch_sparse_solver import problems
        if kind == "poisson3d":
            A = problems.poisson3d_csr(6)
        elif kind == "slab":
            A = problems.stencil3d_csr(4, nz=4 * world)
        else:  # random coupling: every rank talks to every rank, some rows have no ghost entries
            g = torch.Generator().manual_seed(7)
            n = 97
            D = torch.randn(n, n, dtype=torch.float64, generator=g)
            D[torch.rand(n, n, generator=g) > 0.08] = 0.0
            D += torch.eye(n, dtype=torch.float64) * 3
            A = D.to_sparse_csr()
        n = A.shape[0]
        offsets = bkd.partition_rows(n, world)
        assert offsets[0] == 0 and offsets[-1] == n and all(b >= a for a, b in zip(offsets, offsets[1:]))
        rb, re_ = offsets[rank], offsets[rank + 1]
        crow, col, val = A.crow_indices(), A.col_indices(), A.values()
        lcrow = crow[rb:re_ + 1] - crow[rb]
        sl = slice(int(crow[rb]), int(crow[re_]))
        sp = bkd.split_local_ghost(lcrow, col[sl], val[sl], rb, re_)
        plan = bkd.build_halo_plan(sp.ghost_ids, offsets, rank, world)
        # structure checks
        assert sp.loc_rowptr.dtype == torch.int32 and sp.gh_col.dtype == torch.int32
        assert int(sp.loc_rowptr[-1]) + int(sp.gh_rowptr[-1]) == val[sl].numel()
        assert sum(plan.recv_counts) == sp.ghost_ids.numel()
        assert plan.send_idx.numel() == sum(plan.send_counts)
        assert rank not in plan.peers
        if plan.send_idx.numel():
            assert int(plan.send_idx.min()) >= 0 and int(plan.send_idx.max()) < sp.n_local
        if kind == "slab":  # 1-D slabs: only nearest neighbours, one n x n plane each way
            assert plan.peers == [q for q in (rank - 1, rank + 1) if 0 <= q < world]
            assert all(c == 16 for c in plan.send_counts) and all(c == 16 for c in plan.recv_counts)
        # numerical check against the global product
        x = torch.randn(n, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
        y_ref = torch.matmul(A, x)[rb:re_]
        y, ghost = _emulated_dist_spmv(sp, plan, x[rb:re_].contiguous(), rank)
        assert torch.equal(ghost, x[sp.ghost_ids])
        assert float((y - y_ref).abs().max()) <= 1e-13 * float(y_ref.abs().max() + 1)
        # the rows as ONE matrix over [local | ghost] (what bk_dist_set_extended registers): same product, and the
        # entries keep the order of the global matrix's rows
        assert sp.ext_col.dtype == torch.int32 and sp.ext_col.numel() == val[sl].numel()
        A_ext = torch.sparse_csr_tensor(sp.ext_rowptr.long(), sp.ext_col.long(), val[sl],
                                        size=(sp.n_local, sp.n_local + max(sp.ghost_ids.numel(), 1)))
        x_ext = torch.cat([x[rb:re_], ghost if ghost.numel() else torch.zeros(1, dtype=x.dtype)])
        assert float((torch.matmul(A_ext, x_ext) - y_ref).abs().max()) <= 1e-13 * float(y_ref.abs().max() + 1)
        gid = torch.where(sp.ext_col.long() < sp.n_local, sp.ext_col.long() + rb,
                          sp.ghost_ids[(sp.ext_col.long() - sp.n_local).clamp(min=0)])
        assert torch.equal(gid, col[sl])
        # the transposed slab (adjoint solves): exchange by column owner, against the global transpose
        tcrow, tcol, tval = bkd.transpose_slab(lcrow, col[sl], val[sl], offsets, rank, world)
        At = A.to_dense().T.contiguous().to_sparse_csr()
        tc = At.crow_indices()
        assert torch.equal(tcrow, tc[rb:re_ + 1] - tc[rb])
        assert torch.equal(tcol, At.col_indices()[int(tc[rb]):int(tc[re_])])
        assert torch.equal(tval, At.values()[int(tc[rb]):int(tc[re_])])
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kind", [(2, "poisson3d"), (2, "slab"), (3, "slab"), (2, "random"), (3, "random")])
def test_partition_and_halo_plan_gloo(world, kind):
    mp.spawn(_worker, args=(world, _free_port(), kind), nprocs=world, join=True)


def test_single_rank_split_is_all_local():
    sys.path.insert(0, str(PKG_DIR))
    from pytorch_sparse_solver import distributed as bkd
    from pytorch_sparse_solver import problems
    A = problems.poisson3d_csr(5)
    sp = bkd.split_local_ghost(A.crow_indices(), A.col_indices(), A.values(), 0, A.shape[0])
    assert sp.ghost_ids.numel() == 0 and sp.brow_ids.numel() == 0
    assert torch.equal(sp.loc_col.long(), A.col_indices()) and torch.equal(sp.loc_rowptr.long(), A.crow_indices())
    plan = bkd.build_halo_plan(sp.ghost_ids, [0, A.shape[0]], 0, 1)
    assert plan.peers == [] and plan.send_idx.numel() == 0


def test_slab_rows_match_global_matrix():
    sys.path.insert(0, str(PKG_DIR))
    from pytorch_sparse_solver import problems
    A = problems.stencil3d_csr(4, nz=12)
    for q in range(3):
        crow, col, val = problems.stencil3d_rows(4, 12, 4 * q, 4 * (q + 1))
        rb, re_ = 64 * q, 64 * (q + 1)
        gc = A.crow_indices()
        assert torch.equal(crow, gc[rb:re_ + 1] - gc[rb])
        assert torch.equal(col, A.col_indices()[int(gc[rb]):int(gc[re_])])
        assert torch.equal(val, A.values()[int(gc[rb]):int(gc[re_])])
