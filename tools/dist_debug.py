#!/usr/bin/env python3
"""Debug helper (torchrun, one rank per GPU): prints distributed vs single-GPU BiCGStab/GMRES results side by side."""
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
    from pytorch_sparse_solver import _native, problems
    from pytorch_sparse_solver import distributed as bkd
    from pytorch_sparse_solver.module_a.krylov import _gmres_effective_tolerances
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    rows = n ** 3
    offsets = [q * rows for q in range(world + 1)]
    g3 = (1.0, 0.5, 0.25)
    cd = dict(lower=(-(1 + g3[0]), -(1 + g3[1]), -(1 + g3[2])), diag=6 + sum(g3))
    crow, col, val = problems.stencil3d_rows(n, world * n, rank * n, (rank + 1) * n, device=dev, **cd)
    Dc = bkd.DistMatrix(crow, col, val, offsets, rank, world)
    Ac = problems.stencil3d_csr(n, nz=world * n, device=dev, **cd)
    mc = _native.register_matrix(Ac)
    sl = slice(rank * rows, (rank + 1) * rows)
    xg = torch.randn(world * rows, dtype=torch.float64, device=dev, generator=torch.Generator(dev).manual_seed(5))
    dist.broadcast(xg, 0)
    bc = mc.spmv(xg)
    Ng = world * rows

    def rel(a, b):
        return float(torch.linalg.norm(a - b) / torch.linalg.norm(b))

    def show(name, r, rr, x, xr):
        if rank == 0:
            keys = ("iterations", "matvecs", "info", "status", "final_residual", "threshold")
            print(name, {k: r[k] for k in keys}, "| ref", {k: rr[k] for k in keys}, "| xrel %.3e" % rel(x, xr[sl]),
                  flush=True)

    te, ae = _gmres_effective_tolerances(1e-10, 0.0, Ng, 'cuda')
    for mode, p2p in ((1, 0), (2, 0), (1, 1), (2, 1)):
        Dc.handle.set_option("loop_mode", mode)
        Dc.handle.set_option("dist_p2p", p2p)
        xr, rr = mc.bicgstab(bc, None, 1e-10, 0.0, None)
        x, r = Dc.bicgstab(bc[sl].contiguous(), None, 1e-10, 0.0, None)
        show(f"bicgstab mode{mode} p2p{p2p}", r, rr, x, xr)
        for name, meth in (("batched", 0), ("incremental", 1)):
            xr, rr = mc.gmres(bc, None, te, ae, 30, 1000, meth)
            x, r = Dc.gmres(bc[sl].contiguous(), None, 1e-10, 0.0, 30, 1000, name)
            show(f"gmres-{name} mode{mode} p2p{p2p}", r, rr, x, xr)
    Dc.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
