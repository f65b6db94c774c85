#!/usr/bin/env python3
"""Latency-bound configs of BASELINE.json (1: CG on 2-D Poisson 256^2, 4: GMRES(30) on the LDC pressure systems):
solve times through the public API on one GPU.  Writes JSON lines to gpurun_out/small.jsonl."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pytorch_sparse_solver import _native, module_a, problems  # noqa: E402
from pytorch_sparse_solver.module_a import krylov  # noqa: E402

OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
LOG = open(OUT / "small.jsonl", "a")


def emit(**kw):
    s = json.dumps(kw)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


def main():
    dev = torch.device("cuda", 0)
    h = _native.Handle.get(dev)
    A = problems.poisson2d_csr(256, 256, device=dev)
    b = torch.ones(A.shape[0], dtype=torch.float64, device=dev)
    for mode, pers in ((0, 1), (0, 0), (1, 0)):
        h.set_option("loop_mode", mode)
        h.set_option("persistent", pers)
        dt, (x, info) = timed(lambda: module_a.cg(A, b, tol=1e-8))
        r = krylov.last_result
        emit(what="cg_p2d256", loop_mode=mode, persistent=pers, ms=1e3 * dt, iterations=r["iterations"], it_s=r["iterations"] / dt,
             us_per_iter=1e6 * dt / r["iterations"], info=info, launches=r["kernel_launches"])
    h.set_option("loop_mode", 0)
    h.set_option("persistent", 1)
    for nx, step in ((32, 1), (100, 0), (100, 1)):
        name = f"gmres_ldc{nx}_step{step}_batched"
        z = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
        bb = torch.from_numpy(z["b"]).to(dev)
        xr = torch.from_numpy(z["x"]).to(dev)
        L = problems.ldc_pressure_csr(nx, device=dev)
        for sm in ("batched", "incremental"):
            dt, (x, info) = timed(lambda: module_a.gmres(L, bb, tol=1e-10, maxiter=1000, restart=30, solve_method=sm), reps=3)
            r = krylov.last_result
            emit(what=f"gmres_ldc{nx}_step{step}", solve_method=sm, ms=1e3 * dt, cycles=r["iterations"],
                 matvecs=r["matvecs"], matvecs_per_s=r["matvecs"] / dt, info=info,
                 rel_vs_ref=float(torch.linalg.norm(x - xr) / torch.linalg.norm(xr)) if sm == "batched" else None)
        dt, (x, info) = timed(lambda: module_a.bicgstab(L, bb, tol=1e-10, maxiter=1000), reps=3)
        emit(what=f"bicgstab_ldc{nx}_step{step}", ms=1e3 * dt, iterations=krylov.last_result["iterations"], info=info)
    # LDC-512 (bandwidth-relevant variant, N = 262144)
    L = problems.ldc_pressure_csr(512, device=dev)
    g = torch.Generator().manual_seed(0)
    bb = torch.randn(L.shape[0], dtype=torch.float64, generator=g).to(dev)
    bb -= bb.mean()
    dt, (x, info) = timed(lambda: module_a.gmres(L, bb, tol=1e-8, maxiter=200, restart=30), reps=2)
    r = krylov.last_result
    emit(what="gmres_ldc512_rand", ms=1e3 * dt, cycles=r["iterations"], matvecs=r["matvecs"], info=info,
         ms_per_cycle=1e3 * dt / max(r["iterations"], 1))


if __name__ == "__main__":
    main()
