import sys, time
sys.path.insert(0, "/root/repo/pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200")
import torch
from pytorch_sparse_solver import _native, module_a, problems
from pytorch_sparse_solver.module_a import krylov
dev = torch.device("cuda", 0)
h = _native.Handle.get(dev)
for nx in (64, 256, 512, 1024):
    A = problems.poisson2d_csr(nx, nx, device=dev)
    b = torch.ones(A.shape[0], dtype=torch.float64, device=dev)
    for opts in (dict(), dict(fuse_xpay=1), dict(fuse_xpay=1, use_tma=0), dict(use_tma=0), dict(chunk=64), dict(fuse_xpay=1, chunk=64)):
        for k in ("fuse_xpay", "use_tma", "chunk"):
            h.set_option(k, {"fuse_xpay": 0, "use_tma": 1, "chunk": 0}[k])
        for k, v in opts.items():
            h.set_option(k, v)
        _native.clear_cache()
        module_a.cg(A, b, tol=1e-8)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            x, info = module_a.cg(A, b, tol=1e-8)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        r = krylov.last_result
        print(f"nx={nx} {opts}: {1e3*dt:.2f} ms, {r['iterations']} it, {1e6*dt/r['iterations']:.2f} us/it info={info}")
