#!/usr/bin/env python3
"""Where the end-to-end (host-buffer) time goes: H2D, registration (index narrowing, statistics, TMA plan,
dictionary coding), solve, D2H."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
import torch  # noqa: E402

from pytorch_sparse_solver import _native, problems  # noqa: E402

dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
A = problems.poisson3d_csr(n, device=dev)
h = _native.Handle.get(dev)
N = A.shape[0]
b = torch.ones(N, dtype=torch.float64, device=dev)


def t(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for cmp_ in (2, 1, 0):
    h.set_option("use_compress", cmp_)

    def reg():
        _native.clear_cache()
        return _native.register_matrix(A)
    print(f"use_compress={cmp_}: register_matrix (device int64 CSR) {t(reg):8.2f} ms")
crow_h = A.crow_indices().cpu().pin_memory()
col_h = A.col_indices().cpu().pin_memory()
val_h = A.values().cpu().pin_memory()
b_h = b.cpu().pin_memory()
nbytes = sum(x.numel() * x.element_size() for x in (crow_h, col_h, val_h, b_h))
print(f"H2D of {nbytes/1e9:.2f} GB: {t(lambda: [x.to(dev, non_blocking=True) for x in (crow_h, col_h, val_h, b_h)]):8.2f} ms")
for cmp_ in (2, 1, 0):
    h.set_option("use_compress", cmp_)
    ms = t(lambda: _native.solve_host(_native.METHOD_CG, crow_h, col_h, val_h, b_h, None, 1e-8, 0.0, None))
    _native.clear_cache()
    m = _native.register_matrix(A)
    ms_dev = t(lambda: m.cg(b, None, 1e-8, 0.0, None))
    print(f"use_compress={cmp_}: solve_host {ms:8.2f} ms ; device-resident solve {ms_dev:8.2f} ms")
