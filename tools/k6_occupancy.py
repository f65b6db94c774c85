#!/usr/bin/env python3
"""GPU probe: kernel 6 (row-bitmask SpMV) at 3..5 CTAs/SM, fp32 (40 registers, no spills) and fp64 (spills at 5)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
import torch  # noqa: E402

from pytorch_sparse_solver import _native, problems  # noqa: E402


def time_gpu(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
h = _native.Handle.get(dev)
h.set_option("mask_const", 0)   # kernel 6 itself (kernel 7 would serve the fp64 stencil otherwise)
for dt in (torch.float32, torch.float64):
    A = problems.poisson3d_csr(n, device=dev, dtype=dt)
    m = _native.register_matrix(A, dt)
    x = torch.randn(A.shape[0], dtype=dt, device=dev)
    for ctas in (3, 4, 5):
        for grp in (4, 8, 16):
            h.set_option("mask_ctas", ctas)
            h.set_option("mask_group", grp)
            us = 1e3 * time_gpu(lambda: m.spmv_dot(x, x))
            print(json.dumps(dict(what="k6_occ", dtype=str(dt), kernel=m.info()["kernel"], mask_ctas=ctas, mask_group=grp,
                                  us=round(us, 2))), flush=True)
    del m, A
    _native.clear_cache()
h.set_option("mask_ctas", 4)
h.set_option("mask_group", 8)
h.set_option("mask_const", 1)
