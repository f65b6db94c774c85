#!/usr/bin/env python3
"""GPU probe: kernel 6G (group-unrolled, patterns in the parameter block) vs kernel 6, fp64: bitwise check + timing sweep."""
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
for kind in ("p3d", "cd3d"):
    A = problems.poisson3d_csr(n, device=dev) if kind == "p3d" else problems.convdiff3d_csr(n, device=dev)
    m = _native.register_matrix(A)
    x = torch.randn(A.shape[0], dtype=torch.float64, device=dev)
    h.set_option("mask_const", 0)
    y0, d0 = m.spmv_dot(x, x)
    y0 = y0.clone()
    us0 = 1e3 * time_gpu(lambda: m.spmv_dot(x, x))
    print(json.dumps(dict(what="k6", kind=kind, patterns=m.info().get("mask_patterns"), us=round(us0, 2))), flush=True)
    h.set_option("mask_const", 1)
    for cctas in (4, 5, 6):
        h.set_option("mask_cctas", cctas)
        y1, d1 = m.spmv_dot(x, x)
        same = bool(torch.equal(y1, y0))
        us = 1e3 * time_gpu(lambda: m.spmv_dot(x, x))
        print(json.dumps(dict(what="k7", kind=kind, mask_cctas=cctas, us=round(us, 2), bitwise_y=same,
                              dot_rel=float(abs(d1 - d0) / abs(d0)))), flush=True)
    h.set_option("mask_cctas", 5)
    del m, A
    _native.clear_cache()
