#!/usr/bin/env python3
"""GPU probe: is kernel 6 / 6G bound by L2->SM traffic?  Same row count (2^24), same kernels, but a 5-point pattern whose
gathers all stay near (P2D 65536 x 256: offsets +-1, +-256) vs the 7-point 3-D pattern (far offsets +-65536)."""
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
h = _native.Handle.get(dev)
h.set_option("mask_cctas", 4)
for kind in ("p2d_65536x256", "p2d_4096x4096", "p3d_256"):
    if kind == "p2d_65536x256":
        A = problems.poisson2d_csr(65536, 256, device=dev)
    elif kind == "p2d_4096x4096":
        A = problems.poisson2d_csr(4096, 4096, device=dev)
    else:
        A = problems.poisson3d_csr(256, device=dev)
    m = _native.register_matrix(A)
    x = torch.randn(A.shape[0], dtype=torch.float64, device=dev)
    for const in (0, 1):
        h.set_option("mask_const", const)
        for pf in (1, 0):
            h.set_option("mask_prefetch", pf)
            us = 1e3 * time_gpu(lambda: m.spmv_dot(x, x))
            print(json.dumps(dict(what="k6_traffic", kind=kind, kernel=m.info()["kernel"], group_unrolled=const, l2_prefetch=pf,
                                  n=A.shape[0], us=round(us, 2))), flush=True)
    h.set_option("mask_prefetch", 1)
    del m, A
    _native.clear_cache()
