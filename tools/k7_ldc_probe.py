#!/usr/bin/env python3
"""GPU probe: kernel 7 vs kernel 6 on the LDC pressure matrices (bitwise y, dots) and the multi-kernel GMRES path."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
import torch  # noqa: E402

from pytorch_sparse_solver import _native, problems, module_a  # noqa: E402
from pytorch_sparse_solver.module_a import krylov  # noqa: E402

dev = torch.device("cuda", 0)
h = _native.Handle.get(dev)
for nx in (32, 100, 257):
    A = problems.ldc_pressure_csr(nx, device=dev)
    m = _native.register_matrix(A)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(A.shape[0], dtype=torch.float64, generator=g).to(dev)
    out = {}
    for const in (0, 1):
        h.set_option("mask_const", const)
        y, d = m.spmv_dot(x, x)
        out[const] = (y.clone(), float(d), m.info()["kernel"])
    print(json.dumps(dict(what="ldc_spmv", nx=nx, kernels=[out[0][2], out[1][2]], bitwise_y=bool(torch.equal(out[0][0], out[1][0])),
                          dot_rel=abs(out[0][1] - out[1][1]) / abs(out[0][1]))), flush=True)
with np.load(ROOT / "tests" / "golden" / "gmres_ldc100_step1_batched.npz") as z:
    b = torch.from_numpy(z["b"].copy()).to(dev)
    xref = torch.from_numpy(z["x"].copy()).to(dev)
A = problems.ldc_pressure_csr(100, device=dev)
for persistent in (1, 0):
    for const in (0, 1):
        h.set_option("persistent", persistent)
        h.set_option("mask_const", const)
        x, info = module_a.gmres(A, b, tol=1e-10, maxiter=1000, restart=30)
        r = krylov.last_result
        res = float(torch.linalg.norm(b - torch.mv(A, x)) / torch.linalg.norm(b))
        print(json.dumps(dict(what="ldc100_gmres", persistent=persistent, mask_const=const, cycles=int(r["iterations"]),
                              matvecs=int(r["matvecs"]), info=int(info), relres=res,
                              x_vs_ref=float(torch.linalg.norm(x - xref) / torch.linalg.norm(xref)))), flush=True)
h.set_option("persistent", 1)
h.set_option("mask_const", 1)
