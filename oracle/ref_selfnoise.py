#!/usr/bin/env python3
"""Self-noise of the UNMODIFIED reference on BASELINE configs[2] at full size (build container only): BiCGStab on
CD3D-256, manufactured RHS, tol 1e-10, run with different OpenMP thread counts (different summation orders inside
torch's CPU dot / SpMV).  Reports iterations (matvec count via a counting callable), ||x|| and the relative difference
between the runs' solutions; results go to tests/golden/manifest.json -> round2.ref_selfnoise_bicgstab_cd3d256.
    python oracle/ref_selfnoise.py [threads ...]      (default: 8 3)"""
import json
import sys
import time
import warnings
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
warnings.filterwarnings("ignore")
from pin_reference import load_reference, PKG_DIR  # noqa: E402


def main():
    threads = [int(a) for a in sys.argv[1:]] or [8, 3]
    ref, _ = load_reference()
    sys.path.insert(0, str(PKG_DIR))
    from pytorch_sparse_solver import problems
    A = problems.convdiff3d_csr(256)
    b, _ = problems.manufactured_rhs(A, 0)
    out, xs = {}, []
    for th in threads:
        torch.set_num_threads(th)
        calls = {"n": 0}

        def Aop(v):
            calls["n"] += 1
            return torch.matmul(A, v)
        t0 = time.time()
        x, info = ref.bicgstab(Aop, b, tol=1e-10)
        out[str(th)] = dict(matvecs=calls["n"], iterations_est=(calls["n"] - 2) / 2.0, info=int(info),
                            x_norm=float(torch.linalg.norm(x)), seconds=round(time.time() - t0, 1))
        xs.append(x)
        print(th, out[str(th)], flush=True)
    rel = float(torch.linalg.norm(xs[0] - xs[-1]) / torch.linalg.norm(xs[0])) if len(xs) > 1 else 0.0
    out["rel_diff_between_runs"] = rel
    print("rel diff between runs", rel)
    mf = ROOT / "tests" / "golden" / "manifest.json"
    m = json.loads(mf.read_text())
    m.setdefault("round2", {})["ref_selfnoise_bicgstab_cd3d256"] = out
    mf.write_text(json.dumps(m, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
