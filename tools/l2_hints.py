#!/usr/bin/env python3
"""GPU sweep (not part of the product): CG on P3D-n with the L2 streaming hints of the vector kernels (option
`l2_hints`, bit mask) and the snake order on/off.  JSON lines -> gpurun_out/l2_hints.jsonl."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from pytorch_sparse_solver import _native, problems  # noqa: E402

OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
LOG = open(OUT / "l2_hints.jsonl", "a")


def emit(**kw):
    s = json.dumps(kw)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--hints", type=str, default="0,1,2,4,6,7,15,31,63")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--loop-mode", type=int, default=0, help="1 = plain stream launches (for ncu)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    A = problems.poisson3d_csr(args.n, device=dev)
    N = A.shape[0]
    b = torch.ones(N, dtype=torch.float64, device=dev)
    h = _native.Handle.get(dev)
    m = _native.register_matrix(A)
    if args.loop_mode:
        h.set_option("loop_mode", args.loop_mode)
    ref = None
    for snake in (1, 0):
        h.set_option("snake", snake)
        for hints in [int(t) for t in args.hints.split(",")]:
            h.set_option("l2_hints", hints)
            best = 1e30
            for _ in range(args.reps + 1):
                x, res = m.cg(b, None, 1e-30, 0.0, args.iters)
                best = min(best, res["device_ms"])
            if ref is None:
                ref = x.clone()
            emit(what="cg", n=args.n, snake=snake, l2_hints=hints, iterations=int(res["iterations"]),
                 us_per_iter=1e3 * best / max(int(res["iterations"]), 1), bitwise_same=bool(torch.equal(x, ref)))
    h.set_option("snake", 1)
    h.set_option("l2_hints", 0)


if __name__ == "__main__":
    main()
