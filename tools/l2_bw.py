#!/usr/bin/env python3
"""GPU probe (not part of the product): streaming bandwidth of read-only / copy passes as a function of the footprint —
does an L2-resident vector stream faster than one in HBM?  JSON lines -> gpurun_out/l2_bw.jsonl."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
import torch  # noqa: E402

from pytorch_sparse_solver import _native  # noqa: E402

OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
LOG = open(OUT / "l2_bw.jsonl", "a")


def timeit(fn, reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda", 0)
    _native.Handle.get(dev)
    for mb in (4, 8, 16, 32, 48, 64, 96, 128, 256, 512):
        n = mb * (1 << 20) // 8
        x = torch.randn(n, dtype=torch.float64, device=dev)
        y = torch.randn(n, dtype=torch.float64, device=dev)
        z = torch.empty_like(x)
        reps = max(20, 4096 // mb)
        t_dot = timeit(lambda: _native.dot(x, y), reps)          # reads 2 vectors
        t_nrm = timeit(lambda: _native.dot(x, x), reps)          # reads 1 vector (twice the same line)
        t_axpby = timeit(lambda: _native.axpby(1.0, x, 2.0, y, out=z), reps)   # 2 reads + 1 write
        t_copy = timeit(lambda: z.copy_(x), reps)                # torch: 1 read + 1 write
        rec = dict(what="l2_bw", mb_per_vector=mb,
                   dot_gbs=2 * n * 8 / t_dot / 1e6, dot_us=1e3 * t_dot,
                   nrm_gbs=n * 8 / t_nrm / 1e6, nrm_us=1e3 * t_nrm,
                   axpby_gbs=3 * n * 8 / t_axpby / 1e6, axpby_us=1e3 * t_axpby,
                   copy_gbs=2 * n * 8 / t_copy / 1e6, copy_us=1e3 * t_copy)
        s = json.dumps(rec)
        print(s, flush=True)
        LOG.write(s + "\n")


if __name__ == "__main__":
    main()
