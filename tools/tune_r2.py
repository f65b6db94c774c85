#!/usr/bin/env python3
"""Round-2 GPU sweep (not part of the product): kernel 6 (row-bitmask SpMV) against kernels 5 / 3 / 2 on P3D-n, its
tunables (CTAs per SM, block-group size), and whole CG iterations.  JSON lines -> gpurun_out/tune_r2.jsonl."""
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
LOG = open(OUT / "tune_r2.jsonl", "a")


def emit(**kw):
    s = json.dumps(kw)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def time_gpu(fn, reps=20, warm=3):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--window", type=int, default=100)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n = args.n
    N = n ** 3
    A = problems.poisson3d_csr(n, device=dev)
    nnz = A.values().numel()
    x = torch.randn(N, dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    b = torch.ones(N, dtype=torch.float64, device=dev)
    bytes_alg = nnz * 12 + (N + 1) * 4 + 2 * N * 8
    h = _native.Handle.get(dev)
    for uc in (3, 2, 1, 0):
        h.set_option("use_compress", uc)
        _native.clear_cache()
        m = _native.register_matrix(A)
        info = m.info()
        actual = info["bytes_stream"] + 2 * N * 8
        ms = time_gpu(lambda: m.spmv_dot(x, x))
        ms0 = time_gpu(lambda: m.spmv(x, out=y))
        emit(what="spmv", use_compress=uc, kernel=info["kernel"], bytes_stream=info["bytes_stream"], us_dot=1e3 * ms,
             us_plain=1e3 * ms0, actual_gbs=actual / ms / 1e6, algorithmic_gbs=bytes_alg / ms / 1e6)
        if info["kernel"] == 6 and not args.quick:
            d_ctas, d_grp = h.get_option("mask_ctas"), h.get_option("mask_group")
            for ctas in (2, 3, 4):
                for wg in (2, 4, 8):
                    for st in (0, 2, 3):
                        h.set_option("mask_window", 1)
                        h.set_option("mask_ctas", ctas)
                        h.set_option("mask_wgroup", wg)
                        h.set_option("tma_stages", st)
                        ms = time_gpu(lambda: m.spmv_dot(x, x), reps=20)
                        emit(what="spmv_maskw", mask_ctas=ctas, mask_wgroup=wg, tma_stages=st, us_dot=1e3 * ms,
                             actual_gbs=actual / ms / 1e6)
            h.set_option("tma_stages", 0)
            h.set_option("mask_wgroup", 4)
            h.set_option("mask_window", 0)
            for ctas in (2, 3, 4, 5):
                for grp in (2, 4, 8, 16):
                    for pf in (0, 1):
                        h.set_option("mask_ctas", ctas)
                        h.set_option("mask_group", grp)
                        h.set_option("mask_prefetch", pf)
                        ms = time_gpu(lambda: m.spmv_dot(x, x), reps=20)
                        emit(what="spmv_mask", mask_ctas=ctas, mask_group=grp, mask_prefetch=pf, us_dot=1e3 * ms,
                             actual_gbs=actual / ms / 1e6)
            h.set_option("mask_prefetch", 1)
            h.set_option("mask_ctas", d_ctas)
            h.set_option("mask_group", d_grp)
            h.set_option("mask_window", 1)
        W = args.window
        ms = time_gpu(lambda: m.cg(b, None, 0.0, 0.0, W), reps=3, warm=1)
        emit(what="cg_window", use_compress=uc, kernel=info["kernel"], us_per_iter=1e3 * ms / W, it_s=W / ms * 1e3)
        for opts in ((dict(snake=0),) if not args.quick else ()):
            for k, v in opts.items():
                h.set_option(k, v)
            ms = time_gpu(lambda: m.cg(b, None, 0.0, 0.0, W), reps=3, warm=1)
            emit(what="cg_window", use_compress=uc, kernel=info["kernel"], opts=opts, us_per_iter=1e3 * ms / W)
            h.set_option("snake", 1)
    h.set_option("use_compress", 3)
    _native.clear_cache()
    m = _native.register_matrix(A)
    xs, res = m.cg(b, None, 1e-8, 0.0, None)
    emit(what="cg_full", **res)
    if not args.quick:
        C = problems.convdiff3d_csr(n, device=dev)
        mc = _native.register_matrix(C)
        bc, xt = problems.manufactured_rhs(C, 0)
        ms = time_gpu(lambda: mc.bicgstab(bc, None, 0.0, 0.0, 50), reps=2, warm=1)
        emit(what="bicgstab_window", kernel=mc.info()["kernel"], us_per_iter=1e3 * ms / 50)
        ms = time_gpu(lambda: mc.gmres(bc, None, 0.0, 0.0, 30, 2, 0), reps=2, warm=1)
        emit(what="gmres30_cycle", ms_per_cycle=ms / 2)


if __name__ == "__main__":
    main()
