#!/usr/bin/env python3
"""GPU tuning sweep (not part of the product): times the hot kernels and whole CG iterations under the library's
tunables and writes JSON lines to gpurun_out/tune.jsonl.  Usage: python tools/tune.py [--n 256] [--quick]"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"))
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from pytorch_sparse_solver import _native, problems  # noqa: E402

OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
LOG = open(OUT / "tune.jsonl", "a")


def emit(**kw):
    s = json.dumps(kw)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def time_gpu(fn, reps=10, warm=3):
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
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--window", type=int, default=100)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n = args.n
    N = n ** 3
    emit(what="device", name=torch.cuda.get_device_name(0), sms=torch.cuda.get_device_properties(0).multi_processor_count,
         l2=torch.cuda.get_device_properties(0).L2_cache_size)

    # reference points: copy and read bandwidth with torch
    a = torch.empty(1 << 28, dtype=torch.float64, device=dev)   # 2 GiB
    bb = torch.empty_like(a)
    ms = time_gpu(lambda: bb.copy_(a), reps=5)
    emit(what="torch_copy", gbs=2 * a.numel() * 8 / ms / 1e6)
    ms = time_gpu(lambda: a.sum(), reps=5)
    emit(what="torch_sum_read", gbs=a.numel() * 8 / ms / 1e6)
    del a, bb

    A = problems.poisson3d_csr(n, device=dev)
    nnz = A.values().numel()
    m = _native.register_matrix(A)
    h = m.handle
    emit(what="matrix", **m.info())
    x = torch.randn(N, dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    b = torch.ones(N, dtype=torch.float64, device=dev)
    bytes_spmv = nnz * 12 + (N + 1) * 4 + 2 * N * 8
    bytes_iter = problems.cg_bytes_per_iteration(N, nnz)

    # our dot / axpby (2N read / 3N) as vector-kernel bandwidth probes
    for gm in (2, 3, 4, 6):
        h.set_option("grid_mult_vec", gm)
        ms = time_gpu(lambda: _native.dot(x, b), reps=10)
        emit(what="bk_dot", grid_mult_vec=gm, gbs=2 * N * 8 / ms / 1e6, ms=ms)
        ms = time_gpu(lambda: _native.axpby(1.0, x, 2.0, b, out=y), reps=10)
        emit(what="bk_axpby", grid_mult_vec=gm, gbs=3 * N * 8 / ms / 1e6, ms=ms)
    h.set_option("grid_mult_vec", 3)

    for ctas, st in ((2, 0), (2, 2), (2, 3), (3, 0), (3, 2), (4, 0)):
        h.set_option("tma_ctas", ctas)
        h.set_option("tma_stages", st)
        ms = time_gpu(lambda: m.spmv_dot(x, x), reps=10)
        emit(what="bk_spmv_dot_tma", kernel=m.info()["kernel"], tma_ctas=ctas, tma_stages=st, gbs=bytes_spmv / ms / 1e6, ms=ms)
    h.set_option("tma_ctas", 4)
    h.set_option("tma_stages", 0)
    for pf in (0, 1):
        h.set_option("prefetch_x", pf)
        ms = time_gpu(lambda: m.spmv_dot(x, x), reps=10)
        emit(what="bk_spmv_dot_prefetch", prefetch_x=pf, kernel=m.info()["kernel"], gbs_algorithmic=bytes_spmv / ms / 1e6, ms=ms)
    for cmp_ in (0, 1, 2):
        h.set_option("use_compress", cmp_)
        _native.clear_cache()
        mm = _native.register_matrix(A)
        for ctas in (2, 3, 4):
            h.set_option("tma_ctas", ctas)
            ms = time_gpu(lambda: mm.spmv_dot(x, x), reps=10)
            emit(what="bk_spmv_dot_compress", kernel=mm.info()["kernel"], use_compress=cmp_, tma_ctas=ctas,
                 gbs_algorithmic=bytes_spmv / ms / 1e6, ms=ms)
        if mm.info()["kernel"] == 5:
            for ctas in (3, 4, 5, 6):
                h.set_option("pair_ctas", ctas)
                for st in (0, 4):
                    h.set_option("tma_stages", st)
                    ms = time_gpu(lambda: mm.spmv_dot(x, x), reps=10)
                    emit(what="bk_spmv_dot_pair", kernel=5, pair_ctas=ctas, tma_stages=st,
                         gbs_algorithmic=bytes_spmv / ms / 1e6, ms=ms)
            h.set_option("pair_ctas", 4)
            h.set_option("tma_stages", 0)
    h.set_option("tma_ctas", 4)
    h.set_option("use_compress", 2)
    _native.clear_cache()
    m = _native.register_matrix(A)
    for gm in (() if m.info()["kernel"] == 2 else (2, 3, 4, 6, 8)):
        h.set_option("grid_mult_spmv", gm)
        ms = time_gpu(lambda: m.spmv(x, out=y), reps=10)
        emit(what="bk_spmv", grid_mult_spmv=gm, gbs=bytes_spmv / ms / 1e6, ms=ms)
        ms = time_gpu(lambda: m.spmv_dot(x, x), reps=10)
        emit(what="bk_spmv_dot", grid_mult_spmv=gm, gbs=bytes_spmv / ms / 1e6, ms=ms)
    h.set_option("grid_mult_spmv", 4)

    # whole iterations: fixed window of CG iterations
    W = args.window

    def cg_window():
        return m.cg(b, None, 0.0, 0.0, W)

    combos = [dict(), dict(prefetch_x=0), dict(snake=0), dict(snake=0, prefetch_x=0), dict(loop_mode=1), dict(grid_mult_vec=2), dict(grid_mult_vec=4),
              dict(tma_ctas=3), dict(tma_ctas=3, snake=1), dict(chunk=16), dict(fuse_xpay=1)]
    if args.quick:
        combos = combos[:6]
    defaults = {k: h.get_option(k) for k in ("fuse_xpay", "snake", "loop_mode", "grid_mult_vec", "grid_mult_spmv", "tma_ctas", "tma_stages", "chunk", "prefetch_x")}
    for c in combos:
        try:
            for k, v in defaults.items():
                h.set_option(k, v)
            for k, v in c.items():
                h.set_option(k, v)
            ms = time_gpu(cg_window, reps=3, warm=1)
            emit(what="cg_window", opts=c, us_per_iter=1e3 * ms / W, it_s=W / ms * 1e3, gbs=bytes_iter * W / ms / 1e6,
                 frac_8tbs=bytes_iter * W / ms / 1e6 / 8000.0)
        except Exception as e:  # keep sweeping
            emit(what="cg_window", opts=c, error=str(e))
    for k, v in defaults.items():
        h.set_option(k, v)

    # full solve
    t0 = time.perf_counter()
    xs, res = m.cg(b, None, 1e-8, 0.0, None)
    torch.cuda.synchronize()
    emit(what="cg_full", seconds=time.perf_counter() - t0, **res)

    if not args.quick:
        C = problems.convdiff3d_csr(n, device=dev)
        mc = _native.register_matrix(C)
        bc, xt = problems.manufactured_rhs(C, 0)
        ms = time_gpu(lambda: mc.bicgstab(bc, None, 0.0, 0.0, 50), reps=2, warm=1)
        bi = problems.bicgstab_bytes_per_iteration(N, nnz)
        emit(what="bicgstab_window", us_per_iter=1e3 * ms / 50, it_s=50 / ms * 1e3, gbs=bi * 50 / ms / 1e6)
        t0 = time.perf_counter()
        xs, res = mc.bicgstab(bc, None, 1e-8, 0.0, None)
        emit(what="bicgstab_full", seconds=time.perf_counter() - t0, **res)
        ms = time_gpu(lambda: mc.gmres(bc, None, 0.0, 0.0, 30, 2, 0), reps=2, warm=1)
        gb = problems.gmres_bytes_per_cycle(N, nnz, 30)
        emit(what="gmres30_cycle", ms_per_cycle=ms / 2, gbs=gb * 2 / ms / 1e6)


if __name__ == "__main__":
    main()
