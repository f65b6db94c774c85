#!/usr/bin/env python3
"""bench.py — headline benchmark of the Module A Krylov hot path (contract: see the task statement / DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 256]

Metric (BASELINE.json): CG iterations/s, fp64, 7-point Poisson n^3 (n=256 on one GPU), b = ones, tol = 1e-8.
A "step" is ONE full CG solve of that system (611 iterations); value = iterations done in the K timed solves /
their device time (CUDA events), with the matrix and b already resident in HBM.  Inputs (2.9 GB touched per
iteration) are far larger than L2, so no explicit L2 flush is needed between steps.

Extra objects on the JSON line:
  roofline      the dominant kernel (SpMV fused with p.Ap): ALGORITHMIC (CSR) bytes per launch / its mean launch time,
                timed live with CUDA events in a loop of back-to-back launches; peak = MEASURED_PEAKS.json hbm_gbs;
                traffic = DRAM bytes per launch from the committed ncu capture (the coded SpMV kernels move far fewer
                bytes than the CSR count, so frac can exceed 1).
  iteration     the whole CG iteration: algorithmic bytes per iteration * value, vs the same peak.
  e2e           the same solve through the host-buffer C-ABI entry (bk_solve_host): pinned HOST CSR arrays and b are
                copied H2D, solved, x copied D2H, all inside the timed region.
  cpu_baseline  the oracle port of the reference (torch CPU, all host threads) on a bounded fixed-iteration window
                of the same system.
With --impl reference the oracle port itself is the thing timed (rank 0 only).
For N > 1 (torchrun, one rank per GPU) the matrix is row-partitioned into N slabs of n^3 rows each (weak scaling;
BASELINE configs[4] geometry: 2n x 2n planes, n/4 planes per GPU, so N = 8 is exactly the (2n)^3 system), halos and
the dot-product all-reduces go through CUDA-IPC peer memory over NVLink (NCCL with BK_DIST_P2P=0); value = N * global
iterations/s, i.e. "n^3-row CG iterations per second" summed over ranks.  --strong splits the SAME (2n)^3 system.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG_DIR = ROOT / "pytorch-sparse-linalg-torch-amgx.cg.bicg.gmres_b200"
for p in (str(PKG_DIR), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

KERNEL_NAMES = {0: "bk_spmv_stream_kernel", 1: "bk_spmv_vector_kernel", 2: "bk_spmv_tma_kernel<int32 columns>",
                3: "bk_spmv_tma_kernel<8-bit dictionary-coded columns>", 4: "row-split view + bk_vrow_reduce_kernel",
                5: "bk_spmv_pair_kernel<8-bit (offset,value) pair codes, SELL-32-4>"}

METRIC = "cg_iterations_per_second"
UNIT = "it/s"
FALLBACK_HBM_GBS = 6650.0


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def oracle_window(A_cpu, b_cpu, target_s=15.0):
    """Time the oracle port (torch CPU, all threads) on a fixed-iteration CG window of the same system."""
    from oracle import krylov_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t0 = time.perf_counter()
    orc.cg(A_cpu, b_cpu, tol=0.0, atol=0.0, maxiter=2)
    t2 = time.perf_counter() - t0          # 2 iterations + 2 extra matvecs
    per_it = max(t2 / 4.0, 1e-6)
    iters = int(min(max(target_s / per_it, 5), 200))
    t0 = time.perf_counter()
    _x, _info, st = orc.cg(A_cpu, b_cpu, tol=0.0, atol=0.0, maxiter=iters)
    dt = time.perf_counter() - t0
    return st["iterations"] / dt, cores, iters, dt


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU path (its oracle port: same torch-CPU primitives in the same order,
    pinned bit-exact to the reference in oracle/pin_reference.py), all host threads, rank 0 only."""
    if rank != 0:
        return
    from pytorch_sparse_solver import problems
    n = args.n
    A = problems.poisson3d_csr(n)
    b = torch.ones(A.shape[0], dtype=torch.float64)
    from oracle import krylov_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    window = args.ref_window
    for _ in range(args.warmup):
        orc.cg(A, b, tol=0.0, atol=0.0, maxiter=2)
    t0 = time.perf_counter()
    its = 0
    for _ in range(args.steps):
        _x, _i, st = orc.cg(A, b, tol=0.0, atol=0.0, maxiter=window)
        its += st["iterations"]
    dt = time.perf_counter() - t0
    value = its / dt
    sample = f"{args.steps} x fixed window of {window} CG iterations (tol=0) on P3D-{n}, torch CPU, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"CG fp64, 7-pt Poisson {n}^3 CSR (int64 idx), b=ones, oracle port of reference Module A"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_single(args):
    from pytorch_sparse_solver import _native, problems
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n = args.n
    N = n ** 3
    A = problems.poisson3d_csr(n, device=dev)                      # torch CSR, int64 indices, fp64 values
    nnz = A.values().numel()
    b = torch.ones(N, dtype=torch.float64, device=dev)
    m = _native.register_matrix(A)
    h = m.handle
    bytes_iter = problems.cg_bytes_per_iteration(N, nnz)
    bytes_k1 = nnz * 12 + (N + 1) * 4 + 2 * N * 8                  # matrix + read p + write Ap
    peak, peak_kind = measured_peak()

    for _ in range(max(args.warmup, 3)):
        m.cg(b, None, args.tol, 0.0, None)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    its = launches = 0
    with ClockSampler(0) as clk:
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(args.steps):
            x, res = m.cg(b, None, args.tol, 0.0, None)
            its += res["iterations"]
            launches += res["kernel_launches"]
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    value = its / (ms * 1e-3)
    last = res

    # dominant kernel, timed alone (back-to-back launches, CUDA events on the launching stream)
    p = torch.randn(N, dtype=torch.float64, device=dev)
    for _ in range(3):
        m.spmv_dot(p, p)
    reps = 20
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        m.spmv_dot(p, p)
    ev1.record()
    torch.cuda.synchronize()
    k1_ms = ev0.elapsed_time(ev1) / reps
    achieved = bytes_k1 / (k1_ms * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "roofline_traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get("spmv_dot_dram_bytes_per_launch")
        except Exception:
            traffic = None

    # end to end through the host-buffer C-ABI entry: pinned host arrays in, x out
    e2e = None
    if not args.no_e2e:
        crow_h = A.crow_indices().cpu().pin_memory()
        col_h = A.col_indices().cpu().pin_memory()
        val_h = A.values().cpu().pin_memory()
        b_h = b.cpu().pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in (crow_h, col_h, val_h, b_h))
        d2h = N * 8
        _native.solve_host(_native.METHOD_CG, crow_h, col_h, val_h, b_h, None, args.tol, 0.0, None)
        torch.cuda.synchronize()
        k_e2e = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        its_e = 0
        for _ in range(k_e2e):
            xh, r = _native.solve_host(_native.METHOD_CG, crow_h, col_h, val_h, b_h, None, args.tol, 0.0, None)
            its_e += r["iterations"]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": its_e / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * dt / k_e2e, "steps": k_e2e}
        del crow_h, col_h, val_h

    cpu = None
    if not args.no_cpu:
        A_cpu = torch.sparse_csr_tensor(A.crow_indices().cpu(), A.col_indices().cpu(), A.values().cpu(), size=A.shape)
        v, cores, iters, dt = oracle_window(A_cpu, b.cpu(), args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"fixed window of {iters} CG iterations (tol=0) of the same P3D-{n} system, torch CPU, "
                         f"{cores} threads, {dt:.1f} s"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"CG fp64, 7-pt Poisson {n}^3 CSR, b=ones, tol={args.tol:g} (BASELINE configs[1])",
                   "n": N, "nnz": nnz, "iterations_per_solve": last["iterations"], "info": last["info"],
                   "relres": last["final_residual"] / last["b_norm"],
                   "l2_policy": "inputs (2.9 GB/iteration) exceed L2; no flush needed",
                   "spmv_kernel": m.info()["kernel"],
                   "options": {k: h.get_option(k) for k in ("use_tma", "use_compress", "tma_ctas", "pair_ctas", "tma_stages", "grid_mult_spmv", "grid_mult_vec",
                                                            "fuse_xpay", "snake", "loop_mode", "chunk")}},
        "roofline": {"bound": "hbm", "kernel": KERNEL_NAMES.get(m.info()["kernel"], "?") + " (CSR SpMV fused with p.Ap)",
                     "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "bytes_per_launch": bytes_k1, "ms_per_launch": k1_ms,
                     "traffic_gbs": (traffic / (k1_ms * 1e-3) / 1e9) if traffic else None,
                     "note": "achieved = ALGORITHMIC CSR bytes (nnz*12 + (n+1)*4 + 2n*8) / time; the coded kernels "
                             "(3, 5) are lossless re-encodings that move fewer bytes than that, so frac can exceed 1 "
                             "- `traffic` is the DRAM bytes ncu measured per launch"},
        "iteration": {"bytes_per_iteration": bytes_iter, "achieved_gbs": bytes_iter * value / 1e9,
                      "frac_of_peak": bytes_iter * value / 1e9 / peak, "frac_of_8tbs": bytes_iter * value / 8e12,
                      "us_per_iteration": 1e3 * ms / its},
        "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": int(launches), "clocks": clk.summary(),
    }
    print(json.dumps(line))


def run_dist(args, rank, world):
    import torch.distributed as dist
    from pytorch_sparse_solver import distributed as bkd
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = bkd.bench_weak_scaling(args, rank, world, local, METRIC, UNIT, measured_peak(), ClockSampler)
    if rank == 0 and line is not None:
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=256, help="grid edge (n^3 rows per GPU)")
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--dist-window", type=int, default=200, help="N>1: CG iterations per timed step")
    ap.add_argument("--strong", action="store_true", help="N>1: split the SAME (2n)^3 system over N GPUs (strong scaling)")
    ap.add_argument("--ref-window", type=int, default=10, help="--impl reference: CG iterations per step")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1 or args.gpus > 1 or args.strong or os.environ.get("BK_BENCH_FORCE_DIST"):  # env: 1-rank run of the multi-GPU path
        if "RANK" not in os.environ:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
        run_dist(args, rank, world)
    else:
        run_single(args)


if __name__ == "__main__":
    main()
