"""Row-partitioned (multi-GPU) Module A: set-up logic + binding of the bk_dist_* C entries.

One process per GPU (torchrun); `torch.distributed` is used for rendezvous and for the set-up exchange of index
lists only — the per-iteration halo exchange and scalar all-reduces are issued by the C library on NCCL.

Set-up (device-agnostic torch index arithmetic, exercised on CPU by tests/test_dist_gloo.py with world_size 2):
  partition_rows      contiguous 1-D row partition
  split_local_ghost   local rows (CSR, global columns) -> LOCAL block (own columns, renumbered) + GHOST block
                      (CSR over the boundary rows, columns renumbered into a compact ghost vector ordered by owner)
  build_halo_plan     who sends which of its entries to whom (the ParCSR / VecScatter pattern)
The reference has no distributed code at all; nothing here mirrors a reference file.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _native


def partition_rows(n_global: int, world: int) -> List[int]:
    """Row offsets [o_0=0, ..., o_world=n_global] of a contiguous, balanced 1-D partition."""
    return [(n_global * r) // world for r in range(world + 1)]


@dataclass
class LocalSplit:
    n_local: int
    loc_rowptr: torch.Tensor   # int32 [n_local+1]
    loc_col: torch.Tensor      # int32, local column ids
    loc_val: torch.Tensor
    brow_ids: torch.Tensor     # int32 [n_brows] local row ids that own ghost entries (ascending)
    gh_rowptr: torch.Tensor    # int32 [n_brows+1]
    gh_col: torch.Tensor       # int32, index into the ghost vector
    gh_val: torch.Tensor
    ghost_ids: torch.Tensor    # int64 [n_ghost] global ids of the ghost vector entries (ascending => grouped by owner)


def split_local_ghost(crow: torch.Tensor, col: torch.Tensor, val: torch.Tensor, row_begin: int,
                      row_end: int) -> LocalSplit:
    n_local = row_end - row_begin
    dev = col.device
    crow = crow.to(torch.int64)
    col = col.to(torch.int64)
    lens = crow[1:] - crow[:-1]
    row_of = torch.repeat_interleave(torch.arange(n_local, device=dev), lens)
    is_loc = (col >= row_begin) & (col < row_end)
    # local block
    loc_counts = torch.zeros(n_local, dtype=torch.int64, device=dev)
    loc_counts.index_add_(0, row_of, is_loc.to(torch.int64))
    loc_rowptr = torch.zeros(n_local + 1, dtype=torch.int64, device=dev)
    loc_rowptr[1:] = loc_counts.cumsum(0)
    loc_col = (col[is_loc] - row_begin).to(torch.int32)
    loc_val = val[is_loc].contiguous()
    # ghost block
    gmask = ~is_loc
    gcols = col[gmask]
    ghost_ids = torch.unique(gcols)  # sorted ascending
    gh_col = torch.searchsorted(ghost_ids, gcols).to(torch.int32)
    gh_val = val[gmask].contiguous()
    gh_rows = row_of[gmask]
    if gh_rows.numel():
        brow_ids, counts = torch.unique_consecutive(gh_rows, return_counts=True)
    else:
        brow_ids = torch.zeros(0, dtype=torch.int64, device=dev)
        counts = torch.zeros(0, dtype=torch.int64, device=dev)
    gh_rowptr = torch.zeros(brow_ids.numel() + 1, dtype=torch.int64, device=dev)
    gh_rowptr[1:] = counts.cumsum(0)
    return LocalSplit(n_local, loc_rowptr.to(torch.int32), loc_col, loc_val, brow_ids.to(torch.int32),
                      gh_rowptr.to(torch.int32), gh_col, gh_val, ghost_ids)


@dataclass
class HaloPlan:
    peers: List[int]             # ranks exchanged with, ascending
    send_counts: List[int]       # entries sent to each peer
    recv_counts: List[int]       # entries received from each peer (ghost vector is the concatenation, in peer order)
    send_idx: torch.Tensor       # int32: local indices to send, concatenated per peer


def build_halo_plan(ghost_ids: torch.Tensor, offsets: Sequence[int], rank: int, world: int, group=None) -> HaloPlan:
    """Exchange 'which of your entries I need' with every rank (torch.distributed all_gather_object; set-up only)."""
    import torch.distributed as dist
    off = torch.tensor(list(offsets[1:]), dtype=torch.int64)
    gids = ghost_ids.cpu()
    owner = torch.bucketize(gids, off, right=True)
    needs = {}
    for o in torch.unique(owner).tolist():
        needs[int(o)] = gids[owner == o]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, needs, group=group)
    else:
        gathered = [needs]
    peers = sorted(set(needs.keys()) | {q for q in range(world) if q != rank and rank in gathered[q]})
    send_counts, recv_counts, send_lists = [], [], []
    for q in peers:
        req = gathered[q].get(rank) if q != rank else None
        if req is None:
            req = torch.zeros(0, dtype=torch.int64)
        send_lists.append((req - offsets[rank]).to(torch.int32))
        send_counts.append(int(req.numel()))
        recv_counts.append(int(needs[q].numel()) if q in needs else 0)
    send_idx = torch.cat(send_lists) if send_lists else torch.zeros(0, dtype=torch.int32)
    return HaloPlan(peers, send_counts, recv_counts, send_idx)


class DistMatrix:
    """A row-partitioned matrix registered with the library on this rank's GPU."""

    def __init__(self, crow: torch.Tensor, col: torch.Tensor, val: torch.Tensor, offsets: Sequence[int], rank: int,
                 world: int, group=None):
        import torch.distributed as dist
        if not val.is_cuda:
            raise _native.NativeLibraryError("DistMatrix needs CUDA tensors (one GPU per rank)")
        self.rank, self.world = rank, world
        self.offsets = list(offsets)
        self.n_global = int(offsets[-1])
        self.device = val.device
        self.dtype = val.dtype
        sp = split_local_ghost(crow, col, val, offsets[rank], offsets[rank + 1])
        plan = build_halo_plan(sp.ghost_ids, offsets, rank, world, group)
        self.split, self.plan = sp, plan
        self.handle = _native.Handle.get(self.device)
        lib = self.handle.lib
        # NCCL unique id: rank 0 creates, everybody receives
        idbuf = (C.c_char * 128)()
        if rank == 0:
            _native._check(lib.bk_dist_unique_id(idbuf), "bk_dist_unique_id")
        if world > 1:
            box = [bytes(idbuf.raw)]
            dist.broadcast_object_list(box, src=0, group=group)
            idbuf = (C.c_char * 128).from_buffer_copy(box[0])
        npeers = len(plan.peers)
        peers = (C.c_int32 * max(npeers, 1))(*plan.peers)
        sc = (C.c_int64 * max(npeers, 1))(*plan.send_counts)
        rc = (C.c_int64 * max(npeers, 1))(*plan.recv_counts)
        self._send_idx = plan.send_idx.to(self.device)
        self._keep = (sp.loc_rowptr, sp.loc_col, sp.loc_val, sp.brow_ids, sp.gh_rowptr, sp.gh_col, sp.gh_val,
                      self._send_idx)
        p = C.c_void_p()
        with torch.cuda.device(self.device):
            _native._check(lib.bk_dist_create(
                self.handle.ptr, idbuf, rank, world, sp.n_local, sp.loc_val.numel(), sp.loc_rowptr.data_ptr(),
                sp.loc_col.data_ptr(), sp.loc_val.data_ptr(), sp.brow_ids.numel(), sp.brow_ids.data_ptr(),
                sp.gh_val.numel(), sp.gh_rowptr.data_ptr(), sp.gh_col.data_ptr(), sp.gh_val.data_ptr(),
                sp.ghost_ids.numel(), npeers, peers, sc, rc, self._send_idx.data_ptr(),
                _native._dtype_code(self.dtype), _native._stream_ptr(self.device), C.byref(p)), "bk_dist_create")
        self.ptr = p
        self.p2p = False
        import os
        if os.environ.get("BK_DIST_P2P", "1") != "0":
            self._connect_peer_memory(plan, group)

    def _connect_peer_memory(self, plan: HaloPlan, group=None):
        """Exchange CUDA-IPC handles of the communication windows and map every rank's window (NVLink peer memory)."""
        import warnings
        import torch.distributed as dist
        lib = self.handle.lib
        hbuf = (C.c_char * 64)()
        try:
            _native._check(lib.bk_dist_p2p_export(self.ptr, hbuf), "bk_dist_p2p_export")
        except _native.NativeLibraryError as e:  # pragma: no cover
            warnings.warn(f"peer-memory path unavailable ({e}); using NCCL")
            return
        recv_off, o = {}, 0
        for q, c in zip(plan.peers, plan.recv_counts):
            recv_off[q] = o
            o += c
        mine = {"handle": bytes(hbuf.raw), "recv_off": recv_off}
        if self.world > 1:
            gathered = [None] * self.world
            dist.all_gather_object(gathered, mine, group=group)
        else:
            gathered = [mine]
        handles = b"".join(g["handle"] for g in gathered)
        roffs = (C.c_int64 * max(len(plan.peers), 1))(*[gathered[q]["recv_off"].get(self.rank, 0) for q in plan.peers])
        rc = lib.bk_dist_p2p_connect(self.ptr, handles, roffs)
        ok = torch.tensor([1 if rc == 0 else 0], device=self.device)
        if self.world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)   # all ranks or none
        if int(ok) == 1:
            self.p2p = True
        else:
            warnings.warn("peer-memory path could not be connected on every rank; using NCCL")
            self.handle.set_option("dist_p2p", 0)

    def close(self):
        if getattr(self, "ptr", None) is not None:
            self.handle.lib.bk_dist_destroy(self.ptr)
            self.ptr = None

    def spmv(self, x_local: torch.Tensor) -> torch.Tensor:
        x = x_local.contiguous()
        y = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _native._check(self.handle.lib.bk_dist_spmv(self.handle.ptr, self.ptr, x.data_ptr(), y.data_ptr(),
                                                        _native._stream_ptr(self.device)), "bk_dist_spmv")
        return y

    def cg(self, b_local: torch.Tensor, x0: Optional[torch.Tensor] = None, tol: float = 1e-5, atol: float = 0.0,
           maxiter: Optional[int] = None) -> Tuple[torch.Tensor, dict]:
        b = b_local.to(self.dtype).contiguous()
        if x0 is None:
            x, has = torch.empty_like(b), 0
        else:
            x, has = x0.to(self.dtype).contiguous().clone(), 1
        res = _native.bk_result()
        with torch.cuda.device(self.device):
            _native._check(self.handle.lib.bk_dist_cg(
                self.handle.ptr, self.ptr, b.data_ptr(), x.data_ptr(), has, float(tol), float(atol),
                -1 if maxiter is None else int(maxiter), self.n_global, C.byref(res),
                _native._stream_ptr(self.device)), "bk_dist_cg")
        return x, res.as_dict()

    def _vectors(self, b_local, x0):
        b = b_local.to(self.dtype).contiguous()
        if x0 is None:
            return b, torch.empty_like(b), 0
        return b, x0.to(self.dtype).contiguous().clone(), 1

    def bicgstab(self, b_local: torch.Tensor, x0: Optional[torch.Tensor] = None, tol: float = 1e-5,
                 atol: float = 0.0, maxiter: Optional[int] = None) -> Tuple[torch.Tensor, dict]:
        """Row-partitioned BiCGStab (reference _bicgstab_solve :859-964 on the global system)."""
        b, x, has = self._vectors(b_local, x0)
        res = _native.bk_result()
        with torch.cuda.device(self.device):
            _native._check(self.handle.lib.bk_dist_bicgstab(
                self.handle.ptr, self.ptr, b.data_ptr(), x.data_ptr(), has, float(tol), float(atol),
                -1 if maxiter is None else int(maxiter), self.n_global, C.byref(res),
                _native._stream_ptr(self.device)), "bk_dist_bicgstab")
        return x, res.as_dict()

    def gmres(self, b_local: torch.Tensor, x0: Optional[torch.Tensor] = None, tol: float = 1e-5, atol: float = 0.0,
              restart: int = 20, maxiter: Optional[int] = None,
              solve_method: str = 'batched') -> Tuple[torch.Tensor, dict]:
        """Row-partitioned restarted GMRES (reference gmres :641-784 on the global system; the tolerance constants
        use the GLOBAL size, as one process holding the whole matrix would)."""
        from .module_a.krylov import _gmres_effective_tolerances
        if solve_method not in ('batched', 'incremental'):
            raise ValueError(f"invalid solve_method {solve_method}, must be either 'incremental' or 'batched'")
        method = 1 if solve_method == 'incremental' else 0
        restart = min(int(restart), self.n_global)
        tol_eff, atol_eff = _gmres_effective_tolerances(tol, atol, self.n_global, 'cuda')
        b, x, has = self._vectors(b_local, x0)
        res = _native.bk_result()
        with torch.cuda.device(self.device):
            _native._check(self.handle.lib.bk_dist_gmres(
                self.handle.ptr, self.ptr, b.data_ptr(), x.data_ptr(), has, float(tol_eff), float(atol_eff),
                restart, -1 if maxiter is None else int(maxiter), method, self.n_global, C.byref(res),
                _native._stream_ptr(self.device)), "bk_dist_gmres")
        return x, res.as_dict()


    # ---- the reference API on a row-partitioned matrix ("API stays additive, e.g. a DistCSR wrapper
    # passed as A"): module_a.cg / bicgstab / gmres(A=DistMatrix, b=this rank's slab) -> (x_local, info) -----------
    def _module_a_solve(self, name: str, b, x0=None, *, tol=1e-5, atol=0.0, maxiter=None, M=None, restart=20,
                        solve_method='batched', _result=None):
        from .module_a import krylov
        if not isinstance(b, torch.Tensor) or b.ndim != 1 or b.shape[0] != self.split.n_local:
            raise ValueError(f"b must be this rank's slab: a vector of length {self.split.n_local}")
        if x0 is not None and tuple(x0.shape) != tuple(b.shape):
            raise ValueError(f'arrays in x0 and b must have matching shapes: {x0.shape} vs {b.shape}')
        if M is not None:
            raise NotImplementedError("preconditioners on a DistMatrix: not wired yet")
        with torch.no_grad():
            bw = b.detach()
            x0w = None if x0 is None else x0.detach()
            if name == "cg":
                x, res = self.cg(bw, x0w, tol, atol, maxiter)
            elif name == "bicgstab":
                x, res = self.bicgstab(bw, x0w, tol, atol, maxiter)
            else:
                if restart < 1:
                    raise ValueError("restart must be >= 1")
                x, res = self.gmres(bw, x0w, tol, atol, restart, maxiter, solve_method)
        krylov._publish(dict(res, solver=name, route="dist"), _result)
        return x, int(res["info"])


# ---- weak-scaling benchmark used by bench.py --gpus N --------------------------------------------------------
def bench_weak_scaling(args, rank, world, local, metric, unit, peak, ClockSampler):
    """N slabs of n^3 rows each: a (N*n) x n x n Poisson grid, slab q on rank q.  value = N * global iterations/s
    (n^3-row CG iterations per second summed over ranks)."""
    import time
    import torch.distributed as dist
    from . import problems
    dev = torch.device("cuda", local)
    n = args.n
    rows = n ** 3
    # BASELINE config 5 geometry: planes of (2n) x (2n), n/4 planes per GPU (n = 256: 64 planes of 512 x 512, so
    # 8 GPUs hold exactly the 512^3 system and every halo is one 512^2 plane = 2 MiB); per-GPU rows stay n^3.
    npl, ppg = 2 * n, max(n // 4, 1)
    if ppg * npl * npl != rows:
        npl, ppg = n, n
    strong = bool(getattr(args, "strong", False))
    if strong:  # BASELINE configs[4] as written: the SAME (2n)^3 system split over N GPUs (N must divide 2n)
        ppg = npl // world
        rows = ppg * npl * npl
    offsets = [q * rows for q in range(world + 1)]
    crow, col, val = problems.stencil3d_rows(npl, world * ppg, rank * ppg, (rank + 1) * ppg, device=dev)
    nnz_local = val.numel()
    D = DistMatrix(crow, col, val, offsets, rank, world)
    del crow, col
    b = torch.ones(rows, dtype=torch.float64, device=dev)
    window = args.dist_window
    for _ in range(max(args.warmup, 3)):
        D.cg(b, None, 0.0, 0.0, min(window, 50))
    torch.cuda.synchronize()
    dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    its = launches = 0
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        dist.barrier()
        ev0.record()
        for _ in range(args.steps):
            x, res = D.cg(b, None, 0.0, 0.0, window)
            its += res["iterations"]
            launches += res["kernel_launches"]
        ev1.record()
        torch.cuda.synchronize()
        dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    it_s = its / (ms * 1e-3)
    value = it_s * (rows * world) / float(n ** 3)   # n^3-row CG iterations per second over all ranks
    bytes_iter = problems.cg_bytes_per_iteration(rows, nnz_local)
    # end to end: host slab -> device, registration, solve window, x back to host
    crow_h, col_h, val_h = problems.stencil3d_rows(npl, world * ppg, rank * ppg, (rank + 1) * ppg)
    crow_h, col_h, val_h, b_h = crow_h.pin_memory(), col_h.pin_memory(), val_h.pin_memory(), b.cpu().pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in (crow_h, col_h, val_h, b_h))
    D.close()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    D2 = DistMatrix(crow_h.to(dev, non_blocking=True), col_h.to(dev, non_blocking=True), val_h.to(dev, non_blocking=True),
                    offsets, rank, world)
    xe, re_ = D2.cg(b_h.to(dev, non_blocking=True), None, 0.0, 0.0, window)
    xh = xe.cpu()
    torch.cuda.synchronize()
    dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e = {"value": world * re_["iterations"] / float(dt), "unit": unit, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": rows * 8, "ms_per_step": 1e3 * float(dt), "steps": 1,
           "note": "includes partition set-up, halo-plan exchange and IPC window mapping (the NCCL communicator of "
                   "the process is reused)"}
    D2.close()
    if rank != 0:
        return None
    pk, pk_kind = peak
    return {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"CG fp64, 7-pt Poisson {world * ppg}x{npl}x{npl} CSR row-partitioned over {world} GPUs "
                               f"({n}^3 rows per GPU; 8 GPUs = BASELINE configs[4] 512^3), b=ones, fixed window of "
                               f"{window} iterations per step",
                   "comm": "peer-memory (CUDA IPC over NVLink): kernel halo push + one-shot all-reduce" if D.p2p
                   else "NCCL send/recv + allreduce",
                   "value_definition": f"{world} x global iterations/s = {n}^3-row CG iterations per second over all ranks",
                   "global_iterations_per_second": it_s, "n_local": rows, "nnz_local": nnz_local,
                   "halo_bytes_per_neighbour": npl * npl * 8, "peers_rank0": D.plan.peers,
                   "l2_policy": "inputs exceed L2; no flush needed"},
        "roofline": {"bound": "hbm", "kernel": "whole distributed CG iteration (per GPU)",
                     "achieved": bytes_iter * it_s / 1e9, "peak": pk, "peak_kind": pk_kind, "unit": "GB/s",
                     "frac": bytes_iter * it_s / 1e9 / pk, "traffic": None},
        "iteration": {"bytes_per_iteration_per_gpu": bytes_iter, "us_per_iteration": 1e3 * ms / its,
                      "frac_of_8tbs": bytes_iter * it_s / 8e12},
        "e2e": e2e, "cpu_baseline": None, "gpu_launches": int(launches), "clocks": clk.summary(),
    }
