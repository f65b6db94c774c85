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
    ext_rowptr: torch.Tensor   # int32 [n_local+1]: the rows as ONE matrix over [local | ghost] (original entry order)
    ext_col: torch.Tensor      # int32: local column, or n_local + ghost index


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
    ext_col = torch.where(is_loc, col - row_begin, n_local + torch.searchsorted(ghost_ids, col)).to(torch.int32)
    return LocalSplit(n_local, loc_rowptr.to(torch.int32), loc_col, loc_val, brow_ids.to(torch.int32),
                      gh_rowptr.to(torch.int32), gh_col, gh_val, ghost_ids, crow.to(torch.int32), ext_col)


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

    def pack(t):  # a contiguous run of ids (the usual case: whole grid planes) travels as (first, count), not as 2 MB
        n = int(t.numel())
        if n > 0 and int(t[-1]) - int(t[0]) == n - 1:
            return ("range", int(t[0]), n)
        return ("ids", t)

    def unpack(m):
        if m[0] == "range":
            return torch.arange(m[1], m[1] + m[2], dtype=torch.int64)
        return m[1]
    if world > 1:
        wire = [None] * world
        dist.all_gather_object(wire, {o: pack(t) for o, t in needs.items()}, group=group)
        gathered = [{o: unpack(m) for o, m in w.items()} for w in wire]
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


def transpose_slab(crow: torch.Tensor, col: torch.Tensor, val: torch.Tensor, offsets: Sequence[int], rank: int,
                   world: int, group=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """This rank's slab of A^T (same row partition), as CSR arrays with GLOBAL columns, from this rank's slab of A:
    every entry (r, c, v) goes to the owner of column c (torch.distributed all_to_all_single; set-up only), which
    sorts what it receives by (c, r) — rows of A^T ascending, columns ascending inside a row.  Device-agnostic
    (exercised with gloo on CPU by tests/test_dist_gloo.py)."""
    import torch.distributed as dist
    dev = col.device
    off = torch.tensor(list(offsets), dtype=torch.int64, device=dev)
    n_local = int(offsets[rank + 1] - offsets[rank])
    n_global = int(offsets[-1])
    lens = (crow[1:] - crow[:-1]).long()
    grow = torch.repeat_interleave(torch.arange(n_local, device=dev), lens) + int(offsets[rank])
    gcol = col.long()
    owner = torch.bucketize(gcol, off[1:], right=True)
    order = torch.argsort(owner, stable=True)
    send_counts = torch.bincount(owner, minlength=world)
    if world > 1:
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=group)
        sc, rc = send_counts.tolist(), recv_counts.tolist()

        def exchange(t):
            out = torch.empty(sum(rc), dtype=t.dtype, device=dev)
            dist.all_to_all_single(out, t[order].contiguous(), rc, sc, group=group)
            return out
        t_row, t_col, t_val = exchange(gcol), exchange(grow), exchange(val)   # transposed: (col, row, val)
    else:
        t_row, t_col, t_val = gcol, grow, val
    lrow = t_row - int(offsets[rank])
    perm = torch.argsort(lrow * n_global + t_col, stable=True)
    lrow, t_col, t_val = lrow[perm], t_col[perm].contiguous(), t_val[perm].contiguous()
    tcrow = torch.zeros(n_local + 1, dtype=torch.int64, device=dev)
    tcrow[1:] = torch.bincount(lrow, minlength=n_local).cumsum(0)
    return tcrow, t_col, t_val


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
        import os
        import time
        timing = os.environ.get("BK_DIST_TIMING", "0") != "0"   # per-phase set-up times (adds device syncs)
        self.setup_ms = {}
        t_last = [time.perf_counter()]

        def mark(name):
            if timing:
                torch.cuda.synchronize(self.device)
                now = time.perf_counter()
                self.setup_ms[name] = round(1e3 * (now - t_last[0]), 2)
                t_last[0] = now
        sp = split_local_ghost(crow, col, val, offsets[rank], offsets[rank + 1])
        mark("split_local_ghost")
        plan = build_halo_plan(sp.ghost_ids, offsets, rank, world, group)
        mark("build_halo_plan")
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
        mark("bk_dist_create")
        self.p2p = False
        self.folded = False
        self.fold_kernel = 0
        self._group = group
        self._src = (crow, col, val)      # the caller's slab (global columns): needed to build the transpose
        self._transpose = None
        self._diag = None
        if os.environ.get("BK_DIST_P2P", "1") != "0":
            self._connect_peer_memory(plan, group)
        mark("connect_peer_memory")
        if self.p2p and os.environ.get("BK_DIST_FOLD", "1") != "0":
            # the rows as one matrix over [local | ghost]: one SpMV kernel per matvec when the row-bitmask plan fits
            folded = C.c_int32(0)
            vals = val.contiguous()
            gid = sp.ghost_ids.to(torch.int64).contiguous()
            with torch.cuda.device(self.device):
                _native._check(lib.bk_dist_set_extended(
                    self.ptr, vals.numel(), sp.ext_rowptr.data_ptr(), sp.ext_col.data_ptr(), vals.data_ptr(),
                    gid.data_ptr() if gid.numel() else None, int(offsets[rank]), _native._stream_ptr(self.device),
                    C.byref(folded)), "bk_dist_set_extended")
            self.folded = bool(folded.value)
            self.fold_kernel = 7 if folded.value == 2 else (6 if folded.value else 0)
        mark("set_extended")
        sp.ext_col = None                 # only read during registration
        sp.ext_rowptr = None

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
        if getattr(self, "_transpose", None) is not None:
            self._transpose.close()
            self._transpose = None
        if getattr(self, "ptr", None) is not None:
            self.handle.lib.bk_dist_destroy(self.ptr)
            self.ptr = None

    def local_info(self) -> dict:
        return {"folded_single_kernel_spmv": self.folded, "folded_spmv_kernel": self.fold_kernel, "p2p": self.p2p,
                "n_local": self.split.n_local,
                "n_ghost": int(self.split.ghost_ids.numel()), "n_boundary_rows": int(self.split.brow_ids.numel())}

    def diagonal(self) -> torch.Tensor:
        """This rank's slice of diag(A) (the diagonal block of a row partition is local)."""
        if self._diag is None:
            sp = self.split
            n = sp.n_local
            rows = torch.repeat_interleave(torch.arange(n, device=self.device),
                                           (sp.loc_rowptr[1:] - sp.loc_rowptr[:-1]).long())
            m = rows == sp.loc_col.long()
            d = torch.zeros(n, dtype=self.dtype, device=self.device)
            d.index_add_(0, rows[m], sp.loc_val[m])
            self._diag = d
        return self._diag

    def transpose(self) -> "DistMatrix":
        """A^T with the same row partition (cached): every rank sends each peer the entries of its rows whose columns
        that peer owns (torch.distributed all_to_all, set-up only); the owner sorts them by (row, col) into its slab
        of A^T.  Used by the implicit-differentiation backward (reference :1237-1248 solves with A^T)."""
        if self._transpose is not None:
            return self._transpose
        tcrow, t_col, t_val = transpose_slab(*self._src, self.offsets, self.rank, self.world, self._group)
        self._transpose = DistMatrix(tcrow, t_col, t_val, self.offsets, self.rank, self.world, self._group)
        return self._transpose

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
    def _solve_local(self, name, bw, x0w, tol, atol, maxiter, restart, solve_method, diag):
        from .module_a.krylov import _gmres_effective_tolerances
        lib = self.handle.lib
        b, x, has = self._vectors(bw, x0w)
        res = _native.bk_result()
        mi = -1 if maxiter is None else int(maxiter)
        with torch.cuda.device(self.device):
            s = _native._stream_ptr(self.device)
            if name in ("cg", "bicgstab"):
                if diag is None:
                    fn = lib.bk_dist_cg if name == "cg" else lib.bk_dist_bicgstab
                    rc = fn(self.handle.ptr, self.ptr, b.data_ptr(), x.data_ptr(), has, float(tol), float(atol), mi,
                            self.n_global, C.byref(res), s)
                else:
                    fn = lib.bk_dist_cg_jacobi if name == "cg" else lib.bk_dist_bicgstab_jacobi
                    rc = fn(self.handle.ptr, self.ptr, diag.data_ptr(), b.data_ptr(), x.data_ptr(), has, float(tol),
                            float(atol), mi, self.n_global, C.byref(res), s)
            else:
                method = 1 if solve_method == 'incremental' else 0
                rst = min(int(restart), self.n_global)
                te, ae = _gmres_effective_tolerances(tol, atol, self.n_global, 'cuda')
                if diag is None:
                    rc = lib.bk_dist_gmres(self.handle.ptr, self.ptr, b.data_ptr(), x.data_ptr(), has, float(te),
                                           float(ae), rst, mi, method, self.n_global, C.byref(res), s)
                else:
                    rc = lib.bk_dist_gmres_jacobi(self.handle.ptr, self.ptr, diag.data_ptr(), b.data_ptr(), x.data_ptr(),
                                                  has, float(te), float(ae), rst, mi, method, self.n_global,
                                                  C.byref(res), s)
        _native._check(rc, f"bk_dist_{name}")
        return x, res.as_dict()

    # ---- the reference API on a row-partitioned matrix ("API stays additive, e.g. a DistCSR wrapper
    # passed as A"): module_a.cg / bicgstab / gmres(A=DistMatrix, b=this rank's slab) -> (x_local, info), with the
    # built-in Jacobi preconditioner (M = module_a.JacobiPreconditioner(D)) and the implicit-diff backward for b --------
    def _module_a_solve(self, name: str, b, x0=None, *, tol=1e-5, atol=0.0, maxiter=None, M=None, restart=20,
                        solve_method='batched', _result=None):
        from .module_a import krylov
        from .module_a.preconditioners import JacobiPreconditioner
        if not isinstance(b, torch.Tensor) or b.ndim != 1 or b.shape[0] != self.split.n_local:
            raise ValueError(f"b must be this rank's slab: a vector of length {self.split.n_local}")
        if x0 is not None and tuple(x0.shape) != tuple(b.shape):
            raise ValueError(f'arrays in x0 and b must have matching shapes: {x0.shape} vs {b.shape}')
        if name == "gmres" and restart < 1:
            raise ValueError("restart must be >= 1")
        diag = None
        if M is not None:
            if not isinstance(M, JacobiPreconditioner):
                raise NotImplementedError("a DistMatrix takes the built-in JacobiPreconditioner (or M=None)")
            diag = M.diagonal(self.dtype, self.device).contiguous()
        with torch.no_grad():
            x, res = self._solve_local(name, b.detach(), None if x0 is None else x0.detach(), tol, atol, maxiter,
                                       restart, solve_method, diag)
        krylov._publish(dict(res, solver=name, route="dist"), _result)
        if b.requires_grad:
            x = _DistAdjoint.apply(b, x, self, name, x0, tol, atol, maxiter, restart, solve_method, diag)
        return x, int(res["info"])


class _DistAdjoint(torch.autograd.Function):
    """grad_b = solve(A^T, grad_x) on the row-partitioned transpose (reference ImplicitAdjointFunction :1227-1248:
    same x0 / tolerances, M re-used, no gradient for A)."""

    @staticmethod
    def forward(ctx, b, x, D, name, x0, tol, atol, maxiter, restart, solve_method, diag):
        ctx.D = D
        ctx.meta = (name, x0, tol, atol, maxiter, restart, solve_method, diag)
        return x.clone()

    @staticmethod
    def backward(ctx, grad_output):
        name, x0, tol, atol, maxiter, restart, solve_method, diag = ctx.meta
        with torch.no_grad():
            Dt = ctx.D.transpose()          # collective: every rank's backward reaches this point
            g, _ = Dt._solve_local(name, grad_output.detach().contiguous(), None if x0 is None else x0.detach(), tol,
                                   atol, maxiter, restart, solve_method, diag)
        return (g.to(grad_output.dtype),) + (None,) * 10
