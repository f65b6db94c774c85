// bk_dist.cu — life cycle of a row-partitioned (multi-GPU) matrix: NCCL loading and communicator, the communication
// window and its CUDA-IPC mapping between ranks, and the plain distributed SpMV.  One process per GPU; SURVEY §8e.
//
// Each rank owns a contiguous slab of rows and the matching slices of every vector.  Its rows are split (by the
// Python set-up code, pytorch_sparse_solver/distributed.py) into
//   * a LOCAL block  (columns inside the slab, renumbered)  -> a normal bk_csr, run by the same SpMV kernels, and
//   * a GHOST block  (columns owned by other ranks, renumbered into a compact ghost vector) stored as a CSR over the
//     few "boundary rows" that have such entries.
// The kernels and the solver-facing interface live in bk_dist.cuh; the distributed solvers are the single-GPU
// drivers instantiated on it (bk_dist_cg in bk_cg.cu, bk_dist_bicgstab in bk_bicgstab.cu, bk_dist_gmres in
// bk_gmres.cu).  NCCL is dlopen'ed (torch's bundled libnccl.so.2 is already in the process), never linked.
#include <dlfcn.h>
#include <stdlib.h>

#include "bk_internal.cuh"
#include "bk_loop.cuh"
#include "bk_p2p.cuh"
#include "bk_spmv.cuh"
#include "bk_vec.cuh"
#include "bk_dist.cuh"

bk_nccl_api g_nccl = {nullptr};

static int bk_nccl_load() {
  if (g_nccl.lib) return BK_OK;
  const char* names[3] = {getenv("BK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (int i = 0; i < 3 && !lib; ++i)
    if (names[i] && *names[i]) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return bk_fail(BK_ERR_NCCL, "cannot dlopen libnccl.so.2 (set BK_NCCL_LIB): %s", dlerror());
#define BK_SYM(field, sym)                                                                  \
  *(void**)(&g_nccl.field) = dlsym(lib, sym);                                               \
  if (!g_nccl.field) return bk_fail(BK_ERR_NCCL, "libnccl lacks %s", sym);
  BK_SYM(GetUniqueId, "ncclGetUniqueId")
  BK_SYM(CommInitRank, "ncclCommInitRank")
  BK_SYM(CommDestroy, "ncclCommDestroy")
  BK_SYM(AllReduce, "ncclAllReduce")
  BK_SYM(Send, "ncclSend")
  BK_SYM(Recv, "ncclRecv")
  BK_SYM(GroupStart, "ncclGroupStart")
  BK_SYM(GroupEnd, "ncclGroupEnd")
  BK_SYM(GetErrorString, "ncclGetErrorString")
#undef BK_SYM
  g_nccl.lib = lib;
  return BK_OK;
}

extern "C" int bk_dist_unique_id(void* id128) {
  if (!id128) return bk_fail(BK_ERR_ARG, "bk_dist_unique_id: null buffer");
  BK_TRY(bk_nccl_load());
  bk_nccl_id id;
  BK_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return BK_OK;
}

void bk_dist_release_comm(bk_handle* h) {
  if (h && h->dist_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((bk_nccl_comm)h->dist_comm);
  if (h) h->dist_comm = nullptr;
}

extern "C" int bk_dist_destroy(bk_dist* D) {
  if (!D) return BK_OK;
  if (D->h) {
    cudaSetDevice(D->h->device);
    cudaDeviceSynchronize();
    bk_graphs_invalidate(D->h);
  }
  // D->comm belongs to the handle (shared by all matrices of this rank set): see bk_dist_release_comm
  if (D->Aloc) bk_csr_destroy(D->Aloc);
  if (D->Aext) bk_csr_destroy(D->Aext);
  for (int q = 0; q < BK_P2P_MAXP; ++q)
    if (D->peer_mapped[q]) cudaIpcCloseMemHandle(D->peer_mapped[q]);
  if (D->sendbuf) cudaFree(D->sendbuf);
  if (D->window) cudaFree(D->window);
  if (D->p2p.counters) cudaFree(D->p2p.counters);
  if (D->d_seg_start) cudaFree(D->d_seg_start);
  if (D->d_remote_ghost) cudaFree(D->d_remote_ghost);
  if (D->d_remote_flag) cudaFree(D->d_remote_flag);
  if (D->d_peer_ranks) cudaFree(D->d_peer_ranks);
  if (D->red) cudaFree(D->red);
  if (D->comm_stream) cudaStreamDestroy(D->comm_stream);
  if (D->ev_ready) cudaEventDestroy(D->ev_ready);
  if (D->ev_halo) cudaEventDestroy(D->ev_halo);
  free(D->peer_ranks);
  free(D->send_counts);
  free(D->recv_counts);
  free(D);
  return BK_OK;
}

extern "C" int bk_dist_create(bk_handle* h, const void* id128, int rank, int nranks, int64_t n_local, int64_t nnz_loc,
                              const void* loc_rowptr, const void* loc_col, const void* loc_val, int64_t n_brows,
                              const void* brow_ids, int64_t nnz_gh, const void* gh_rowptr, const void* gh_col,
                              const void* gh_val, int64_t n_ghost, int npeers, const int32_t* peer_ranks,
                              const int64_t* send_counts, const int64_t* recv_counts, const void* send_idx, int dtype,
                              void* stream, bk_dist** out) {
  if (!h || !out || !id128) return bk_fail(BK_ERR_ARG, "bk_dist_create: null handle/out/id");
  *out = nullptr;
  if (nranks < 1 || rank < 0 || rank >= nranks) return bk_fail(BK_ERR_ARG, "bk_dist_create: bad rank %d/%d", rank, nranks);
  if (npeers < 0 || (npeers > 0 && (!peer_ranks || !send_counts || !recv_counts)))
    return bk_fail(BK_ERR_ARG, "bk_dist_create: bad peer description");
  if (n_brows > 0 && (!brow_ids || !gh_rowptr)) return bk_fail(BK_ERR_ARG, "bk_dist_create: null ghost block");
  BK_CUDA(cudaSetDevice(h->device));
  BK_TRY(bk_nccl_load());
  bk_dist* D = (bk_dist*)calloc(1, sizeof(bk_dist));
  if (!D) return bk_fail(BK_ERR_ALLOC, "bk_dist_create: host allocation failed");
  D->h = h;
  D->rank = rank;
  D->nranks = nranks;
  D->n_local = n_local;
  D->n_ghost = n_ghost;
  D->n_brows = n_brows;
  D->nnz_gh = nnz_gh;
  D->dtype = dtype;
  D->uid = h->next_uid++;
  int rc = bk_csr_create(h, n_local, nnz_loc, loc_rowptr, loc_col, 32, loc_val, dtype, 0, stream, &D->Aloc);
  if (rc != BK_OK) {
    bk_dist_destroy(D);
    return rc;
  }
  D->brow_ids = (const int*)brow_ids;
  D->gh_rowptr = (const int*)gh_rowptr;
  D->gh_col = (const int*)gh_col;
  D->gh_val = gh_val;
  D->npeers = npeers;
  D->peer_ranks = (int*)malloc(sizeof(int) * (npeers + 1));
  D->send_counts = (int64_t*)malloc(sizeof(int64_t) * (npeers + 1));
  D->recv_counts = (int64_t*)malloc(sizeof(int64_t) * (npeers + 1));
  int64_t rtot = 0;
  for (int i = 0; i < npeers; ++i) {
    D->peer_ranks[i] = peer_ranks[i];
    D->send_counts[i] = send_counts[i];
    D->recv_counts[i] = recv_counts[i];
    D->send_total += send_counts[i];
    rtot += recv_counts[i];
  }
  if (rtot != n_ghost) {
    bk_dist_destroy(D);
    return bk_fail(BK_ERR_ARG, "bk_dist_create: recv counts sum to %lld, ghost vector has %lld", (long long)rtot,
                   (long long)n_ghost);
  }
  D->send_idx = (const int*)send_idx;
  const size_t vs = bk_dtype_size(dtype);
  cudaError_t e = cudaMalloc(&D->sendbuf, vs * (size_t)(D->send_total > 0 ? D->send_total : 1));
  D->window_bytes = BK_P2P_GHOST_OFF + ((vs * (size_t)(n_ghost > 0 ? n_ghost : 1) + 255) & ~(size_t)255);
  if (e == cudaSuccess) e = cudaMalloc((void**)&D->window, D->window_bytes);
  if (e == cudaSuccess) e = cudaMemset(D->window, 0, D->window_bytes);
  if (e == cudaSuccess) D->ghost = D->window + BK_P2P_GHOST_OFF;
  if (e == cudaSuccess) e = cudaMalloc((void**)&D->p2p.counters, sizeof(unsigned int) * 16);
  if (e == cudaSuccess) e = cudaMemset(D->p2p.counters, 0, sizeof(unsigned int) * 16);
  if (e == cudaSuccess) e = cudaMalloc(&D->red, sizeof(double) * BK_DIST_RED_DOUBLES);
  if (e == cudaSuccess) e = cudaMemset(D->red, 0, sizeof(double) * BK_DIST_RED_DOUBLES);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&D->comm_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&D->ev_ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&D->ev_halo, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    bk_dist_destroy(D);
    return bk_fail(BK_ERR_CUDA, "bk_dist_create: %s", cudaGetErrorString(e));
  }
  if (h->dist_comm && h->dist_comm_rank == rank && h->dist_comm_nranks == nranks) {
    D->comm = (bk_nccl_comm)h->dist_comm;  // every rank takes this branch together (same registration sequence)
  } else {
    bk_dist_release_comm(h);
    bk_nccl_id id;
    memcpy(&id, id128, 128);
    int nr = g_nccl.CommInitRank(&D->comm, nranks, id, rank);
    if (nr != 0) {
      const char* msg = g_nccl.GetErrorString(nr);
      D->comm = nullptr;
      bk_dist_destroy(D);
      return bk_fail(BK_ERR_NCCL, "ncclCommInitRank failed: %s", msg);
    }
    h->dist_comm = D->comm;
    h->dist_comm_rank = rank;
    h->dist_comm_nranks = nranks;
  }
  *out = D;
  return BK_OK;
}

extern "C" int bk_dist_set_extended(bk_dist* D, int64_t nnz_ext, const void* ext_rowptr, const void* ext_col,
                                    const void* ext_val, const void* ghost_gid, int64_t row_begin, void* stream,
                                    int32_t* folded) {
  if (!D || !folded) return bk_fail(BK_ERR_ARG, "bk_dist_set_extended: null argument");
  *folded = 0;
  if (nnz_ext < 0 || !ext_rowptr || (nnz_ext > 0 && (!ext_col || !ext_val)) || (D->n_ghost > 0 && !ghost_gid))
    return bk_fail(BK_ERR_ARG, "bk_dist_set_extended: null array");
  bk_handle* h = D->h;
  BK_CUDA(cudaSetDevice(h->device));
  if (D->Aext) {
    bk_csr_destroy(D->Aext);
    D->Aext = nullptr;
  }
  if (D->n_local + D->n_ghost >= 2147483647LL || nnz_ext >= 2147483647LL) return BK_OK;  // two-kernel path stays
  bk_csr* E = (bk_csr*)calloc(1, sizeof(bk_csr));
  if (!E) return bk_fail(BK_ERR_ALLOC, "bk_dist_set_extended: host allocation failed");
  E->h = h;
  E->n = D->n_local;
  E->n_cols = D->n_local + D->n_ghost;
  E->nnz = nnz_ext;
  E->dtype = D->dtype;
  E->rowptr = (const int*)ext_rowptr;
  E->col = (const int*)ext_col;
  E->val = ext_val;
  E->uid = h->next_uid++;
  E->reg_ghost_gid = (const long long*)ghost_gid;
  E->reg_row_begin = row_begin;
  int rc = bk_csr_finish_plan(h, E, (cudaStream_t)stream);
  // the plan (masks, pattern ids, pattern table) is self-contained: the CSR arrays are not referenced afterwards
  E->rowptr = nullptr;
  E->col = nullptr;
  E->val = nullptr;
  E->reg_ghost_gid = nullptr;
  if (rc != BK_OK || E->kernel != 6) {
    bk_csr_destroy(E);
    return rc;
  }
  D->Aext = E;
  *folded = bk_mask2_usable(h, E) ? 2 : 1;  // 2: the interior steps run kernel 7 (stencil fast path)
  bk_graphs_invalidate(h);
  return BK_OK;
}

// ---- peer-memory path: export this rank's window, map everybody else's ---------------------------------------
extern "C" int bk_dist_p2p_export(bk_dist* D, void* handle64) {
  if (!D || !handle64) return bk_fail(BK_ERR_ARG, "bk_dist_p2p_export: null argument");
  BK_CUDA(cudaSetDevice(D->h->device));
  cudaIpcMemHandle_t hd;
  BK_CUDA(cudaIpcGetMemHandle(&hd, D->window));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle64, &hd, 64);
  return BK_OK;
}

// handles: nranks x 64 bytes (rank order).  remote_ghost_offsets[i]: element offset, inside halo peer i's ghost
// vector, at which the entries this rank sends to it must land (that peer's receive offset for this rank).
extern "C" int bk_dist_p2p_connect(bk_dist* D, const void* handles, const int64_t* remote_ghost_offsets) {
  if (!D || !handles) return bk_fail(BK_ERR_ARG, "bk_dist_p2p_connect: null argument");
  if (D->nranks > BK_P2P_MAXP) return bk_fail(BK_ERR_UNSUPPORTED, "peer-memory path supports up to %d ranks", BK_P2P_MAXP);
  if (D->npeers > 0 && !remote_ghost_offsets) return bk_fail(BK_ERR_ARG, "bk_dist_p2p_connect: null offsets");
  BK_CUDA(cudaSetDevice(D->h->device));
  D->p2p.P = 0;
  D->p2p.rank = D->rank;
  for (int q = 0; q < D->nranks; ++q) {
    if (q == D->rank) {
      D->p2p.win[q] = D->window;
      continue;
    }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, (const char*)handles + (size_t)q * 64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return bk_fail(BK_ERR_UNSUPPORTED, "cudaIpcOpenMemHandle(rank %d) failed: %s (falling back to NCCL)", q,
                     cudaGetErrorString(e));
    }
    D->peer_mapped[q] = p;
    D->p2p.win[q] = (char*)p;
  }
  const size_t vs = bk_dtype_size(D->dtype);
  const int np = D->npeers;
  long long seg[BK_P2P_MAXP + 1];
  void* rg[BK_P2P_MAXP];
  unsigned long long* rf[BK_P2P_MAXP];
  if (np > BK_P2P_MAXP) return bk_fail(BK_ERR_UNSUPPORTED, "too many halo peers");
  seg[0] = 0;
  for (int i = 0; i < np; ++i) {
    seg[i + 1] = seg[i] + D->send_counts[i];
    char* w = D->p2p.win[D->peer_ranks[i]];
    rg[i] = w + BK_P2P_GHOST_OFF + (size_t)remote_ghost_offsets[i] * vs;
    rf[i] = (unsigned long long*)(w + BK_P2P_FLAG_OFF) + D->rank;
  }
  BK_CUDA(cudaMalloc((void**)&D->d_seg_start, sizeof(long long) * (np + 1)));
  BK_CUDA(cudaMalloc((void**)&D->d_remote_ghost, sizeof(void*) * (np > 0 ? np : 1)));
  BK_CUDA(cudaMalloc((void**)&D->d_remote_flag, sizeof(void*) * (np > 0 ? np : 1)));
  BK_CUDA(cudaMalloc((void**)&D->d_peer_ranks, sizeof(int) * (np > 0 ? np : 1)));
  BK_CUDA(cudaMemcpy(D->d_seg_start, seg, sizeof(long long) * (np + 1), cudaMemcpyHostToDevice));
  if (np > 0) {
    BK_CUDA(cudaMemcpy(D->d_remote_ghost, rg, sizeof(void*) * np, cudaMemcpyHostToDevice));
    BK_CUDA(cudaMemcpy(D->d_remote_flag, rf, sizeof(void*) * np, cudaMemcpyHostToDevice));
    BK_CUDA(cudaMemcpy(D->d_peer_ranks, D->peer_ranks, sizeof(int) * np, cudaMemcpyHostToDevice));
  }
  // fused push: possible when every neighbour's send list is one ascending contiguous index range
  D->push_nranges = -1;
  if (np <= BK_PUSH_MAXR && D->send_total > 0) {
    int* idx_h = (int*)malloc(sizeof(int) * (size_t)D->send_total);
    if (idx_h && cudaMemcpy(idx_h, D->send_idx, sizeof(int) * (size_t)D->send_total, cudaMemcpyDeviceToHost) ==
                     cudaSuccess) {
      bool ok = true;
      int nr = 0;
      for (int i = 0; i < np && ok; ++i) {
        const long long c = D->send_counts[i];
        if (c == 0) continue;
        const int* seg_i = idx_h + seg[i];
        for (long long k = 1; k < c && ok; ++k) ok = (seg_i[k] == seg_i[0] + (int)k);
        D->push_lo[nr] = seg_i[0];
        D->push_hi[nr] = seg_i[0] + c;
        D->push_dst[nr] = rg[i];
        ++nr;
      }
      if (ok) D->push_nranges = nr;
    } else {
      cudaGetLastError();
    }
    free(idx_h);
  } else if (D->send_total == 0 && np <= BK_PUSH_MAXR) {
    D->push_nranges = 0;
  }
  D->p2p.P = D->nranks;
  D->p2p_enabled = 1;
  bk_graphs_invalidate(D->h);
  return BK_OK;
}

extern "C" int bk_dist_spmv(bk_handle* h, bk_dist* D, const void* x_local, void* y_local, void* stream) {
  if (!h || !D || !x_local || !y_local) return bk_fail(BK_ERR_ARG, "bk_dist_spmv: null argument");
  BK_CUDA(cudaSetDevice(h->device));
  const bk_sys_dist sys{h, D, false, 0};
  if (D->dtype == BK_F64) return sys.spmv<double>(x_local, y_local, (cudaStream_t)stream);
  return sys.spmv<float>(x_local, y_local, (cudaStream_t)stream);
}
