// bk_dist.cu — row-partitioned (multi-GPU) SpMV and CG.  One process per GPU; SURVEY §8e.
//
// Each rank owns a contiguous slab of rows and the matching slices of every vector.  Its rows are split (by the
// Python set-up code, pytorch_sparse_solver/distributed.py) into
//   * a LOCAL block  (columns inside the slab, renumbered)  -> a normal bk_csr, run by the same SpMV kernels, and
//   * a GHOST block  (columns owned by other ranks, renumbered into a compact ghost vector) stored as a CSR over the
//     few "boundary rows" that have such entries.
// One distributed SpMV:
//   main stream : pack boundary entries of x into the send buffer ........ K_local SpMV (+ fused dot) .. wait .. K_ghost rows
//   comm stream :            \-> ncclSend/ncclRecv with every neighbour (grouped) --------------------------/
// so the halo exchange (NVLink, 2 MiB per neighbour for a 512^2 plane) overlaps the interior SpMV.  The two scalar
// reductions of a CG iteration go through ncclAllReduce on one double each; the scalars that depend on them are
// formed by 1-thread kernels, so the whole iteration stays on the device and is captured in a CUDA graph like the
// single-GPU loop.  NCCL is dlopen'ed (torch's bundled libnccl.so.2 is already in the process), never linked.
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

extern "C" int bk_dist_destroy(bk_dist* D) {
  if (!D) return BK_OK;
  if (D->h) {
    cudaSetDevice(D->h->device);
    cudaDeviceSynchronize();
    bk_graphs_invalidate(D->h);
  }
  if (D->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(D->comm);
  if (D->Aloc) bk_csr_destroy(D->Aloc);
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
  bk_nccl_id id;
  memcpy(&id, id128, 128);
  int nr = g_nccl.CommInitRank(&D->comm, nranks, id, rank);
  if (nr != 0) {
    const char* msg = g_nccl.GetErrorString(nr);
    D->comm = nullptr;
    bk_dist_destroy(D);
    return bk_fail(BK_ERR_NCCL, "ncclCommInitRank failed: %s", msg);
  }
  *out = D;
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
  D->p2p.P = D->nranks;
  D->p2p_enabled = 1;
  bk_graphs_invalidate(D->h);
  return BK_OK;
}

// boundary rows with the halo wait in front and the p.Ap all-reduce + alpha behind (peer-memory path of CG's K1b)
template <typename T>
__global__ void __launch_bounds__(BK_BLOCK)
bk_ghost_rows_p2p_kernel(const int* __restrict__ brow_ids, const int* __restrict__ rowptr, const int* __restrict__ col,
                         const T* __restrict__ val, const T* ghost, T* __restrict__ y, const T* __restrict__ w,
                         long long n_brows, const bk_scratch sc, const double* local_partial, bk_dev_state* st,
                         const bk_p2p_ctx p2p, const int* __restrict__ peer_ranks, int npeers) {
  if (st->done) return;
  __shared__ int s_fail;
  if (threadIdx.x == 0) s_fail = 0;
  __syncthreads();
  const unsigned int want = p2p.counters[2] + 1u;  // halos consumed so far + 1
  if (threadIdx.x < npeers) {
    const unsigned long long* flag =
        reinterpret_cast<const unsigned long long*>(p2p.win[p2p.rank] + BK_P2P_FLAG_OFF) + peer_ranks[threadIdx.x];
    const long long t0 = clock64();
    while ((unsigned int)bk_ld_acquire_sys_u64(flag) != want) {
      if (clock64() - t0 > BK_P2P_TIMEOUT_CYCLES) {
        s_fail = 1;
        break;
      }
    }
  }
  __syncthreads();
  double acc[1] = {0.0};
  if (!s_fail) {
    const long long stride = (long long)gridDim.x * BK_BLOCK;
    for (long long b = (long long)blockIdx.x * BK_BLOCK + threadIdx.x; b < n_brows; b += stride) {
      const int r = brow_ids[b];
      T sum = T(0);
      for (int k = rowptr[b]; k < rowptr[b + 1]; ++k) sum = fma(val[k], __ldcg(ghost + col[k]), sum);
      y[r] = y[r] + sum;
      acc[0] += (double)w[r] * (double)sum;
    }
  } else if (threadIdx.x == 0) {
    p2p.counters[4] = 1u;
  }
  bk_grid_reduce<1>(acc, sc, [&](const double* s) {
    p2p.counters[2] = want;
    const double total = bk_p2p_allreduce(p2p, local_partial[0] + s[0]);
    if (p2p.counters[4]) {
      st->done = 1;
      st->status = BK_ST_COMM_TIMEOUT;
      return;
    }
    st->pAp = total;
    st->alpha = st->gamma / total;
  });
}

// ---- kernels ---------------------------------------------------------------------------------------
struct bk_epi_dist_store {  // out[0] = base[0] + s   (adds the local-block partial to the ghost-block partial)
  double* out;
  const double* base;
  __device__ __forceinline__ void operator()(const double* s) const { out[0] = (base ? base[0] : 0.0) + s[0]; }
};

// boundary rows: y[r] += sum_k gh_val[k] * ghost[gh_col[k]] ; optional dot partial w[r] * sum
template <typename T, int DOT>
__global__ void __launch_bounds__(BK_BLOCK)
bk_ghost_rows_cg_kernel(const int* __restrict__ brow_ids, const int* __restrict__ rowptr, const int* __restrict__ col,
                     const T* __restrict__ val, const T* __restrict__ ghost, T* __restrict__ y, const T* __restrict__ w,
                     long long n_brows, const bk_scratch sc, bk_epi_dist_store epi, const bk_dev_state* st, int guard) {
  if (guard && st->done) return;
  double acc[1] = {0.0};
  const long long stride = (long long)gridDim.x * BK_BLOCK;
  for (long long b = (long long)blockIdx.x * BK_BLOCK + threadIdx.x; b < n_brows; b += stride) {
    const int r = brow_ids[b];
    T sum = T(0);
    for (int k = rowptr[b]; k < rowptr[b + 1]; ++k) sum = fma(val[k], ghost[col[k]], sum);
    y[r] = y[r] + sum;
    if (DOT) acc[0] += (double)w[r] * (double)sum;
  }
  if (DOT) bk_grid_reduce<1>(acc, sc, epi);
}

struct bk_epi_store_pap {  // local-block partial of p.Ap
  double* out;
  __device__ __forceinline__ void operator()(const double* s) const { out[0] = s[0]; }
};

__global__ void bk_dist_alpha_kernel(bk_dev_state* st, const double* red) {  // after allreduce of p.Ap
  if (st->done) return;
  st->pAp = red[0];
  st->alpha = st->gamma / red[0];
}

__global__ void bk_dist_beta_kernel(bk_dev_state* st, const double* red) {  // after allreduce of r.r   (:849-853, :841)
  if (st->done) return;
  const double gamma_new = red[1];
  st->beta = gamma_new / st->gamma;
  st->gamma = gamma_new;
  const long long k = st->k + 1;
  st->k = k;
  st->parity ^= 1;
  if (k >= st->maxiter) {
    st->done = 1;
    st->status = BK_ST_MAXITER;
  }
  if (gamma_new <= st->atol2) {
    st->done = 1;
    st->status = BK_ST_CONVERGED;
  }
}

__global__ void bk_dist_init_kernel(bk_dev_state* st, const double* red) {  // red[2] = b.b, red[3] = r0.r0 (global)
  const double bs = red[2];
  st->bs = bs;
  st->atol2 = fmax(st->tolsq32 * bs, st->atolsq32);
  st->gamma = red[3];
  if (st->maxiter <= 0) {
    st->done = 1;
    st->status = BK_ST_MAXITER;
  }
  if (st->gamma <= st->atol2) {
    st->done = 1;
    st->status = BK_ST_CONVERGED;
  }
}

__global__ void bk_dist_final_kernel(bk_dev_state* st, const double* red) {  // red[2] = |b-Ax|^2, red[3] = x.x
  st->rtrue2 = red[2];
  st->xx = red[3];
}

template <typename T>
struct bk_op_dot_to {
  static constexpr int R = 1;
  using Ctx = bk_noctx;
  template <int W>
  struct In {
    bk_vec<T, W> a, b;
  };
  const T* x;
  const T* y;
  double* out;
  __device__ bool skip() const { return false; }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const { return Ctx(); }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.a = bk_ld<T, W>(x + i);
    in.b = bk_ld<T, W>(y + i);
  }
  template <int W>
  __device__ void apply(long long, const In<W>& in, const Ctx&, double (&acc)[1]) const {
#pragma unroll
    for (int j = 0; j < W; ++j) acc[0] += (double)in.a.v[j] * (double)in.b.v[j];
  }
  __device__ void epilogue(const double* s) const { out[0] = s[0]; }
};

// y = A x (distributed).  If dot_out != nullptr also leaves the LOCAL partial of w.y there (caller all-reduces).
template <typename T>
static int bk_dist_spmv_t(bk_handle* h, bk_dist* D, const T* x, T* y, const T* w, double* dot_out, int guard,
                          cudaStream_t s) {
  const long long n = D->n_local;
  bk_dev_state* st = h->st;
  if (D->npeers > 0) {
    if (D->send_total > 0) {
      int g = (int)((D->send_total + 255) / 256);
      if (g > h->num_sms * 4) g = h->num_sms * 4;
      bk_halo_pack_kernel<T><<<g, 256, 0, s>>>(x, D->send_idx, (T*)D->sendbuf, D->send_total, st, guard);
      BK_KERNEL_CHECK();
    }
    BK_CUDA(cudaEventRecord(D->ev_ready, s));
    BK_CUDA(cudaStreamWaitEvent(D->comm_stream, D->ev_ready, 0));
    const int nt = sizeof(T) == 8 ? BK_NCCL_F64 : BK_NCCL_F32;
    BK_NCCL(g_nccl.GroupStart());
    int64_t so = 0, ro = 0;
    for (int i = 0; i < D->npeers; ++i) {
      if (D->send_counts[i] > 0)
        BK_NCCL(g_nccl.Send((const T*)D->sendbuf + so, (size_t)D->send_counts[i], nt, D->peer_ranks[i], D->comm,
                            D->comm_stream));
      if (D->recv_counts[i] > 0)
        BK_NCCL(g_nccl.Recv((T*)D->ghost + ro, (size_t)D->recv_counts[i], nt, D->peer_ranks[i], D->comm,
                            D->comm_stream));
      so += D->send_counts[i];
      ro += D->recv_counts[i];
    }
    BK_NCCL(g_nccl.GroupEnd());
    BK_CUDA(cudaEventRecord(D->ev_halo, D->comm_stream));
  }
  {  // local block (overlaps the exchange)
    bk_spmv_args a = bk_spmv_base(D->Aloc, st);
    a.x = x;
    a.y = y;
    a.w = w;
    a.guard = guard;
    if (dot_out) {
      bk_epi_store_pap epi{D->red + 4};
      BK_TRY((bk_launch_spmv<0, 1, 0>(h, D->Aloc, a, bk_slot(h, 0), epi, s)));
    } else {
      BK_TRY((bk_launch_spmv<0, 0, 0>(h, D->Aloc, a, bk_slot(h, 0), bk_epi_none(), s)));
    }
  }
  if (D->npeers > 0) BK_CUDA(cudaStreamWaitEvent(s, D->ev_halo, 0));
  {  // ghost block rows (also finishes the dot: out = local partial + ghost partial)
    int g = (int)((D->n_brows + BK_BLOCK - 1) / BK_BLOCK);
    if (g < 1) g = 1;
    if (g > h->num_sms * 4) g = h->num_sms * 4;
    bk_epi_dist_store epi{dot_out, D->red + 4};
    if (dot_out) {
      bk_ghost_rows_cg_kernel<T, 1><<<g, BK_BLOCK, 0, s>>>(D->brow_ids, D->gh_rowptr, D->gh_col, (const T*)D->gh_val,
                                                       (const T*)D->ghost, y, w, D->n_brows, bk_slot(h, 3), epi, st,
                                                       guard);
    } else if (D->n_brows > 0) {
      bk_ghost_rows_cg_kernel<T, 0><<<g, BK_BLOCK, 0, s>>>(D->brow_ids, D->gh_rowptr, D->gh_col, (const T*)D->gh_val,
                                                       (const T*)D->ghost, y, w, D->n_brows, bk_slot(h, 3), epi, st,
                                                       guard);
    }
    BK_KERNEL_CHECK();
  }
  return BK_OK;
}

extern "C" int bk_dist_spmv(bk_handle* h, bk_dist* D, const void* x_local, void* y_local, void* stream) {
  if (!h || !D || !x_local || !y_local) return bk_fail(BK_ERR_ARG, "bk_dist_spmv: null argument");
  BK_CUDA(cudaSetDevice(h->device));
  if (D->dtype == BK_F64)
    return bk_dist_spmv_t<double>(h, D, (const double*)x_local, (double*)y_local, nullptr, nullptr, 0,
                                  (cudaStream_t)stream);
  return bk_dist_spmv_t<float>(h, D, (const float*)x_local, (float*)y_local, nullptr, nullptr, 0, (cudaStream_t)stream);
}

template <typename T>
static int bk_dist_cg_t(bk_handle* h, bk_dist* D, const void* b, void* x_user, int has_x0, double tol, double atol,
                        int64_t maxiter, int64_t n_global, bk_result* res, cudaStream_t s) {
  const long long n = D->n_local;
  const size_t npad = ((size_t)n + 63) & ~(size_t)63;
  BK_TRY(bk_ws_reserve(h, (size_t)4 * npad * sizeof(T)));
  T* x = (T*)h->ws;
  T* r = x + npad;
  T* p = r + npad;
  T* ap = p + npad;
  bk_dev_state* st = h->st;
  const size_t vbytes = (size_t)n * sizeof(T);
  double* red = D->red;

  bk_dev_state init;
  memset(&init, 0, sizeof(init));
  init.maxiter = maxiter < 0 ? 10 * n_global : maxiter;
  init.status = BK_ST_MAXITER;
  bk_state_fill_tol(&init, tol, atol);
  bk_state_set_kernel<<<1, 1, 0, s>>>(st, init);
  BK_KERNEL_CHECK();

  auto dot_to = [&](const T* a, const T* c, double* out) -> int {
    bk_op_dot_to<T> op;
    op.x = a;
    op.y = c;
    op.out = out;
    return bk_launch_ew<T>(h, op, n, bk_aligned16(a) && bk_aligned16(c), bk_slot(h, 1), s);
  };
  auto axpby = [&](double ca, const T* a, double cb, const T* c, T* z) -> int {
    bk_op_axpby<T> op;
    op.x = a;
    op.y = c;
    op.z = z;
    op.ca = (T)ca;
    op.cb = (T)cb;
    return bk_launch_ew<T>(h, op, n, bk_aligned16(a) && bk_aligned16(c) && bk_aligned16(z), bk_slot(h, 1), s);
  };

  if (has_x0) {
    BK_CUDA(cudaMemcpyAsync(x, x_user, vbytes, cudaMemcpyDeviceToDevice, s));
    BK_TRY(bk_dist_spmv_t<T>(h, D, x, ap, nullptr, nullptr, 0, s));
    BK_TRY(axpby(1.0, (const T*)b, -1.0, ap, r));  // r0 = b - A x0
  } else {
    BK_CUDA(cudaMemsetAsync(x, 0, vbytes, s));
    BK_CUDA(cudaMemcpyAsync(r, b, vbytes, cudaMemcpyDeviceToDevice, s));
  }
  BK_TRY(dot_to((const T*)b, (const T*)b, red + 2));
  BK_TRY(dot_to(r, r, red + 3));
  BK_NCCL(g_nccl.AllReduce(red + 2, red + 2, 2, BK_NCCL_F64, BK_NCCL_SUM, D->comm, s));
  bk_dist_init_kernel<<<1, 1, 0, s>>>(st, red);
  BK_KERNEL_CHECK();
  BK_CUDA(cudaMemcpyAsync(p, r, vbytes, cudaMemcpyDeviceToDevice, s));

  const bool p2p = D->p2p_enabled && h->dist_p2p;
  auto enqueue_iter_p2p = [&](cudaStream_t cs) -> int {
    if (D->send_total > 0) {
      int g = (int)((D->send_total + 255) / 256);
      if (g > h->num_sms * 2) g = h->num_sms * 2;
      bk_halo_push_kernel<T><<<g, 256, 0, cs>>>(p, D->send_idx, D->d_seg_start, D->d_remote_ghost, D->d_remote_flag,
                                                D->npeers, D->send_total, D->p2p.counters, st, 1);
      BK_KERNEL_CHECK();
    }
    {
      bk_spmv_args a = bk_spmv_base(D->Aloc, st);
      a.x = p;
      a.y = ap;
      a.w = p;
      a.guard = 1;
      bk_epi_store_pap epi{red + 4};
      BK_TRY((bk_launch_spmv<0, 1, 0>(h, D->Aloc, a, bk_slot(h, 0), epi, cs)));
    }
    {
      int g = (int)((D->n_brows + BK_BLOCK - 1) / BK_BLOCK);
      if (g < 1) g = 1;
      if (g > h->num_sms * 4) g = h->num_sms * 4;
      bk_ghost_rows_p2p_kernel<T><<<g, BK_BLOCK, 0, cs>>>(D->brow_ids, D->gh_rowptr, D->gh_col, (const T*)D->gh_val,
                                                         (const T*)D->ghost, ap, p, D->n_brows, bk_slot(h, 3), red + 4,
                                                         st, D->p2p, D->d_peer_ranks, D->npeers);
      BK_KERNEL_CHECK();
    }
    {
      bk_op_cg_update<T> op;
      op.p = p;
      op.ap = ap;
      op.x = x;
      op.r = r;
      op.st = st;
      op.snake = 0;
      op.dist_out = nullptr;
      op.p2p = D->p2p;
      BK_TRY(bk_launch_ew<T>(h, op, n, true, bk_slot(h, 1), cs));
    }
    {
      bk_op_xpay<T> op;
      op.r = r;
      op.p = p;
      op.st = st;
      op.snake = 0;
      BK_TRY(bk_launch_ew<T>(h, op, n, true, bk_slot(h, 2), cs));
    }
    return BK_OK;
  };
  auto enqueue_iter = [&](cudaStream_t cs) -> int {
    if (p2p) return enqueue_iter_p2p(cs);
    BK_TRY(bk_dist_spmv_t<T>(h, D, p, ap, p, red, 1, cs));                                  // Ap, local p.Ap
    BK_NCCL(g_nccl.AllReduce(red, red, 1, BK_NCCL_F64, BK_NCCL_SUM, D->comm, cs));
    bk_dist_alpha_kernel<<<1, 1, 0, cs>>>(st, red);
    BK_KERNEL_CHECK();
    {
      bk_op_cg_update<T> op;
      op.p = p;
      op.ap = ap;
      op.x = x;
      op.r = r;
      op.st = st;
      op.snake = 0;
      op.dist_out = red + 1;
      op.p2p.P = 0;
      BK_TRY(bk_launch_ew<T>(h, op, n, true, bk_slot(h, 1), cs));
    }
    BK_NCCL(g_nccl.AllReduce(red + 1, red + 1, 1, BK_NCCL_F64, BK_NCCL_SUM, D->comm, cs));
    bk_dist_beta_kernel<<<1, 1, 0, cs>>>(st, red);
    BK_KERNEL_CHECK();
    {
      bk_op_xpay<T> op;
      op.r = r;
      op.p = p;
      op.st = st;
      op.snake = 0;
      BK_TRY(bk_launch_ew<T>(h, op, n, true, bk_slot(h, 2), cs));
    }
    return BK_OK;
  };
  const double bytes_iter = (double)D->Aloc->nnz * (sizeof(T) + 4) + 4.0 * (n + 1) + 11.0 * n * sizeof(T);
  const int chunk = bk_pick_chunk(h, bytes_iter, 8);
  const bool use_graph = h->loop_mode != BK_LOOP_STREAM;  // NCCL calls are captured into the iteration graph too
  uint64_t key[6] = {4 /*dist cg*/, D->uid, (uint64_t)(uintptr_t)h->ws, (uint64_t)n,
                     (uint64_t)D->dtype | ((uint64_t)p2p << 8) | ((uint64_t)chunk << 16), (uint64_t)bk_grid_spmv(h) | ((uint64_t)bk_grid_vec(h) << 32)};
  auto enqueue_chunk = [&](cudaStream_t cs) -> int {
    for (int it = 0; it < chunk; ++it) BK_TRY(enqueue_iter(cs));
    return BK_OK;
  };
  int64_t chunks = 0;
  BK_TRY(bk_run_loop(h, s, use_graph, key, enqueue_chunk, &chunks));

  // final true residual and ||x|| (global)
  BK_TRY(bk_dist_spmv_t<T>(h, D, x, ap, nullptr, nullptr, 0, s));
  BK_TRY(axpby(1.0, (const T*)b, -1.0, ap, ap));
  BK_TRY(dot_to(ap, ap, red + 2));
  BK_TRY(dot_to(x, x, red + 3));
  BK_NCCL(g_nccl.AllReduce(red + 2, red + 2, 2, BK_NCCL_F64, BK_NCCL_SUM, D->comm, s));
  bk_dist_final_kernel<<<1, 1, 0, s>>>(st, red);
  BK_KERNEL_CHECK();
  BK_CUDA(cudaMemcpyAsync(x_user, x, vbytes, cudaMemcpyDeviceToDevice, s));
  BK_CUDA(cudaMemcpyAsync(&h->st_host[3], st, sizeof(bk_dev_state), cudaMemcpyDeviceToHost, s));
  BK_CUDA(cudaStreamSynchronize(s));
  const bk_dev_state* fin = &h->st_host[3];
  bk_fill_result_isolve(fin, res, fin->k + (has_x0 ? 1 : 0));
  res->rr_last = fin->gamma;
  res->kernel_launches = chunks * chunk * (p2p ? 5 : 7) + 12;
  if (fin->status == BK_ST_COMM_TIMEOUT)
    return bk_fail(BK_ERR_NCCL, "bk_dist_cg: a peer did not arrive within the timeout (peer-memory path)");
  return BK_OK;
}

extern "C" int bk_dist_cg(bk_handle* h, bk_dist* D, const void* b_local, void* x_local, int has_x0, double tol,
                          double atol, int64_t maxiter, int64_t n_global, bk_result* result, void* stream) {
  if (!h || !D || !result) return bk_fail(BK_ERR_ARG, "bk_dist_cg: null handle/matrix/result");
  if (D->n_local > 0 && (!b_local || !x_local)) return bk_fail(BK_ERR_ARG, "bk_dist_cg: null vector");
  memset(result, 0, sizeof(*result));
  BK_CUDA(cudaSetDevice(h->device));
  if (D->dtype == BK_F64)
    return bk_dist_cg_t<double>(h, D, b_local, x_local, has_x0, tol, atol, maxiter, n_global, result,
                                (cudaStream_t)stream);
  return bk_dist_cg_t<float>(h, D, b_local, x_local, has_x0, tol, atol, maxiter, n_global, result,
                             (cudaStream_t)stream);
}
