// bk_dist.cuh — the row-partitioned matrix (one slab per GPU) and the kernels every distributed solver shares:
// halo pack / push, boundary ("ghost block") rows, and bk_sys_dist, the multi-GPU implementation of the small
// interface (bk_sys.cuh) the Krylov drivers are written against.  SURVEY §8e.
//
// One distributed matvec y = A x (+ fused dots), on the solver's stream:
//   NCCL path : pack boundary entries -> [comm stream: grouped ncclSend/ncclRecv] | local-block SpMV (stores the local
//               dot partials) -> wait -> ghost-block rows (adds its partial) -> ncclAllReduce -> 1-thread scalar step
//   peer path : push boundary entries straight into the neighbours' ghost vectors (NVLink stores + arrival flags)
//               | local-block SpMV -> ghost-block rows: wait for the flags, finish y, and in its epilogue all-reduce
//               the dots through the peer windows and run the scalar step — no NCCL call, no extra kernel.
#pragma once

#include "bk_internal.cuh"
#include "bk_p2p.cuh"
#include "bk_spmv.cuh"
#include "bk_sys.cuh"
#include "bk_vec.cuh"

// ---- minimal NCCL surface (nccl.h is not required at build time; the library is dlopen'ed in bk_dist.cu) ------
typedef struct { char internal[128]; } bk_nccl_id;
typedef void* bk_nccl_comm;
struct bk_nccl_api {
  void* lib;
  int (*GetUniqueId)(bk_nccl_id*);
  int (*CommInitRank)(bk_nccl_comm*, int, bk_nccl_id, int);
  int (*CommDestroy)(bk_nccl_comm);
  int (*AllReduce)(const void*, void*, size_t, int, int, bk_nccl_comm, cudaStream_t);
  int (*Send)(const void*, size_t, int, int, bk_nccl_comm, cudaStream_t);
  int (*Recv)(void*, size_t, int, int, bk_nccl_comm, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char* (*GetErrorString)(int);
};
extern bk_nccl_api g_nccl;
enum { BK_NCCL_SUM = 0, BK_NCCL_F32 = 7, BK_NCCL_F64 = 8 };

#define BK_NCCL(expr)                                                                                   \
  do {                                                                                                  \
    int _r = (expr);                                                                                    \
    if (_r != 0) return bk_fail(BK_ERR_NCCL, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,             \
                                g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?");              \
  } while (0)

#define BK_PUSH_MAXR 4
#define BK_DIST_RED_DOUBLES 16  // [0..3] sums being all-reduced  [4] CG local partial  [8..9] local-block partials

struct bk_dist {
  bk_handle* h;
  int rank, nranks;
  int64_t n_local, n_ghost, n_brows, nnz_gh;
  int dtype;
  bk_csr* Aloc;
  bk_csr* Aext;  // [local | ghost] rows with a row-bitmask plan (kernel 6): one SpMV kernel per matvec on the peer path
  const int* brow_ids;
  const int* gh_rowptr;
  const int* gh_col;
  const void* gh_val;
  int npeers;
  int* peer_ranks;
  int64_t* send_counts;
  int64_t* recv_counts;
  int64_t send_total;
  const int* send_idx;
  void* sendbuf;
  void* ghost;  // = window + BK_P2P_GHOST_OFF
  char* window; // IPC-exportable: all-reduce slots, halo flags, ghost vector (bk_p2p.cuh)
  size_t window_bytes;
  int p2p_enabled;
  bk_p2p_ctx p2p;
  void* peer_mapped[BK_P2P_MAXP];     // cudaIpcOpenMemHandle results (to close)
  long long* d_seg_start;              // device [npeers+1]: prefix sums of send_counts
  void** d_remote_ghost;               // device [npeers]: where my entries land in each peer's ghost vector
  unsigned long long** d_remote_flag;  // device [npeers]: my arrival flag inside each peer's window
  int* d_peer_ranks;                   // device [npeers]
  // fused push (peer path): when every neighbour's boundary entries form ONE contiguous index range (slab
  // partitions), the kernel that PRODUCES a vector can store those entries straight into the neighbours' ghost
  // vectors while it streams — no separate push kernel.  push_nranges < 0: pattern not contiguous, not available.
  int push_nranges;
  long long push_lo[BK_PUSH_MAXR], push_hi[BK_PUSH_MAXR];
  void* push_dst[BK_PUSH_MAXR];
  double* red;                         // BK_DIST_RED_DOUBLES doubles
  bk_nccl_comm comm;
  cudaStream_t comm_stream;
  cudaEvent_t ev_ready, ev_halo;
  uint64_t uid;
};

// push: every boundary entry of x is stored straight into the owning peer's ghost vector (NVLink stores); the last
// CTA then publishes this rank's arrival flag (sequence number) in every peer's window with release semantics.
template <typename T>
__global__ void __launch_bounds__(256)
bk_halo_push_kernel(const T* __restrict__ x, const int* __restrict__ idx, const long long* __restrict__ seg_start,
                    void* const* __restrict__ remote_ghost, unsigned long long* const* __restrict__ remote_flag,
                    int npeers, long long count, unsigned int* counters, const bk_dev_state* st, int guard) {
  if (bk_guard_skip(st, guard)) return;
  __shared__ int s_last;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    int p = 0;
    while (p + 1 < npeers && i >= seg_start[p + 1]) ++p;
    static_cast<T*>(remote_ghost[p])[i - seg_start[p]] = x[idx[i]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicInc(&counters[3], gridDim.x - 1);
    s_last = (t == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    const unsigned int seq = counters[1] + 1u;
    if (threadIdx.x < npeers) bk_st_release_sys_u64(remote_flag[threadIdx.x], (unsigned long long)seq);
    __syncthreads();
    if (threadIdx.x == 0) counters[1] = seq;
  }
}

template <typename T>
__global__ void bk_halo_pack_kernel(const T* __restrict__ x, const int* __restrict__ idx, T* __restrict__ out,
                                    long long count, const bk_dev_state* st, int guard) {
  if (bk_guard_skip(st, guard)) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) out[i] = x[idx[i]];
}

template <int R>
struct bk_epi_store_n {  // local-block partials of the fused dots
  double* out;
  __device__ __forceinline__ void operator()(const double* s) const {
#pragma unroll
    for (int i = 0; i < R; ++i) out[i] = s[i];
  }
};

template <typename T>
struct bk_ghost_args {
  const int* brow_ids;
  const int* rowptr;
  const int* col;
  const T* val;
  const T* ghost;
  T* y;
  const T* w;
  long long n_brows;
  const int* peer_ranks;
  int npeers;
};

// Boundary rows: y[r] (+|-)= sum_k gh_val[k] * ghost[gh_col[k]], plus the corrections that turn the local block's
// dot partials (`base`) into this rank's full partials: w.y changes by w[r]*(+|-sum), y.y by y_new^2 - y_old^2.
// WAIT (peer path): first wait for every neighbour's arrival flag.  The epilogue makes the sums global (bk_gsum)
// and, unless that is deferred to NCCL, runs the solver's scalar step.
template <typename T, int MODE, int DOTS, bool WAIT, typename Epi>
__global__ void __launch_bounds__(BK_BLOCK)
bk_ghost_rows_kernel(const bk_ghost_args<T> ga, const bk_scratch sc, const double* base, const bk_gsum gs,
                     bk_dev_state* st, int guard, Epi epi) {
  if (bk_guard_skip(st, guard)) return;
  constexpr int R = bk_ndots<DOTS>::value;
  __shared__ int s_fail;
  unsigned int want = 0;
  if (WAIT) {
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    want = gs.p2p.counters[2] + 1u;  // halos consumed so far + 1
    if (threadIdx.x < ga.npeers) {
      const unsigned long long* flag =
          reinterpret_cast<const unsigned long long*>(gs.p2p.win[gs.p2p.rank] + BK_P2P_FLAG_OFF) +
          ga.peer_ranks[threadIdx.x];
      const long long t0 = clock64();
      while ((unsigned int)bk_ld_acquire_sys_u64(flag) != want) {
        if (clock64() - t0 > BK_P2P_TIMEOUT_CYCLES) {
          s_fail = 1;
          break;
        }
      }
    }
    __syncthreads();
  }
  double acc[R];
#pragma unroll
  for (int i = 0; i < R; ++i) acc[i] = 0.0;
  if (!WAIT || !s_fail) {
    const long long stride = (long long)gridDim.x * BK_BLOCK;
    for (long long b = (long long)blockIdx.x * BK_BLOCK + threadIdx.x; b < ga.n_brows; b += stride) {
      const int r = ga.brow_ids[b];
      T sum = T(0);
      for (int k = ga.rowptr[b]; k < ga.rowptr[b + 1]; ++k)
        sum = fma(ga.val[k], WAIT ? __ldcg(ga.ghost + ga.col[k]) : ga.ghost[ga.col[k]], sum);
      const T y_old = ga.y[r];
      const T y_new = MODE ? bk_sub(y_old, sum) : bk_add(y_old, sum);
      ga.y[r] = y_new;
      int a = 0;
      if (DOTS & 1) acc[a++] += (double)ga.w[r] * (MODE ? -(double)sum : (double)sum);
      if (DOTS & 2) acc[a] += (double)y_new * (double)y_new - (double)y_old * (double)y_old;
    }
  } else if (threadIdx.x == 0) {
    gs.p2p.counters[4] = 1u;
  }
  bk_grid_reduce<R>(acc, sc, [&](const double* s) {
    if (WAIT) {
      gs.p2p.counters[2] = want;
      if (gs.p2p.counters[4]) {
        st->done = 1;
        st->status = BK_ST_COMM_TIMEOUT;
        return;
      }
    }
    if (DOTS != 0) {
      double tot[R], g[R];
#pragma unroll
      for (int i = 0; i < R; ++i) tot[i] = base[i] + s[i];
      if (bk_gsum_finish<R>(gs, st, tot, g)) epi(g);
    }
  });
}

// CG's K3 (x += alpha p ; p = r + beta p, bk_op_cg_xp) with the halo push of the NEW p folded in: elements inside a
// neighbour's range are also stored into that neighbour's ghost vector over NVLink while the kernel streams; the last
// CTA publishes the arrival flags.  In the iteration that ended the solve (`done` just set) only x is still needed:
// nothing is pushed, so pushes and halo waits stay paired on every rank.
template <typename T>
struct bk_push_ranges {
  int nr;
  long long lo[BK_PUSH_MAXR], hi[BK_PUSH_MAXR];
  T* dst[BK_PUSH_MAXR];
};

// lag: 0  x += alpha p ; p = r + beta p (in place: pin == pout)
//      1  even iteration of the lagged-x cut (bk_op_cg_p_lag): pout = r + beta pin — or, when the stop test has just
//         fired, the pending x += alpha pin instead
//      2  odd iteration (bk_op_cg_xp_lag): x = (x + alpha_lag pout) + alpha pin ; pout = r + beta pin  (pout held p_{k-1})
template <typename T>
__global__ void __launch_bounds__(BK_BLOCK, 3)
bk_cg_xp_push_kernel(T* __restrict__ x, const T* pin, T* pout, const T* __restrict__ r, const long long n,
                     const bk_dev_state* st, const bk_push_ranges<T> pr,
                     unsigned long long* const* __restrict__ remote_flag, const int npeers, unsigned int* counters,
                     const int snake, const int lag) {
  const int done = st->done;
  if (done != 0 && st->just_done == 0) return;
  const bool push = (done == 0);
  const bool flush = (lag == 1) && !push;                // x receives the term the even iteration still owes it
  const bool upd_x = (lag != 1) || flush;
  const bool upd_p = !flush;
  (void)snake;
  __shared__ int s_last;
  const T alpha = static_cast<T>(st->alpha), beta = static_cast<T>(st->beta);
  const T alpha_lag = static_cast<T>(st->alpha_lag);
  constexpr int W = bk_native_w<T>::value;
  const long long npack = n / W;
  const long long stride = (long long)gridDim.x * BK_BLOCK;
  // Sweep order: the packs of the push ranges FIRST.  The per-phase trace of the 2-GPU iteration (globaltimer stamps)
  // showed 16 us between the end of this kernel's loop and the start of the next SpMV, 7 of them in the system-scope
  // fence below: with a plain sweep every CTA stores its share of the trailing boundary plane to the neighbour at the
  // very END of its loop and then waits for NVLink to acknowledge.  Rotating the sweep so that it starts at the last
  // range (a suffix of the slab; the first range, a prefix, follows at once) gives those stores the whole kernel to land.
  long long rot = 0;
  if (pr.nr > 0 && pr.hi[pr.nr - 1] >= npack * W) rot = pr.lo[pr.nr - 1] / W;
  auto pack_of = [&](long long kk) -> long long {
    long long p = kk + rot;
    return p >= npack ? p - npack : p;
  };
  // In the rotated order the packs that can hold push elements come first: [0, nbnd); the rest of the sweep runs without
  // the range tests (the trace showed the loop 10 us slower than the plain x/p update).
  long long nbnd = npack;
  if (pr.nr == 0) {
    nbnd = 0;
  } else if (rot > 0) {  // the last range is a suffix; a second range that is a prefix follows it in the rotated order
    nbnd = npack - rot;
    if (pr.nr >= 2) nbnd = (pr.nr == 2 && pr.lo[0] == 0) ? nbnd + (pr.hi[0] + W - 1) / W : npack;
    if (nbnd > npack) nbnd = npack;
  } else if (pr.nr == 1 && pr.lo[0] == 0) {  // a prefix only
    nbnd = (pr.hi[0] + W - 1) / W;
    if (nbnd > npack) nbnd = npack;
  }
  auto one = [&](long long i, T xv, T ppv, T pv, T rv, T& xo, T& po, const bool chk) {
    if (lag == 2) xv = bk_add(xv, bk_mul(alpha_lag, ppv));
    xo = bk_add(xv, bk_mul(alpha, pv));
    po = bk_add(rv, bk_mul(beta, pv));
    if (chk) {
#pragma unroll
      for (int q = 0; q < BK_PUSH_MAXR; ++q)
        if (q < pr.nr && i >= pr.lo[q] && i < pr.hi[q]) pr.dst[q][i - pr.lo[q]] = po;
    }
  };
  // two packs per thread and step, all loads first (like bk_ew_kernel: with one pack in flight the kernel was latency-
  // bound — the per-phase trace of the 2-GPU iteration showed this kernel ~30 % slower than the plain x/p update)
  constexpr int UN = 2;
  const long long tid0 = (long long)blockIdx.x * BK_BLOCK + threadIdx.x;
  for (int phase = 0; phase < 2; ++phase) {
  const long long kend = phase == 0 ? nbnd : npack;
  const bool chk = push && phase == 0;
  long long k = (phase == 0 ? 0 : nbnd) + tid0;
  for (; k + (UN - 1) * stride < kend; k += UN * stride) {
    long long ii[UN];
    bk_vec<T, W> xv[UN], ppv[UN], rv[UN], pv[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long kk = k + u * stride;
      ii[u] = pack_of(kk) * W;
      pv[u] = bk_ld<T, W>(pin + ii[u]);
#pragma unroll
      for (int j = 0; j < W; ++j) xv[u].v[j] = ppv[u].v[j] = rv[u].v[j] = T(0);
      if (upd_x) xv[u] = bk_ld<T, W>(x + ii[u]);
      if (lag == 2) ppv[u] = bk_ld<T, W>(pout + ii[u]);
      if (upd_p) rv[u] = bk_ld<T, W>(r + ii[u]);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      bk_vec<T, W> xo, po;
#pragma unroll
      for (int j = 0; j < W; ++j)
        one(ii[u] + j, xv[u].v[j], ppv[u].v[j], pv[u].v[j], rv[u].v[j], xo.v[j], po.v[j], chk);
      if (upd_x) bk_st<T, W>(x + ii[u], xo);
      if (upd_p) bk_st<T, W>(pout + ii[u], po);
    }
  }
  for (; k < kend; k += stride) {
    const long long i = pack_of(k) * W;
    bk_vec<T, W> xv, ppv, rv;
    const bk_vec<T, W> pv = bk_ld<T, W>(pin + i);
#pragma unroll
    for (int j = 0; j < W; ++j) xv.v[j] = ppv.v[j] = rv.v[j] = T(0);
    if (upd_x) xv = bk_ld<T, W>(x + i);
    if (lag == 2) ppv = bk_ld<T, W>(pout + i);
    if (upd_p) rv = bk_ld<T, W>(r + i);
    bk_vec<T, W> xo, po;
#pragma unroll
    for (int j = 0; j < W; ++j) one(i + j, xv.v[j], ppv.v[j], pv.v[j], rv.v[j], xo.v[j], po.v[j], chk);
    if (upd_x) bk_st<T, W>(x + i, xo);
    if (upd_p) bk_st<T, W>(pout + i, po);
  }
  }
  {
    const long long t = npack * W + (long long)blockIdx.x * BK_BLOCK + threadIdx.x;
    if (t < n) {
      T xo, po;
      one(t, upd_x ? x[t] : T(0), lag == 2 ? pout[t] : T(0), pin[t], upd_p ? r[t] : T(0), xo, po, push);
      if (upd_x) x[t] = xo;
      if (upd_p) pout[t] = po;
    }
  }
  if (!push) return;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicInc(&counters[3], gridDim.x - 1);
    s_last = (t == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    const unsigned int seq = counters[1] + 1u;
    if (threadIdx.x < npeers) bk_st_release_sys_u64(remote_flag[threadIdx.x], (unsigned long long)seq);
    __syncthreads();
    if (threadIdx.x == 0) counters[1] = seq;
  }
}

// Epilogue of the folded SpMV (bk_spmv_mask_kernel<GHOST>): the halo of this matvec is consumed, then the sums become
// global and the solver's scalar step runs — what the boundary-row kernel's epilogue did on the two-kernel path.
template <int R, typename Epi>
struct bk_epi_fold {
  Epi epi;
  bk_gsum gs;
  bk_dev_state* st;
  int waited;
  __device__ __forceinline__ void operator()(const double* s) const {
    if (waited) gs.p2p.counters[2] += 1u;
    if (gs.p2p.counters[4]) {
      st->done = 1;
      st->status = BK_ST_COMM_TIMEOUT;
      return;
    }
    double g[R];
    if (bk_gsum_finish<R>(gs, st, s, g)) epi(g);
  }
};

// ---- the multi-GPU "system" ----------------------------------------------------------------------------------
struct bk_sys_dist {
  static constexpr bool kDist = true;
  bk_handle* h;
  bk_dist* D;
  bool p2p;
  int64_t n_glob;
  bool snake = false;  // CG: alternate the sweep direction of consecutive kernels (st->parity) for L2 reuse

  long long n() const { return D->n_local; }
  long long n_global() const { return n_glob; }
  int dtype() const { return D->dtype; }
  uint64_t uid() const { return D->uid | (1ull << 62) | ((uint64_t)p2p << 61); }
  double matrix_bytes() const {
    return (double)(D->Aloc->nnz + D->nnz_gh) * (bk_dtype_size(D->dtype) + 4) + 4.0 * (D->n_local + 1);
  }
  bk_gsum gsum() const {
    bk_gsum g;
    g.mode = p2p ? 1 : 2;
    g.red = D->red;
    g.p2p = D->p2p;
    return g;
  }
  // sum `count` doubles over the ranks in place (NCCL path only; the peer path reduces inside its kernels)
  int allreduce(double* v, int count, cudaStream_t cs) const {
    BK_NCCL(g_nccl.AllReduce(v, v, (size_t)count, BK_NCCL_F64, BK_NCCL_SUM, D->comm, cs));
    return BK_OK;
  }

  // start the halo exchange of x: peer path = push kernel on `cs`; NCCL path = pack on `cs`, send/recv on the side stream
  template <typename T>
  int halo_begin(const void* x, int guard, bool peer, cudaStream_t cs) const {
    if (D->npeers == 0) return BK_OK;
    bk_dev_state* st = h->st;
    int g = (int)((D->send_total + 255) / 256);
    if (g < 1) g = 1;
    if (peer) {
      if (g > h->num_sms * 2) g = h->num_sms * 2;
      bk_halo_push_kernel<T><<<g, 256, 0, cs>>>((const T*)x, D->send_idx, D->d_seg_start, D->d_remote_ghost,
                                                D->d_remote_flag, D->npeers, D->send_total, D->p2p.counters, st, guard);
      BK_KERNEL_CHECK();
      return BK_OK;
    }
    if (g > h->num_sms * 4) g = h->num_sms * 4;
    if (D->send_total > 0) {
      bk_halo_pack_kernel<T><<<g, 256, 0, cs>>>((const T*)x, D->send_idx, (T*)D->sendbuf, D->send_total, st, guard);
      BK_KERNEL_CHECK();
    }
    BK_CUDA(cudaEventRecord(D->ev_ready, cs));
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
    return BK_OK;
  }

  // boundary rows once the halo is there (NCCL path: stream wait; peer path: the kernel waits for the flags itself)
  template <typename T, int MODE, int KD, typename Epi>
  int ghost_rows(void* y, const void* w, int guard, bool peer, Epi epi, cudaStream_t cs) const {
    if (!peer && D->npeers > 0) BK_CUDA(cudaStreamWaitEvent(cs, D->ev_halo, 0));
    int g = (int)((D->n_brows + BK_BLOCK - 1) / BK_BLOCK);
    if (g < 1) g = 1;
    if (g > h->num_sms * 8) g = h->num_sms * 8;  // one or two rows per thread: the kernel is a chain of dependent loads
    bk_ghost_args<T> ga;
    ga.brow_ids = D->brow_ids;
    ga.rowptr = D->gh_rowptr;
    ga.col = D->gh_col;
    ga.val = (const T*)D->gh_val;
    ga.ghost = (const T*)D->ghost;
    ga.y = (T*)y;
    ga.w = (const T*)w;
    ga.n_brows = D->n_brows;
    ga.peer_ranks = D->d_peer_ranks;
    ga.npeers = D->npeers;
    bk_gsum gs = gsum();
    if (!peer) gs.mode = 2;
    const double* part = D->red + 8;
    if (peer) {
      bk_ghost_rows_kernel<T, MODE, KD, true, Epi><<<g, BK_BLOCK, 0, cs>>>(ga, bk_slot(h, 3), part, gs, h->st, guard, epi);
    } else if (KD != 0 || D->n_brows > 0) {
      bk_ghost_rows_kernel<T, MODE, KD, false, Epi><<<g, BK_BLOCK, 0, cs>>>(ga, bk_slot(h, 3), part, gs, h->st, guard, epi);
    }
    BK_KERNEL_CHECK();
    return BK_OK;
  }

  // y = A x without any reduction (bk_dist_spmv).  Always the NCCL exchange: the peer path needs an all-reduce
  // between two halo exchanges to know that every neighbour has consumed the previous one.
  template <typename T>
  int spmv(const void* x, void* y, cudaStream_t cs) const {
    BK_TRY(halo_begin<T>(x, 0, false, cs));
    bk_spmv_args a = bk_spmv_base(D->Aloc, h->st);
    a.x = x;
    a.y = y;
    BK_TRY((bk_launch_spmv_t<T, 0, 0, 0>(h, D->Aloc, a, bk_slot(h, 0), bk_epi_none(), cs)));
    return ghost_rows<T, 0, 0>(y, nullptr, 0, false, bk_epi_none(), cs);
  }

  // halo_pushed: the kernel that produced x already stored the boundary entries into the neighbours (fused push)
  template <typename T, int MODE, int DOTS, typename Epi>
  int matvec(const void* x, void* y, const void* w, const void* b, int guard, Epi epi, cudaStream_t cs,
             bool halo_pushed = false) const {
    static_assert(DOTS != 0, "distributed solver matvecs always carry a reduction");
    static_assert(MODE == 0 || DOTS == 2, "residual matvecs carry ||y||^2 only");
    // A residual's norm is tiny next to the pieces it is assembled from, so correcting the local block's ||y||^2
    // partial for the boundary rows would cancel catastrophically: reduce it in a pass of its own once y is complete.
    bk_dev_state* st = h->st;
    if (p2p && D->Aext != nullptr && h->dist_fold) {
      // ---- folded: ONE kernel over [local | ghost]; boundary chunks last, after the arrival flags ----------------
      if (!halo_pushed) BK_TRY(halo_begin<T>(x, guard, true, cs));
      constexpr int RF = bk_ndots<DOTS>::value;
      const bk_csr* E = D->Aext;
      bk_spmv_args a = bk_spmv_base(E, st);
      a.x = x;
      a.y = y;
      a.w = w;
      a.b = b;
      a.guard = guard;
      a.use_parity = snake ? 1 : 0;
      bk_mask_plan plan;
      memset(&plan, 0, sizeof(plan));
      plan.masks = E->mmasks;
      plan.pids = E->mpids;
      plan.ptab = (const bk_pair_entry*)E->mptab;
      plan.xg = D->ghost;
      plan.deferred = E->mdeferred;
      plan.n_deferred = E->n_mdeferred;
      {
        int gs_ = 1;
        while ((2 << gs_) <= h->mask_group && gs_ < 5) ++gs_;
        plan.group = gs_;
      }
      plan.prefetch = (h->mask_prefetch && bk_aligned16(x)) ? 1 : 0;
      if (D->npeers > 0) {
        plan.flags = reinterpret_cast<const unsigned long long*>(D->p2p.win[D->p2p.rank] + BK_P2P_FLAG_OFF);
        plan.flag_peers = D->d_peer_ranks;
        plan.n_flag_peers = D->npeers;
        plan.halo_seq = D->p2p.counters + 2;
        plan.err_flag = D->p2p.counters + 4;
      }
      bk_epi_fold<RF, Epi> fe{epi, gsum(), st, D->npeers > 0 ? 1 : 0};
      if constexpr (std::is_same<T, double>::value) {
        // kernel 7 on the interior steps (bk_spmv_mask.cuh), kernel 6's chunk code on the chunks with ghost entries
        if (bk_mask2_usable(h, E) && bk_aligned16(x) && bk_aligned16(y) && (MODE == 0 || bk_aligned16(b)) &&
            ((DOTS & 1) == 0 || bk_aligned16(w))) {
          bk_mask_utab ct;
          memcpy(&ct, E->mctab, sizeof(ct));
          bk_mask2_plan up;
          up.usum = E->musum;
          up.umasks = E->mumasks;
          plan.prefetch = 0;
          int g7 = h->num_sms * 4;
          if (g7 > BK_MAXB) g7 = BK_MAXB;
          g7 = bk_grid_rows(g7, E->n, BK_BLOCK << 3);
          if (E->mu_len == 7)
            bk_spmv_mask2_kernel<MODE, DOTS, 7, 0x14u, true, 4, bk_epi_fold<RF, Epi>><<<g7, BK_BLOCK, 0, cs>>>(
                a, plan, up, ct, bk_slot(h, 0), fe);
          else
            bk_spmv_mask2_kernel<MODE, DOTS, 5, 0x0au, true, 4, bk_epi_fold<RF, Epi>><<<g7, BK_BLOCK, 0, cs>>>(
                a, plan, up, ct, bk_slot(h, 0), fe);
          BK_KERNEL_CHECK();
          return BK_OK;
        }
      }
      int g = h->num_sms * 4;
      if (g > BK_MAXB) g = BK_MAXB;
      g = bk_grid_rows(g, E->n, BK_BLOCK << plan.group);
      bk_spmv_mask_kernel<T, MODE, DOTS, true, 4, bk_epi_fold<RF, Epi>><<<g, BK_BLOCK, 0, cs>>>(a, plan, bk_slot(h, 0), fe);
      BK_KERNEL_CHECK();
      return BK_OK;
    }
    constexpr int KD = (MODE == 1) ? 0 : DOTS;
    constexpr int R = bk_ndots<KD>::value;
    if (!halo_pushed) BK_TRY(halo_begin<T>(x, guard, p2p, cs));
    {  // local block (overlaps the exchange)
      bk_spmv_args a = bk_spmv_base(D->Aloc, st);
      a.x = x;
      a.y = y;
      a.w = w;
      a.b = b;
      a.guard = guard;
      a.use_parity = snake ? 1 : 0;
      if constexpr (KD != 0) {
        bk_epi_store_n<R> store{D->red + 8};
        BK_TRY((bk_launch_spmv_t<T, MODE, KD, 0>(h, D->Aloc, a, bk_slot(h, 0), store, cs)));
      } else {
        BK_TRY((bk_launch_spmv_t<T, MODE, 0, 0>(h, D->Aloc, a, bk_slot(h, 0), bk_epi_none(), cs)));
      }
    }
    BK_TRY((ghost_rows<T, MODE, KD>(y, w, guard, p2p, epi, cs)));
    if constexpr (KD == 0) {
      bk_op_dot_epi<T, Epi> op;
      op.x = (const T*)y;
      op.y = (const T*)y;
      op.epi = epi;
      op.st = st;
      op.guard = guard;
      return ew<T>(op, bk_aligned16(y), 1, cs);
    } else {
      if (!p2p) {
        BK_TRY(allreduce(D->red, R, cs));
        bk_epi_kernel<Epi><<<1, 1, 0, cs>>>(epi, D->red, st, guard);
        BK_KERNEL_CHECK();
      }
      return BK_OK;
    }
  }

  template <typename T, typename Op>
  int ew(const Op& op, bool aligned, int slot, cudaStream_t cs) const {
    if constexpr (Op::R == 0) {
      return bk_launch_ew<T>(h, op, n(), aligned, bk_slot(h, slot), cs);
    } else {
      bk_op_global<Op> g(op, gsum(), h->st);
      BK_TRY(bk_launch_ew<T>(h, g, n(), aligned, bk_slot(h, slot), cs));
      if (!p2p) {
        BK_TRY(allreduce(D->red, Op::R, cs));
        bk_op_epilogue_kernel<Op><<<1, 1, 0, cs>>>(op, D->red);
        BK_KERNEL_CHECK();
      }
      return BK_OK;
    }
  }

  template <typename T, typename Epi>
  int dot(const void* a, const void* b, Epi epi, int slot, cudaStream_t cs) const {
    bk_op_dot_epi<T, Epi> op;
    op.x = (const T*)a;
    op.y = (const T*)b;
    op.epi = epi;
    return ew<T>(op, bk_aligned16(a) && bk_aligned16(b), slot, cs);
  }

  bool can_fuse_push() const { return p2p && D->npeers > 0 && D->push_nranges >= 0; }
  template <typename T>
  int cg_xp_push(T* x, const T* pin, T* pout, const T* r, int lag, cudaStream_t cs) const {
    bk_push_ranges<T> pr;
    pr.nr = D->push_nranges;
    for (int q = 0; q < BK_PUSH_MAXR; ++q) {
      pr.lo[q] = q < pr.nr ? D->push_lo[q] : 0;
      pr.hi[q] = q < pr.nr ? D->push_hi[q] : 0;
      pr.dst[q] = q < pr.nr ? (T*)D->push_dst[q] : nullptr;
    }
    constexpr int NW = bk_native_w<T>::value;
    const int grid = bk_grid_vec_n(h, n(), 2 * NW);
    bk_cg_xp_push_kernel<T><<<grid, BK_BLOCK, 0, cs>>>(x, pin, pout, r, n(), h->st, pr, D->d_remote_flag, D->npeers,
                                                      D->p2p.counters, snake ? 1 : 0, lag);
    BK_KERNEL_CHECK();
    return BK_OK;
  }

  int check_comm(const bk_dev_state* fin, const char* who) const {
    if (fin->status == BK_ST_COMM_TIMEOUT)
      return bk_fail(BK_ERR_NCCL, "%s: a peer did not arrive within the timeout (peer-memory path)", who);
    return BK_OK;
  }
};
