// bk_sys.cuh — the small interface the Krylov drivers (BiCGStab, GMRES) are written against, so that ONE driver
// body serves one GPU (bk_sys_local, here) and a row-partitioned matrix (bk_sys_dist, bk_dist.cuh):
//   matvec<T, MODE, DOTS>(x, y, w, b, guard, epi)   y = A x | b - A x, fused dots, epi(GLOBAL sums) on one thread
//   ew<T>(op, aligned, slot)                        fused BLAS-1 pass; op.epilogue(GLOBAL sums)
//   dot<T>(a, b, epi, slot)                         epi(GLOBAL a.b)
//   gsum() / allreduce()                            for kernels with their own reduction tail (GMRES)
// On one GPU the sums a kernel reduces already are global and everything below compiles to the plain launches.
#pragma once

#include "bk_internal.cuh"
#include "bk_p2p.cuh"
#include "bk_spmv.cuh"
#include "bk_vec.cuh"

// An element-wise op whose epilogue first makes its sums global (multi-GPU).
template <typename Op>
struct bk_op_global : Op {
  bk_gsum gs;
  bk_dev_state* gst;
  bk_op_global(const Op& op, const bk_gsum& g, bk_dev_state* st) : Op(op), gs(g), gst(st) {}
  __device__ void epilogue(const double* s) const {
    double g[Op::R > 0 ? Op::R : 1];
    if (bk_gsum_finish<(Op::R > 0 ? Op::R : 1)>(gs, gst, s, g)) Op::epilogue(g);
  }
};

// NCCL path: the scalar step that follows an ncclAllReduce of the parked sums.
template <typename Op>
__global__ void bk_op_epilogue_kernel(const Op op, const double* red) {
  if (op.skip()) return;
  op.epilogue(red);
}
template <typename Epi>
__global__ void bk_epi_kernel(const Epi epi, const double* red, const bk_dev_state* st, int guard) {
  if (bk_guard_skip(st, guard)) return;
  epi(red);
}

struct bk_sys_local {
  static constexpr bool kDist = false;
  bk_handle* h;
  const bk_csr* A;

  long long n() const { return A->n; }
  long long n_global() const { return A->n; }
  int dtype() const { return A->dtype; }
  uint64_t uid() const { return A->uid; }
  double matrix_bytes() const { return (double)A->nnz * (bk_dtype_size(A->dtype) + 4) + 4.0 * (A->n + 1); }
  bk_gsum gsum() const {
    bk_gsum g;
    memset(&g, 0, sizeof(g));
    return g;
  }
  int allreduce(double*, int, cudaStream_t) const { return BK_OK; }

  template <typename T, int MODE, int DOTS, typename Epi>
  int matvec(const void* x, void* y, const void* w, const void* b, int guard, Epi epi, cudaStream_t cs) const {
    bk_spmv_args a = bk_spmv_base(A, h->st);
    a.x = x;
    a.y = y;
    a.w = w;
    a.b = b;
    a.guard = guard;
    return bk_launch_spmv_t<T, MODE, DOTS, 0>(h, A, a, bk_slot(h, 0), epi, cs);
  }
  template <typename T, typename Op>
  int ew(const Op& op, bool aligned, int slot, cudaStream_t cs) const {
    return bk_launch_ew<T>(h, op, A->n, aligned, bk_slot(h, slot), cs);
  }
  template <typename T, typename Epi>
  int dot(const void* a, const void* b, Epi epi, int slot, cudaStream_t cs) const {
    return bk_dot_epi<T>(h, A->n, a, b, epi, slot, cs);
  }
  int check_comm(const bk_dev_state*, const char*) const { return BK_OK; }
};
