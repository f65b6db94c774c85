// bk_host.cu — end-to-end entry from HOST buffers: H2D of the CSR arrays and b, the device-resident
// solve, D2H of x.  This is what the reference-facing Python API calls for CPU tensors and what
// bench.py times as `e2e` (host<->device copies inside the timed region).  It is still the CUDA path:
// there is no CPU solver in this library.
#include "bk_internal.cuh"

#include <chrono>
#include <cstdlib>

extern "C" int bk_solve_host(bk_handle* h, int method, int64_t n, int64_t nnz, const void* rowptr, const void* col,
                             int idx_bits, const void* val, int dtype, const void* b, void* x_inout, int has_x0,
                             double tol, double atol, int64_t maxiter, int restart, int gmres_method,
                             bk_result* result) {
  if (!h || !result) return bk_fail(BK_ERR_ARG, "bk_solve_host: null handle/result");
  if (n < 0 || nnz < 0 || !rowptr || (nnz > 0 && (!col || !val)) || (n > 0 && (!b || !x_inout)))
    return bk_fail(BK_ERR_ARG, "bk_solve_host: bad argument");
  if (idx_bits != 32 && idx_bits != 64) return bk_fail(BK_ERR_ARG, "bk_solve_host: idx_bits must be 32 or 64");
  if (dtype != BK_F64 && dtype != BK_F32) return bk_fail(BK_ERR_ARG, "bk_solve_host: bad dtype");
  if (method < 0 || method > 2) return bk_fail(BK_ERR_ARG, "bk_solve_host: method must be 0 (cg), 1 (bicgstab), 2 (gmres)");
  BK_CUDA(cudaSetDevice(h->device));
  memset(result, 0, sizeof(*result));
  if (n == 0) return BK_OK;
  cudaStream_t s = h->io_stream;
  const size_t is = idx_bits / 8, vs = bk_dtype_size(dtype);
  bk_csr* A = nullptr;
  int rc = BK_OK;
  // device staging area cached in the handle (grow-only): repeated host solves pay no cudaMalloc/cudaFree
  const size_t nn = (size_t)(nnz > 0 ? nnz : 1);
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t conv = (idx_bits == 64) ? al(4 * (size_t)(n + 1)) + al(4 * nn) : 0;  // int32 copies of the indices
  const size_t need = al(is * (size_t)(n + 1)) + al(is * nn) + al(vs * nn) + 2 * al(vs * (size_t)n) + conv;
  if (need > h->stage_bytes) {
    if (h->stage) {
      BK_CUDA(cudaDeviceSynchronize());
      cudaFree(h->stage);
      h->stage = nullptr;
      h->stage_bytes = 0;
    }
    if (cudaMalloc(&h->stage, need) != cudaSuccess) {
      cudaGetLastError();
      return bk_fail(BK_ERR_ALLOC, "bk_solve_host: device staging allocation of %zu bytes failed", need);
    }
    h->stage_bytes = need;
  }
  char* base = (char*)h->stage;
  void* d_rp = base;
  void* d_col = base + al(is * (size_t)(n + 1));
  void* d_val = (char*)d_col + al(is * nn);
  void* d_b = (char*)d_val + al(vs * nn);
  void* d_x = (char*)d_b + al(vs * (size_t)n);
  void* d_rp32 = (char*)d_x + al(vs * (size_t)n);
  void* d_col32 = (char*)d_rp32 + al(4 * (size_t)(n + 1));
  auto cleanup = [&]() {
    if (A) bk_csr_destroy(A);
  };
  // BK_HOST_TIMING=1: per-phase wall times on stderr (adds stream syncs; diagnosis only)
  static const bool timing = getenv("BK_HOST_TIMING") != nullptr && atoi(getenv("BK_HOST_TIMING")) != 0;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_prev = timing ? now() : 0.0;
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(s);
    const double t = now();
    fprintf(stderr, "[bk_solve_host] %-14s %8.2f ms\n", what, t - t_prev);
    t_prev = t;
  };
  cudaError_t e = cudaMemcpyAsync(d_rp, rowptr, is * (size_t)(n + 1), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && nnz > 0) e = cudaMemcpyAsync(d_col, col, is * (size_t)nnz, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && nnz > 0) e = cudaMemcpyAsync(d_val, val, vs * (size_t)nnz, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_b, b, vs * (size_t)n, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && has_x0) e = cudaMemcpyAsync(d_x, x_inout, vs * (size_t)n, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) {
    cleanup();
    return bk_fail(BK_ERR_CUDA, "bk_solve_host: H2D copy failed: %s", cudaGetErrorString(e));
  }
  lap("h2d");
  if (idx_bits == 64) {  // narrow the indices inside the staging area (overlaps the value copy still in flight)
    if (n >= 2147483647LL || nnz >= 2147483647LL) {
      cleanup();
      return bk_fail(BK_ERR_UNSUPPORTED, "bk_solve_host: n and nnz must be < 2^31");
    }
    bk_convert_i64_i32(h, d_rp, d_rp32, n + 1, s);
    bk_convert_i64_i32(h, d_col, d_col32, nnz, s);
    rc = bk_csr_create(h, n, nnz, d_rp32, d_col32, 32, d_val, dtype, 0, s, &A);
  } else {
    rc = bk_csr_create(h, n, nnz, d_rp, d_col, 32, d_val, dtype, 0, s, &A);
  }
  lap("registration");
  if (rc == BK_OK) {
    if (method == 0)
      rc = bk_cg(h, A, d_b, d_x, has_x0, tol, atol, maxiter, result, s);
    else if (method == 1)
      rc = bk_bicgstab(h, A, d_b, d_x, has_x0, tol, atol, maxiter, result, s);
    else
      rc = bk_gmres(h, A, d_b, d_x, has_x0, tol, atol, restart, maxiter, gmres_method, result, s);
  }
  lap("solve");
  if (rc == BK_OK) {
    e = cudaMemcpyAsync(x_inout, d_x, vs * (size_t)n, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    lap("d2h");
    if (e != cudaSuccess) rc = bk_fail(BK_ERR_CUDA, "bk_solve_host: D2H copy failed: %s", cudaGetErrorString(e));
  }
  cleanup();
  return rc;
}
