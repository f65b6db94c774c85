// bk_host.cu — end-to-end entry from HOST buffers: H2D of the CSR arrays and b, the device-resident
// solve, D2H of x.  This is what the reference-facing Python API calls for CPU tensors and what
// bench.py times as `e2e` (host<->device copies inside the timed region).  It is still the CUDA path:
// there is no CPU solver in this library.
#include "bk_internal.cuh"

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
  void *d_rp = nullptr, *d_col = nullptr, *d_val = nullptr, *d_b = nullptr, *d_x = nullptr;
  bk_csr* A = nullptr;
  int rc = BK_OK;
  auto cleanup = [&]() {
    if (A) bk_csr_destroy(A);
    if (d_rp) cudaFree(d_rp);
    if (d_col) cudaFree(d_col);
    if (d_val) cudaFree(d_val);
    if (d_b) cudaFree(d_b);
    if (d_x) cudaFree(d_x);
  };
  const size_t nn = (size_t)(nnz > 0 ? nnz : 1);
  if (cudaMalloc(&d_rp, is * (size_t)(n + 1)) != cudaSuccess || cudaMalloc(&d_col, is * nn) != cudaSuccess ||
      cudaMalloc(&d_val, vs * nn) != cudaSuccess || cudaMalloc(&d_b, vs * (size_t)n) != cudaSuccess ||
      cudaMalloc(&d_x, vs * (size_t)n) != cudaSuccess) {
    cudaGetLastError();
    cleanup();
    return bk_fail(BK_ERR_ALLOC, "bk_solve_host: device allocation failed");
  }
  cudaError_t e = cudaMemcpyAsync(d_rp, rowptr, is * (size_t)(n + 1), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && nnz > 0) e = cudaMemcpyAsync(d_col, col, is * (size_t)nnz, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && nnz > 0) e = cudaMemcpyAsync(d_val, val, vs * (size_t)nnz, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_b, b, vs * (size_t)n, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && has_x0) e = cudaMemcpyAsync(d_x, x_inout, vs * (size_t)n, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) {
    cleanup();
    return bk_fail(BK_ERR_CUDA, "bk_solve_host: H2D copy failed: %s", cudaGetErrorString(e));
  }
  rc = bk_csr_create(h, n, nnz, d_rp, d_col, idx_bits, d_val, dtype, 0, s, &A);
  if (rc == BK_OK) {
    if (method == 0)
      rc = bk_cg(h, A, d_b, d_x, has_x0, tol, atol, maxiter, result, s);
    else if (method == 1)
      rc = bk_bicgstab(h, A, d_b, d_x, has_x0, tol, atol, maxiter, result, s);
    else
      rc = bk_gmres(h, A, d_b, d_x, has_x0, tol, atol, restart, maxiter, gmres_method, result, s);
  }
  if (rc == BK_OK) {
    e = cudaMemcpyAsync(x_inout, d_x, vs * (size_t)n, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) rc = bk_fail(BK_ERR_CUDA, "bk_solve_host: D2H copy failed: %s", cudaGetErrorString(e));
  }
  cleanup();
  return rc;
}
