// bk_ingest.cu — dense / COO -> CSR on the device.
// The reference's tests and its LDC example hand the solvers DENSE matrices (test_module_a.py:93-124,
// ldc_solver_module_a.py), and `_normalize_matvec` (torch_sparse_linalg.py:176-208) multiplies with whatever layout
// it gets.  Here every layout becomes the library's CSR once, at registration, with our own kernels:
//   dense : [count non-zeros per row] -> [exclusive scan] -> [ordered compaction]      (two passes over the matrix)
//   COO   : stable radix sort by (row, col) -> runs of equal (row, col) summed in input order (what coalesce()
//           means) -> row pointers by binary search
// Entries are kept iff `!= 0` (NaN is kept), columns ascending — the same structure torch's to_sparse_csr() gives.
#include "bk_internal.cuh"

// ---- dense ---------------------------------------------------------------------------------------------------
template <typename TI>
__global__ void __launch_bounds__(256)
bk_dense_count_kernel(const TI* __restrict__ a, long long n, long long ld, unsigned int* __restrict__ cnt,
                      unsigned long long* __restrict__ total /* 64-bit count: the 32-bit scan below could wrap */) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < n; r += nwarps) {
    const TI* __restrict__ row = a + r * ld;
    unsigned int c = 0;
    for (long long j = lane; j < n; j += 32) c += (row[j] != TI(0)) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) {
      cnt[r] = c;
      if (c) atomicAdd(total, (unsigned long long)c);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) cnt[n] = 0u;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
bk_dense_fill_kernel(const TI* __restrict__ a, long long n, long long ld, const int* __restrict__ rowptr,
                     int* __restrict__ col, TO* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const unsigned int lt = (1u << lane) - 1u;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < n; r += nwarps) {
    const TI* __restrict__ row = a + r * ld;
    int base = rowptr[r];
    for (long long j0 = 0; j0 < n; j0 += 32) {
      const long long j = j0 + lane;
      const TI v = (j < n) ? row[j] : TI(0);
      const bool nz = v != TI(0);
      const unsigned int m = __ballot_sync(0xffffffffu, nz);
      if (nz) {
        const int pos = base + __popc(m & lt);
        col[pos] = (int)j;
        val[pos] = (TO)v;
      }
      base += __popc(m);
    }
  }
}

static bk_csr* bk_new_owned(bk_handle* h, long long n, int dtype) {
  bk_csr* A = (bk_csr*)calloc(1, sizeof(bk_csr));
  if (!A) return nullptr;
  A->h = h;
  A->n = n;
  A->dtype = dtype;
  A->uid = h->next_uid++;
  return A;
}

extern "C" int bk_csr_from_dense(bk_handle* h, int64_t n, const void* dense, int64_t ld, int in_dtype, int dtype,
                                 void* stream, bk_csr** out) {
  if (!h || !out) return bk_fail(BK_ERR_ARG, "bk_csr_from_dense: null handle/out");
  *out = nullptr;
  if (n < 0 || ld < n) return bk_fail(BK_ERR_ARG, "bk_csr_from_dense: bad size (n=%lld ld=%lld)", (long long)n, (long long)ld);
  if (n >= 2147483647LL) return bk_fail(BK_ERR_UNSUPPORTED, "bk_csr_from_dense: n must be < 2^31");
  if (n > 0 && !dense) return bk_fail(BK_ERR_ARG, "bk_csr_from_dense: null matrix");
  if ((in_dtype != BK_F64 && in_dtype != BK_F32) || (dtype != BK_F64 && dtype != BK_F32))
    return bk_fail(BK_ERR_ARG, "bk_csr_from_dense: dtypes must be BK_F64 or BK_F32");
  BK_CUDA(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  bk_csr* A = bk_new_owned(h, n, dtype);
  if (!A) return bk_fail(BK_ERR_ALLOC, "bk_csr_from_dense: host allocation failed");
  auto fail = [&](int code) {
    bk_csr_destroy(A);
    return code;
  };
  if (bk_pool_alloc(&A->own_rowptr, sizeof(int) * (size_t)(n + 1), s) != cudaSuccess) {
    cudaGetLastError();
    return fail(bk_fail(BK_ERR_ALLOC, "bk_csr_from_dense: allocation failed"));
  }
  const int g = h->num_sms * 8;
  cudaMemsetAsync(h->cksum, 0, sizeof(unsigned long long), s);
  if (in_dtype == BK_F64)
    bk_dense_count_kernel<double><<<g, 256, 0, s>>>((const double*)dense, n, ld, (unsigned int*)A->own_rowptr, h->cksum);
  else
    bk_dense_count_kernel<float><<<g, 256, 0, s>>>((const float*)dense, n, ld, (unsigned int*)A->own_rowptr, h->cksum);
  {  // the row-pointer scan is 32-bit: reject matrices with >= 2^31 non-zeros on the 64-bit total BEFORE trusting it
    unsigned long long total = 0;
    cudaMemcpyAsync(&total, h->cksum, sizeof(total), cudaMemcpyDeviceToHost, s);
    cudaError_t e0 = cudaStreamSynchronize(s);
    if (e0 == cudaSuccess) e0 = cudaGetLastError();
    if (e0 != cudaSuccess) return fail(bk_fail(BK_ERR_CUDA, "bk_csr_from_dense: %s", cudaGetErrorString(e0)));
    if (total >= 2147483647ull)
      return fail(bk_fail(BK_ERR_UNSUPPORTED, "bk_csr_from_dense: nnz must be < 2^31 (matrix has %llu non-zeros)", total));
  }
  int rc = bk_exclusive_scan_u32((unsigned int*)A->own_rowptr, n + 1, s);
  if (rc != BK_OK) return fail(rc);
  unsigned int nnz_u = 0;
  cudaMemcpyAsync(&nnz_u, (const int*)A->own_rowptr + n, sizeof(unsigned int), cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(bk_fail(BK_ERR_CUDA, "bk_csr_from_dense: %s", cudaGetErrorString(e)));
  if (nnz_u >= 2147483647u) return fail(bk_fail(BK_ERR_UNSUPPORTED, "bk_csr_from_dense: nnz must be < 2^31"));
  const long long nnz = (long long)nnz_u;
  A->nnz = nnz;
  const size_t vs = bk_dtype_size(dtype);
  if (bk_pool_alloc(&A->own_col, sizeof(int) * (size_t)(nnz > 0 ? nnz : 1), s) != cudaSuccess ||
      bk_pool_alloc(&A->own_val, vs * (size_t)(nnz > 0 ? nnz : 1), s) != cudaSuccess) {
    cudaGetLastError();
    return fail(bk_fail(BK_ERR_ALLOC, "bk_csr_from_dense: allocation failed (nnz=%lld)", nnz));
  }
  if (nnz > 0) {
    const int* rp = (const int*)A->own_rowptr;
    if (in_dtype == BK_F64 && dtype == BK_F64)
      bk_dense_fill_kernel<double, double><<<g, 256, 0, s>>>((const double*)dense, n, ld, rp, (int*)A->own_col, (double*)A->own_val);
    else if (in_dtype == BK_F64)
      bk_dense_fill_kernel<double, float><<<g, 256, 0, s>>>((const double*)dense, n, ld, rp, (int*)A->own_col, (float*)A->own_val);
    else if (dtype == BK_F64)
      bk_dense_fill_kernel<float, double><<<g, 256, 0, s>>>((const float*)dense, n, ld, rp, (int*)A->own_col, (double*)A->own_val);
    else
      bk_dense_fill_kernel<float, float><<<g, 256, 0, s>>>((const float*)dense, n, ld, rp, (int*)A->own_col, (float*)A->own_val);
  }
  A->rowptr = (const int*)A->own_rowptr;
  A->col = (const int*)A->own_col;
  A->val = A->own_val;
  rc = bk_csr_finish_plan(h, A, s);
  if (rc != BK_OK) return fail(rc);
  *out = A;
  return BK_OK;
}

// ---- COO -----------------------------------------------------------------------------------------------------
__global__ void bk_coo_gather_kernel(const int* __restrict__ src, const int* __restrict__ perm, long long n,
                                     int* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = src[perm[i]];
}

// on the ORIGINAL index arrays (before any narrowing to int32, which would hide indices >= 2^32)
template <typename I>
__global__ void bk_coo_range_check_kernel(const I* __restrict__ rows, const I* __restrict__ cols, long long nnz,
                                          long long n, int* __restrict__ bad) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
    const long long r = (long long)rows[i], c = (long long)cols[i];
    if (r < 0 || r >= n || c < 0 || c >= n) *bad = 1;
  }
}

__device__ __forceinline__ bool bk_coo_is_head(const int* srow, const int* scol, long long i) {
  return i == 0 || srow[i] != srow[i - 1] || scol[i] != scol[i - 1];
}

// flag[i] = 1 when sorted entry i starts a new (row, col) run (flag[nnz] = 0 closes the scan)
__global__ void bk_coo_heads_kernel(const int* __restrict__ srow, const int* __restrict__ scol, long long nnz,
                                    unsigned int* __restrict__ flag) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride)
    flag[i] = bk_coo_is_head(srow, scol, i) ? 1u : 0u;
  if (blockIdx.x == 0 && threadIdx.x == 0) flag[nnz] = 0u;
}

// one thread per run: sum its values in sorted (= input) order, emit (row, col, value) at the run's output slot
template <typename TI, typename TO>
__global__ void bk_coo_merge_kernel(const int* __restrict__ srow, const int* __restrict__ scol,
                                    const int* __restrict__ perm, const TI* __restrict__ val, long long nnz,
                                    const unsigned int* __restrict__ slot, int* __restrict__ orow,
                                    int* __restrict__ ocol, TO* __restrict__ oval) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
    if (!bk_coo_is_head(srow, scol, i)) continue;
    TI sum = val[perm[i]];
    for (long long j = i + 1; j < nnz && !bk_coo_is_head(srow, scol, j); ++j) sum += val[perm[j]];
    const unsigned int o = slot[i];
    orow[o] = srow[i];
    ocol[o] = scol[i];
    oval[o] = (TO)sum;
  }
}

extern "C" int bk_csr_from_coo(bk_handle* h, int64_t n, int64_t nnz, const void* rows, const void* cols,
                               int idx_bits, const void* val, int in_dtype, int dtype, void* stream, bk_csr** out) {
  if (!h || !out) return bk_fail(BK_ERR_ARG, "bk_csr_from_coo: null handle/out");
  *out = nullptr;
  if (n < 0 || nnz < 0) return bk_fail(BK_ERR_ARG, "bk_csr_from_coo: negative size");
  if (n >= 2147483647LL || nnz >= 2147483647LL)
    return bk_fail(BK_ERR_UNSUPPORTED, "bk_csr_from_coo: n and nnz must be < 2^31");
  if (idx_bits != 32 && idx_bits != 64) return bk_fail(BK_ERR_ARG, "bk_csr_from_coo: idx_bits must be 32 or 64");
  if (nnz > 0 && (!rows || !cols || !val)) return bk_fail(BK_ERR_ARG, "bk_csr_from_coo: null array");
  if ((in_dtype != BK_F64 && in_dtype != BK_F32) || (dtype != BK_F64 && dtype != BK_F32))
    return bk_fail(BK_ERR_ARG, "bk_csr_from_coo: dtypes must be BK_F64 or BK_F32");
  BK_CUDA(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  bk_csr* A = bk_new_owned(h, n, dtype);
  if (!A) return bk_fail(BK_ERR_ALLOC, "bk_csr_from_coo: host allocation failed");
  const size_t nn = (size_t)(nnz > 0 ? nnz : 1);
  int *r32 = nullptr, *c32 = nullptr, *k0 = nullptr, *k1 = nullptr, *v0 = nullptr, *v1 = nullptr, *scol = nullptr,
      *orow = nullptr;
  unsigned int* slot = nullptr;
  auto cleanup = [&]() {
    void* ps[] = {r32, c32, k0, k1, v0, v1, scol, orow, slot};
    for (void* q : ps)
      if (q) bk_pool_free(q);
  };
  auto fail = [&](int code) {
    cudaStreamSynchronize(s);
    cleanup();
    bk_csr_destroy(A);
    return code;
  };
  bool ok = bk_pool_alloc(&A->own_rowptr, sizeof(int) * (size_t)(n + 1), s) == cudaSuccess &&
            bk_pool_alloc((void**)&k0, sizeof(int) * nn, s) == cudaSuccess &&
            bk_pool_alloc((void**)&k1, sizeof(int) * nn, s) == cudaSuccess &&
            bk_pool_alloc((void**)&v0, sizeof(int) * nn, s) == cudaSuccess &&
            bk_pool_alloc((void**)&v1, sizeof(int) * nn, s) == cudaSuccess &&
            bk_pool_alloc((void**)&scol, sizeof(int) * nn, s) == cudaSuccess &&
            bk_pool_alloc((void**)&orow, sizeof(int) * nn, s) == cudaSuccess &&
            bk_pool_alloc((void**)&slot, sizeof(unsigned int) * (nn + 1), s) == cudaSuccess;
  if (ok && idx_bits == 64)
    ok = bk_pool_alloc((void**)&r32, sizeof(int) * nn, s) == cudaSuccess &&
         bk_pool_alloc((void**)&c32, sizeof(int) * nn, s) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    return fail(bk_fail(BK_ERR_ALLOC, "bk_csr_from_coo: device allocation failed (nnz=%lld)", (long long)nnz));
  }
  const int g = h->num_sms * 8;
  const int* rows32 = (const int*)rows;
  const int* cols32 = (const int*)cols;
  if (nnz == 0) {
    cudaMemsetAsync(A->own_rowptr, 0, sizeof(int) * (size_t)(n + 1), s);
    A->nnz = 0;
    if (bk_pool_alloc(&A->own_col, sizeof(int), s) != cudaSuccess ||
        bk_pool_alloc(&A->own_val, bk_dtype_size(dtype), s) != cudaSuccess) {
      cudaGetLastError();
      return fail(bk_fail(BK_ERR_ALLOC, "bk_csr_from_coo: allocation failed"));
    }
  } else {
    int* bad = (int*)(h->counters + 8);
    cudaMemsetAsync(bad, 0, sizeof(int), s);
    if (idx_bits == 64) {
      bk_coo_range_check_kernel<long long><<<g, 256, 0, s>>>((const long long*)rows, (const long long*)cols, nnz, n, bad);
      bk_convert_i64_i32(h, rows, r32, nnz, s);
      bk_convert_i64_i32(h, cols, c32, nnz, s);
      rows32 = r32;
      cols32 = c32;
    } else {
      bk_coo_range_check_kernel<int><<<g, 256, 0, s>>>(rows32, cols32, nnz, n, bad);
    }
    int bits = 1;
    while ((1LL << bits) < n) ++bits;
    // stable sort by column, then by row: (row, col) order with ties in input order
    cudaMemcpyAsync(k0, cols32, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToDevice, s);
    bk_iota_i32(h, v0, nnz, s);
    int *kc = nullptr, *pc = nullptr;
    int rc = bk_sort_pairs_i32(h, k0, k1, v0, v1, nnz, bits, s, &kc, &pc);
    if (rc != BK_OK) return fail(rc);
    int* kfree = (kc == k0) ? k1 : k0;  // the key buffer the sorted permutation does not live next to
    bk_coo_gather_kernel<<<g, 256, 0, s>>>(rows32, pc, nnz, kfree);
    int* vfree = (pc == v0) ? v1 : v0;
    int *srow = nullptr, *perm = nullptr;
    // second sort: keys = kfree (rows in column order), payload = pc; ping-pong partners are the other two buffers
    rc = bk_sort_pairs_i32(h, kfree, kc, pc, vfree, nnz, bits, s, &srow, &perm);
    if (rc != BK_OK) return fail(rc);
    bk_coo_gather_kernel<<<g, 256, 0, s>>>(cols32, perm, nnz, scol);
    bk_coo_heads_kernel<<<g, 256, 0, s>>>(srow, scol, nnz, slot);
    rc = bk_exclusive_scan_u32(slot, nnz + 1, s);
    if (rc != BK_OK) return fail(rc);
    unsigned int host[2] = {0, 0};
    cudaMemcpyAsync(&host[0], slot + nnz, sizeof(unsigned int), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(&host[1], bad, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(bk_fail(BK_ERR_CUDA, "bk_csr_from_coo: %s", cudaGetErrorString(e)));
    if (host[1]) return fail(bk_fail(BK_ERR_ARG, "bk_csr_from_coo: an index is outside [0, %lld)", (long long)n));
    const long long nout = (long long)host[0];
    A->nnz = nout;
    const size_t vs = bk_dtype_size(dtype);
    if (bk_pool_alloc(&A->own_col, sizeof(int) * (size_t)nout, s) != cudaSuccess ||
        bk_pool_alloc(&A->own_val, vs * (size_t)nout, s) != cudaSuccess) {
      cudaGetLastError();
      return fail(bk_fail(BK_ERR_ALLOC, "bk_csr_from_coo: allocation failed (nnz=%lld)", nout));
    }
    int* ocol = (int*)A->own_col;
    if (in_dtype == BK_F64 && dtype == BK_F64)
      bk_coo_merge_kernel<double, double><<<g, 256, 0, s>>>(srow, scol, perm, (const double*)val, nnz, slot, orow, ocol, (double*)A->own_val);
    else if (in_dtype == BK_F64)
      bk_coo_merge_kernel<double, float><<<g, 256, 0, s>>>(srow, scol, perm, (const double*)val, nnz, slot, orow, ocol, (float*)A->own_val);
    else if (dtype == BK_F64)
      bk_coo_merge_kernel<float, double><<<g, 256, 0, s>>>(srow, scol, perm, (const float*)val, nnz, slot, orow, ocol, (double*)A->own_val);
    else
      bk_coo_merge_kernel<float, float><<<g, 256, 0, s>>>(srow, scol, perm, (const float*)val, nnz, slot, orow, ocol, (float*)A->own_val);
    bk_lower_bound_i32(h, orow, nout, n, (int*)A->own_rowptr, s);
  }
  cudaError_t e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  cleanup();
  if (e != cudaSuccess) {
    bk_csr_destroy(A);
    return bk_fail(BK_ERR_CUDA, "bk_csr_from_coo: %s", cudaGetErrorString(e));
  }
  A->rowptr = (const int*)A->own_rowptr;
  A->col = (const int*)A->own_col;
  A->val = A->own_val;
  int rc = bk_csr_finish_plan(h, A, s);
  if (rc != BK_OK) {
    bk_csr_destroy(A);
    return rc;
  }
  *out = A;
  return BK_OK;
}
