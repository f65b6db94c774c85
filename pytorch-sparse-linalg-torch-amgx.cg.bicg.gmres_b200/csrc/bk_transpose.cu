// bk_transpose.cu — cached device transpose (CSC of A == CSR of A^T) for the adjoint solve.
// Replaces `A_matrix.T` in ImplicitAdjointFunction.backward (reference torch_sparse_linalg.py:1245),
// which raises for CSR tensors on torch 2.11; the adjoint Krylov solve then runs the same SpMV
// kernels on the transposed arrays.
//
// Method: the entries of A in CSR order are (row-sorted) triples; A^T in CSR order is the same
// triples STABLY sorted by column.  A hand-written LSD radix sort (8-bit digits, only as many
// passes as the column index needs) on (col, position) pairs does that:
//   pass = [per-tile digit histogram] -> [exclusive scan over (digit, tile)] -> [stable scatter]
// The scatter ranks equal digits inside a tile in input order (warp match + per-warp counters +
// cross-warp prefix), so the result is a deterministic function of the input — no atomics decide
// any position.  Row pointers of A^T and the row index of every moved entry come from binary
// searches.  Set-up cost only: one transpose per matrix, cached on the bk_csr.
#include "bk_internal.cuh"

#define RS_ITEMS 8
#define RS_TILE (BK_BLOCK * RS_ITEMS)

__global__ void bk_iota_kernel(int* __restrict__ out, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (int)i;
}

// hist[d * ntiles + tile] = number of keys of this tile whose digit is d
__global__ void __launch_bounds__(BK_BLOCK)
bk_rs_hist_kernel(const int* __restrict__ keys, long long n, int shift, int ntiles, unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * RS_TILE;
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    const long long i = base + (long long)it * BK_BLOCK + threadIdx.x;
    if (i < n) atomicAdd(&sh[(keys[i] >> shift) & 255], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = sh[threadIdx.x];
}

// ---- exclusive scan of a uint32 array (three small phases; total fits in uint32 since nnz < 2^31) ----
#define SC_CHUNK 4096
__global__ void __launch_bounds__(BK_BLOCK)
bk_scan_sums_kernel(const unsigned int* __restrict__ in, long long n, unsigned int* __restrict__ sums) {
  __shared__ unsigned int sh[BK_WARPS];
  const long long base = (long long)blockIdx.x * SC_CHUNK;
  unsigned int a = 0;
  for (int i = threadIdx.x; i < SC_CHUNK; i += BK_BLOCK) {
    const long long j = base + i;
    if (j < n) a += in[j];
  }
  for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = 0;
    for (int w = 0; w < BK_WARPS; ++w) t += sh[w];
    sums[blockIdx.x] = t;
  }
}

__global__ void bk_scan_serial_kernel(unsigned int* sums, long long m) {  // one block, exclusive, in place
  __shared__ unsigned int carry;
  __shared__ unsigned int sh[BK_BLOCK];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (long long base = 0; base < m; base += BK_BLOCK) {
    const long long i = base + threadIdx.x;
    const unsigned int v = (i < m) ? sums[i] : 0u;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < BK_BLOCK; o <<= 1) {  // Hillis-Steele inclusive
      unsigned int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0u;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    const unsigned int incl = sh[threadIdx.x];
    if (i < m) sums[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == BK_BLOCK - 1) carry += incl;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(BK_BLOCK)
bk_scan_apply_kernel(unsigned int* __restrict__ data, long long n, const unsigned int* __restrict__ sums) {
  // exclusive scan of one SC_CHUNK chunk, offset by the scanned chunk sum; thread t owns 16 consecutive items
  __shared__ unsigned int sh[BK_BLOCK];
  const long long base = (long long)blockIdx.x * SC_CHUNK;
  constexpr int PER = SC_CHUNK / BK_BLOCK;
  unsigned int v[PER];
  unsigned int tot = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const long long j = base + (long long)threadIdx.x * PER + k;
    v[k] = (j < n) ? data[j] : 0u;
    tot += v[k];
  }
  sh[threadIdx.x] = tot;
  __syncthreads();
  for (int o = 1; o < BK_BLOCK; o <<= 1) {
    unsigned int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0u;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  unsigned int run = sums[blockIdx.x] + sh[threadIdx.x] - tot;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const long long j = base + (long long)threadIdx.x * PER + k;
    if (j < n) data[j] = run;
    run += v[k];
  }
}

// In-place exclusive scan of `n` uint32 values (set-up helper shared with the long-row splitter).
int bk_exclusive_scan_u32(unsigned int* data, long long n, cudaStream_t s) {
  if (n <= 0) return BK_OK;
  const long long nchunks = (n + SC_CHUNK - 1) / SC_CHUNK;
  unsigned int* sums = nullptr;
  if (bk_pool_alloc((void**)&sums, sizeof(unsigned int) * (size_t)(nchunks + 1), s) != cudaSuccess)
    return bk_fail(BK_ERR_ALLOC, "scan scratch allocation failed");
  bk_scan_sums_kernel<<<(unsigned)nchunks, BK_BLOCK, 0, s>>>(data, n, sums);
  bk_scan_serial_kernel<<<1, BK_BLOCK, 0, s>>>(sums, nchunks);
  bk_scan_apply_kernel<<<(unsigned)nchunks, BK_BLOCK, 0, s>>>(data, n, sums);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(sums, s);
  if (e != cudaSuccess) return bk_fail(BK_ERR_CUDA, "scan: %s", cudaGetErrorString(e));
  return BK_OK;
}

// stable scatter of one tile: position = scanned_hist[digit][tile] + rank of the item among the
// tile's earlier items with the same digit (tile order = warp, then iteration, then lane).
__global__ void __launch_bounds__(BK_BLOCK)
bk_rs_scatter_kernel(const int* __restrict__ keys_in, const int* __restrict__ vals_in, int* __restrict__ keys_out,
                     int* __restrict__ vals_out, long long n, int shift, int ntiles,
                     const unsigned int* __restrict__ offs) {
  __shared__ unsigned int cnt[BK_WARPS][256];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < BK_WARPS * 256; i += BK_BLOCK) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const long long wbase = (long long)blockIdx.x * RS_TILE + (long long)wid * (32 * RS_ITEMS);
  int key[RS_ITEMS], val[RS_ITEMS];
  unsigned int rank[RS_ITEMS];
  const unsigned int lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    const long long i = wbase + it * 32 + lane;
    const bool valid = i < n;
    key[it] = valid ? keys_in[i] : 0;
    val[it] = valid ? vals_in[i] : 0;
    const int digit = valid ? ((key[it] >> shift) & 255) : (0x1000 | lane);  // invalid lanes match nobody
    const unsigned int peers = __match_any_sync(0xffffffffu, digit);
    const int leader = __ffs(peers) - 1;
    unsigned int old = 0;
    if (valid && lane == leader) {
      old = cnt[wid][digit];
      cnt[wid][digit] = old + __popc(peers);
    }
    __syncwarp();
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[it] = old + __popc(peers & lt_mask);
  }
  __syncthreads();
  {  // thread d: turn per-warp counts of digit d into starting positions
    const int d = threadIdx.x;
    unsigned int run = offs[(size_t)d * ntiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < BK_WARPS; ++w) {
      const unsigned int t = cnt[w][d];
      cnt[w][d] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    const long long i = wbase + it * 32 + lane;
    if (i < n) {
      const int digit = (key[it] >> shift) & 255;
      const unsigned int pos = cnt[wid][digit] + rank[it];
      keys_out[pos] = key[it];
      vals_out[pos] = val[it];
    }
  }
}

// t_rowptr[c] = first position in the sorted keys with key >= c   (c = 0..n)
__global__ void bk_lower_bound_kernel(const int* __restrict__ keys, long long nnz, long long n, int* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c <= n; c += stride) {
    long long lo = 0, hi = nnz;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (keys[mid] < (int)c) lo = mid + 1; else hi = mid;
    }
    out[c] = (int)lo;
  }
}

// for every moved entry q: original position p = perm[q]; its row = upper_bound(rowptr, p) - 1; copy the value
template <typename T>
__global__ void bk_gather_transposed_kernel(const int* __restrict__ perm, const int* __restrict__ rowptr, long long n,
                                            long long nnz, const T* __restrict__ val, int* __restrict__ tcol,
                                            T* __restrict__ tval) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += stride) {
    const int p = perm[q];
    long long lo = 0, hi = n;  // find largest r with rowptr[r] <= p
    while (lo < hi) {
      const long long mid = (lo + hi + 1) >> 1;
      if (rowptr[mid] <= p) lo = mid; else hi = mid - 1;
    }
    // skip empty rows that share the same rowptr value: the largest such r is correct because rowptr[r+1] > p
    tcol[q] = (int)lo;
    tval[q] = val[p];
  }
}

// Stable LSD radix sort of (key, payload) int32 pairs on keys in [0, 2^bits).  Ping-pongs between the caller's two
// buffer pairs; *ks / *vs point at the sorted keys / payloads on return (one of the two pairs).  Asynchronous on `s`.
int bk_sort_pairs_i32(bk_handle* h, int* k0, int* k1, int* v0, int* v1, long long n, int bits, cudaStream_t s,
                      int** ks, int** vs) {
  *ks = k0;
  *vs = v0;
  if (n <= 0) return BK_OK;
  const int ntiles = (int)((n + RS_TILE - 1) / RS_TILE);
  const long long hist_n = 256LL * ntiles;
  const long long nchunks = (hist_n + SC_CHUNK - 1) / SC_CHUNK;
  unsigned int *hist = nullptr, *sums = nullptr;
  if (bk_pool_alloc((void**)&hist, sizeof(unsigned int) * (size_t)hist_n, s) != cudaSuccess ||
      bk_pool_alloc((void**)&sums, sizeof(unsigned int) * (size_t)(nchunks + 1), s) != cudaSuccess) {
    cudaGetLastError();
    if (hist) bk_pool_free(hist);
    return bk_fail(BK_ERR_ALLOC, "radix sort scratch allocation failed");
  }
  int *ki = k0, *ko = k1, *vi = v0, *vo = v1;
  for (int shift = 0; shift < bits; shift += 8) {
    bk_rs_hist_kernel<<<ntiles, BK_BLOCK, 0, s>>>(ki, n, shift, ntiles, hist);
    bk_scan_sums_kernel<<<(unsigned)nchunks, BK_BLOCK, 0, s>>>(hist, hist_n, sums);
    bk_scan_serial_kernel<<<1, BK_BLOCK, 0, s>>>(sums, nchunks);
    bk_scan_apply_kernel<<<(unsigned)nchunks, BK_BLOCK, 0, s>>>(hist, hist_n, sums);
    bk_rs_scatter_kernel<<<ntiles, BK_BLOCK, 0, s>>>(ki, vi, ko, vo, n, shift, ntiles, hist);
    int* t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
  }
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(hist, s);
  cudaFreeAsync(sums, s);
  if (e != cudaSuccess) return bk_fail(BK_ERR_CUDA, "radix sort: %s", cudaGetErrorString(e));
  *ks = ki;
  *vs = vi;
  return BK_OK;
}

// out[c] = first position in the ascending `keys` with key >= c, c = 0..n   (row pointers from sorted row ids)
void bk_lower_bound_i32(bk_handle* h, const int* keys, long long count, long long n, int* out, cudaStream_t s) {
  bk_lower_bound_kernel<<<h->num_sms * 8, 256, 0, s>>>(keys, count, n, out);
}

void bk_iota_i32(bk_handle* h, int* out, long long n, cudaStream_t s) {
  bk_iota_kernel<<<h->num_sms * 8, 256, 0, s>>>(out, n);
}

int bk_csr_finish_plan(bk_handle* h, bk_csr* A, cudaStream_t s);  // bk_core.cu

extern "C" int bk_csr_transpose(bk_handle* h, bk_csr* A, void* stream, bk_csr** out) {
  if (!h || !A || !out) return bk_fail(BK_ERR_ARG, "bk_csr_transpose: null argument");
  if (A->transpose) {
    *out = A->transpose;
    return BK_OK;
  }
  BK_CUDA(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = A->n, nnz = A->nnz;
  bk_csr* Tm = (bk_csr*)calloc(1, sizeof(bk_csr));
  if (!Tm) return bk_fail(BK_ERR_ALLOC, "bk_csr_transpose: host allocation failed");
  Tm->h = h;
  Tm->n = n;
  Tm->nnz = nnz;
  Tm->dtype = A->dtype;
  Tm->uid = h->next_uid++;
  const size_t vs = bk_dtype_size(A->dtype);
  const size_t nn = (size_t)(nnz > 0 ? nnz : 1);
  int *k0 = nullptr, *k1 = nullptr, *v0 = nullptr, *v1 = nullptr;
  auto cleanup = [&]() {
    if (k0) bk_pool_free(k0);
    if (k1) bk_pool_free(k1);
    if (v0) bk_pool_free(v0);
    if (v1) bk_pool_free(v1);
  };
  bool ok = bk_pool_alloc(&Tm->own_rowptr, sizeof(int) * (size_t)(n + 1), s) == cudaSuccess &&
            bk_pool_alloc(&Tm->own_col, sizeof(int) * nn, s) == cudaSuccess &&
            bk_pool_alloc(&Tm->own_val, vs * nn, s) == cudaSuccess &&
            bk_pool_alloc((void**)&k0, sizeof(int) * nn, s) == cudaSuccess &&
            bk_pool_alloc((void**)&k1, sizeof(int) * nn, s) == cudaSuccess &&
            bk_pool_alloc((void**)&v0, sizeof(int) * nn, s) == cudaSuccess &&
            bk_pool_alloc((void**)&v1, sizeof(int) * nn, s) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    cleanup();
    bk_csr_destroy(Tm);
    return bk_fail(BK_ERR_ALLOC, "bk_csr_transpose: device allocation failed (nnz=%lld)", nnz);
  }
  const int g = h->num_sms * 8;
  if (nnz > 0) {
    cudaMemcpyAsync(k0, A->col, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToDevice, s);
    bk_iota_kernel<<<g, 256, 0, s>>>(v0, nnz);
    int bits = 1;
    while ((1LL << bits) < n) ++bits;
    int *ki = nullptr, *vi = nullptr;
    int src = bk_sort_pairs_i32(h, k0, k1, v0, v1, nnz, bits, s, &ki, &vi);
    if (src != BK_OK) {
      cudaStreamSynchronize(s);
      cleanup();
      bk_csr_destroy(Tm);
      return src;
    }
    // ki: sorted columns (= rows of A^T), vi: original positions
    bk_lower_bound_kernel<<<g, 256, 0, s>>>(ki, nnz, n, (int*)Tm->own_rowptr);
    if (A->dtype == BK_F64)
      bk_gather_transposed_kernel<double><<<g, 256, 0, s>>>(vi, A->rowptr, n, nnz, (const double*)A->val,
                                                           (int*)Tm->own_col, (double*)Tm->own_val);
    else
      bk_gather_transposed_kernel<float><<<g, 256, 0, s>>>(vi, A->rowptr, n, nnz, (const float*)A->val,
                                                          (int*)Tm->own_col, (float*)Tm->own_val);
  } else {
    cudaMemsetAsync(Tm->own_rowptr, 0, sizeof(int) * (size_t)(n + 1), s);
  }
  cudaError_t e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  cleanup();
  if (e != cudaSuccess) {
    bk_csr_destroy(Tm);
    return bk_fail(BK_ERR_CUDA, "bk_csr_transpose: %s", cudaGetErrorString(e));
  }
  Tm->rowptr = (const int*)Tm->own_rowptr;
  Tm->col = (const int*)Tm->own_col;
  Tm->val = Tm->own_val;
  int rc = bk_csr_finish_plan(h, Tm, s);
  if (rc != BK_OK) {
    bk_csr_destroy(Tm);
    return rc;
  }
  A->transpose = Tm;
  *out = Tm;
  return BK_OK;
}
