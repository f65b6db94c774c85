// bk_spmv.cuh — CSR SpMV kernels for sm_100a, fused with the dot products that follow them in
// the Krylov recurrences.  Replaces torch.matmul(CSR, v) (reference torch_sparse_linalg.py:191)
// + the torch.vdot that consumes its result (:845 p.Ap, :910 rhat.q, :926-930 t.s/t.t).
//
// Kernels, chosen per matrix at registration (bk_csr_finish_plan, bk_core.cu) from its row-length statistics and
// from what its entries allow:
//
//  * row-stream (short rows, mean <= 32 nnz: stencils, FEM/FVM):  a warp owns 32 consecutive rows.  Three stagings:
//      - kernel 5, bk_spmv_pair.cuh: one byte per entry — a code of the entry's (column - row, value) pair, sliced-ELL
//        stream, TMA-staged; for constant-coefficient stencils (every BASELINE config)
//      - kernels 2 / 3, bk_spmv_tma.cuh: val/col (or val + 8-bit column codes) spans staged by 1-D TMA bulk copies
//      - kernel 0, below: LDG-staged — the span [rowptr[r0], rowptr[r0+32]) of val/col is read fully coalesced with
//        streaming loads, multiplied by the gathered x[col] and parked in the warp's shared-memory strip; lane l
//        then adds the products of row r0+l in CSR order.  Serves misaligned arrays, rows too wide for a TMA stage
//        and the small-system CG variant that forms p = r + beta p inside the gather.
//  * sub-warp vector (kernel 1; long rows: the dense-ish matrices the reference tests use): LPR lanes per row,
//    strided coalesced reads along the row, fixed shuffle tree.
//  * kernel 4: matrices with a short mean row but a few very long rows run any of the above on a virtual-row view
//    (rows cut at BK_SPLIT_LEN entries) followed by an ordered per-row reduction.
//
// All are persistent (grid = SMs x k, blocked-cyclic over row blocks so the whole chip sweeps one contiguous window
// of the matrix, which keeps the x-gather window L2 resident) and end in bk_grid_reduce, whose last CTA runs the
// solver's scalar epilogue.  Summation order is fixed => bitwise reproducible fp64; kernels 2, 3 and 5 (fma chains in
// CSR order) give identical bits, kernel 0 (rounded products, then adds) agrees with them to the last bit or two.
#pragma once

#include "bk_internal.cuh"

struct bk_spmv_args {
  const int* rowptr;
  const int* col;
  const void* val;
  const void* x;    // SpMV input (XMODE 0) / r (XMODE 1)
  const void* x2;   // XMODE 1: previous p
  void* xout;       // XMODE 1: new p = r + beta p, written for the rows this warp owns
  void* y;
  const void* b;    // MODE 1: y = b - A x
  const void* w;    // DOTS bit0: acc0 += w . y   (XMODE 1 uses the new p instead)
  long long n;
  int nnz;
  const bk_dev_state* st;
  int guard;        // 0 none | 1 st->done | 2 st->done || st->g_cycle_over | 3 st->done || st->exit_early
  int reverse;      // 1: sweep row blocks from the end (snake order for L2 reuse)
  int use_parity;   // 1: reverse ^= st->parity
  int l2_hints;     // kernel 6: bit 3 masks / pattern ids stream through L2 | bit 4 y is stored "evict last"
};

__device__ __forceinline__ bool bk_guard_skip(const bk_dev_state* st, int guard) {
  if (guard == 0) return false;
  if (st->done) return true;
  if (guard == 2 && st->g_cycle_over) return true;
  if (guard == 3 && st->exit_early) return true;
  return false;
}
__device__ __forceinline__ bool bk_spmv_skip(const bk_spmv_args& a) { return bk_guard_skip(a.st, a.guard); }

template <int DOTS>
struct bk_ndots {
  static constexpr int value = ((DOTS & 1) + ((DOTS >> 1) & 1)) > 0 ? ((DOTS & 1) + ((DOTS >> 1) & 1)) : 1;
};

// MODE: 0 y = A x, 1 y = b - A x.   DOTS: bit0 w.y, bit1 y.y.   XMODE: 0 plain, 1 x := r + beta*p_old.
template <typename T, int CAP, int MODE, int DOTS, int XMODE, typename Epi>
__global__ void __launch_bounds__(BK_BLOCK, 4)
bk_spmv_stream_kernel(const bk_spmv_args a, const bk_scratch sc, Epi epi) {
  if (bk_spmv_skip(a)) return;
  extern __shared__ __align__(16) unsigned char bk_smem[];
  constexpr int UN = 8;
  constexpr int R = bk_ndots<DOTS>::value;
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  T* __restrict__ prod = reinterpret_cast<T*>(bk_smem) + wid * CAP;
  const int* __restrict__ rowptr = a.rowptr;
  const int* __restrict__ col = a.col;
  const T* __restrict__ val = static_cast<const T*>(a.val);
  const T* __restrict__ x = static_cast<const T*>(a.x);
  const T* __restrict__ x2 = static_cast<const T*>(a.x2);
  T* __restrict__ y = static_cast<T*>(a.y);
  const long long n = a.n;
  const int nnz = a.nnz;
  const long long nblk = (n + BK_BLOCK - 1) / BK_BLOCK;
  int reverse = a.reverse;
  if (a.use_parity) reverse ^= (a.st->parity & 1);
  T beta = T(0);
  if constexpr (XMODE == 1) beta = static_cast<T>(a.st->beta);

  double acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.0;

  auto row_of = [&](long long blk) -> long long {
    const long long bb = reverse ? (nblk - 1 - blk) : blk;
    return bb * BK_BLOCK + wid * 32 + lane;
  };

  long long blk = blockIdx.x;
  int rs = nnz, re = nnz;
  if (blk < nblk) {
    const long long r = row_of(blk);
    if (r < n) {
      rs = __ldg(rowptr + r);
      re = __ldg(rowptr + r + 1);
    }
  }
  for (; blk < nblk; blk += gridDim.x) {
    const long long row = row_of(blk);
    // prefetch the next block's row extents so their latency hides behind this block
    int rs_n = nnz, re_n = nnz;
    {
      const long long nb = blk + gridDim.x;
      if (nb < nblk) {
        const long long r = row_of(nb);
        if (r < n) {
          rs_n = __ldg(rowptr + r);
          re_n = __ldg(rowptr + r + 1);
        }
      }
    }
    const int s = __shfl_sync(0xffffffffu, rs, 0);
    const int e = __shfl_sync(0xffffffffu, re, 31);
    T sum = T(0);
    for (int ws = s; ws < e; ws += CAP) {
      const int we = (e - ws > CAP) ? ws + CAP : e;
      for (int base = ws; base < we; base += 32 * UN) {
        int c[UN];
        T v[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int i = base + u * 32 + lane;
          const bool p = i < we;
          c[u] = p ? __ldcs(col + i) : -1;
          v[u] = p ? __ldcs(val + i) : T(0);
        }
        T xv[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          xv[u] = T(0);
          if (c[u] >= 0) {
            if constexpr (XMODE == 1) {
              xv[u] = bk_add(__ldg(x + c[u]), bk_mul(beta, __ldg(x2 + c[u])));
            } else {
              xv[u] = __ldg(x + c[u]);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          if (c[u] >= 0) prod[base - ws + u * 32 + lane] = v[u] * xv[u];
        }
      }
      __syncwarp();
      const int lo = rs > ws ? rs : ws;
      const int hi = re < we ? re : we;
      for (int k = lo; k < hi; ++k) sum += prod[k - ws];
      __syncwarp();
    }
    if (row < n) {
      T out = sum;
      if constexpr (MODE == 1) out = bk_sub(__ldg(static_cast<const T*>(a.b) + row), sum);
      y[row] = out;
      T wv = T(0);
      if constexpr (XMODE == 1) {
        wv = bk_add(__ldg(x + row), bk_mul(beta, __ldg(x2 + row)));
        static_cast<T*>(a.xout)[row] = wv;
      } else if constexpr ((DOTS & 1) != 0) {
        wv = __ldg(static_cast<const T*>(a.w) + row);
      }
      if constexpr ((DOTS & 1) != 0) acc[0] += static_cast<double>(wv) * static_cast<double>(out);
      if constexpr ((DOTS & 2) != 0) acc[DOTS & 1] += static_cast<double>(out) * static_cast<double>(out);
    }
    rs = rs_n;
    re = re_n;
  }
  if constexpr (DOTS != 0) {
    bk_grid_reduce<R>(acc, sc, epi);
  }
}

template <typename T, int LPR, int MODE, int DOTS, typename Epi>
__global__ void __launch_bounds__(BK_BLOCK, 4)
bk_spmv_vector_kernel(const bk_spmv_args a, const bk_scratch sc, Epi epi) {
  if (bk_spmv_skip(a)) return;
  constexpr int R = bk_ndots<DOTS>::value;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int sub = lane % LPR;
  const int rw = lane / LPR;
  const int* __restrict__ rowptr = a.rowptr;
  const int* __restrict__ col = a.col;
  const T* __restrict__ val = static_cast<const T*>(a.val);
  const T* __restrict__ x = static_cast<const T*>(a.x);
  T* __restrict__ y = static_cast<T*>(a.y);
  const long long n = a.n;
  constexpr long long RPB = BK_WARPS * RPW;
  const long long nblk = (n + RPB - 1) / RPB;
  int reverse = a.reverse;
  if (a.use_parity) reverse ^= (a.st->parity & 1);

  double acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.0;

  for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const long long bb = reverse ? (nblk - 1 - blk) : blk;
    const long long row = bb * RPB + wid * RPW + rw;
    T sum = T(0);
    if (row < n) {
      const int rs = __ldg(rowptr + row);
      const int re = __ldg(rowptr + row + 1);
#pragma unroll 4
      for (int k = rs + sub; k < re; k += LPR) sum = fma(__ldcs(val + k), __ldg(x + __ldcs(col + k)), sum);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o, LPR);
    if (row < n && sub == 0) {
      T out = sum;
      if constexpr (MODE == 1) out = bk_sub(__ldg(static_cast<const T*>(a.b) + row), sum);
      y[row] = out;
      if constexpr ((DOTS & 1) != 0)
        acc[0] += static_cast<double>(__ldg(static_cast<const T*>(a.w) + row)) * static_cast<double>(out);
      if constexpr ((DOTS & 2) != 0) acc[DOTS & 1] += static_cast<double>(out) * static_cast<double>(out);
    }
  }
  if constexpr (DOTS != 0) {
    bk_grid_reduce<R>(acc, sc, epi);
  }
}

#include "bk_spmv_tma.cuh"
#include "bk_spmv_pair.cuh"
#include "bk_spmv_mask.cuh"

// Second half of the long-row path: y[r] = (b[r] -) sum of the partial sums of r's virtual rows, added in order
// (deterministic), fused with the requested dots and the solver's scalar epilogue.
template <typename T, int MODE, int DOTS, typename Epi>
__global__ void __launch_bounds__(BK_BLOCK)
bk_vrow_reduce_kernel(const bk_spmv_args a, const int* __restrict__ vstart, const T* __restrict__ yv,
                      const bk_scratch sc, Epi epi) {
  if (bk_spmv_skip(a)) return;
  constexpr int R = bk_ndots<DOTS>::value;
  double acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.0;
  T* __restrict__ y = static_cast<T*>(a.y);
  const long long stride = (long long)gridDim.x * BK_BLOCK;
  for (long long row = (long long)blockIdx.x * BK_BLOCK + threadIdx.x; row < a.n; row += stride) {
    T sum = T(0);
    for (int v = vstart[row]; v < vstart[row + 1]; ++v) sum += yv[v];
    T out = sum;
    if constexpr (MODE == 1) out = bk_sub(__ldg(static_cast<const T*>(a.b) + row), sum);
    y[row] = out;
    if constexpr ((DOTS & 1) != 0)
      acc[0] += static_cast<double>(__ldg(static_cast<const T*>(a.w) + row)) * static_cast<double>(out);
    if constexpr ((DOTS & 2) != 0) acc[DOTS & 1] += static_cast<double>(out) * static_cast<double>(out);
  }
  if constexpr (DOTS != 0) {
    bk_grid_reduce<R>(acc, sc, epi);
  }
}

struct bk_epi_none {
  __device__ __forceinline__ void operator()(const double*) const {}
};

// Opt a kernel in to > 48 KB of dynamic shared memory once per (kernel, size) — keyed by the kernel's address
// (all instantiations share one function-pointer TYPE, so a per-type static would be wrong).
static inline int bk_ensure_dyn_smem(const void* func, size_t bytes) {
  // This header is compiled into several translation units, each with its own table, while the attribute is
  // per kernel and process-wide: always opt in to one fixed ceiling (the most a 2-CTA/SM plan can ask for, leaving
  // room for static shared memory) so that no unit can lower another's limit.
  constexpr size_t kMaxOptIn = 112 * 1024;
  static const void* funcs[128];
  static int count = 0;
  if (bytes > kMaxOptIn) return bk_fail(BK_ERR_ARG, "dynamic shared memory request exceeds 112 KB");
  for (int i = 0; i < count; ++i)
    if (funcs[i] == func) return BK_OK;
  BK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxOptIn));
  if (count < 128) funcs[count++] = func;
  return BK_OK;
}

template <typename K>
static inline int bk_set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    BK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  return BK_OK;
}

// kernel 7 runs the structures it is instantiated for: 7 union offsets with the 3rd and 5th odd (3-D 7-point stencil, even
// line length), 5 with the 2nd and 4th odd (2-D 5-point)
static inline bool bk_mask2_usable(const bk_handle* h, const bk_csr* A) {
  if (!h->mask_const || A->musum == nullptr || A->dtype != BK_F64) return false;
  if (!A->mu_center) return false;  // offsets -1, 0, +1 at the union positions K - 1, K, K + 1
  return (A->mu_len == 7 && A->mu_odd == 0x14) || (A->mu_len == 5 && A->mu_odd == 0x0a);
}

template <typename T, int MODE, int DOTS, int XMODE, typename Epi>
static int bk_launch_spmv_t(bk_handle* h, const bk_csr* A, const bk_spmv_args& a, const bk_scratch& sc,
                            Epi epi, cudaStream_t s);

// Skewed matrices: SpMV over the virtual-row view (every lane's work bounded by BK_SPLIT_LEN entries), then the
// ordered per-row reduction of the partial sums.
template <typename T, int MODE, int DOTS, typename Epi>
static int bk_launch_spmv_split(bk_handle* h, const bk_csr* A, const bk_spmv_args& a, const bk_scratch& sc, Epi epi,
                                cudaStream_t s) {
  bk_spmv_args av = a;
  av.rowptr = A->split->rowptr;
  av.n = A->split->n;
  av.y = A->yv;
  av.b = nullptr;
  av.w = nullptr;
  BK_TRY((bk_launch_spmv_t<T, 0, 0, 0, bk_epi_none>(h, A->split, av, sc, bk_epi_none(), s)));
  const int grid = bk_grid_rows(bk_grid_vec(h), A->n, BK_BLOCK);
  bk_vrow_reduce_kernel<T, MODE, DOTS, Epi><<<grid, BK_BLOCK, 0, s>>>(a, A->vstart, (const T*)A->yv, sc, epi);
  BK_KERNEL_CHECK();
  return BK_OK;
}

template <typename T, int MODE, int DOTS, int XMODE, typename Epi>
static int bk_launch_spmv_t(bk_handle* h, const bk_csr* A, const bk_spmv_args& a, const bk_scratch& sc,
                            Epi epi, cudaStream_t s) {
  if (A->split != nullptr) {
    if constexpr (XMODE != 0) {
      return bk_fail(BK_ERR_UNSUPPORTED, "fused p-update is not available for row-split matrices");
    } else {
      return bk_launch_spmv_split<T, MODE, DOTS, Epi>(h, A, a, sc, epi, s);
    }
  }
  if (A->kernel == 6 && XMODE == 0) {
    if constexpr (XMODE == 0) {
      // row-bitmask stream (bk_spmv_mask.cuh): no shared memory, occupancy set by registers
      bk_mask_plan plan;
      memset(&plan, 0, sizeof(plan));
      plan.masks = A->mmasks;
      plan.pids = A->mpids;
      plan.ptab = (const bk_pair_entry*)A->mptab;
      {  // option mask_group = blocks per visit, rounded down to a power of two (the kernel shifts and masks)
        int gs = 1;  // (the kernel works on pairs of consecutive blocks)
        while ((2 << gs) <= h->mask_group && gs < 5) ++gs;
        plan.group = gs;
      }
      plan.prefetch = (h->mask_prefetch && bk_aligned16(a.x)) ? 1 : 0;
      // kernel 6W: gathers from TMA-staged shared-memory windows (x 16-byte aligned, n a multiple of the 16-byte pack)
      if (h->mask_window && A->mw_win > 0 && bk_aligned16(a.x) && (A->n % (16 / sizeof(T))) == 0) {
        int gsw = 0;
        while ((2 << gsw) <= h->mask_wgroup && gsw < 5) ++gsw;
        const int rows_g = 256 << gsw;
        bk_maskw_plan wp;
        wp.win = A->mw_win;
        wp.nfar = A->mw_nfar;
        wp.far_off[0] = A->mw_far[0];
        wp.far_off[1] = A->mw_far[1];
        const size_t raw = (size_t)(rows_g + 2 * wp.win) * sizeof(T) + (size_t)wp.nfar * rows_g * sizeof(T) + rows_g +
                           (size_t)(rows_g >> 8) * 32;
        wp.stage_bytes = (uint32_t)((raw + 127) & ~(size_t)127);
        int wctas = h->mask_ctas < 2 ? 2 : (h->mask_ctas > 4 ? 4 : h->mask_ctas);
        int stages = 0;
        for (; wctas >= 2; --wctas) {
          stages = (int)(((size_t)224 * 1024 / wctas - 2048) / wp.stage_bytes);
          if (stages >= 2) break;
        }
        if (stages >= 2) {
          if (stages > BK_TMA_MAX_STAGES) stages = BK_TMA_MAX_STAGES;
          if (h->tma_stages >= 2 && h->tma_stages < stages) stages = h->tma_stages;
          wp.stages = stages;
          plan.group = gsw;
          const size_t sm = (size_t)stages * wp.stage_bytes;
          int g = h->num_sms * wctas;
          if (g > BK_MAXB) g = BK_MAXB;
          g = bk_grid_rows(g, A->n, rows_g);
          auto launch = [&](auto k) -> int {
            BK_TRY(bk_ensure_dyn_smem((const void*)k, sm));
            k<<<g, BK_TMA_THREADS, sm, s>>>(a, plan, wp, sc, epi);
            return BK_OK;
          };
          switch (wctas) {
            case 2: BK_TRY(launch(bk_spmv_maskw_kernel<T, MODE, DOTS, false, 2, Epi>)); break;
            case 3: BK_TRY(launch(bk_spmv_maskw_kernel<T, MODE, DOTS, false, 3, Epi>)); break;
            default: BK_TRY(launch(bk_spmv_maskw_kernel<T, MODE, DOTS, false, 4, Epi>)); break;
          }
          BK_KERNEL_CHECK();
          return BK_OK;
        }
      }
      if constexpr (std::is_same<T, double>::value) {
        // kernel 7: two rows per lane, 128-bit gathers over the union of the patterns' offsets (operands 16-byte aligned;
        // instantiated for 7-point 3-D and 5-point 2-D stencils with even line lengths)
        if (bk_mask2_usable(h, A) && bk_aligned16(a.x) && bk_aligned16(a.y) && (MODE == 0 || bk_aligned16(a.b)) &&
            ((DOTS & 1) == 0 || bk_aligned16(a.w))) {
          bk_mask_utab ct;
          memcpy(&ct, A->mctab, sizeof(ct));
          bk_mask2_plan up;
          plan.prefetch = (h->mask2_prefetch && bk_aligned16(a.x)) ? 1 : 0;
          up.usum = A->musum;
          up.umasks = A->mumasks;
          const int cctas = h->mask_cctas < 4 ? 4 : (h->mask_cctas > 6 ? 6 : h->mask_cctas);
          int g = h->num_sms * cctas;
          if (g > BK_MAXB) g = BK_MAXB;
          g = bk_grid_rows(g, A->n, BK_BLOCK << 3);
          const bool s7 = A->mu_len == 7;
#define BK_MASK2_LAUNCH(MINB)                                                                                     \
  if (s7) bk_spmv_mask2_kernel<MODE, DOTS, 7, 0x14u, false, MINB, Epi><<<g, BK_BLOCK, 0, s>>>(a, plan, up, ct, sc, epi);    \
  else bk_spmv_mask2_kernel<MODE, DOTS, 5, 0x0au, false, MINB, Epi><<<g, BK_BLOCK, 0, s>>>(a, plan, up, ct, sc, epi);
          switch (cctas) {
            case 4: BK_MASK2_LAUNCH(4) break;
            case 5: BK_MASK2_LAUNCH(5) break;
            default: BK_MASK2_LAUNCH(6) break;
          }
#undef BK_MASK2_LAUNCH
          BK_KERNEL_CHECK();
          return BK_OK;
        }
      }
      int ctas = h->mask_ctas < 3 ? 3 : (h->mask_ctas > 5 ? 5 : h->mask_ctas);
      int g = h->num_sms * ctas;
      if (g > BK_MAXB) g = BK_MAXB;
      g = bk_grid_rows(g, A->n, BK_BLOCK << plan.group);
      switch (ctas) {
        case 3: bk_spmv_mask_kernel<T, MODE, DOTS, false, 3, Epi><<<g, BK_BLOCK, 0, s>>>(a, plan, sc, epi); break;
        case 4: bk_spmv_mask_kernel<T, MODE, DOTS, false, 4, Epi><<<g, BK_BLOCK, 0, s>>>(a, plan, sc, epi); break;
        default: bk_spmv_mask_kernel<T, MODE, DOTS, false, 5, Epi><<<g, BK_BLOCK, 0, s>>>(a, plan, sc, epi); break;
      }
    }
  } else if (A->kernel == 5 && XMODE == 0) {
    if constexpr (XMODE == 0) {
      // pair-coded SELL stream: stages are tiny (~2 KB for a 7-point stencil), so occupancy is set by registers
      const size_t stage_bytes = (size_t)A->pair_cap + BK_PAIR_DICT_BYTES;
      int ctas = h->pair_ctas < 2 ? 2 : (h->pair_ctas > 6 ? 6 : h->pair_ctas);
      int stages = 0;
      for (; ctas >= 2; --ctas) {
        stages = (int)(((size_t)224 * 1024 / ctas - 2048) / stage_bytes);
        if (stages >= 2) break;
      }
      if (stages > BK_TMA_MAX_STAGES) stages = BK_TMA_MAX_STAGES;
      if (h->tma_stages >= 2 && h->tma_stages < stages) stages = h->tma_stages;
      const size_t sm = (size_t)stages * stage_bytes;
      bk_pair_plan plan;
      plan.bptr = A->pbptr;
      plan.codes = A->pcodes;
      plan.dict = (const bk_pair_entry*)A->pdict;
      plan.cap = A->pair_cap;
      plan.stages = stages;
      int g = h->num_sms * ctas;
      if (g > BK_MAXB) g = BK_MAXB;
      g = bk_grid_rows(g, A->n, BK_TMA_RPB);
      auto launch = [&](auto k) -> int {
        BK_TRY(bk_ensure_dyn_smem((const void*)k, sm));
        k<<<g, BK_TMA_THREADS, sm, s>>>(a, plan, sc, epi);
        return BK_OK;
      };
      switch (ctas) {
        case 2: BK_TRY(launch(bk_spmv_pair_kernel<T, MODE, DOTS, 2, Epi>)); break;
        case 3: BK_TRY(launch(bk_spmv_pair_kernel<T, MODE, DOTS, 3, Epi>)); break;
        case 4: BK_TRY(launch(bk_spmv_pair_kernel<T, MODE, DOTS, 4, Epi>)); break;
        case 5: BK_TRY(launch(bk_spmv_pair_kernel<T, MODE, DOTS, 5, Epi>)); break;
        default: BK_TRY(launch(bk_spmv_pair_kernel<T, MODE, DOTS, 6, Epi>)); break;
      }
    }
  } else if ((A->kernel == 2 || A->kernel == 3) && XMODE == 0) {
    if constexpr (XMODE == 0) {
      // CTAs per SM (2..4) trade pipeline depth for consumer warps; stages fill the per-CTA share of shared memory
      int ctas = h->tma_ctas < 2 ? 2 : (h->tma_ctas > 4 ? 4 : h->tma_ctas);
      const bool cmp = (A->kernel == 3);
      const int cap = cmp ? A->cmp_cap : A->tma_cap;
      const size_t stage_bytes = (size_t)cap * (sizeof(T) + (cmp ? 1 : 4)) + (cmp ? 128 : 0);
      int stages = 0;
      for (; ctas >= 2; --ctas) {
        stages = (int)(((size_t)224 * 1024 / ctas - 2048) / stage_bytes);
        if (stages >= 2) break;
      }
      if (stages > BK_TMA_MAX_STAGES) stages = BK_TMA_MAX_STAGES;
      if (h->tma_stages >= 2 && h->tma_stages < stages) stages = h->tma_stages;
      const size_t sm = (size_t)stages * stage_bytes;
      bk_tma_plan plan;
      plan.cap = cap;
      plan.stages = stages;
      plan.nnz_al = (int)(A->nnz & ~(int64_t)(cmp ? 15 : 3));
      plan.tail_val = cmp ? A->tail_val16 : A->tail_val;
      plan.tail_idx = cmp ? (const void*)A->tail_code16 : (const void*)A->tail_col;
      plan.idx = cmp ? (const void*)A->codes : (const void*)A->col;
      plan.dict = A->dict;
      plan.prefetch_x = h->prefetch_x && bk_aligned16(a.x);
      int g = h->num_sms * ctas;
      if (g > BK_MAXB) g = BK_MAXB;
      g = bk_grid_rows(g, A->n, BK_TMA_RPB);
      auto launch = [&](auto k) -> int {
        BK_TRY(bk_ensure_dyn_smem((const void*)k, sm));
        k<<<g, BK_TMA_THREADS, sm, s>>>(a, plan, sc, epi);
        return BK_OK;
      };
      if (cmp) {
        if (ctas == 2) {
          BK_TRY(launch(bk_spmv_tma_kernel<T, MODE, DOTS, 2, 1, Epi>));
        } else if (ctas == 3) {
          BK_TRY(launch(bk_spmv_tma_kernel<T, MODE, DOTS, 3, 1, Epi>));
        } else {
          BK_TRY(launch(bk_spmv_tma_kernel<T, MODE, DOTS, 4, 1, Epi>));
        }
      } else if (ctas == 2) {
        BK_TRY(launch(bk_spmv_tma_kernel<T, MODE, DOTS, 2, 0, Epi>));
      } else if (ctas == 3) {
        BK_TRY(launch(bk_spmv_tma_kernel<T, MODE, DOTS, 3, 0, Epi>));
      } else {
        BK_TRY(launch(bk_spmv_tma_kernel<T, MODE, DOTS, 4, 0, Epi>));
      }
    }
  } else if (A->kernel == 0 || A->kernel == 2 || A->kernel == 3 || A->kernel == 5 || A->kernel == 6) {
    const int grid = bk_grid_rows(bk_grid_spmv(h), A->n, BK_BLOCK);
    if (A->cap <= 256) {
      auto k = bk_spmv_stream_kernel<T, 256, MODE, DOTS, XMODE, Epi>;
      const size_t sm = (size_t)BK_WARPS * 256 * sizeof(T);
      k<<<grid, BK_BLOCK, sm, s>>>(a, sc, epi);
    } else {
      auto k = bk_spmv_stream_kernel<T, 1024, MODE, DOTS, XMODE, Epi>;
      const size_t sm = (size_t)BK_WARPS * 1024 * sizeof(T);
      BK_TRY(bk_set_smem(k, sm));
      k<<<grid, BK_BLOCK, sm, s>>>(a, sc, epi);
    }
  } else {
    if constexpr (XMODE != 0) {
      return bk_fail(BK_ERR_UNSUPPORTED, "fused p-update needs the row-stream SpMV kernel");
    } else {
      const int grid = bk_grid_rows(bk_grid_spmv(h), A->n, BK_WARPS * (32 / A->lanes_per_row));
      switch (A->lanes_per_row) {
        case 8:
          bk_spmv_vector_kernel<T, 8, MODE, DOTS, Epi><<<grid, BK_BLOCK, 0, s>>>(a, sc, epi);
          break;
        case 16:
          bk_spmv_vector_kernel<T, 16, MODE, DOTS, Epi><<<grid, BK_BLOCK, 0, s>>>(a, sc, epi);
          break;
        default:
          bk_spmv_vector_kernel<T, 32, MODE, DOTS, Epi><<<grid, BK_BLOCK, 0, s>>>(a, sc, epi);
          break;
      }
    }
  }
  BK_KERNEL_CHECK();
  return BK_OK;
}

// Fill the matrix part of the argument block.
static inline bk_spmv_args bk_spmv_base(const bk_csr* A, const bk_dev_state* st) {
  bk_spmv_args a;
  memset(&a, 0, sizeof(a));
  a.rowptr = A->rowptr;
  a.col = A->col;
  a.val = A->val;
  a.n = A->n;
  a.nnz = (int)A->nnz;
  a.st = st;
  return a;
}

template <int MODE, int DOTS, int XMODE, typename Epi>
static int bk_launch_spmv(bk_handle* h, const bk_csr* A, const bk_spmv_args& a, const bk_scratch& sc, Epi epi,
                          cudaStream_t s) {
  if (A->dtype == BK_F64) return bk_launch_spmv_t<double, MODE, DOTS, XMODE, Epi>(h, A, a, sc, epi, s);
  return bk_launch_spmv_t<float, MODE, DOTS, XMODE, Epi>(h, A, a, sc, epi, s);
}
