// bk_spmv_tma.cuh — row-stream CSR SpMV with the matrix tiles staged by the TMA engine (sm_100a).
//
// Why: ncu on the first row-stream kernel (profiles/r01_*) showed it bound by the L1TEX data pipe
// (l1tex__data_pipe_lsu_wavefronts 85 % of peak; DRAM only 65 %): the LSU had to move every matrix byte
// (LDG val/col -> registers -> STS products -> LDS) AND serve x-gathers whose lanes were spread over all
// seven stencil diagonals (7-8 cache lines per instruction).  Here:
//   * a producer warp streams each 256-row block's val[] and col[] spans global -> shared memory with
//     cp.async.bulk (1-D TMA, mbarrier complete_tx, L2 evict-first hint): no LSU wavefronts, no registers,
//     NSTAGE tiles in flight per CTA regardless of occupancy;
//   * 8 consumer warps take 32 rows each, lane <-> row: the k-th entries of 32 consecutive rows are read
//     from shared memory at an odd stride (bank-conflict free) and their x-gathers fall on 32 consecutive
//     elements of ONE diagonal (2-3 cache lines per instruction instead of 7-8);
//   * the row sum is a sequential FMA chain in CSR order (deterministic), the dot partials go through the
//     same fixed-order grid reduction, whose last CTA runs the solver's scalar epilogue.
// L1TEX wavefronts per 32-row chunk drop from ~148 to ~45; the kernel becomes DRAM-bound.
//
// Alignment: bulk copies need 16-byte aligned addresses and sizes, so a block's span [s, e) is widened to
// [s & ~3, (e+3) & ~3); the last (nnz % 4) entries of the matrix live in a zero-padded 4-entry tail buffer
// owned by the bk_csr so that no copy ever reads past the caller's arrays.
#pragma once

#include "bk_internal.cuh"

#define BK_TMA_RPB 256                 // rows per block (8 consumer warps x 32 rows)
#define BK_TMA_THREADS (BK_BLOCK + 32) // + one producer warp
#define BK_TMA_MAX_STAGES 8

__device__ __forceinline__ uint32_t bk_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void bk_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bk_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bk_mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bk_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bk_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bk_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bk_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bk_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bk_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t bk_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bk_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          bk_smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(bk_smem_u32(bar)), "l"(policy)
      : "memory");
}

// L2 bulk prefetch (no data returned to the SM; SASS: UBLKPF)
__device__ __forceinline__ void bk_bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}

struct bk_tma_plan {
  int cap;            // entries per stage (multiple of 32)
  int stages;
  int nnz_al;         // nnz rounded down to the copy alignment: entries below come from the main arrays, the rest from the tail buffers
  const void* tail_val;
  const void* tail_idx;
  const void* idx;    // IDX 0: int32 column indices; IDX 1: 8-bit dictionary codes
  const int* dict;    // IDX 1: per 256-row block, 32 int32 (column - row) offsets
  int prefetch_x;     // IDX 1: the producer warp prefetches the x ranges of each block's forward diagonals into L2
};

// IDX selects how column indices are stored for this matrix (decided at registration, bk_csr_plan_tma):
//   0  int32 columns (4 B per entry; the general case)
//   1  one byte per entry: a code into the block's dictionary of distinct (column - row) offsets (<= 32 per 256-row
//      block — every stencil / structured-grid matrix qualifies).  The SpMV then moves nnz*(sizeof T + 1) instead of
//      nnz*(sizeof T + 4) matrix bytes: 20 % less HBM traffic per fp64 SpMV than the algorithmic CSR byte count.
// MODE: 0 y = A x, 1 y = b - A x.   DOTS: bit0 w.y, bit1 y.y.
template <typename T, int MODE, int DOTS, int MINB, int IDX, typename Epi>
__global__ void __launch_bounds__(BK_TMA_THREADS, MINB)
bk_spmv_tma_kernel(const bk_spmv_args a, const bk_tma_plan plan, const bk_scratch sc, Epi epi) {
  if (bk_spmv_skip(a)) return;
  extern __shared__ __align__(128) unsigned char bk_smem_tma[];
  __shared__ __align__(8) uint64_t full_bar[BK_TMA_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[BK_TMA_MAX_STAGES];
  __shared__ int s_base[BK_TMA_MAX_STAGES];
  constexpr int R = bk_ndots<DOTS>::value;
  constexpr int AL = IDX ? 16 : 4;        // entries per 16 bytes of the index stream
  constexpr int IB = IDX ? 1 : 4;         // bytes per index entry
  constexpr int DICT_BYTES = IDX ? 128 : 0;
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int nstage = plan.stages;
  const int cap = plan.cap;
  const size_t stage_bytes = (size_t)cap * (sizeof(T) + IB) + DICT_BYTES;
  const int* __restrict__ rowptr = a.rowptr;
  const long long n = a.n;
  const int nnz = a.nnz;
  const long long nblk = (n + BK_TMA_RPB - 1) / BK_TMA_RPB;
  int reverse = a.reverse;
  if (a.use_parity) reverse ^= (a.st->parity & 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstage; ++s) {
      bk_mbar_init(&full_bar[s], 1);
      bk_mbar_init(&empty_bar[s], BK_WARPS);
    }
    bk_mbar_fence_init();
  }
  __syncthreads();

  double acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.0;

  auto block_of = [&](long long it) -> long long {
    const long long blk = (long long)blockIdx.x + it * gridDim.x;
    return reverse ? (nblk - 1 - blk) : blk;
  };
  const long long my_iters = (nblk > (long long)blockIdx.x) ? (nblk - 1 - blockIdx.x) / gridDim.x + 1 : 0;

  if (wid == BK_WARPS) {
    // ------------------------------ producer warp ------------------------------------------------
    const T* __restrict__ val = static_cast<const T*>(a.val);
    const unsigned char* __restrict__ idx = static_cast<const unsigned char*>(plan.idx);
    const uint64_t pol = bk_policy_evict_first();
    int s_cur = 0, e_cur = 0;  // lane j holds the span of iteration (batch*32 + j)
    // x-gather prefetch (IDX 1): lane l owns dictionary entry l of the block being staged; for every forward
    // diagonal (offset > 0: the entries no earlier block has touched yet) it asks the L2 for the 256-row x range
    // that block will gather from, ~NSTAGE block-times before the consumers need it.
    const T* __restrict__ xpf = static_cast<const T*>(a.x);
    int d_next = 0;
    if (IDX == 1 && plan.prefetch_x && my_iters > 0) d_next = __ldg(plan.dict + block_of(0) * 32 + lane);
    for (long long it0 = 0; it0 < my_iters; it0 += 32) {
      {
        const long long it = it0 + lane;
        if (it < my_iters) {
          const long long r0 = block_of(it) * BK_TMA_RPB;
          const long long r1 = (r0 + BK_TMA_RPB < n) ? r0 + BK_TMA_RPB : n;
          s_cur = __ldg(rowptr + r0);
          e_cur = __ldg(rowptr + r1);
        }
      }
      const int lim = (my_iters - it0 < 32) ? (int)(my_iters - it0) : 32;
      for (int j = 0; j < lim; ++j) {
        const long long it = it0 + j;
        const int s = __shfl_sync(0xffffffffu, s_cur, j);
        const int e = __shfl_sync(0xffffffffu, e_cur, j);
        if constexpr (IDX == 1) {
          if (plan.prefetch_x) {
            const int d = d_next;
            if (it + 1 < my_iters) d_next = __ldg(plan.dict + block_of(it + 1) * 32 + lane);
            if (reverse ? (d < 0) : (d > 0)) {  // diagonals pointing at rows this sweep has not reached yet
              constexpr long long EA = 16 / sizeof(T);                    // elements per 16 bytes
              long long lo = block_of(it) * BK_TMA_RPB + d;
              if (lo < 0) lo = 0;
              lo &= ~(EA - 1);
              long long hi = block_of(it) * BK_TMA_RPB + BK_TMA_RPB + d;
              const long long nal = n & ~(EA - 1);
              if (hi > nal) hi = nal;
              hi = (hi + EA - 1) & ~(EA - 1);
              if (hi > nal) hi = nal;
              if (hi > lo) bk_bulk_prefetch_l2(xpf + lo, (uint32_t)((hi - lo) * sizeof(T)));
            }
          }
        }
        if (lane == 0) {
          const int stage = (int)(it % nstage);
          if (it >= nstage) bk_mbar_wait(&empty_bar[stage], (uint32_t)(((it / nstage) - 1) & 1));
          const int s_al = s & ~(AL - 1);
          int e_al = (e + AL - 1) & ~(AL - 1);
          const bool has_tail = e_al > plan.nnz_al;  // this block reaches the unaligned end of the matrix
          if (has_tail) e_al = plan.nnz_al;
          const int main_cnt = e_al > s_al ? e_al - s_al : 0;
          const int tail_cnt = (has_tail && e > s) ? AL : 0;
          unsigned char* sv = bk_smem_tma + (size_t)stage * stage_bytes;
          unsigned char* sidx = sv + (size_t)cap * sizeof(T);
          s_base[stage] = s_al;
          bk_mbar_expect_tx(&full_bar[stage], (uint32_t)((main_cnt + tail_cnt) * (sizeof(T) + IB) + DICT_BYTES));
          if (main_cnt > 0) {
            bk_bulk_g2s(sv, val + s_al, (uint32_t)(main_cnt * sizeof(T)), &full_bar[stage], pol);
            bk_bulk_g2s(sidx, idx + (size_t)s_al * IB, (uint32_t)(main_cnt * IB), &full_bar[stage], pol);
          }
          if (tail_cnt > 0) {
            const int off = plan.nnz_al - s_al;  // >= 0: s_al <= nnz_al whenever the block reaches the tail
            bk_bulk_g2s(sv + (size_t)off * sizeof(T), plan.tail_val, (uint32_t)(AL * sizeof(T)), &full_bar[stage], pol);
            bk_bulk_g2s(sidx + (size_t)off * IB, plan.tail_idx, (uint32_t)(AL * IB), &full_bar[stage], pol);
          }
          if constexpr (IDX == 1) {
            bk_bulk_g2s(sidx + (size_t)cap * IB, plan.dict + block_of(it) * 32, 128u, &full_bar[stage], pol);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ consumer warps -----------------------------------------------
    // (32-bit row / block arithmetic: n < 2^31 is guaranteed by bk_csr_create)
    const T* __restrict__ x = static_cast<const T*>(a.x);
    T* __restrict__ y = static_cast<T*>(a.y);
    const int n32 = (int)n;
    const int nblk32 = (int)nblk;
    const int iters32 = (int)my_iters;
    const int gstep = (int)gridDim.x;
    const int lane_row = wid * 32 + lane;
    auto row_of = [&](int it) -> int {
      const int blk = (int)blockIdx.x + it * gstep;
      return (reverse ? (nblk32 - 1 - blk) : blk) * BK_TMA_RPB + lane_row;
    };
    int rs = nnz, re = nnz;
    if (iters32 > 0) {
      const int r = row_of(0);
      if (r < n32) {
        rs = __ldg(rowptr + r);
        re = __ldg(rowptr + r + 1);
      }
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < iters32; ++it) {
      const int row = row_of(it);
      int rs_n = nnz, re_n = nnz;
      if (it + 1 < iters32) {  // prefetch the next block's row extents
        const int r = row_of(it + 1);
        if (r < n32) {
          rs_n = __ldg(rowptr + r);
          re_n = __ldg(rowptr + r + 1);
        }
      }
      bk_mbar_wait(&full_bar[stage], phase);
      const unsigned char* sbase = bk_smem_tma + (size_t)stage * stage_bytes;
      const T* __restrict__ sval = reinterpret_cast<const T*>(sbase);
      const unsigned char* __restrict__ sidx = sbase + (size_t)cap * sizeof(T);
      const int* __restrict__ sdict = reinterpret_cast<const int*>(sidx + (size_t)cap * IB);
      const int off = rs - s_base[stage];
      const int len = re - rs;
      T sum = T(0);
      // Branch-free batches of 8 entries: every shared-memory load of a batch is issued before any dependent one
      // (slots past the row end re-read the row's last entry and contribute 0), then all 8 gathers, then the FMA chain.
      const int last = off + len - 1;
      for (int k0 = 0; k0 < len; k0 += 8) {
        int kk[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) kk[u] = min(off + k0 + u, last);
        int c[8];
        if constexpr (IDX == 1) {
          unsigned int code[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) code[u] = sidx[kk[u]];
#pragma unroll
          for (int u = 0; u < 8; ++u) c[u] = row + sdict[code[u] & 31u];
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) c[u] = reinterpret_cast<const int*>(sidx)[kk[u]];
        }
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = sval[kk[u]];
        T xv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) xv[u] = __ldg(x + c[u]);
#pragma unroll
        for (int u = 0; u < 8; ++u) sum = fma((k0 + u < len) ? v[u] : T(0), xv[u], sum);
      }
      __syncwarp();
      if (lane == 0) bk_mbar_arrive(&empty_bar[stage]);
      if (row < n32) {
        T out = sum;
        if constexpr (MODE == 1) out = bk_sub(__ldg(static_cast<const T*>(a.b) + row), sum);
        y[row] = out;
        if constexpr ((DOTS & 1) != 0)
          acc[0] += static_cast<double>(__ldg(static_cast<const T*>(a.w) + row)) * static_cast<double>(out);
        if constexpr ((DOTS & 2) != 0) acc[DOTS & 1] += static_cast<double>(out) * static_cast<double>(out);
      }
      rs = rs_n;
      re = re_n;
      if (++stage == nstage) {
        stage = 0;
        phase ^= 1u;
      }
    }
  }
  if constexpr (DOTS != 0) {
    bk_grid_reduce<R, Epi, BK_WARPS + 1>(acc, sc, epi);
  }
}
