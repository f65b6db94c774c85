// bk_spmv_pair.cuh — SpMV for matrices whose 256-row blocks each hold at most 31 distinct (column - row, value) PAIRS
// (constant-coefficient stencils: Poisson, upwind convection-diffusion, the LDC pressure matrix ...): kernel 5.
//
// At registration (bk_csr_plan_pairs, bk_core.cu) every entry is replaced by ONE BYTE — a code into its block's
// dictionary of pairs — and the codes are stored sliced-ELL style ("SELL-32-4"): for each chunk of 32 rows, padded to
// the longest row rounded up to 4, word j of lane l holds the codes of entries 4j..4j+3 of row chunk*32 + l, words of
// the 32 lanes adjacent.  Padding uses code 31, a dictionary slot that always holds {value 0, offset 0}.  Every
// 256-row block's span starts with a 128-byte header: word w = (end << 16 | start) of chunk w in 128-byte units from
// the span start, so a consumer warp learns its chunk's place from the staged tile itself (one shared load).
//   * the SpMV reads 1 byte per entry instead of sizeof(T) + 4, and neither `val`, `col` nor `rowptr` at all:
//     matrix traffic drops from 12 B to ~1.15 B per entry (P3D-256: 1.40 GB -> 0.17 GB per SpMV), the kernel is then
//     bound by the x / y vectors and by instruction issue, not by the matrix;
//   * a lane fetches 4 codes with one conflict-free 32-bit shared load, each code selects its {value, offset} with
//     one 16-byte shared load (a broadcast when the 32 rows sit on the same stencil diagonal); no masking, no row
//     pointer, no tail handling (chunks are 128-byte multiples, so every bulk copy is aligned);
//   * lossless and bit-identical to the CSR kernels: the same values are multiplied with the same x entries in CSR
//     order (padding adds +0 * x[row]; like the other kernels' padding this turns a non-finite x[row] into NaN).
// Staging is the same as bk_spmv_tma.cuh: one producer warp, cp.async.bulk + mbarrier ring, 8 consumer warps.
#pragma once

#include "bk_internal.cuh"
#include "bk_spmv_tma.cuh"

struct __align__(16) bk_pair_entry {  // dictionary entry (val holds a T in its low bytes)
  unsigned long long val;
  int off;
  int pad;
};
#define BK_PAIR_ZERO 31      // dictionary slot that always holds {0, 0}
#define BK_PAIR_DICT_BYTES (32 * 16)

#define BK_PAIR_HDR_BYTES 128

struct bk_pair_plan {
  const int* bptr;             // [nblk + 1] byte offset of every 256-row block's span in `codes` (multiples of 128)
  const unsigned char* codes;  // per block: 128-byte header + SELL-32-4 code stream of its 8 chunks
  const bk_pair_entry* dict;   // [nblk][32]
  int cap;                     // bytes of codes per stage (multiple of 128)
  int stages;
};

template <typename T, int MODE, int DOTS, int MINB, typename Epi>
__global__ void __launch_bounds__(BK_TMA_THREADS, MINB)
bk_spmv_pair_kernel(const bk_spmv_args a, const bk_pair_plan plan, const bk_scratch sc, Epi epi) {
  if (bk_spmv_skip(a)) return;
  extern __shared__ __align__(128) unsigned char bk_smem_pair[];
  __shared__ __align__(8) uint64_t full_bar[BK_TMA_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[BK_TMA_MAX_STAGES];
  constexpr int R = bk_ndots<DOTS>::value;
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int nstage = plan.stages;
  const int cap = plan.cap;
  const uint32_t stage_bytes = (uint32_t)cap + BK_PAIR_DICT_BYTES;
  const int* __restrict__ bptr = plan.bptr;
  const int n32 = (int)a.n;
  const int nblk = (n32 + BK_TMA_RPB - 1) / BK_TMA_RPB;
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

  // this CTA's blocks: blk0, blk0 + bstep, ... (my_iters of them); a reversed sweep walks down from the last block
  const int gstep = (int)gridDim.x;
  const int my_iters = (nblk > (int)blockIdx.x) ? (nblk - 1 - (int)blockIdx.x) / gstep + 1 : 0;
  const int blk0 = reverse ? (nblk - 1 - (int)blockIdx.x) : (int)blockIdx.x;
  const int bstep = reverse ? -gstep : gstep;

  if (wid == BK_WARPS) {
    // ------------------------------ producer warp ------------------------------------------------
    const uint64_t pol = bk_policy_evict_first();
    int s_cur = 0, e_cur = 0;  // lane j holds the byte span of iteration (batch*32 + j)
    for (int it0 = 0; it0 < my_iters; it0 += 32) {
      {
        const int it = it0 + lane;
        if (it < my_iters) {
          const int blk = blk0 + it * bstep;
          s_cur = __ldg(bptr + blk);
          e_cur = __ldg(bptr + blk + 1);
        }
      }
      const int lim = (my_iters - it0 < 32) ? (my_iters - it0) : 32;
      for (int j = 0; j < lim; ++j) {
        const int it = it0 + j;
        const int s = __shfl_sync(0xffffffffu, s_cur, j);
        const int e = __shfl_sync(0xffffffffu, e_cur, j);
        if (lane == 0) {
          const int stage = it % nstage;
          if (it >= nstage) bk_mbar_wait(&empty_bar[stage], (uint32_t)(((it / nstage) - 1) & 1));
          unsigned char* sc_ = bk_smem_pair + (size_t)stage * stage_bytes;
          bk_mbar_expect_tx(&full_bar[stage], (uint32_t)(e - s) + BK_PAIR_DICT_BYTES);
          bk_bulk_g2s(sc_, plan.codes + s, (uint32_t)(e - s), &full_bar[stage], pol);
          bk_bulk_g2s(sc_ + cap, plan.dict + (size_t)(blk0 + it * bstep) * 32, BK_PAIR_DICT_BYTES, &full_bar[stage],
                      pol);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ consumer warps: warp <-> 32-row chunk, lane <-> row -------------
    const T* __restrict__ x = static_cast<const T*>(a.x);
    T* __restrict__ y = static_cast<T*>(a.y);
    const uint32_t smem0 = bk_smem_u32(bk_smem_pair);
    const uint32_t full0 = bk_smem_u32(&full_bar[0]);
    const uint32_t empty0 = bk_smem_u32(&empty_bar[0]);
    int row = blk0 * BK_TMA_RPB + wid * 32 + lane;
    const int rstep = bstep * BK_TMA_RPB;
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < my_iters; ++it) {
      {  // wait for the tile
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "PAIR_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
            "@P1 bra PAIR_DONE;\n"
            "bra PAIR_WAIT;\n"
            "PAIR_DONE:\n"
            "}\n" ::"r"(full0 + stage * 8),
            "r"(phase)
            : "memory");
      }
      const uint32_t sbase = smem0 + (uint32_t)stage * stage_bytes;
      uint32_t hdr;  // (end << 16 | start) of my chunk, 128-byte units from the start of the tile
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hdr) : "r"(sbase + wid * 4));
      const int c_words = (int)(hdr >> 16) - (int)(hdr & 0xffffu);
      const uint32_t wp = sbase + ((hdr & 0xffffu) << 7) + lane * 4;  // word j of this lane at wp + j * 128
      const uint32_t sp = sbase + cap;
      const int xrow = min(row, n32 - 1);  // rows past the end exist only as padding (value 0)
      T sum = T(0);
      for (int j = 0; j < c_words; j += 2) {
        // 8 entries: two conflict-free code words, eight 16-byte dictionary loads, eight gathers, the FMA chain
        uint32_t w0, w1 = 0x1f1f1f1fu;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(wp + j * 128));
        if (j + 1 < c_words) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(wp + j * 128 + 128));
        unsigned int lo[8], hi[8];
        int po[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t code = __byte_perm(u < 4 ? w0 : w1, 0u, 0x4440u + (unsigned)(u & 3));
          unsigned int pad;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(lo[u]), "=r"(hi[u]), "=r"(po[u]), "=r"(pad)
                       : "r"(sp + code * 16u));
        }
        T xv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) xv[u] = __ldg(x + (xrow + po[u]));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          T v;
          if constexpr (sizeof(T) == 8) v = __hiloint2double((int)hi[u], (int)lo[u]);
          else v = __uint_as_float(lo[u]);
          sum = fma(v, xv[u], sum);
        }
      }
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + stage * 8) : "memory");
      if (row < n32) {
        T out = sum;
        if constexpr (MODE == 1) out = bk_sub(__ldg(static_cast<const T*>(a.b) + row), sum);
        y[row] = out;
        if constexpr ((DOTS & 1) != 0)
          acc[0] += static_cast<double>(__ldg(static_cast<const T*>(a.w) + row)) * static_cast<double>(out);
        if constexpr ((DOTS & 2) != 0) acc[DOTS & 1] += static_cast<double>(out) * static_cast<double>(out);
      }
      row += rstep;
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
