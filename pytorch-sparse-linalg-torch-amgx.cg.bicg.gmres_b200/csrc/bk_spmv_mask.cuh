// bk_spmv_mask.cuh — kernel 6: SpMV for matrices whose 32-row chunks each hold at most 8 distinct
// (column - row, value) PAIRS (constant-coefficient stencils: Poisson, upwind convection-diffusion, the LDC pressure
// matrix ...), the successor of the pair-coded kernel 5 for that class.
//
// ncu on kernel 5 (profiles/r01_ncu_full_final_cg.txt) showed it bound by the L1/LSU data pipe (72 %) and by issue
// slots (59 %), not by HBM (41 %): every entry cost a 16-byte dictionary fetch from shared memory PER LANE (4 LSU
// wavefronts — the return path moves 128 B per cycle whether or not the 32 lanes read the same address), a gather
// (2-3 wavefronts) and ~18 instructions.  Kernel 6 removes the per-entry matrix traffic from the SM altogether:
//   * registration (bk_csr_plan_mask, bk_core.cu) sorts every chunk's distinct pairs by (offset, value) into a PATTERN
//     of <= 8 entries, de-duplicates the patterns of the whole matrix in a small table (a 7-point stencil has a dozen),
//     and stores per row ONE BYTE: the bitmask of pattern entries the row actually has; per chunk the pattern id.
//     Matrix stream: 1.125 B per ROW instead of 12 B per ENTRY (P3D-256: 19 MB instead of 1.40 GB per SpMV);
//   * a warp owns a 32-row chunk, lane <-> row; the pattern {value, offset} x 8 lives in REGISTERS and is reloaded only
//     when the chunk's pattern id differs from the previous chunk's (warp-uniform branch, a few times per 256 blocks);
//   * per entry the warp issues one predicated, fully coalesced gather (lanes of a stencil diagonal read 32 consecutive
//     elements) and one FMA — ~4 instructions; absent entries load nothing and contribute fma(v, 0, sum) = sum;
//   * row blocks are dealt to CTAs in groups of consecutive blocks, so neighbouring grid lines are gathered from L1
//     while the chip as a whole still sweeps one window of the vectors (L2 reuse between consecutive kernels);
//   * the row sum is the same FMA chain in CSR order as kernels 2 / 3 / 5 => bit-identical results (tested).
// Requirements (else registration falls through to kernel 5 / 3 / 2): columns ascending within every row, <= 8 entries
// per row, <= 8 distinct pairs per 32-row chunk, <= BK_MASK_HT/2 distinct patterns in the matrix.
//
// Multi-GPU (bk_dist.cuh): a pattern entry can be flagged GHOST — its gather then reads the halo vector instead of x,
// which folds the boundary rows of a row-partitioned matrix into this kernel (no separate ghost-row kernel).
#pragma once

#include <type_traits>

#include "bk_internal.cuh"
#include "bk_p2p.cuh"
#include "bk_spmv_tma.cuh"

#define BK_MASK_L 8          // pattern entries per chunk (mask bits per row)
#define BK_MASK_HT 4096      // slots of the pattern table (open addressing; at most half may fill)
#define BK_MASK_GHOST 1      // bk_pair_entry.pad flag: gather from the ghost vector

struct bk_mask_plan {
  const unsigned char* masks;  // [nchunks * 32] presence bits of every row over its chunk's pattern
  const int* pids;             // [nchunks] pattern table slot of every chunk (bit 30: chunk has ghost entries)
  const bk_pair_entry* ptab;   // [BK_MASK_HT][BK_MASK_L]
  const void* xg;              // ghost vector (multi-GPU) or nullptr
  const int* deferred;         // chunks with ghost entries, processed after the halo has arrived (multi-GPU)
  int n_deferred;
  int group;                   // log2 of the consecutive 256-row blocks dealt to a CTA at a time (0..5)
  // halo arrival (peer-memory path): flags[peer] == want  (nullptr: no wait, e.g. NCCL path orders by stream)
  const unsigned long long* flags;
  const int* flag_peers;
  int n_flag_peers;
  unsigned int* halo_seq;      // counters[2] of the peer context: halos consumed so far (want = *halo_seq + 1)
  unsigned int* err_flag;      // counters[4]
  int prefetch;                // 1: bulk-prefetch the x range of this CTA's NEXT group into L2 (x 16-byte aligned)
};

#define BK_MASK_PID_GHOST (1 << 30)

#define BK_MASK_FULL_SHIFT 8  // bk_pair_entry.pad of a pattern's entry 0: bits 8..16 = mask of a row that has every entry

template <typename T>
struct bk_mask_pat {
  T val[BK_MASK_L];
  int off[BK_MASK_L];
  unsigned int ghost;  // bit e: entry e gathers from the ghost vector
  unsigned int full;   // mask of a row that has every entry of the pattern (0xffffffff for an empty pattern)
};

template <typename T>
__device__ __forceinline__ void bk_mask_load_pattern(const bk_pair_entry* __restrict__ ptab, int slot, bk_mask_pat<T>& p) {
  const uint4* src = reinterpret_cast<const uint4*>(ptab + (size_t)slot * BK_MASK_L);
  p.ghost = 0u;
#pragma unroll
  for (int e = 0; e < BK_MASK_L; ++e) {
    const uint4 q = __ldg(src + e);  // same address in every lane: one sector
    if constexpr (sizeof(T) == 8) p.val[e] = __hiloint2double((int)q.y, (int)q.x);
    else p.val[e] = __uint_as_float(q.x);
    p.off[e] = (int)q.z;
    p.ghost |= (q.w & BK_MASK_GHOST) ? (1u << e) : 0u;
    if (e == 0) p.full = (q.w >> BK_MASK_FULL_SHIFT) & 0x1ffu;
  }
  if (p.full == 0u) p.full = 0xffffffffu;
}

// Gathers: the lane's row pointer lives in a register pair, so an address is ONE 64-bit multiply-add of the pattern
// offset (ncu on the first version, whose addresses the compiler derived from `x[row + off]`: 145 instructions per
// 32-row chunk, issue-bound at 116 us).  bk_ld_masked: bit ? base[off] : 0 (predicated load); bk_ld_plain: base[off].
__device__ __forceinline__ double bk_ld_masked(const double* base, int off, unsigned int bit) {
  double v;
  asm("{\n\t.reg .pred p;\n\t.reg .u64 a;\n\tsetp.ne.u32 p, %2, 0;\n\tmad.wide.s32 a, %3, 8, %1;\n\t"
      "mov.f64 %0, 0d0000000000000000;\n\t@p ld.global.nc.f64 %0, [a];\n\t}"
      : "=d"(v)
      : "l"(base), "r"(bit), "r"(off));
  return v;
}
__device__ __forceinline__ float bk_ld_masked(const float* base, int off, unsigned int bit) {
  float v;
  asm("{\n\t.reg .pred p;\n\t.reg .u64 a;\n\tsetp.ne.u32 p, %2, 0;\n\tmad.wide.s32 a, %3, 4, %1;\n\t"
      "mov.f32 %0, 0f00000000;\n\t@p ld.global.nc.f32 %0, [a];\n\t}"
      : "=f"(v)
      : "l"(base), "r"(bit), "r"(off));
  return v;
}
__device__ __forceinline__ double bk_ld_masked_cg(const double* base, int off, unsigned int bit) {
  double v;
  asm volatile("{\n\t.reg .pred p;\n\t.reg .u64 a;\n\tsetp.ne.u32 p, %2, 0;\n\tmad.wide.s32 a, %3, 8, %1;\n\t"
               "mov.f64 %0, 0d0000000000000000;\n\t@p ld.global.cg.f64 %0, [a];\n\t}"
               : "=d"(v)
               : "l"(base), "r"(bit), "r"(off)
               : "memory");
  return v;
}
__device__ __forceinline__ float bk_ld_masked_cg(const float* base, int off, unsigned int bit) {
  float v;
  asm volatile("{\n\t.reg .pred p;\n\t.reg .u64 a;\n\tsetp.ne.u32 p, %2, 0;\n\tmad.wide.s32 a, %3, 4, %1;\n\t"
               "mov.f32 %0, 0f00000000;\n\t@p ld.global.cg.f32 %0, [a];\n\t}"
               : "=f"(v)
               : "l"(base), "r"(bit), "r"(off)
               : "memory");
  return v;
}
__device__ __forceinline__ double bk_ld_plain(const double* base, int off) {
  double v;
  asm("{\n\t.reg .u64 a;\n\tmad.wide.s32 a, %2, 8, %1;\n\tld.global.nc.f64 %0, [a];\n\t}" : "=d"(v) : "l"(base), "r"(off));
  return v;
}
__device__ __forceinline__ float bk_ld_plain(const float* base, int off) {
  float v;
  asm("{\n\t.reg .u64 a;\n\tmad.wide.s32 a, %2, 4, %1;\n\tld.global.nc.f32 %0, [a];\n\t}" : "=f"(v) : "l"(base), "r"(off));
  return v;
}

// one 32-row chunk: y[row] = sum_e [mask bit e] val_e * x[row + off_e]   (+ fused residual / dots).
// FAST: every row of the chunk has every entry of the pattern (warp-uniform test by the caller; ~3/4 of the chunks of
// a stencil matrix): no predicates, no zero-fill, no bounds test (a row with a non-empty mask exists).
template <typename T, int MODE, int DOTS, bool GHOST, bool FAST>
__device__ __forceinline__ void bk_mask_chunk(const bk_spmv_args& a, const bk_mask_pat<T>& p, const T* x, const T* xg,
                                              const int row, const unsigned int m, const int n32, double* acc,
                                              const unsigned long long pol = 0ull) {
  const T* xr = x + row;
  const T* xgr = GHOST ? xg + row : nullptr;
  T xv[BK_MASK_L];
#pragma unroll
  for (int e = 0; e < BK_MASK_L; ++e) {
    if (GHOST && (p.ghost & (1u << e))) xv[e] = bk_ld_masked_cg(xgr, p.off[e], m & (1u << e));
    else if (FAST) xv[e] = bk_ld_plain(xr, p.off[e]);  // entries past the pattern's length: value 0, offset 0
    else xv[e] = bk_ld_masked(xr, p.off[e], m & (1u << e));
  }
  T sum = T(0);
#pragma unroll
  for (int e = 0; e < BK_MASK_L; ++e) sum = fma(p.val[e], xv[e], sum);
  if (FAST || row < n32) {
    T out = sum;
    if constexpr (MODE == 1) out = bk_sub(__ldg(static_cast<const T*>(a.b) + row), sum);
    if (pol != 0ull) bk_st_hint1(static_cast<T*>(a.y) + row, out, pol);
    else static_cast<T*>(a.y)[row] = out;
    if constexpr ((DOTS & 1) != 0)
      acc[0] += static_cast<double>(__ldg(static_cast<const T*>(a.w) + row)) * static_cast<double>(out);
    if constexpr ((DOTS & 2) != 0) acc[DOTS & 1] += static_cast<double>(out) * static_cast<double>(out);
  }
}

template <typename T, int MODE, int DOTS, bool GHOST, int MINB, typename Epi>
__global__ void __launch_bounds__(BK_BLOCK, MINB)
bk_spmv_mask_kernel(const bk_spmv_args a, const bk_mask_plan plan, const bk_scratch sc, Epi epi) {
  if (bk_spmv_skip(a)) return;
  constexpr int R = bk_ndots<DOTS>::value;
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int n32 = (int)a.n;
  const int nblk = (n32 + 255) >> 8;
  const int gshift = plan.group;  // log2 of the consecutive 256-row blocks a CTA takes per visit (0..5)
  const int gmask = (1 << gshift) - 1;
  const int ngroups = (nblk + gmask) >> gshift;
  int reverse = a.reverse;
  if (a.use_parity) reverse ^= (a.st->parity & 1);
  const T* x = static_cast<const T*>(a.x);
  const T* xg = static_cast<const T*>(plan.xg);
  const unsigned char* __restrict__ masks = plan.masks;
  const int* __restrict__ pids = plan.pids;

  double acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.0;
  const unsigned long long pol = (a.l2_hints & 16) ? bk_policy_evict_last() : 0ull;

  bk_mask_pat<T> pat;
  int cur = -1;
  // Groups of 2^gshift consecutive blocks are dealt round-robin to the CTAs (the chip sweeps one window of the vectors;
  // inside a group neighbouring grid lines are gathered from L1).  A warp takes chunk `wid` of every block of its
  // groups: chunk t = 0 .. T-1 of its sequence.  `masks` / `pids` are padded to 32 whole blocks beyond the matrix (zero
  // masks gather nothing), so the tail of the last group needs no test.
  const int my_groups = (ngroups > (int)blockIdx.x) ? (ngroups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int T_ = my_groups << gshift;
  // this lane's row in chunk t of the warp's sequence; inside a group consecutive chunks are 256 rows apart, so the
  // hot loop advances with one add and re-derives the row only when it enters a new group
  auto row_at = [&](int t) -> int {
    const int g = (int)blockIdx.x + (t >> gshift) * (int)gridDim.x;
    const int gg = reverse ? (ngroups - 1 - g) : g;
    return ((((gg << gshift) + (t & gmask)) << 3) + wid) * 32 + lane;
  };
  int t = 0;
  int row_next = 0, pid_next = 0;
  unsigned int m_next = 0u;
  if (T_ > 0) {
    row_next = row_at(0);
    m_next = __ldg(masks + row_next);
    pid_next = __ldg(pids + (row_next >> 5));
  }
  // L2 prefetch, one visit ahead: ncu showed the kernel waiting on DRAM latency (one first-touch line of x per chunk,
  // ~1700 cycles per chunk with 8 warps per scheduler; DRAM 32 %, issue 44 %).  When a CTA starts a group, one thread
  // asks the L2 for the x range (and the masks) of the group the CTA will visit NEXT with a single bulk-prefetch
  // instruction each (SASS UBLKPF); every gather of that visit — centre lines and the neighbours' lines, which are
  // other CTAs' centre lines — then hits L2 (measured: 99 -> 84 us).
  auto prefetch_group = [&](int visit) {
    const int g = (int)blockIdx.x + visit * (int)gridDim.x;
    if (g >= ngroups) return;
    const int gg = reverse ? (ngroups - 1 - g) : g;
    const long long r0 = (long long)(gg << gshift) * 256;
    long long r1 = r0 + ((long long)256 << gshift);
    constexpr long long EA = 16 / sizeof(T);
    const long long nal = (long long)n32 & ~(EA - 1);
    if (r1 > nal) r1 = nal;
    if (r1 > r0) bk_bulk_prefetch_l2(x + r0, (uint32_t)((r1 - r0) * sizeof(T)));
    bk_bulk_prefetch_l2(masks + r0, (uint32_t)(256 << gshift));  // (padded to whole groups)
  };
  const bool pf = plan.prefetch != 0 && threadIdx.x == 0;
  if (pf) prefetch_group(1);
  while (t < T_) {
    // a run of chunks with the same pattern: the pattern is (re)loaded here, outside the hot loop
    const int slot = pid_next & (BK_MASK_PID_GHOST - 1);
    if (slot != cur) {
      bk_mask_load_pattern<T>(plan.ptab, slot, pat);
      cur = slot;
    }
    do {
      const int row = row_next;
      const unsigned int m = m_next;
      const bool ghost_chunk = GHOST && (pid_next & BK_MASK_PID_GHOST);
      ++t;
      if ((t & gmask) != 0) {
        row_next += 256;
      } else {  // next group (1 step in 2^gshift)
        if (pf) prefetch_group((t >> gshift) + 1);
        row_next = (t < T_) ? row_at(t) : row_next;
      }
      if (t < T_) {  // the next chunk's mask / pattern id are in flight while this one is computed
        m_next = __ldg(masks + row_next);
        pid_next = __ldg(pids + (row_next >> 5));
      }
      if (!ghost_chunk) {  // (chunks with ghost entries: second phase)
        if (__all_sync(0xffffffffu, m == pat.full))
          bk_mask_chunk<T, MODE, DOTS, false, true>(a, pat, x, xg, row, m, n32, acc, pol);
        else
          bk_mask_chunk<T, MODE, DOTS, false, false>(a, pat, x, xg, row, m, n32, acc, pol);
      }
    } while (t < T_ && (pid_next & (BK_MASK_PID_GHOST - 1)) == slot);
  }
  if constexpr (GHOST) {
    // ---- second phase: chunks with ghost entries, after the neighbours' halos have landed ---------------------
    if (plan.n_deferred > 0 || plan.flags != nullptr) {
      bool ok = true;
      if (plan.flags != nullptr) {
        __shared__ int s_fail;
        if (threadIdx.x == 0) s_fail = 0;
        __syncthreads();
        const unsigned int want = *plan.halo_seq + 1u;
        if ((int)threadIdx.x < plan.n_flag_peers) {
          const unsigned long long* flag = plan.flags + plan.flag_peers[threadIdx.x];
          const long long t0 = clock64();
          while ((unsigned int)bk_ld_acquire_sys_u64(flag) != want) {
            if (clock64() - t0 > BK_P2P_TIMEOUT_CYCLES) {
              s_fail = 1;
              break;
            }
          }
        }
        __syncthreads();
        if (s_fail) {
          if (threadIdx.x == 0) *plan.err_flag = 1u;
          ok = false;
        }
      }
      if (ok) {
        const int warps = (int)gridDim.x * BK_WARPS;
        for (int i = (int)blockIdx.x * BK_WARPS + wid; i < plan.n_deferred; i += warps) {
          const int c0 = __ldg(plan.deferred + i);
          const unsigned int m = __ldg(masks + (size_t)c0 * 32 + lane);
          const int slot = __ldg(pids + c0) & (BK_MASK_PID_GHOST - 1);
          if (slot != cur) {
            bk_mask_load_pattern<T>(plan.ptab, slot, pat);
            cur = slot;
          }
          bk_mask_chunk<T, MODE, DOTS, true, false>(a, pat, x, xg, c0 * 32 + lane, m, n32, acc);
        }
      }
    }
  }
  if constexpr (DOTS != 0) {
    bk_grid_reduce<R, Epi, BK_WARPS>(acc, sc, epi);
  }
}

// =====================================================================================================================
// Kernel 6W — the same row-bitmask SpMV with its gathers served from SHARED MEMORY: a producer warp stages, per group of
// 2^gshift consecutive 256-row blocks, the windows of x the group's rows can touch with TMA bulk copies (cp.async.bulk +
// mbarrier ring, as kernels 2 / 3 / 5 stage the matrix stream):
//   near window   x[row0 - W, row0 + rows + W)          every offset with |off| <= W (W from the pattern table)
//   far windows   x[row0 + F_k, row0 + F_k + rows)      up to two far offsets F_k (the +-n^2 planes of a 3-D stencil)
// plus the group's masks and pattern ids.  Why: ncu on kernel 6 (profiles/r02_ncu_k6_*.txt) shows it bound by the L1TEX
// pipe, not by HBM — a warp-wide 8-byte gather touches 2-3 cache lines and replays cost ~2 cycles per line, ~40 LSU
// cycles per 32-row chunk (85 us at 256^3) against an HBM floor of 44 us; bulk copies bypass that pipe, a conflict-free
// LDS.64 costs 2 cycles flat, and the ring keeps NSTAGE groups of loads in flight per CTA regardless of occupancy.
// A gather becomes `lds [lane_base + disp_e]` with one per-pattern displacement per entry.  Patterns with an offset that
// no window covers (and chunks with ghost entries) take the LDG path of kernel 6 inside the same kernel.
// =====================================================================================================================
#define BK_MW_MAXFAR 2

struct bk_maskw_plan {
  int win;                  // W: elements on each side of the near window (multiple of 8)
  int nfar;
  int far_off[BK_MW_MAXFAR];  // F_k (multiples of 16 / sizeof(T))
  int stages;
  uint32_t stage_bytes;
};

__device__ __forceinline__ void bk_bulk_g2s_plain(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   bk_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(bk_smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ double bk_lds_plain(uint32_t addr, double) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float bk_lds_plain(uint32_t addr, float) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double bk_lds_masked(uint32_t addr, unsigned int bit, double) {
  double v;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@p ld.shared.f64 %0, [%1];\n\t}"
               : "=d"(v)
               : "r"(addr), "r"(bit));
  return v;
}
__device__ __forceinline__ float bk_lds_masked(uint32_t addr, unsigned int bit, float) {
  float v;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t@p ld.shared.f32 %0, [%1];\n\t}"
               : "=f"(v)
               : "r"(addr), "r"(bit));
  return v;
}

// one chunk from the staged windows: xv[e] = smem[lane_base + disp[e]]
template <typename T, int MODE, int DOTS, bool FAST, int CNT>
__device__ __forceinline__ void bk_maskw_chunk(const bk_spmv_args& a, const bk_mask_pat<T>& p, const int (&disp)[BK_MASK_L],
                                               const uint32_t lane_base, const int row, const unsigned int m,
                                               const int n32, double* acc) {
  T xv[BK_MASK_L];
#pragma unroll
  for (int e = 0; e < BK_MASK_L; ++e) {
    if (e < CNT) {
      if (FAST) xv[e] = bk_lds_plain(lane_base + (uint32_t)disp[e], T(0));
      else xv[e] = bk_lds_masked(lane_base + (uint32_t)disp[e], m & (1u << e), T(0));
    } else {
      xv[e] = T(0);
    }
  }
  T sum = T(0);
#pragma unroll
  for (int e = 0; e < BK_MASK_L; ++e)
    if (e < CNT) sum = fma(p.val[e], xv[e], sum);
  if (FAST || row < n32) {
    T out = sum;
    if constexpr (MODE == 1) out = bk_sub(__ldg(static_cast<const T*>(a.b) + row), sum);
    static_cast<T*>(a.y)[row] = out;
    if constexpr ((DOTS & 1) != 0)
      acc[0] += static_cast<double>(__ldg(static_cast<const T*>(a.w) + row)) * static_cast<double>(out);
    if constexpr ((DOTS & 2) != 0) acc[DOTS & 1] += static_cast<double>(out) * static_cast<double>(out);
  }
}

template <typename T, int MODE, int DOTS, bool GHOST, int MINB, typename Epi>
__global__ void __launch_bounds__(BK_TMA_THREADS, MINB)
bk_spmv_maskw_kernel(const bk_spmv_args a, const bk_mask_plan plan, const bk_maskw_plan wp, const bk_scratch sc,
                     Epi epi) {
  if (bk_spmv_skip(a)) return;
  extern __shared__ __align__(128) unsigned char bk_smem_mw[];
  __shared__ __align__(8) uint64_t full_bar[BK_TMA_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[BK_TMA_MAX_STAGES];
  constexpr int R = bk_ndots<DOTS>::value;
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int n32 = (int)a.n;
  const int nblk = (n32 + 255) >> 8;
  const int gshift = plan.group;
  const int G = 1 << gshift;
  const int rows_g = G << 8;
  const int ngroups = (nblk + G - 1) >> gshift;
  const int nstage = wp.stages;
  const int W = wp.win;
  int reverse = a.reverse;
  if (a.use_parity) reverse ^= (a.st->parity & 1);
  const T* x = static_cast<const T*>(a.x);
  const T* xg = static_cast<const T*>(plan.xg);
  // stage layout: [near window][far windows][masks][pattern ids]
  const uint32_t near_bytes = (uint32_t)(rows_g + 2 * W) * (uint32_t)sizeof(T);
  const uint32_t far_bytes = (uint32_t)rows_g * (uint32_t)sizeof(T);
  const uint32_t mask_off = near_bytes + (uint32_t)wp.nfar * far_bytes;
  const uint32_t pid_off = mask_off + (uint32_t)rows_g;

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

  const int my_groups = (ngroups > (int)blockIdx.x) ? (ngroups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto row0_of = [&](int u) -> int {
    const int g = (int)blockIdx.x + u * (int)gridDim.x;
    const int gg = reverse ? (ngroups - 1 - g) : g;
    return (gg << gshift) << 8;
  };

  bk_mask_pat<T> pat;
  int cur = -1;
  if (wid == BK_WARPS) {
    // ------------------------------ producer warp ------------------------------------------------
    if (lane == 0) {
      for (int u = 0; u < my_groups; ++u) {
        const int stage = u % nstage;
        if (u >= nstage) bk_mbar_wait(&empty_bar[stage], (uint32_t)(((u / nstage) - 1) & 1));
        unsigned char* sb = bk_smem_mw + (size_t)stage * wp.stage_bytes;
        const int row0 = row0_of(u);
        uint32_t tx = (uint32_t)rows_g + (uint32_t)(G * 32);
        // windows, clamped to [0, n) (n is a multiple of the 16-byte pack: checked by the launcher)
        int lo[1 + BK_MW_MAXFAR], hi[1 + BK_MW_MAXFAR], base[1 + BK_MW_MAXFAR];
        uint32_t soff[1 + BK_MW_MAXFAR];
        base[0] = row0 - W;
        lo[0] = max(base[0], 0);
        hi[0] = min(row0 + rows_g + W, n32);
        soff[0] = 0u;
        for (int k = 0; k < wp.nfar; ++k) {
          base[1 + k] = row0 + wp.far_off[k];
          lo[1 + k] = max(base[1 + k], 0);
          hi[1 + k] = min(base[1 + k] + rows_g, n32);
          soff[1 + k] = near_bytes + (uint32_t)k * far_bytes;
        }
        for (int k = 0; k <= wp.nfar; ++k)
          if (hi[k] > lo[k]) tx += (uint32_t)(hi[k] - lo[k]) * (uint32_t)sizeof(T);
        bk_mbar_expect_tx(&full_bar[stage], tx);
        for (int k = 0; k <= wp.nfar; ++k)
          if (hi[k] > lo[k])
            bk_bulk_g2s_plain(sb + soff[k] + (size_t)(lo[k] - base[k]) * sizeof(T), x + lo[k],
                              (uint32_t)(hi[k] - lo[k]) * (uint32_t)sizeof(T), &full_bar[stage]);
        bk_bulk_g2s_plain(sb + mask_off, plan.masks + row0, (uint32_t)rows_g, &full_bar[stage]);
        bk_bulk_g2s_plain(sb + pid_off, plan.pids + (row0 >> 5), (uint32_t)(G * 32), &full_bar[stage]);
      }
    }
  } else {
    // ------------------------------ consumer warps: warp <-> chunk `wid` of every block of the group ------------
    int disp[BK_MASK_L];      // byte displacement of every pattern entry inside a stage, relative to the lane's slot
    bool windowed = false;    // every entry of the current pattern is covered by a staged window
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t smem0 = bk_smem_u32(bk_smem_mw);
    for (int u = 0; u < my_groups; ++u) {
      bk_mbar_wait(&full_bar[stage], phase);
      const uint32_t sb = smem0 + (uint32_t)stage * wp.stage_bytes;
      const int row0 = row0_of(u);
      for (int j = 0; j < G; ++j) {
        const int lrow = (j << 8) + (wid << 5) + lane;
        const int row = row0 + lrow;
        unsigned int m;
        int pid;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(m) : "r"(sb + mask_off + (uint32_t)lrow));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(pid) : "r"(sb + pid_off + (uint32_t)(((j << 3) + wid) << 2)));
        if (GHOST && (pid & BK_MASK_PID_GHOST)) continue;  // second phase
        const int slot = pid & (BK_MASK_PID_GHOST - 1);
        if (slot != cur) {
          bk_mask_load_pattern<T>(plan.ptab, slot, pat);
          cur = slot;
          windowed = true;
#pragma unroll
          for (int e = 0; e < BK_MASK_L; ++e) {
            const int off = pat.off[e];
            int d = -1;
            if (!(pat.ghost & (1u << e))) {
              if (off >= -W && off <= W) d = (W + off) * (int)sizeof(T);
              for (int k = 0; k < wp.nfar; ++k)
                if (off == wp.far_off[k]) d = (int)(near_bytes + (uint32_t)k * far_bytes);
            }
            if (d < 0 && (pat.full & (1u << e)) && pat.full != 0xffffffffu) windowed = false;  // a real entry nobody stages
            disp[e] = d < 0 ? 0 : d;
          }
        }
        const bool fast = __all_sync(0xffffffffu, m == pat.full);
        if (windowed) {
          const uint32_t lane_base = sb + (uint32_t)lrow * (uint32_t)sizeof(T);
          const int cnt = __popc(pat.full);  // pattern length (full is (1 << count) - 1)
          if (fast) {
            if (cnt <= 5) bk_maskw_chunk<T, MODE, DOTS, true, 5>(a, pat, disp, lane_base, row, m, n32, acc);
            else if (cnt <= 7) bk_maskw_chunk<T, MODE, DOTS, true, 7>(a, pat, disp, lane_base, row, m, n32, acc);
            else bk_maskw_chunk<T, MODE, DOTS, true, 8>(a, pat, disp, lane_base, row, m, n32, acc);
          } else {
            bk_maskw_chunk<T, MODE, DOTS, false, 8>(a, pat, disp, lane_base, row, m, n32, acc);
          }
        } else if (fast) {
          bk_mask_chunk<T, MODE, DOTS, false, true>(a, pat, x, xg, row, m, n32, acc);
        } else {
          bk_mask_chunk<T, MODE, DOTS, false, false>(a, pat, x, xg, row, m, n32, acc);
        }
      }
      __syncwarp();
      if (lane == 0) bk_mbar_arrive(&empty_bar[stage]);
      if (++stage == nstage) {
        stage = 0;
        phase ^= 1u;
      }
    }
  }
  if constexpr (GHOST) {
    // ---- second phase: chunks with ghost entries, after the neighbours' halos have landed (LDG path) -----------
    if (plan.n_deferred > 0 || plan.flags != nullptr) {
      __shared__ int s_fail;
      bool ok = true;
      if (plan.flags != nullptr) {
        if (threadIdx.x == 0) s_fail = 0;
        __syncthreads();
        const unsigned int want = *plan.halo_seq + 1u;
        if ((int)threadIdx.x < plan.n_flag_peers) {
          const unsigned long long* flag = plan.flags + plan.flag_peers[threadIdx.x];
          const long long t0 = clock64();
          while ((unsigned int)bk_ld_acquire_sys_u64(flag) != want) {
            if (clock64() - t0 > BK_P2P_TIMEOUT_CYCLES) {
              s_fail = 1;
              break;
            }
          }
        }
        __syncthreads();
        if (s_fail) {
          if (threadIdx.x == 0) *plan.err_flag = 1u;
          ok = false;
        }
      }
      if (ok && wid < BK_WARPS) {
        const int warps = (int)gridDim.x * BK_WARPS;
        for (int i = (int)blockIdx.x * BK_WARPS + wid; i < plan.n_deferred; i += warps) {
          const int c0 = __ldg(plan.deferred + i);
          const unsigned int m = __ldg(plan.masks + (size_t)c0 * 32 + lane);
          const int slot = __ldg(plan.pids + c0) & (BK_MASK_PID_GHOST - 1);
          if (slot != cur) {
            bk_mask_load_pattern<T>(plan.ptab, slot, pat);
            cur = slot;
          }
          bk_mask_chunk<T, MODE, DOTS, true, false>(a, pat, x, xg, c0 * 32 + lane, m, n32, acc);
        }
      }
    }
  }
  if constexpr (DOTS != 0) {
    bk_grid_reduce<R, Epi, BK_WARPS + 1>(acc, sc, epi);
  }
}

// =====================================================================================================================
// Kernel 7 — the stencil fast path of the row-bitmask SpMV, rebuilt around what the round-2 probes measured
// (tools/micro/gather_width.cu, tools/k6_occupancy.py, tools/k6_traffic_probe.py; summaries in profiles/):
//   * a 7-gather stencil kernel is bound neither by HBM nor by L2 traffic (the same time whether the gathers hit L1, L2
//     or all read the same line) but by the number of load instructions per row and the warps in flight: a bare 7-point
//     gather takes 131 / 78 / 62 us at 2 / 4 / 6 CTAs per SM with 64-bit loads, 82 / 56 / 52 us with 128-bit loads
//     (copy: 48 us); kernel 6 holds the pattern in 26 registers and issues ~85 instructions per 32-row chunk (84-105 us).
// Hence, for matrices whose patterns are all sub-patterns of ONE offset set (the union; every constant-coefficient
// stencil — boundary patterns only lack entries):
//   * a lane owns TWO consecutive rows and gathers with 128-bit loads; an entry with an odd offset takes the aligned
//     pair one element to the left, the second row's operand comes from the next lane by shuffle, lane 31 loads its own;
//   * the byte offsets of the union travel in the kernel's parameter block and are constant-bank operands of the 64-bit
//     address adds; the VALUES of the step's pattern (0 where the pattern lacks a union entry) are fetched from the
//     parameter block by pattern id — no pattern registers, no per-pattern code, no pattern switches;
//   * registration classifies every 64-row step once (bk_mask_usum_kernel): pattern id, "some row lacks an entry its
//     pattern has" (such steps read two mask bytes per lane — in union numbering — and zero the missing operands),
//     "two patterns / within reach of the matrix ends" (such steps run kernel 6's chunk code; none in the interior);
//   * the structure (number of union entries, which are odd) is a template parameter: 7-point 3-D and 5-point 2-D
//     stencils with even line lengths are instantiated; everything else stays on kernel 6.
// An operand multiplied by a zero value contributes fma(0, x, s) = s, as the padded entries of kernels 5 / 6 do: the row
// sums are the same FMA chain in CSR order => y is bit-identical to kernels 2 / 3 / 5 / 6 (tested).  fp64, single GPU.
// =====================================================================================================================
#define BK_MASK_CP 12        // patterns the parameter block holds
#define BK_MASK_US_DIRTY 0x100   // step summary: some row lacks an entry of its pattern -> masks are read
#define BK_MASK_US_MIXED 0x200   // step summary: two patterns in the 64 rows / matrix end within reach -> chunk code
#define BK_MASK_US_LO 0x400      // step summary: the ONLY missing entry is the -1 neighbour of the step's first row (a grid
#define BK_MASK_US_HI 0x800      // line starts there) / the +1 neighbour of its last row: the edge loads return 0, no masks

struct bk_mask_utab {
  double val[BK_MASK_CP][BK_MASK_L];  // value of union entry e in pattern p (0: the pattern has no such entry)
  long long offb[BK_MASK_L];          // byte offset of the 16-byte pair to load: 8*off (even off), 8*(off-1) (odd off)
};

struct bk_mask2_plan {
  const unsigned short* usum;     // [ngroups][8 warps][4 steps] step summaries (BK_MASK_US_*, low byte: pattern id)
  const unsigned char* umasks;    // [rows] presence bits in UNION numbering
};

__device__ __forceinline__ double2 bk_ldg_pair(const double* base, long long byte_off) {
  return __ldg(reinterpret_cast<const double2*>(reinterpret_cast<const char*>(base) + byte_off));
}

// One step: rows i, i + 1 of this lane (xr = x + i), every union entry.  The instantiated structures have the offsets
// -1, 0, +1 at the union positions K - 1, K, K + 1 (K = LEN / 2; checked at registration), so the two odd entries need no
// loads of their own: x[i - 1] is the previous lane's centre .y, x[i + 2] the next lane's centre .x (lanes 0 / 31 load
// theirs) — ncu on the first version showed the L1 wavefront pipe at 67 %, the top unit; this removes 8 of ~42 wavefronts.
template <int MODE, int DOTS, int LEN, unsigned int ODD, bool DIRTY>
__device__ __forceinline__ void bk_mask2_step(const bk_mask_utab& ct, const unsigned int summ, const double* xr,
                                              const unsigned char* __restrict__ pm, double* __restrict__ py,
                                              const double* __restrict__ pb, const double* __restrict__ pw,
                                              const int lane, double* acc) {
  const int pat = (int)(summ & 0xffu);
  constexpr int K = LEN / 2;
  static_assert(ODD == ((1u << (K - 1)) | (1u << (K + 1))), "kernel 7: the odd entries are the centre's two neighbours");
  double2 P[LEN];
#pragma unroll
  for (int e = 0; e < LEN; ++e)
    if (!((ODD >> e) & 1u)) P[e] = bk_ldg_pair(xr, ct.offb[e]);
  // the step's first row has no -1 neighbour when a grid line starts there (its last row no +1 when one ends): the usual
  // reason for an incomplete row, recorded in the summary so that such steps need no masks — the edge operand is 0
  double lo = 0.0, hi = 0.0;
  if (lane == 0 && !(summ & BK_MASK_US_LO)) lo = __ldg(xr - 1);
  if (lane == 31 && !(summ & BK_MASK_US_HI)) hi = __ldg(xr + 2);
  unsigned int m2 = 0xffffu;
  if (DIRTY) m2 = __ldg(reinterpret_cast<const unsigned short*>(pm));  // this lane's two mask bytes
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int e = 0; e < LEN; ++e) {
    double a0, a1;
    if (e == K - 1) {         // offset -1: rows i, i + 1 take x[i - 1], x[i]
      a0 = __shfl_up_sync(0xffffffffu, P[K].y, 1);
      if (lane == 0) a0 = lo;
      a1 = P[K].x;
    } else if (e == K + 1) {  // offset +1: x[i + 1], x[i + 2]
      a0 = P[K].y;
      a1 = __shfl_down_sync(0xffffffffu, P[K].x, 1);
      if (lane == 31) a1 = hi;
    } else {
      a0 = P[e].x;
      a1 = P[e].y;
    }
    if (DIRTY) {
      a0 = (m2 & (1u << e)) ? a0 : 0.0;
      a1 = (m2 & (0x100u << e)) ? a1 : 0.0;
    }
    const double v = ct.val[pat][e];
    s0 = fma(v, a0, s0);
    s1 = fma(v, a1, s1);
  }
  double o0 = s0, o1 = s1;
  if constexpr (MODE == 1) {
    const double2 bv = __ldg(reinterpret_cast<const double2*>(pb));
    o0 = bk_sub(bv.x, s0);
    o1 = bk_sub(bv.y, s1);
  }
  *reinterpret_cast<double2*>(py) = make_double2(o0, o1);
  if constexpr ((DOTS & 1) != 0) {
    const double2 wv = __ldg(reinterpret_cast<const double2*>(pw));
    acc[0] += wv.x * o0;
    acc[0] += wv.y * o1;
  }
  if constexpr ((DOTS & 2) != 0) {
    acc[DOTS & 1] += o0 * o0;
    acc[DOTS & 1] += o1 * o1;
  }
}

// A step with two patterns / within reach of the matrix ends: kernel 6's chunk code on its two 32-row chunks.
// Not inlined: its pattern registers must not weigh on the hot path.
// (returns its dot contributions by value: an accumulator whose address escapes would live in local memory)
template <int MODE, int DOTS, bool GHOST>
__device__ __noinline__ double2 bk_mask2_mixed_step(const bk_spmv_args a, const bk_mask_plan plan, const int row_first,
                                                    const int lane) {
  const double* x = static_cast<const double*>(a.x);
  const int n32 = (int)a.n;
  bk_mask_pat<double> pat;
  double t[2] = {0.0, 0.0};
#pragma unroll 1
  for (int u = 0; u < 2; ++u) {
    const int row = row_first + u * 32 + lane;
    const unsigned int m = __ldg(plan.masks + row);  // (masks / pids are padded by 32 blocks)
    const int pid = __ldg(plan.pids + (row >> 5));
    if (GHOST && (pid & BK_MASK_PID_GHOST)) continue;  // chunks with ghost entries: second phase, after the halo
    const int slot = pid & (BK_MASK_PID_GHOST - 1);
    bk_mask_load_pattern<double>(plan.ptab, slot, pat);
    bk_mask_chunk<double, MODE, DOTS, false, false>(a, pat, x, nullptr, row, m, n32, t);
  }
  return make_double2(t[0], t[1]);
}

// Multi-GPU (folded SpMV over [local | ghost], bk_dist.cuh): the chunks that gather ghost entries, after the neighbours'
// halos have landed — kernel 6's second phase (flag poll with acquire loads, then the compact list of those chunks).
template <int MODE, int DOTS>
__device__ __noinline__ double2 bk_mask2_ghost_phase(const bk_spmv_args a, const bk_mask_plan plan) {
  double t[2] = {0.0, 0.0};
  if (plan.n_deferred == 0 && plan.flags == nullptr) return make_double2(0.0, 0.0);
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  bool ok = true;
  if (plan.flags != nullptr) {
    __shared__ int s_fail2;
    if (threadIdx.x == 0) s_fail2 = 0;
    __syncthreads();
    const unsigned int want = *plan.halo_seq + 1u;
    if ((int)threadIdx.x < plan.n_flag_peers) {
      const unsigned long long* flag = plan.flags + plan.flag_peers[threadIdx.x];
      const long long t0 = clock64();
      while ((unsigned int)bk_ld_acquire_sys_u64(flag) != want) {
        if (clock64() - t0 > BK_P2P_TIMEOUT_CYCLES) {
          s_fail2 = 1;
          break;
        }
      }
    }
    __syncthreads();
    if (s_fail2) {
      if (threadIdx.x == 0) *plan.err_flag = 1u;
      ok = false;
    }
  }
  if (ok) {
    const double* x = static_cast<const double*>(a.x);
    const double* xg = static_cast<const double*>(plan.xg);
    const int n32 = (int)a.n;
    bk_mask_pat<double> pat;
    int cur = -1;
    const int warps = (int)gridDim.x * BK_WARPS;
    for (int i = (int)blockIdx.x * BK_WARPS + wid; i < plan.n_deferred; i += warps) {
      const int c0 = __ldg(plan.deferred + i);
      const unsigned int m = __ldg(plan.masks + (size_t)c0 * 32 + lane);
      const int slot = __ldg(plan.pids + c0) & (BK_MASK_PID_GHOST - 1);
      if (slot != cur) {
        bk_mask_load_pattern<double>(plan.ptab, slot, pat);
        cur = slot;
      }
      bk_mask_chunk<double, MODE, DOTS, true, false>(a, pat, x, xg, c0 * 32 + lane, m, n32, t);
    }
  }
  return make_double2(t[0], t[1]);
}

template <int MODE, int DOTS, int LEN, unsigned int ODD, bool GHOST, int MINB, typename Epi>
__global__ void __launch_bounds__(BK_BLOCK, MINB)
bk_spmv_mask2_kernel(const bk_spmv_args a, const bk_mask_plan plan, const bk_mask2_plan up,
                     const __grid_constant__ bk_mask_utab ct, const bk_scratch sc, Epi epi) {
  if (bk_spmv_skip(a)) return;
  using T = double;
  constexpr int R = bk_ndots<DOTS>::value;
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int n32 = (int)a.n;
  const int nblk = (n32 + 255) >> 8;
  const int ngroups = (nblk + 7) >> 3;
  int reverse = a.reverse;
  if (a.use_parity) reverse ^= (a.st->parity & 1);
  const T* __restrict__ x = static_cast<const T*>(a.x);

  double acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.0;

  const int my_visits = (ngroups > (int)blockIdx.x) ? (ngroups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto group_of = [&](int visit) -> int {
    const int g = (int)blockIdx.x + visit * (int)gridDim.x;
    return reverse ? (ngroups - 1 - g) : g;
  };
  auto prefetch_group = [&](int visit) {  // see kernel 6: the x range of the group this CTA visits next -> L2
    if (visit >= my_visits) return;
    const long long r0 = (long long)group_of(visit) * 2048;
    long long r1 = r0 + 2048;
    const long long nal = (long long)n32 & ~1LL;
    if (r1 > nal) r1 = nal;
    if (r1 > r0) bk_bulk_prefetch_l2(x + r0, (uint32_t)((r1 - r0) * sizeof(T)));
  };
  const bool pf = plan.prefetch != 0 && threadIdx.x == 0;
  if (pf) prefetch_group(1);
  // the warp's four step summaries of a group are stored together (one 8-byte load); those of the NEXT visit are loaded
  // one visit ahead
  uint2 sm_next = make_uint2(0u, 0u);
  if (my_visits > 0) sm_next = __ldg(reinterpret_cast<const uint2*>(up.usum) + group_of(0) * 8 + wid);
  for (int visit = 0; visit < my_visits; ++visit) {
    const int gg = group_of(visit);
    if (pf) prefetch_group(visit + 2);
    // this lane's rows in step j: row0 + 512 j, row0 + 512 j + 1
    const int row0 = gg * 2048 + wid * 64 + 2 * lane;
    const uint2 smq = sm_next;
    if (visit + 1 < my_visits) sm_next = __ldg(reinterpret_cast<const uint2*>(up.usum) + group_of(visit + 1) * 8 + wid);
    unsigned int sm[4];
    sm[0] = smq.x & 0xffffu;
    sm[1] = smq.x >> 16;
    sm[2] = smq.y & 0xffffu;
    sm[3] = smq.y >> 16;
    const T* xr = x + row0;
    const unsigned char* pm = up.umasks + row0;
    T* py = static_cast<T*>(a.y) + row0;
    const T* pb = MODE == 1 ? static_cast<const T*>(a.b) + row0 : nullptr;
    const T* pw = (DOTS & 1) ? static_cast<const T*>(a.w) + row0 : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned int s = sm[j];
      if ((s & (BK_MASK_US_DIRTY | BK_MASK_US_MIXED)) == 0) {
        bk_mask2_step<MODE, DOTS, LEN, ODD, false>(ct, s, xr + j * 512, pm + j * 512, py + j * 512,
                                                  MODE == 1 ? pb + j * 512 : nullptr, (DOTS & 1) ? pw + j * 512 : nullptr,
                                                  lane, acc);
      } else if ((s & BK_MASK_US_MIXED) == 0) {
        bk_mask2_step<MODE, DOTS, LEN, ODD, true>(ct, s, xr + j * 512, pm + j * 512, py + j * 512,
                                                 MODE == 1 ? pb + j * 512 : nullptr, (DOTS & 1) ? pw + j * 512 : nullptr,
                                                 lane, acc);
      } else {
        const double2 t = bk_mask2_mixed_step<MODE, DOTS, GHOST>(a, plan, gg * 2048 + j * 512 + wid * 64, lane);
        if constexpr (DOTS != 0) acc[0] += t.x;
        if constexpr (R == 2) acc[1] += t.y;
      }
    }
  }
  if constexpr (GHOST) {
    const double2 t = bk_mask2_ghost_phase<MODE, DOTS>(a, plan);
    if constexpr (DOTS != 0) acc[0] += t.x;
    if constexpr (R == 2) acc[1] += t.y;
  }
  if constexpr (DOTS != 0) {
    bk_grid_reduce<R, Epi, BK_WARPS>(acc, sc, epi);
  }
}
