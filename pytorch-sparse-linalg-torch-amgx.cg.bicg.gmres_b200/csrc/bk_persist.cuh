// bk_persist.cuh — shared pieces of the persistent (one-kernel-per-solve) solvers for launch-latency-bound systems:
// the kernel-wide barrier (cooperative grid barrier, or the hardware barrier of a thread-block cluster when the whole
// grid is one cluster), deterministic block sums into per-CTA partials and their replicated, fixed-order gather.
#pragma once

#include <cooperative_groups.h>

#include "bk_internal.cuh"

#define BK_GP_BLOCK 1024
#define BK_GP_WARPS (BK_GP_BLOCK / 32)
#define BK_GP_NV 8  // projection coefficients reduced per pass

// block-wide sum of NV (1 or 8) values per thread -> partials[(base + v) * BK_MAXB + blockIdx.x], v < nv.
// NV = 8 uses a butterfly: after the exchanges over lane distances 16, 8, 4 every lane holds ONE of the 8 values summed
// over 8 lanes, two more steps finish the warp sum — 9 double-word shuffles instead of 40.  (With 1024-thread CTAs the
// shuffle pipe, not the barrier, bounded the first version of the persistent GMRES: 16 of 18 us per Arnoldi step.)
template <int NV>
__device__ __forceinline__ void bk_gp_block_sums(double (&acc)[NV], int nv, double* sh /* NV * BK_GP_WARPS */,
                                                 double* partials, int base) {
  static_assert(NV == 1 || NV == 8, "1 or 8 values");
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if constexpr (NV == 1) {
    double t = acc[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) sh[wid] = t;
  } else {
    const bool h1 = (lane & 16) != 0, h2 = (lane & 8) != 0, h3 = (lane & 4) != 0;
    double a4[4], a2[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double send = h1 ? acc[k] : acc[k + 4];
      const double keep = h1 ? acc[k + 4] : acc[k];
      a4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double send = h2 ? a4[k] : a4[k + 2];
      const double keep = h2 ? a4[k + 2] : a4[k];
      a2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    double a1 = (h3 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, h3 ? a2[0] : a2[1], 4);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
    const int vi = (h1 ? 4 : 0) + (h2 ? 2 : 0) + (h3 ? 1 : 0);
    if ((lane & 3) == 0) sh[vi * BK_GP_WARPS + wid] = a1;
  }
  __syncthreads();
  if (wid < nv) {
    double t = sh[wid * BK_GP_WARPS + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane == 0) __stcg(&partials[(size_t)(base + wid) * BK_MAXB + blockIdx.x], t);
  }
  __syncthreads();
}

// s_out[i] = sum over CTAs of partials[(base + i)][cta], i < count — executed by every CTA in the same order
__device__ __forceinline__ void bk_gp_gather_sums(const double* partials, int base, int count, double* s_out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = wid; i < count; i += BK_GP_WARPS) {
    double a = 0.0;
    for (int c = lane; c < (int)gridDim.x; c += 32) a += __ldcg(&partials[(size_t)(base + i) * BK_MAXB + c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if (lane == 0) s_out[i] = a;
  }
  __syncthreads();
}

// Barrier of the whole kernel.  CLUSTER = false: cooperative-groups grid barrier (an atomic round trip through L2,
// ~3 us measured).  CLUSTER = true: the grid IS one thread-block cluster (<= 16 CTAs on one GPC, n <= 16384 rows): the
// hardware cluster barrier (barrier.cluster, ~0.2 us) with release/acquire semantics replaces it — two barriers per
// Arnoldi step make this the difference between 18 and ~11 us per step on the LDC-100 system.
template <bool CLUSTER>
struct bk_gp_barrier {
  cooperative_groups::grid_group grid;
  __device__ bk_gp_barrier() : grid(cooperative_groups::this_grid()) {}
  __device__ __forceinline__ void sync() {
    if constexpr (CLUSTER) {
      asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
      grid.sync();
    }
  }
};


// Launch `kern(args)` with one row per thread: as ONE thread-block cluster when the grid fits 16 CTAs (hardware
// barrier), else as a cooperative grid.  Returns 1 in *launched when a persistent launch was made.
template <typename KC, typename KG, typename Args>
static int bk_launch_persistent(bk_handle* h, KC kern_cluster, KG kern_grid, const Args& ga, long long n,
                                cudaStream_t s, bool* launched) {
  *launched = false;
  int coop = 0, per_sm = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
  if (!h->persistent || !coop || n > (long long)h->persistent_max_n) return BK_OK;
  const long long grid = (n + BK_GP_BLOCK - 1) / BK_GP_BLOCK;
  if (h->persistent_cluster && grid <= 16) {
    const unsigned csize = grid <= 8 ? (unsigned)grid : 16u;  // > 8 CTAs per cluster is the opt-in (non-portable) size
    cudaError_t ce = cudaSuccess;
    if (csize > 8) ce = cudaFuncSetAttribute(kern_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (ce == cudaSuccess) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(csize);
      cfg.blockDim = dim3(BK_GP_BLOCK);
      cfg.stream = s;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = csize;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int nclusters = 0;
      ce = cudaOccupancyMaxActiveClusters(&nclusters, kern_cluster, &cfg);
      if (ce == cudaSuccess && nclusters >= 1) ce = cudaLaunchKernelEx(&cfg, kern_cluster, ga);
      else if (ce == cudaSuccess) ce = cudaErrorInvalidConfiguration;
    }
    if (ce == cudaSuccess) {
      *launched = true;
      h->last_loop_mode = 3;
      return BK_OK;
    }
    cudaGetLastError();  // cluster launch unavailable: the cooperative grid version below
  }
  BK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern_grid, BK_GP_BLOCK, 0));
  if (per_sm > 0 && grid <= (long long)per_sm * h->num_sms && grid <= BK_MAXB) {
    void* args[] = {(void*)&ga};
    BK_CUDA(cudaLaunchCooperativeKernel((const void*)kern_grid, dim3((unsigned)grid), dim3(BK_GP_BLOCK), args, 0, s));
    *launched = true;
    h->last_loop_mode = 3;
  }
  return BK_OK;
}
