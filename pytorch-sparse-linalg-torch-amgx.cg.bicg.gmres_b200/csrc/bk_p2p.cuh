// bk_p2p.cuh — peer-memory (NVLink / NVSwitch) primitives of the multi-GPU path: a one-shot all-reduce of a few
// doubles and halo-arrival flags, both through windows of device memory that every rank maps with CUDA IPC.
//
// Window layout (identical on every rank, so a peer's addresses follow from its base pointer):
//   [   0,  512)   reserved
//   [ 512,  640)   halo flags: one 64-bit sequence number per SENDER rank (st.release.sys / ld.acquire.sys)
//   [1024, 140288) all-reduce slots: [2 buffers][16 ranks][272 values] x 16 bytes, "LL" format — each 8-byte word
//                  carries 32 bits of payload + the 32-bit sequence number, so one atomic 8-byte store publishes data
//                  and flag together and no fence is needed on the payload path
//   [140288, ...)  ghost vector (halo landing zone: neighbours store their boundary entries straight into it)
//
// All-reduce: every rank stores its values into slot [seq & 1][own rank] of EVERY rank's window, then reads the P
// slots of its own window in rank order and adds them in that order => the sums are bitwise identical on all ranks
// and from run to run (the stop tests derived from them therefore agree everywhere).  Two buffers are enough: a rank
// can be at most one all-reduce ahead of any other because completing all-reduce s needs every rank's value s.
// Every wait has a ~10 s timeout that turns a lost peer into an error status instead of a hung GPU.
#pragma once

#include "bk_internal.cuh"

#define BK_P2P_MAXP 16
#define BK_P2P_FLAG_OFF 512
#define BK_P2P_VEC_OFF 1024
#define BK_P2P_VEC_MAX 272  // >= largest GMRES restart + 1
#define BK_P2P_GHOST_OFF (BK_P2P_VEC_OFF + 2 * BK_P2P_MAXP * BK_P2P_VEC_MAX * 16)
#define BK_P2P_TIMEOUT_CYCLES 20000000000LL
#define BK_ST_COMM_TIMEOUT (-20)

struct bk_p2p_ctx {
  int P;                     // 0 => peer path disabled
  int rank;
  unsigned int* counters;    // device-local: [1] halo send seq [2] halo recv seq [3] push ticket [4] error
                             //               [5] all-reduce seq
  char* win[BK_P2P_MAXP];    // window base of every rank, own entry included
};

#ifdef __CUDACC__
__device__ __forceinline__ void bk_st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long bk_ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void bk_st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long bk_ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// All-reduce (sum) of `count` <= BK_P2P_VEC_MAX doubles, run by `nthreads` threads of ONE CTA per rank: thread `tid`
// owns elements tid, tid + nthreads, ...  `seq` is this all-reduce's sequence number (counters[5] + 1; the caller
// publishes it afterwards).  `in` and `out` may alias.  Two all-reduces must be separated by a kernel boundary or a
// block barrier.
__device__ __forceinline__ void bk_p2p_allreduce_vec(const bk_p2p_ctx& c, unsigned int seq, const double* in,
                                                     double* out, int count, int tid, int nthreads) {
  const size_t buf = (size_t)(seq & 1u);
  const unsigned long long tag = (unsigned long long)seq << 32;
  for (int i = tid; i < count; i += nthreads) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(in[i]);
    const unsigned long long w0 = tag | (bits & 0xffffffffULL);
    const unsigned long long w1 = tag | (bits >> 32);
    for (int q = 0; q < c.P; ++q) {
      unsigned long long* slot = reinterpret_cast<unsigned long long*>(
          c.win[q] + BK_P2P_VEC_OFF + ((buf * BK_P2P_MAXP + c.rank) * BK_P2P_VEC_MAX + i) * 16);
      bk_st_volatile_u64(slot, w0);
      bk_st_volatile_u64(slot + 1, w1);
    }
  }
  const long long t0 = clock64();
  for (int i = tid; i < count; i += nthreads) {
    double sum = 0.0;
    for (int q = 0; q < c.P; ++q) {
      const unsigned long long* slot = reinterpret_cast<const unsigned long long*>(
          c.win[c.rank] + BK_P2P_VEC_OFF + ((buf * BK_P2P_MAXP + q) * BK_P2P_VEC_MAX + i) * 16);
      unsigned long long a, b;
      for (;;) {
        a = bk_ld_volatile_u64(slot);
        b = bk_ld_volatile_u64(slot + 1);
        if ((unsigned int)(a >> 32) == seq && (unsigned int)(b >> 32) == seq) break;
        if (clock64() - t0 > BK_P2P_TIMEOUT_CYCLES) {
          c.counters[4] = 1u;
          a = 0;
          b = 0x7ff80000ULL;  // NaN
          break;
        }
      }
      sum += __longlong_as_double((long long)((b << 32) | (a & 0xffffffffULL)));
    }
    out[i] = sum;
  }
}

// ---- how a reducing kernel's LOCAL sums become GLOBAL ones ------------------------------------------------------
// mode 0: one GPU, the sums already are global.  mode 1: peer memory — the epilogue thread all-reduces them in place
// and goes on with the solver's scalar step.  mode 2: NCCL — the epilogue parks them in `red`; the host enqueues
// ncclAllReduce and a one-thread kernel that runs the same scalar step on the global sums.
struct bk_gsum {
  int mode;
  double* red;
  bk_p2p_ctx p2p;
};

// Called by ONE thread.  Returns true when `out` holds the global sums and the scalar step should run now.
template <int R>
__device__ __forceinline__ bool bk_gsum_finish(const bk_gsum& g, bk_dev_state* st, const double* s, double* out) {
  if (g.mode == 0) {
#pragma unroll
    for (int i = 0; i < R; ++i) out[i] = s[i];
    return true;
  }
  if (g.mode == 2) {
#pragma unroll
    for (int i = 0; i < R; ++i) g.red[i] = s[i];
    return false;
  }
  const unsigned int seq = g.p2p.counters[5] + 1u;
  bk_p2p_allreduce_vec(g.p2p, seq, s, out, R, 0, 1);
  g.p2p.counters[5] = seq;
  if (g.p2p.counters[4]) {
    st->done = 1;
    st->status = BK_ST_COMM_TIMEOUT;
    return false;
  }
  return true;
}
#endif
