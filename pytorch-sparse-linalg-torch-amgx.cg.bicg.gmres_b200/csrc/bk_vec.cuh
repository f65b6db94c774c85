// bk_vec.cuh — fused BLAS-1 kernels: every axpy-type update of the recurrences is ONE pass over
// its vectors with the norm/dot that follows it folded in, ending in the deterministic grid
// reduction whose last CTA computes the next scalars (alpha, beta, omega, stop flags) on device.
// Replaces the reference's _add/_sub/_mul temporaries (torch_sparse_linalg.py:165-173) and the
// separate torch.vdot launches (:86-139).
//
// One generic persistent kernel, parameterised by an Op:
//   Op::R                      number of fused reductions (0 = none)
//   Op::In<W>                  the input packs of one step
//   skip()                     device-side guard (flag set by an earlier epilogue)
//   prepare() -> Ctx           load the scalars of this step from bk_dev_state
//   load<W>(i, in)             all global loads of pack i (issued for UN packs before any store)
//   apply<W>(i, in, ctx, acc)  arithmetic + stores + reduction terms
//   epilogue(sums)             runs once, on the last CTA
// fp64 packs are double2, fp32 packs float4 (16-byte accesses); W=1 instantiations serve
// misaligned user pointers and the n % W tail.
#pragma once

#include "bk_internal.cuh"
#include "bk_p2p.cuh"

#include <type_traits>

// An op that declares `static constexpr bool kCtxLoad = true` gets its Ctx passed to load() as well (ops whose operands
// depend on a device-side flag).
template <typename Op, typename = void>
struct bk_has_ctx_load : std::false_type {};
template <typename Op>
struct bk_has_ctx_load<Op, std::void_t<decltype(Op::kCtxLoad)>> : std::true_type {};

template <int W, typename Op, typename In, typename Ctx>
__device__ __forceinline__ void bk_ew_load(const Op& op, long long i, In& in, const Ctx& ctx) {
  if constexpr (bk_has_ctx_load<Op>::value) op.template load<W>(i, in, ctx);
  else op.template load<W>(i, in);
}

template <typename T, int W, typename Op>
__global__ void __launch_bounds__(BK_BLOCK, 3) bk_ew_kernel(Op op, const long long n, const bk_scratch sc) {
  if (op.skip()) return;
  const typename Op::Ctx ctx = op.prepare();
  constexpr int R = Op::R > 0 ? Op::R : 1;
  constexpr int UN = 2;
  double acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.0;
  const long long npack = n / W;
  const long long stride = (long long)gridDim.x * BK_BLOCK;
  const bool rev = op.reverse();
  long long i = (long long)blockIdx.x * BK_BLOCK + threadIdx.x;
  for (; i + (UN - 1) * stride < npack; i += UN * stride) {
    typename Op::template In<W> in[UN];
    long long p[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      p[u] = i + u * stride;
      if (rev) p[u] = npack - 1 - p[u];
      bk_ew_load<W>(op, p[u] * W, in[u], ctx);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) op.template apply<W>(p[u] * W, in[u], ctx, acc);
  }
  for (; i < npack; i += stride) {
    typename Op::template In<W> in;
    const long long p = rev ? (npack - 1 - i) : i;
    bk_ew_load<W>(op, p * W, in, ctx);
    op.template apply<W>(p * W, in, ctx, acc);
  }
  if constexpr (W > 1) {
    const long long t = npack * W + (long long)blockIdx.x * BK_BLOCK + threadIdx.x;
    if (t < n) {
      typename Op::template In<1> in;
      bk_ew_load<1>(op, t, in, ctx);
      op.template apply<1>(t, in, ctx, acc);
    }
  }
  if constexpr (Op::R > 0) {
    bk_grid_reduce<R>(acc, sc, [&](const double* s) { op.epilogue(s); });
  }
}

template <typename T, typename Op>
static int bk_launch_ew(bk_handle* h, const Op& op, long long n, bool aligned, const bk_scratch& sc,
                        cudaStream_t s) {
  constexpr int NW = bk_native_w<T>::value;
  const int grid = bk_grid_vec_n(h, n, 2 * NW);
  if (aligned) {
    bk_ew_kernel<T, NW, Op><<<grid, BK_BLOCK, 0, s>>>(op, n, sc);
  } else {
    bk_ew_kernel<T, 1, Op><<<grid, BK_BLOCK, 0, s>>>(op, n, sc);
  }
  BK_KERNEL_CHECK();
  return BK_OK;
}

struct bk_noctx {};

// ---- generic building blocks ---------------------------------------------------------------
// out[0] = x . y  (sqrt_out: out[0] = sqrt(max(x.y, 0)) — _norm :154-162)
template <typename T>
struct bk_op_dot {
  static constexpr int R = 1;
  using Ctx = bk_noctx;
  template <int W>
  struct In {
    bk_vec<T, W> a, b;
  };
  const T* x;
  const T* y;
  double* out;
  int sqrt_out;
  __device__ bool skip() const { return false; }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const { return Ctx(); }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.a = bk_ld<T, W>(x + i);
    in.b = bk_ld<T, W>(y + i);
  }
  template <int W>
  __device__ void apply(long long, const In<W>& in, const Ctx&, double (&acc)[1]) const {
#pragma unroll
    for (int j = 0; j < W; ++j) acc[0] += (double)in.a.v[j] * (double)in.b.v[j];
  }
  __device__ void epilogue(const double* s) const { out[0] = sqrt_out ? sqrt(fmax(s[0], 0.0)) : s[0]; }
};

// x . y reduced on the device, then `epi(sums)` on the last CTA (solver-specific scalar work)
template <typename T, typename Epi>
struct bk_op_dot_epi {
  static constexpr int R = 1;
  using Ctx = bk_noctx;
  template <int W>
  struct In {
    bk_vec<T, W> a, b;
  };
  const T* x;
  const T* y;
  Epi epi;
  const bk_dev_state* st = nullptr;  // with guard != 0: skip like the other guarded kernels (bk_guard_skip)
  int guard = 0;
  __device__ bool skip() const {
    if (guard == 0) return false;
    if (st->done) return true;
    if (guard == 2 && st->g_cycle_over) return true;
    if (guard == 3 && st->exit_early) return true;
    return false;
  }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const { return Ctx(); }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.a = bk_ld<T, W>(x + i);
    in.b = bk_ld<T, W>(y + i);
  }
  template <int W>
  __device__ void apply(long long, const In<W>& in, const Ctx&, double (&acc)[1]) const {
#pragma unroll
    for (int j = 0; j < W; ++j) acc[0] += (double)in.a.v[j] * (double)in.b.v[j];
  }
  __device__ void epilogue(const double* s) const { epi(s); }
};

template <typename T, typename Epi>
static int bk_dot_epi(bk_handle* h, long long n, const void* x, const void* y, Epi epi, int slot, cudaStream_t s) {
  bk_op_dot_epi<T, Epi> op;
  op.x = (const T*)x;
  op.y = (const T*)y;
  op.epi = epi;
  return bk_launch_ew<T>(h, op, n, bk_aligned16(x) && bk_aligned16(y), bk_slot(h, slot), s);
}

// z = a x + b y
template <typename T>
struct bk_op_axpby {
  static constexpr int R = 0;
  using Ctx = bk_noctx;
  template <int W>
  struct In {
    bk_vec<T, W> a, b;
  };
  const T* x;
  const T* y;
  T* z;
  T ca, cb;
  __device__ bool skip() const { return false; }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const { return Ctx(); }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.a = bk_ld<T, W>(x + i);
    in.b = bk_ld<T, W>(y + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx&, double (&)[1]) const {
    bk_vec<T, W> o;
#pragma unroll
    for (int j = 0; j < W; ++j) o.v[j] = bk_add(bk_mul(ca, in.a.v[j]), bk_mul(cb, in.b.v[j]));
    bk_st<T, W>(z + i, o);
  }
  __device__ void epilogue(const double*) const {}
};

// z = (*a) x + (*b) y with the scalars read from DEVICE memory (fp64; a null pointer means 1): lets the callable-A route
// keep alpha / beta on the device instead of synchronising for every dot product (generic.py)
template <typename T>
struct bk_op_axpby_dev {
  static constexpr int R = 0;
  struct Ctx {
    T ca, cb;
  };
  template <int W>
  struct In {
    bk_vec<T, W> a, b;
  };
  const T* x;
  const T* y;
  T* z;
  const double* pa;
  const double* pb;
  double sa, sb;  // sign / constant factor applied to the loaded scalar (exact: +-1)
  __device__ bool skip() const { return false; }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const {
    Ctx c;
    c.ca = static_cast<T>(pa ? sa * pa[0] : sa);
    c.cb = static_cast<T>(pb ? sb * pb[0] : sb);
    return c;
  }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.a = bk_ld<T, W>(x + i);
    in.b = bk_ld<T, W>(y + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx& c, double (&)[1]) const {
    bk_vec<T, W> o;
#pragma unroll
    for (int j = 0; j < W; ++j) o.v[j] = bk_add(bk_mul(c.ca, in.a.v[j]), bk_mul(c.cb, in.b.v[j]));
    bk_st<T, W>(z + i, o);
  }
  __device__ void epilogue(const double*) const {}
};

// z = x / d  (true division, as the reference's `y / norm` in _safe_normalize :268)
template <typename T>
struct bk_op_div {
  static constexpr int R = 0;
  using Ctx = bk_noctx;
  template <int W>
  struct In {
    bk_vec<T, W> a;
  };
  const T* x;
  T* z;
  T d;
  __device__ bool skip() const { return false; }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const { return Ctx(); }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.a = bk_ld<T, W>(x + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx&, double (&)[1]) const {
    bk_vec<T, W> o;
#pragma unroll
    for (int j = 0; j < W; ++j) o.v[j] = in.a.v[j] / d;
    bk_st<T, W>(z + i, o);
  }
  __device__ void epilogue(const double*) const {}
};

// ---- complex128 vectors as interleaved (re, im) doubles: one fp64 pack (W = 2) is one complex number ----------------
// out[0] + i out[1] = sum conj(x_k) y_k      (torch.vdot on complex tensors, reference _vdot :86-91)
struct bk_op_cdot {
  static constexpr int R = 2;
  using Ctx = bk_noctx;
  template <int W>
  struct In {
    bk_vec<double, W> a, b;
  };
  const double* x;
  const double* y;
  double* out;
  __device__ bool skip() const { return false; }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const { return Ctx(); }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.a = bk_ld<double, W>(x + i);
    in.b = bk_ld<double, W>(y + i);
  }
  template <int W>
  __device__ void apply(long long, const In<W>& in, const Ctx&, double (&acc)[2]) const {
    if constexpr (W == 2) {  // one (re, im) pack; the scalar instantiation is never launched (2n is even, aligned)
      acc[0] += in.a.v[0] * in.b.v[0] + in.a.v[1] * in.b.v[1];
      acc[1] += in.a.v[0] * in.b.v[1] - in.a.v[1] * in.b.v[0];
    }
  }
  __device__ void epilogue(const double* s) const {
    out[0] = s[0];
    out[1] = s[1];
  }
};

// z = a x + b y with complex scalars a, b (each product and sum rounded separately, like the reference's torch ops)
struct bk_op_caxpby {
  static constexpr int R = 0;
  using Ctx = bk_noctx;
  template <int W>
  struct In {
    bk_vec<double, W> a, b;
  };
  const double* x;
  const double* y;
  double* z;
  double ar, ai, br, bi;
  __device__ bool skip() const { return false; }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const { return Ctx(); }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.a = bk_ld<double, W>(x + i);
    in.b = bk_ld<double, W>(y + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx&, double (&)[1]) const {
    if constexpr (W == 2) {  // one (re, im) pack; the scalar instantiation is never launched (2n is even, aligned)
      const double xr = in.a.v[0], xi = in.a.v[1], yr = in.b.v[0], yi = in.b.v[1];
      const double axr = bk_sub(bk_mul(ar, xr), bk_mul(ai, xi)), axi = bk_add(bk_mul(ar, xi), bk_mul(ai, xr));
      const double byr = bk_sub(bk_mul(br, yr), bk_mul(bi, yi)), byi = bk_add(bk_mul(br, yi), bk_mul(bi, yr));
      bk_vec<double, W> o;
      o.v[0] = bk_add(axr, byr);
      o.v[1] = bk_add(axi, byi);
      bk_st<double, W>(z + i, o);
    }
  }
  __device__ void epilogue(const double*) const {}
};

// ---- CG --------------------------------------------------------------------------------------
// x += alpha p ; r -= alpha Ap ; gamma' = r.r          (_cg_solve :846-850)
// epilogue: beta = gamma'/gamma, gamma = gamma', k += 1, stop test of :841 for the NEXT iteration.
// (used by the small-system variant whose SpMV forms p = r + beta p on the fly; large systems use the cut below)
template <typename T>
struct bk_op_cg_update {
  static constexpr int R = 1;
  struct Ctx {
    T alpha;
  };
  template <int W>
  struct In {
    bk_vec<T, W> p, ap, x, r;
  };
  const T* p;
  const T* ap;
  T* x;
  T* r;
  bk_dev_state* st;
  int snake;
  __device__ bool skip() const { return st->done != 0; }
  __device__ bool reverse() const { return snake && ((st->parity & 1) == 0); }
  __device__ Ctx prepare() const {
    Ctx c;
    c.alpha = static_cast<T>(st->alpha);
    return c;
  }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.p = bk_ld<T, W>(p + i);
    in.ap = bk_ld<T, W>(ap + i);
    in.x = bk_ld<T, W>(x + i);
    in.r = bk_ld<T, W>(r + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx& c, double (&acc)[1]) const {
    bk_vec<T, W> xo, ro;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      xo.v[j] = bk_add(in.x.v[j], bk_mul(c.alpha, in.p.v[j]));
      ro.v[j] = bk_sub(in.r.v[j], bk_mul(c.alpha, in.ap.v[j]));
      acc[0] += (double)ro.v[j] * (double)ro.v[j];
    }
    bk_st<T, W>(x + i, xo);
    bk_st<T, W>(r + i, ro);
  }
  __device__ void epilogue(const double* s) const {
    const double gamma_new = s[0];
    st->beta = gamma_new / st->gamma;
    st->gamma = gamma_new;
    const long long k = st->k + 1;
    st->k = k;
    st->parity ^= 1;
    if (k >= st->maxiter) {
      st->done = 1;
      st->status = BK_ST_MAXITER;
    }
    if (gamma_new <= st->atol2) {
      st->done = 1;
      st->status = BK_ST_CONVERGED;
    }
  }
};

// The same CG iteration cut differently (large systems): K2 touches only r, K3 updates x together with p — both read
// p_k anyway, so x costs one vector pass less (8n instead of 9n per iteration).  Bitwise the same x, r, p.
//   K2  r -= alpha Ap ; gamma' = r.r ; epilogue as above                       (:847, :849-851)
//   K3  x += alpha p ; p = r + beta p                                          (:846, :852)
// K3 must also run in the iteration whose K2 set `done` (x is not final before): K2's epilogue raises `just_done`,
// and a K2 that finds `done` already set (every later, skipped iteration) lowers it again.
template <typename T>
struct bk_op_cg_r {
  static constexpr int R = 1;
  struct Ctx {
    T alpha;
    unsigned long long pol;
  };
  template <int W>
  struct In {
    bk_vec<T, W> ap, r;
  };
  const T* ap;
  T* r;
  bk_dev_state* st;
  int snake;
  int hints = 0;  // bit 0: Ap is dead after this pass -> streaming loads | bit 4: r is stored "evict last"
  __device__ bool skip() const {
    if (st->done == 0) return false;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->just_done = 0;
    return true;
  }
  __device__ bool reverse() const { return snake && ((st->parity & 1) == 0); }
  __device__ Ctx prepare() const {
    Ctx c;
    c.alpha = static_cast<T>(st->alpha);
    c.pol = (hints & 16) ? bk_policy_evict_last() : 0ull;
    return c;
  }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.ap = (hints & 1) ? bk_ld_cs<T, W>(ap + i) : bk_ld<T, W>(ap + i);
    in.r = bk_ld<T, W>(r + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx& c, double (&acc)[1]) const {
    bk_vec<T, W> ro;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      ro.v[j] = bk_sub(in.r.v[j], bk_mul(c.alpha, in.ap.v[j]));
      acc[0] += (double)ro.v[j] * (double)ro.v[j];
    }
    if (hints & 16) bk_st_keep<T, W>(r + i, ro, c.pol); else bk_st<T, W>(r + i, ro);
  }
  __device__ void epilogue(const double* s) const {
    const double gamma_new = s[0];
    st->beta = gamma_new / st->gamma;
    st->gamma = gamma_new;
    const long long k = st->k + 1;
    st->k = k;
    st->parity ^= 1;
    if (k >= st->maxiter) {
      st->done = 1;
      st->just_done = 1;
      st->status = BK_ST_MAXITER;
    }
    if (gamma_new <= st->atol2) {
      st->done = 1;
      st->just_done = 1;
      st->status = BK_ST_CONVERGED;
    }
  }
};

template <typename T>
struct bk_op_cg_xp {
  static constexpr int R = 0;
  struct Ctx {
    T alpha, beta;
    unsigned long long pol;
  };
  template <int W>
  struct In {
    bk_vec<T, W> x, p, r;
  };
  T* x;
  T* p;
  const T* r;
  const bk_dev_state* st;
  int snake;
  int hints = 0;  // bit 1: x streams through (next touched one iteration later) | bit 2: so does r
  __device__ bool skip() const { return st->done != 0 && st->just_done == 0; }
  __device__ bool reverse() const { return snake && ((st->parity & 1) == 0); }
  __device__ Ctx prepare() const {
    Ctx c;
    c.alpha = static_cast<T>(st->alpha);
    c.beta = static_cast<T>(st->beta);
    c.pol = (hints & 16) ? bk_policy_evict_last() : 0ull;
    return c;
  }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.x = (hints & 2) ? bk_ld_cs<T, W>(x + i) : bk_ld<T, W>(x + i);
    in.p = bk_ld<T, W>(p + i);
    in.r = (hints & 4) ? bk_ld_cs<T, W>(r + i) : bk_ld<T, W>(r + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx& c, double (&)[1]) const {
    bk_vec<T, W> xo, po;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      xo.v[j] = bk_add(in.x.v[j], bk_mul(c.alpha, in.p.v[j]));
      po.v[j] = bk_add(in.r.v[j], bk_mul(c.beta, in.p.v[j]));
    }
    if (hints & 2) bk_st_cs<T, W>(x + i, xo); else bk_st<T, W>(x + i, xo);
    if (hints & 16) bk_st_keep<T, W>(p + i, po, c.pol); else bk_st<T, W>(p + i, po);
  }
  __device__ void epilogue(const double*) const {}
};

// The same iteration with x lagging (large systems, option cg_lag_x): x_{k+1} = x_k + alpha_k p_k feeds nothing but the
// final result, so two consecutive updates are applied together — in the same order, with the same roundings — by the
// K3 of every second iteration, while p ping-pongs between two buffers so that p_k is still there:
//   even iteration  K3e  p_{k+1} = r + beta p_k               (reads p_k, r; writes the other buffer)           3n
//   odd  iteration  K3o  x = (x + alpha_{k-1} p_{k-1}) + alpha_k p_k ; p_{k+1} = r + beta p_k  (into p_{k-1}'s buffer) 6n
// 9n per two iterations instead of 10n.  If the stop test fires in an even iteration, K3e applies the pending term
// instead (x += alpha p_k; `just_done`), so x is complete whenever the loop ends.  Bitwise the same x, r, p.
template <typename T>
struct bk_op_cg_p_lag {
  static constexpr int R = 0;
  static constexpr bool kCtxLoad = true;
  struct Ctx {
    T alpha, beta;
    int flush;
  };
  template <int W>
  struct In {
    bk_vec<T, W> p, q;
  };
  T* x;
  const T* pcur;
  T* pnext;
  const T* r;
  const bk_dev_state* st;
  int snake;
  __device__ bool skip() const { return st->done != 0 && st->just_done == 0; }
  __device__ bool reverse() const { return snake && ((st->parity & 1) == 0); }
  __device__ Ctx prepare() const {
    Ctx c;
    c.alpha = static_cast<T>(st->alpha);
    c.beta = static_cast<T>(st->beta);
    c.flush = st->just_done;
    return c;
  }
  template <int W>
  __device__ void load(long long i, In<W>& in, const Ctx& c) const {
    in.p = bk_ld<T, W>(pcur + i);
    in.q = bk_ld<T, W>((c.flush ? static_cast<const T*>(x) : r) + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx& c, double (&)[1]) const {
    bk_vec<T, W> o;
    if (c.flush) {
#pragma unroll
      for (int j = 0; j < W; ++j) o.v[j] = bk_add(in.q.v[j], bk_mul(c.alpha, in.p.v[j]));
      bk_st<T, W>(x + i, o);
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) o.v[j] = bk_add(in.q.v[j], bk_mul(c.beta, in.p.v[j]));
      bk_st<T, W>(pnext + i, o);
    }
  }
  __device__ void epilogue(const double*) const {}
};

template <typename T>
struct bk_op_cg_xp_lag {
  static constexpr int R = 0;
  struct Ctx {
    T alpha, alpha_lag, beta;
  };
  template <int W>
  struct In {
    bk_vec<T, W> x, pp, pc, r;
  };
  T* x;
  T* pprev;        // p_{k-1}; receives p_{k+1}
  const T* pcur;   // p_k
  const T* r;
  const bk_dev_state* st;
  int snake;
  __device__ bool skip() const { return st->done != 0 && st->just_done == 0; }
  __device__ bool reverse() const { return snake && ((st->parity & 1) == 0); }
  __device__ Ctx prepare() const {
    Ctx c;
    c.alpha = static_cast<T>(st->alpha);
    c.beta = static_cast<T>(st->beta);
    c.alpha_lag = static_cast<T>(st->alpha_lag);
    return c;
  }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.x = bk_ld<T, W>(x + i);
    in.pp = bk_ld<T, W>(pprev + i);
    in.pc = bk_ld<T, W>(pcur + i);
    in.r = bk_ld<T, W>(r + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx& c, double (&)[1]) const {
    bk_vec<T, W> xo, po;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const T x1 = bk_add(in.x.v[j], bk_mul(c.alpha_lag, in.pp.v[j]));
      xo.v[j] = bk_add(x1, bk_mul(c.alpha, in.pc.v[j]));
      po.v[j] = bk_add(in.r.v[j], bk_mul(c.beta, in.pc.v[j]));
    }
    bk_st<T, W>(x + i, xo);
    bk_st<T, W>(pprev + i, po);
  }
  __device__ void epilogue(const double*) const {}
};

// ---- Jacobi-preconditioned CG (M = diag(A)^-1, applied as r / d like the reference's `M = lambda r: r / d`) ------
// _cg_solve with M (:820-853): z = M r ; gamma = r.z ; the stop test uses rs = r.r (:838) ; p = z + beta p.
// z is never stored: K2 forms it on the fly for r.z, K3 again for the p-update (reading d once more costs one
// vector pass; storing and re-reading z would cost two).
template <typename T>
struct bk_op_pcg_init {  // p0 = z0 = r0 / d ; sums: [0] r0.z0  [1] r0.r0    (:821-826)
  static constexpr int R = 2;
  using Ctx = bk_noctx;
  template <int W>
  struct In {
    bk_vec<T, W> r, d;
  };
  const T* r;
  const T* d;
  T* p;
  bk_dev_state* st;
  __device__ bool skip() const { return false; }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const { return Ctx(); }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.r = bk_ld<T, W>(r + i);
    in.d = bk_ld<T, W>(d + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx&, double (&acc)[2]) const {
    bk_vec<T, W> z;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      z.v[j] = in.r.v[j] / in.d.v[j];
      acc[0] += (double)in.r.v[j] * (double)z.v[j];
      acc[1] += (double)in.r.v[j] * (double)in.r.v[j];
    }
    bk_st<T, W>(p + i, z);
  }
  __device__ void epilogue(const double* s) const {  // first stop test (:841) — atol2 was set by the b.b reduction
    st->gamma = s[0];
    st->rs = s[1];
    if (st->maxiter <= 0) {
      st->done = 1;
      st->status = BK_ST_MAXITER;
    }
    if (s[1] <= st->atol2) {
      st->done = 1;
      st->status = BK_ST_CONVERGED;
    }
  }
};

template <typename T>
struct bk_op_pcg_r {  // r -= alpha Ap ; z = r / d ; sums: [0] r.z  [1] r.r    (:847-850, :838)
  static constexpr int R = 2;
  struct Ctx {
    T alpha;
  };
  template <int W>
  struct In {
    bk_vec<T, W> ap, r, d;
  };
  const T* ap;
  T* r;
  const T* d;
  bk_dev_state* st;
  __device__ bool skip() const {
    if (st->done == 0) return false;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->just_done = 0;
    return true;
  }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const {
    Ctx c;
    c.alpha = static_cast<T>(st->alpha);
    return c;
  }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.ap = bk_ld<T, W>(ap + i);
    in.r = bk_ld<T, W>(r + i);
    in.d = bk_ld<T, W>(d + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx& c, double (&acc)[2]) const {
    bk_vec<T, W> ro;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      ro.v[j] = bk_sub(in.r.v[j], bk_mul(c.alpha, in.ap.v[j]));
      const T z = ro.v[j] / in.d.v[j];
      acc[0] += (double)ro.v[j] * (double)z;
      acc[1] += (double)ro.v[j] * (double)ro.v[j];
    }
    bk_st<T, W>(r + i, ro);
  }
  __device__ void epilogue(const double* s) const {
    const double gamma_new = s[0];
    st->beta = gamma_new / st->gamma;
    st->gamma = gamma_new;
    st->rs = s[1];
    const long long k = st->k + 1;
    st->k = k;
    if (k >= st->maxiter) {
      st->done = 1;
      st->just_done = 1;
      st->status = BK_ST_MAXITER;
    }
    if (s[1] <= st->atol2) {
      st->done = 1;
      st->just_done = 1;
      st->status = BK_ST_CONVERGED;
    }
  }
};

template <typename T>
struct bk_op_pcg_xp {  // x += alpha p ; p = r / d + beta p    (:846, :852)
  static constexpr int R = 0;
  struct Ctx {
    T alpha, beta;
  };
  template <int W>
  struct In {
    bk_vec<T, W> x, p, r, d;
  };
  T* x;
  T* p;
  const T* r;
  const T* d;
  const bk_dev_state* st;
  __device__ bool skip() const { return st->done != 0 && st->just_done == 0; }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const {
    Ctx c;
    c.alpha = static_cast<T>(st->alpha);
    c.beta = static_cast<T>(st->beta);
    return c;
  }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.x = bk_ld<T, W>(x + i);
    in.p = bk_ld<T, W>(p + i);
    in.r = bk_ld<T, W>(r + i);
    in.d = bk_ld<T, W>(d + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx& c, double (&)[1]) const {
    bk_vec<T, W> xo, po;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      xo.v[j] = bk_add(in.x.v[j], bk_mul(c.alpha, in.p.v[j]));
      po.v[j] = bk_add(in.r.v[j] / in.d.v[j], bk_mul(c.beta, in.p.v[j]));
    }
    bk_st<T, W>(x + i, xo);
    bk_st<T, W>(p + i, po);
  }
  __device__ void epilogue(const double*) const {}
};

template <typename T, typename Epi>
struct bk_op_scaled_sq {  // sum (t / d)^2 — the final ||M (b - A x)||^2 of _isolve (:1008)
  static constexpr int R = 1;
  using Ctx = bk_noctx;
  template <int W>
  struct In {
    bk_vec<T, W> t, d;
  };
  const T* t;
  const T* d;
  Epi epi;
  __device__ bool skip() const { return false; }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const { return Ctx(); }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.t = bk_ld<T, W>(t + i);
    in.d = bk_ld<T, W>(d + i);
  }
  template <int W>
  __device__ void apply(long long, const In<W>& in, const Ctx&, double (&acc)[1]) const {
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const T z = in.t.v[j] / in.d.v[j];
      acc[0] += (double)z * (double)z;
    }
  }
  __device__ void epilogue(const double* s) const { epi(s); }
};

// ---- BiCGStab ----------------------------------------------------------------------------------
// p = r + beta (p - omega q)                               (_bicgstab_solve :906-907)
template <typename T>
struct bk_op_bicg_p {
  static constexpr int R = 0;
  struct Ctx {
    T beta, omega;
  };
  template <int W>
  struct In {
    bk_vec<T, W> r, p, q;
  };
  const T* r;
  T* p;
  const T* q;
  const bk_dev_state* st;
  __device__ bool skip() const { return st->done != 0; }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const {
    Ctx c;
    c.beta = static_cast<T>(st->beta);
    c.omega = static_cast<T>(st->omega);
    return c;
  }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.r = bk_ld<T, W>(r + i);
    in.p = bk_ld<T, W>(p + i);
    in.q = bk_ld<T, W>(q + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx& c, double (&)[1]) const {
    bk_vec<T, W> o;
#pragma unroll
    for (int j = 0; j < W; ++j)
      o.v[j] = bk_add(in.r.v[j], bk_mul(c.beta, bk_sub(in.p.v[j], bk_mul(c.omega, in.q.v[j]))));
    bk_st<T, W>(p + i, o);
  }
  __device__ void epilogue(const double*) const {}
};

// s = r - alpha q ; ss = s.s ; exit_early = ss < atol2      (:917-920)
template <typename T>
struct bk_op_bicg_s {
  static constexpr int R = 1;
  struct Ctx {
    T alpha;
  };
  template <int W>
  struct In {
    bk_vec<T, W> r, q;
  };
  const T* r;
  const T* q;
  T* s;
  bk_dev_state* st;
  __device__ bool skip() const { return st->done != 0; }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const {
    Ctx c;
    c.alpha = static_cast<T>(st->alpha);
    return c;
  }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.r = bk_ld<T, W>(r + i);
    in.q = bk_ld<T, W>(q + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx& c, double (&acc)[1]) const {
    bk_vec<T, W> o;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      o.v[j] = bk_sub(in.r.v[j], bk_mul(c.alpha, in.q.v[j]));
      acc[0] += (double)o.v[j] * (double)o.v[j];
    }
    bk_st<T, W>(s + i, o);
  }
  __device__ void epilogue(const double* sums) const {
    st->ss = sums[0];
    st->exit_early = (sums[0] < st->atol2) ? 1 : 0;
  }
};

// exit_early ? (x += alpha p ; r = s) : (x += alpha p + omega s ; r = s - omega t)   (:942-950)
// fused with rs = r.r and rho' = rhat.r for the next iteration's tests (:894-904).
template <typename T>
struct bk_op_bicg_xr {
  static constexpr int R = 2;
  struct Ctx {
    T alpha, omega;
    int early;
  };
  template <int W>
  struct In {
    bk_vec<T, W> x, p, s, t, rh;
  };
  T* x;
  const T* p;
  const T* s;
  const T* t;
  const T* rhat;
  T* r;
  bk_dev_state* st;
  __device__ bool skip() const { return st->done != 0; }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const {
    Ctx c;
    c.alpha = static_cast<T>(st->alpha);
    c.omega = static_cast<T>(st->omega);
    c.early = st->exit_early;
    return c;
  }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.x = bk_ld<T, W>(x + i);
    in.p = bk_ld<T, W>(p + i);
    in.s = bk_ld<T, W>(s + i);
    in.t = bk_ld<T, W>(t + i);
    in.rh = bk_ld<T, W>(rhat + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx& c, double (&acc)[2]) const {
    bk_vec<T, W> xo, ro;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const T ap = bk_mul(c.alpha, in.p.v[j]);
      if (c.early) {
        xo.v[j] = bk_add(in.x.v[j], ap);
        ro.v[j] = in.s.v[j];
      } else {
        xo.v[j] = bk_add(in.x.v[j], bk_add(ap, bk_mul(c.omega, in.s.v[j])));
        ro.v[j] = bk_sub(in.s.v[j], bk_mul(c.omega, in.t.v[j]));
      }
      acc[0] += (double)ro.v[j] * (double)ro.v[j];
      acc[1] += (double)in.rh.v[j] * (double)ro.v[j];
    }
    bk_st<T, W>(x + i, xo);
    bk_st<T, W>(r + i, ro);
  }
  // End of iteration k and top-of-loop tests of iteration k+1 (:892-905, :952-962).
  __device__ void epilogue(const double* sums) const {
    const double eps = (sizeof(T) == 8) ? 2.220446049250313e-16 : 1.1920928955078125e-07;
    const long long k = st->k + 1;
    st->k = k;
    const double rho_prev = st->rho_new;  // rho = rho_new
    st->rho = rho_prev;
    if (st->exit_early) {  // "if exit_early: break" (:961-962)
      st->rs = sums[0];
      st->done = 1;
      st->status = BK_ST_CONVERGED;
      return;
    }
    if (k >= st->maxiter) {
      st->rs = sums[0];
      st->done = 1;
      st->status = BK_ST_MAXITER;
      return;
    }
    st->rs = sums[0];
    if (sums[0] <= st->atol2) {
      st->done = 1;
      st->status = BK_ST_CONVERGED;
      return;
    }
    const double rho_new = sums[1];
    st->rho_new = rho_new;
    if (fabs(rho_new) < eps * fabs(rho_prev)) {
      st->done = 1;
      st->status = BK_ST_BREAKDOWN_RHO;
      return;
    }
    st->beta = rho_new / rho_prev * st->alpha / st->omega;
  }
};

// w = w / d in place, sum w^2 -> epi   (left Jacobi preconditioning of GMRES: v = M(A v) :351, r = M(b - A x) :491)
template <typename T, typename Epi>
struct bk_op_scale_sq {
  static constexpr int R = 1;
  using Ctx = bk_noctx;
  template <int W>
  struct In {
    bk_vec<T, W> w, d;
  };
  T* w;
  const T* d;
  Epi epi;
  const bk_dev_state* st;
  int guard;  // 0 none | 1 st->done | 2 st->done || st->g_cycle_over
  __device__ bool skip() const {
    if (guard == 0) return false;
    if (st->done) return true;
    if (guard == 2 && st->g_cycle_over) return true;
    return false;
  }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const { return Ctx(); }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.w = bk_ld<T, W>(w + i);
    in.d = bk_ld<T, W>(d + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx&, double (&acc)[1]) const {
    bk_vec<T, W> o;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      o.v[j] = in.w.v[j] / in.d.v[j];
      acc[0] += (double)o.v[j] * (double)o.v[j];
    }
    bk_st<T, W>(w + i, o);
  }
  __device__ void epilogue(const double* s) const { epi(s); }
};

// ---- Jacobi-preconditioned BiCGStab (right preconditioning as in _bicgstab_solve :907-946: phat = M p, q = A phat,
// shat = M s, t = A shat, x += alpha phat + omega shat; r, s, t and all dots stay in residual space) -------------
// The three element-wise kernels additionally read d and write the preconditioned copy the next SpMV gathers from.
template <typename T>
struct bk_op_bicg_p_pc : bk_op_bicg_p<T> {
  using Base = bk_op_bicg_p<T>;
  using Ctx = typename Base::Ctx;
  template <int W>
  struct In {
    bk_vec<T, W> r, p, q, d;
  };
  const T* d;
  T* phat;
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.r = bk_ld<T, W>(this->r + i);
    in.p = bk_ld<T, W>(this->p + i);
    in.q = bk_ld<T, W>(this->q + i);
    in.d = bk_ld<T, W>(d + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx& c, double (&)[1]) const {
    bk_vec<T, W> o, oh;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      o.v[j] = bk_add(in.r.v[j], bk_mul(c.beta, bk_sub(in.p.v[j], bk_mul(c.omega, in.q.v[j]))));
      oh.v[j] = o.v[j] / in.d.v[j];
    }
    bk_st<T, W>(this->p + i, o);
    bk_st<T, W>(phat + i, oh);
  }
};

template <typename T>
struct bk_op_bicg_s_pc : bk_op_bicg_s<T> {
  using Base = bk_op_bicg_s<T>;
  using Ctx = typename Base::Ctx;
  template <int W>
  struct In {
    bk_vec<T, W> r, q, d;
  };
  const T* d;
  T* shat;
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.r = bk_ld<T, W>(this->r + i);
    in.q = bk_ld<T, W>(this->q + i);
    in.d = bk_ld<T, W>(d + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx& c, double (&acc)[1]) const {
    bk_vec<T, W> o, oh;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      o.v[j] = bk_sub(in.r.v[j], bk_mul(c.alpha, in.q.v[j]));
      oh.v[j] = o.v[j] / in.d.v[j];
      acc[0] += (double)o.v[j] * (double)o.v[j];
    }
    bk_st<T, W>(this->s + i, o);
    bk_st<T, W>(shat + i, oh);
  }
};

template <typename T>
struct bk_op_bicg_xr_pc : bk_op_bicg_xr<T> {  // base fields p / s hold phat / s; shat is extra
  using Base = bk_op_bicg_xr<T>;
  using Ctx = typename Base::Ctx;
  template <int W>
  struct In {
    bk_vec<T, W> x, ph, sh, s, t, rh;
  };
  const T* shat;
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.x = bk_ld<T, W>(this->x + i);
    in.ph = bk_ld<T, W>(this->p + i);
    in.sh = bk_ld<T, W>(shat + i);
    in.s = bk_ld<T, W>(this->s + i);
    in.t = bk_ld<T, W>(this->t + i);
    in.rh = bk_ld<T, W>(this->rhat + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx& c, double (&acc)[2]) const {
    bk_vec<T, W> xo, ro;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const T ap = bk_mul(c.alpha, in.ph.v[j]);
      if (c.early) {
        xo.v[j] = bk_add(in.x.v[j], ap);
        ro.v[j] = in.s.v[j];
      } else {
        xo.v[j] = bk_add(in.x.v[j], bk_add(ap, bk_mul(c.omega, in.sh.v[j])));
        ro.v[j] = bk_sub(in.s.v[j], bk_mul(c.omega, in.t.v[j]));
      }
      acc[0] += (double)ro.v[j] * (double)ro.v[j];
      acc[1] += (double)in.rh.v[j] * (double)ro.v[j];
    }
    bk_st<T, W>(this->x + i, xo);
    bk_st<T, W>(this->r + i, ro);
  }
};

// ---- GMRES -------------------------------------------------------------------------------------
// v = use ? w / norm : 0                                  (_safe_normalize :266-272)
template <typename T>
struct bk_op_normalize {
  static constexpr int R = 0;
  struct Ctx {
    T norm;
    int use;
  };
  template <int W>
  struct In {
    bk_vec<T, W> w;
  };
  const T* w;
  T* v;
  const bk_dev_state* st;
  int guard;  // 0 none | 1 st->done | 2 st->done || st->g_cycle_over
  __device__ bool skip() const {
    if (guard == 0) return false;
    if (st->done) return true;
    if (guard == 2 && st->g_cycle_over) return true;
    return false;
  }
  __device__ bool reverse() const { return false; }
  __device__ Ctx prepare() const {
    Ctx c;
    c.norm = static_cast<T>(st->g_scale);
    c.use = st->g_use;
    return c;
  }
  template <int W>
  __device__ void load(long long i, In<W>& in) const {
    in.w = bk_ld<T, W>(w + i);
  }
  template <int W>
  __device__ void apply(long long i, const In<W>& in, const Ctx& c, double (&)[1]) const {
    bk_vec<T, W> o;
#pragma unroll
    for (int j = 0; j < W; ++j) o.v[j] = c.use ? in.w.v[j] / c.norm : T(0);
    bk_st<T, W>(v + i, o);
  }
  __device__ void epilogue(const double*) const {}
};
