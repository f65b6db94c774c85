// bk_bicgstab_persist.cuh — BiCGStab for launch-latency-bound (L2-resident) systems as ONE persistent kernel (a
// thread-block cluster with its hardware barrier when the grid fits 16 CTAs, else a cooperative grid): four barriers per
// iteration replace five kernel launches (26 us per iteration as a graph at n = 10^4).  Same recurrences, operation
// order, roundings (separate mul / add) and breakdown tests as bk_bicg_enqueue_iter (reference _bicgstab_solve :859-964):
//   phase A  p' = r + beta (p - omega q) formed on the fly in the gather of q' = A p' (the owner of a row stores p'[row];
//            p and q are double-buffered), partial rhat.q'                        -> barrier -> alpha, |alpha| < eps: -11
//   phase B  s = r - alpha q', partial s.s                                        -> barrier -> exit_early = s.s < atol2
//   phase C  t = A s, partials t.s, t.t (skipped when exit_early)                 -> barrier -> omega, |omega| < eps: -11
//   phase D  x += alpha p' (+ omega s), r = s (- omega t), partials r.r, rhat.r   -> barrier -> k += 1, stop tests, beta
// Every CTA adds the per-CTA partials in the same fixed order, so all threads take identical decisions.
#pragma once

#include "bk_internal.cuh"
#include "bk_persist.cuh"

struct bk_bp_args {
  const int* rowptr;
  const int* col;
  const void* val;
  long long n;
  void* x;
  void* r;
  const void* rhat;
  void* p0;  // current p at entry
  void* p1;
  void* q0;  // current q at entry
  void* q1;
  void* s;
  void* t;
  bk_dev_state* st;
  double* partials;  // 6 rows x BK_MAXB
};

template <typename T, bool CLUSTER>
__global__ void __launch_bounds__(BK_GP_BLOCK, 1) bk_bicgstab_persistent_kernel(const bk_bp_args a) {
  bk_gp_barrier<CLUSTER> grid;
  __shared__ double s_red[BK_GP_NV * BK_GP_WARPS];
  __shared__ double s_sum[2];
  bk_dev_state* st = a.st;
  if (st->done) return;  // uniform: set by the set-up
  const double eps = sizeof(T) == 8 ? 2.220446049250313e-16 : 1.1920928955078125e-07;
  const long long n = a.n;
  const long long row = (long long)blockIdx.x * BK_GP_BLOCK + threadIdx.x;  // one row per thread
  const bool active = row < n;
  const int* __restrict__ col = a.col;
  const T* __restrict__ val = static_cast<const T*>(a.val);
  T* x = static_cast<T*>(a.x);
  T* r = static_cast<T*>(a.r);
  const T* __restrict__ rhat = static_cast<const T*>(a.rhat);
  T* pb[2] = {static_cast<T*>(a.p0), static_cast<T*>(a.p1)};
  T* qb[2] = {static_cast<T*>(a.q0), static_cast<T*>(a.q1)};
  T* s = static_cast<T*>(a.s);
  T* t = static_cast<T*>(a.t);
  double* PA = a.partials;
  double* PB = a.partials + (size_t)1 * BK_MAXB;
  double* PC = a.partials + (size_t)2 * BK_MAXB;  // 2 rows
  double* PD = a.partials + (size_t)4 * BK_MAXB;  // 2 rows
  const int rs_ = active ? a.rowptr[row] : 0, re_ = active ? a.rowptr[row + 1] : 0;
  const T rh = active ? rhat[row] : T(0);

  const double atol2 = st->atol2;
  const long long maxiter = st->maxiter;
  double rho = st->rho, rho_new = st->rho_new, alpha = st->alpha, omega = st->omega, beta = st->beta, rs = st->rs;
  long long k = 0;
  int status = BK_ST_MAXITER, early = 0;
  int cur = 0;
  for (;;) {
    const T* pc = pb[cur];
    const T* qc = qb[cur];
    T* pn = pb[cur ^ 1];
    T* qn = qb[cur ^ 1];
    const T tb = (T)beta, tw = (T)omega;
    // ---- phase A: q' = A p' with p' = r + beta (p - omega q) on the fly (:906-909) -------------------------------
    T qrow = T(0), prow = T(0);
    if (active) {
      T sum = T(0);
      for (int e = rs_; e < re_; ++e) {
        const int c = col[e];
        const T pv = bk_add(r[c], bk_mul(tb, bk_sub(pc[c], bk_mul(tw, qc[c]))));
        sum = fma(val[e], pv, sum);
      }
      prow = bk_add(r[row], bk_mul(tb, bk_sub(pc[row], bk_mul(tw, qc[row]))));
      pn[row] = prow;
      qn[row] = sum;
      qrow = sum;
    }
    {
      double acc[1] = {(double)rh * (double)qrow};
      bk_gp_block_sums<1>(acc, 1, s_red, PA, 0);
    }
    grid.sync();
    bk_gp_gather_sums(PA, 0, 1, s_sum);
    alpha = rho_new / s_sum[0];  // :910-911
    if (fabs(alpha) < eps) {     // :913-915
      status = BK_ST_BREAKDOWN_AW;
      break;
    }
    const T ta = (T)alpha;
    // ---- phase B: s = r - alpha q', s.s (:917-920) -----------------------------------------------------------------
    T srow = T(0);
    if (active) {
      srow = bk_sub(r[row], bk_mul(ta, qrow));
      s[row] = srow;
    }
    {
      double acc[1] = {(double)srow * (double)srow};
      bk_gp_block_sums<1>(acc, 1, s_red, PB, 0);
    }
    grid.sync();
    bk_gp_gather_sums(PB, 0, 1, s_sum);
    early = (s_sum[0] < atol2) ? 1 : 0;
    T trow = T(0);
    if (!early) {
      // ---- phase C: t = A s, t.s, t.t (:922-930) -------------------------------------------------------------------
      if (active) {
        T sum = T(0);
        for (int e = rs_; e < re_; ++e) sum = fma(val[e], s[col[e]], sum);
        t[row] = sum;
        trow = sum;
      }
      {
        double acc[BK_GP_NV];
#pragma unroll
        for (int v = 0; v < BK_GP_NV; ++v) acc[v] = 0.0;
        acc[0] = (double)srow * (double)trow;
        acc[1] = (double)trow * (double)trow;
        bk_gp_block_sums<BK_GP_NV>(acc, 2, s_red, PC, 0);
      }
      grid.sync();
      bk_gp_gather_sums(PC, 0, 2, s_sum);
      omega = (fabs(s_sum[1]) < eps) ? 0.0 : s_sum[0] / s_sum[1];
      if (fabs(omega) < eps) {  // :934-936 (not exit_early here)
        status = BK_ST_BREAKDOWN_AW;
        break;
      }
    }
    // ---- phase D: x, r update; r.r, rhat.r (:942-950) --------------------------------------------------------------
    const T to = (T)omega;
    T rrow = T(0);
    if (active) {
      const T ap = bk_mul(ta, prow);
      if (early) {
        x[row] = bk_add(x[row], ap);
        rrow = srow;
      } else {
        x[row] = bk_add(x[row], bk_add(ap, bk_mul(to, srow)));
        rrow = bk_sub(srow, bk_mul(to, trow));
      }
      r[row] = rrow;
    }
    {
      double acc[BK_GP_NV];
#pragma unroll
      for (int v = 0; v < BK_GP_NV; ++v) acc[v] = 0.0;
      acc[0] = (double)rrow * (double)rrow;
      acc[1] = (double)rh * (double)rrow;
      bk_gp_block_sums<BK_GP_NV>(acc, 2, s_red, PD, 0);
    }
    grid.sync();
    bk_gp_gather_sums(PD, 0, 2, s_sum);
    // ---- end of iteration k and top-of-loop tests of iteration k+1 (bk_op_bicg_xr::epilogue) ------------------------
    k += 1;
    const double rho_prev = rho_new;
    rho = rho_prev;
    rs = s_sum[0];
    if (early) {
      status = BK_ST_CONVERGED;
      break;
    }
    if (k >= maxiter) {
      status = BK_ST_MAXITER;
      break;
    }
    if (rs <= atol2) {
      status = BK_ST_CONVERGED;
      break;
    }
    rho_new = s_sum[1];
    if (fabs(rho_new) < eps * fabs(rho_prev)) {
      status = BK_ST_BREAKDOWN_RHO;
      break;
    }
    beta = rho_new / rho_prev * alpha / omega;
    cur ^= 1;
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->k = k;
    st->rs = rs;
    st->rho = rho;
    st->rho_new = rho_new;
    st->alpha = alpha;
    st->omega = omega;
    st->beta = beta;
    st->exit_early = early;
    st->status = status;
    st->done = 1;
  }
}
