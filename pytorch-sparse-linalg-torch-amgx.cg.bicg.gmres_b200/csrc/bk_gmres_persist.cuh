// bk_gmres_persist.cuh — restarted GMRES for launch-latency-bound (L2-resident) systems as ONE cooperative persistent
// kernel: the whole solve — every restart cycle, every Arnoldi step, the Givens least-squares update, the restart and
// stop tests — runs inside it; grid-wide barriers replace kernel launches.
//
// Why: BASELINE configs[3] (GMRES(30) on the LDC 100 x 100 pressure system, n = 10^4, inside a 1000-step time loop,
// reference FVM_example/LDC_by_torchsp/ldc_solver_common.py:232-236) costs 24 us per Arnoldi step as four graph-launched
// kernels — pure launch latency; everything it touches sits in L2.  Here an Arnoldi step is TWO grid barriers:
//   phase A  (one row per thread)  w = A v_j with v_j = w_prev / ||w_prev|| formed on the fly in the gather (the owner of
//            a row stores v_j[row]; same IEEE division as the stand-alone normalisation), then — row-local, so no
//            barrier in between — the partials of ||w||^2 and of h_i = v_i . w for i <= j
//   barrier  every CTA adds the per-CTA partials in the same fixed order (=> identical h, norms and decisions everywhere)
//   phase B  w -= sum_i h_i v_i (same FMA order as bk_multiaxpy_kernel), partial ||w||^2
//   barrier  thread 0 of EVERY CTA replays the small dense step (stored rotations, new Givens rotation, rotated rhs,
//            residual estimate, cycle flags — the code of bk_gm_dense_step on shared-memory copies); CTA 0 also keeps R.
// End of a cycle: CTA 0 back-substitutes R y = g, barrier, x += V y, barrier, r = b - A x and ||r||^2, barrier, restart
// test — all on the device; the host sees one kernel.  Same recurrences, thresholds and operation order as bk_gmres_t
// (reference gmres :641-784, _kth_arnoldi_iteration :331-388, _gmres_incremental :557-638); sums are reduced over a
// different partition than the multi-kernel path, so results agree to rounding, deterministically.
#pragma once

#include "bk_internal.cuh"
#include "bk_persist.cuh"

struct bk_gp_args {
  const int* rowptr;
  const int* col;
  const void* val;
  const void* diag;  // left Jacobi preconditioner (divide every A-product by diag) or nullptr
  long long n;
  void* V;           // (m + 1) x ldv; slot m doubles as the second w buffer
  size_t ldv;
  void* w;           // w buffer 0 (holds the unnormalised r_0 / M r_0 of the set-up when the kernel starts)
  void* x;
  const void* b;
  bk_dev_state* st;
  double* R;         // (m+1) x m column-major (CTA 0 only)
  double* y;         // m
  double* partials;  // [(m + 3)][BK_MAXB]
  int m;
};

template <typename T, bool CLUSTER>
__global__ void __launch_bounds__(BK_GP_BLOCK, 1) bk_gmres_persistent_kernel(const bk_gp_args a) {
  bk_gp_barrier<CLUSTER> grid;
  __shared__ double s_red[BK_GP_NV * BK_GP_WARPS];
  __shared__ double s_h[BK_GM_MAXM + 2];   // [0..j] projection coefficients (then the rotated column), scratch sums
  __shared__ double s_cs[BK_GM_MAXM];
  __shared__ double s_sn[BK_GM_MAXM];
  __shared__ double s_g[BK_GM_MAXM + 1];
  __shared__ double s_sum[2];
  __shared__ int s_flags[4];                // [0] cycle_over  [1] use  [2] done  [3] kcur
  __shared__ double s_scale;
  bk_dev_state* st = a.st;
  if (st->done) return;  // uniform: set by the set-up (zero rhs, maxiter 0, converged x0)
  const double eps = sizeof(T) == 8 ? 2.220446049250313e-16 : 1.1920928955078125e-07;
  const int m = a.m;
  const long long n = a.n;
  const long long row = (long long)blockIdx.x * BK_GP_BLOCK + threadIdx.x;  // one row per thread
  const bool active = row < n;
  const int* __restrict__ rowptr = a.rowptr;
  const int* __restrict__ col = a.col;
  const T* __restrict__ val = static_cast<const T*>(a.val);
  const T* __restrict__ dg = static_cast<const T*>(a.diag);
  T* V = static_cast<T*>(a.V);
  T* x = static_cast<T*>(a.x);
  const T* __restrict__ b = static_cast<const T*>(a.b);
  T* wbuf[2] = {static_cast<T*>(a.w), V + (size_t)m * a.ldv};
  double* P = a.partials;                       // rows 0 .. m+1: ||w||^2 then h_0..h_m
  double* P2 = a.partials + (size_t)(m + 2) * BK_MAXB;  // row m+2: ||w||^2 after the projection / ||r||^2
  const int rs = active ? rowptr[row] : 0, re = active ? rowptr[row + 1] : 0;
  const T dinv_row = (dg != nullptr && active) ? dg[row] : T(1);

  // replicated scalar state
  const double atol = st->g_atol, ptol = st->g_ptol;
  const long long maxiter = st->maxiter;
  const int incremental = st->g_method == BK_GMRES_INCREMENTAL;
  long long k = st->k, matvecs = st->matvecs;
  double resnorm = st->g_resnorm, scale = st->g_scale, rr_last = st->rtrue2;
  int use = st->g_use;
  int status = BK_ST_MAXITER;
  int cur = 0;             // index of the w buffer holding the latest UNnormalised vector
  bool pending = false;    // v_j must still be formed as wbuf[cur] / scale (false at entry: the set-up stored V[0])

  for (;;) {  // restart cycles
    if (threadIdx.x <= m) s_g[threadIdx.x] = (threadIdx.x == 0) ? resnorm : 0.0;
    __syncthreads();
    int kcur = 0;
    bool cycle_over = incremental && !(resnorm > ptol);  // 'incremental' enters the Arnoldi loop only while err > ptol
    while (!cycle_over) {
      const int j = kcur;
      T* vj = V + (size_t)j * a.ldv;
      const T* wprev = wbuf[cur];
      T* wout = wbuf[cur ^ 1];
      const T sc = (T)scale;
      // ---- phase A: w = [M] A v_j (v_j formed on the fly when pending), ||w||^2, h_i = v_i . w ---------------------
      T wrow = T(0);
      if (active) {
        T sum = T(0);
        if (pending) {
          for (int e = rs; e < re; ++e) {
            const int c = col[e];
            const T vc = use ? wprev[c] / sc : T(0);
            sum = fma(val[e], vc, sum);
          }
          vj[row] = use ? wprev[row] / sc : T(0);
        } else {
          for (int e = rs; e < re; ++e) sum = fma(val[e], vj[col[e]], sum);
        }
        if (dg != nullptr) sum = sum / dinv_row;
        wout[row] = sum;
        wrow = sum;
      }
      {  // values 0 .. j+1 of this step: ||w||^2, h_0 .. h_j — reduced 8 at a time
        double acc[BK_GP_NV];
        for (int g0 = 0; g0 < j + 2; g0 += BK_GP_NV) {
          const int nv = (j + 2 - g0 < BK_GP_NV) ? (j + 2 - g0) : BK_GP_NV;
#pragma unroll
          for (int v = 0; v < BK_GP_NV; ++v) {
            const int idx = g0 + v;  // 0: ||w||^2, 1 + i: h_i
            double t = 0.0;
            if (v < nv && active) t = (idx == 0) ? (double)wrow : (double)V[(size_t)(idx - 1) * a.ldv + row];
            acc[v] = t * (double)wrow;
          }
          bk_gp_block_sums<BK_GP_NV>(acc, nv, s_red, P, g0);
        }
      }
      grid.sync();
      // ---- phase B: every CTA adds the partials in the same order; w -= sum h_i v_i; ||w||^2 ------------------------
      bk_gp_gather_sums(P, 0, j + 2, s_h);  // s_h[0] = ||w||^2, s_h[1 + i] = h_i
      const double vn0 = sqrt(fmax(s_h[0], 0.0));
      const double vnorm0 = (vn0 > eps) ? vn0 : 0.0;
      matvecs += 1;
      T wnew = T(0);
      if (active) {
        T acc = T(0);
        for (int i = 0; i <= j; ++i) acc = fma((T)s_h[1 + i], V[(size_t)i * a.ldv + row], acc);
        wnew = bk_sub(wrow, acc);
        wout[row] = wnew;
      }
      {
        double acc[1] = {(double)wnew * (double)wnew};
        bk_gp_block_sums<1>(acc, 1, s_red, P2, 0);
      }
      grid.sync();
      // ---- the small dense step, replayed by thread 0 of every CTA (bk_gm_dense_step on shared copies) ---------------
      bk_gp_gather_sums(P2, 0, 1, s_sum);
      if (threadIdx.x == 0) {
        double* hc = s_h + 1;  // h_0 .. h_j
        const double norm1 = sqrt(fmax(s_sum[0], 0.0));
        const double thresh = eps * vnorm0;   // tol = eps * v_norm_0   (:358)
        const int use1 = norm1 > thresh;
        const double vnorm1 = use1 ? norm1 : 0.0;
        const bool breakdown = (vnorm1 == 0.0);  // :387
        hc[j + 1] = vnorm1;
        for (int i = 0; i < j; ++i) {  // stored rotations (:599-603)
          const double t = s_cs[i] * hc[i] - s_sn[i] * hc[i + 1];
          hc[i + 1] = s_sn[i] * hc[i] + s_cs[i] * hc[i + 1];
          hc[i] = t;
        }
        double c_new, s_new;
        bk_givens(hc[j], hc[j + 1], c_new, s_new);  // :606
        s_cs[j] = c_new;
        s_sn[j] = s_new;
        hc[j] = c_new * hc[j] - s_new * hc[j + 1];  // :611
        hc[j + 1] = 0.0;
        if (blockIdx.x == 0) {
          double* Rcol = a.R + (size_t)j * (m + 1);
          for (int i = 0; i <= j; ++i) Rcol[i] = hc[i];  // :615
        }
        const double gj = s_g[j], gj1 = s_g[j + 1];
        s_g[j] = c_new * gj - s_new * gj1;  // :618-620
        const double gnext = s_new * gj + c_new * gj1;
        s_g[j + 1] = gnext;
        const double err = fabs(gnext);
        bool over = breakdown || (j + 1 >= m);
        if (incremental && !(err > ptol)) over = true;  // :591
        s_flags[0] = over ? 1 : 0;
        s_flags[1] = use1;
        s_flags[3] = j + 1;
        s_scale = norm1;
      }
      __syncthreads();
      cycle_over = s_flags[0] != 0;
      use = s_flags[1];
      kcur = s_flags[3];
      scale = s_scale;
      cur ^= 1;
      pending = true;
      __syncthreads();
    }
    // ---- end of the cycle: y = R^-1 g (CTA 0, one warp), x += V[:, :kcur] y, r = b - A x ------------------------------
    if (blockIdx.x == 0 && threadIdx.x < 32) {
      const int lane = threadIdx.x;
      const int ld = m + 1;
      for (int i = kcur - 1; i >= 0; --i) {
        double acc = 0.0;
        for (int c = i + 1 + lane; c < kcur; c += 32) acc += a.R[(size_t)c * ld + i] * a.y[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) a.y[i] = (s_g[i] - acc) / a.R[(size_t)i * ld + i];
        __syncwarp();
      }
      __threadfence();
    }
    grid.sync();
    if (active && kcur > 0) {
      T acc = T(0);
      for (int i = 0; i < kcur; ++i) acc = fma((T)__ldcg(&a.y[i]), V[(size_t)i * a.ldv + row], acc);
      x[row] = bk_add(x[row], acc);
    }
    grid.sync();
    T* rbuf = wbuf[cur ^ 1];
    T rrow = T(0);
    if (active) {
      T sum = T(0);
      for (int e = rs; e < re; ++e) sum = fma(val[e], x[col[e]], sum);
      rrow = bk_sub(b[row], sum);
      if (dg != nullptr) rrow = rrow / dinv_row;
      rbuf[row] = rrow;
    }
    {
      double acc[1] = {(double)rrow * (double)rrow};
      bk_gp_block_sums<1>(acc, 1, s_red, P2, 0);
    }
    grid.sync();
    bk_gp_gather_sums(P2, 0, 1, s_sum);
    {  // bk_gm_after_residual, replayed everywhere
      const double rr = s_sum[0];
      matvecs += 1;
      rr_last = rr;
      const double norm = sqrt(fmax(rr, 0.0));
      use = norm > eps;
      resnorm = use ? norm : 0.0;
      scale = norm;
      k += 1;
      cur ^= 1;
      pending = true;
      const bool go = (k < maxiter) && (resnorm > atol);
      if (!go) {
        status = (resnorm > atol) ? BK_ST_MAXITER : BK_ST_CONVERGED;
        break;
      }
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->k = k;
    st->matvecs = matvecs;
    st->rtrue2 = rr_last;
    st->g_resnorm = resnorm;
    st->g_scale = scale;
    st->g_use = use;
    st->status = status;
    st->done = 1;
  }
}
