// rollout_kernel.cu -- the hot path of librbo.so: one CTA per (quasi-)Monte-Carlo trajectory, persistent grid
// with a dynamic trajectory counter, FP64 throughout, written for sm_100a.
//
// What one CTA does for one sample index m (reference: rollout.jl:279-340 -> rollout! :39-74, resolve :108-111,
// gradient :233-277):
//   * keeps the trajectory's state in shared memory: the (<= 8) fantasy rows of the Cholesky factor as one
//     8-row panel, the coefficient tape cs[0..h+1], u = L^-1 y, fantasy locations / draws;
//   * every surrogate evaluation is expressed as column operations on a shared-memory matrix V (rows =
//     observations, columns = right-hand sides): build kernel columns, triangular solves against
//     [L0 (global, pre-packed 8-row panels with inverted 8x8 diagonal blocks, read through L1/L2) ; fantasy panel],
//     then row reductions (dot products of column pairs; lanes walk the rows, warps own blocks of pairs);
//   * the multi-start inner solve (replacing rbf_optim.jl:68-101 / Optim.IPNewton) keeps W start slots busy in
//     lock-step rounds: each round evaluates (alpha, grad alpha, Hess alpha) at one trial point per active slot,
//     one warp per slot then runs the regularised projected Newton logic, finished slots are refilled from the
//     start queue;
//   * the adjoint (rollout.jl:233-277) replays the tape: each policy solve i = t..1 is re-evaluated ONCE and its
//     perturbation columns (rbs.jl:633-764) are pushed into the right-hand sides of the earlier duals.
#include "rbo_kernel.cuh"

namespace rbo {

namespace {

// int-area layout
enum { I_M = 0, I_NACT = 1, I_TSTATUS = 2, I_BEST = 3, I_EVALS = 4, I_T = 5, I_CASE = 6, I_NEXT = 7, I_ARR = 64 };
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, off));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, off));
  return v;
}
// index of (p, q), p <= q, in the row-major upper triangle of an n x n matrix
__device__ __forceinline__ int tri_idx(int p, int q, int n) { return p * n - (p * (p - 1)) / 2 + (q - p); }

struct K {
  const DevProblem& P;
  SmemPlan pl;
  double* sm;
  int* si;
  int tid, lane, warp;
  int nf;  // number of fantasy rows that are active for the current operation (uniform over the CTA)
  int CCOL, UCOL;  // columns of V that hold the current coefficients c and u = L^-1 y
  double *V, *Fp, *G, *cs, *u, *Xf, *yf, *gyf, *misc, *adj, *bestx;
  int *alist, *phase, *sstat, *siter, *stry, *sstart, *sevals, *sfr, *colidx, *pairs, *tblq, *tbld;

  __device__ K(const DevProblem& P_, double* sm_) : P(P_), sm(sm_) {
    pl = make_plan(P.d, P.N8, P.h, P.W, P.RP, P.NR, P.RSmax, P.NPmax);
    tid = threadIdx.x; lane = tid & 31; warp = tid >> 5;
    nf = 0;
    CCOL = P.RP - 1; UCOL = P.RP - 2;
    V = sm + pl.V; Fp = sm + pl.Fp; G = sm + pl.G; cs = sm + pl.cs; u = sm + pl.u;
    Xf = sm + pl.Xf; yf = sm + pl.yf; gyf = sm + pl.gyf; misc = sm + pl.misc; adj = sm + pl.adj; bestx = sm + pl.bestx;
    si = reinterpret_cast<int*>(sm + pl.ints);
    pairs = reinterpret_cast<int*>(sm + pl.pairs);
    tblq = reinterpret_cast<int*>(sm + pl.tbl);
    const int W = P.W, q1 = P.d + 1;
    tbld = tblq + q1 * (q1 + 1) / 2;
    alist = si + I_ARR; phase = alist + W; sstat = phase + W; siter = sstat + W; stry = siter + W;
    sstart = stry + W; sevals = sstart + W; sfr = sevals + W; colidx = sfr + 32 * W;
  }

  __device__ __forceinline__ double xcoord(int j, int p) const {
    return j < P.N8 ? __ldg(P.Xb + (size_t)p * P.N8 + j) : Xf[(j - P.N8) * P.d + p];
  }
  __device__ __forceinline__ bool row_active(int j) const { return j < P.N || (j >= P.N8 && j < P.N8 + nf); }
  __device__ __forceinline__ int nact_rows() const { return P.N + nf; }
  __device__ __forceinline__ int act_row(int a) const { return a < P.N ? a : P.N8 + (a - P.N); }

  // (p, q) tables of the upper triangles of size d+1 (tblq) and d (tbld), packed p | q << 8
  __device__ void build_tables() {
    const int d = P.d, q1 = d + 1;
    for (int e = tid; e < q1 * (q1 + 1) / 2 + d * (d + 1) / 2; e += RBO_THREADS) {
      int n = q1, t = e;
      if (e >= q1 * (q1 + 1) / 2) { n = d; t = e - q1 * (q1 + 1) / 2; }
      int p = 0;
      while (t >= n - p) { t -= n - p; ++p; }
      tblq[e] = p | ((p + t) << 8);
    }
  }

  // column `col` of V <- vector v (length NR)
  __device__ void set_column(int col, const double* v) {
    for (int j = tid; j < P.NR; j += RBO_THREADS) V[(size_t)j * P.RP + col] = v[j];
  }

  // ------------------------------------------------------------------------------------------------
  // Kernel columns for `np` points: column block at cb(s) gets [kx | b*r (d columns) | a | b] for every row
  // (rbf.jl:180-208 eval_KxX / eval_gradKxX and the eval_Hk coefficients of rbf.jl:141-150). Inactive rows get 0.
  // ------------------------------------------------------------------------------------------------
  template <class PT, class CB>
  __device__ void fill_columns(int np, PT pt, CB cb) {
    const int d = P.d, NR = P.NR, RP = P.RP;
    for (int idx = tid; idx < np * NR; idx += RBO_THREADS) {
      int s = idx / NR, j = idx - s * NR;
      double* row = V + (size_t)j * RP + cb(s);
      if (!row_active(j)) {
        for (int q = 0; q < d + 3; ++q) row[q] = 0.0;
        continue;
      }
      const double* x = pt(s);
      double rho2 = 0.0;
      for (int p = 0; p < d; ++p) { double r = x[p] - xcoord(j, p); rho2 = fma(r, r, rho2); }
      double psi, a, b, gb;
      kern_radial(P.kern, rho2, psi, a, b, gb);
      row[0] = psi;
      for (int p = 0; p < d; ++p) row[1 + p] = b * (x[p] - xcoord(j, p));
      row[d + 1] = a;
      row[d + 2] = b;
    }
  }

  // ------------------------------------------------------------------------------------------------
  // Row reductions.  out[rs * npairs + i] = sum over the rows of split rs of V[j][c1_i] * V[j][c2_i], pairs[i] =
  // c1 | c2 << 16.  A warp owns a block of 72 pairs and one row split; lane = (row group rg = lane & 3, pair group
  // eg = lane >> 2): 9 accumulators per lane, rows a = rg + 4 * (rs + RS * i). The RS partial sums are added in a
  // fixed order by the consumer, so results do not depend on scheduling.
  // ------------------------------------------------------------------------------------------------
  __device__ int choose_rs(int nblocks) const {
    int rs = RBO_NWARPS / (nblocks > 0 ? nblocks : 1);
    return rs < 1 ? 1 : (rs > P.RSmax ? P.RSmax : rs);
  }

  __device__ void reduce_pairs(int npairs, double* out, int RS) {
    const int RP = P.RP, na = nact_rows(), rg = lane & 3, eg = lane >> 2;
    const int nblk = (npairs + 71) / 72;
    for (int task = warp; task < nblk * RS; task += RBO_NWARPS) {
      const int blk = task / RS, rs = task - blk * RS;
      int c1[9], c2[9];
      double acc[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        int e = blk * 72 + eg * 9 + i;
        int pr = e < npairs ? pairs[e] : 0;
        c1[i] = pr & 0xffff; c2[i] = pr >> 16;
        acc[i] = 0.0;
      }
      for (int a = rg + 4 * rs; a < na; a += 4 * RS) {
        const double* row = V + (size_t)act_row(a) * RP;
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[i] = fma(row[c1[i]], row[c2[i]], acc[i]);
      }
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        acc[i] += __shfl_xor_sync(FULL, acc[i], 1);
        acc[i] += __shfl_xor_sync(FULL, acc[i], 2);
      }
      if (rg == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          int e = blk * 72 + eg * 9 + i;
          if (e < npairs) out[(size_t)rs * npairs + e] = acc[i];
        }
      }
    }
  }

  // HC = sum_j c_j Hk(x - X_j) (rbs.jl:516-523) and HW = sum_j w_j Hk(x - X_j) (rbs.jl:542-545) for np points.
  // Hk = a r r' + b I: entries (p <= q) accumulate a r_p r_q, the b-weighted sums go to the extra entry T2.
  // phess[((rs * np + s) * 2 + which) * (T2 + 1) + e], which = 0 (c-weighted) / 1 (w-weighted).
  // Columns: (a, b) at cab(s) + d+1, d+2 ; w at cw(s) ; c in CCOL.
  template <class PT, class CAB, class CW>
  __device__ void reduce_hess(int np, PT pt, CAB cab, CW cw, int RS) {
    const int d = P.d, RP = P.RP, T2 = d * (d + 1) / 2, na = nact_rows(), rg = lane & 3, eg = lane >> 2;
    const int nblk = (T2 + 55) / 56;
    double* out = sm + pl.phess;
    for (int task = warp; task < np * nblk * RS; task += RBO_NWARPS) {
      const int s = task / (nblk * RS), rem = task - s * nblk * RS, blk = rem / RS, rs = rem - blk * RS;
      const double* x = pt(s);
      const int colab = cab(s) + d + 1, colw = cw(s);
      int pp[7], qq[7];
      double xp[7], xq[7], accC[7], accW[7], bC = 0.0, bW = 0.0;
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        int e = blk * 56 + eg * 7 + i;
        int pq = e < T2 ? tbld[e] : 0;
        pp[i] = pq & 0xff; qq[i] = pq >> 8;
        xp[i] = x[pp[i]]; xq[i] = x[qq[i]];
        accC[i] = 0.0; accW[i] = 0.0;
      }
      for (int a = rg + 4 * rs; a < na; a += 4 * RS) {
        const int j = act_row(a);
        const double* row = V + (size_t)j * RP;
        const double aj = row[colab], bj = row[colab + 1], wj = row[colw], cj = row[CCOL];
        const double ca = cj * aj, wa = wj * aj;
        bC = fma(cj, bj, bC); bW = fma(wj, bj, bW);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          double rr = (xp[i] - xcoord(j, pp[i])) * (xq[i] - xcoord(j, qq[i]));
          accC[i] = fma(ca, rr, accC[i]);
          accW[i] = fma(wa, rr, accW[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        accC[i] += __shfl_xor_sync(FULL, accC[i], 1); accC[i] += __shfl_xor_sync(FULL, accC[i], 2);
        accW[i] += __shfl_xor_sync(FULL, accW[i], 1); accW[i] += __shfl_xor_sync(FULL, accW[i], 2);
      }
      bC += __shfl_xor_sync(FULL, bC, 1); bC += __shfl_xor_sync(FULL, bC, 2);
      bW += __shfl_xor_sync(FULL, bW, 1); bW += __shfl_xor_sync(FULL, bW, 2);
      if (rg == 0) {
        double* oC = out + ((size_t)(rs * np + s) * 2 + 0) * (T2 + 1);
        double* oW = out + ((size_t)(rs * np + s) * 2 + 1) * (T2 + 1);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          int e = blk * 56 + eg * 7 + i;
          if (e < T2) { oC[e] = accC[i]; oW[e] = accW[i]; }
        }
        if (blk == 0 && eg == 0) { oC[T2] = bC; oW[T2] = bW; }
      }
    }
  }

  // ------------------------------------------------------------------------------------------------
  // Triangular solves of `ncols` columns of V (indices in colidx[]) against L = [L0 0; F G]:
  // FWD: V <- L^-1 V, else V <- L^-T V.  A group of KS adjacent lanes owns one column and splits the
  // k-range of every 8-row panel; partial sums are combined with xor-shuffles, the 8x8 diagonal blocks are
  // pre-inverted so the diagonal step is a small mat-vec.  L0 panels come from global memory (read-only path),
  // the fantasy panel from shared memory.  `nfan` = number of active fantasy rows.
  // ------------------------------------------------------------------------------------------------
  template <bool GLOBAL>
  __device__ __forceinline__ void panel_rows(const double* pan, int nk, int part, int KS, const double* vcol, double acc[8]) const {
    const int RP = P.RP;
#pragma unroll 4
    for (int k = part; k < nk; k += KS) {
      double v = vcol[(size_t)k * RP];
      const double2* lp = reinterpret_cast<const double2*>(pan + (size_t)k * 8);
      double2 l0, l1, l2, l3;
      if (GLOBAL) { l0 = __ldg(lp); l1 = __ldg(lp + 1); l2 = __ldg(lp + 2); l3 = __ldg(lp + 3); }
      else { l0 = lp[0]; l1 = lp[1]; l2 = lp[2]; l3 = lp[3]; }
      acc[0] = fma(l0.x, v, acc[0]); acc[1] = fma(l0.y, v, acc[1]);
      acc[2] = fma(l1.x, v, acc[2]); acc[3] = fma(l1.y, v, acc[3]);
      acc[4] = fma(l2.x, v, acc[4]); acc[5] = fma(l2.y, v, acc[5]);
      acc[6] = fma(l3.x, v, acc[6]); acc[7] = fma(l3.y, v, acc[7]);
    }
  }

  __device__ __forceinline__ void group_reduce(double acc[8], int KS) const {
    for (int off = KS >> 1; off > 0; off >>= 1) {
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r] += shfl_xor_d(acc[r], off);
    }
  }

  // out rows r == part (mod KS): v_r = sum_kk D[kk*8 + r] * t[kk]; rows >= rmax are forced to 0.
  template <bool GLOBAL>
  __device__ __forceinline__ void diag_apply(const double* dg, const double t[8], int part, int KS, double* vout /* row 0 of the block */,
                                              bool valid, int rmax) const {
    const int RP = P.RP;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if ((r & (KS - 1)) == part) {
        double s = 0.0;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          double dv = GLOBAL ? __ldg(dg + kk * 8 + r) : dg[kk * 8 + r];
          s = fma(dv, t[kk], s);
        }
        if (r >= rmax) s = 0.0;
        if (valid && part < 8) vout[(size_t)r * RP] = s;
      }
    }
  }

  template <bool FWD>
  __device__ void tri_solve(int ncols, int nfan) {
    if (ncols <= 0) return;
    const int RP = P.RP, N8 = P.N8, nb8 = P.nb8;
    int KS = 1;
    while (KS < 32 && ncols * (KS * 2) <= RBO_THREADS) KS *= 2;
    const int ngroups = RBO_THREADS / KS, gid = tid / KS, part = tid & (KS - 1);
    const int warp_first_gid = (tid & ~31) / KS;
    for (int cb = 0; cb < ncols; cb += ngroups) {
      if (cb + warp_first_gid >= ncols) continue;  // whole warp idle (warp-uniform)
      const int ci = cb + gid;
      const bool valid = ci < ncols;
      double* vcol = V + colidx[valid ? ci : 0];
      double acc[8], t[8];
      if (FWD) {
        for (int ib = 0; ib < nb8; ++ib) {
          const double* pan = P.Lf + (size_t)32 * ib * (ib + 1);
          const int nk = 8 * ib;
#pragma unroll
          for (int r = 0; r < 8; ++r) acc[r] = 0.0;
          panel_rows<true>(pan, nk, part, KS, vcol, acc);
          group_reduce(acc, KS);
#pragma unroll
          for (int r = 0; r < 8; ++r) t[r] = vcol[(size_t)(nk + r) * RP] - acc[r];
          __syncwarp();
          diag_apply<true>(pan + (size_t)nk * 8, t, part, KS, vcol + (size_t)nk * RP, valid, 8);
          __syncwarp();
        }
        if (nfan > 0) {
#pragma unroll
          for (int r = 0; r < 8; ++r) acc[r] = 0.0;
          panel_rows<false>(Fp, N8, part, KS, vcol, acc);
          group_reduce(acc, KS);
#pragma unroll
          for (int r = 0; r < 8; ++r) t[r] = vcol[(size_t)(N8 + r) * RP] - acc[r];
          __syncwarp();
          diag_apply<false>(Fp + (size_t)N8 * 8, t, part, KS, vcol + (size_t)N8 * RP, valid, nfan);
          __syncwarp();
        }
      } else {
        if (nfan > 0) {
          // w_bot = Ginv^T t restricted to the active rows: w_r = sum_{kk} Ginv[kk][r] t[kk], Ginv[kk][r] = Fp[(N8 + r)*8 + kk]
#pragma unroll
          for (int r = 0; r < 8; ++r) t[r] = (r < nfan) ? vcol[(size_t)(N8 + r) * RP] : 0.0;
          __syncwarp();
          double wb[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            double s = 0.0;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) s = fma(Fp[(size_t)(N8 + r) * 8 + kk], t[kk], s);
            wb[r] = (r < nfan) ? s : 0.0;
          }
          if (valid && part == 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r) vcol[(size_t)(N8 + r) * RP] = wb[r];
          }
          // top rows: v_i -= sum_r F[r][i] w_bot[r]
          for (int i = part; i < N8; i += KS) {
            const double* f = Fp + (size_t)i * 8;
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) s = fma(f[r], wb[r], s);
            if (valid) vcol[(size_t)i * RP] -= s;
          }
          __syncwarp();
        }
        for (int ib = nb8 - 1; ib >= 0; --ib) {
          const double* pan = P.Lb + ((size_t)8 * N8 * ib - (size_t)32 * ib * (ib - 1));
          const int k0 = 8 * (ib + 1), nk = N8 - k0;
#pragma unroll
          for (int r = 0; r < 8; ++r) acc[r] = 0.0;
          panel_rows<true>(pan, nk, part, KS, vcol + (size_t)k0 * RP, acc);
          group_reduce(acc, KS);
#pragma unroll
          for (int r = 0; r < 8; ++r) t[r] = vcol[(size_t)(8 * ib + r) * RP] - acc[r];
          __syncwarp();
          diag_apply<true>(pan + (size_t)nk * 8, t, part, KS, vcol + (size_t)(8 * ib) * RP, valid, 8);
          __syncwarp();
        }
      }
    }
  }

  // ------------------------------------------------------------------------------------------------
  // One warp assembles the surrogate evaluation of slot `sl` from the partial sums (rbs.jl:513-577).
  //   aidx / np: position of the slot in the reduction outputs; RS*: row splits used by each reduction.
  // Writes sdmu, sdsig (grad sigma), sga (grad alpha), sHt (-(H alpha + mu-sigma cross term)), sHref (H alpha as the
  // reference computes it, Q1) and sgh = [alpha, g_mu, g_sig, g_muth, g_sigth, sigma, mu, finite].
  // ------------------------------------------------------------------------------------------------
  __device__ void assemble_warp(int sl, int aidx, int np, int npairs_pre, int npairs_post, int post_off, int RSpre, int RSpost, int RShess, double fstar) {
    const int d = P.d, dd = d * d, q1 = d + 1, T2 = d * (d + 1) / 2;
    const double* ppre = sm + pl.ppre; const double* ppost = sm + pl.ppost; const double* phess = sm + pl.phess;
    double* dmu = sm + pl.sdmu + sl * d; double* dsig = sm + pl.sdsig + sl * d; double* ga = sm + pl.sga + sl * d;
    double* Ht = sm + pl.sHt + sl * dd; double* Href = sm + pl.sHref + sl * dd; double* gh = sm + pl.sgh + sl * 8;
    auto pre = [&](int e) { double s = 0.0; for (int r = 0; r < RSpre; ++r) s += ppre[(size_t)r * npairs_pre + aidx * q1 + e]; return s; };
    auto post = [&](int e) { double s = 0.0; for (int r = 0; r < RSpost; ++r) s += ppost[(size_t)r * npairs_post + post_off + e]; return s; };
    auto hes = [&](int which, int e) { double s = 0.0; for (int r = 0; r < RShess; ++r) s += phess[((size_t)(r * np + aidx) * 2 + which) * (T2 + 1) + e]; return s; };
    const double mu = pre(0);
    const double var = P.k0 - post(0);  // rbs.jl:528 (kx.w == |L^-1 kx|^2)
    const double sigma = sqrt(var), isg = 1.0 / sigma;
    const GPart g = rule_eval(P.rule_id, P.sigma_tol, mu, sigma, P.theta1, fstar);
    bool fin = isfinite(g.g);
    for (int p = lane; p < d; p += 32) {
      double m = pre(1 + p), sg = -post(1 + p) * isg;  // rbs.jl:514, 529
      double a = g.g_mu * m + g.g_sig * sg;            // rbs.jl:567
      dmu[p] = m; dsig[p] = sg; ga[p] = a;
      fin = fin && isfinite(a);
    }
    __syncwarp();
    const double bC = hes(0, T2), bW = hes(1, T2);
    for (int e = lane; e < T2; e += 32) {
      const int pq = tbld[e], p = pq & 0xff, q = pq >> 8;
      double gram = post(tri_idx(p + 1, q + 1, q1)), hc = hes(0, e), hw = hes(1, e);
      if (p == q) { hc += bC; hw += bW; }
      double hs = (-dsig[p] * dsig[q] - gram - hw) * isg;                                                            // rbs.jl:541-546
      double href = g.g_mumu * dmu[p] * dmu[q] + g.g_mu * hc + g.g_sigsig * dsig[p] * dsig[q] + g.g_sig * hs;        // rbs.jl:568
      double htrue = href + g.g_musig * (dmu[p] * dsig[q] + dsig[p] * dmu[q]);
      Href[p * d + q] = href; Href[q * d + p] = href;
      Ht[p * d + q] = -htrue; Ht[q * d + p] = -htrue;
      fin = fin && isfinite(htrue);
    }
    fin = __all_sync(FULL, fin);
    if (lane == 0) {
      gh[0] = g.g; gh[1] = g.g_mu; gh[2] = g.g_sig; gh[3] = g.g_muth; gh[4] = g.g_sigth; gh[5] = sigma; gh[6] = mu;
      gh[7] = fin ? 1.0 : 0.0;
    }
    __syncwarp();
  }

  // Cholesky of the n x n row-major matrix A (n <= 32) by one warp; lane i owns row i. Returns false if not PD.
  __device__ bool chol_warp(double* A, int n) const {
    for (int j = 0; j < n; ++j) {
      double t = 0.0;
      if (lane >= j && lane < n) {
        t = A[lane * n + j];
        for (int k = 0; k < j; ++k) t = fma(-A[lane * n + k], A[j * n + k], t);
      }
      const double tj = __shfl_sync(FULL, t, j);
      if (!(tj > 0.0) || !isfinite(tj)) return false;
      const double ljj = sqrt(tj);
      if (lane == j) A[j * n + j] = ljj;
      else if (lane > j && lane < n) A[lane * n + j] = t / ljj;
      __syncwarp();
    }
    return true;
  }

  // ------------------------------------------------------------------------------------------------
  // One step of the per-start state machine (regularised projected Newton, specified in DESIGN.md section 4),
  // run by ONE WARP for slot `sl` right after assemble_warp(). Returns true (uniformly) if the slot has a new trial
  // point in sxt and stays active.
  // ------------------------------------------------------------------------------------------------
  __device__ bool slot_logic_warp(int sl) {
    const int d = P.d, dd = d * d;
    const rbo_solver_opts& o = P.so;
    double* x = sm + pl.sx + sl * d; double* xt = sm + pl.sxt + sl * d; double* g = sm + pl.sg + sl * d;
    double* H = sm + pl.sH + sl * dd; double* A = sm + pl.sA + sl * dd;
    const double* Ht = sm + pl.sHt + sl * dd; const double* ga = sm + pl.sga + sl * d; const double* gh = sm + pl.sgh + sl * 8;
    int* fr = sfr + 32 * sl;
    double f = (sm + pl.sf)[sl], lam = (sm + pl.slam)[sl], pred = (sm + pl.spred)[sl], hs_st = (sm + pl.shs)[sl];
    int iters = siter[sl], tries = stry[sl];
    const double ft = -gh[0];
    const bool fin = gh[7] != 0.0;
    const int ph = phase[sl];
    __syncwarp();
    auto store = [&](int status, bool keep) {
      if (lane == 0) {
        (sm + pl.sf)[sl] = f; (sm + pl.slam)[sl] = lam; (sm + pl.spred)[sl] = pred; (sm + pl.shs)[sl] = hs_st;
        siter[sl] = iters; stry[sl] = tries; sevals[sl] += 1;
        if (!keep) sstat[sl] = status;
        phase[sl] = keep ? 1 : 2;
      }
      __syncwarp();
      return keep;
    };
    auto accept_state = [&]() {
      for (int a = lane; a < d; a += 32) { x[a] = xt[a]; g[a] = -ga[a]; }
      for (int i = lane; i < dd; i += 32) H[i] = Ht[i];
      f = ft;
      __syncwarp();
    };
    bool fresh;
    if (ph == 0) {
      if (!fin) {
        for (int a = lane; a < d; a += 32) x[a] = xt[a];
        f = nan("");
        return store(RBO_SOLVE_NAN, false);
      }
      accept_state();
      lam = 0.0; iters = 0;
      fresh = true;
    } else {
      const double ared = f - ft;
      if (fin && ared >= o.eta * pred) {
        accept_state();
        if (ared >= 0.75 * pred) { lam *= o.lam_down; if (lam < o.lam_min * hs_st) lam = 0.0; }
        iters += 1;
        if (iters >= o.maxit) return store(RBO_SOLVE_MAXIT, false);
        fresh = true;
      } else {
        lam = fmax(o.lam_up * lam, o.lam_min * hs_st);
        tries += 1;
        fresh = false;
      }
    }
    // active set (lane a <-> coordinate a), projected gradient, Hessian scale
    const bool in = lane < d;
    const double xa = in ? x[lane] : 0.0, gg = in ? g[lane] : 0.0, haa = in ? H[lane * d + lane] : 0.0;
    const bool act = in && ((xa <= P.lbs[lane] && gg > 0.0) || (xa >= P.ubs[lane] && gg < 0.0));
    const unsigned fmask = __ballot_sync(FULL, in && !act);
    const int nfree = __popc(fmask);
    const bool isfree = in && !act;
    const double pg = warp_max(isfree ? fabs(gg) : 0.0);
    double hs = warp_max(isfree ? fabs(haa) : 0.0);
    const double mind = warp_min(isfree ? haa : INFINITY);
    if (!(hs > 0.0)) hs = 1.0;
    hs_st = hs;
    if (lane < nfree) fr[lane] = __fns(fmask, 0, lane + 1);
    __syncwarp();
    if (fresh) {
      if (pg <= o.gtol * fmax(1.0, fabs(f))) return store(RBO_SOLVE_CONVERGED, false);
      tries = 0;
    }
    const int myc = lane < nfree ? fr[lane] : 0;  // coordinate owned by this lane in the reduced system
    while (tries < o.maxtry) {
      if (mind + lam <= 0.0) lam = fmax(lam, -mind + o.lam_min * hs);
      for (int e = lane; e < nfree * nfree; e += 32) {
        int i = e / nfree, j = e - i * nfree;
        A[e] = H[fr[i] * d + fr[j]] + (i == j ? lam : 0.0);
      }
      __syncwarp();
      if (!chol_warp(A, nfree)) { lam = fmax(o.lam_up * lam, o.lam_min * hs); tries += 1; __syncwarp(); continue; }
      // solve (H_FF + lam I) p = -g_F : forward then backward substitution, lane i holds component i
      double t = lane < nfree ? -g[myc] : 0.0;
      for (int i = 0; i < nfree; ++i) {
        const double pi = __shfl_sync(FULL, t, i) / A[i * nfree + i];
        if (lane == i) t = pi;
        else if (lane > i && lane < nfree) t = fma(-A[lane * nfree + i], pi, t);
      }
      for (int i = nfree - 1; i >= 0; --i) {
        const double pi = __shfl_sync(FULL, t, i) / A[i * nfree + i];
        if (lane == i) t = pi;
        else if (lane < i) t = fma(-A[i * nfree + lane], pi, t);
      }
      for (int a = lane; a < d; a += 32) xt[a] = x[a];
      __syncwarp();
      if (lane < nfree) xt[myc] = fmin(fmax(x[myc] + t, P.lbs[myc]), P.ubs[myc]);
      __syncwarp();
      const double sa = in ? xt[lane] - xa : 0.0;
      const double smax = warp_max(fabs(sa)), xmax = warp_max(fabs(xa));
      if (smax <= o.xtol * fmax(1.0, xmax)) return store(RBO_SOLVE_STEP_TINY, false);
      double hsv = 0.0;
      if (in) for (int b = 0; b < d; ++b) hsv = fma(H[lane * d + b], xt[b] - x[b], hsv);
      const double gs = warp_sum(gg * sa), sHs = warp_sum(sa * hsv);
      const double pr = -(gs + 0.5 * sHs);
      if (!(pr > 0.0)) { lam = fmax(o.lam_up * lam, o.lam_min * hs); tries += 1; continue; }
      if (pr <= o.pred_tol * fmax(1.0, fabs(f))) return store(RBO_SOLVE_PRED_TINY, false);
      pred = pr;
      return store(0, true);
    }
    return store(RBO_SOLVE_STALLED, false);
  }

  // loads start `sid` into slot `sl`
  __device__ void load_start(int sl, int sid) {
    const int d = P.d;
    phase[sl] = 0; sstat[sl] = RBO_SOLVE_MAXIT; siter[sl] = 0; stry[sl] = 0; sevals[sl] = 0; sstart[sl] = sid;
    for (int a = 0; a < d; ++a) {
      double v = __ldg(P.starts + (size_t)sid * d + a);
      (sm + pl.sxt)[sl * d + a] = fmin(fmax(v, P.lbs[a]), P.ubs[a]);
    }
  }

  // ------------------------------------------------------------------------------------------------
  // multistart_base_solve! (rbf_optim.jl:68-101 / :103-134): all S starts through W slots in lock-step rounds.
  // The coefficients of the active surrogate must be in column CCOL. Result: bestx (argmax), misc[0] = -alpha
  // there, si[I_BEST] (or -1), si[I_EVALS]. Ties resolve to the lowest start index (findmin: first minimum).
  // ------------------------------------------------------------------------------------------------
  __device__ void multistart(size_t tape_off) {
    const int d = P.d, W = P.W, q1 = d + 1, T = q1 * (q1 + 1) / 2;
    __syncthreads();
    if (tid == 0) {
      si[I_BEST] = -1; si[I_EVALS] = 0; misc[0] = 0.0;
      const int n0 = min(W, P.S);
      for (int i = 0; i < n0; ++i) { alist[i] = i; load_start(i, i); }
      si[I_NACT] = n0; si[I_NEXT] = n0;
    }
    __syncthreads();
    int nact = si[I_NACT];
    while (nact > 0) {
      auto pt = [&](int s) { return (const double*)(sm + pl.sxt + alist[s] * d); };
      auto cb = [&](int s) { return alist[s] * P.CS; };
      fill_columns(nact, pt, cb);
      for (int i = tid; i < nact * q1; i += RBO_THREADS) {
        int s = i / q1, q = i - s * q1, col = alist[s] * P.CS + q;
        colidx[i] = col;
        pairs[i] = col | (CCOL << 16);  // mu = kx.c, grad mu = grad_kx c (rbs.jl:513-514)
      }
      __syncthreads();
      const int RSpre = choose_rs((nact * q1 + 71) / 72);
      reduce_pairs(nact * q1, sm + pl.ppre, RSpre);
      __syncthreads();  // the solve below overwrites the raw columns in place
      tri_solve<true>(nact * q1, nf);
      for (int i = tid; i < nact * T; i += RBO_THREADS) {
        int s = i / T, e = i - s * T, pq = tblq[e], col = alist[s] * P.CS;
        pairs[i] = (col + (pq & 0xff)) | ((col + (pq >> 8)) << 16);  // |v0|^2, V_p.v0, V_p.V_q
      }
      __syncthreads();
      const int RSpost = choose_rs((nact * T + 71) / 72);
      reduce_pairs(nact * T, sm + pl.ppost, RSpost);
      for (int i = tid; i < nact; i += RBO_THREADS) colidx[i] = alist[i] * P.CS;
      __syncthreads();
      tri_solve<false>(nact, nf);  // w = L^-T v0 (rbs.jl:525)
      __syncthreads();
      const int RShess = choose_rs(nact * ((d * (d + 1) / 2 + 55) / 56));
      reduce_hess(nact, pt, cb, cb, RShess);
      __syncthreads();
      // per-start logic: one warp per active slot
      for (int s = warp; s < nact; s += RBO_NWARPS) {
        const int sl = alist[s];
        assemble_warp(sl, s, nact, nact * q1, nact * T, s * T, RSpre, RSpost, RShess, misc[1]);
        slot_logic_warp(sl);
      }
      __syncthreads();
      if (tid == 0) {
        int na2 = 0;
        for (int i = 0; i < nact; ++i) {
          const int sl = alist[i];
          if (phase[sl] == 1) { alist[na2++] = sl; continue; }
          // finished start: candidate (discard NaN, rbf_optim.jl:96), first minimum wins (rbf_optim.jl:97)
          const int sid = sstart[sl];
          const double f = (sm + pl.sf)[sl];
          const double* x = sm + pl.sx + sl * d;
          bool bad = !isfinite(f);
          for (int a = 0; a < d; ++a) bad = bad || isnan(x[a]);
          si[I_EVALS] += sevals[sl];
          if (P.start_status) P.start_status[tape_off + sid] = sstat[sl];
          if (P.start_iters) P.start_iters[tape_off + sid] = siter[sl];
          if (!bad && (si[I_BEST] < 0 || f < misc[0] || (f == misc[0] && sid < si[I_BEST]))) {
            si[I_BEST] = sid; misc[0] = f;
            for (int a = 0; a < d; ++a) bestx[a] = x[a];
          }
          if (si[I_NEXT] < P.S) { load_start(sl, si[I_NEXT]); si[I_NEXT] += 1; alist[na2++] = sl; }
        }
        si[I_NACT] = na2;
      }
      __syncthreads();
      nact = si[I_NACT];
    }
    if (tid == 0 && si[I_BEST] < 0) {
      if (si[I_TSTATUS] == RBO_TRAJ_OK) si[I_TSTATUS] = RBO_TRAJ_ALL_STARTS_NAN;
      for (int a = 0; a < d; ++a) bestx[a] = nan("");
    }
    __syncthreads();
  }
};

}  // namespace

__global__ void __launch_bounds__(RBO_THREADS, 1) rbo_rollout_kernel(const __grid_constant__ DevProblem P) {
  extern __shared__ __align__(16) double smem[];
  K k(P, smem);
  const int tid = k.tid, d = P.d, N8 = P.N8, NR = P.NR, RP = P.RP, h = P.h, q1 = d + 1, T = q1 * (q1 + 1) / 2;
  double* bestx = k.bestx;
  double* misc = k.misc;  // misc[0] best f, misc[1] fstar, misc[2..7] scalars, misc[8..] scratch
  int* si = k.si;
  k.build_tables();

  for (;;) {
    __syncthreads();
    if (tid == 0) si[I_M] = atomicAdd(P.work_counter, 1);
    __syncthreads();
    const int m = si[I_M];
    if (m >= P.M) break;

    // ---- trajectory init: reset!(fs) (rbs.jl:476-480) ----
    for (int i = tid; i < (N8 + RBO_MAXFAN) * 8; i += RBO_THREADS) k.Fp[i] = 0.0;
    for (int i = tid; i < 64; i += RBO_THREADS) k.G[i] = ((i >> 3) == (i & 7)) ? 1.0 : 0.0;
    for (int i = tid; i < (h + 2) * NR; i += RBO_THREADS) k.cs[i] = (i < N8) ? __ldg(P.c0 + i) : 0.0;
    for (int i = tid; i < NR; i += RBO_THREADS) k.u[i] = (i < N8) ? __ldg(P.u0 + i) : 0.0;
    __syncthreads();
    if (tid < 8) k.Fp[(size_t)(N8 + tid) * 8 + tid] = 1.0;  // inverse of the identity fantasy block
    if (tid == 0) { misc[1] = P.ymin_base; si[I_TSTATUS] = RBO_TRAJ_OK; }
    k.nf = 0;
    k.set_column(k.CCOL, k.cs);
    __syncthreads();

    if (P.flags & RBO_FLAG_MYOPIC_INTERNAL) {
      // multistart_base_solve!(::Surrogate, ...) (rbf_optim.jl:103-134): the base surrogate, no fantasies
      k.multistart((size_t)m * P.S);
      if (tid < d) P.xs[(size_t)m * d + tid] = bestx[tid];
      if (tid == 0) {
        P.values[m] = -misc[0];
        if (P.n_evals) P.n_evals[m] = si[I_EVALS];
        if (P.status) P.status[m] = si[I_TSTATUS];
        if (P.best_index) P.best_index[m] = si[I_BEST];
        if (P.grad_case) P.grad_case[m] = 0;
      }
      continue;
    }

    for (int step = 0; step <= h; ++step) {
      // ============ choose the location x_step ============
      if (step == 0) {
        if (tid < d) bestx[tid] = P.x0[tid];  // rollout.jl:46
      } else if (P.flags & RBO_FLAG_TEACHER_FORCED) {
        if (tid < d) bestx[tid] = P.x_forced[((size_t)m * h + (step - 1)) * d + tid];
        if (tid == 0) { si[I_EVALS] = 0; misc[0] = nan(""); }
      } else {
        // multistart_base_solve!(fs, xnext; fantasy_index = step-1) (rollout.jl:58-66, rbf_optim.jl:68-101);
        // column CCOL holds cs[fantasy_index + 2] (1-based) = the coefficients after `step` fantasies
        k.multistart(((size_t)m * h + (step - 1)) * P.S);
      }
      __syncthreads();
      if (step > 0 && tid == 0) {
        if (P.n_evals) P.n_evals[(size_t)m * h + step - 1] = si[I_EVALS];
        if (P.alphas) P.alphas[(size_t)m * h + step - 1] = -misc[0];
      }

      // ============ joint draw at x_step (observables.jl:106-121, rbs.jl:588-611) and condition! (rbs.jl:431-441) ============
      {
        auto pt = [&](int) { return (const double*)bestx; };
        auto cb0 = [&](int) { return 0; };
        k.fill_columns(1, pt, cb0);
        k.set_column(k.UCOL, k.u);
        for (int i = tid; i < q1; i += RBO_THREADS) { k.colidx[i] = i; k.pairs[i] = i | (k.CCOL << 16); }
        __syncthreads();
        const int RSpre = k.choose_rs(1);
        k.reduce_pairs(q1, smem + k.pl.ppre, RSpre);  // mu, grad mu
        __syncthreads();
        k.tri_solve<true>(q1, k.nf);
        for (int e = tid; e <= T; e += RBO_THREADS) {
          if (e < T) { int pq = k.tblq[e]; k.pairs[e] = (pq & 0xff) | ((pq >> 8) << 16); }
          else k.pairs[e] = 0 | (k.UCOL << 16);  // l . u
        }
        __syncthreads();
        const int RSpost = k.choose_rs((T + 1 + 71) / 72);
        k.reduce_pairs(T + 1, smem + k.pl.ppost, RSpost);
        __syncthreads();
        // Sigma = Dk(0) - A K^-1 A' = Dk(0) - V'V (rbs.jl:531-536)
        double* Sg = misc + 8;  // (d+1) x (d+1)
        for (int e = tid; e <= T; e += RBO_THREADS) {
          double acc = 0.0;
          for (int r = 0; r < RSpost; ++r) acc += (smem + k.pl.ppost)[(size_t)r * (T + 1) + e];
          if (e < T) {
            const int pq = k.tblq[e], p = pq & 0xff, q = pq >> 8;
            const double dk = (p == q) ? (p == 0 ? P.k0 : -P.d2k0) : 0.0;  // eval_Dk(kernel, 0) rbf.jl:152-159
            Sg[p * q1 + q] = dk - acc;
            Sg[q * q1 + p] = dk - acc;
            if (e == 0) misc[4] = acc;  // |l|^2
          } else misc[5] = acc;         // l . u
        }
        if (tid < q1) {
          double acc = 0.0;
          for (int r = 0; r < RSpre; ++r) acc += (smem + k.pl.ppre)[(size_t)r * q1 + tid];
          misc[8 + q1 * q1 + tid] = acc;  // [mu, grad mu]
        }
        __syncthreads();
        if (tid == 0) {
          const int r = k.nf;  // index of the new fantasy row
          const double* dmu = misc + 8 + q1 * q1;
          bool pd = chol_inplace(Sg, q1, q1);
          if (!pd && si[I_TSTATUS] == RBO_TRAJ_OK) si[I_TSTATUS] = RBO_TRAJ_NOT_PD_JOINT;
          const double* rnm = P.rn + (size_t)m + (size_t)P.M * q1 * step;
          double yv = dmu[0] + Sg[0] * __ldg(rnm);
          for (int a = 0; a < d; ++a) {
            double v = dmu[1 + a];
            for (int j = 0; j <= a + 1; ++j) v += Sg[(a + 1) * q1 + j] * __ldg(rnm + (size_t)P.M * j);
            k.gyf[r * d + a] = v;
          }
          k.yf[r] = yv;
          for (int a = 0; a < d; ++a) k.Xf[r * d + a] = bestx[a];
          misc[1] = fmin(misc[1], yv);
          // new Cholesky row (rbs.jl:403-420): l = L^-1 k (already in column 0), l_rr = sqrt(k0 + sigma_n2 - l.l)
          double s = (P.k0 + P.sigma_n2) - misc[4];
          if (!(s > 0.0) && si[I_TSTATUS] == RBO_TRAJ_OK) si[I_TSTATUS] = RBO_TRAJ_NOT_PD_ROW;
          double lrr = sqrt(s);
          for (int j = 0; j < r; ++j) k.G[r * 8 + j] = k.V[(size_t)(N8 + j) * RP];
          k.G[r * 8 + r] = lrr;
          // inverse of the lower-triangular 8x8 fantasy block, stored k-major: Fp[(N8 + kk)*8 + rr] = Ginv[rr][kk]
          for (int cc = 0; cc < 8; ++cc) {
            double col[8];
            for (int rr = 0; rr < 8; ++rr) {
              double t = (rr == cc) ? 1.0 : 0.0;
              for (int j = cc; j < rr; ++j) t -= k.G[rr * 8 + j] * col[j];
              col[rr] = (rr < cc) ? 0.0 : t / k.G[rr * 8 + rr];
            }
            for (int rr = 0; rr < 8; ++rr) k.Fp[(size_t)(N8 + cc) * 8 + rr] = col[rr];
          }
          k.u[N8 + r] = (yv - misc[5]) / lrr;
        }
        __syncthreads();
        {
          const int r = k.nf;
          for (int j = tid; j < N8; j += RBO_THREADS) k.Fp[(size_t)j * 8 + r] = k.V[(size_t)j * RP];
          for (int j = tid; j < NR; j += RBO_THREADS) k.V[(size_t)j * RP + k.CCOL] = k.u[j];
          if (tid == 0) k.colidx[0] = k.CCOL;
        }
        k.nf += 1;
        __syncthreads();
        // coefficients: c = L^-T (L^-1 y) (rbs.jl:422-429); L^-1 y is maintained incrementally in u
        k.tri_solve<false>(1, k.nf);
        __syncthreads();
        for (int j = tid; j < NR; j += RBO_THREADS) k.cs[(size_t)(step + 1) * NR + j] = k.V[(size_t)j * RP + k.CCOL];
        __syncthreads();
      }
    }

    // ============ resolve (rollout.jl:108-111) and bookkeeping ============
    if (tid == 0) {
      double best = k.yf[0];
      int t = 0;
      for (int j = 1; j <= h; ++j) if (k.yf[j] < best) { best = k.yf[j]; t = j; }  // findmin: first minimum (rollout.jl:77-82)
      P.values[m] = fmax(P.fmini - best, 0.0);
      si[I_T] = t;
      int tc = 0;
      if (P.mode == RBO_MODE_VALUE_GRAD) tc = (P.fmini <= best) ? 1 : (t == 0 ? 2 : 3);  // rollout.jl:241-251
      si[I_CASE] = tc;
      if (P.best_index) P.best_index[m] = t;
      if (P.grad_case) P.grad_case[m] = tc;
    }
    if (P.xs) for (int i = tid; i < (h + 1) * d; i += RBO_THREADS) P.xs[(size_t)m * (h + 1) * d + i] = k.Xf[i];
    if (P.gys) for (int i = tid; i < (h + 1) * d; i += RBO_THREADS) P.gys[(size_t)m * (h + 1) * d + i] = k.gyf[i];
    if (P.ys) for (int i = tid; i <= h; i += RBO_THREADS) P.ys[(size_t)m * (h + 1) + i] = k.yf[i];
    __syncthreads();

    // ============ gradient(T) (rollout.jl:233-277) ============
    if (P.mode == RBO_MODE_VALUE_GRAD) {
      const int tc = si[I_CASE], t = si[I_T], nth = P.ntheta;
      if (tc == 1) {
        for (int i = tid; i < d; i += RBO_THREADS) P.grad_x[(size_t)m * d + i] = 0.0;
        for (int i = tid; i < nth; i += RBO_THREADS) P.grad_theta[(size_t)m * nth + i] = 0.0;
      } else if (tc == 2) {
        for (int i = tid; i < d; i += RBO_THREADS) P.grad_x[(size_t)m * d + i] = -k.gyf[i];  // -get_gradient(at = 1)
        for (int i = tid; i < nth; i += RBO_THREADS) P.grad_theta[(size_t)m * nth + i] = 0.0;
      } else {
        // adjoint work area: xbars[(j)*d], acc[(j)*d] for j = 0..8 ; ybars[0..11] ; gx[d] ; gth
        double* xbars = k.adj; double* accr = k.adj + 9 * d; double* ybars = k.adj + 18 * d; double* gxa = ybars + 12; double* gtha = gxa + d;
        for (int i = tid; i < 19 * d + 16; i += RBO_THREADS) k.adj[i] = 0.0;
        __syncthreads();
        if (tid == 0) ybars[t + 1] = 1.0;  // rollout.jl:256
        const int CB_RAW = 0, CB_SOL = d + 3, CB_U = 2 * d + 4, CB_Q = 3 * d + 5;
        const int nd = 2 * d + 1;
        const double* dd_m = P.dual_dirs ? P.dual_dirs + (size_t)m * h * d : nullptr;
        for (int i = t; i >= 1; --i) {
          // ---- re-evaluate policy solve i: fs(x_i, theta; fantasy_index = i-1) (rollout.jl:114-124) ----
          k.nf = i;
          const double* c = k.cs + (size_t)i * NR;
          const double* xi = k.Xf + (size_t)i * d;
          auto pt = [&](int) { return xi; };
          auto cbr = [&](int) { return CB_RAW; };
          auto cbs = [&](int) { return CB_SOL; };
          __syncthreads();
          k.fill_columns(1, pt, cbr);
          k.set_column(k.CCOL, c);
          __syncthreads();
          for (int idx = tid; idx < NR * q1; idx += RBO_THREADS) {
            int j = idx / q1, q = idx - j * q1;
            k.V[(size_t)j * RP + CB_SOL + q] = k.V[(size_t)j * RP + CB_RAW + q];
          }
          for (int q = tid; q < q1; q += RBO_THREADS) { k.colidx[q] = CB_SOL + q; k.pairs[q] = (CB_RAW + q) | (k.CCOL << 16); }
          __syncthreads();
          const int RSpre = k.choose_rs(1);
          k.reduce_pairs(q1, smem + k.pl.ppre, RSpre);
          k.tri_solve<true>(q1, k.nf);
          for (int e = tid; e < T; e += RBO_THREADS) { int pq = k.tblq[e]; k.pairs[e] = (CB_SOL + (pq & 0xff)) | ((CB_SOL + (pq >> 8)) << 16); }
          __syncthreads();
          const int RSpost = k.choose_rs((T + 71) / 72);
          k.reduce_pairs(T, smem + k.pl.ppost, RSpost);
          __syncthreads();
          k.tri_solve<false>(q1, k.nf);  // w = SOL[:,0], Dw = SOL[:,1..d] (rbs.jl:525-526)
          __syncthreads();
          const int RShess = k.choose_rs((d * (d + 1) / 2 + 55) / 56);
          k.reduce_hess(1, pt, cbr, cbs, RShess);
          __syncthreads();
          if (k.warp == 0) {
            double fst = P.ymin_base;  // f* over the active slice y[1:N+i]
            for (int j = 0; j < i; ++j) fst = fmin(fst, k.yf[j]);
            k.assemble_warp(0, 0, 1, q1, T, 0, RSpre, RSpost, RShess, fst);
            if (tid == 0) {
              misc[2] = fst;
              // ---- solve_dual_x for j = i (rollout.jl:150-191) with the contributions of later solves already pushed ----
              double* Hlu = smem + k.pl.sA;  // slot-0 scratch (d x d)
              const double* Href = smem + k.pl.sHref;
              int piv[RBO_MAXD];
              double det;
              for (int a = 0; a < d; ++a) for (int b = 0; b < d; ++b) Hlu[a * d + b] = Href[a * d + b];
              lu_factor(Hlu, d, piv, &det);
              double* xb = xbars + (size_t)i * d;
              if (det < P.htol) {  // rollout.jl:159-161 (Q3)
                for (int a = 0; a < d; ++a) xb[a] = 0.0;
                misc[3] = 0.0;
              } else {
                for (int a = 0; a < d; ++a) xb[a] = -k.gyf[(size_t)(i - 1) * d + a] * ybars[i + 1] - accr[(size_t)i * d + a];  // rollout.jl:164-165 (Q4)
                for (int a = 0; a < d; ++a) for (int b = 0; b < d; ++b) Hlu[a * d + b] = Href[b * d + a];  // hessian(sx)' (rollout.jl:188)
                double det2;
                if (!lu_factor(Hlu, d, piv, &det2) && si[I_TSTATUS] == RBO_TRAJ_OK) si[I_TSTATUS] = RBO_TRAJ_SINGULAR_HESSIAN;
                lu_solve(Hlu, d, piv, xb);
                misc[3] = 1.0;
                // gather_q (rollout.jl:220-231): d2alpha/dx dtheta = grad_mu g_mu_theta + grad_sigma g_sigma_theta (rbs.jl:575-577)
                const double* gh = smem + k.pl.sgh; const double* dmu = smem + k.pl.sdmu; const double* dsg = smem + k.pl.sdsig;
                double s = 0.0;
                for (int a = 0; a < d; ++a) s += (dmu[a] * gh[3] + dsg[a] * gh[4]) * xb[a];
                gtha[0] += s;
              }
            }
          }
          __syncthreads();
          if (misc[3] == 0.0) continue;  // xbar_i = 0: every term it feeds vanishes
          const double* xb = xbars + (size_t)i * d;
          const double sigma = (smem + k.pl.sgh)[5];
          for (int p = 0; p < i; ++p) {
            // perturbation of fantasy location x_p inside policy solve i: d unit directions (spatial, rbs.jl:652-694)
            // and one direction dual_dirs[:, p] (data perturbation surrogate, rbs.jl:711-760)
            const int rowp = N8 + p;
            const double* xp = k.Xf + (size_t)p * d;
            // phase A: u_a = grad_k(X_a - X_p) . (-dx) per row (rbf.jl:210-228 restricted to the moved column)
            for (int j = tid; j < NR; j += RBO_THREADS) {
              double* rowU = k.V + (size_t)j * RP + CB_U; double* rowQ = k.V + (size_t)j * RP + CB_Q;
              if (!k.row_active(j) || j == rowp) { for (int q = 0; q < q1; ++q) { rowU[q] = 0.0; rowQ[q] = 0.0; } continue; }
              double rho2 = 0.0;
              for (int a = 0; a < d; ++a) { double r = k.xcoord(j, a) - xp[a]; rho2 = fma(r, r, rho2); }
              double psi, a_, b_, gb_;
              kern_radial(P.kern, rho2, psi, a_, b_, gb_);
              if (!(rho2 > 0.0)) b_ = 0.0;
              double cp = c[rowp], rd = 0.0;
              for (int a = 0; a < d; ++a) {
                double r = k.xcoord(j, a) - xp[a];
                double uu = -b_ * r;
                rowU[a] = uu; rowQ[a] = uu * cp;
                if (dd_m) rd += r * dd_m[(size_t)p * d + a];
              }
              double ud = -b_ * rd;
              rowU[d] = ud; rowQ[d] = ud * cp;
            }
            for (int e = tid; e < 2 * q1; e += RBO_THREADS) {
              int q = e >> 1;
              k.pairs[e] = (CB_U + q) | (((e & 1) ? CB_SOL : k.CCOL) << 16);  // u.c (even) ; u.w (odd)
            }
            __syncthreads();
            // phase B: (dK c)_p = u.c ; u.w
            const int RSb = k.choose_rs(1);
            k.reduce_pairs(2 * q1, smem + k.pl.ppost, RSb);
            __syncthreads();
            double* uw = misc + 8;  // [q1]
            for (int e = tid; e < 2 * q1; e += RBO_THREADS) {
              double acc = 0.0;
              for (int r = 0; r < RSb; ++r) acc += (smem + k.pl.ppost)[(size_t)r * 2 * q1 + e];
              if (e & 1) uw[e >> 1] = acc; else k.V[(size_t)rowp * RP + CB_Q + (e >> 1)] = acc;
            }
            for (int q = tid; q < q1; q += RBO_THREADS) k.colidx[q] = CB_Q + q;
            __syncthreads();
            // phase C: dc = -K^-1 (dK c) (rbs.jl:675)
            k.tri_solve<true>(q1, k.nf);
            __syncthreads();
            k.tri_solve<false>(q1, k.nf);
            for (int e = tid; e < q1 * nd; e += RBO_THREADS) {
              int q = e / nd, w = e - q * nd;
              int c1 = (w <= d) ? CB_RAW + w : CB_SOL + (w - d);  // w in [0,d]: raw kx / grad_kx ; w in (d, 2d]: Dw column
              int c2 = (w <= d) ? CB_Q + q : CB_U + q;
              k.pairs[e] = c1 | (c2 << 16);
            }
            __syncthreads();
            // phase D: dots kx.Q, grad_kx.Q, Dw'U per direction
            const int RSd = k.choose_rs((q1 * nd + 71) / 72);
            k.reduce_pairs(q1 * nd, smem + k.pl.ppost, RSd);
            __syncthreads();
            // phase E: assemble delta grad alpha per direction and push it into the earlier duals
            if (tid < q1) {
              const int q = tid;
              const double* gh = smem + k.pl.sgh; const double* dmu = smem + k.pl.sdmu; const double* dsg = smem + k.pl.sdsig;
              const double* rowp_v = k.V + (size_t)rowp * RP;
              auto dq = [&](int w) { double acc = 0.0; for (int r = 0; r < RSd; ++r) acc += (smem + k.pl.ppost)[(size_t)r * q1 * nd + q * nd + w]; return acc; };
              const double cp = c[rowp], wp = rowp_v[CB_SOL];
              double dxv[RBO_MAXD];
              for (int a = 0; a < d; ++a) dxv[a] = (q < d) ? (a == q ? 1.0 : 0.0) : (dd_m ? dd_m[(size_t)p * d + a] : 0.0);
              // dkx_p = grad_k(x - X_p).(-dx) (rbf.jl:230-245); dgkx_p = Hk(x - X_p)(-dx) (rbf.jl:247-262)
              double dkx = 0.0, rdx = 0.0;
              for (int a = 0; a < d; ++a) { dkx -= rowp_v[CB_RAW + 1 + a] * dxv[a]; rdx += (xi[a] - xp[a]) * dxv[a]; }
              const double a_ = rowp_v[CB_RAW + d + 1], b_ = rowp_v[CB_RAW + d + 2];
              double dmu_v = dkx * cp - dq(0);                                        // rbs.jl:680 (dc = -Q)
              double dsig_v = (-2.0 * dkx * wp + 2.0 * wp * uw[q]) / (2.0 * sigma);   // rbs.jl:683, w'dK w = 2 w_p (u.w)
              GPart ghat = rule_eval(P.rule_id, P.sigma_tol, dmu_v, dsig_v, P.theta1, misc[2]);  // rbs.jl:687-688 (Q6)
              double push = 0.0;
              for (int a = 0; a < d; ++a) {
                double dgkx = -(a_ * rdx * (xi[a] - xp[a]) + b_ * dxv[a]);
                double dgmu = dgkx * cp - dq(1 + a);                                  // rbs.jl:681
                double val = gh[1] * dgmu + ghat.g_mu * dmu[a] + ghat.g_sig * dsg[a];
                if (q < d) {                                                          // spatial: + g_sigma * delta grad sigma (rbs.jl:684,690)
                  double dwdkw = wp * dq(d + 1 + a) + rowp_v[CB_SOL + 1 + a] * uw[q];
                  double dgsig = (dwdkw - dgkx * wp - rowp_v[CB_SOL + 1 + a] * dkx - dsig_v * dsg[a]) / sigma;
                  val += gh[2] * dgsig;
                }
                push += val * xb[a];
              }
              if (q < d) {
                if (p == 0) gxa[q] += push;               // gather_g: g[i+1]' * xbars[i] (rollout.jl:199-215, 271)
                else accr[(size_t)p * d + q] += push;     // solve_dual_x(p): x_dual -= dri' * xbars[i] (rollout.jl:185)
              } else {
                ybars[p + 1] += push;                     // solve_dual_y(solve_index = p) (rollout.jl:144)
              }
            }
            __syncthreads();
          }
        }
        // g[1] = grad mu(x_0) under the base GP (rollout.jl:194-196), final assembly rollout.jl:267-276
        k.nf = 0;
        {
          const double* x0p = k.Xf;
          auto pt = [&](int) { return x0p; };
          auto cb0 = [&](int) { return 0; };
          __syncthreads();
          k.fill_columns(1, pt, cb0);
          k.set_column(k.CCOL, k.cs);
          for (int q = tid; q < q1; q += RBO_THREADS) k.pairs[q] = q | (k.CCOL << 16);
          __syncthreads();
          const int RSpre = k.choose_rs(1);
          k.reduce_pairs(q1, smem + k.pl.ppre, RSpre);
          __syncthreads();
          for (int a = tid; a < d; a += RBO_THREADS) {
            double dmu = 0.0;
            for (int r = 0; r < RSpre; ++r) dmu += (smem + k.pl.ppre)[(size_t)r * q1 + 1 + a];
            P.grad_x[(size_t)m * d + a] = -(dmu * ybars[1] + gxa[a]);
          }
          for (int a = tid; a < nth; a += RBO_THREADS) P.grad_theta[(size_t)m * nth + a] = (a == 0) ? -gtha[0] : 0.0;
        }
        k.nf = h + 1;
      }
    }
    __syncthreads();
    if (tid == 0 && P.status) P.status[m] = si[I_TSTATUS];
  }
}

// ----------------------------------------------------------------------------------------------------
// gen_low_discrepancy_sequence (utils.jl:65-74) on the device: Sobol (utils.jl:4-13, Joe-Kuo direction numbers,
// Gray-code order, origin skipped) -> Box-Muller with log10 and pair indexing (utils.jl:23-43, Q8) -> the
// column-major reshape(N, M, D, H) and removal of the padding coordinate.
// ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned sobol_point(const unsigned* dirs, int dim, unsigned k) {
  unsigned g = k ^ (k >> 1), x = 0;
  const unsigned* v = dirs + dim * 32;
  for (int b = 0; g != 0u; ++b, g >>= 1)
    if (g & 1u) x ^= v[b];
  return x;
}

__global__ void rbo_normals_kernel(const unsigned* __restrict__ dirs, double* __restrict__ out, int M_total, int d, int H, int m_begin, int m_count) {
  const int q1 = d + 1, D = q1 + ((q1 & 1) ? 1 : 0);
  const size_t total = (size_t)m_count * q1 * H;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    int ml = (int)(idx % m_count);
    int kq = (int)((idx / m_count) % q1);
    int t = (int)(idx / ((size_t)m_count * q1));
    size_t q = (size_t)(m_begin + ml) + (size_t)M_total * kq + (size_t)M_total * D * t;  // flat index into the D x (M*H) normals
    int coord = (int)(q % D);
    unsigned pnt = (unsigned)(q / D) + 1u;  // Sobol point index (the origin is skipped)
    int c0 = coord & ~1;                    // pair (c0, c0+1)
    double u1 = (double)sobol_point(dirs, c0, pnt) * (1.0 / 4294967296.0);
    double u2 = (double)sobol_point(dirs, c0 + 1, pnt) * (1.0 / 4294967296.0);
    double rad = sqrt(-2.0 * log10(u1));
    const double two_pi = 6.283185307179586;
    out[idx] = (coord & 1) ? rad * sin(two_pi * u2) : rad * cos(two_pi * u2);
  }
}

__global__ void rbo_sobol_kernel(const unsigned* __restrict__ dirs, unsigned* __restrict__ out_u32, double* __restrict__ out_f64, int dim, int npoints,
                                 const double* __restrict__ lbs, const double* __restrict__ ubs) {
  const size_t total = (size_t)dim * npoints;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    int a = (int)(idx % dim);
    unsigned pnt = (unsigned)(idx / dim) + 1u;
    unsigned x = sobol_point(dirs, a, pnt);
    if (out_u32) out_u32[idx] = x;
    if (out_f64) {
      double uu = (double)x * (1.0 / 4294967296.0);
      out_f64[idx] = lbs ? lbs[a] + (ubs[a] - lbs[a]) * uu : uu;
    }
  }
}

// Per-handle statistics of the last rollout: sums[0] = n, then [n*mean, M2, n*mean^2] for the value and every gradient
// row (two-pass: mean first, then centred squares -- rollout.jl:328-337), then the acquisition evaluations per step
// (max(h,1) entries), then a histogram over t = 0..h of the case-3 trajectories, then the number of failed trajectories.
// One CTA; M is at most a few 10^5.
__global__ void rbo_stats_kernel(const double* __restrict__ values, const double* __restrict__ gx, const double* __restrict__ gth,
                                 const int* __restrict__ n_evals, const int* __restrict__ best_index, const int* __restrict__ grad_case,
                                 const int* __restrict__ status, int M, int d, int nth, int h, double* __restrict__ sums) {
  __shared__ double red[32];
  __shared__ double s_mean;
  const int nrows = 1 + d + nth, nw = blockDim.x >> 5;
  auto block_sum = [&](double acc) -> double {
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < nw; ++w) tot += red[w];
    return tot;
  };
  for (int row = 0; row < nrows; ++row) {
    const double* base; int stride;
    if (row == 0) { base = values; stride = 1; }
    else if (row <= d) { base = gx ? gx + (row - 1) : nullptr; stride = d; }
    else { base = gth ? gth + (row - 1 - d) : nullptr; stride = nth; }
    double mean = 0.0, m2 = 0.0;
    if (base) {
      double acc = 0.0;
      for (int i = threadIdx.x; i < M; i += blockDim.x) acc += base[(size_t)i * stride];
      mean = block_sum(acc) / M;
      if (threadIdx.x == 0) s_mean = mean;
      __syncthreads();
      mean = s_mean;
      acc = 0.0;
      for (int i = threadIdx.x; i < M; i += blockDim.x) { double v = base[(size_t)i * stride] - mean; acc += v * v; }
      m2 = block_sum(acc);
    }
    if (threadIdx.x == 0) {
      if (row == 0) sums[0] = (double)M;
      sums[1 + 3 * row + 0] = M * mean;
      sums[1 + 3 * row + 1] = m2;
      sums[1 + 3 * row + 2] = M * mean * mean;
    }
  }
  const int hh = h > 1 ? h : 1;
  double* ev = sums + 1 + 3 * nrows;
  for (int j = 0; j < hh; ++j) {
    double acc = 0.0;
    if (n_evals && h > 0) for (int i = threadIdx.x; i < M; i += blockDim.x) acc += n_evals[(size_t)i * h + j];
    double tot = block_sum(acc);
    if (threadIdx.x == 0) ev[j] = tot;
  }
  double* hist = ev + hh;
  for (int t = 0; t <= h + 1; ++t) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
      if (t <= h) acc += (grad_case && best_index && grad_case[i] == 3 && best_index[i] == t) ? 1.0 : 0.0;
      else acc += (status && status[i] != 0) ? 1.0 : 0.0;
    }
    double tot = block_sum(acc);
    if (threadIdx.x == 0) hist[t + (t > h ? 1 : 0)] = tot;
  }
}

// FP64 FMA micro-benchmark: the roofline denominator for this path (MEASURED_PEAKS.json has no FP64 figure).
__global__ void rbo_fp64_peak_kernel(double* out, int iters) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace rbo
