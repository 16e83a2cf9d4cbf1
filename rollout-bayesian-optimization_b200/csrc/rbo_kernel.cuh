// rbo_kernel.cuh -- parameter block and shared-memory plan of the rollout kernel (host + device).
#pragma once
#include "rbo_device.cuh"

namespace rbo {

// Everything the rollout kernel needs, passed by value as a __grid_constant__ parameter.
struct DevProblem {
  // sizes
  int d, N, N8, nb8;      // input dim; base observations; N rounded up to RBO_PR; N8 / RBO_PR
  int h, S, W, nwaves;    // horizon; start columns; starts processed together (wave); ceil(S / W)
  int CS, RP, NR;         // columns per start slot (d+3); padded V row pitch (doubles); V rows (N8 + RBO_MAXFAN)
  int M;                  // trajectories owned by this handle
  int hp1;                // third dimension of the normals tensor
  int mode, flags, ntheta;
  // model
  KernelSpec kern;
  int rule_id;
  double sigma_tol, sigma_n2, k0, d2k0, ymin_base;
  double fmini, theta1, htol;
  rbo_solver_opts so;
  double x0[RBO_MAXD], lbs[RBO_MAXD], ubs[RBO_MAXD];
  // resident inputs (device)
  const double* Xb;      // [d][N8] coordinate-major base locations (pad columns 0)
  const double* yb;      // [N]
  const double* c0;      // [N8] base coefficients K^-1 y (pad 0)
  const double* u0;      // [N8] L0^-1 y (pad 0)
  const double* Lf;      // forward panels of L0 with inverted diagonal blocks
  const double* Lb;      // backward (transposed) panels of L0 with inverted diagonal blocks
  const double* rn;      // [M][(d+1)][hp1] column-major (sample fastest)
  const double* starts;  // [S][d]
  const double* dual_dirs;  // [M][h][d] or nullptr
  const double* x_forced;   // [M][h][d] or nullptr
  // outputs (device)
  double* values;      // [M]
  double* grad_x;      // [M][d]
  double* grad_theta;  // [M][ntheta]
  int* best_index;     // [M]
  int* grad_case;      // [M]
  int* status;         // [M]
  double* xs;          // [M][h+1][d]
  double* ys;          // [M][h+1]
  double* gys;         // [M][h+1][d]
  double* alphas;      // [M][h]
  int* n_evals;        // [M][h]
  int* start_status;   // [M][h][S]
  int* start_iters;    // [M][h][S]
  int* work_counter;   // dynamic trajectory scheduler
};

// Shared-memory plan (offsets in doubles from the start of dynamic shared memory).
struct SmemPlan {
  int V, Fp, G, cs, u, Xf, yf, gyf;
  int sx, sxt, sg, sH, sA, sp;     // per slot: x, trial x, gradient, Hessian, Cholesky scratch, step
  int e_mu, e_dmu, e_s2, e_tq, e_G, e_HC, e_HW, e_gh;  // per slot evaluation scratch
  int sf, slam, spred, shs;        // per slot scalars
  int bestx, misc, adj;            // [d] best candidate ; scalar/scratch area ; adjoint duals
  int ints;                        // int area (in doubles)
  int total;                       // total doubles
};

__host__ __device__ inline int ncols_adjoint(int d) { return 4 * (d + 1) + 2; }

__host__ __device__ inline SmemPlan make_plan(int d, int N8, int h, int W, int RP, int NR) {
  SmemPlan p;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };  // keep 16-byte alignment
  p.V = take(NR * RP);
  p.Fp = take((N8 + RBO_MAXFAN) * RBO_PR);
  p.G = take(RBO_MAXFAN * RBO_MAXFAN);
  p.cs = take((h + 2) * NR);
  p.u = take(NR);
  p.Xf = take(RBO_MAXFAN * d);
  p.yf = take(RBO_MAXFAN);
  p.gyf = take(RBO_MAXFAN * d);
  int dd = d * d;
  p.sx = take(W * d); p.sxt = take(W * d); p.sg = take(W * d); p.sH = take(W * dd); p.sA = take(W * dd); p.sp = take(W * d);
  p.e_mu = take(W); p.e_dmu = take(W * d); p.e_s2 = take(W); p.e_tq = take(W * d);
  p.e_G = take(W * dd); p.e_HC = take(W * dd); p.e_HW = take(W * dd); p.e_gh = take(W * 8);
  p.sf = take(W); p.slam = take(W); p.spred = take(W); p.shs = take(W);
  p.bestx = take(d);
  p.misc = take(64 + 4 * (d + 1) * (d + 1) + 8 * d);
  p.adj = take(19 * d + 32);
  p.ints = take(64 + 8 * W + ncols_adjoint(d) + W * (d + 1));
  p.total = o;
  return p;
}

}  // namespace rbo
