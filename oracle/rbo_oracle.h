/*
 * rbo_oracle.h -- C interface of the CPU restatement ("oracle") of the reference's
 * Monte-Carlo rollout acquisition estimator and adjoint gradient.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under the product package may include, link or call
 * this. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the reference (Julia) cannot run in this environment and ships no golden
 * vectors for this path. The restatement follows the reference files
 * line by line (citations at each function in rbo_oracle.cpp) and is validated by finite
 * differences, closed forms and an independent numpy restatement (oracle/py_restatement.py).
 */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_KERNEL_MATERN12 = 0, ORC_KERNEL_MATERN32 = 1, ORC_KERNEL_MATERN52 = 2, ORC_KERNEL_SE = 3, ORC_KERNEL_PERIODIC = 4 };
enum { ORC_RULE_EI = 0, ORC_RULE_POI = 1, ORC_RULE_LCB = 2 };
enum { ORC_MODE_VALUE = 0, ORC_MODE_VALUE_GRAD = 1 };
enum {
  ORC_FLAG_TEACHER_FORCED = 1,  /* x_1..x_h taken from x_forced instead of the inner solve */
  ORC_FLAG_FAST_PERTURB = 2,    /* rank-2 shortcut for dK*v instead of the reference's dense build */
  ORC_FLAG_FACTORED = 4,        /* sigma^2 = k0 - |L^-1 kx|^2 etc. (forward-solve formulation) */
  ORC_FLAG_GAUSS_HERMITE = 8    /* GaussHermiteObservable (observables.jl:32-81,157) instead of StochasticObservable */
};
/* per-trajectory status */
enum {
  ORC_OK = 0,
  ORC_NOT_PD_ROW = 1,     /* rbs.jl:412 cholesky(C - L21 L21') would throw */
  ORC_NEG_VARIANCE = 2,   /* rbs.jl:528 sqrt of a negative number would throw */
  ORC_NOT_PD_JOINT = 3,   /* rbs.jl:537 cholesky of the joint covariance would throw */
  ORC_ALL_STARTS_NAN = 4, /* rbf_optim.jl:96-97 findmin over an empty candidate list */
  ORC_SINGULAR_HESSIAN = 5 /* rollout.jl:188 LU solve hit an exact zero pivot */
};
/* per-start solver status */
enum { ORC_SOLVE_CONVERGED = 0, ORC_SOLVE_MAXIT = 1, ORC_SOLVE_STEP_TINY = 2, ORC_SOLVE_PRED_TINY = 3, ORC_SOLVE_STALLED = 4, ORC_SOLVE_NAN = 5, ORC_SOLVE_FINAL_STEP = 6 };

typedef struct {
  int maxit;      /* outer iterations (accepted steps) per start */
  int maxtry;     /* consecutive rejected / non-descent trust-region steps before a start is declared stalled */
  double gtol;    /* stop when max |projected gradient| <= gtol * max(1, |alpha|) */
  double xtol;    /* stop when max |step| <= xtol * max(1, max |x|) */
  double pred_tol;/* stop when predicted decrease <= pred_tol * max(1, |alpha|) */
  double eta;     /* acceptance ratio rho >= eta (optim.jl:99 uses 0.1) */
  double delta0_box; /* initial trust-region radius: min(delta0_box * widest box side, */
  double delta0_ell; /*                                  delta0_ell * kernel length-scale theta_k[0]) */
  double stol;    /* an interior Newton step with max |s| <= stol * max(1, max |x|) is taken without re-evaluation and ends the start */
} orc_solver_opts;

typedef struct {
  int d, N, h, M, S;  /* S = number of start columns (the reference passes S+2) */
  int kernel_id, nktheta;
  double ktheta[4];
  int rule_id;
  double sigma_tol;
  double sigma_n2;
  const double* X; int ldX;   /* d x N, column-major */
  const double* L; int ldL;   /* N x N lower, column-major */
  const double* y;            /* N */
  const double* c;            /* N  (= K^-1 y) */
  const double* x0;           /* d */
  const double* theta; int ntheta; /* decision-rule hyper-parameters (EI: xi) */
  const double* lbs; const double* ubs;
  double fmini;               /* rollout.jl:109 minimum over the zero-padded y of the base surrogate */
  const double* rn; int rn_hp1; /* M x (d+1) x rn_hp1, column-major (sample index fastest) */
  const double* starts;       /* d x S */
  const double* dual_dirs;    /* d x h x M : the rand(dim) draws of rollout.jl:133, indexed [k, solve_index, m] */
  const double* x_forced;     /* d x h x M, only with ORC_FLAG_TEACHER_FORCED */
  int mode, flags;
  double htol;                /* rollout.jl:156 */
  orc_solver_opts solver;
  int nthreads;               /* <=0: all */
  const double* gh_nodes;     /* (h+1) x M: nodes[indices[m]] of simulate_trajectory_ghq (rollout.jl:431-432), step fastest */
  const double* gh_weights;   /* (h+1) x M */
} orc_problem;

typedef struct {
  double* values;      /* M */
  double* grad_x;      /* d x M or NULL */
  double* grad_theta;  /* ntheta x M or NULL */
  int* best_index;     /* M: t of rollout.jl:235 */
  int* grad_case;      /* M: 1,2,3 of rollout.jl:239-251 (0 if value only) */
  int* status;         /* M */
  /* optional tape (may be NULL) */
  double* xs;          /* d x (h+1) x M   fantasy locations x_0..x_h */
  double* ys;          /* (h+1) x M       sampled observations */
  double* gys;         /* d x (h+1) x M   sampled gradients */
  double* alphas;      /* h x M           acquisition value at the chosen x_j */
  int* n_evals;        /* h x M           acquisition evaluations spent at step j (all starts) */
  int* start_status;   /* S x h x M */
  int* start_iters;    /* S x h x M */
  /* optional extended tape (may be NULL): the surrogate evaluation sx_j = fs(x_j; fantasy_index = j-1) at the chosen x_j */
  double* t_mu;        /* h x M */
  double* t_sigma;     /* h x M */
  double* t_dmu;       /* d x h x M */
  double* t_dsigma;    /* d x h x M */
  double* t_Halpha;    /* d x d x h x M   the reference's H alpha (rbs.jl:568, no mu-sigma cross term) */
} orc_outputs;

void orc_default_solver_opts(orc_solver_opts* o);

/* a1: simulate_trajectory_mc (rollout.jl:279-340), per-trajectory part */
int orc_rollout(const orc_problem* p, orc_outputs* out);

/* rollout.jl:328-337: mean and corrected std of a length-M vector with stride */
void orc_mean_std(const double* v, int M, int stride, double* mean, double* std);

/* rbs.jl:77-118: K = eval_KXX + sigma_n2 I, L = chol(K), c = L'\(L\y).  K, L: N x N column-major (ld N) */
int orc_fit_surrogate(int d, int N, const double* X, int ldX, const double* y, int kernel_id, const double* ktheta,
                      double sigma_n2, double* K, double* L, double* c);

/* rbs.jl:482-581 at one point against the base surrogate extended by nf fantasy points.
 * out layout (doubles): [mu, sigma, alpha, fstar, g_mu, g_sigma, g_mumu, g_sigsig, g_theta, g_thth, g_muth, g_sigth,
 *   dmu[d], dsigma[d], dalpha[d], Hmu[d*d], Hsigma[d*d], Halpha_ref[d*d], Halpha_true[d*d], d2alpha_dxdtheta[d]] */
int orc_eval_point(const orc_problem* p, int nf, const double* Xf, const double* yf, const double* x, double* out);

/* rbf_optim.jl:68-101 against the surrogate extended by nf fantasy points: returns argmax x and -alpha */
int orc_multistart_solve(const orc_problem* p, int nf, const double* Xf, const double* yf, double* xbest, double* fbest,
                         int* start_status, int* start_iters, double* start_x, double* start_f);

/* utils.jl:4-13 */
void orc_sobol_uniform(int dim, int npoints, double* out /* dim x npoints col-major */);
void orc_sobol_uint32(int dim, int npoints, unsigned* out /* dim x npoints col-major */);
/* utils.jl:65-74: out is M x (d+1) x H column-major */
void orc_gen_low_discrepancy_sequence(int M, int d, int H, double* out);
/* utils.jl:145-153: out is d x (S+2) */
void orc_generate_initial_guesses(int S, int d, const double* lbs, const double* ubs, double* out);
/* decision-rule value and partials: out = [g, g_mu, g_sig, g_mumu, g_sigsig, g_muth, g_sigth, g_musig] */
void orc_rule_partials(int rule_id, double sigma_tol, double mu, double sigma, double theta1, double fstar, double* out);
/* exact trust-region step of the inner solve (solve_tr, optim.jl:9-51); returns 1 when the constraint is active */
int orc_tr_step(int n, const double* H, const double* g, double Delta, double* p);
/* kernel scalar functions, for the FD ladder: out = [psi, dpsi, d2psi] */
void orc_kernel_scalars(int kernel_id, const double* ktheta, double rho, double* out);

#ifdef __cplusplus
}
#endif
