// rbo_api.cu -- the C ABI of librbo.so (include/rbo.h): handle, device-resident inputs, launches.
// No CPU fallback anywhere: every entry point that computes does so with the sm_100a kernels of rollout_kernel.cu.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "rbo_kernel.cuh"
#include "sobol_dirs.h"

namespace rbo {
__global__ void rbo_rollout_kernel(const __grid_constant__ DevProblem P);
__global__ void rbo_rollout_kernel_largen(const __grid_constant__ DevProblem P);  // same source, work matrix in global memory
__global__ void rbo_normals_kernel(const unsigned* dirs, double* out, int M_total, int d, int H, int m_begin, int m_count);
__global__ void rbo_sobol_kernel(const unsigned* dirs, unsigned* out_u32, double* out_f64, int dim, int npoints, const double* lbs, const double* ubs);
__global__ void rbo_stats_kernel(const double* values, const double* gx, const double* gth, const int* n_evals, const int* best_index, const int* grad_case,
                                 const int* status, int M, int d, int nth, int h, double* sums);
__global__ void rbo_fp64_peak_kernel(double* out, int iters);
__global__ void rbo_tr_step_kernel(const double* H, const double* g, const double* Delta, int n, int B, double* p, int* hit);
__global__ void rbo_gather_sums_kernel(const double* sums, int need, int idx_failed, const int* work_counter, double* out);
__global__ void rbo_lpt_order_kernel(const int* n_evals, const int* grad_case, const int* best_index, int M, int hh, int* order);
// surrogate_kernels.cu
__global__ void rbo_trinv_kernel(const double* L, int N, int N32, double* Linv, int ldi);
__global__ void rbo_pack_fwd_kernel(const double* Linv, int ldi, int nb32, double* Lf);
__global__ void rbo_pack_bwd_kernel(const double* Linv, int ldi, int nb32, double* Lb, double* Lbf);
__global__ void rbo_matvec_lower_kernel(const double* Linv, int ldi, int n, const double* v, double* out);
__global__ void rbo_matvec_lower_t_kernel(const double* Linv, int ldi, int n, const double* v, double scale, double* out);
__global__ void rbo_layout_x_kernel(const double* Xpts, int d, int N, int N8, double* Xb);
__global__ void rbo_kvec_kernel(const double* Xpts, int d, int N, const double* x, KernelSpec kern, double* out);
__global__ void rbo_dots_kernel(const double* l, const double* u, int n, double* scal);
__global__ void rbo_append_row_kernel(double* Linv, int ldi, int n, const double* tmp, const double* scal, double kdiag, double ynew, double* u, int* status);
}  // namespace rbo

using namespace rbo;

static thread_local std::string g_create_error;

struct rbo_handle {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int num_sms = 0, max_smem = 0;
  std::string err;
  rbo_solver_opts so;
  // surrogate
  bool have_sur = false;
  int d = 0, N = 0, N8 = 0, nb8 = 0, nb32 = 0;
  KernelSpec kern;
  int rule_id = 0;
  double sigma_tol = 1e-8, sigma_n2 = 1e-6, k0 = 1, d2k0 = 0, ymin_base = 0;
  double *Xb = nullptr, *yb = nullptr, *c0 = nullptr, *u0 = nullptr, *Lf = nullptr, *Lb = nullptr, *Lbf = nullptr, *x0_batch = nullptr;
  // normals / starts
  int M = 0, hp1 = 0;
  double* rn = nullptr;
  int S = 0;
  double* starts = nullptr;
  unsigned* sobol_dirs = nullptr;
  // outputs
  int outM = 0, outh = -1, outS = 0, outd = 0, outnth = 0;
  double *values = nullptr, *grad_x = nullptr, *grad_theta = nullptr, *xs = nullptr, *ys = nullptr, *gys = nullptr, *alphas = nullptr;
  int *best_index = nullptr, *grad_case = nullptr, *status = nullptr, *n_evals = nullptr, *start_status = nullptr, *start_iters = nullptr;
  int* work_counter = nullptr;
  double* sums = nullptr;
  int sums_len = 0;
  double *dual_dirs = nullptr, *x_forced = nullptr, *cs_tape = nullptr, *gh_nodes = nullptr, *gh_weights = nullptr;
  size_t cap_ghn = 0, cap_ghw = 0;
  int gh_depth = 0, gh_M = 0;
  size_t dual_cap = 0, forced_cap = 0, tape_cap = 0;
  size_t cap_Xb = 0, cap_yb = 0, cap_c0 = 0, cap_u0 = 0, cap_Lf = 0, cap_Lb = 0, cap_Lbf = 0, cap_x0b = 0, cap_rn = 0, cap_starts = 0;
  // last call
  int last_h = 0, last_mode = 0, last_nth = 1;
  bool tape_enabled = true;
  double htol = 1e-4;
  int rs_cap = 0;           // development knob: cap on the row splits of the Gram reductions (0 = default)
  int vglob_wmax = 0;       // development knob: cap on the start slots of the large-n variant (0 = default)
  int force_vglob = 0;      // development / test knob: use the large-n variant even when shared memory would do
  double* Vscratch = nullptr, *Bscratch = nullptr;
  double *t_mu = nullptr, *t_sigma = nullptr, *t_dmu = nullptr, *t_dsigma = nullptr, *t_Halpha = nullptr;
  size_t cap_tex = 0;  // trajectories x steps the extended tape holds
  bool tex_valid = false;
  bool xs_valid = false;
  int* order = nullptr;     // longest-first order of the trajectories of the last launch (scheduling hint for the next one)
  size_t cap_order = 0;
  int order_M = 0;          // 0: no valid order
  int lpt = 1;              // RBO_TUNE_LPT
  unsigned long long* cta_done = nullptr;  // the x-path of the last rollout is on the device (RBO_FLAG_REPLAY_TAPE)
  // resident surrogate in its canonical device form (rbo_set_surrogate / rbo_condition): L0^-1 row-major with pitch ldi,
  // the observation sites point-major, work vectors
  double *Ld = nullptr, *Linv = nullptr, *Xpts = nullptr, *wk = nullptr;
  int* dstat = nullptr;
  size_t cap_Ld = 0, cap_Linv = 0, cap_Xpts = 0, cap_wk = 0;
  int ldi = 0, cap_rows = 0;
  size_t cap_Vscratch = 0, cap_Bscratch = 0;
};

static int fail(rbo_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

#define CK(h, call)                                                                                       \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) return fail(h, RBO_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

template <class T>
static cudaError_t dev_realloc(T** p, size_t n) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  if (n == 0) return cudaSuccess;
  return cudaMalloc((void**)p, n * sizeof(T));
}
// grows only: cudaFree / cudaMalloc synchronise the device and cost ~0.1 s on a busy context
template <class T>
static cudaError_t dev_reserve(T** p, size_t* cap, size_t n) {
  if (*p && *cap >= n) return cudaSuccess;
  cudaError_t e = dev_realloc(p, n);
  *cap = (e == cudaSuccess) ? n : 0;
  return e;
}

extern "C" {

int rbo_abi_version(void) { return RBO_ABI_VERSION; }

void rbo_default_solver_opts(rbo_solver_opts* o) {
  o->maxit = 100;
  o->maxtry = 30;
  o->gtol = 1e-10;
  o->xtol = 1e-15;
  o->pred_tol = 1e-13;
  o->eta = 0.1;
  o->delta0_box = 0.5;
  o->delta0_ell = 1.0;
  o->stol = 1e-5;
}

const char* rbo_last_error(const rbo_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int rbo_create(rbo_handle** out, int device_id) {
  if (!out) return fail(nullptr, RBO_ERR_ARG, "rbo_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, RBO_ERR_CUDA, "rbo_create: no CUDA device (%s); librbo has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device_id < 0 || device_id >= ndev) return fail(nullptr, RBO_ERR_ARG, "rbo_create: device %d out of range (have %d)", device_id, ndev);
  rbo_handle* h = new rbo_handle();
  h->device = device_id;
  rbo_default_solver_opts(&h->so);
  if (const char* e = getenv("RBO_FORCE_LARGE_N")) h->force_vglob = atoi(e) != 0;  // test knob: whole suites under the large-n variant
#define CKC(call)                                                                                    \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) { fail(nullptr, RBO_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); delete h; return RBO_ERR_CUDA; } \
  } while (0)
  CKC(cudaSetDevice(device_id));
  cudaDeviceProp prop;
  CKC(cudaGetDeviceProperties(&prop, device_id));
  if (prop.major < 10) { fail(nullptr, RBO_ERR_UNSUPPORTED, "rbo_create: device sm_%d%d is not Blackwell (sm_100a required)", prop.major, prop.minor); delete h; return RBO_ERR_UNSUPPORTED; }
  h->num_sms = prop.multiProcessorCount;
  h->max_smem = (int)prop.sharedMemPerBlockOptin;
  CKC(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  CKC(cudaEventCreate(&h->ev0));
  CKC(cudaEventCreate(&h->ev1));
  CKC(cudaMalloc((void**)&h->work_counter, 16 * sizeof(int)));  // [0] trajectory scheduler, [1..2] kernel watchdog flags
  CKC(cudaMalloc((void**)&h->cta_done, 1024 * sizeof(unsigned long long)));
  CKC(cudaMalloc((void**)&h->sobol_dirs, sizeof(rbo_sobol_dirs_host)));
  CKC(cudaMemcpy(h->sobol_dirs, rbo_sobol_dirs_host, sizeof(rbo_sobol_dirs_host), cudaMemcpyHostToDevice));
  CKC(cudaFuncSetAttribute(rbo_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
  CKC(cudaFuncSetAttribute(rbo_rollout_kernel_largen, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
#undef CKC
  *out = h;
  return RBO_SUCCESS;
}

int rbo_destroy(rbo_handle* h) {
  if (!h) return RBO_SUCCESS;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  void* ptrs[] = {h->Xb, h->yb, h->c0, h->u0, h->Lf, h->Lb, h->Lbf, h->x0_batch, h->rn, h->starts, h->sobol_dirs, h->values, h->grad_x, h->grad_theta, h->xs, h->ys,
                  h->gys, h->alphas, h->best_index, h->grad_case, h->status, h->n_evals, h->start_status, h->start_iters, h->work_counter, h->sums,
                  h->dual_dirs, h->x_forced, h->cs_tape, h->gh_nodes, h->gh_weights, h->Vscratch, h->Bscratch, h->Ld, h->Linv, h->Xpts, h->wk, h->dstat, h->t_mu, h->t_sigma, h->t_dmu, h->t_dsigma, h->t_Halpha, h->order, h->cta_done};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return RBO_SUCCESS;
}

int rbo_set_stream(rbo_handle* h, void* cuda_stream) {
  if (!h) return RBO_ERR_ARG;
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return RBO_SUCCESS;
}

int rbo_set_solver_opts(rbo_handle* h, const rbo_solver_opts* o) {
  if (!h || !o) return RBO_ERR_ARG;
  if (o->maxit < 1 || o->maxtry < 1 || !(o->delta0_box > 0.0) || !(o->delta0_ell > 0.0) || !(o->eta > 0.0 && o->eta < 0.25) || !(o->stol >= 0.0))
    return fail(h, RBO_ERR_ARG, "rbo_set_solver_opts: invalid options");
  h->so = *o;
  return RBO_SUCCESS;
}

int rbo_set_tuning(rbo_handle* h, int key, int value) {
  if (!h) return RBO_ERR_ARG;
  switch (key) {
    case RBO_TUNE_LARGE_N: h->force_vglob = value != 0; return RBO_SUCCESS;
    case RBO_TUNE_LARGE_N_SLOTS: if (value < 0 || value > RBO_NCONS) break; h->vglob_wmax = value; return RBO_SUCCESS;
    case RBO_TUNE_LPT: h->lpt = value != 0; h->order_M = 0; return RBO_SUCCESS;
    case RBO_TUNE_ROW_SPLITS: if (value < 0 || value > 4) break; h->rs_cap = value; return RBO_SUCCESS;
    default: break;
  }
  return fail(h, RBO_ERR_ARG, "rbo_set_tuning: unknown key %d or value %d out of range", key, value);
}

int rbo_set_htol(rbo_handle* h, double htol) {
  if (!h) return RBO_ERR_ARG;
  h->htol = htol;
  return RBO_SUCCESS;
}

// (Re)packs everything the rollout kernel streams from the canonical device form (Linv, Xpts): Xb, Lf, Lb, Lbf. Stream-ordered.
static int repack_surrogate(rbo_handle* h) {
  const int N = h->N, d = h->d, BR = RBO_BR, LP = RBO_LP;
  const int N8 = (N + RBO_PR - 1) / RBO_PR * RBO_PR, nb32 = (N + BR - 1) / BR;
  h->N8 = N8; h->nb8 = N8 / RBO_PR; h->nb32 = nb32;
  const size_t nchunks = (size_t)nb32 * (nb32 + 1) / 2, nLf = (size_t)LP * BR * nchunks, nLbf = (size_t)1024 * nchunks;
  CK(h, dev_reserve(&h->Xb, &h->cap_Xb, (size_t)d * N8));
  CK(h, dev_reserve(&h->Lf, &h->cap_Lf, nLf));
  CK(h, dev_reserve(&h->Lb, &h->cap_Lb, nLf));
  CK(h, dev_reserve(&h->Lbf, &h->cap_Lbf, nLbf));
  CK(h, cudaMemsetAsync(h->Lf, 0, nLf * 8, h->stream));  // the 4 pad doubles per k
  CK(h, cudaMemsetAsync(h->Lb, 0, nLf * 8, h->stream));
  rbo_layout_x_kernel<<<(d * N8 + 255) / 256, 256, 0, h->stream>>>(h->Xpts, d, N, N8, h->Xb);
  dim3 grid(8, nb32);
  rbo_pack_fwd_kernel<<<grid, 256, 0, h->stream>>>(h->Linv, h->ldi, nb32, h->Lf);
  rbo_pack_bwd_kernel<<<grid, 256, 0, h->stream>>>(h->Linv, h->ldi, nb32, h->Lb, h->Lbf);
  CK(h, cudaGetLastError());
  return RBO_SUCCESS;
}

// Capacity of the canonical form: rows_needed rows of L0^-1 (pitch = capacity, so that rbo_condition appends rows in place).
static int reserve_surrogate(rbo_handle* h, int d, int rows_needed, bool keep) {
  const int N32 = (rows_needed + RBO_BR - 1) / RBO_BR * RBO_BR;
  if (N32 <= h->cap_rows && h->Linv && (size_t)rows_needed * d <= h->cap_Xpts) return RBO_SUCCESS;
  const int cap = (rows_needed + 64 + RBO_BR - 1) / RBO_BR * RBO_BR;
  double *nLinv = nullptr, *nX = nullptr;
  CK(h, cudaMalloc((void**)&nLinv, (size_t)cap * cap * 8));
  if (cudaMalloc((void**)&nX, (size_t)cap * d * 8) != cudaSuccess) { cudaFree(nLinv); return fail(h, RBO_ERR_CUDA, "reserve_surrogate: out of device memory"); }
  cudaMemsetAsync(nLinv, 0, (size_t)cap * cap * 8, h->stream);
  if (keep && h->Linv) {
    const int rows = (h->N + RBO_BR - 1) / RBO_BR * RBO_BR;
    cudaMemcpy2DAsync(nLinv, (size_t)cap * 8, h->Linv, (size_t)h->ldi * 8, (size_t)rows * 8, rows, cudaMemcpyDeviceToDevice, h->stream);
    cudaMemcpyAsync(nX, h->Xpts, (size_t)h->N * d * 8, cudaMemcpyDeviceToDevice, h->stream);
  }
  cudaStreamSynchronize(h->stream);
  if (h->Linv) cudaFree(h->Linv);
  if (h->Xpts) cudaFree(h->Xpts);
  h->Linv = nLinv; h->Xpts = nX; h->ldi = cap; h->cap_rows = cap; h->cap_Xpts = (size_t)cap * d; h->cap_Linv = (size_t)cap * cap;
  // vectors that grow with the observation count
  double *ny = nullptr, *nc = nullptr, *nu = nullptr;
  if (cudaMalloc((void**)&ny, (size_t)cap * 8) != cudaSuccess || cudaMalloc((void**)&nc, (size_t)cap * 8) != cudaSuccess || cudaMalloc((void**)&nu, (size_t)cap * 8) != cudaSuccess)
    return fail(h, RBO_ERR_CUDA, "reserve_surrogate: out of device memory");
  cudaMemsetAsync(nc, 0, (size_t)cap * 8, h->stream); cudaMemsetAsync(nu, 0, (size_t)cap * 8, h->stream); cudaMemsetAsync(ny, 0, (size_t)cap * 8, h->stream);
  if (keep && h->yb) {
    cudaMemcpyAsync(ny, h->yb, (size_t)h->N * 8, cudaMemcpyDeviceToDevice, h->stream);
    cudaMemcpyAsync(nc, h->c0, (size_t)h->N * 8, cudaMemcpyDeviceToDevice, h->stream);
    cudaMemcpyAsync(nu, h->u0, (size_t)h->N * 8, cudaMemcpyDeviceToDevice, h->stream);
  }
  cudaStreamSynchronize(h->stream);
  if (h->yb) cudaFree(h->yb);
  if (h->c0) cudaFree(h->c0);
  if (h->u0) cudaFree(h->u0);
  h->yb = ny; h->c0 = nc; h->u0 = nu; h->cap_yb = h->cap_c0 = h->cap_u0 = (size_t)cap;
  CK(h, dev_reserve(&h->wk, &h->cap_wk, (size_t)4 * cap + 64));
  if (!h->dstat) CK(h, cudaMalloc((void**)&h->dstat, 4 * sizeof(int)));
  return RBO_SUCCESS;
}

int rbo_set_surrogate(rbo_handle* h, int d, int N, const double* X, int ldX, const double* L, int ldL, const double* y, const double* c,
                      double sigma_n2, int kernel_id, const double* ktheta, int nktheta, int rule_id, double sigma_tol) {
  if (!h) return RBO_ERR_ARG;
  if (d < 1 || d > RBO_MAXD) return fail(h, RBO_ERR_UNSUPPORTED, "rbo_set_surrogate: d = %d outside [1, %d]", d, RBO_MAXD);
  if (N < 1 || !X || !L || !y || !c || ldX < d || ldL < N) return fail(h, RBO_ERR_ARG, "rbo_set_surrogate: bad arguments");
  if (kernel_id < RBO_KERNEL_MATERN12 || kernel_id > RBO_KERNEL_PERIODIC) return fail(h, RBO_ERR_ARG, "rbo_set_surrogate: unknown kernel id %d", kernel_id);
  if (rule_id < RBO_RULE_EI || rule_id > RBO_RULE_LCB) return fail(h, RBO_ERR_ARG, "rbo_set_surrogate: unknown decision rule id %d", rule_id);
  if (nktheta < 1 || nktheta > 4 || !ktheta) return fail(h, RBO_ERR_ARG, "rbo_set_surrogate: kernel hyper-parameters missing");
  if ((size_t)N * 16 > 200 * 1024) return fail(h, RBO_ERR_UNSUPPORTED, "rbo_set_surrogate: N = %d observations exceed the inversion kernel's shared-memory column (12800)", N);
  CK(h, cudaSetDevice(h->device));
  if (d != h->d) {
    // buffers laid out for another input dimension are no longer valid (normals are M x (d+1) x hp1, starts d x S, ...)
    h->M = 0; h->hp1 = 0; h->S = 0; h->gh_M = 0; h->gh_depth = 0;
    if (h->rn) { cudaFree(h->rn); h->rn = nullptr; h->cap_rn = 0; }
    if (h->starts) { cudaFree(h->starts); h->starts = nullptr; h->cap_starts = 0; }
    if (h->gh_nodes) { cudaFree(h->gh_nodes); h->gh_nodes = nullptr; h->cap_ghn = 0; }
    if (h->gh_weights) { cudaFree(h->gh_weights); h->gh_weights = nullptr; h->cap_ghw = 0; }
    h->cap_rows = 0;  // Xpts is sized with d
  }
  h->have_sur = false;
  int rc = reserve_surrogate(h, d, N, false);
  if (rc) return rc;
  h->d = d; h->N = N;
  h->kern.id = kernel_id;
  for (int i = 0; i < 4; ++i) h->kern.th[i] = i < nktheta ? ktheta[i] : 0.0;
  h->rule_id = rule_id; h->sigma_tol = sigma_tol; h->sigma_n2 = sigma_n2;
  double a, b;
  kern_eval(h->kern, 0.0, h->k0, a, b);
  h->d2k0 = b;
  double ymin = y[0];
  for (int j = 0; j < N; ++j) ymin = std::min(ymin, y[j]);
  h->ymin_base = ymin;
  const int N32 = (N + RBO_BR - 1) / RBO_BR * RBO_BR;
  // H2D of what FantasySurrogate(s, h) copies (rbs.jl:345-381): X (d x N), the lower factor, y, c
  CK(h, dev_reserve(&h->Ld, &h->cap_Ld, (size_t)N * N));
  CK(h, cudaMemcpy2DAsync(h->Xpts, (size_t)d * 8, X, (size_t)ldX * 8, (size_t)d * 8, N, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpy2DAsync(h->Ld, (size_t)N * 8, L, (size_t)ldL * 8, (size_t)N * 8, N, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemsetAsync(h->c0, 0, h->cap_c0 * 8, h->stream));
  CK(h, cudaMemsetAsync(h->u0, 0, h->cap_u0 * 8, h->stream));
  CK(h, cudaMemcpyAsync(h->yb, y, (size_t)N * 8, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->c0, c, (size_t)N * 8, cudaMemcpyHostToDevice, h->stream));
  // Explicit inverse of the base factor on the device in double-double arithmetic (one rounding per stored entry): every later
  // "triangular solve" against L0 is then a dependency-free panel product on the FP64 tensor cores.
  CK(h, cudaMemsetAsync(h->Linv, 0, (size_t)N32 * h->ldi * 8, h->stream));
  {
    const int wpc = std::max(1, std::min(4, (int)((200 * 1024) / ((size_t)N * 16))));
    const size_t smem = (size_t)wpc * 2 * N * 8;
    CK(h, cudaFuncSetAttribute(rbo_trinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
    rbo_trinv_kernel<<<(N32 + wpc - 1) / wpc, 32 * wpc, smem, h->stream>>>(h->Ld, N, N32, h->Linv, h->ldi);
    CK(h, cudaGetLastError());
  }
  // u0 = L^-1 y (kept so that the coefficient re-solve of rbs.jl:422-429 only needs the backward half)
  rbo_matvec_lower_kernel<<<(N + 7) / 8, 256, 0, h->stream>>>(h->Linv, h->ldi, N, h->yb, h->u0);
  CK(h, cudaGetLastError());
  rc = repack_surrogate(h);
  if (rc) return rc;
  CK(h, cudaStreamSynchronize(h->stream));  // the caller's buffers may change after return
  h->have_sur = true;
  return RBO_SUCCESS;
}

int rbo_condition(rbo_handle* h, const double* x, double y) {
  if (!h || !x) return RBO_ERR_ARG;
  if (!h->have_sur) return fail(h, RBO_ERR_STATE, "rbo_condition: no surrogate (rbo_set_surrogate)");
  CK(h, cudaSetDevice(h->device));
  const int n = h->N, d = h->d;
  if ((size_t)(n + 1) * 16 > 200 * 1024) return fail(h, RBO_ERR_UNSUPPORTED, "rbo_condition: too many observations");
  int rc = reserve_surrogate(h, d, n + 1, true);
  if (rc) return rc;
  double* kv = h->wk; double* lv = kv + h->cap_rows; double* tmp = lv + h->cap_rows; double* scal = tmp + h->cap_rows; double* xd = scal + 8;
  CK(h, cudaMemcpyAsync(xd, x, (size_t)d * 8, cudaMemcpyHostToDevice, h->stream));
  if (n % RBO_BR == 0) {
    // the factor grows by a panel: its padding rows are the identity
    std::vector<double> one(1, 1.0);
    CK(h, cudaMemsetAsync(h->Linv + (size_t)n * h->ldi, 0, (size_t)RBO_BR * h->ldi * 8, h->stream));
    for (int r = n + 1; r < n + RBO_BR; ++r) CK(h, cudaMemcpyAsync(h->Linv + (size_t)r * h->ldi + r, one.data(), 8, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
  }
  // update_covariance! (rbs.jl:166-183): k = psi(|x - X_j|); update_cholesky! (rbs.jl:185-203): l = L^-1 k, l_nn = sqrt(k0 + sigma_n2 - l.l)
  rbo_kvec_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(h->Xpts, d, n, xd, h->kern, kv);
  rbo_matvec_lower_kernel<<<(n + 7) / 8, 256, 0, h->stream>>>(h->Linv, h->ldi, n, kv, lv);
  rbo_dots_kernel<<<1, 32, 0, h->stream>>>(lv, h->u0, n, scal);
  rbo_matvec_lower_t_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(h->Linv, h->ldi, n, lv, 1.0, tmp);
  rbo_append_row_kernel<<<(n + 256) / 256, 256, 0, h->stream>>>(h->Linv, h->ldi, n, tmp, scal, h->k0 + h->sigma_n2, y, h->u0, h->dstat);
  CK(h, cudaGetLastError());
  int st = 0;
  CK(h, cudaMemcpyAsync(&st, h->dstat, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  if (st != 0) return fail(h, RBO_ERR_NUMERIC, "rbo_condition: PosDefException -- the extended kernel matrix is not positive definite (rbs.jl:196)");
  CK(h, cudaMemcpyAsync(h->Xpts + (size_t)n * d, xd, (size_t)d * 8, cudaMemcpyDeviceToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->yb + n, &y, 8, cudaMemcpyHostToDevice, h->stream));
  h->N = n + 1;
  h->ymin_base = std::min(h->ymin_base, y);
  // update_coefficients! (rbs.jl:205-212): c = L^-T (L^-1 y), a full re-solve in the reference as well
  rbo_matvec_lower_t_kernel<<<(n + 1 + 127) / 128, 128, 0, h->stream>>>(h->Linv, h->ldi, n + 1, h->u0, 1.0, h->c0);
  CK(h, cudaGetLastError());
  rc = repack_surrogate(h);
  if (rc) return rc;
  CK(h, cudaStreamSynchronize(h->stream));
  return RBO_SUCCESS;
}

int rbo_get_surrogate(rbo_handle* h, int* N_out, double* X, int ldX, double* y, double* c) {
  if (!h) return RBO_ERR_ARG;
  if (!h->have_sur) return fail(h, RBO_ERR_STATE, "rbo_get_surrogate: no surrogate");
  CK(h, cudaSetDevice(h->device));
  if (N_out) *N_out = h->N;
  if (X) { if (ldX < h->d) return fail(h, RBO_ERR_ARG, "rbo_get_surrogate: ldX < d"); CK(h, cudaMemcpy2DAsync(X, (size_t)ldX * 8, h->Xpts, (size_t)h->d * 8, (size_t)h->d * 8, h->N, cudaMemcpyDeviceToHost, h->stream)); }
  if (y) CK(h, cudaMemcpyAsync(y, h->yb, (size_t)h->N * 8, cudaMemcpyDeviceToHost, h->stream));
  if (c) CK(h, cudaMemcpyAsync(c, h->c0, (size_t)h->N * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return RBO_SUCCESS;
}

int rbo_set_normals(rbo_handle* h, const double* rn, int M_total, int hp1, int m_begin, int m_count) {
  if (!h) return RBO_ERR_ARG;
  if (!h->have_sur) return fail(h, RBO_ERR_STATE, "rbo_set_normals: call rbo_set_surrogate first");
  if (!rn || M_total < 1 || hp1 < 1 || m_begin < 0 || m_count < 1 || m_begin + m_count > M_total) return fail(h, RBO_ERR_ARG, "rbo_set_normals: bad arguments");
  CK(h, cudaSetDevice(h->device));
  const int q1 = h->d + 1;
  CK(h, dev_reserve(&h->rn, &h->cap_rn, (size_t)m_count * q1 * hp1));
  // strided slice [m_begin, m_begin+m_count) of the sample-fastest tensor
  CK(h, cudaMemcpy2DAsync(h->rn, (size_t)m_count * 8, rn + m_begin, (size_t)M_total * 8, (size_t)m_count * 8, (size_t)q1 * hp1, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  h->M = m_count; h->hp1 = hp1; h->order_M = 0;
  return RBO_SUCCESS;
}

int rbo_generate_normals(rbo_handle* h, int M_total, int hp1, int m_begin, int m_count) {
  if (!h) return RBO_ERR_ARG;
  if (!h->have_sur) return fail(h, RBO_ERR_STATE, "rbo_generate_normals: call rbo_set_surrogate first");
  if (M_total < 1 || hp1 < 1 || m_begin < 0 || m_count < 1 || m_begin + m_count > M_total) return fail(h, RBO_ERR_ARG, "rbo_generate_normals: bad arguments");
  const int q1 = h->d + 1, D = q1 + (q1 & 1);
  if (D > RBO_SOBOL_MAXDIM) return fail(h, RBO_ERR_UNSUPPORTED, "rbo_generate_normals: %d Sobol dimensions > %d", D, RBO_SOBOL_MAXDIM);
  if ((double)M_total * hp1 >= 4294967295.0) return fail(h, RBO_ERR_UNSUPPORTED, "rbo_generate_normals: more than 2^32 Sobol points");
  CK(h, cudaSetDevice(h->device));
  CK(h, dev_reserve(&h->rn, &h->cap_rn, (size_t)m_count * q1 * hp1));
  size_t total = (size_t)m_count * q1 * hp1;
  int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 8);
  rbo_normals_kernel<<<blocks, 256, 0, h->stream>>>(h->sobol_dirs, h->rn, M_total, h->d, hp1, m_begin, m_count);
  CK(h, cudaGetLastError());
  h->M = m_count; h->hp1 = hp1; h->order_M = 0;
  return RBO_SUCCESS;
}

int rbo_get_normals(rbo_handle* h, double* out) {
  if (!h || !out) return RBO_ERR_ARG;
  if (!h->rn) return fail(h, RBO_ERR_STATE, "rbo_get_normals: no normals resident");
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaMemcpyAsync(out, h->rn, (size_t)h->M * (h->d + 1) * h->hp1 * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return RBO_SUCCESS;
}

int rbo_set_quadrature(rbo_handle* h, const double* nodes, const double* weights, int depth, int m_count) {
  if (!h) return RBO_ERR_ARG;
  if (!nodes || !weights || depth < 1 || m_count < 1) return fail(h, RBO_ERR_ARG, "rbo_set_quadrature: bad arguments");
  CK(h, cudaSetDevice(h->device));
  const size_t n = (size_t)depth * m_count;
  CK(h, dev_reserve(&h->gh_nodes, &h->cap_ghn, n));
  CK(h, dev_reserve(&h->gh_weights, &h->cap_ghw, n));
  CK(h, cudaMemcpyAsync(h->gh_nodes, nodes, n * 8, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->gh_weights, weights, n * 8, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  h->gh_depth = depth; h->gh_M = m_count;
  return RBO_SUCCESS;
}

int rbo_set_starts(rbo_handle* h, const double* starts, int S) {
  if (!h) return RBO_ERR_ARG;
  if (!h->have_sur) return fail(h, RBO_ERR_STATE, "rbo_set_starts: call rbo_set_surrogate first");
  if (!starts || S < 1) return fail(h, RBO_ERR_ARG, "rbo_set_starts: bad arguments");
  if (S > 65535) return fail(h, RBO_ERR_UNSUPPORTED, "rbo_set_starts: %d starts (the kernel keeps per-start counters and the hand-out order as 16-bit values: at most 65535)", S);
  CK(h, cudaSetDevice(h->device));
  CK(h, dev_reserve(&h->starts, &h->cap_starts, (size_t)S * h->d));
  CK(h, cudaMemcpyAsync(h->starts, starts, (size_t)S * h->d * 8, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  h->S = S;
  return RBO_SUCCESS;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
static int ensure_outputs(rbo_handle* h, int M, int hor, int S, int d, int nth) {
  if (h->outM == M && h->outh == hor && h->outS == S && h->outd == d && h->outnth == nth) return RBO_SUCCESS;
  const int hh = std::max(hor, 1);
  h->outM = 0; h->outh = -1; h->xs_valid = false;  // the cache key is only valid once every buffer below exists
  CK(h, dev_realloc(&h->values, (size_t)M));
  CK(h, dev_realloc(&h->grad_x, (size_t)M * d));
  CK(h, dev_realloc(&h->grad_theta, (size_t)M * nth));
  CK(h, dev_realloc(&h->best_index, (size_t)M));
  CK(h, dev_realloc(&h->grad_case, (size_t)M));
  CK(h, dev_realloc(&h->status, (size_t)M));
  CK(h, dev_realloc(&h->xs, (size_t)M * (hor + 1) * d));
  CK(h, dev_realloc(&h->ys, (size_t)M * (hor + 1)));
  CK(h, dev_realloc(&h->gys, (size_t)M * (hor + 1) * d));
  CK(h, dev_realloc(&h->alphas, (size_t)M * hh));
  CK(h, dev_realloc(&h->n_evals, (size_t)M * hh));
  CK(h, dev_realloc(&h->start_status, (size_t)M * hh * S));
  CK(h, dev_realloc(&h->start_iters, (size_t)M * hh * S));
  h->sums_len = 1 + 3 * (1 + d + nth) + hh + (hor + 2) + 2;
  CK(h, dev_realloc(&h->sums, (size_t)h->sums_len + 1 + 3 * (1 + d + nth) + 2));  // + the gathered partial-sums vector (rbo_partial_sums_host)
  h->outM = M; h->outh = hor; h->outS = S; h->outd = d; h->outnth = nth;
  return RBO_SUCCESS;
}

// Chooses the number of start slots W (starts evaluated in lock-step) so that the shared-memory plan fits.
struct PlanChoice { int W, RP, NR, RSmax, NPmax, xsm; size_t bytes; int RSh; int vglob; };
static bool choose_plan(const rbo_handle* h, int hor, int S, int mode, PlanChoice* pc) {
  const int d = h->d, N8 = h->N8, CS = d + 3, NR = std::max(N8 + RBO_MAXFAN, h->nb32 * RBO_BR);
  const int nadj = (mode == RBO_MODE_VALUE_GRAD) ? ncols_adjoint(d) : 0;  // the adjoint's column plan is only needed with gradients
  // Preference: (i) row splits >= 2 and the base locations in shared memory, with the largest W that still allows it
  // (provided that W covers at least half of the start list); (ii) otherwise the largest W with whatever fits.
  auto try_plan = [&](int W, int RSmax, int xsm, int vglob = 0) {
    int RP = std::max(W * CS, nadj) + 2;
    if ((RP & 1) == 0) RP += 1;  // odd pitch: conflict-free column walks
    int NPmax = npairs_max(d, W);
    SmemPlan pl = make_plan(d, N8, hor, W, RP, NR, RSmax, NPmax, xsm, RSmax, vglob, S);
    size_t bytes = (size_t)pl.total * 8;
    if (bytes > (size_t)h->max_smem) return false;
    *pc = {W, RP, NR, RSmax, NPmax, xsm, bytes, RSmax, vglob};
    // the Hessian sums are tensor-pipe bound per scheduler: give them enough row splits to occupy every warp if that still fits
    const int want = std::min(4, std::max(RSmax, RBO_NWARPS / W));
    for (int rsh = want; rsh > RSmax; --rsh) {
      SmemPlan p2 = make_plan(d, N8, hor, W, RP, NR, RSmax, NPmax, xsm, rsh, vglob, S);
      if ((size_t)p2.total * 8 <= (size_t)h->max_smem) { pc->RSh = rsh; pc->bytes = (size_t)p2.total * 8; break; }
    }
    return true;
  };
  const int Wmax = std::min(S, RBO_NCONS);
  for (int W = Wmax; !h->force_vglob && W >= std::max(1, std::min(Wmax, (S + 1) / 2)); --W) {
    int rs = std::min(h->rs_cap > 0 ? h->rs_cap : 4, std::min(4, std::max(2, RBO_NWARPS / W)));
    if (try_plan(W, rs, 1)) return true;
    if (rs > 2 && try_plan(W, 2, 1)) return true;
  }
  for (int W = Wmax; !h->force_vglob && W >= 1; --W) {
    const int tries[4][2] = {{2, 1}, {1, 1}, {2, 0}, {1, 0}};
    for (auto& t : tries) if (try_plan(W, t[0], t[1])) return true;
  }
  // Large-n variant (north_star: "L0 ... read through L2 when n is large"; BASELINE config C5): the work matrix moves to a
  // per-CTA scratch in global memory (it stays L2-resident), everything else keeps its place in shared memory. The base
  // locations are then read through L1 as well. More start slots per round amortise the stream of L0^-1 over more columns.
  // measured at the C5 shape (n = 1000, d = 20): 4-5 slots are best; more slots enlarge the scratch (148 CTAs x NR x RP doubles)
  // beyond what stays L2-resident and leave fewer row splits for the reductions
  const int Wg = std::min(Wmax, h->vglob_wmax > 0 ? h->vglob_wmax : 5);
  for (int W = Wg; W >= 1; --W)
    for (int rs = std::min(4, std::max(2, RBO_NWARPS / W)); rs >= 1; --rs)
      if (try_plan(W, rs, 0, 1)) return true;
  return false;
}

static double f_eval(double n, double d) { return 2 * n * n * (d + 1) + n * (4 * d * d + 11 * d + 31); }
static double f_step(double n, double d) { return 2 * n * n * (d + 1) + 2 * n * (d + 1) * (d + 1) + 3 * n * n + (d + 1) * (d + 1) * (d + 1) / 3.0; }
// what this implementation executes per evaluation: (d+1) forward + 1 backward substitutions, Gram + two Hessian sums
static double f_eval_exec(double n, double d) { return n * n * (d + 2) + n * (3 * d * (d + 1) + 4 * d + 40); }

static int launch_rollout(rbo_handle* h, const double* x0, const double* theta, int ntheta, const double* lbs, const double* ubs, int horizon,
                          double fmini, int mode, int flags, const double* dual_dirs_dev, const double* x_forced_dev, bool want_summary,
                          rbo_summary* summary, const double* x0_batch_dev = nullptr, int B = 1) {
  const bool myopic = (flags & RBO_FLAG_MYOPIC_INTERNAL) != 0, ghq = (flags & RBO_FLAG_GAUSS_HERMITE) != 0;
  if (!h->have_sur) return fail(h, RBO_ERR_STATE, "rbo_rollout: no surrogate (rbo_set_surrogate)");
  if (ghq && !h->gh_nodes) return fail(h, RBO_ERR_STATE, "rbo_rollout: no quadrature data (rbo_set_quadrature)");
  if (!h->rn && !myopic && !ghq) return fail(h, RBO_ERR_STATE, "rbo_rollout: no normals (rbo_set_normals / rbo_generate_normals)");
  if (!h->starts && !(flags & RBO_FLAG_TEACHER_FORCED)) return fail(h, RBO_ERR_STATE, "rbo_rollout: no starts (rbo_set_starts)");
  if (!x0 || !theta || !lbs || !ubs || ntheta < 1) return fail(h, RBO_ERR_ARG, "rbo_rollout: bad arguments");
  if (horizon < 0 || horizon + 1 > RBO_MAXFAN) return fail(h, RBO_ERR_UNSUPPORTED, "rbo_rollout: horizon %d not in [0, %d]", horizon, RBO_MAXFAN - 1);
  if (!myopic && !ghq && horizon + 1 > h->hp1) return fail(h, RBO_ERR_ARG, "rbo_rollout: normals hold %d steps, horizon + 1 = %d needed", h->hp1, horizon + 1);
  if (ghq && horizon + 1 > h->gh_depth) return fail(h, RBO_ERR_ARG, "rbo_rollout: quadrature depth %d < horizon + 1 = %d", h->gh_depth, horizon + 1);
  const int Ms = myopic ? 1 : (ghq ? h->gh_M : h->M);  // sample indices
  const int M = Ms * std::max(B, 1);                    // trajectories of this launch
  if (flags & RBO_FLAG_REPLAY_TAPE) {
    // second phase of the two-phase call: replay the x-path of the previous rollout of this handle (still on the device)
    if (myopic || h->outh != horizon || h->outM != M || !h->xs_valid) return fail(h, RBO_ERR_STATE, "rbo_rollout: RBO_FLAG_REPLAY_TAPE needs a preceding rollout with the same samples and horizon on this handle");
    flags |= RBO_FLAG_TEACHER_FORCED;
  } else if ((flags & RBO_FLAG_TEACHER_FORCED) && !x_forced_dev) return fail(h, RBO_ERR_ARG, "rbo_rollout: teacher forcing without x_forced");
  if (mode != RBO_MODE_VALUE && mode != RBO_MODE_VALUE_GRAD) return fail(h, RBO_ERR_ARG, "rbo_rollout: bad mode");
  for (int a = 0; a < h->d; ++a)
    if (!(lbs[a] <= ubs[a])) return fail(h, RBO_ERR_ARG, "rbo_rollout: lower bound above upper bound in dimension %d", a);
  CK(h, cudaSetDevice(h->device));
  const int S = std::max(h->S, 1);
  PlanChoice pc;
  if (!choose_plan(h, horizon, S, mode, &pc))
    return fail(h, RBO_ERR_UNSUPPORTED, "rbo_rollout: problem (d=%d, N=%d, h=%d) needs more than %d bytes of shared memory per CTA", h->d, h->N, horizon, h->max_smem);
  if (getenv("RBO_DEBUG")) fprintf(stderr, "[rbo] plan: W=%d RP=%d NR=%d RSmax=%d NPmax=%d xsm=%d smem=%zu B (limit %d)\n", pc.W, pc.RP, pc.NR, pc.RSmax, pc.NPmax, pc.xsm, pc.bytes, h->max_smem);
  if (getenv("RBO_DEBUG")) fprintf(stderr, "[rbo] plan: RSh=%d vglob=%d\n", pc.RSh, pc.vglob);
  int rc = ensure_outputs(h, M, horizon, S, h->d, ntheta);
  if (rc) return rc;
  DevProblem P;
  memset(&P, 0, sizeof(P));
  P.d = h->d; P.N = h->N; P.N8 = h->N8; P.nb8 = h->nb8; P.nb32 = h->nb32; P.h = horizon; P.S = S; P.W = pc.W; P.RSmax = pc.RSmax; P.NPmax = pc.NPmax; P.xsm = pc.xsm; P.XP = h->N8 + 1;
  P.RSh = pc.RSh;
  P.pl = make_plan(P.d, P.N8, horizon, pc.W, pc.RP, pc.NR, pc.RSmax, pc.NPmax, pc.xsm, pc.RSh, pc.vglob, S);
  P.vglob = pc.vglob;
  P.CS = h->d + 3; P.RP = pc.RP; P.NR = pc.NR; P.M = M; P.Ms = Ms; P.B = std::max(B, 1); P.x0_batch = (B > 1 || x0_batch_dev) ? x0_batch_dev : nullptr; P.hp1 = h->hp1; P.mode = mode; P.flags = flags; P.ntheta = ntheta;
  P.kern = h->kern; P.rule_id = h->rule_id; P.sigma_tol = h->sigma_tol; P.sigma_n2 = h->sigma_n2; P.k0 = h->k0; P.d2k0 = h->d2k0;
  P.ymin_base = h->ymin_base; P.m52_c = std::sqrt(5.0) / h->kern.th[0]; P.fmini = fmini; P.theta1 = theta[0]; P.htol = h->htol; P.so = h->so;
  for (int a = 0; a < h->d; ++a) { P.x0[a] = x0[a]; P.lbs[a] = lbs[a]; P.ubs[a] = ubs[a]; }
  P.Xb = h->Xb; P.yb = h->yb; P.c0 = h->c0; P.u0 = h->u0; P.Lf = h->Lf; P.Lb = h->Lb; P.Lbf = h->Lbf; P.rn = h->rn; P.starts = h->starts;
  P.dual_dirs = dual_dirs_dev; P.x_forced = x_forced_dev;
  P.gh_nodes = h->gh_nodes; P.gh_weights = h->gh_weights; P.gh_depth = h->gh_depth;
  P.values = h->values; P.grad_x = h->grad_x; P.grad_theta = h->grad_theta; P.best_index = h->best_index; P.grad_case = h->grad_case;
  P.status = h->status; P.xs = h->xs; P.ys = h->ys; P.gys = h->gys; P.alphas = h->alphas; P.n_evals = h->n_evals;
  P.start_status = h->tape_enabled ? h->start_status : nullptr; P.start_iters = h->tape_enabled ? h->start_iters : nullptr;
  P.work_counter = h->work_counter;
  h->tex_valid = false;
  if (flags & RBO_FLAG_TAPE_EX) {
    if (myopic) return fail(h, RBO_ERR_ARG, "rbo_rollout: RBO_FLAG_TAPE_EX does not apply to the myopic solve");
    const size_t n = (size_t)M * std::max(horizon, 1), dd = (size_t)h->d * h->d;
    if (h->cap_tex < n * dd) {
      CK(h, dev_realloc(&h->t_mu, n)); CK(h, dev_realloc(&h->t_sigma, n)); CK(h, dev_realloc(&h->t_dmu, n * h->d)); CK(h, dev_realloc(&h->t_dsigma, n * h->d));
      CK(h, dev_realloc(&h->t_Halpha, n * dd));
      h->cap_tex = n * dd;
    }
    P.t_mu = h->t_mu; P.t_sigma = h->t_sigma; P.t_dmu = h->t_dmu; P.t_dsigma = h->t_dsigma; P.t_Halpha = h->t_Halpha;
    h->tex_valid = true;
  }
  const int grid = std::min(M, h->num_sms);
  // longest-first hand-out when the previous launch ran the same trajectories (same normals, nearby x0): only worth it beyond two waves
  P.order = (h->lpt && !myopic && h->order_M == M && M > 2 * grid && !(flags & RBO_FLAG_TEACHER_FORCED)) ? h->order : nullptr;
  P.cta_done = grid <= 1024 ? h->cta_done : nullptr;
  {
    size_t need = (size_t)grid * (horizon + 2) * pc.NR;
    if (h->tape_cap < need) { CK(h, dev_realloc(&h->cs_tape, need)); h->tape_cap = need; }
    P.cs_tape = h->cs_tape;
    if (pc.vglob) {
      CK(h, dev_reserve(&h->Vscratch, &h->cap_Vscratch, (size_t)grid * pc.NR * pc.RP));
      P.Vscratch = h->Vscratch;
      P.bscratch_len = (size_t)h->N8 * 8 * ((std::max(h->d + 1, pc.W) + 7) / 8);  // widest backward pass: q1 columns (adjoint) or W (inner solve)
      CK(h, dev_reserve(&h->Bscratch, &h->cap_Bscratch, (size_t)grid * P.bscratch_len));
      P.Bscratch = h->Bscratch;
    }
  }
  CK(h, cudaMemsetAsync(h->work_counter, 0, 16 * sizeof(int), h->stream));
  CK(h, cudaEventRecord(h->ev0, h->stream));
  if (pc.vglob) rbo_rollout_kernel_largen<<<grid, RBO_THREADS, pc.bytes, h->stream>>>(P);
  else rbo_rollout_kernel<<<grid, RBO_THREADS, pc.bytes, h->stream>>>(P);
  CK(h, cudaGetLastError());
  rbo_stats_kernel<<<1, 1024, 0, h->stream>>>(h->values, mode == RBO_MODE_VALUE_GRAD ? h->grad_x : nullptr, mode == RBO_MODE_VALUE_GRAD ? h->grad_theta : nullptr,
                                              h->n_evals, h->best_index, h->grad_case, h->status, M, h->d, ntheta, horizon, h->sums);
  CK(h, cudaGetLastError());
  CK(h, cudaEventRecord(h->ev1, h->stream));
  if (h->lpt && !myopic && horizon > 0 && M > 2 * grid && !(flags & RBO_FLAG_TEACHER_FORCED)) {
    CK(h, dev_reserve(&h->order, &h->cap_order, (size_t)M));
    rbo_lpt_order_kernel<<<1, 1024, 0, h->stream>>>(h->n_evals, h->grad_case, h->best_index, M, std::max(horizon, 1), h->order);
    CK(h, cudaGetLastError());
    h->order_M = M;
  } else if (!(flags & RBO_FLAG_TEACHER_FORCED)) h->order_M = 0;
  h->last_h = horizon; h->last_mode = mode; h->last_nth = ntheta;
  h->xs_valid = !myopic;
  if (want_summary && summary) {
    std::vector<double> sums(h->sums_len);
    int wd[16] = {0};
    CK(h, cudaMemcpyAsync(sums.data(), h->sums, (size_t)h->sums_len * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(wd, h->work_counter, sizeof(wd), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (wd[1] || wd[2])
      return fail(h, RBO_ERR_CUDA, "rollout kernel watchdog: %s%s (wait kind %d, chunk %d, warp %d, consumers %d, block %d)", wd[1] ? "[panel pipeline wait never completed] " : "",
                  wd[2] ? "[inner solve exceeded its evaluation bound]" : "", wd[4], wd[5], wd[6], wd[7], wd[8]);
    float ms = 0;
    CK(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    const int d = h->d, hh = std::max(horizon, 1), nrows = 1 + d + ntheta;
    const double n = sums[0];
    summary->n_traj = M;
    summary->mean = n > 0 ? sums[1] / n : NAN;
    summary->std = n > 1 ? std::sqrt(sums[2] / (n - 1)) : NAN;
    summary->kernel_ms = ms;
    summary->gpu_launches = 2 + (h->order_M == M ? 1 : 0);
    summary->tail_ms = 0.0;
    if (P.cta_done) {
      std::vector<unsigned long long> td(grid);
      CK(h, cudaMemcpy(td.data(), h->cta_done, (size_t)grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      std::sort(td.begin(), td.end());
      summary->tail_ms = (double)(td.back() - td[grid / 2]) * 1e-6;  // last CTA to finish minus the median CTA
    }
    const double* ev = sums.data() + 1 + 3 * nrows;
    const double* hist = ev + hh;
    summary->n_failed = (int)hist[horizon + 2];
    double fl = 0, fx = 0;
    long long ne = 0;
    if (myopic) {
      ne = (long long)ev[0];
      fl = ev[0] * f_eval(h->N, d);
      fx = ev[0] * f_eval_exec(h->N, d);
    } else {
      fl = M * f_step(h->N, d); fx = M * f_step(h->N, d) * 0.5;
      for (int j = 1; j <= horizon; ++j) {
        double nj = h->N + j, e = ev[j - 1];
        ne += (long long)e;
        fl += e * f_eval(nj, d) + M * f_step(nj, d);
        fx += e * f_eval_exec(nj, d) + M * f_step(nj, d) * 0.5;
      }
    }
    if (mode == RBO_MODE_VALUE_GRAD)
      for (int t = 1; t <= horizon; ++t) {  // F_adj(t) for the hist[t] trajectories whose best step is t (case 3)
        double fa = t * (2.0 / 3.0) * d * d * d;
        for (int i = 1; i <= t; ++i) { double ni = h->N + i; fa += f_eval(ni, d) + i * (d + 1) * (2 * ni * ni + 6 * ni * d); }
        fl += hist[t] * fa;
        fx += hist[t] * fa;
      }
    summary->flops = fl;
    summary->flops_executed = fx;
    summary->n_evals = ne;
  }
  return RBO_SUCCESS;
}

static int upload_opt(rbo_handle* h, const double* src, size_t n, double** dst, size_t* cap) {
  if (!src) return RBO_SUCCESS;
  if (*cap < n) { CK(h, dev_realloc(dst, n)); *cap = n; }
  CK(h, cudaMemcpyAsync(*dst, src, n * 8, cudaMemcpyHostToDevice, h->stream));
  return RBO_SUCCESS;
}

extern "C" {

int rbo_rollout(rbo_handle* h, const double* x0, const double* theta, int ntheta, const double* lbs, const double* ubs, int horizon, double fmini,
                int mode, int flags, const double* dual_dirs, const double* x_forced, double* values, double* grad_x, double* grad_theta,
                int32_t* best_index, int32_t* grad_case, int32_t* status, rbo_summary* summary) {
  if (!h) return RBO_ERR_ARG;
  if (!values) return fail(h, RBO_ERR_ARG, "rbo_rollout: values is NULL");
  if (mode == RBO_MODE_VALUE_GRAD && (!grad_x || !grad_theta)) return fail(h, RBO_ERR_ARG, "rbo_rollout: gradient containers missing in VALUE_GRAD mode");
  CK(h, cudaSetDevice(h->device));
  const size_t Mrun = (flags & RBO_FLAG_GAUSS_HERMITE) ? (size_t)h->gh_M : (size_t)h->M;
  const size_t nd = Mrun * std::max(horizon, 0) * h->d;
  int rc = upload_opt(h, (mode == RBO_MODE_VALUE_GRAD) ? dual_dirs : nullptr, nd, &h->dual_dirs, &h->dual_cap);
  if (rc) return rc;
  rc = upload_opt(h, ((flags & RBO_FLAG_TEACHER_FORCED) && !(flags & RBO_FLAG_REPLAY_TAPE)) ? x_forced : nullptr, nd, &h->x_forced, &h->forced_cap);
  if (rc) return rc;
  rbo_summary local;
  rc = launch_rollout(h, x0, theta, ntheta, lbs, ubs, horizon, fmini, mode, flags, (mode == RBO_MODE_VALUE_GRAD && dual_dirs) ? h->dual_dirs : nullptr,
                      ((flags & RBO_FLAG_TEACHER_FORCED) && !(flags & RBO_FLAG_REPLAY_TAPE)) ? h->x_forced : nullptr, true, summary ? summary : &local);
  if (rc) return rc;
  const size_t M = Mrun;
  CK(h, cudaMemcpyAsync(values, h->values, M * 8, cudaMemcpyDeviceToHost, h->stream));
  if (mode == RBO_MODE_VALUE_GRAD) {
    CK(h, cudaMemcpyAsync(grad_x, h->grad_x, M * h->d * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(grad_theta, h->grad_theta, M * ntheta * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  if (best_index) CK(h, cudaMemcpyAsync(best_index, h->best_index, M * 4, cudaMemcpyDeviceToHost, h->stream));
  if (grad_case) CK(h, cudaMemcpyAsync(grad_case, h->grad_case, M * 4, cudaMemcpyDeviceToHost, h->stream));
  if (status) CK(h, cudaMemcpyAsync(status, h->status, M * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return RBO_SUCCESS;
}

int rbo_rollout_device(rbo_handle* h, const double* x0, const double* theta, int ntheta, const double* lbs, const double* ubs, int horizon,
                       double fmini, int mode, int flags, const double* dual_dirs_device, const double* x_forced_device, rbo_summary* summary) {
  if (!h) return RBO_ERR_ARG;
  return launch_rollout(h, x0, theta, ntheta, lbs, ubs, horizon, fmini, mode, flags, dual_dirs_device, x_forced_device, summary != nullptr, summary);
}

int rbo_rollout_batch(rbo_handle* h, const double* x0s, int n_x0, const double* theta, int ntheta, const double* lbs, const double* ubs, int horizon,
                      double fmini, int mode, const double* dual_dirs, double* values, double* grad_x, double* grad_theta, int32_t* status,
                      rbo_summary* summary) {
  if (!h) return RBO_ERR_ARG;
  if (!x0s || n_x0 < 1 || !values) return fail(h, RBO_ERR_ARG, "rbo_rollout_batch: bad arguments");
  if (mode == RBO_MODE_VALUE_GRAD && (!grad_x || !grad_theta)) return fail(h, RBO_ERR_ARG, "rbo_rollout_batch: gradient containers missing in VALUE_GRAD mode");
  if (!h->rn) return fail(h, RBO_ERR_STATE, "rbo_rollout_batch: no normals (rbo_set_normals / rbo_generate_normals)");
  CK(h, cudaSetDevice(h->device));
  const size_t Ms = (size_t)h->M, M = Ms * (size_t)n_x0, nd = Ms * std::max(horizon, 0) * h->d;
  int rc = upload_opt(h, (mode == RBO_MODE_VALUE_GRAD) ? dual_dirs : nullptr, nd, &h->dual_dirs, &h->dual_cap);
  if (rc) return rc;
  CK(h, dev_reserve(&h->x0_batch, &h->cap_x0b, (size_t)n_x0 * h->d));
  CK(h, cudaMemcpyAsync(h->x0_batch, x0s, (size_t)n_x0 * h->d * 8, cudaMemcpyHostToDevice, h->stream));
  rbo_summary local;
  rc = launch_rollout(h, x0s, theta, ntheta, lbs, ubs, horizon, fmini, mode, 0, (mode == RBO_MODE_VALUE_GRAD && dual_dirs) ? h->dual_dirs : nullptr, nullptr,
                      true, summary ? summary : &local, h->x0_batch, n_x0);
  if (rc) return rc;
  CK(h, cudaMemcpyAsync(values, h->values, M * 8, cudaMemcpyDeviceToHost, h->stream));
  if (mode == RBO_MODE_VALUE_GRAD) {
    CK(h, cudaMemcpyAsync(grad_x, h->grad_x, M * h->d * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(grad_theta, h->grad_theta, M * ntheta * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  if (status) CK(h, cudaMemcpyAsync(status, h->status, M * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return RBO_SUCCESS;
}

int rbo_partial_sums_device(rbo_handle* h, double* sums_device, int len) {
  if (!h || !sums_device) return RBO_ERR_ARG;
  const int need = 1 + 3 * (1 + h->outd + h->outnth);
  if (!h->sums || len < need + 2) return fail(h, RBO_ERR_ARG, "rbo_partial_sums_device: need %d doubles (1 + 3 (1 + d + ntheta) + 2)", need + 2);
  CK(h, cudaSetDevice(h->device));
  const int hh = std::max(h->outh, 1);
  const int idx_failed = need + hh + (h->outh + 2);  // sums layout of rbo_stats_kernel: rows, evaluations per step, case-3 histogram, failures
  rbo_gather_sums_kernel<<<1, 64, 0, h->stream>>>(h->sums, need, idx_failed, h->work_counter, sums_device);
  CK(h, cudaGetLastError());
  return RBO_SUCCESS;
}

int rbo_partial_sums_host(rbo_handle* h, double* sums, int len) {
  if (!h || !sums) return RBO_ERR_ARG;
  const int need = 1 + 3 * (1 + h->outd + h->outnth) + 2;
  if (!h->sums || len < need) return fail(h, RBO_ERR_ARG, "rbo_partial_sums_host: need %d doubles (1 + 3 (1 + d + ntheta) + 2)", need);
  CK(h, cudaSetDevice(h->device));
  double* tmp = h->sums + h->sums_len;  // the sums buffer is allocated with room for the gathered vector
  int rc = rbo_partial_sums_device(h, tmp, need);
  if (rc) return rc;
  CK(h, cudaMemcpyAsync(sums, tmp, (size_t)need * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));  // this is the call that waits for the (asynchronous) rbo_rollout_device of this handle
  return RBO_SUCCESS;
}

int rbo_get_results(rbo_handle* h, double* values, double* grad_x, double* grad_theta, int32_t* best_index, int32_t* grad_case, int32_t* status) {
  if (!h) return RBO_ERR_ARG;
  if (h->outh < 0 || h->outM <= 0) return fail(h, RBO_ERR_STATE, "rbo_get_results: no rollout has run");
  CK(h, cudaSetDevice(h->device));
  const size_t M = h->outM;
  if (values) CK(h, cudaMemcpyAsync(values, h->values, M * 8, cudaMemcpyDeviceToHost, h->stream));
  if ((grad_x || grad_theta) && h->last_mode != RBO_MODE_VALUE_GRAD) return fail(h, RBO_ERR_STATE, "rbo_get_results: the last rollout computed no gradients");
  if (grad_x) CK(h, cudaMemcpyAsync(grad_x, h->grad_x, M * h->outd * 8, cudaMemcpyDeviceToHost, h->stream));
  if (grad_theta) CK(h, cudaMemcpyAsync(grad_theta, h->grad_theta, M * h->outnth * 8, cudaMemcpyDeviceToHost, h->stream));
  if (best_index) CK(h, cudaMemcpyAsync(best_index, h->best_index, M * 4, cudaMemcpyDeviceToHost, h->stream));
  if (grad_case) CK(h, cudaMemcpyAsync(grad_case, h->grad_case, M * 4, cudaMemcpyDeviceToHost, h->stream));
  if (status) CK(h, cudaMemcpyAsync(status, h->status, M * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return RBO_SUCCESS;
}

int rbo_finalize_sums(const double* sums, int d, int ntheta, double* mean, double* std_, double* gx_mean, double* gx_std, double* gth_mean, double* gth_std) {
  if (!sums) return RBO_ERR_ARG;
  const double n = sums[0];
  const int need = 1 + 3 * (1 + d + ntheta);
  // a failed trajectory (where the reference would have thrown) or a kernel watchdog flag on ANY rank poisons the estimate
  if (sums[need] > 0.0 || sums[need + 1] > 0.0) return RBO_ERR_NUMERIC;
  if (!(n > 0)) return RBO_ERR_ARG;
  for (int row = 0; row < 1 + d + ntheta; ++row) {
    const double sm = sums[1 + 3 * row], m2 = sums[2 + 3 * row], smm = sums[3 + 3 * row];
    const double mu = sm / n;
    // sum over groups of [M2_g + n_g (mean_g - mean)^2] = sum M2_g + sum n_g mean_g^2 - n mean^2
    const double tot = std::max(m2 + smm - n * mu * mu, 0.0);
    const double sd = n > 1 ? std::sqrt(tot / (n - 1)) : NAN;
    if (row == 0) { if (mean) *mean = mu; if (std_) *std_ = sd; }
    else if (row <= d) { if (gx_mean) gx_mean[row - 1] = mu; if (gx_std) gx_std[row - 1] = sd; }
    else { if (gth_mean) gth_mean[row - 1 - d] = mu; if (gth_std) gth_std[row - 1 - d] = sd; }
  }
  return RBO_SUCCESS;
}

int rbo_get_tape(rbo_handle* h, double* xs, double* ys, double* gys, double* alphas, int32_t* n_evals, int32_t* start_status, int32_t* start_iters) {
  if (!h) return RBO_ERR_ARG;
  if (h->outh < 0) return fail(h, RBO_ERR_STATE, "rbo_get_tape: no rollout has run");
  CK(h, cudaSetDevice(h->device));
  const size_t M = h->outM, hor = h->outh, hh = std::max(h->outh, 1), d = h->outd, S = h->outS;
  if (xs) CK(h, cudaMemcpyAsync(xs, h->xs, M * (hor + 1) * d * 8, cudaMemcpyDeviceToHost, h->stream));
  if (ys) CK(h, cudaMemcpyAsync(ys, h->ys, M * (hor + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
  if (gys) CK(h, cudaMemcpyAsync(gys, h->gys, M * (hor + 1) * d * 8, cudaMemcpyDeviceToHost, h->stream));
  if (alphas) CK(h, cudaMemcpyAsync(alphas, h->alphas, M * hh * 8, cudaMemcpyDeviceToHost, h->stream));
  if (n_evals) CK(h, cudaMemcpyAsync(n_evals, h->n_evals, M * hh * 4, cudaMemcpyDeviceToHost, h->stream));
  if (start_status) CK(h, cudaMemcpyAsync(start_status, h->start_status, M * hh * S * 4, cudaMemcpyDeviceToHost, h->stream));
  if (start_iters) CK(h, cudaMemcpyAsync(start_iters, h->start_iters, M * hh * S * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return RBO_SUCCESS;
}

extern "C" int rbo_get_tape_ex(rbo_handle* h, double* mu, double* sigma, double* dmu, double* dsigma, double* Halpha) {
  if (!h) return RBO_ERR_ARG;
  if (!h->tex_valid) return fail(h, RBO_ERR_STATE, "rbo_get_tape_ex: the last rollout did not run with RBO_FLAG_TAPE_EX");
  CK(h, cudaSetDevice(h->device));
  const size_t n = (size_t)h->outM * std::max(h->outh, 0), d = h->outd;
  if (mu) CK(h, cudaMemcpyAsync(mu, h->t_mu, n * 8, cudaMemcpyDeviceToHost, h->stream));
  if (sigma) CK(h, cudaMemcpyAsync(sigma, h->t_sigma, n * 8, cudaMemcpyDeviceToHost, h->stream));
  if (dmu) CK(h, cudaMemcpyAsync(dmu, h->t_dmu, n * d * 8, cudaMemcpyDeviceToHost, h->stream));
  if (dsigma) CK(h, cudaMemcpyAsync(dsigma, h->t_dsigma, n * d * 8, cudaMemcpyDeviceToHost, h->stream));
  if (Halpha) CK(h, cudaMemcpyAsync(Halpha, h->t_Halpha, n * d * d * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return RBO_SUCCESS;
}

static int sobol_common(rbo_handle* h, int dim, int npoints, uint32_t* out_u32, double* out_f64, const double* lbs, const double* ubs) {
  if (!h) return RBO_ERR_ARG;
  if (dim < 1 || dim > RBO_SOBOL_MAXDIM || npoints < 1) return fail(h, RBO_ERR_ARG, "sobol: dim %d / npoints %d out of range", dim, npoints);
  CK(h, cudaSetDevice(h->device));
  const size_t total = (size_t)dim * npoints;
  unsigned* du = nullptr; double* df = nullptr; double* db = nullptr;
  struct Guard { unsigned*& a; double*& b; double*& c; ~Guard() { if (a) cudaFree(a); if (b) cudaFree(b); if (c) cudaFree(c); } } guard{du, df, db};  // temporaries die on every path
  if (out_u32) CK(h, cudaMalloc((void**)&du, total * 4));
  if (out_f64) CK(h, cudaMalloc((void**)&df, total * 8));
  if (lbs) {
    CK(h, cudaMalloc((void**)&db, (size_t)2 * dim * 8));
    CK(h, cudaMemcpyAsync(db, lbs, (size_t)dim * 8, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(db + dim, ubs, (size_t)dim * 8, cudaMemcpyHostToDevice, h->stream));
  }
  int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 8);
  rbo_sobol_kernel<<<blocks, 256, 0, h->stream>>>(h->sobol_dirs, du, df, dim, npoints, db, db ? db + dim : nullptr);
  CK(h, cudaGetLastError());
  if (out_u32) CK(h, cudaMemcpyAsync(out_u32, du, total * 4, cudaMemcpyDeviceToHost, h->stream));
  if (out_f64) CK(h, cudaMemcpyAsync(out_f64, df, total * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return RBO_SUCCESS;
}

int rbo_sobol_uniform(rbo_handle* h, int dim, int npoints, double* out) { return sobol_common(h, dim, npoints, nullptr, out, nullptr, nullptr); }
int rbo_sobol_uint32(rbo_handle* h, int dim, int npoints, uint32_t* out) { return sobol_common(h, dim, npoints, out, nullptr, nullptr, nullptr); }

int rbo_generate_initial_guesses(rbo_handle* h, int S, int d, const double* lbs, const double* ubs, double* out) {
  if (!h || !lbs || !ubs || !out || S < 0 || d < 1) return RBO_ERR_ARG;
  if (S > 0) {
    int rc = sobol_common(h, d, S, nullptr, out, lbs, ubs);
    if (rc) return rc;
  }
  const double eps = 1e-6;  // utils.jl:146
  for (int a = 0; a < d; ++a) { out[(size_t)S * d + a] = lbs[a] + eps; out[(size_t)(S + 1) * d + a] = ubs[a] - eps; }
  return RBO_SUCCESS;
}

int rbo_multistart_base_solve(rbo_handle* h, const double* theta, int ntheta, const double* lbs, const double* ubs, double* xfinal, double* alpha,
                              rbo_summary* summary) {
  if (!h) return RBO_ERR_ARG;
  if (!xfinal) return fail(h, RBO_ERR_ARG, "rbo_multistart_base_solve: xfinal is NULL");
  if (!h->have_sur) return fail(h, RBO_ERR_STATE, "rbo_multistart_base_solve: no surrogate");
  double x0[RBO_MAXD] = {0};
  rbo_summary local;
  int rc = launch_rollout(h, x0, theta, ntheta, lbs, ubs, 0, 0.0, RBO_MODE_VALUE, RBO_FLAG_MYOPIC_INTERNAL, nullptr, nullptr, true, summary ? summary : &local);
  if (rc) return rc;
  double a = 0;
  int st = 0;
  CK(h, cudaMemcpyAsync(xfinal, h->xs, (size_t)h->d * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaMemcpyAsync(&a, h->values, 8, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaMemcpyAsync(&st, h->status, 4, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  if (alpha) *alpha = a;
  if (st != RBO_TRAJ_OK) return fail(h, RBO_ERR_NUMERIC, "rbo_multistart_base_solve: every start produced NaN (rbf_optim.jl:129-130 would throw)");
  return RBO_SUCCESS;
}

int rbo_num_sms(const rbo_handle* h) { return h ? h->num_sms : 0; }



int rbo_fp64_peak(rbo_handle* h, double* tflops) {
  if (!h || !tflops) return RBO_ERR_ARG;
  CK(h, cudaSetDevice(h->device));
  const int blocks = h->num_sms * 4, threads = 512, iters = 4096;
  double* out = nullptr;
  struct Guard { double*& p; ~Guard() { if (p) cudaFree(p); } } guard{out};
  CK(h, cudaMalloc((void**)&out, (size_t)blocks * threads * 8));
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(h, cudaEventRecord(h->ev0, h->stream));
    rbo_fp64_peak_kernel<<<blocks, threads, 0, h->stream>>>(out, iters);
    CK(h, cudaEventRecord(h->ev1, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    float ms;
    CK(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    double fl = (double)blocks * threads * iters * 16.0 * 2.0;
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  *tflops = best;
  return RBO_SUCCESS;
}

int rbo_tr_step_batch(rbo_handle* h, int n, int B, const double* H, const double* g, const double* Delta, double* p, int* hit) {
  if (!h || !H || !g || !Delta || !p || !hit || B < 1) return RBO_ERR_ARG;
  if (n < 1 || n > RBO_MAXD) return fail(h, RBO_ERR_UNSUPPORTED, "rbo_tr_step_batch: n = %d outside [1, %d]", n, RBO_MAXD);
  CK(h, cudaSetDevice(h->device));
  double *dH = nullptr, *dg = nullptr, *dD = nullptr, *dp = nullptr; int* dh = nullptr;
  struct Guard { void** p[5]; ~Guard() { for (auto q : p) if (*q) cudaFree(*q); } } guard{{(void**)&dH, (void**)&dg, (void**)&dD, (void**)&dp, (void**)&dh}};
  CK(h, cudaMalloc((void**)&dH, (size_t)B * n * n * 8)); CK(h, cudaMalloc((void**)&dg, (size_t)B * n * 8)); CK(h, cudaMalloc((void**)&dD, (size_t)B * 8));
  CK(h, cudaMalloc((void**)&dp, (size_t)B * n * 8)); CK(h, cudaMalloc((void**)&dh, (size_t)B * 4));
  CK(h, cudaMemcpyAsync(dH, H, (size_t)B * n * n * 8, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(dg, g, (size_t)B * n * 8, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(dD, Delta, (size_t)B * 8, cudaMemcpyHostToDevice, h->stream));
  const int wpb = n > 16 ? 2 : 4;  // <= 48 KB of dynamic shared memory
  const size_t smem = (size_t)wpb * (2 * n * n + 4 * n + 32) * 8;
  rbo_tr_step_kernel<<<(B + wpb - 1) / wpb, 32 * wpb, smem, h->stream>>>(dH, dg, dD, n, B, dp, dh);
  CK(h, cudaGetLastError());
  CK(h, cudaMemcpyAsync(p, dp, (size_t)B * n * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaMemcpyAsync(hit, dh, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return RBO_SUCCESS;
}

}  // extern "C"
