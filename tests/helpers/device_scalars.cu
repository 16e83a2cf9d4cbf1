// Host-side build of the PRODUCT's scalar helpers (csrc/rbo_device.cuh: kern_eval, kern_radial, rule_eval are
// __host__ __device__) so that the CPU test-suite can finite-difference the very code the CUDA kernels inline.
// Test infrastructure; built on demand by tests/test_oracle_cpu.py with nvcc (no GPU needed).
#include "../../rollout-bayesian-optimization_b200/csrc/rbo_device.cuh"
extern "C" {
void dev_rule_partials(int rule_id, double sigma_tol, double mu, double sigma, double theta1, double fstar, double* out) {
  const rbo::GPart g = rbo::rule_eval(rule_id, sigma_tol, mu, sigma, theta1, fstar);
  out[0] = g.g; out[1] = g.g_mu; out[2] = g.g_sig; out[3] = g.g_mumu; out[4] = g.g_sigsig; out[5] = g.g_muth; out[6] = g.g_sigth; out[7] = g.g_musig;
}
void dev_kernel_scalars(int kernel_id, const double* ktheta, double rho, double* out) {
  rbo::KernelSpec k;
  k.id = kernel_id;
  for (int i = 0; i < 4; ++i) k.th[i] = ktheta[i];
  rbo::kern_eval(k, rho, out[0], out[1], out[2]);
  double psi, a, b, gb;
  rbo::kern_radial(k, rho * rho, psi, a, b, gb);
  out[3] = a; out[4] = b; out[5] = gb;
}
}
