// api_mimo.cu -- C ABI of the multi-input MPC layer (smpc_mimo_*, include/solvempc_b200.h): BASELINE config 3.
// Assembly on the device (mimo_assembly.cu), the solve through the shared-factor batched solver (api.cu).
#include <cfloat>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "handles.hpp"
#include "json_min.hpp"

using smpc::cuda_fail;
using smpc::DeviceBuf;
using smpc::fail;

struct smpc_mimo {
  int device = 0, B = 0;
  smpc::MimoDims dims{};
  cudaStream_t stream = nullptr;
  DeviceBuf buf;
  double *d_Ad = nullptr, *d_Bd = nullptr, *d_Q = nullptr, *d_R = nullptr, *d_umin = nullptr, *d_umax = nullptr;
  double *d_AB = nullptr, *d_AP = nullptr;
  smpc::MimoMatsDev mats{};
  double *d_x0 = nullptr, *d_xr = nullptr, *d_u0 = nullptr;
  smpc_solver *solver = nullptr;
  long long launches = 0;
};

extern "C" {

int smpc_mimo_create(smpc_mimo **out, int device, const smpc_mimo_config *cfg, int batch, const smpc_settings *settings) {
  if (!out) return fail(SMPC_ERR_ARG, "out is null");
  *out = nullptr;
  if (!cfg || !settings || !cfg->Ad || !cfg->Bd || !cfg->Q || !cfg->R || !cfg->umin || !cfg->umax) return fail(SMPC_ERR_ARG, "null config field");
  if (cfg->horizon < 1 || cfg->nx < 1 || cfg->nx > 16 || cfg->nu < 1 || batch < 1) return fail(SMPC_ERR_ARG, "need horizon >= 1, 1 <= nx <= 16, nu >= 1, batch >= 1");
  for (int c = 0; c < cfg->nu; ++c)
    if (!(cfg->umin[c] <= cfg->umax[c])) return fail(SMPC_ERR_DATA, "umin greater than umax");
  for (int c = 0; c < cfg->nx; ++c)
    if (!(cfg->Q[c] >= 0.0)) return fail(SMPC_ERR_DATA, "Q must be non-negative");
  for (int c = 0; c < cfg->nu; ++c)
    if (!(cfg->R[c] > 0.0)) return fail(SMPC_ERR_DATA, "R must be positive");
  if (int rc = smpc::select_device(device)) return rc;
  smpc_mimo *M = new smpc_mimo;
  M->device = device; M->B = batch; M->dims.N = cfg->horizon; M->dims.nx = cfg->nx; M->dims.nu = cfg->nu;
  const size_t N = cfg->horizon, nx = cfg->nx, nu = cfg->nu, n = N * nu, m = 2 * n, B = batch;
  auto body = [&]() -> int {
    size_t bytes = 0;
    for (size_t c : {nx * nx, nx * nu, nx, nu, nu, nu, N * nx * nu, N * nx * nx, n * n, m * n, m, n * nx, n * nx, N * nx * n, N * nx * nx,
                     B * nx, B * nx, B * nu}) bytes += DeviceBuf::need(c * sizeof(double));
    CK(M->buf.alloc(bytes));
    DeviceBuf &b = M->buf;
    M->d_Ad = b.take<double>(nx * nx); M->d_Bd = b.take<double>(nx * nu); M->d_Q = b.take<double>(nx); M->d_R = b.take<double>(nu);
    M->d_umin = b.take<double>(nu); M->d_umax = b.take<double>(nu); M->d_AB = b.take<double>(N * nx * nu); M->d_AP = b.take<double>(N * nx * nx);
    smpc::MimoMatsDev &t = M->mats;
    t.H = b.take<double>(n * n); t.A = b.take<double>(m * n); t.ub = b.take<double>(m); t.Fx = b.take<double>(n * nx); t.Fr = b.take<double>(n * nx);
    t.Su = b.take<double>(N * nx * n); t.Sx = b.take<double>(N * nx * nx);
    M->d_x0 = b.take<double>(B * nx); M->d_xr = b.take<double>(B * nx); M->d_u0 = b.take<double>(B * nu);
    if (!M->d_u0) return fail(SMPC_ERR_CUDA, "internal: mimo buffer carve-out overflow");
    CK(cudaMemset(b.base, 0, b.size));
    CK(cudaMemcpy(M->d_Ad, cfg->Ad, nx * nx * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(M->d_Bd, cfg->Bd, nx * nu * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(M->d_Q, cfg->Q, nx * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(M->d_R, cfg->R, nu * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(M->d_umin, cfg->umin, nu * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(M->d_umax, cfg->umax, nu * sizeof(double), cudaMemcpyHostToDevice));
    CK(smpc::launch_mimo_assemble(M->dims, M->d_Ad, M->d_Bd, M->d_Q, M->d_R, M->d_umin, M->d_umax, M->d_AB, M->d_AP, M->mats, nullptr));
    M->launches += 2;
    std::vector<double> H(n * n), A(m * n), ub(m), lb(m, -DBL_MAX);
    CK(cudaMemcpy(H.data(), t.H, H.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(A.data(), t.A, A.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ub.data(), t.ub, ub.size() * sizeof(double), cudaMemcpyDeviceToHost));
    // the solver is set up with q = 0 (x0 = xr = 0), l = -DBL_MAX, u = ub, as the reference constructor does (cpp:42-43,54-64)
    return smpc_solver_create_shared(&M->solver, device, (int)n, (int)m, batch, H.data(), A.data(), nullptr, lb.data(), ub.data(), settings);
  };
  int rc = body();
  if (rc != SMPC_OK) { std::string keep = smpc_last_error(); smpc_mimo_destroy(M); fail(rc, keep); return rc; }
  *out = M;
  return SMPC_OK;
}

int smpc_mimo_create_from_json(smpc_mimo **out, int device, const char *path, int batch, const smpc_settings *settings) {
  if (!out) return fail(SMPC_ERR_ARG, "out is null");
  *out = nullptr;
  if (!path) return fail(SMPC_ERR_ARG, "null path");
  std::ifstream f(path);
  if (!f) return fail(SMPC_ERR_IO, std::string("cannot open config '") + path + "'");
  std::stringstream ss; ss << f.rdbuf();
  std::vector<double> Ad, Bd, Q, R, umin, umax;
  smpc_mimo_config cfg{};
  try {
    const std::string text = ss.str();
    smpc::JsonParser parser(text);
    smpc::JsonValue j = parser.parse();
    int r, c;
    j.at("Ad").flatten(Ad, r, c);
    if (r != c) throw std::runtime_error("Ad must be square");
    cfg.nx = r;
    j.at("Bd").flatten(Bd, r, c);
    if (r != cfg.nx) throw std::runtime_error("Bd must have nx rows");
    cfg.nu = c;
    j.at("Q").flatten(Q, r, c); if (r * c != cfg.nx) throw std::runtime_error("Q must hold the nx diagonal weights");
    j.at("R").flatten(R, r, c); if (r * c != cfg.nu) throw std::runtime_error("R must hold the nu diagonal weights");
    j.at("umin").flatten(umin, r, c); if (r * c != cfg.nu) throw std::runtime_error("umin must have nu entries");
    j.at("umax").flatten(umax, r, c); if (r * c != cfg.nu) throw std::runtime_error("umax must have nu entries");
    if (j.at("horizon").kind != smpc::JsonValue::Number) throw std::runtime_error("horizon must be a number");
    cfg.horizon = (int)j.at("horizon").num;
  } catch (const std::exception &e) {
    return fail(SMPC_ERR_IO, std::string("config '") + path + "': " + e.what());
  }
  cfg.Ad = Ad.data(); cfg.Bd = Bd.data(); cfg.Q = Q.data(); cfg.R = R.data(); cfg.umin = umin.data(); cfg.umax = umax.data();
  return smpc_mimo_create(out, device, &cfg, batch, settings);
}

int smpc_mimo_destroy(smpc_mimo *M) {
  if (!M) return SMPC_OK;
  cudaSetDevice(M->device);
  if (M->solver) smpc_solver_destroy(M->solver);
  cudaStreamSynchronize(M->stream);
  M->buf.release();
  delete M;
  return SMPC_OK;
}

int smpc_mimo_set_stream(smpc_mimo *M, void *stream) {
  if (!M) return fail(SMPC_ERR_ARG, "null handle");
  M->stream = (cudaStream_t)stream;
  return smpc_solver_set_stream(M->solver, stream);
}

int smpc_mimo_dims(const smpc_mimo *M, int *horizon, int *nx, int *nu, int *n, int *mrows, int *batch) {
  if (!M) return fail(SMPC_ERR_ARG, "null handle");
  if (horizon) *horizon = M->dims.N;
  if (nx) *nx = M->dims.nx;
  if (nu) *nu = M->dims.nu;
  if (n) *n = M->dims.N * M->dims.nu;
  if (mrows) *mrows = 2 * M->dims.N * M->dims.nu;
  if (batch) *batch = M->B;
  return SMPC_OK;
}

smpc_solver *smpc_mimo_solver(smpc_mimo *M) { return M ? M->solver : nullptr; }

int smpc_mimo_get_matrix(smpc_mimo *M, const char *name, double *out, int capacity) {
  if (!M || !name || !out) return fail(SMPC_ERR_ARG, "null argument");
  const size_t N = M->dims.N, nx = M->dims.nx, nu = M->dims.nu, n = N * nu, m = 2 * n;
  struct { const char *nm; const double *p; size_t count; } tab[] = {
      {"H", M->mats.H, n * n}, {"A", M->mats.A, m * n}, {"ub", M->mats.ub, m}, {"Fx", M->mats.Fx, n * nx}, {"Fr", M->mats.Fr, n * nx},
      {"Su", M->mats.Su, N * nx * n}, {"Sx", M->mats.Sx, N * nx * nx}};
  for (auto &t : tab)
    if (!std::strcmp(name, t.nm)) {
      if ((size_t)capacity < t.count) return fail(SMPC_ERR_ARG, "output capacity too small");
      CK(cudaSetDevice(M->device));
      CK(cudaStreamSynchronize(M->stream));
      CK(cudaMemcpy(out, t.p, t.count * sizeof(double), cudaMemcpyDeviceToHost));
      return SMPC_OK;
    }
  return fail(SMPC_ERR_ARG, std::string("unknown matrix '") + name + "'");
}

int smpc_mimo_set_state(smpc_mimo *M, const double *x0, const double *xr, int loc) {
  if (!M) return fail(SMPC_ERR_ARG, "null handle");
  if (loc != SMPC_HOST && loc != SMPC_DEVICE) return fail(SMPC_ERR_ARG, "loc must be SMPC_HOST or SMPC_DEVICE");
  CK(cudaSetDevice(M->device));
  cudaMemcpyKind k = loc == SMPC_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  const size_t bytes = sizeof(double) * M->B * M->dims.nx;
  if (x0) CK(cudaMemcpyAsync(M->d_x0, x0, bytes, k, M->stream));
  if (xr) CK(cudaMemcpyAsync(M->d_xr, xr, bytes, k, M->stream));
  return SMPC_OK;
}

int smpc_mimo_controller_step(smpc_mimo *M) {
  if (!M) return fail(SMPC_ERR_ARG, "null handle");
  CK(cudaSetDevice(M->device));
  smpc_solver *s = M->solver;
  CK(smpc::launch_mimo_step_vectors(M->dims, M->B, M->mats.Fx, M->mats.Fr, M->d_x0, M->d_xr, s->d_q, M->stream));   // updateGradient (cpp:96)
  M->launches++;
  s->have_q = true;
  if (int rc = smpc_solver_solve(s)) return rc;                                                                     // cpp:102
  CK(smpc::launch_mimo_first_move(M->B, s->n, M->dims.nu, s->d_x, s->d_status, M->d_u0, M->stream));                 // cpp:105
  M->launches++;
  return SMPC_OK;
}

int smpc_mimo_get_control(smpc_mimo *M, double *u0, int loc) {
  if (!M || !u0) return fail(SMPC_ERR_ARG, "null argument");
  if (loc != SMPC_HOST && loc != SMPC_DEVICE) return fail(SMPC_ERR_ARG, "loc must be SMPC_HOST or SMPC_DEVICE");
  CK(cudaSetDevice(M->device));
  CK(cudaMemcpyAsync(u0, M->d_u0, sizeof(double) * M->B * M->dims.nu, loc == SMPC_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, M->stream));
  if (loc == SMPC_HOST) CK(cudaStreamSynchronize(M->stream));
  return SMPC_OK;
}

long long smpc_mimo_launch_count(const smpc_mimo *M) { return M ? M->launches + (M->solver ? M->solver->launches : 0) : 0; }

}  // extern "C"
