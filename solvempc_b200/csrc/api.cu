// api.cu -- the C ABI of include/solvempc_b200.h: handles, device memory, launches.
// No CPU fallback: every compute entry point needs a CUDA device and returns SMPC_ERR_CUDA otherwise.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "../../include/solvempc_b200.h"
#include "json_min.hpp"
#include "handles.hpp"
#include "kernels.cuh"
#include "plan.hpp"

namespace smpc {

thread_local std::string g_err;

int fail(int code, const std::string &msg) { g_err = msg; return code; }
int cuda_fail(cudaError_t e, const char *what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return SMPC_ERR_CUDA;
}

}  // namespace smpc

namespace smpc {
int select_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_err = "no CUDA device available (this library has no CPU fallback)";
    if (e != cudaSuccess) { g_err += std::string(": ") + cudaGetErrorString(e); cudaGetLastError(); }
    return SMPC_ERR_CUDA;
  }
  if (device < 0 || device >= count) return fail(SMPC_ERR_ARG, "device index out of range");
  CK(cudaSetDevice(device));
  return SMPC_OK;
}
}  // namespace smpc

using smpc::cuda_fail;
using smpc::DeviceBuf;
using smpc::fail;
using smpc::g_err;
using smpc::select_device;

struct smpc_mpc {
  int device = 0, B = 0, plants = 1;
  smpc::MpcDims dims{};
  int per_instance = 0;
  double xref = 0.0;
  cudaStream_t stream = nullptr;
  DeviceBuf buf;
  double *d_Ad = nullptr, *d_Bd = nullptr, *d_Cd = nullptr, *d_K = nullptr;
  smpc::MpcMatsDev mats{};
  double *d_X = nullptr, *d_U = nullptr, *d_ref = nullptr;
  int *d_phase = nullptr, *d_step = nullptr;          // closed-loop driver: per-instance phase, device step counter
  unsigned long long *d_stats = nullptr;              // [0] solves that did not reach SOLVED, [1] ADMM iterations
  smpc_solver *solver = nullptr;
  double *bound_U = nullptr;       // smpc_mpc_bind_results: device views of the caller's result buffers
  int *bound_status = nullptr;
  long long launches = 0;
};

namespace {

// device view of a host pointer when it is pinned (cudaHostAlloc / cudaHostRegister) and mapped; NULL for pageable memory
template <typename T>
T *pinned_device_view(T *p) {
  if (!p) return nullptr;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return (a.type == cudaMemoryTypeHost && a.devicePointer) ? static_cast<T *>(a.devicePointer) : nullptr;
}

smpc::SettingsDev to_dev(const smpc_settings &s) {
  smpc::SettingsDev d;
  d.rho0 = s.rho; d.sigma = s.sigma; d.alpha = s.alpha; d.eps_abs = s.eps_abs; d.eps_rel = s.eps_rel;
  d.eps_prim_inf = s.eps_prim_inf; d.eps_dual_inf = s.eps_dual_inf; d.rho_tol = s.adaptive_rho_tolerance;
  d.max_iter = s.max_iter; d.check_every = s.check_termination; d.adaptive_rho = s.adaptive_rho;
  d.rho_interval = s.adaptive_rho_interval; d.warm_start = s.warm_start; d.scaled_termination = s.scaled_termination;
  return d;
}

int check_settings(const smpc_settings &s) {
  if (!(s.rho > 0) || !(s.sigma > 0) || !(s.alpha > 0 && s.alpha < 2)) return fail(SMPC_ERR_ARG, "rho, sigma must be > 0 and 0 < alpha < 2");
  if (s.eps_abs < 0 || s.eps_rel < 0 || (s.eps_abs == 0 && s.eps_rel == 0)) return fail(SMPC_ERR_ARG, "eps_abs/eps_rel must be >= 0 and not both 0");
  if (s.max_iter <= 0 || s.check_termination < 0 || s.scaling < 0 || s.adaptive_rho_interval < 0) return fail(SMPC_ERR_ARG, "max_iter > 0; check_termination, scaling, adaptive_rho_interval >= 0");
  if (s.adaptive_rho && !(s.adaptive_rho_tolerance >= 1.0)) return fail(SMPC_ERR_ARG, "adaptive_rho_tolerance must be >= 1");
  return SMPC_OK;
}

int upload_plan(smpc_solver *s) {
  const smpc::SharedPlan &p = s->plan;
  const size_t n = p.n, m = p.m;
  size_t bytes = 0;
  for (size_t c : {n * n, m * n, n * m, n * n, n * n, n * n, n * n, m * n, n * n, n, n, n, m, m, m, m}) bytes += DeviceBuf::need(c * sizeof(double));
  bytes += DeviceBuf::need(m);
  CK(s->planbuf.alloc(bytes));
  auto put = [&](const std::vector<double> &v, size_t count, const double **dst) -> cudaError_t {
    double *d = s->planbuf.take<double>(count ? count : 1);
    *dst = d;
    if (!count) return cudaSuccess;
    return cudaMemcpy(d, v.data(), count * sizeof(double), cudaMemcpyHostToDevice);
  };
  smpc::SharedPlanDev &d = s->dplan;
  d.n = p.n; d.m = p.m; d.c = p.c; d.cinv = p.cinv;
  d.dx_bound = 0.0;
  for (size_t i = 0; i < n; ++i) {
    double rs = 0.0;
    for (size_t j = 0; j < n; ++j) rs += std::fabs(p.V[i * n + j]);
    d.dx_bound = std::max(d.dx_bound, (s->st.scaled_termination ? 1.0 : p.D[i]) * rs);
  }
  CK(put(p.SG, n * n, &d.SG)); CK(put(p.W, m * n, &d.W)); CK(put(p.WT, n * m, &d.WT));
  CK(put(p.V, n * n, &d.V)); CK(put(p.VT, n * n, &d.VT)); CK(put(p.PVT, n * n, &d.PVT));
  CK(put(p.VinvT, n * n, &d.VinvT)); CK(put(p.Abar, m * n, &d.Abar)); CK(put(p.Pbar, n * n, &d.Pbar));
  CK(put(p.lam, n, &d.lam)); CK(put(p.D, n, &d.D)); CK(put(p.Dinv, n, &d.Dinv));
  CK(put(p.E, m, &d.E)); CK(put(p.Einv, m, &d.Einv));
  std::vector<double> l0(m), u0(m);   // unscaled setup bounds
  for (size_t i = 0; i < m; ++i) { l0[i] = p.l0bar[i] * p.Einv[i]; u0[i] = p.u0bar[i] * p.Einv[i]; }
  CK(put(l0, m, &d.l0)); CK(put(u0, m, &d.u0));
  signed char *ct = s->planbuf.take<signed char>(m ? m : 1);
  if (m) CK(cudaMemcpy(ct, p.ctype.data(), m, cudaMemcpyHostToDevice));
  d.ctype = ct;
  return SMPC_OK;
}

// zero-padded k-major packs for admm_shared_small_kernel
int upload_small_pack(smpc_solver *s) {
  const smpc::SharedPlan &p = s->plan;
  const int n = p.n, m = p.m, NP = 16, MP = 32;
  std::vector<double> M1T((NP + MP) * NP, 0.0), WT(NP * MP, 0.0), VT(NP * NP, 0.0), PVT(NP * NP, 0.0), Ab(MP * NP, 0.0),
      V(NP * NP, 0.0), lam(NP, 0.0), D(NP, 1.0), Dinv(NP, 1.0), E(MP, 1.0), Einv(MP, 1.0);
  std::vector<int> ct(MP, 0);
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < n; ++k) {
      M1T[k * NP + i] = p.SG[(size_t)i * n + k];
      VT[k * NP + i] = p.V[(size_t)i * n + k];
      PVT[k * NP + i] = p.PVT[(size_t)k * n + i];
      V[k * NP + i] = p.V[(size_t)k * n + i];
    }
    lam[i] = p.lam[i]; D[i] = p.D[i]; Dinv[i] = p.Dinv[i];
  }
  for (int r = 0; r < m; ++r) {
    for (int k = 0; k < n; ++k) {
      M1T[(NP + r) * NP + k] = p.W[(size_t)r * n + k];   // [sigma G | W'](i=k, NP+r) = W(r,k)
      WT[k * MP + r] = p.W[(size_t)r * n + k];
      Ab[r * NP + k] = p.Abar[(size_t)r * n + k];
    }
    E[r] = p.E[r]; Einv[r] = p.Einv[r]; ct[r] = p.ctype[r];
  }
  size_t bytes = 0;
  for (size_t c : {M1T.size(), WT.size(), VT.size(), PVT.size(), Ab.size(), V.size(), lam.size(), D.size(), Dinv.size(), E.size(), Einv.size()})
    bytes += DeviceBuf::need(c * sizeof(double));
  bytes += 5 * DeviceBuf::need(1024 * sizeof(double));   // fused-kernel packs
  bytes += DeviceBuf::need(MP * sizeof(int)) + DeviceBuf::need(sizeof(int) * smpc::small_queue_ints()) +
           DeviceBuf::need(sizeof(int) * (size_t)smpc::small_sched_classes() * s->B);
  CK(s->packbuf.alloc(bytes));
  auto put = [&](const std::vector<double> &v, const double **dst) -> cudaError_t {
    double *d = s->packbuf.take<double>(v.size());
    *dst = d;
    return cudaMemcpy(d, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice);
  };
  smpc::SmallPackDev &k = s->dpack;
  CK(put(M1T, &k.M1T)); CK(put(WT, &k.WT)); CK(put(VT, &k.VT)); CK(put(PVT, &k.PVT)); CK(put(Ab, &k.Ab)); CK(put(V, &k.V));
  CK(put(lam, &k.lam)); CK(put(D, &k.D)); CK(put(Dinv, &k.Dinv)); CK(put(E, &k.E)); CK(put(Einv, &k.Einv));
  int *dct = s->packbuf.take<int>(MP);
  CK(cudaMemcpy(dct, ct.data(), MP * sizeof(int), cudaMemcpyHostToDevice));
  k.ctype = dct;
  k.mp = (p.pairs > 0 && 2 * p.pairs == p.m && !getenv("SMPC_SMALL_NO_PAIRS")) ? p.pairs : 0;
  k.M1x = k.WTt = k.C1 = k.C2 = k.cst = nullptr;
  if (k.mp > 0 && k.mp <= NP) {   // packs of the fused one-phase kernel (device_types.cuh)
    const int mp = k.mp;
    std::vector<double> M1x(NP * 32, 0.0), WTt(NP * NP, 0.0), C1(8 * 32 * 2, 0.0), C2(8 * 32 * 2, 0.0), cst(32 * 4, 0.0);
    for (int kk = 0; kk < n; ++kk) {
      for (int e = 0; e < n; ++e) M1x[kk * 32 + smpc::small_pos32(e)] = p.SG[(size_t)kk * n + e];
      for (int j = 0; j < mp; ++j) { M1x[kk * 32 + smpc::small_pos32(16 + j)] = p.W[(size_t)j * n + kk]; WTt[kk * NP + j] = p.W[(size_t)j * n + kk]; }
    }
    auto op1 = [&](int r, int c) -> double {   // [V; Wtop]
      if (c >= n) return 0.0;
      if (r < 16) return r < n ? p.V[(size_t)r * n + c] : 0.0;
      return r - 16 < mp ? p.W[(size_t)(r - 16) * n + c] : 0.0;
    };
    auto op2 = [&](int r, int c) -> double {   // [P̄V; A̅top']: rows 16 + i take the pair vector y_top - y_bot as input
      if (r < 16) return (r < n && c < n) ? p.PVT[(size_t)c * n + r] : 0.0;
      return (r - 16 < n && c < mp) ? p.Abar[(size_t)c * n + (r - 16)] : 0.0;
    };
    for (int lane = 0; lane < 32; ++lane) {
      const int a = lane >> 2, b = lane & 3;
      for (int l = 0; l < 4; ++l) {
        const int bt = b ^ smpc::small_rowmix(l), row = (bt < 2 ? 0 : 16) + 2 * a + (bt & 1);
        for (int c = 0; c < 4; ++c) {
          const size_t at = ((size_t)(l * 2 + (c >> 1)) * 32 + lane) * 2 + (c & 1);
          C1[at] = op1(row, 4 * b + c); C2[at] = op2(row, 4 * b + c);
        }
      }
      const int idx = 2 * a + (b & 1);
      if (b < 2) { cst[lane * 4] = idx < n ? p.D[idx] : 1.0; cst[lane * 4 + 1] = idx < n ? p.Dinv[idx] : 1.0; cst[lane * 4 + 2] = idx < n ? p.lam[idx] : 0.0; }
      else {
        cst[lane * 4] = idx < mp ? p.E[idx] : 1.0; cst[lane * 4 + 1] = idx < mp ? p.Einv[idx] : 1.0;
        cst[lane * 4 + 2] = idx < mp ? p.E[mp + idx] : 1.0; cst[lane * 4 + 3] = idx < mp ? p.Einv[mp + idx] : 1.0;
      }
    }
    CK(put(M1x, &k.M1x)); CK(put(WTt, &k.WTt)); CK(put(C1, &k.C1)); CK(put(C2, &k.C2)); CK(put(cst, &k.cst));
  }
  s->d_queue = s->packbuf.take<int>(smpc::small_queue_ints());
  s->d_lists = s->packbuf.take<int>((size_t)smpc::small_sched_classes() * s->B);
  if (!s->d_lists) return fail(SMPC_ERR_CUDA, "internal: pack buffer carve-out overflow");
  CK(cudaMemset(s->d_queue, 0, sizeof(int) * smpc::small_queue_ints()));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
  s->num_sms = sms > 0 ? sms : 148;
  return SMPC_OK;
}

// DMMA A-fragment packs for admm_shared_tile_kernel (layout: device_types.cuh TilePackDev)
int upload_tile_pack(smpc_solver *s) {
  const smpc::SharedPlan &p = s->plan;
  const int n = p.n, m = p.m, n8 = (n + 7) & ~7, m8 = (m + 7) & ~7;
  auto pack = [](int rows8, int K8, auto &&at) {
    std::vector<double> out((size_t)rows8 * K8);
    const int kpt = K8 / 8;
    for (int rb = 0; rb < rows8 / 8; ++rb)
      for (int kp = 0; kp < kpt; ++kp)
        for (int lane = 0; lane < 32; ++lane)
          for (int j = 0; j < 2; ++j)
            out[(((size_t)rb * kpt + kp) * 32 + lane) * 2 + j] = at(8 * rb + lane / 4, 8 * kp + 4 * j + lane % 4);
    return out;
  };
  auto M1 = pack(n8, n8 + m8, [&](int i, int k) -> double {
    if (i >= n) return 0.0;
    if (k < n8) return k < n ? p.SG[(size_t)i * n + k] : 0.0;
    const int r = k - n8;
    return r < m ? p.W[(size_t)r * n + i] : 0.0;
  });
  auto Wp = pack(m8, n8, [&](int r, int k) -> double { return (r < m && k < n) ? p.W[(size_t)r * n + k] : 0.0; });
  auto VTp = pack(n8, n8, [&](int i, int k) -> double { return (i < n && k < n) ? p.V[(size_t)k * n + i] : 0.0; });
  auto Vp = pack(n8, n8, [&](int i, int k) -> double { return (i < n && k < n) ? p.V[(size_t)i * n + k] : 0.0; });
  auto PVp = pack(n8, n8, [&](int i, int k) -> double { return (i < n && k < n) ? p.PVT[(size_t)k * n + i] : 0.0; });
  auto ATp = pack(n8, m8, [&](int i, int r) -> double { return (i < n && r < m) ? p.Abar[(size_t)r * n + i] : 0.0; });
  // paired rows [G; -G] (plan.pairs = m / 2): iteration operators on the top half only
  const int mp = p.pairs, mp8 = (mp + 7) & ~7;
  std::vector<double> M1p, Wtop, ATtop;
  if (mp > 0) {
    ATtop = pack(n8, mp8, [&](int i, int r) -> double { return (i < n && r < mp) ? p.Abar[(size_t)r * n + i] : 0.0; });
    M1p = pack(n8, n8 + mp8, [&](int i, int k) -> double {
      if (i >= n) return 0.0;
      if (k < n8) return k < n ? p.SG[(size_t)i * n + k] : 0.0;
      const int r = k - n8;
      return r < mp ? p.W[(size_t)r * n + i] : 0.0;
    });
    Wtop = pack(mp8, n8, [&](int r, int k) -> double { return (r < mp && k < n) ? p.W[(size_t)r * n + k] : 0.0; });
  }
  // x-space variant: P̄ itself (termination checks) and the diagonal of the top block
  const bool xd = !p.xdiag.empty() && !getenv("SMPC_TILE_NO_XSPACE");
  std::vector<double> Pp, adiag;
  double dmax = 1.0;
  if (xd) {
    Pp = pack(n8, n8, [&](int i, int k) -> double { return (i < n && k < n) ? p.Pbar[(size_t)i * n + k] : 0.0; });
    adiag.assign(n8, 0.0);
    for (int i = 0; i < n; ++i) adiag[i] = p.xdiag[i];
    if (!s->st.scaled_termination) { dmax = 0.0; for (int i = 0; i < n; ++i) dmax = std::max(dmax, p.D[i]); }
  }
  size_t bytes = DeviceBuf::need(sizeof(int) * 4);
  for (const std::vector<double> *v : {&M1, &Wp, &VTp, &Vp, &PVp, &ATp, &M1p, &Wtop, &ATtop, &Pp, &adiag}) bytes += DeviceBuf::need((v->size() ? v->size() : 1) * sizeof(double));
  CK(s->tilebuf.alloc(bytes));
  auto put = [&](const std::vector<double> &v, const double **dst) -> cudaError_t {
    double *d = s->tilebuf.take<double>(v.size() ? v.size() : 1);
    *dst = d;
    if (v.empty()) return cudaSuccess;
    return cudaMemcpy(d, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice);
  };
  smpc::TilePackDev &k = s->dtile;
  k.n8 = n8; k.m8 = m8;
  CK(put(M1, &k.M1)); CK(put(Wp, &k.Wp)); CK(put(VTp, &k.VTp)); CK(put(Vp, &k.Vp)); CK(put(PVp, &k.PVp)); CK(put(ATp, &k.ATp));
  k.mp = mp; k.mp8 = mp8;
  CK(put(M1p, &k.M1p)); CK(put(Wtop, &k.Wtop)); CK(put(ATtop, &k.ATtop));
  k.xd = xd ? 1 : 0; k.dmax = dmax;
  CK(put(Pp, &k.Pp)); CK(put(adiag, &k.adiag));
  s->d_queue = s->tilebuf.take<int>(4);
  if (!s->d_queue) return fail(SMPC_ERR_CUDA, "internal: tile buffer carve-out overflow");
  CK(cudaMemset(s->d_queue, 0, sizeof(int) * 4));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
  s->num_sms = sms > 0 ? sms : 148;
  s->tile_nb = smpc::tile_kernel_nb(n, m);
  if (const char *e = getenv("SMPC_TILE_NB")) { if (atoi(e) == 1 && s->tile_nb == 2) s->tile_nb = 1; }   // development knob: one 8-slot block per tile
  return SMPC_OK;
}

int alloc_batch(smpc_solver *s) {
  const size_t n = s->n, m = s->m, B = s->B;
  size_t bytes = 0;
  for (size_t c : {B * n, B * m, B * m, B * n, B * m, B * m, B, B * n, B * m, B, B, B, B * n, B * m}) bytes += DeviceBuf::need(c * sizeof(double));
  bytes += 3 * DeviceBuf::need(B * sizeof(int));
  CK(s->batchbuf.alloc(bytes));
  DeviceBuf &b = s->batchbuf;
  s->d_q = b.take<double>(B * n); s->d_l = b.take<double>(B * m); s->d_u = b.take<double>(B * m);
  s->d_xi = b.take<double>(B * n); s->d_z = b.take<double>(B * m); s->d_y = b.take<double>(B * m);
  s->d_rho = b.take<double>(B); s->d_x = b.take<double>(B * n); s->d_yout = b.take<double>(B * m);
  s->d_obj = b.take<double>(B); s->d_pri = b.take<double>(B); s->d_dua = b.take<double>(B);
  s->d_stage_x = b.take<double>(B * n); s->d_stage_y = b.take<double>(B * m);
  s->d_status = b.take<int>(B); s->d_iter = b.take<int>(B); s->d_rhoup = b.take<int>(B);
  if (!s->d_rhoup) return fail(SMPC_ERR_CUDA, "internal: batch buffer carve-out overflow");
  CK(cudaMemset(b.base, 0, b.size));
  return SMPC_OK;
}

int reset_state(smpc_solver *s, bool reset_rho) {
  const size_t n = s->n, m = s->m, B = s->B;
  CK(cudaMemsetAsync(s->d_xi, 0, B * n * sizeof(double), s->stream));
  CK(cudaMemsetAsync(s->d_z, 0, B * m * sizeof(double), s->stream));
  CK(cudaMemsetAsync(s->d_y, 0, B * m * sizeof(double), s->stream));
  if (reset_rho) {
    double rho = std::min(std::max(s->st.rho, smpc::kRhoMin), smpc::kRhoMax);
    CK(smpc::launch_fill(s->d_rho, rho, B, s->stream));
    s->launches++;
  }
  return SMPC_OK;
}

int copy_in(smpc_solver *s, double *dst, const double *src, size_t count, int loc) {
  if (!src) return fail(SMPC_ERR_ARG, "null input pointer");
  if (loc != SMPC_HOST && loc != SMPC_DEVICE) return fail(SMPC_ERR_ARG, "loc must be SMPC_HOST or SMPC_DEVICE");
  CK(cudaMemcpyAsync(dst, src, count * sizeof(double), loc == SMPC_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s->stream));
  return SMPC_OK;
}
template <typename T>
int copy_out(smpc_solver *s, T *dst, const T *src, size_t count, int loc) {
  if (!dst) return SMPC_OK;
  CK(cudaMemcpyAsync(dst, src, count * sizeof(T), loc == SMPC_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s->stream));
  return SMPC_OK;
}

int create_shared_common(smpc_solver **out, int device, int n, int m, int batch, const double *P, const double *A,
                         const double *q0, const double *l0, const double *u0, const smpc_settings *settings) {
  if (!out) return fail(SMPC_ERR_ARG, "out is null");
  *out = nullptr;
  if (!P || (!A && m > 0) || !settings) return fail(SMPC_ERR_ARG, "P, A and settings must not be null");
  if (n <= 0 || m < 0 || batch <= 0) return fail(SMPC_ERR_ARG, "need n > 0, m >= 0, batch > 0");
  if (int rc = check_settings(*settings)) return rc;
  smpc_solver *s = new smpc_solver;
  s->device = device; s->n = n; s->m = m; s->B = batch; s->st = *settings; s->regime = 0;
  std::string err;
  int rc = smpc::build_shared_plan(n, m, P, A, q0, l0, u0, s->st, s->plan, err);   // host maths, no device needed
  if (rc != SMPC_OK) { delete s; return fail(rc, err); }
  rc = select_device(device);
  if (rc == SMPC_OK) rc = upload_plan(s);
  if (rc == SMPC_OK) rc = alloc_batch(s);
  if (rc == SMPC_OK) rc = reset_state(s, true);
  if (rc == SMPC_OK && q0) {   // until updateGradient is called every instance keeps the setup gradient
    bool nz = false;
    for (int j = 0; j < n; ++j) nz = nz || q0[j] != 0.0;
    if (nz) {
      std::vector<double> rep((size_t)batch * n);
      for (int b = 0; b < batch; ++b) std::memcpy(&rep[(size_t)b * n], q0, sizeof(double) * n);
      cudaError_t e = cudaMemcpy(s->d_q, rep.data(), rep.size() * sizeof(double), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) rc = cuda_fail(e, "setup gradient upload"); else s->have_q = true;
    }
  }
  if (rc == SMPC_OK) { cudaError_t e = cudaStreamSynchronize(s->stream); if (e != cudaSuccess) rc = cuda_fail(e, "setup sync"); }
  s->kernel = 1;
  if (rc == SMPC_OK && settings->kernel != 0 && settings->kernel != 1 && settings->kernel != 2 && settings->kernel != 4 && settings->kernel != 5)
    rc = fail(SMPC_ERR_ARG, "settings.kernel must be 0 (auto), 1 (generic), 2 (small), 4 (tile) or 5 (small, DMMA)");
  if (rc == SMPC_OK && (settings->kernel == 2 || settings->kernel == 5 || (settings->kernel == 0 && smpc::small_kernel_supports(n, m)))) {
    if (!smpc::small_kernel_supports(n, m)) rc = fail(SMPC_ERR_ARG, "kernels 2 and 5 (small QPs) support n <= 16, m <= 32 only");
    else { s->kernel = settings->kernel == 5 ? 5 : 2; rc = upload_small_pack(s); }
  } else if (rc == SMPC_OK && (settings->kernel == 4 || (settings->kernel == 0 && smpc::tile_kernel_supports(n, m)))) {
    if (!smpc::tile_kernel_supports(n, m)) rc = fail(SMPC_ERR_ARG, "kernel 4 (DMMA tile): the iterates of 8 QPs do not fit shared memory");
    else { s->kernel = 4; rc = upload_tile_pack(s); }
  }
  if (rc != SMPC_OK) { s->planbuf.release(); s->batchbuf.release(); s->packbuf.release(); s->tilebuf.release(); delete s; return rc; }
  *out = s;
  return SMPC_OK;
}

}  // namespace

extern "C" {

void smpc_default_settings(smpc_settings *s) {
  if (!s) return;
  s->rho = 0.1; s->sigma = 1e-6; s->alpha = 1.6;
  s->eps_abs = 1e-3; s->eps_rel = 1e-3; s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4;
  s->adaptive_rho_tolerance = 5.0;
  s->max_iter = 4000; s->check_termination = 25; s->scaling = 10;
  s->adaptive_rho = 1; s->adaptive_rho_interval = 25;
  s->warm_start = 1; s->scaled_termination = 0; s->kernel = 0;
}

const char *smpc_last_error(void) { return g_err.c_str(); }
const char *smpc_version(void) { return "solvempc_b200 0.1 (sm_100a)"; }

int smpc_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
  return count;
}

int smpc_solver_create_shared(smpc_solver **out, int device, int n, int m, int batch, const double *P, const double *A,
                              const double *q0, const double *l0, const double *u0, const smpc_settings *settings) {
  return create_shared_common(out, device, n, m, batch, P, A, q0, l0, u0, settings);
}

int smpc_solver_create_shared_csc(smpc_solver **out, int device, int n, int m, int batch, const int *Pp, const int *Pi,
                                  const double *Px, const int *Ap, const int *Ai, const double *Ax, const double *q0,
                                  const double *l0, const double *u0, const smpc_settings *settings) {
  if (!Pp || !Pi || !Px || (m > 0 && (!Ap || !Ai || !Ax))) return fail(SMPC_ERR_ARG, "null CSC array");
  if (n <= 0 || m < 0) return fail(SMPC_ERR_ARG, "need n > 0, m >= 0");
  std::vector<double> P((size_t)n * n, 0.0), A((size_t)m * n, 0.0);
  for (int j = 0; j < n; ++j)
    for (int k = Pp[j]; k < Pp[j + 1]; ++k) {
      int i = Pi[k];
      if (i < 0 || i >= n) return fail(SMPC_ERR_ARG, "P row index out of range");
      if (i <= j) P[(size_t)i * n + j] = Px[k];   // upper triangle only, as osqp-eigen passes
    }
  for (int j = 0; j < n && m > 0; ++j)
    for (int k = Ap[j]; k < Ap[j + 1]; ++k) {
      int i = Ai[k];
      if (i < 0 || i >= m) return fail(SMPC_ERR_ARG, "A row index out of range");
      A[(size_t)i * n + j] = Ax[k];
    }
  return create_shared_common(out, device, n, m, batch, P.data(), A.data(), q0, l0, u0, settings);
}

int smpc_solver_create_batched(smpc_solver **out, int device, int n, int m, int batch, const double *P, const double *A, int loc,
                               const double *l0, const double *u0, const smpc_settings *settings) {
  if (!out) return fail(SMPC_ERR_ARG, "out is null");
  *out = nullptr;
  if (!P || (!A && m > 0) || !settings) return fail(SMPC_ERR_ARG, "P, A and settings must not be null");
  if (n <= 0 || m < 0 || batch <= 0) return fail(SMPC_ERR_ARG, "need n > 0, m >= 0, batch > 0");
  if (loc != SMPC_HOST && loc != SMPC_DEVICE) return fail(SMPC_ERR_ARG, "loc must be SMPC_HOST or SMPC_DEVICE");
  if (int rc = check_settings(*settings)) return rc;
  if (!smpc::instance_kernel_supports(n, m)) return fail(SMPC_ERR_ARG, "per-instance regime: n, m too large for one warp's shared memory");
  for (int i = 0; i < m; ++i) {
    double lo = l0 ? l0[i] : -INFINITY, hi = u0 ? u0[i] : INFINITY;
    if (lo > hi) return fail(SMPC_ERR_DATA, "lower bound greater than upper bound");
  }
  if (int rc = select_device(device)) return rc;
  smpc_solver *s = new smpc_solver;
  s->device = device; s->n = n; s->m = m; s->B = batch; s->st = *settings; s->regime = 1; s->kernel = 3;
  s->plan.n = n; s->plan.m = m;
  auto body = [&]() -> int {
    const size_t N = n, M = m, B = batch;
    size_t bytes = 0;
    for (size_t c : {B * N * N, B * M * N, B * N, B * M, B, M, M}) bytes += DeviceBuf::need(c * sizeof(double));
    CK(s->instbuf.alloc(bytes));
    smpc::InstanceDataDev &d = s->dinst;
    d.n = n; d.m = m; d.B = batch;
    d.P = s->instbuf.take<double>(B * N * N); d.A = s->instbuf.take<double>(B * M * N ? B * M * N : 1);
    d.D = s->instbuf.take<double>(B * N); d.E = s->instbuf.take<double>(B * M ? B * M : 1); d.c = s->instbuf.take<double>(B);
    double *dl0 = s->instbuf.take<double>(M ? M : 1), *du0 = s->instbuf.take<double>(M ? M : 1);
    if (!du0) return fail(SMPC_ERR_CUDA, "internal: instance buffer carve-out overflow");
    d.S0 = d.T = d.Minv0 = nullptr; d.rho_prepared = 0.0; d.paired = 0;
    d.pack = d.pack_rho = nullptr; d.queue = nullptr;
    std::vector<double> hl(M), hu(M);
    for (size_t i = 0; i < M; ++i) { hl[i] = l0 ? l0[i] : -INFINITY; hu[i] = u0 ? u0[i] : INFINITY; }
    if (M) { CK(cudaMemcpy(dl0, hl.data(), M * sizeof(double), cudaMemcpyHostToDevice)); CK(cudaMemcpy(du0, hu.data(), M * sizeof(double), cudaMemcpyHostToDevice)); }
    d.l0 = dl0; d.u0 = du0;
    cudaMemcpyKind k = loc == SMPC_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    CK(cudaMemcpy(d.P, P, B * N * N * sizeof(double), k));
    if (M) CK(cudaMemcpy(d.A, A, B * M * N * sizeof(double), k));
    // scale_data per instance (scaling == 0: zero passes, which still mirror P's upper triangle and set D = E = c = 1)
    CK(smpc::launch_ruiz_instance(d, s->st.scaling > 0 ? s->st.scaling : 0, nullptr)); s->launches++;
    if (int rc = alloc_batch(s)) return rc;
    if (int rc = reset_state(s, true)) return rc;
    const bool prep = smpc::instance_reg_supports(n, m);
    if (prep && m >= 2 && m % 2 == 0) {   // [G; -G] rows in every instance?  (checked on the scaled data; d_status is free here)
      int one = 1, flag = 0;
      CK(cudaMemcpy(s->d_status, &one, sizeof(int), cudaMemcpyHostToDevice));
      CK(smpc::launch_instance_pairs(d, s->d_status, nullptr));
      s->launches++;
      CK(cudaMemcpy(&flag, s->d_status, sizeof(int), cudaMemcpyDeviceToHost));
      CK(cudaMemset(s->d_status, 0, sizeof(int)));
      d.paired = flag;
    }
    // what osqp_setup does once per solver -- the factorisation for rho0 and the rho-independent parts of M -- is prepared
    // here, in the layout of the kernel that will run the solves
    const bool pair_kernel = d.paired && smpc::instance_pair_supports(n, m) && !getenv("SMPC_INSTANCE_NO_PAIR_KERNEL");
    const size_t tri = N * (N + 1) / 2;
    if (pair_kernel) {
      const size_t pk = smpc::instance_pair_pack_doubles(n);
      CK(s->prepbuf.alloc(DeviceBuf::need(B * pk * sizeof(double)) + DeviceBuf::need(B * sizeof(double)) + DeviceBuf::need(2 * sizeof(int))));
      d.pack = s->prepbuf.take<double>(B * pk); d.pack_rho = s->prepbuf.take<double>(B); d.queue = s->prepbuf.take<int>(2);
      if (!d.queue) return fail(SMPC_ERR_CUDA, "internal: instance pack carve-out overflow");
      CK(cudaMemset(d.queue, 0, 2 * sizeof(int)));
      int sms = 0;
      CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
      s->num_sms = sms > 0 ? sms : 148;
    } else if (prep) {
      size_t pb = 0;
      for (size_t c : {B * tri, B * tri, B * N * 32}) pb += DeviceBuf::need(c * sizeof(double));
      CK(s->prepbuf.alloc(pb));
      d.S0 = s->prepbuf.take<double>(B * tri); d.T = s->prepbuf.take<double>(B * tri); d.Minv0 = s->prepbuf.take<double>(B * N * 32);
      if (!d.Minv0) return fail(SMPC_ERR_CUDA, "internal: instance buffer carve-out overflow");
      d.rho_prepared = std::min(std::max(s->st.rho, smpc::kRhoMin), smpc::kRhoMax);
    }
    if (pair_kernel || prep) {
      smpc::BatchDev pb{};
      pb.B = batch;
      CK(pair_kernel ? smpc::launch_admm_instance_pair(d, pb, to_dev(s->st), s->num_sms, nullptr, 1)
                     : smpc::launch_admm_instance(d, pb, to_dev(s->st), nullptr, 1));
      s->launches++;
    }
    CK(cudaDeviceSynchronize());
    return SMPC_OK;
  };
  int rc = body();
  if (rc != SMPC_OK) { s->instbuf.release(); s->prepbuf.release(); s->batchbuf.release(); delete s; return rc; }
  *out = s;
  return SMPC_OK;
}

int smpc_solver_destroy(smpc_solver *s) {
  if (!s) return SMPC_OK;
  cudaSetDevice(s->device);
  cudaStreamSynchronize(s->stream);
  s->planbuf.release(); s->batchbuf.release(); s->packbuf.release(); s->instbuf.release(); s->prepbuf.release(); s->tilebuf.release();
  if (s->d_polish) cudaFree(s->d_polish);
  if (s->d_polish_scratch) cudaFree(s->d_polish_scratch);
  delete s;
  return SMPC_OK;
}

int smpc_solver_set_stream(smpc_solver *s, void *stream) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  s->stream = (cudaStream_t)stream;
  return SMPC_OK;
}

int smpc_solver_dims(const smpc_solver *s, int *n, int *m, int *batch) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  if (n) *n = s->n; if (m) *m = s->m; if (batch) *batch = s->B;
  return SMPC_OK;
}

int smpc_solver_update_lin_cost(smpc_solver *s, const double *q, int loc) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  CK(cudaSetDevice(s->device));
  if (int rc = copy_in(s, s->d_q, q, (size_t)s->B * s->n, loc)) return rc;
  s->have_q = true;
  return SMPC_OK;
}
int smpc_solver_update_upper_bound(smpc_solver *s, const double *u, int loc) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  CK(cudaSetDevice(s->device));
  if (int rc = copy_in(s, s->d_u, u, (size_t)s->B * s->m, loc)) return rc;
  s->have_u = true;
  return SMPC_OK;
}
int smpc_solver_update_lower_bound(smpc_solver *s, const double *l, int loc) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  CK(cudaSetDevice(s->device));
  if (int rc = copy_in(s, s->d_l, l, (size_t)s->B * s->m, loc)) return rc;
  s->have_l = true;
  return SMPC_OK;
}
int smpc_solver_update_bounds(smpc_solver *s, const double *l, const double *u, int loc) {
  if (int rc = smpc_solver_update_lower_bound(s, l, loc)) return rc;
  return smpc_solver_update_upper_bound(s, u, loc);
}

int smpc_solver_warm_start(smpc_solver *s, const double *x, const double *y, int loc) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  CK(cudaSetDevice(s->device));
  if (int rc = copy_in(s, s->d_stage_x, x, (size_t)s->B * s->n, loc)) return rc;
  if (int rc = copy_in(s, s->d_stage_y, y, (size_t)s->B * s->m, loc)) return rc;
  if (s->regime == 1) CK(smpc::launch_warm_start_instance(s->dinst, s->d_stage_x, s->d_stage_y, s->d_xi, s->d_z, s->d_y, s->stream));
  else CK(smpc::launch_warm_start(s->dplan, s->B, s->d_stage_x, s->d_stage_y, s->d_xi, s->d_z, s->d_y, s->stream, s->kernel == 4 && s->dtile.xd));
  s->launches++;
  return SMPC_OK;
}
int smpc_solver_cold_start(smpc_solver *s) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  CK(cudaSetDevice(s->device));
  return reset_state(s, false);
}
int smpc_solver_reset(smpc_solver *s) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  CK(cudaSetDevice(s->device));
  return reset_state(s, true);
}

int smpc_solver_solve(smpc_solver *s) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  CK(cudaSetDevice(s->device));
  smpc::BatchDev b{};
  b.B = s->B;
  b.q = s->have_q ? s->d_q : nullptr; b.l = s->have_l ? s->d_l : nullptr; b.u = s->have_u ? s->d_u : nullptr;
  b.xi = s->d_xi; b.z = s->d_z; b.y = s->d_y; b.rho = s->d_rho;
  b.x_out = s->d_x; b.y_out = s->d_yout; b.status = s->d_status; b.iter = s->d_iter; b.rho_updates = s->d_rhoup;
  b.obj = s->d_obj; b.pri_res = s->d_pri; b.dua_res = s->d_dua;
  b.fresh = s->cold_solves ? 1 : 0;
  b.u_apply = (s->regime == 0 && (s->kernel == 2 || s->kernel == 5) && !s->polish) ? s->u_apply : nullptr;   // polish may still change x
  b.u_export = (s->regime == 0 && s->kernel == 2 && b.u_apply) ? s->u_export : nullptr;
  b.status_export = (s->regime == 0 && s->kernel == 2 && !s->polish) ? s->status_export : nullptr;
  smpc::SettingsDev sd = to_dev(s->st);
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (s->timing) {
    CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
    CK(cudaEventRecord(ev0, s->stream));
  }
  cudaError_t e = s->regime == 1 ? (s->dinst.pack ? smpc::launch_admm_instance_pair(s->dinst, b, sd, s->num_sms, s->stream, 0)
                                                  : smpc::launch_admm_instance(s->dinst, b, sd, s->stream))
                  : s->kernel == 2 ? smpc::launch_admm_shared_small(s->dpack, s->dplan, b, sd, s->d_queue, s->schedule ? s->d_lists : nullptr, s->classified, s->num_sms, s->stream)
                  : s->kernel == 5 ? smpc::launch_admm_shared_small_mma(s->dpack, s->dplan, b, sd, s->d_queue, s->schedule ? s->d_lists : nullptr, s->classified, s->num_sms, s->stream)
                  : s->kernel == 4 ? smpc::launch_admm_shared_tile(s->dtile, s->dplan, b, sd, s->d_queue, s->tile_nb, s->num_sms, s->stream)
                                   : smpc::launch_admm_shared_generic(s->dplan, b, sd, s->stream);
  if (e != cudaSuccess) return cuda_fail(e, "ADMM kernel launch");
  if (s->timing) { CK(cudaEventRecord(ev1, s->stream)); s->events.emplace_back(ev0, ev1); }
  if (s->polish) {   // osqp_solve: polish(work) after the ADMM loop when settings->polish and the status is SOLVED
    smpc::PolishDataDev pd{};
    pd.n = s->n; pd.m = s->m;
    if (s->regime == 1) {
      pd.Pbar = s->dinst.P; pd.Abar = s->dinst.A; pd.D = s->dinst.D; pd.E = s->dinst.E; pd.c_inst = s->dinst.c; pd.c = 1.0;
      pd.strideP = (size_t)s->n * s->n; pd.strideA = (size_t)s->m * s->n; pd.strideD = s->n; pd.strideE = s->m;
      pd.l0 = s->dinst.l0; pd.u0 = s->dinst.u0; pd.VinvT = nullptr;
    } else {
      pd.Pbar = s->dplan.Pbar; pd.Abar = s->dplan.Abar; pd.D = s->dplan.D; pd.E = s->dplan.E; pd.c = s->dplan.c;
      pd.l0 = s->dplan.l0; pd.u0 = s->dplan.u0;
      pd.VinvT = (s->kernel == 4 && s->dtile.xd) ? nullptr : s->dplan.VinvT;
    }
    e = smpc::launch_polish(pd, b, sd, s->polish_delta, s->polish_refine, s->d_polish, s->d_polish_scratch, s->num_sms, s->stream);
    if (e != cudaSuccess) return cuda_fail(e, "polish kernel launch");
    s->launches++;
  }
  s->launches += (s->regime == 0 && (s->kernel == 2 || s->kernel == 5) && s->schedule && !s->classified) ? 2 : 1;
  s->classified = false;
  s->solved_once = true;
  return SMPC_OK;
}

int smpc_solver_set_polish(smpc_solver *s, int on, double delta, int refine_iter) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  if (on && (!(delta > 0) || refine_iter < 0)) return fail(SMPC_ERR_ARG, "polish needs delta > 0 and polish_refine_iter >= 0");
  if (on && !s->d_polish) {
    CK(cudaSetDevice(s->device));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
    s->num_sms = sms > 0 ? sms : 148;
    const size_t scratch = smpc::polish_scratch_doubles(s->n, s->m, s->num_sms);
    CK(cudaMalloc((void **)&s->d_polish, sizeof(int) * (size_t)s->B));
    CK(cudaMemset(s->d_polish, 0, sizeof(int) * (size_t)s->B));
    if (scratch) CK(cudaMalloc((void **)&s->d_polish_scratch, scratch * sizeof(double)));
  }
  s->polish = on != 0;
  if (on) { s->polish_delta = delta; s->polish_refine = refine_iter; }
  return SMPC_OK;
}
int smpc_solver_get_polish_status(smpc_solver *s, int *status_polish, int loc) {
  if (!s || !status_polish) return fail(SMPC_ERR_ARG, "null argument");
  if (!s->solved_once) return fail(SMPC_ERR_STATE, "get_polish_status before solve");
  CK(cudaSetDevice(s->device));
  if (!s->d_polish) {   // polish never enabled: 0 = not run, as OSQP reports
    if (loc == SMPC_HOST) { std::memset(status_polish, 0, sizeof(int) * (size_t)s->B); return SMPC_OK; }
    CK(cudaMemsetAsync(status_polish, 0, sizeof(int) * (size_t)s->B, s->stream));
    return SMPC_OK;
  }
  if (int rc = copy_out(s, status_polish, s->d_polish, (size_t)s->B, loc)) return rc;
  if (loc == SMPC_HOST) CK(cudaStreamSynchronize(s->stream));
  return SMPC_OK;
}

int smpc_solver_set_cold_solves(smpc_solver *s, int on) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  s->cold_solves = on != 0;
  return SMPC_OK;
}
int smpc_solver_set_scheduling(smpc_solver *s, int on) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  s->schedule = on != 0;
  return SMPC_OK;
}
int smpc_solver_enable_timing(smpc_solver *s, int on) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  s->timing = on != 0;
  return SMPC_OK;
}
int smpc_solver_kernel_ms(smpc_solver *s, double *total_ms, int *launches, int reset) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  CK(cudaSetDevice(s->device));
  for (auto &ev : s->events) {
    CK(cudaEventSynchronize(ev.second));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ev.first, ev.second));
    s->timed_ms += ms; s->timed_launches++;
    cudaEventDestroy(ev.first); cudaEventDestroy(ev.second);
  }
  s->events.clear();
  if (total_ms) *total_ms = s->timed_ms;
  if (launches) *launches = s->timed_launches;
  if (reset) { s->timed_ms = 0.0; s->timed_launches = 0; }
  return SMPC_OK;
}

int smpc_solver_get_solution(smpc_solver *s, double *x, double *y, int loc) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  if (!s->solved_once) return fail(SMPC_ERR_STATE, "get_solution before solve");
  CK(cudaSetDevice(s->device));
  if (int rc = copy_out(s, x, s->d_x, (size_t)s->B * s->n, loc)) return rc;
  if (int rc = copy_out(s, y, s->d_yout, (size_t)s->B * s->m, loc)) return rc;
  if (loc == SMPC_HOST) CK(cudaStreamSynchronize(s->stream));
  return SMPC_OK;
}

int smpc_solver_get_info(smpc_solver *s, int *status, int *iter, double *obj, double *pri_res, double *dua_res,
                         double *rho, int *rho_updates, int loc) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  if (!s->solved_once) return fail(SMPC_ERR_STATE, "get_info before solve");
  CK(cudaSetDevice(s->device));
  const size_t B = s->B;
  if (int rc = copy_out(s, status, s->d_status, B, loc)) return rc;
  if (int rc = copy_out(s, iter, s->d_iter, B, loc)) return rc;
  if (int rc = copy_out(s, obj, s->d_obj, B, loc)) return rc;
  if (int rc = copy_out(s, pri_res, s->d_pri, B, loc)) return rc;
  if (int rc = copy_out(s, dua_res, s->d_dua, B, loc)) return rc;
  if (int rc = copy_out(s, rho, s->d_rho, B, loc)) return rc;
  if (int rc = copy_out(s, rho_updates, s->d_rhoup, B, loc)) return rc;
  if (loc == SMPC_HOST) CK(cudaStreamSynchronize(s->stream));
  return SMPC_OK;
}

int smpc_solver_count_solved(smpc_solver *s, int *count) {
  if (!s || !count) return fail(SMPC_ERR_ARG, "null argument");
  if (!s->solved_once) return fail(SMPC_ERR_STATE, "count_solved before solve");
  std::vector<int> st(s->B);
  CK(cudaSetDevice(s->device));
  CK(cudaMemcpyAsync(st.data(), s->d_status, sizeof(int) * s->B, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  int c = 0;
  for (int v : st) c += (v == SMPC_SOLVED);
  *count = c;
  return SMPC_OK;
}

int smpc_solver_sync(smpc_solver *s) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  CK(cudaSetDevice(s->device));
  CK(cudaStreamSynchronize(s->stream));
  return SMPC_OK;
}

int smpc_solver_get_scaling(smpc_solver *s, double *D, double *E, double *c) {
  if (!s) return fail(SMPC_ERR_ARG, "null handle");
  if (s->regime == 1) {   // per-instance regime: the scaling of instance 0
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    if (D) CK(cudaMemcpy(D, s->dinst.D, sizeof(double) * s->n, cudaMemcpyDeviceToHost));
    if (E && s->m) CK(cudaMemcpy(E, s->dinst.E, sizeof(double) * s->m, cudaMemcpyDeviceToHost));
    if (c) CK(cudaMemcpy(c, s->dinst.c, sizeof(double), cudaMemcpyDeviceToHost));
    return SMPC_OK;
  }
  if (D) std::memcpy(D, s->plan.D.data(), sizeof(double) * s->n);
  if (E) std::memcpy(E, s->plan.E.data(), sizeof(double) * s->m);
  if (c) *c = s->plan.c;
  return SMPC_OK;
}

long long smpc_solver_launch_count(const smpc_solver *s) { return s ? s->launches : 0; }
int smpc_solver_row_pairs(const smpc_solver *s) {
  if (!s) return 0;
  if (s->regime == 1) return s->dinst.paired ? s->m / 2 : 0;
  return s->kernel == 4 ? s->dtile.mp : s->kernel == 2 ? s->dpack.mp : 0;
}
const char *smpc_solver_kernel_name(const smpc_solver *s) {
  if (!s) return "";
  return s->regime == 1 ? (s->dinst.pack ? "admm_instance_pair_kernel" : "admm_instance_kernel") : s->kernel == 2 ? (smpc::small_fused_supports(s->dpack, s->dplan) ? "admm_shared_small_fused_kernel" : "admm_shared_small_kernel")
         : s->kernel == 4 ? (s->dtile.xd ? "admm_shared_tile_kernel<x-space>" : "admm_shared_tile_kernel") : s->kernel == 5 ? "admm_shared_small_mma_kernel" : "admm_shared_generic_kernel";
}

/* host-only inspection of the shared plan (no device needed): used by the CPU tests of the host logic */
int smpc_shared_plan_inspect(int n, int m, const double *P, const double *A, const double *q0, const double *l0, const double *u0,
                             const smpc_settings *settings, double *D, double *E, double *c, double *lam, double *V,
                             double *SG, double *W, double *PVT, double *VinvT, signed char *ctype) {
  if (!P || !settings || (!A && m > 0)) return fail(SMPC_ERR_ARG, "null argument");
  if (int rc = check_settings(*settings)) return rc;
  smpc::SharedPlan pl; std::string err;
  int rc = smpc::build_shared_plan(n, m, P, A, q0, l0, u0, *settings, pl, err);
  if (rc != SMPC_OK) return fail(rc, err);
  auto cp = [](double *dst, const std::vector<double> &v) { if (dst) std::memcpy(dst, v.data(), v.size() * sizeof(double)); };
  cp(D, pl.D); cp(E, pl.E); if (c) *c = pl.c; cp(lam, pl.lam); cp(V, pl.V); cp(SG, pl.SG); cp(W, pl.W); cp(PVT, pl.PVT); cp(VinvT, pl.VinvT);
  if (ctype) std::memcpy(ctype, pl.ctype.data(), pl.ctype.size());
  return SMPC_OK;
}

/* ------------------------------------------------------------------------------------ MPC layer */
static int mpc_alloc(smpc_mpc *M, const smpc_mpc_config *cfg) {
  const size_t N = M->dims.N, nx = M->dims.nx, P = M->plants, B = M->B;
  size_t bytes = 0;
  for (size_t c : {P * nx * nx, P * nx, nx, nx, P * N * N, P * 2 * N * N, P * N * nx, P * N, P * N * N, P * 2 * N * nx, P * 2 * N,
                   P * 2 * N, P * N * nx, P * N * N, P * N, P * N * N, B * nx, B, B}) bytes += DeviceBuf::need(c * sizeof(double));
  bytes += DeviceBuf::need(B * sizeof(int)) + DeviceBuf::need(sizeof(int)) + DeviceBuf::need(2 * sizeof(unsigned long long));
  CK(M->buf.alloc(bytes));
  DeviceBuf &b = M->buf;
  M->d_Ad = b.take<double>(P * nx * nx); M->d_Bd = b.take<double>(P * nx); M->d_Cd = b.take<double>(nx); M->d_K = b.take<double>(nx);
  smpc::MpcMatsDev &t = M->mats;
  t.H = b.take<double>(P * N * N); t.Gbar = b.take<double>(P * 2 * N * N); t.Fx = b.take<double>(P * N * nx); t.Fu = b.take<double>(P * N);
  t.Fr = b.take<double>(P * N * N); t.Sbar = b.take<double>(P * 2 * N * nx); t.Ku = b.take<double>(P * 2 * N); t.W0 = b.take<double>(P * 2 * N);
  t.Sx = b.take<double>(P * N * nx); t.Su = b.take<double>(P * N * N); t.CAB = b.take<double>(P * N); t.FrT = b.take<double>(P * N * N);
  M->d_X = b.take<double>(B * nx); M->d_U = b.take<double>(B); M->d_ref = b.take<double>(B);
  M->d_phase = b.take<int>(B); M->d_step = b.take<int>(1); M->d_stats = b.take<unsigned long long>(2);
  if (!M->d_stats) return fail(SMPC_ERR_CUDA, "internal: mpc buffer carve-out overflow");
  CK(cudaMemset(b.base, 0, b.size));
  CK(cudaMemcpy(M->d_Ad, cfg->Ad, P * nx * nx * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(M->d_Bd, cfg->Bd, P * nx * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(M->d_Cd, cfg->Cd, nx * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(M->d_K, cfg->K, nx * sizeof(double), cudaMemcpyHostToDevice));
  return SMPC_OK;
}

int smpc_mpc_create(smpc_mpc **out, int device, const smpc_mpc_config *cfg, int batch, const smpc_settings *settings) {
  if (!out) return fail(SMPC_ERR_ARG, "out is null");
  *out = nullptr;
  if (!cfg || !settings || !cfg->Ad || !cfg->Bd || !cfg->Cd || !cfg->K) return fail(SMPC_ERR_ARG, "null config field");
  if (cfg->horizon < 1 || cfg->nx < 1 || cfg->nx > 16 || batch < 1) return fail(SMPC_ERR_ARG, "need horizon >= 1, 1 <= nx <= 16, batch >= 1");
  if (int rc = check_settings(*settings)) return rc;
  if (int rc = select_device(device)) return rc;
  smpc_mpc *M = new smpc_mpc;
  M->device = device; M->B = batch; M->per_instance = cfg->per_instance; M->plants = cfg->per_instance ? batch : 1;
  M->dims.N = cfg->horizon; M->dims.nx = cfg->nx; M->dims.n_state_rows = cfg->n_state_rows; M->dims.Q = cfg->Q; M->dims.R = cfg->R;
  M->dims.RD = cfg->RD; M->dims.u_limit = cfg->u_limit; M->xref = cfg->xref;
  int rc = mpc_alloc(M, cfg);
  if (rc == SMPC_OK) {
    cudaError_t e = smpc::launch_mpc_assemble(M->dims, M->plants, M->d_Ad, M->d_Bd, M->d_Cd, M->d_K, M->mats, nullptr);
    if (e != cudaSuccess) rc = cuda_fail(e, "mpc_assemble launch"); else M->launches++;
  }
  const int N = cfg->horizon;
  std::vector<double> H((size_t)N * N), G((size_t)2 * N * N), W0(2 * N), lb(2 * N, -DBL_MAX);
  if (rc == SMPC_OK) {
    cudaError_t e = cudaMemcpy(H.data(), M->mats.H, H.size() * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(G.data(), M->mats.Gbar, G.size() * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(W0.data(), M->mats.W0, W0.size() * sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "assembly readback");
  }
  // cpp:42-43,54-64: n = N, m = 2N, q = f(X=U=ref=0) = 0, l = -DBL_MAX, u = W0 (X = U = 0)
  if (rc == SMPC_OK) {
    if (M->per_instance) rc = smpc_solver_create_batched(&M->solver, device, N, 2 * N, batch, M->mats.H, M->mats.Gbar, SMPC_DEVICE, lb.data(), W0.data(), settings);
    else rc = smpc_solver_create_shared(&M->solver, device, N, 2 * N, batch, H.data(), G.data(), nullptr, lb.data(), W0.data(), settings);
  }
  if (rc == SMPC_OK) {
    std::vector<double> ref(batch, cfg->xref);
    cudaError_t e = cudaMemcpy(M->d_ref, ref.data(), sizeof(double) * batch, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = cuda_fail(e, "ref upload");
  }
  if (rc != SMPC_OK) { std::string keep = g_err; smpc_mpc_destroy(M); g_err = keep; return rc; }
  *out = M;
  return SMPC_OK;
}

int smpc_mpc_create_from_json(smpc_mpc **out, int device, const char *path, int batch, const smpc_settings *settings) {
  if (!out) return fail(SMPC_ERR_ARG, "out is null");
  *out = nullptr;
  if (!path) return fail(SMPC_ERR_ARG, "null path");
  std::ifstream f(path);
  if (!f) return fail(SMPC_ERR_IO, std::string("cannot open config '") + path + "'");
  std::stringstream ss; ss << f.rdbuf();
  std::vector<double> Ad, Bd, Cd, K, tmp;
  smpc_mpc_config cfg{};
  try {
    const std::string text = ss.str();
    smpc::JsonParser parser(text);
    smpc::JsonValue j = parser.parse();
    int r, c;
    j.at("Ad").flatten(Ad, r, c);
    if (r != c) throw std::runtime_error("Ad must be square");
    cfg.nx = r;
    // shape checks as the reference's from_json does (cpp:465-472)
    j.at("Bd").flatten(Bd, r, c); if (r * c != cfg.nx) throw std::runtime_error("Bd must have nx entries");
    j.at("Cd").flatten(Cd, r, c); if (r * c != cfg.nx) throw std::runtime_error("Cd must be 1 x nx");
    j.at("K").flatten(K, r, c); if (r * c != cfg.nx) throw std::runtime_error("K must be 1 x nx");
    j.at("Dd").flatten(tmp, r, c); if (r * c != 1) throw std::runtime_error("Dd must be 1 x 1");   // loaded, never used (cpp:116)
    j.at("Q").flatten(tmp, r, c); if (r * c != 1) throw std::runtime_error("Q must be 1 x 1"); cfg.Q = tmp[0];
    j.at("R").flatten(tmp, r, c); if (r * c != 1) throw std::runtime_error("R must be 1 x 1"); cfg.R = tmp[0];
    j.at("RD").flatten(tmp, r, c); if (r * c != 1) throw std::runtime_error("RD must be 1 x 1"); cfg.RD = tmp[0];
    if (j.at("xref").kind != smpc::JsonValue::Number) throw std::runtime_error("xref must be a number");
    cfg.xref = j.at("xref").num;
    cfg.horizon = j.has("horizon") ? (int)j.at("horizon").num : 15;            // mpcWindow, h:26
    cfg.n_state_rows = j.has("n_state_rows") ? (int)j.at("n_state_rows").num : 10;  // literal at cpp:185
    cfg.u_limit = j.has("u_limit") ? j.at("u_limit").num : 255.0;              // literal at cpp:368
  } catch (const std::exception &e) {
    return fail(SMPC_ERR_IO, std::string("config '") + path + "': " + e.what());
  }
  cfg.Ad = Ad.data(); cfg.Bd = Bd.data(); cfg.Cd = Cd.data(); cfg.K = K.data(); cfg.per_instance = 0;
  return smpc_mpc_create(out, device, &cfg, batch, settings);
}

int smpc_mpc_destroy(smpc_mpc *M) {
  if (!M) return SMPC_OK;
  cudaSetDevice(M->device);
  if (M->solver) smpc_solver_destroy(M->solver);
  cudaStreamSynchronize(M->stream);
  M->buf.release();
  delete M;
  return SMPC_OK;
}

int smpc_mpc_set_stream(smpc_mpc *M, void *stream) {
  if (!M) return fail(SMPC_ERR_ARG, "null handle");
  M->stream = (cudaStream_t)stream;
  return smpc_solver_set_stream(M->solver, stream);
}

int smpc_mpc_dims(const smpc_mpc *M, int *horizon, int *nx, int *n, int *mrows, int *batch) {
  if (!M) return fail(SMPC_ERR_ARG, "null handle");
  if (horizon) *horizon = M->dims.N; if (nx) *nx = M->dims.nx; if (n) *n = M->dims.N; if (mrows) *mrows = 2 * M->dims.N; if (batch) *batch = M->B;
  return SMPC_OK;
}

smpc_solver *smpc_mpc_solver(smpc_mpc *M) { return M ? M->solver : nullptr; }

int smpc_mpc_get_matrix(smpc_mpc *M, const char *name, int index, double *out, int capacity) {
  if (!M || !name || !out) return fail(SMPC_ERR_ARG, "null argument");
  if (index < 0 || index >= M->plants) return fail(SMPC_ERR_ARG, "plant index out of range");
  const size_t N = M->dims.N, nx = M->dims.nx;
  struct { const char *nm; const double *p; size_t count; } tab[] = {
      {"H", M->mats.H, N * N}, {"Gbar", M->mats.Gbar, 2 * N * N}, {"Fx", M->mats.Fx, N * nx}, {"Fu", M->mats.Fu, N},
      {"Fr", M->mats.Fr, N * N}, {"Sbar", M->mats.Sbar, 2 * N * nx}, {"Ku", M->mats.Ku, 2 * N}, {"W0", M->mats.W0, 2 * N},
      {"Sx", M->mats.Sx, N * nx}, {"Su", M->mats.Su, N * N}, {"CAB", M->mats.CAB, N}};
  for (auto &t : tab)
    if (!std::strcmp(name, t.nm)) {
      if ((size_t)capacity < t.count) return fail(SMPC_ERR_ARG, "output capacity too small");
      CK(cudaSetDevice(M->device));
      CK(cudaStreamSynchronize(M->stream));
      CK(cudaMemcpy(out, t.p + (size_t)index * t.count, t.count * sizeof(double), cudaMemcpyDeviceToHost));
      return SMPC_OK;
    }
  return fail(SMPC_ERR_ARG, std::string("unknown matrix '") + name + "'");
}

int smpc_mpc_set_state(smpc_mpc *M, const double *X, const double *U, const double *ref, int loc) {
  if (!M) return fail(SMPC_ERR_ARG, "null handle");
  if (loc != SMPC_HOST && loc != SMPC_DEVICE) return fail(SMPC_ERR_ARG, "loc must be SMPC_HOST or SMPC_DEVICE");
  CK(cudaSetDevice(M->device));
  if (loc == SMPC_DEVICE) {   // one gather kernel instead of up to three copy nodes
    CK(smpc::launch_mpc_copy_state(M->B, M->dims.nx, X, U, ref, M->d_X, M->d_U, M->d_ref, M->stream));
    M->launches++;
    return SMPC_OK;
  }
  // pinned host buffers: the gather kernel reads them over PCIe directly (one launch instead of up to three DMA copies);
  // like cudaMemcpyAsync it reads the host memory asynchronously, on the stream
  const double *vX = pinned_device_view(X), *vU = pinned_device_view(U), *vr = pinned_device_view(ref);
  if ((!X || vX) && (!U || vU) && (!ref || vr) && (X || U || ref)) {
    CK(smpc::launch_mpc_copy_state(M->B, M->dims.nx, vX, vU, vr, M->d_X, M->d_U, M->d_ref, M->stream));
    M->launches++;
    return SMPC_OK;
  }
  if (X) CK(cudaMemcpyAsync(M->d_X, X, sizeof(double) * M->B * M->dims.nx, cudaMemcpyHostToDevice, M->stream));
  if (U) CK(cudaMemcpyAsync(M->d_U, U, sizeof(double) * M->B, cudaMemcpyHostToDevice, M->stream));
  if (ref) CK(cudaMemcpyAsync(M->d_ref, ref, sizeof(double) * M->B, cudaMemcpyHostToDevice, M->stream));
  return SMPC_OK;
}

namespace {
// the small-QP kernels' step: f, ub and the scheduling lists in one launch
bool fused_step_available(const smpc_mpc *M) {
  const smpc_solver *s = M->solver;
  const bool small = s->regime == 0 && (s->kernel == 2 || s->kernel == 5);
  return small && s->schedule && !M->per_instance && !s->have_l && M->dims.N <= 16;
}
// controllerStep reading the state from (X, U, ref); `keep`: they are the caller's buffers, copy them into the controller's own
int mpc_step_impl(smpc_mpc *M, const double *X, const double *U, const double *ref, bool keep);
}  // namespace

int smpc_mpc_controller_step(smpc_mpc *M) {
  if (!M) return fail(SMPC_ERR_ARG, "null handle");
  CK(cudaSetDevice(M->device));
  return mpc_step_impl(M, M->d_X, M->d_U, M->d_ref, false);
}

int smpc_mpc_controller_step_from(smpc_mpc *M, const double *X, const double *U, const double *ref, int loc) {
  if (!M) return fail(SMPC_ERR_ARG, "null handle");
  if (loc != SMPC_HOST && loc != SMPC_DEVICE) return fail(SMPC_ERR_ARG, "loc must be SMPC_HOST or SMPC_DEVICE");
  if (!X || !U || !ref) return fail(SMPC_ERR_ARG, "controller_step_from: X, U and ref are all required");
  CK(cudaSetDevice(M->device));
  if (fused_step_available(M)) {
    // device buffers, or pinned host buffers through their device view: the step's first kernel reads them where they lie
    const double *vX = loc == SMPC_DEVICE ? X : pinned_device_view(X), *vU = loc == SMPC_DEVICE ? U : pinned_device_view(U),
                 *vr = loc == SMPC_DEVICE ? ref : pinned_device_view(ref);
    if (vX && vU && vr) return mpc_step_impl(M, vX, vU, vr, true);
  }
  const int rc = smpc_mpc_set_state(M, X, U, ref, loc);
  return rc ? rc : mpc_step_impl(M, M->d_X, M->d_U, M->d_ref, false);
}

namespace {
int mpc_step_impl(smpc_mpc *M, const double *X, const double *U, const double *ref, bool keep) {
  smpc_solver *s = M->solver;
  // setF (cpp:90), W0 + Sbar X + Ku U (cpp:99) written straight into the solver's q / u (updateGradient, updateUpperBound);
  // for the small-QP kernels the same launch also fills their longest-expected-first scheduling lists
  if (fused_step_available(M)) {
    CK(smpc::launch_mpc_step_classify(M->dims, M->B, M->mats, X, U, ref, s->d_q, s->d_u, s->dpack, s->dplan,
                                      s->d_queue + 1, s->d_lists, keep ? M->d_X : nullptr, keep ? M->d_U : nullptr,
                                      keep ? M->d_ref : nullptr, M->stream));
    s->classified = true;
  } else {
    CK(smpc::launch_mpc_step_vectors(M->dims, M->B, M->per_instance, M->mats, M->d_X, M->d_U, M->d_ref, s->d_q, s->d_u, M->stream));
  }
  M->launches++;
  s->have_q = true; s->have_u = true;
  // cpp:102 solve, cpp:105 U += dU*[0]: inside the small-QP kernels' store_solution, a separate kernel otherwise
  const bool fused = s->regime == 0 && (s->kernel == 2 || s->kernel == 5) && !s->polish;
  // bound result buffers (smpc_mpc_bind_results): written by the one-warp kernel as each instance ends, by an export
  // kernel behind the solve otherwise
  const bool bound = M->bound_U || M->bound_status, export_fused = bound && fused && s->kernel == 2;
  s->u_apply = fused ? M->d_U : nullptr;
  s->u_export = export_fused ? M->bound_U : nullptr;
  s->status_export = export_fused ? M->bound_status : nullptr;
  const int rc = smpc_solver_solve(s);
  s->u_apply = nullptr; s->u_export = nullptr; s->status_export = nullptr;
  if (rc) {
    // the ADMM kernel's last warp zeroes the class counters; if it never ran they must not survive into the next step
    if (s->classified) { cudaMemsetAsync(s->d_queue, 0, sizeof(int) * smpc::small_queue_ints(), M->stream); s->classified = false; }
    return rc;
  }
  if (!fused) {
    CK(smpc::launch_mpc_apply_control(M->B, s->n, s->d_x, s->d_status, M->d_U, M->stream));
    M->launches++;
  }
  if (bound && !export_fused) {
    CK(smpc::launch_mpc_export(M->B, M->d_U, s->d_status, M->bound_U, M->bound_status, M->stream));
    M->launches++;
  }
  return SMPC_OK;
}
}  // namespace

int smpc_mpc_plant_step(smpc_mpc *M) {
  if (!M) return fail(SMPC_ERR_ARG, "null handle");
  CK(cudaSetDevice(M->device));
  CK(smpc::launch_mpc_plant_step(M->B, M->dims.nx, M->per_instance, M->d_Ad, M->d_Bd, M->d_X, M->d_U, M->stream));
  M->launches++;
  return SMPC_OK;
}

int smpc_mpc_closed_loop(smpc_mpc *M, int steps, double ref_amplitude, int ref_period, const int *phase, int use_graph,
                         long long *not_solved, long long *iterations) {
  if (!M) return fail(SMPC_ERR_ARG, "null handle");
  if (steps < 0 || (ref_period != 0 && ref_period < 2)) return fail(SMPC_ERR_ARG, "need steps >= 0 and ref_period 0 (constant reference) or >= 2");
  CK(cudaSetDevice(M->device));
  smpc_solver *s = M->solver;
  if (s->timing && use_graph) return fail(SMPC_ERR_STATE, "kernel timing events cannot be captured: disable timing or run without the graph");
  cudaStream_t st = M->stream;
  cudaStream_t own = nullptr;
  if (use_graph && st == nullptr) {   // the legacy default stream cannot be captured
    CK(cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking));
    CK(cudaDeviceSynchronize());
    st = own;
    smpc_mpc_set_stream(M, st);
  }
  auto restore = [&]() { if (own) { cudaStreamSynchronize(own); smpc_mpc_set_stream(M, nullptr); cudaStreamDestroy(own); } };
  auto body = [&]() -> int {
    CK(cudaMemsetAsync(M->d_stats, 0, 2 * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(M->d_step, 0, sizeof(int), st));
    if (phase) CK(cudaMemcpyAsync(M->d_phase, phase, sizeof(int) * M->B, cudaMemcpyHostToDevice, st));
    else CK(cudaMemsetAsync(M->d_phase, 0, sizeof(int) * M->B, st));
    auto one_step = [&]() -> int {
      if (ref_period) { CK(smpc::launch_mpc_square_ref(M->B, ref_amplitude, ref_period, M->d_phase, M->d_step, M->d_ref, st)); M->launches++; }
      CK(smpc::launch_mpc_step_vectors(M->dims, M->B, M->per_instance, M->mats, M->d_X, M->d_U, M->d_ref, s->d_q, s->d_u, st));
      M->launches++;
      s->have_q = true; s->have_u = true;
      if (int rc = smpc_solver_solve(s)) return rc;
      CK(smpc::launch_mpc_advance(M->B, s->n, M->dims.nx, M->per_instance, M->d_Ad, M->d_Bd, s->d_x, s->d_status, s->d_iter, M->d_X,
                                  M->d_U, M->d_stats, M->d_step, st));
      M->launches++;
      return SMPC_OK;
    };
    if (use_graph && steps > 1) {
      // one controller step + plant step captured once, replayed `steps` times (launch-bound for small QPs)
      cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
      CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      const long long l0 = M->launches, sl0 = s->launches;
      int rc = one_step();
      cudaError_t e = cudaStreamEndCapture(st, &graph);
      if (rc != SMPC_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture");
      const long long per_step = M->launches - l0, s_per_step = s->launches - sl0;
      e = cudaGraphInstantiate(&exec, graph, 0);
      if (e != cudaSuccess) { cudaGraphDestroy(graph); return cuda_fail(e, "cudaGraphInstantiate"); }
      for (int k = 0; k < steps && e == cudaSuccess; ++k) e = cudaGraphLaunch(exec, st);
      M->launches += per_step * (steps - 1); s->launches += s_per_step * (steps - 1);
      cudaError_t e2 = cudaStreamSynchronize(st);
      cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
      if (e != cudaSuccess) return cuda_fail(e, "cudaGraphLaunch");
      if (e2 != cudaSuccess) return cuda_fail(e2, "closed-loop sync");
    } else {
      for (int k = 0; k < steps; ++k) if (int rc = one_step()) return rc;
    }
    unsigned long long h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, M->d_stats, sizeof(h), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (not_solved) *not_solved = (long long)h[0];
    if (iterations) *iterations = (long long)h[1];
    return SMPC_OK;
  };
  int rc = body();
  restore();
  return rc;
}

int smpc_mpc_get_state(smpc_mpc *M, double *X, double *U, int loc) {
  if (!M) return fail(SMPC_ERR_ARG, "null handle");
  CK(cudaSetDevice(M->device));
  cudaMemcpyKind k = loc == SMPC_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  if (X) CK(cudaMemcpyAsync(X, M->d_X, sizeof(double) * M->B * M->dims.nx, k, M->stream));
  if (U) CK(cudaMemcpyAsync(U, M->d_U, sizeof(double) * M->B, k, M->stream));
  if (loc == SMPC_HOST) CK(cudaStreamSynchronize(M->stream));
  return SMPC_OK;
}

int smpc_mpc_get_control_status(smpc_mpc *M, double *U, int *status, int loc) {
  if (!M) return fail(SMPC_ERR_ARG, "null handle");
  if (loc != SMPC_HOST && loc != SMPC_DEVICE) return fail(SMPC_ERR_ARG, "loc must be SMPC_HOST or SMPC_DEVICE");
  if (status && !M->solver->solved_once) return fail(SMPC_ERR_STATE, "status requested before the first controllerStep");
  CK(cudaSetDevice(M->device));
  if (loc == SMPC_HOST) {
    double *vU = pinned_device_view(U);
    int *vs = pinned_device_view(status);
    if ((!U || vU) && (!status || vs) && (U || status)) {   // pinned destinations: one export kernel writes them over PCIe
      CK(smpc::launch_mpc_export(M->B, M->d_U, M->solver->d_status, vU, vs, M->stream));
      M->launches++;
      CK(cudaStreamSynchronize(M->stream));
      return SMPC_OK;
    }
  }
  cudaMemcpyKind k = loc == SMPC_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  if (U) CK(cudaMemcpyAsync(U, M->d_U, sizeof(double) * M->B, k, M->stream));
  if (status) CK(cudaMemcpyAsync(status, M->solver->d_status, sizeof(int) * M->B, k, M->stream));
  if (loc == SMPC_HOST) CK(cudaStreamSynchronize(M->stream));
  return SMPC_OK;
}

int smpc_mpc_bind_results(smpc_mpc *M, double *U, int *status, int loc) {
  if (!M) return fail(SMPC_ERR_ARG, "null handle");
  if (loc != SMPC_HOST && loc != SMPC_DEVICE) return fail(SMPC_ERR_ARG, "loc must be SMPC_HOST or SMPC_DEVICE");
  CK(cudaSetDevice(M->device));
  double *vU = U;
  int *vs = status;
  if (loc == SMPC_HOST) {
    vU = pinned_device_view(U); vs = pinned_device_view(status);
    if ((U && !vU) || (status && !vs))
      return fail(SMPC_ERR_ARG, "bind_results: host buffers must be pinned (cudaHostAlloc / cudaHostRegister); use smpc_mpc_get_control_status for pageable memory");
  }
  M->bound_U = vU; M->bound_status = vs;
  return SMPC_OK;
}

int smpc_mpc_sync(smpc_mpc *M) {
  if (!M) return fail(SMPC_ERR_ARG, "null handle");
  CK(cudaSetDevice(M->device));
  CK(cudaStreamSynchronize(M->stream));
  return SMPC_OK;
}

int smpc_mpc_get_step_vectors(smpc_mpc *M, double *f, double *ub, int loc) {
  if (!M) return fail(SMPC_ERR_ARG, "null handle");
  CK(cudaSetDevice(M->device));
  cudaMemcpyKind k = loc == SMPC_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  smpc_solver *s = M->solver;
  if (f) CK(cudaMemcpyAsync(f, s->d_q, sizeof(double) * M->B * s->n, k, M->stream));
  if (ub) CK(cudaMemcpyAsync(ub, s->d_u, sizeof(double) * M->B * s->m, k, M->stream));
  if (loc == SMPC_HOST) CK(cudaStreamSynchronize(M->stream));
  return SMPC_OK;
}

long long smpc_mpc_launch_count(const smpc_mpc *M) { return M ? M->launches + (M->solver ? M->solver->launches : 0) : 0; }

}  // extern "C"
