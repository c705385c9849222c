// plan.cpp -- see plan.hpp.  Host-side, executed once per shared plant (the analogue of
// osqp_setup's scaling + factorisation, reference call site src/ModelPredictiveControlAPI.cpp:64).
#include "plan.hpp"

#include <algorithm>
#include <cmath>

namespace smpc {

namespace {
constexpr double kRhoMin = 1e-6, kRhoEqOverIneq = 1e3, kRhoTol = 1e-4;
constexpr double kInfty = 1e30, kMinScaling = 1e-4, kMaxScaling = 1e4;

inline double limit_scaling(double v) {
  v = v < kMinScaling ? 1.0 : v;
  return v > kMaxScaling ? kMaxScaling : v;
}
}  // namespace

void ruiz_scale(int n, int m, int iters, std::vector<double> &P, std::vector<double> &A, std::vector<double> &q,
                std::vector<double> &D, std::vector<double> &E, double &c) {
  D.assign(n, 1.0);
  E.assign(m, 1.0);
  c = 1.0;
  std::vector<double> dt(n), et(m);
  for (int it = 0; it < iters; ++it) {
    // column infinity norms of the KKT matrix [P A'; A 0]
    for (int j = 0; j < n; ++j) {
      double r = 0.0;
      for (int i = 0; i < n; ++i) r = std::max(r, std::fabs(P[(size_t)i * n + j]));
      for (int i = 0; i < m; ++i) r = std::max(r, std::fabs(A[(size_t)i * n + j]));
      dt[j] = 1.0 / std::sqrt(limit_scaling(r));
    }
    for (int i = 0; i < m; ++i) {
      double r = 0.0;
      for (int j = 0; j < n; ++j) r = std::max(r, std::fabs(A[(size_t)i * n + j]));
      et[i] = 1.0 / std::sqrt(limit_scaling(r));
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) P[(size_t)i * n + j] = (dt[i] * P[(size_t)i * n + j]) * dt[j];
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = (et[i] * A[(size_t)i * n + j]) * dt[j];
    for (int j = 0; j < n; ++j) q[j] = dt[j] * q[j];
    for (int j = 0; j < n; ++j) D[j] *= dt[j];
    for (int i = 0; i < m; ++i) E[i] *= et[i];
    // cost normalisation; in the reference the setup gradient is 0 (X = U = ref = 0, cpp:22-23,38-39),
    // whose norm OSQP's limit_scaling replaces by 1
    double mean = 0.0;
    for (int j = 0; j < n; ++j) {
      double r = 0.0;
      for (int i = 0; i < n; ++i) r = std::max(r, std::fabs(P[(size_t)i * n + j]));
      mean += r;
    }
    mean /= n;
    double nq = 0.0;
    for (int j = 0; j < n; ++j) nq = std::max(nq, std::fabs(q[j]));
    double ct = limit_scaling(std::max(mean, limit_scaling(nq)));
    ct = 1.0 / ct;
    for (size_t k = 0; k < (size_t)n * n; ++k) P[k] *= ct;
    for (int j = 0; j < n; ++j) q[j] *= ct;
    c *= ct;
  }
}

void jacobi_eigh(int n, std::vector<double> &C, std::vector<double> &w, std::vector<double> &Q) {
  Q.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) Q[(size_t)i * n + i] = 1.0;
  auto at = [&](int i, int j) -> double & { return C[(size_t)i * n + j]; };
  double scale = 0.0;
  for (size_t k = 0; k < (size_t)n * n; ++k) scale = std::max(scale, std::fabs(C[k]));
  if (scale == 0.0) scale = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) off = std::max(off, std::fabs(at(i, j)));
    if (off <= 1e-15 * scale) break;   // reachable in double: the rotations leave off-diagonals at round-off level
    for (int p = 0; p < n - 1; ++p) {
      for (int q = p + 1; q < n; ++q) {
        double apq = at(p, q);
        if (std::fabs(apq) <= 1e-17 * scale) { at(p, q) = at(q, p) = 0.0; continue; }
        double app = at(p, p), aqq = at(q, q);
        double theta = (aqq - app) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < n; ++k) {  // columns p, q
          double akp = at(k, p), akq = at(k, q);
          at(k, p) = cs * akp - sn * akq;
          at(k, q) = sn * akp + cs * akq;
        }
        for (int k = 0; k < n; ++k) {  // rows p, q
          double apk = at(p, k), aqk = at(q, k);
          at(p, k) = cs * apk - sn * aqk;
          at(q, k) = sn * apk + cs * aqk;
        }
        at(p, q) = at(q, p) = 0.0;   // annihilated exactly by this rotation
        for (int k = 0; k < n; ++k) {
          double qkp = Q[(size_t)k * n + p], qkq = Q[(size_t)k * n + q];
          Q[(size_t)k * n + p] = cs * qkp - sn * qkq;
          Q[(size_t)k * n + q] = sn * qkp + cs * qkq;
        }
      }
    }
  }
  w.resize(n);
  for (int i = 0; i < n; ++i) w[i] = at(i, i);
}

int build_shared_plan(int n, int m, const double *P_in, const double *A_in, const double *q0, const double *l0,
                      const double *u0, const smpc_settings &st, SharedPlan &pl, std::string &err) {
  if (n <= 0 || m < 0) { err = "n must be > 0 and m >= 0"; return SMPC_ERR_ARG; }
  pl.n = n; pl.m = m;
  // P: upper triangle mirrored (what osqp-eigen hands to OSQP)
  pl.Pbar.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {
      double v = P_in[(size_t)i * n + j];
      if (!std::isfinite(v)) { err = "P has a non-finite entry"; return SMPC_ERR_DATA; }
      pl.Pbar[(size_t)i * n + j] = v; pl.Pbar[(size_t)j * n + i] = v;
    }
  pl.Abar.assign(A_in, A_in + (size_t)m * n);
  for (double v : pl.Abar) if (!std::isfinite(v)) { err = "A has a non-finite entry"; return SMPC_ERR_DATA; }
  for (int i = 0; i < m; ++i) {
    double lo = l0 ? l0[i] : -INFINITY, hi = u0 ? u0[i] : INFINITY;
    if (lo > hi) { err = "lower bound greater than upper bound"; return SMPC_ERR_DATA; }
  }
  std::vector<double> qs(n, 0.0);
  if (q0) for (int j = 0; j < n; ++j) { qs[j] = q0[j]; if (!std::isfinite(q0[j])) { err = "q has a non-finite entry"; return SMPC_ERR_DATA; } }
  if (st.scaling > 0) ruiz_scale(n, m, st.scaling, pl.Pbar, pl.Abar, qs, pl.D, pl.E, pl.c);
  else { pl.D.assign(n, 1.0); pl.E.assign(m, 1.0); pl.c = 1.0; }
  pl.cinv = 1.0 / pl.c;
  pl.Dinv.resize(n); pl.Einv.resize(m);
  for (int j = 0; j < n; ++j) pl.Dinv[j] = 1.0 / pl.D[j];
  for (int i = 0; i < m; ++i) pl.Einv[i] = 1.0 / pl.E[i];
  pl.l0bar.resize(m); pl.u0bar.resize(m); pl.ctype.resize(m);
  for (int i = 0; i < m; ++i) {
    pl.l0bar[i] = pl.E[i] * (l0 ? l0[i] : -INFINITY);
    pl.u0bar[i] = pl.E[i] * (u0 ? u0[i] : INFINITY);
    // OSQP set_rho_vec classification on the scaled bounds
    if (pl.l0bar[i] < -kInfty * kMinScaling && pl.u0bar[i] > kInfty * kMinScaling) pl.ctype[i] = -1;
    else if (pl.u0bar[i] - pl.l0bar[i] < kRhoTol) pl.ctype[i] = 1;
    else pl.ctype[i] = 0;
  }
  const double *Pb = pl.Pbar.data(), *Ab = pl.Abar.data();
  // S = P̄ + sigma I + rho_min A_f'A_f ;  T = A̅' diag(kappa) A̅
  std::vector<double> S((size_t)n * n), T((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) S[(size_t)i * n + j] = Pb[(size_t)i * n + j] + (i == j ? st.sigma : 0.0);
  for (int r = 0; r < m; ++r) {
    double kap = pl.ctype[r] == 0 ? 1.0 : (pl.ctype[r] == 1 ? kRhoEqOverIneq : 0.0);
    const double *a = Ab + (size_t)r * n;
    for (int i = 0; i < n; ++i) {
      if (a[i] == 0.0) continue;
      for (int j = 0; j < n; ++j) {
        if (pl.ctype[r] == -1) S[(size_t)i * n + j] += kRhoMin * a[i] * a[j];
        else T[(size_t)i * n + j] += kap * a[i] * a[j];
      }
    }
  }
  // Cholesky S = L L'
  std::vector<double> L((size_t)n * n, 0.0);
  for (int j = 0; j < n; ++j) {
    double d = S[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
    if (!(d > 0.0)) { err = "P + sigma*I is not positive definite"; return SMPC_ERR_DATA; }
    d = std::sqrt(d);
    L[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = S[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
      L[(size_t)i * n + j] = s / d;
    }
  }
  // Li = L^-1 (lower)
  std::vector<double> Li((size_t)n * n, 0.0);
  for (int j = 0; j < n; ++j) {
    Li[(size_t)j * n + j] = 1.0 / L[(size_t)j * n + j];
    for (int i = j + 1; i < n; ++i) {
      double s = 0.0;
      for (int k = j; k < i; ++k) s -= L[(size_t)i * n + k] * Li[(size_t)k * n + j];
      Li[(size_t)i * n + j] = s / L[(size_t)i * n + i];
    }
  }
  // C = Li T Li'  (symmetrised), eigen-decomposition C = Q diag(lam) Q'
  std::vector<double> tmp((size_t)n * n, 0.0), Cm((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k <= i; ++k) {
      double lik = Li[(size_t)i * n + k];
      if (lik == 0.0) continue;
      for (int j = 0; j < n; ++j) tmp[(size_t)i * n + j] += lik * T[(size_t)k * n + j];
    }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int k = 0; k <= j; ++k) s += tmp[(size_t)i * n + k] * Li[(size_t)j * n + k];
      Cm[(size_t)i * n + j] = s;
    }
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      double v = 0.5 * (Cm[(size_t)i * n + j] + Cm[(size_t)j * n + i]);
      Cm[(size_t)i * n + j] = v; Cm[(size_t)j * n + i] = v;
    }
  std::vector<double> Q;
  jacobi_eigh(n, Cm, pl.lam, Q);
  for (double &v : pl.lam) if (v < 0.0) v = 0.0;  // T is PSD; clip round-off
  // V = Li' Q
  pl.V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int k = i; k < n; ++k) {
      double lki = Li[(size_t)k * n + i];
      if (lki == 0.0) continue;
      for (int j = 0; j < n; ++j) pl.V[(size_t)i * n + j] += lki * Q[(size_t)k * n + j];
    }
  const double *V = pl.V.data();
  pl.VT.resize((size_t)n * n); pl.SG.assign((size_t)n * n, 0.0); pl.PVT.assign((size_t)n * n, 0.0);
  pl.VinvT.assign((size_t)n * n, 0.0);
  pl.W.assign((size_t)m * n, 0.0); pl.WT.resize((size_t)n * m);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) pl.VT[(size_t)j * n + i] = V[(size_t)i * n + j];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double g = 0.0, pv = 0.0, sv = 0.0;
      for (int k = 0; k < n; ++k) {
        g += V[(size_t)k * n + i] * V[(size_t)k * n + j];
        pv += Pb[(size_t)i * n + k] * V[(size_t)k * n + j];
        sv += S[(size_t)i * n + k] * V[(size_t)k * n + j];
      }
      pl.SG[(size_t)i * n + j] = st.sigma * g;
      pl.PVT[(size_t)j * n + i] = pv;     // PVT[k][i] = (P̄V)[i][k]
      pl.VinvT[(size_t)i * n + j] = sv;   // Vinv = V'S  =>  Vinv' = S V ; VinvT[k][i] = Vinv[i][k] = (SV)[k][i]
    }
  for (int r = 0; r < m; ++r)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += Ab[(size_t)r * n + k] * V[(size_t)k * n + j];
      pl.W[(size_t)r * n + j] = s; pl.WT[(size_t)j * m + r] = s;
    }
  pl.pairs = 0;
  if (m % 2 == 0) {
    int h = m / 2, ok = 0;
    for (int r = 0; r < h; ++r) {
      bool neg = true;
      for (int j = 0; j < n && neg; ++j) neg = (Ab[(size_t)(r + h) * n + j] == -Ab[(size_t)r * n + j]);
      ok += neg;
    }
    if (ok == h) pl.pairs = h;
  }
  // paired rows whose top block is diagonal (a two-sided box on the variables): the tile kernel can iterate in x-space
  pl.xdiag.clear();
  if (pl.pairs == n && n > 0) {
    bool diag = true;
    for (int r = 0; r < n && diag; ++r)
      for (int j = 0; j < n && diag; ++j) diag = (j == r) || Ab[(size_t)r * n + j] == 0.0;
    if (diag) { pl.xdiag.resize(n); for (int r = 0; r < n; ++r) pl.xdiag[r] = Ab[(size_t)r * n + r]; }
  }
  return SMPC_OK;
}

}  // namespace smpc
