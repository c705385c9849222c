// OsqpEigen/OsqpEigen.h -- drop-in shim: the osqp-eigen surface that solveMPC uses, forwarded to the
// B200 C ABI (include/solvempc_b200.h).
//
// Put this directory in front of the real osqp-eigen on the include path and link libsolvempc_b200.so:
// the reference's src/ModelPredictiveControlAPI.cpp then compiles UNCHANGED (batch = 1) and its QP is
// solved on the GPU.  Members provided are exactly those the reference calls
// (src/ModelPredictiveControlAPI.cpp:51-64,96,99,102,105) plus the usual osqp-eigen setters:
//   settings()->setVerbosity / setWarmStart / setAbsoluteTolerance / setRelativeTolerance / ...
//   data()->setNumberOfVariables / setNumberOfConstraints / setHessianMatrix / setGradient /
//           setLinearConstraintsMatrix / setLowerBound / setUpperBound          (all bool)
//   initSolver(), updateGradient(), updateUpperBound(), updateLowerBound(), updateBounds(),
//   solve() (true only for status SOLVED, as osqp-eigen), getSolution(), getDualSolution().
// Header-only C++11; Eigen comes from the including project.  Errors are reported as `false`
// (osqp-eigen's convention); the text is available from smpc_last_error().
//
// The reference never sets tolerances (OSQP default 1e-3).  To run it at other tolerances WITHOUT editing
// its source, the environment variables SOLVEMPC_EPS (eps_abs = eps_rel) and SOLVEMPC_DEVICE are honoured.
#ifndef SOLVEMPC_B200_OSQPEIGEN_SHIM_H
#define SOLVEMPC_B200_OSQPEIGEN_SHIM_H

#include <Eigen/Dense>
#include <Eigen/Sparse>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <vector>

#include "../solvempc_b200.h"

namespace OsqpEigen {

class Settings {
 public:
  Settings() {
    smpc_default_settings(&m_s);
    if (const char *e = std::getenv("SOLVEMPC_EPS")) m_s.eps_abs = m_s.eps_rel = std::atof(e);
  }
  void resetDefaultSettings() { smpc_default_settings(&m_s); }
  void setVerbosity(bool v) { m_verbose = v; }
  void setWarmStart(bool w) { m_s.warm_start = w ? 1 : 0; }
  void setRho(double v) { m_s.rho = v; }
  void setSigma(double v) { m_s.sigma = v; }
  void setAlpha(double v) { m_s.alpha = v; }
  void setScaling(int v) { m_s.scaling = v; }
  void setAdaptiveRho(bool v) { m_s.adaptive_rho = v ? 1 : 0; }
  void setAdaptiveRhoInterval(int v) { m_s.adaptive_rho_interval = v; }
  void setAdaptiveRhoTolerance(double v) { m_s.adaptive_rho_tolerance = v; }
  void setMaxIteration(int v) { m_s.max_iter = v; }
  void setAbsoluteTolerance(double v) { m_s.eps_abs = v; }
  void setRelativeTolerance(double v) { m_s.eps_rel = v; }
  void setPrimalInfeasibilityTollerance(double v) { m_s.eps_prim_inf = v; }
  void setDualInfeasibilityTollerance(double v) { m_s.eps_dual_inf = v; }
  void setCheckTermination(int v) { m_s.check_termination = v; }
  void setScaledTerimination(bool v) { m_s.scaled_termination = v ? 1 : 0; }
  void setPolish(bool v) { m_polish = v; }
  void setDelta(double v) { m_delta = v; }
  void setPolishRefineIter(int v) { m_refine = v; }
  bool polish() const { return m_polish; }
  double delta() const { return m_delta; }
  int polishRefineIter() const { return m_refine; }
  const smpc_settings *getSettings() const { return &m_s; }
  bool verbose() const { return m_verbose; }

 private:
  smpc_settings m_s;
  bool m_verbose = false, m_polish = false;
  double m_delta = 1e-6;
  int m_refine = 3;
};

class Data {
 public:
  void setNumberOfVariables(int n) { m_n = n; }
  void setNumberOfConstraints(int m) { m_m = m; }
  // P: only the upper triangle is used, as osqp-eigen hands OSQP triangularView<Upper>
  template <typename Derived>
  bool setHessianMatrix(const Eigen::SparseCompressedBase<Derived> &H) {
    if (H.rows() != m_n || H.cols() != m_n) { std::cerr << "[OsqpEigen shim] Hessian must be n x n\n"; return false; }
    toCsc(H.derived(), m_Pp, m_Pi, m_Px);
    m_hasP = true;
    return true;
  }
  template <typename Derived>
  bool setLinearConstraintsMatrix(const Eigen::SparseCompressedBase<Derived> &A) {
    if (A.rows() != m_m || A.cols() != m_n) { std::cerr << "[OsqpEigen shim] constraint matrix must be m x n\n"; return false; }
    toCsc(A.derived(), m_Ap, m_Ai, m_Ax);
    m_hasA = true;
    return true;
  }
  template <typename V> bool setGradient(V &g) { return copyVec(g, m_n, m_q, m_hasQ); }
  template <typename V> bool setLowerBound(V &l) { return copyVec(l, m_m, m_l, m_hasL); }
  template <typename V> bool setUpperBound(V &u) { return copyVec(u, m_m, m_u, m_hasU); }
  bool isSet() const { return m_n > 0 && m_m >= 0 && m_hasP && m_hasQ && (m_m == 0 || (m_hasA && m_hasL && m_hasU)); }

  int m_n = 0, m_m = 0;
  std::vector<int> m_Pp, m_Pi, m_Ap, m_Ai;
  std::vector<double> m_Px, m_Ax, m_q, m_l, m_u;
  bool m_hasP = false, m_hasA = false, m_hasQ = false, m_hasL = false, m_hasU = false;

 private:
  template <typename S>
  static void toCsc(const S &M, std::vector<int> &p, std::vector<int> &i, std::vector<double> &x) {
    Eigen::SparseMatrix<double, Eigen::ColMajor, int> C = M;
    C.makeCompressed();
    p.assign(C.outerIndexPtr(), C.outerIndexPtr() + C.cols() + 1);
    i.assign(C.innerIndexPtr(), C.innerIndexPtr() + C.nonZeros());
    x.assign(C.valuePtr(), C.valuePtr() + C.nonZeros());
  }
  template <typename V>
  static bool copyVec(const V &v, int len, std::vector<double> &dst, bool &flag) {
    if (v.rows() * v.cols() != len) { std::cerr << "[OsqpEigen shim] vector has the wrong size\n"; return false; }
    dst.resize(len);
    for (int k = 0; k < len; ++k) dst[k] = v(k);
    flag = true;
    return true;
  }
};

class Solver {
 public:
  Solver() : m_settings(new Settings), m_data(new Data) {}
  ~Solver() { clearSolver(); }
  Solver(const Solver &) = delete;
  Solver &operator=(const Solver &) = delete;

  const std::unique_ptr<Settings> &settings() const { return m_settings; }
  const std::unique_ptr<Data> &data() const { return m_data; }
  bool isInitialized() const { return m_h != nullptr; }
  void clearSolver() { if (m_h) smpc_solver_destroy(m_h); m_h = nullptr; }

  bool initSolver() {
    if (m_h) { std::cerr << "[OsqpEigen shim] solver already initialised\n"; return false; }
    if (!m_data->isSet()) { std::cerr << "[OsqpEigen shim] data not completely set\n"; return false; }
    int device = 0;
    if (const char *e = std::getenv("SOLVEMPC_DEVICE")) device = std::atoi(e);
    Data &d = *m_data;
    int rc = smpc_solver_create_shared_csc(&m_h, device, d.m_n, d.m_m, 1, d.m_Pp.data(), d.m_Pi.data(), d.m_Px.data(),
                                           d.m_Ap.data(), d.m_Ai.data(), d.m_Ax.data(), d.m_q.data(), d.m_l.data(), d.m_u.data(),
                                           m_settings->getSettings());
    if (rc == SMPC_OK && m_settings->polish()) rc = smpc_solver_set_polish(m_h, 1, m_settings->delta(), m_settings->polishRefineIter());
    if (rc != SMPC_OK) { std::cerr << "[OsqpEigen shim] " << smpc_last_error() << "\n"; if (m_h) smpc_solver_destroy(m_h); m_h = nullptr; return false; }
    m_x = Eigen::VectorXd::Zero(d.m_n);
    m_y = Eigen::VectorXd::Zero(d.m_m);
    return true;
  }

  template <typename V> bool updateGradient(const Eigen::MatrixBase<V> &g) {
    if (!m_h || g.rows() * g.cols() != m_data->m_n) return false;
    Eigen::VectorXd q = g;
    return ok(smpc_solver_update_lin_cost(m_h, q.data(), SMPC_HOST)) && ok(smpc_solver_sync(m_h));
  }
  template <typename V> bool updateUpperBound(const Eigen::MatrixBase<V> &ub) {
    if (!m_h || ub.rows() * ub.cols() != m_data->m_m) return false;
    Eigen::VectorXd u = ub;
    return ok(smpc_solver_update_upper_bound(m_h, u.data(), SMPC_HOST)) && ok(smpc_solver_sync(m_h));
  }
  template <typename V> bool updateLowerBound(const Eigen::MatrixBase<V> &lb) {
    if (!m_h || lb.rows() * lb.cols() != m_data->m_m) return false;
    Eigen::VectorXd l = lb;
    return ok(smpc_solver_update_lower_bound(m_h, l.data(), SMPC_HOST)) && ok(smpc_solver_sync(m_h));
  }
  template <typename V1, typename V2> bool updateBounds(const Eigen::MatrixBase<V1> &lb, const Eigen::MatrixBase<V2> &ub) {
    return updateLowerBound(lb) && updateUpperBound(ub);
  }

  // osqp-eigen: true only when osqp_solve succeeded AND status == OSQP_SOLVED
  bool solve() {
    if (!m_h) return false;
    if (!ok(smpc_solver_solve(m_h))) return false;
    if (!ok(smpc_solver_get_solution(m_h, m_x.data(), m_y.data(), SMPC_HOST))) return false;
    if (!ok(smpc_solver_get_info(m_h, &m_status, &m_iter, nullptr, nullptr, nullptr, nullptr, nullptr, SMPC_HOST))) return false;
    return m_status == SMPC_SOLVED;
  }
  const Eigen::VectorXd &getSolution() const { return m_x; }
  const Eigen::VectorXd &getDualSolution() const { return m_y; }
  int status() const { return m_status; }
  int iterations() const { return m_iter; }
  smpc_solver *handle() { return m_h; }

 private:
  static bool ok(int rc) {
    if (rc != SMPC_OK) std::cerr << "[OsqpEigen shim] " << smpc_last_error() << "\n";
    return rc == SMPC_OK;
  }
  std::unique_ptr<Settings> m_settings;
  std::unique_ptr<Data> m_data;
  smpc_solver *m_h = nullptr;
  Eigen::VectorXd m_x, m_y;
  int m_status = SMPC_UNSOLVED, m_iter = 0;
};

}  // namespace OsqpEigen
#endif
