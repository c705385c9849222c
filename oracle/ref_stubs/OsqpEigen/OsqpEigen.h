// oracle/ref_stubs/OsqpEigen/OsqpEigen.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Recording stand-in for the osqp-eigen surface the reference uses
// (src/ModelPredictiveControlAPI.cpp:51-64,96,99,102,105): it logs every call
// with its payload and forwards the arithmetic to the C oracle (osqp_port.c),
// so the reference's own unmodified ModelPredictiveControlAPI.cpp can be run on
// the CPU to generate golden vectors.  The PRODUCT shim with the same surface
// (forwarding to the CUDA C-ABI) is include/OsqpEigen/OsqpEigen.h.
#ifndef ORACLE_STUB_OSQPEIGEN_H
#define ORACLE_STUB_OSQPEIGEN_H
#include <Eigen/Dense>
#include <Eigen/Sparse>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>
#include "osqp_port.h"

namespace OsqpEigen {

struct CallLog {
  std::vector<std::string> calls;
  Eigen::VectorXd last_q, last_u, last_l;
  Eigen::MatrixXd P, A;
  static CallLog& get() { static CallLog g; return g; }
};

class Settings {
 public:
  orc_settings s;
  bool verbose = false;
  Settings() {
    orc_default_settings(&s);
    // The reference never sets tolerances (OSQP default 1e-3, cpp:51-52); the north
    // star runs 1e-5 on both sides, selected here from the environment so the
    // reference source stays untouched.
    if (const char* e = std::getenv("ORC_EPS")) { s.eps_abs = s.eps_rel = std::atof(e); }
    if (const char* e = std::getenv("ORC_RHO_INTERVAL")) { s.adaptive_rho_interval = std::atoi(e); }
  }
  void setVerbosity(bool v) { verbose = v; CallLog::get().calls.push_back("setVerbosity"); }
  void setWarmStart(bool w) { s.warm_start = w ? 1 : 0; CallLog::get().calls.push_back("setWarmStart"); }
  void setAbsoluteTolerance(double e) { s.eps_abs = e; }
  void setRelativeTolerance(double e) { s.eps_rel = e; }
  void setAdaptiveRhoInterval(int k) { s.adaptive_rho_interval = k; }
};

class Data {
 public:
  int n = 0, m = 0;
  Eigen::MatrixXd P, A;
  Eigen::VectorXd q, l, u;
  void setNumberOfVariables(int n_) { n = n_; CallLog::get().calls.push_back("setNumberOfVariables"); }
  void setNumberOfConstraints(int m_) { m = m_; CallLog::get().calls.push_back("setNumberOfConstraints"); }
  template <typename S> bool setHessianMatrix(const S& H) {
    CallLog::get().calls.push_back("setHessianMatrix");
    if (H.rows() != n || H.cols() != n) return false;
    P = Eigen::MatrixXd(H); CallLog::get().P = P; return true;
  }
  template <typename V> bool setGradient(V& g) {
    CallLog::get().calls.push_back("setGradient");
    if (g.rows() != n) return false; q = g; return true;
  }
  template <typename S> bool setLinearConstraintsMatrix(const S& G) {
    CallLog::get().calls.push_back("setLinearConstraintsMatrix");
    if (G.rows() != m || G.cols() != n) return false;
    A = Eigen::MatrixXd(G); CallLog::get().A = A; return true;
  }
  template <typename V> bool setLowerBound(V& v) {
    CallLog::get().calls.push_back("setLowerBound");
    if (v.rows() != m) return false; l = v; CallLog::get().last_l = l; return true;
  }
  template <typename V> bool setUpperBound(V& v) {
    CallLog::get().calls.push_back("setUpperBound");
    if (v.rows() != m) return false; u = v; return true;
  }
};

class Solver {
  std::unique_ptr<Settings> m_settings{new Settings};
  std::unique_ptr<Data> m_data{new Data};
  orc_solver* m_w = nullptr;
  Eigen::VectorXd m_x, m_y;

 public:
  ~Solver() { if (m_w) orc_cleanup(m_w); }
  const std::unique_ptr<Settings>& settings() const { return m_settings; }
  const std::unique_ptr<Data>& data() const { return m_data; }
  orc_solver* workspace() { return m_w; }
  bool initSolver() {
    CallLog::get().calls.push_back("initSolver");
    typedef Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> RM;
    RM P = m_data->P, A = m_data->A;
    m_w = orc_setup(m_data->n, m_data->m, P.data(), m_data->q.data(), A.data(), m_data->l.data(), m_data->u.data(), &m_settings->s);
    m_x = Eigen::VectorXd::Zero(m_data->n); m_y = Eigen::VectorXd::Zero(m_data->m);
    return m_w != nullptr;
  }
  template <typename V> bool updateGradient(const V& g) {
    CallLog::get().calls.push_back("updateGradient");
    Eigen::VectorXd q = g; if (q.rows() != m_data->n) return false;
    CallLog::get().last_q = q;
    return orc_update_lin_cost(m_w, q.data()) == 0;
  }
  template <typename V> bool updateUpperBound(const V& ub) {
    CallLog::get().calls.push_back("updateUpperBound");
    Eigen::VectorXd u = ub; if (u.rows() != m_data->m) return false;
    CallLog::get().last_u = u;
    return orc_update_upper_bound(m_w, u.data()) == 0;
  }
  bool solve() {
    CallLog::get().calls.push_back("solve");
    if (orc_solve(m_w) != 0) return false;
    orc_get_solution(m_w, m_x.data(), m_y.data());
    double info[8]; orc_get_info(m_w, info);
    return static_cast<int>(info[0]) == ORC_SOLVED;
  }
  const Eigen::VectorXd& getSolution() { CallLog::get().calls.push_back("getSolution"); return m_x; }
  const Eigen::VectorXd& getDualSolution() { return m_y; }
};

}  // namespace OsqpEigen
#endif
