// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Drives the reference's own, unmodified ModelPredictiveControlAPI
// (/root/reference/src/ModelPredictiveControlAPI.cpp, compiled where it lies by
// oracle/build_ref.sh) on the CPU and prints its matrices and per-step vectors
// as JSON for tests/golden/.  The object is placement-constructed in zero-filled
// storage so that the members the reference never initialises (S rows 10-14,
// Su strict upper triangle) are zero.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <sstream>
#include <iomanip>
#include "ModelPredictiveControlAPI.h"

static std::string mat(const Eigen::MatrixXd& M) {
  std::ostringstream o; o << std::setprecision(17) << "[";
  for (int i = 0; i < M.rows(); i++) {
    o << (i ? "," : "") << "[";
    for (int j = 0; j < M.cols(); j++) o << (j ? "," : "") << M(i, j);
    o << "]";
  }
  o << "]"; return o.str();
}
static std::string vec(const Eigen::VectorXd& v) {
  std::ostringstream o; o << std::setprecision(17) << "[";
  for (int i = 0; i < v.size(); i++) o << (i ? "," : "") << v(i);
  o << "]"; return o.str();
}

int main() {
  // silence the reference's own prints
  std::streambuf* old = std::cout.rdbuf(); std::ostringstream sink; std::cout.rdbuf(sink.rdbuf());
  FILE* devnull = fopen("/dev/null", "w"); FILE* real_stdout = stdout; (void)real_stdout; (void)devnull;

  void* mem = std::calloc(1, sizeof(ModelPredictiveControlAPI));
  ModelPredictiveControlAPI* mpc = new (mem) ModelPredictiveControlAPI(false);
  std::cout.rdbuf(old);
  if (!mpc->solverFlag) { std::fprintf(stderr, "reference ctor failed\n"); return 1; }

  std::ostringstream o; o << std::setprecision(17);
  o << "{\n";
  o << "\"sizeof\": " << sizeof(ModelPredictiveControlAPI) << ",\n";
  o << "\"n\": " << mpc->n_variables << ", \"m\": " << mpc->n_constraints << ",\n";
  o << "\"H\": " << mat(Eigen::MatrixXd(mpc->H)) << ",\n";
  o << "\"H_nnz\": " << mpc->H.nonZeros() << ", \"Gbar_nnz\": " << mpc->Gbar.nonZeros() << ",\n";
  o << "\"Gbar\": " << mat(Eigen::MatrixXd(mpc->Gbar)) << ",\n";
  o << "\"Sx\": " << mat(mpc->Sx) << ",\n\"Su\": " << mat(mpc->Su) << ",\n\"CAB\": " << mat(mpc->CAB) << ",\n";
  o << "\"Fu\": " << mat(mpc->Fu) << ",\n\"Fx\": " << mat(mpc->Fx) << ",\n\"Fr\": " << mat(mpc->Fr) << ",\n";
  o << "\"S\": " << mat(mpc->S) << ",\n\"Sbar\": " << mat(mpc->Sbar) << ",\n\"Ku\": " << mat(mpc->Ku) << ",\n\"W0\": " << mat(mpc->W0) << ",\n";
  o << "\"lb\": " << mat(mpc->lb) << ",\n\"ub_init\": " << mat(mpc->ub) << ",\n\"f_init\": " << mat(mpc->f) << ",\n";
  {
    std::ostringstream c; c << "[";
    auto& calls = OsqpEigen::CallLog::get().calls;
    for (size_t i = 0; i < calls.size(); i++) c << (i ? "," : "") << "\"" << calls[i] << "\"";
    c << "]"; o << "\"ctor_calls\": " << c.str() << ",\n";
  }
  // SURVEY 8(c) cases A-D: independent cold solves on a fresh object each
  const double cases[4][6] = {{.01, 0, .02, 0, 0, 0}, {0, 0, .05, 0, 0, 0}, {.02, -.1, .03, .2, 0.5, 0.25}, {.1, .5, .08, -.2, -1, -0.3}};
  o << "\"cases\": [\n";
  for (int k = 0; k < 4; k++) {
    std::cout.rdbuf(sink.rdbuf());
    void* mem2 = std::calloc(1, sizeof(ModelPredictiveControlAPI));
    ModelPredictiveControlAPI* m2 = new (mem2) ModelPredictiveControlAPI(false);
    m2->X << cases[k][0], cases[k][1], cases[k][2], cases[k][3];
    m2->U << cases[k][4];
    m2->xref = cases[k][5];
    OsqpEigen::CallLog::get().calls.clear();
    bool ok = m2->controllerStep();
    std::cout.rdbuf(old);
    double info[8]; orc_get_info(m2->solver.workspace(), info);
    std::ostringstream c; c << "[";
    auto& calls = OsqpEigen::CallLog::get().calls;
    for (size_t i = 0; i < calls.size(); i++) c << (i ? "," : "") << "\"" << calls[i] << "\"";
    c << "]";
    Eigen::VectorXd X0(4); X0 << cases[k][0], cases[k][1], cases[k][2], cases[k][3];
    o << "{\"X\": " << vec(X0) << ", \"U\": " << cases[k][4] << ", \"xref\": " << cases[k][5]
      << ", \"ok\": " << (ok ? "true" : "false") << ", \"q\": " << vec(OsqpEigen::CallLog::get().last_q)
      << ", \"u\": " << vec(OsqpEigen::CallLog::get().last_u) << ", \"x\": " << vec(m2->solver.getSolution())
      << ", \"y\": " << vec(m2->solver.getDualSolution()) << ", \"U_after\": " << m2->U(0, 0)
      << ", \"status\": " << (int)info[0] << ", \"iter\": " << (int)info[1] << ", \"calls\": " << c.str() << "}" << (k < 3 ? ",\n" : "\n");
  }
  o << "],\n";
  // warm-started closed loop on the reference object: X+ = Ad X + Bd U (synthetic plant)
  {
    std::cout.rdbuf(sink.rdbuf());
    void* mem3 = std::calloc(1, sizeof(ModelPredictiveControlAPI));
    ModelPredictiveControlAPI* m3 = new (mem3) ModelPredictiveControlAPI(false);
    m3->X << 0.0, 0.0, 0.05, 0.0; m3->U << 0.0; m3->xref = 0.1;
    std::ostringstream us, xs, its; us << std::setprecision(17) << "["; xs << std::setprecision(17) << "["; its << "[";
    const int steps = 40;
    for (int s = 0; s < steps; s++) {
      bool ok = m3->controllerStep();
      double info[8]; orc_get_info(m3->solver.workspace(), info);
      if (!ok) { std::fprintf(stderr, "closed loop step %d failed status %d\n", s, (int)info[0]); return 2; }
      us << (s ? "," : "") << m3->U(0, 0); its << (s ? "," : "") << (int)info[1];
      Eigen::Matrix<double, 4, 1> Xn = m3->Ad * m3->X + m3->Bd * m3->U(0, 0);
      m3->X = Xn;
      xs << (s ? "," : "") << vec(Eigen::VectorXd(m3->X));
    }
    std::cout.rdbuf(old);
    us << "]"; xs << "]"; its << "]";
    o << "\"closed_loop\": {\"X0\": [0.0,0.0,0.05,0.0], \"U0\": 0.0, \"xref\": 0.1, \"steps\": " << steps
      << ", \"U\": " << us.str() << ", \"X\": " << xs.str() << ", \"iters\": " << its.str() << "}\n";
  }
  o << "}\n";
  std::fputs(o.str().c_str(), stdout);
  return 0;
}
