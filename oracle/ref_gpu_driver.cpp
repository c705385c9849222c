// oracle/ref_gpu_driver.cpp -- TEST INFRASTRUCTURE.
// The reference's own, UNMODIFIED ModelPredictiveControlAPI (compiled from /root/reference where it lies)
// on top of the PRODUCT shim include/OsqpEigen/OsqpEigen.h and libsolvempc_b200.so: the drop-in of
// SURVEY.md 8(b) at batch = 1 (BASELINE config 1).  Prints the known-answer cases and a warm-started closed
// loop as JSON; tests/test_gpu_dropin.py compares it with tests/golden/assembly_ref.json.
// Run from the repo root (the reference opens ./config/MPC_API.json, cpp:12) with SOLVEMPC_EPS=1e-5.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cstdlib>
#include <iomanip>
#include <new>
#include <sstream>
#include "ModelPredictiveControlAPI.h"

static std::string vec(const Eigen::VectorXd &v) {
  std::ostringstream o; o << std::setprecision(17) << "[";
  for (int i = 0; i < v.size(); i++) o << (i ? "," : "") << v(i);
  o << "]"; return o.str();
}

static ModelPredictiveControlAPI *fresh() {
  void *mem = std::calloc(1, sizeof(ModelPredictiveControlAPI));   // zero storage: S rows 10-14, Su upper
  return new (mem) ModelPredictiveControlAPI(false);
}

// `ref_mpc_gpu latency K`: BASELINE config 1 latency -- K warm-started closed-loop controllerSteps of the reference's own
// class (square-wave reference +-0.1, period 200 steps; synthetic plant X <- Ad X + Bd U), each timed on the host clock
// around controllerStep() alone (cpp:81-108: f, ub, updateGradient, updateUpperBound, solve, getSolution).
static int latency(int K) {
  std::streambuf *old = std::cout.rdbuf(); std::ostringstream sink; std::cout.rdbuf(sink.rdbuf());
  ModelPredictiveControlAPI *m = fresh();
  if (!m->solverFlag) { std::cout.rdbuf(old); std::fprintf(stderr, "ctor failed: %s\n", smpc_last_error()); return 1; }
  m->X << 0.0, 0.0, 0.05, 0.0; m->U << 0.0;
  const int warm = 20;
  std::vector<double> us; us.reserve(K);
  long iters = 0; int bad = 0;
  for (int s = 0; s < warm + K; s++) {
    m->xref = (2 * (s % 200) < 200) ? 0.1 : -0.1;
    auto t0 = std::chrono::steady_clock::now();
    bool ok = m->controllerStep();
    auto t1 = std::chrono::steady_clock::now();
    if (s >= warm) { us.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count()); iters += m->solver.iterations(); bad += !ok; }
    Eigen::Matrix<double, 4, 1> Xn = m->Ad * m->X + m->Bd * m->U(0, 0);
    m->X = Xn;
  }
  std::cout.rdbuf(old);
  double sum = 0; for (double v : us) sum += v;
  std::sort(us.begin(), us.end());
  std::printf("{\"steps\": %d, \"latency_us_mean\": %.3f, \"latency_us_median\": %.3f, \"iters_mean\": %.2f, \"not_solved\": %d}\n",
              K, sum / K, us[K / 2], (double)iters / K, bad);
  return 0;
}

int main(int argc, char **argv) {
  if (argc >= 3 && !std::strcmp(argv[1], "latency")) return latency(std::max(1, std::atoi(argv[2])));
  std::streambuf *old = std::cout.rdbuf(); std::ostringstream sink; std::cout.rdbuf(sink.rdbuf());
  const double cases[4][6] = {{.01, 0, .02, 0, 0, 0}, {0, 0, .05, 0, 0, 0}, {.02, -.1, .03, .2, 0.5, 0.25}, {.1, .5, .08, -.2, -1, -0.3}};
  std::ostringstream o; o << std::setprecision(17) << "{\"cases\": [";
  for (int k = 0; k < 4; k++) {
    ModelPredictiveControlAPI *m = fresh();
    if (!m->solverFlag) { std::cout.rdbuf(old); std::fprintf(stderr, "ctor failed: %s\n", smpc_last_error()); return 1; }
    m->X << cases[k][0], cases[k][1], cases[k][2], cases[k][3];
    m->U << cases[k][4];
    m->xref = cases[k][5];
    bool ok = m->controllerStep();
    o << (k ? "," : "") << "{\"ok\": " << (ok ? "true" : "false") << ", \"status\": " << m->solver.status() << ", \"iter\": " << m->solver.iterations()
      << ", \"x\": " << vec(m->solver.getSolution()) << ", \"U_after\": " << m->U(0, 0) << "}";
  }
  o << "], ";
  ModelPredictiveControlAPI *m = fresh();
  m->X << 0.0, 0.0, 0.05, 0.0; m->U << 0.0; m->xref = 0.1;
  std::ostringstream us, its; us << std::setprecision(17) << "["; its << "[";
  for (int s = 0; s < 40; s++) {
    if (!m->controllerStep()) { std::cout.rdbuf(old); std::fprintf(stderr, "closed loop step %d failed (status %d)\n", s, m->solver.status()); return 2; }
    us << (s ? "," : "") << m->U(0, 0); its << (s ? "," : "") << m->solver.iterations();
    Eigen::Matrix<double, 4, 1> Xn = m->Ad * m->X + m->Bd * m->U(0, 0);
    m->X = Xn;
  }
  us << "]"; its << "]";
  o << "\"closed_loop\": {\"U\": " << us.str() << ", \"iters\": " << its.str() << "}}\n";
  std::cout.rdbuf(old);
  std::fputs(o.str().c_str(), stdout);
  return 0;
}
