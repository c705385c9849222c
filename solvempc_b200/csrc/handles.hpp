// handles.hpp -- the opaque handles of the C ABI and the helpers the api_*.cu translation units share.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/solvempc_b200.h"
#include "kernels.cuh"
#include "plan.hpp"

namespace smpc {

// thread-local error message behind smpc_last_error(); both return `code`
int fail(int code, const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);
// cudaSetDevice with the "no CUDA device: no CPU fallback" error
int select_device(int device);

#define CK(call)                                               \
  do {                                                         \
    cudaError_t e__ = (call);                                  \
    if (e__ != cudaSuccess) return smpc::cuda_fail(e__, #call); \
  } while (0)

struct DeviceBuf {   // one cudaMalloc carved into aligned pieces
  char *base = nullptr;
  size_t size = 0, used = 0;
  cudaError_t alloc(size_t bytes) { size = bytes; used = 0; return cudaMalloc((void **)&base, bytes ? bytes : 256); }
  template <typename T> T *take(size_t count) {
    size_t off = (used + 255) & ~size_t(255);
    used = off + count * sizeof(T);
    return used <= size ? reinterpret_cast<T *>(base + off) : nullptr;
  }
  static size_t need(size_t bytes) { return ((bytes + 255) & ~size_t(255)) + 256; }
  void release() { if (base) cudaFree(base); base = nullptr; }
};

}  // namespace smpc

struct smpc_solver {
  int device = 0, n = 0, m = 0, B = 0;
  int regime = 0;  // 0 shared-factor, 1 per-instance
  smpc_settings st{};
  cudaStream_t stream = nullptr;
  smpc::SharedPlan plan;
  smpc::DeviceBuf planbuf, batchbuf;
  smpc::SharedPlanDev dplan{};
  double *d_q = nullptr, *d_l = nullptr, *d_u = nullptr;
  bool have_q = false, have_l = false, have_u = false;
  double *d_xi = nullptr, *d_z = nullptr, *d_y = nullptr, *d_rho = nullptr;
  double *d_x = nullptr, *d_yout = nullptr, *d_obj = nullptr, *d_pri = nullptr, *d_dua = nullptr;
  int *d_status = nullptr, *d_iter = nullptr, *d_rhoup = nullptr;
  double *d_stage_x = nullptr, *d_stage_y = nullptr;  // warm-start staging
  smpc::DeviceBuf instbuf;                                  // per-instance regime: P̄, A̅, D, E, c
  smpc::DeviceBuf prepbuf;                                  // per-instance regime: what osqp_setup prepares (factor for rho0, S0 / T split)
  smpc::InstanceDataDev dinst{};
  smpc::DeviceBuf packbuf;                                  // small-kernel operator pack + work queue
  smpc::SmallPackDev dpack{};
  int *d_queue = nullptr, *d_lists = nullptr;
  smpc::DeviceBuf tilebuf;                                  // tile-kernel operator packs (DMMA A fragments) + work queue
  smpc::TilePackDev dtile{};
  int tile_nb = 0;
  int num_sms = 148;
  bool schedule = true;   // longest-expected-first pre-pass of the small kernel
  long long launches = 0;
  int kernel = 1;
  bool classified = false;     // the scheduling lists of the next solve were filled by the MPC layer's fused step-vector kernel
  double *u_apply = nullptr;   // MPC layer: U to increment inside the small-QP kernels' store (cpp:105), else NULL
  double *u_export = nullptr;  // MPC layer: bound result buffers (smpc_mpc_bind_results) for this solve, kernel 2 only
  int *status_export = nullptr;
  bool polish = false;         // smpc_solver_set_polish
  double polish_delta = 1e-6;
  int polish_refine = 3;
  int *d_polish = nullptr;     // [B] status_polish of the last solve
  double *d_polish_scratch = nullptr;
  bool solved_once = false;
  bool cold_solves = false, timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;   // pending kernel timings
  double timed_ms = 0.0;
  int timed_launches = 0;
};

