// examples/multi_gpu_batch.cpp -- ONE global batch of controllers over several GPUs through the C ABI.
//
// The QPs are independent, so the batch index is split contiguously over the devices (the same partition as
// solvempc_b200/sharding.py: sizes differ by at most one); each shard gets its own smpc_mpc handle on its device and its own
// host thread (a handle is used from one thread at a time, include/solvempc_b200.h); there is no data-path collective, the
// "gather" is every shard writing its rows of the caller's result arrays.  The result is compared bitwise with the whole
// batch stepped on device 0 alone.
//
//   g++ -O2 -std=c++17 -Iinclude examples/multi_gpu_batch.cpp -Lsolvempc_b200 -lsolvempc_b200 -lpthread \
//       -Wl,-rpath,'$ORIGIN/../../solvempc_b200' -o examples/_build/multi_gpu_batch
//   examples/_build/multi_gpu_batch config/MPC_API.json 4099 [shards]      (shards defaults to the device count; with one
//                                                                          device several shards share it)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "solvempc_b200.h"

static void shard_bounds(int batch, int world, int rank, int *lo, int *hi) {
  const int base = batch / world, extra = batch % world;
  *lo = rank * base + (rank < extra ? rank : extra);
  *hi = *lo + base + (rank < extra ? 1 : 0);
}

// controllerStep of rows [lo, hi) of the global arrays on `device`; returns 0 when ok
static int step_shard(const char *json, int device, int lo, int hi, const double *X, const double *U, const double *ref,
                      double *U_out, int *status_out, const smpc_settings *st) {
  smpc_mpc *m = nullptr;
  if (smpc_mpc_create_from_json(&m, device, json, hi - lo, st)) { std::fprintf(stderr, "shard [%d,%d): %s\n", lo, hi, smpc_last_error()); return 1; }
  int rc = smpc_mpc_set_state(m, X + 4 * (size_t)lo, U + lo, ref + lo, SMPC_HOST);
  if (!rc) rc = smpc_mpc_controller_step(m);
  if (!rc) rc = smpc_mpc_get_control_status(m, U_out + lo, status_out + lo, SMPC_HOST);   // synchronises
  if (rc) std::fprintf(stderr, "shard [%d,%d): %s\n", lo, hi, smpc_last_error());
  smpc_mpc_destroy(m);
  return rc;
}

int main(int argc, char **argv) {
  const char *json = argc > 1 ? argv[1] : "config/MPC_API.json";
  const int batch = argc > 2 ? std::atoi(argv[2]) : 4099;
  const int devices = smpc_device_count();
  if (devices < 1) { std::fprintf(stderr, "no CUDA device: solvempc_b200 has no CPU fallback\n"); return 2; }
  const int shards = argc > 3 ? std::atoi(argv[3]) : devices;
  smpc_settings st;
  smpc_default_settings(&st);
  st.eps_abs = st.eps_rel = 1e-5;
  std::mt19937_64 rng(7);
  std::normal_distribution<double> nrm(0.0, 1.0);
  std::vector<double> X(4 * (size_t)batch), U(batch), ref(batch), U_multi(batch), U_one(batch);
  std::vector<int> st_multi(batch), st_one(batch);
  const double sx[4] = {0.05, 0.2, 0.05, 0.3};
  for (int b = 0; b < batch; ++b) {
    for (int c = 0; c < 4; ++c) X[4 * (size_t)b + c] = sx[c] * nrm(rng);
    U[b] = 2.0 * nrm(rng); ref[b] = 0.25 * nrm(rng);
  }
  std::vector<std::thread> th;
  std::vector<int> rcs(shards, 0);
  for (int r = 0; r < shards; ++r) {
    int lo, hi;
    shard_bounds(batch, shards, r, &lo, &hi);
    if (hi == lo) continue;
    th.emplace_back([&, r, lo, hi]() { rcs[r] = step_shard(json, r % devices, lo, hi, X.data(), U.data(), ref.data(), U_multi.data(), st_multi.data(), &st); });
  }
  for (auto &t : th) t.join();
  for (int rc : rcs) if (rc) return 1;
  if (step_shard(json, 0, 0, batch, X.data(), U.data(), ref.data(), U_one.data(), st_one.data(), &st)) return 1;
  int solved = 0;
  for (int b = 0; b < batch; ++b) solved += st_one[b] == SMPC_SOLVED;
  const bool same = !std::memcmp(U_multi.data(), U_one.data(), sizeof(double) * batch) && !std::memcmp(st_multi.data(), st_one.data(), sizeof(int) * batch);
  std::printf("{\"batch\": %d, \"devices\": %d, \"shards\": %d, \"solved\": %d, \"bitwise_equal\": %s}\n", batch, devices, shards, solved, same ? "true" : "false");
  return same ? 0 : 3;
}
