// classify.cuh -- scheduling pre-pass of the small-QP kernels (shared by the stand-alone classify kernel and the fused
// step-vector kernel of the MPC layer).
//
// The expensive instances are the MARGINALLY constrained ones: with x_unc the minimiser of the cost alone,
// key = max_r ((A̅ x_unc)_r - ū_r, l̄_r - (A̅ x_unc)_r) is slightly positive for them (measured on config 2: sorting by
// key / ||bounds|| ascending puts 61 of the 64 instances that need >= 100 iterations among the first 207).  In plan
// coordinates A̅ x_unc = -W q̂ (S^-1 = V V'), i.e. one z-phase product.  Classes (processed in this order): ratio in
// (0, .02], (.02, .05], (.05, .15], > .15 (saturated), <= 0 (unconstrained optimum feasible).  The order changes nothing but
// the schedule: every instance is solved exactly as before.
#pragma once
#include "device_types.cuh"

namespace smpc {

constexpr int kSchedClasses = 5;

// One warp, one instance b (n <= 16, m <= 32).  q_i: UNSCALED gradient entry i = lane & 15 (any value for i >= n);
// lo_r / hi_r: UNSCALED bounds of row r = lane (any value for r >= m).  counts[kSchedClasses], lists[kSchedClasses][B].
__device__ __forceinline__ void classify_instance(const SmallPackDev &K, const SharedPlanDev &P, int B, int b, int lane,
                                                  double q_i, double lo_r, double hi_r, int *counts, int *lists) {
  constexpr int NP = 16, MP = 32;
  constexpr unsigned kFullMask = 0xffffffffu;
  const int n = P.n, m = P.m, i = lane & 15, r = lane;
  const double qb = i < n ? P.c * (K.D[i] * q_i) : 0.0;
  double qh = 0.0;
#pragma unroll
  for (int k = 0; k < NP; ++k) qh = fma(K.V[k * NP + i], __shfl_sync(kFullMask, qb, k), qh);
  double zu = 0.0;
#pragma unroll
  for (int k = 0; k < NP; ++k) zu = fma(-K.WT[k * MP + r], __shfl_sync(kFullMask, qh, k), zu);
  double key = -1e300, ref = 0.0;
  if (r < m) {
    const double lo = K.E[r] * lo_r, hi = K.E[r] * hi_r;
    if (hi < kInfty * kMinScaling) { key = fmax(key, zu - hi); ref = fmax(ref, fabs(hi)); }
    if (lo > -kInfty * kMinScaling) { key = fmax(key, lo - zu); ref = fmax(ref, fabs(lo)); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { key = fmax(key, __shfl_xor_sync(kFullMask, key, o)); ref = fmax(ref, __shfl_xor_sync(kFullMask, ref, o)); }
  if (lane == 0) {
    const double ratio = key / fmax(ref, 1e-300);
    const int cls = !(key > 0.0) ? 4 : (ratio <= 0.02 ? 0 : (ratio <= 0.05 ? 1 : (ratio <= 0.15 ? 2 : 3)));
    const int slot = atomicAdd(counts + cls, 1);
    SMPC_DBG(slot >= 0 && slot < B, "class list slot");
    if (slot < B) lists[(size_t)cls * B + slot] = b;   // (slot >= B only if stale counters survived a failed step)
  }
}

}  // namespace smpc
