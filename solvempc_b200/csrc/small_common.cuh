// small_common.cuh -- helpers shared by the small-QP kernels (admm_shared_small.cu, admm_shared_small_fused.cu): volatile
// shared-memory accessors, exact warp reductions, the work-queue layout of the persistent grids and its release protocol.
#pragma once
#include <cstdint>

#include "classify.cuh"
#include "device_types.cuh"

namespace smpc {
namespace {
constexpr int NP = 16, MP = 32, KH = (NP + MP) / 2;   // padded sizes; 24 concatenated entries per half-warp
constexpr unsigned kFull = 0xffffffffu;
constexpr int kClasses = kSchedClasses;   // difficulty classes of the scheduling pre-pass (classify.cuh)

__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
// exact max over the warp of NON-NEGATIVE doubles: compare the IEEE bit patterns as two 32-bit halves (REDUX)
__device__ __forceinline__ double wmax_nn(double v) {
  const unsigned hi = (unsigned)__double2hiint(v);
  const unsigned mh = __reduce_max_sync(kFull, hi);
  const unsigned lo = hi == mh ? (unsigned)__double2loint(v) : 0u;
  const unsigned ml = __reduce_max_sync(kFull, lo);
  return __hiloint2double((int)mh, (int)ml);
}
// |x| as an IEEE bit pattern; for non-negative doubles the order of the patterns is the order of the values (NaN on top, so it
// propagates like in the oracle's max loops), and max / compare run on the integer pipe instead of DSETP + NaN fix-ups
typedef unsigned long long ull;
__device__ __forceinline__ ull abits(double x) { return (ull)__double_as_longlong(x) & 0x7fffffffffffffffULL; }
__device__ __forceinline__ ull umax2(ull a, ull b) { return a > b ? a : b; }
__device__ __forceinline__ double wmax_bits(ull v) {
  const unsigned hi = (unsigned)(v >> 32);
  const unsigned mh = __reduce_max_sync(kFull, hi);
  const unsigned lo = hi == mh ? (unsigned)v : 0u;
  const unsigned ml = __reduce_max_sync(kFull, lo);
  return __hiloint2double((int)mh, (int)ml);
}
// x / y for y > 0 (rho estimate): reciprocal + one correction step, within an ulp of the IEEE quotient
__device__ __forceinline__ double fast_div(double x, double y) {
  const double r = __drcp_rn(y), q = x * r;
  return fma(fma(-y, q, x), r, q);
}
// v < lo ? lo : (v > hi ? hi : v) as two selects (the compiler otherwise turns the ternaries into divergent branches)
__device__ __forceinline__ double clip_sel(double v, double lo, double hi) {
  double r;
  asm("{\n\t.reg .pred p, q;\n\tsetp.gt.f64 q, %1, %3;\n\tselp.f64 %0, %3, %1, q;\n\tsetp.lt.f64 p, %1, %2;\n\tselp.f64 %0, %2, %0, p;\n\t}"
      : "=&d"(r) : "d"(v), "d"(lo), "d"(hi));
  return r;
}
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
// queue layout: [0] work counter, [1..kClasses] class sizes, [1 + kClasses] warps / CTAs that finished,
// [kQHard] counter of the hardest class's quiet share, [kQSeen] SMs seen, [kQRank + smid] CTAs arrived on an SM,
// [kQTicket + smid] 1 + arrival order of the SM (quiet-SM scheduling of admm_shared_small_kernel)
constexpr int kQHard = 2 + kClasses, kQSeen = 3 + kClasses, kQRank = 16, kSmSlots = 512, kQTicket = kQRank + kSmSlots;
constexpr int kQueueInts = kQTicket + kSmSlots;
// the last warp to leave clears the whole queue block (all lanes of the warp call this)
__device__ __forceinline__ void release_queue_warp(int *queue, int participants, int lane) {
  int last = 0;
  if (lane == 0) {
    __threadfence();
    const int left = atomicAdd(queue + 1 + kClasses, 1);
    SMPC_DBG(left >= 0 && left < participants, "warps that left the persistent grid");
    last = left == participants - 1;
  }
  last = __shfl_sync(kFull, last, 0);
  if (last) {
    for (int k = lane; k < kQueueInts; k += 32) queue[k] = 0;
    __threadfence();
  }
}
__device__ __forceinline__ void release_queue(int *queue, int participants) {
  __threadfence();
  if (atomicAdd(queue + 1 + kClasses, 1) == participants - 1) {
#pragma unroll
    for (int k = 0; k <= 1 + kClasses; ++k) queue[k] = 0;
    __threadfence();
  }
}
__device__ __forceinline__ double rho_row(int ct, double rho) {
  return ct == 0 ? rho : (ct == 1 ? kRhoEqOverIneq * rho : kRhoMin);
}


}  // namespace
}  // namespace smpc
