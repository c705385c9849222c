// mpc_assembly.cu -- on-device condensed-QP assembly and per-step vectors of the MPC layer.
//
// Device restatement of the reference builders (src/ModelPredictiveControlAPI.cpp):
//   setTransformations (cpp:180-208), setH (cpp:247-263), setFVars (cpp:303-307),
//   setLinearConstraints (cpp:326-347), setUpperBound (cpp:360-369), setF (cpp:372-375) and the
//   upper bound sent at cpp:99 -- with mpcWindow / N_S (h:26-30) as run-time N / nx.
// One CTA assembles one plant, so the same kernel serves the shared plant (1 CTA) and per-instance
// linearised plants (one CTA each, BASELINE config 4).
// Reference quirks kept: S carries K in its first n_state_rows rows only;
// Su strictly-upper = 0; Fu uses diag(LL' Rbar') = R.
#include "classify.cuh"
#include "kernels.cuh"

namespace smpc {

constexpr int kMaxNx = 16;

__global__ void __launch_bounds__(256) mpc_assemble_kernel(MpcDims d, int plants, const double *__restrict__ Ad_all,
                                                           const double *__restrict__ Bd_all,
                                                           const double *__restrict__ Cd,
                                                           const double *__restrict__ K, MpcMatsDev o) {
  const int p = blockIdx.x;
  if (p >= plants) return;
  const int N = d.N, nx = d.nx, tid = threadIdx.x, nt = blockDim.x;
  __shared__ double Ak[kMaxNx * kMaxNx], An[kMaxNx * kMaxNx], Ad[kMaxNx * kMaxNx], Bd[kMaxNx], row[kMaxNx];
  extern __shared__ double pc[];  // N prefix sums of CAB
  double *H = o.H + (size_t)p * N * N, *Gbar = o.Gbar + (size_t)p * 2 * N * N, *Fx = o.Fx + (size_t)p * N * nx;
  double *Fu = o.Fu + (size_t)p * N, *Fr = o.Fr + (size_t)p * N * N, *Sbar = o.Sbar + (size_t)p * 2 * N * nx;
  double *Ku = o.Ku + (size_t)p * 2 * N, *W0 = o.W0 + (size_t)p * 2 * N, *Sx = o.Sx + (size_t)p * N * nx;
  double *Su = o.Su + (size_t)p * N * N, *CAB = o.CAB + (size_t)p * N;
  for (int e = tid; e < nx * nx; e += nt) { Ad[e] = Ad_all[(size_t)p * nx * nx + e]; Ak[e] = (e / nx == e % nx) ? 1.0 : 0.0; }
  for (int e = tid; e < nx; e += nt) Bd[e] = Bd_all[(size_t)p * nx + e];
  __syncthreads();
  // cpp:187-190: CAB[i] = Cd Ad^i Bd ; Sx[i] = Cd Ad^(i+1)
  for (int i = 0; i < N; ++i) {
    if (tid < nx) { double s = 0; for (int k = 0; k < nx; ++k) s += Cd[k] * Ak[k * nx + tid]; row[tid] = s; }
    for (int e = tid; e < nx * nx; e += nt) { int r = e / nx, c = e % nx; double s = 0; for (int k = 0; k < nx; ++k) s += Ak[r * nx + k] * Ad[k * nx + c]; An[e] = s; }
    __syncthreads();
    if (tid == 0) { double s = 0; for (int k = 0; k < nx; ++k) s += row[k] * Bd[k]; CAB[i] = s; pc[i] = (i ? pc[i - 1] : 0.0) + s; }
    for (int e = tid; e < nx * nx; e += nt) Ak[e] = An[e];
    __syncthreads();
    if (tid < nx) { double s = 0; for (int k = 0; k < nx; ++k) s += Cd[k] * Ak[k * nx + tid]; Sx[i * nx + tid] = s; }
    __syncthreads();
  }
  // cpp:197-201: Su(i,j) = sum_{k<=i-j} CAB[k] (j<=i), 0 above the diagonal
  for (int e = tid; e < N * N; e += nt) { int i = e / N, j = e % N; Su[e] = j <= i ? pc[i - j] : 0.0; }
  // cpp:185,208: Sbar = [S; -S], S rows < n_state_rows = K ; cpp:332-335: Gbar = [K0 LL; -K0 LL] ; cpp:364-368
  const int ns = d.n_state_rows < N ? d.n_state_rows : N;
  for (int e = tid; e < N * nx; e += nt) { int i = e / nx, c = e % nx; double v = i < ns ? K[c] : 0.0; Sbar[e] = v; Sbar[(size_t)N * nx + e] = -v; }
  for (int e = tid; e < N * N; e += nt) { int i = e / N, j = e % N; double v = j <= i ? K[0] : 0.0; Gbar[e] = v; Gbar[(size_t)N * N + e] = -v; }
  for (int i = tid; i < N; i += nt) { Ku[i] = -K[0]; Ku[N + i] = K[0]; W0[i] = d.u_limit; W0[N + i] = d.u_limit; }
  __syncthreads();
  // cpp:250-251: H = sym(2 (LL' Rbar LL + RbarD + Su' Qbar Su)), (LL'LL)(i,j) = N - max(i,j)
  for (int e = tid; e < N * N; e += nt) {
    int i = e / N, j = e % N;
    if (j < i) continue;
    double s = 0;
    for (int k = j; k < N; ++k) s += Su[(size_t)k * N + i] * Su[(size_t)k * N + j];  // rows k < max(i,j) contribute 0
    double v = 2.0 * (d.R * (double)(N - j) + (i == j ? d.RD : 0.0) + d.Q * s);
    H[(size_t)i * N + j] = v; H[(size_t)j * N + i] = v;
  }
  // cpp:305-307: Fu = 2 (R 1 + Q Su' Su(:,0)) ; Fr = -2 Q Su' ; Fx = 2 Q Su' Sx
  for (int i = tid; i < N; i += nt) {
    double s = 0;
    for (int k = i; k < N; ++k) s += Su[(size_t)k * N + 0] * Su[(size_t)k * N + i];
    Fu[i] = 2.0 * (d.R + d.Q * s);
  }
  double *FrT = o.FrT + (size_t)p * N * N;
  for (int e = tid; e < N * N; e += nt) { int i = e / N, j = e % N; const double v = -2.0 * d.Q * Su[(size_t)j * N + i]; Fr[e] = v; FrT[(size_t)j * N + i] = v; }
  for (int e = tid; e < N * nx; e += nt) {
    int i = e / nx, c = e % nx; double s = 0;
    for (int k = i; k < N; ++k) s += Su[(size_t)k * N + i] * Sx[k * nx + c];
    Fx[e] = 2.0 * d.Q * s;
  }
}

// One warp per instance: f (cpp:374) and ub (cpp:99).
__global__ void mpc_step_vectors_kernel(MpcDims d, int B, int per_instance, MpcMatsDev mt, const double *__restrict__ X,
                                        const double *__restrict__ U, const double *__restrict__ ref,
                                        double *__restrict__ f, double *__restrict__ ub) {
  const int lane = threadIdx.x & 31, b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int N = d.N, nx = d.nx;
  const size_t p = per_instance ? (size_t)b : 0;
  const double *__restrict__ Fx = mt.Fx + p * N * nx, *__restrict__ Fu = mt.Fu + p * N, *__restrict__ FrT = mt.FrT + p * N * N;
  const double *Sbar = mt.Sbar + p * 2 * N * nx, *Ku = mt.Ku + p * 2 * N, *W0 = mt.W0 + p * 2 * N;
  const double *x = X + (size_t)b * nx;
  const double u = U[b], r = ref[b];
  for (int i = lane; i < N; i += 32) {
    double s = 0;
    for (int c = 0; c < nx; ++c) s += Fx[i * nx + c] * x[c];
    s += Fu[i] * u;
    double rr = 0;
#pragma unroll 8
    for (int j = i; j < N; ++j) rr += __ldg(FrT + (size_t)j * N + i) * r;   // Fr(i,j) = 0 for j < i; same order of the sum, coalesced loads
    f[(size_t)b * N + i] = s + rr;
  }
  for (int i = lane; i < 2 * N; i += 32) {
    double s = 0;
    for (int c = 0; c < nx; ++c) s += Sbar[i * nx + c] * x[c];
    ub[(size_t)b * 2 * N + i] = (W0[i] + s) + Ku[i] * u;
  }
}

// Shared plant: one warp carries G consecutive instances through the same sweep, so every operator entry is loaded once per G
// instances (the sweep of Fr is L1-bandwidth bound otherwise).  Per instance the operations and their order are those of
// mpc_step_vectors_kernel: same bits.
template <int G>
__global__ void mpc_step_vectors_shared_kernel(MpcDims d, int B, MpcMatsDev mt, const double *__restrict__ X,
                                               const double *__restrict__ U, const double *__restrict__ ref,
                                               double *__restrict__ f, double *__restrict__ ub) {
  const int lane = threadIdx.x & 31, b0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * G;
  if (b0 >= B) return;
  const int N = d.N, nx = d.nx;
  const double *__restrict__ Fx = mt.Fx, *__restrict__ Fu = mt.Fu, *__restrict__ FrT = mt.FrT;
  const double *__restrict__ Sbar = mt.Sbar, *__restrict__ Ku = mt.Ku, *__restrict__ W0 = mt.W0;
  size_t bg[G];
  double u[G], r[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    bg[g] = (size_t)min(b0 + g, B - 1);      // the tail warp repeats the last instance (stores are guarded)
    u[g] = U[bg[g]]; r[g] = ref[bg[g]];
  }
  for (int i = lane; i < N; i += 32) {
    double s[G], rr[G];
#pragma unroll
    for (int g = 0; g < G; ++g) { s[g] = 0.0; rr[g] = 0.0; }
    for (int c = 0; c < nx; ++c) {
      const double fx = Fx[i * nx + c];
#pragma unroll
      for (int g = 0; g < G; ++g) s[g] += fx * X[bg[g] * nx + c];
    }
    const double fu = Fu[i];
#pragma unroll
    for (int g = 0; g < G; ++g) s[g] += fu * u[g];
#pragma unroll 4
    for (int j = i; j < N; ++j) {
      const double v = __ldg(FrT + (size_t)j * N + i);   // Fr(i,j) = 0 for j < i
#pragma unroll
      for (int g = 0; g < G; ++g) rr[g] += v * r[g];
    }
#pragma unroll
    for (int g = 0; g < G; ++g)
      if (b0 + g < B) f[bg[g] * N + i] = s[g] + rr[g];
  }
  for (int i = lane; i < 2 * N; i += 32) {
    double s[G];
#pragma unroll
    for (int g = 0; g < G; ++g) s[g] = 0.0;
    for (int c = 0; c < nx; ++c) {
      const double sb = Sbar[i * nx + c];
#pragma unroll
      for (int g = 0; g < G; ++g) s[g] += sb * X[bg[g] * nx + c];
    }
    const double w0 = W0[i], ku = Ku[i];
#pragma unroll
    for (int g = 0; g < G; ++g)
      if (b0 + g < B) ub[bg[g] * 2 * N + i] = (w0 + s[g]) + ku * u[g];
  }
}

// The same for a shared plant with N <= 16 (n = N <= 16, m = 2N <= 32: lane = row), fused with the scheduling pre-pass of
// the small-QP kernels: the warp that produced f and ub classifies its instance at once.
__global__ void mpc_step_classify_kernel(MpcDims d, int B, MpcMatsDev mt, const double *__restrict__ X,
                                         const double *__restrict__ U, const double *__restrict__ ref,
                                         double *__restrict__ f, double *__restrict__ ub, SmallPackDev K, SharedPlanDev P,
                                         int *counts, int *lists, double *__restrict__ keepX, double *__restrict__ keepU,
                                         double *__restrict__ keepRef) {
  // let the ADMM kernel behind this one start its prologue now (programmatic dependent launch; it waits for this grid's
  // completion before it reads f, ub and the scheduling lists), then wait for the set_state gather kernel in front
  asm volatile("griddepcontrol.launch_dependents;");
  {
    // the constant operators (plant matrices, scheduling operators) do not depend on the kernel in front: pull their lines
    // into L2 while it is still running (the bench flushes L2 between steps, a real loop evicts them with other work)
    const int N = d.N, nx = d.nx, t = threadIdx.x;
    auto touch = [&](const double *p, int count, int slot) {
      const int line = t - slot;                      // one 128-byte line per thread
      if (line >= 0 && line * 16 < count) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + line * 16));
    };
    touch(mt.Fx, N * nx, 0); touch(mt.FrT, N * N, 16); touch(mt.Sbar, 2 * N * nx, 48); touch(mt.Fu, N, 80); touch(mt.W0, 2 * N, 82);
    touch(mt.Ku, 2 * N, 86); touch(K.V, 256, 96); touch(K.WT, 512, 112); touch(K.D, 16, 144); touch(K.E, 32, 145); touch(P.l0, 2 * N, 147);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int lane = threadIdx.x & 31, b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int N = d.N, nx = d.nx;
  const double *x = X + (size_t)b * nx;
  const double u = U[b], r = ref[b];
  // smpc_mpc_controller_step_from: X, U, ref are the caller's buffers (device or pinned host read over PCIe); the
  // controller's own copies (the members X, U and the reference, h:186-187) are written here instead of by a gather kernel
  if (keepX != nullptr) {
    for (int c = lane; c < nx; c += 32) keepX[(size_t)b * nx + c] = x[c];
    if (lane == 0) { keepU[b] = u; keepRef[b] = r; }
  }
  double f_i = 0.0, ub_r = 0.0;
  if (lane < N) {
    const int i = lane;
    double s = 0;
    for (int c = 0; c < nx; ++c) s += mt.Fx[i * nx + c] * x[c];
    s += mt.Fu[i] * u;
    double rr = 0;
    for (int j = i; j < N; ++j) rr += mt.FrT[(size_t)j * N + i] * r;   // Fr(i,j) = 0 for j < i
    f_i = s + rr;
    f[(size_t)b * N + i] = f_i;
  }
  if (lane < 2 * N) {
    const int i = lane;
    double s = 0;
    for (int c = 0; c < nx; ++c) s += mt.Sbar[i * nx + c] * x[c];
    ub_r = (mt.W0[i] + s) + mt.Ku[i] * u;
    ub[(size_t)b * 2 * N + i] = ub_r;
  }
  const double q_i = __shfl_sync(0xffffffffu, f_i, lane & 15);     // classify_instance wants entry lane & 15 on every lane
  classify_instance(K, P, B, b, lane, q_i, lane < 2 * N ? P.l0[lane] : 0.0, ub_r, counts, lists);
}

// cpp:105: U += dU*[0] -- the reference returns before this line unless the status is SOLVED (cpp:102)
__global__ void mpc_apply_control_kernel(int B, int n, const double *__restrict__ x, const int *__restrict__ status,
                                         double *__restrict__ U) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && status[b] == SMPC_SOLVED) U[b] += x[(size_t)b * n];
}

__global__ void mpc_plant_step_kernel(int B, int nx, int per_instance, const double *__restrict__ Ad,
                                      const double *__restrict__ Bd, double *__restrict__ X,
                                      const double *__restrict__ U) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double *A = Ad + (per_instance ? (size_t)b * nx * nx : 0), *Bv = Bd + (per_instance ? (size_t)b * nx : 0);
  double xo[kMaxNx], xn[kMaxNx];
  for (int c = 0; c < nx; ++c) xo[c] = X[(size_t)b * nx + c];
  for (int r = 0; r < nx; ++r) { double s = 0; for (int c = 0; c < nx; ++c) s += A[r * nx + c] * xo[c]; xn[r] = s + Bv[r] * U[b]; }
  for (int c = 0; c < nx; ++c) X[(size_t)b * nx + c] = xn[c];
}

// Closed-loop driver (the caller of the hot path, reference main loop src/solver.cpp:43-74, with the serial port
// replaced by an on-device reference generator and a synthetic plant).
// ref_b(step) = amplitude * (+1 | -1): square wave of `period` steps, per-instance phase.
__global__ void mpc_square_ref_kernel(int B, double amplitude, int period, const int *__restrict__ phase,
                                      const int *__restrict__ step, double *__restrict__ ref) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int t = (*step + (phase ? phase[b] : 0)) % period;
  ref[b] = (2 * t < period) ? amplitude : -amplitude;
}

// U += dU*[0] (cpp:105, only when SOLVED as cpp:102 returns early otherwise), X <- Ad X + Bd U, statistics, step += 1
__global__ void mpc_advance_kernel(int B, int n, int nx, int per_instance, const double *__restrict__ Ad,
                                   const double *__restrict__ Bd, const double *__restrict__ x,
                                   const int *__restrict__ status, const int *__restrict__ iter, double *__restrict__ X,
                                   double *__restrict__ U, unsigned long long *stats, int *step) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  int bad = 0, it = 0;
  if (b < B) {
    const bool ok = status[b] == SMPC_SOLVED;
    bad = !ok; it = iter[b];
    double u = U[b];
    if (ok) u += x[(size_t)b * n];
    U[b] = u;
    const double *A = Ad + (per_instance ? (size_t)b * nx * nx : 0), *Bv = Bd + (per_instance ? (size_t)b * nx : 0);
    double xo[kMaxNx], xn[kMaxNx];
    for (int c = 0; c < nx; ++c) xo[c] = X[(size_t)b * nx + c];
    for (int r = 0; r < nx; ++r) { double s = 0; for (int c = 0; c < nx; ++c) s += A[r * nx + c] * xo[c]; xn[r] = s + Bv[r] * u; }
    for (int c = 0; c < nx; ++c) X[(size_t)b * nx + c] = xn[c];
  }
  for (int o = 16; o > 0; o >>= 1) { bad += __shfl_xor_sync(0xffffffffu, bad, o); it += __shfl_xor_sync(0xffffffffu, it, o); }
  if ((threadIdx.x & 31) == 0) {
    if (bad) atomicAdd(stats, (unsigned long long)bad);
    atomicAdd(stats + 1, (unsigned long long)it);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(step, 1);
}

__global__ void mpc_copy_state_kernel(int B, int nx, const double *__restrict__ X, const double *__restrict__ U,
                                      const double *__restrict__ ref, double *__restrict__ dX, double *__restrict__ dU,
                                      double *__restrict__ dref) {
  asm volatile("griddepcontrol.launch_dependents;");   // the step-vector kernel may start; it waits for this grid's completion
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (X && e < B * nx) dX[e] = X[e];
  if (e < B) {
    if (U) dU[e] = U[e];
    if (ref) dref[e] = ref[e];
  }
}
cudaError_t launch_mpc_copy_state(int B, int nx, const double *X, const double *U, const double *ref, double *dX, double *dU,
                                  double *dref, cudaStream_t stream) {
  if (!X && !U && !ref) return cudaSuccess;
  const int count = X ? B * nx : B;
  mpc_copy_state_kernel<<<(count + 255) / 256, 256, 0, stream>>>(B, nx, X, U, ref, dX, dU, dref);
  return cudaGetLastError();
}

__global__ void mpc_export_kernel(int B, const double *__restrict__ U, const int *__restrict__ status, double *__restrict__ outU,
                                  int *__restrict__ outStatus) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B) return;
  if (outU) outU[e] = U[e];
  if (outStatus) outStatus[e] = status[e];
}
cudaError_t launch_mpc_export(int B, const double *U, const int *status, double *outU, int *outStatus, cudaStream_t stream) {
  mpc_export_kernel<<<(B + 255) / 256, 256, 0, stream>>>(B, U, status, outU, outStatus);
  return cudaGetLastError();
}

cudaError_t launch_mpc_square_ref(int B, double amplitude, int period, const int *phase, const int *step, double *ref,
                                  cudaStream_t stream) {
  mpc_square_ref_kernel<<<(B + 255) / 256, 256, 0, stream>>>(B, amplitude, period, phase, step, ref);
  return cudaGetLastError();
}
cudaError_t launch_mpc_advance(int B, int n, int nx, int per_instance, const double *Ad, const double *Bd, const double *x,
                               const int *status, const int *iter, double *X, double *U, unsigned long long *stats, int *step,
                               cudaStream_t stream) {
  if (nx > kMaxNx) return cudaErrorInvalidValue;
  mpc_advance_kernel<<<(B + 127) / 128, 128, 0, stream>>>(B, n, nx, per_instance, Ad, Bd, x, status, iter, X, U, stats, step);
  return cudaGetLastError();
}

cudaError_t launch_mpc_assemble(const MpcDims &d, int plants, const double *Ad, const double *Bd, const double *Cd,
                                const double *K, const MpcMatsDev &out, cudaStream_t stream) {
  if (d.nx > kMaxNx || d.nx < 1 || d.N < 1) return cudaErrorInvalidValue;
  mpc_assemble_kernel<<<plants, 256, d.N * sizeof(double), stream>>>(d, plants, Ad, Bd, Cd, K, out);
  return cudaGetLastError();
}
cudaError_t launch_mpc_step_vectors(const MpcDims &d, int B, int per_instance, const MpcMatsDev &mats, const double *X,
                                    const double *U, const double *ref, double *f, double *ub, cudaStream_t stream) {
  const int wpc = 8;
  if (!per_instance && d.N > 32) {   // long horizons: four instances per warp share the operator loads
    constexpr int G = 4;
    const int warps = (B + G - 1) / G;
    mpc_step_vectors_shared_kernel<G><<<(warps + wpc - 1) / wpc, wpc * 32, 0, stream>>>(d, B, mats, X, U, ref, f, ub);
    return cudaGetLastError();
  }
  mpc_step_vectors_kernel<<<(B + wpc - 1) / wpc, wpc * 32, 0, stream>>>(d, B, per_instance, mats, X, U, ref, f, ub);
  return cudaGetLastError();
}
cudaError_t launch_mpc_step_classify(const MpcDims &d, int B, const MpcMatsDev &mats, const double *X, const double *U,
                                     const double *ref, double *f, double *ub, const SmallPackDev &K, const SharedPlanDev &P,
                                     int *counts, int *lists, double *keepX, double *keepU, double *keepRef, cudaStream_t stream) {
  if (d.N > 16 || P.n != d.N || P.m != 2 * d.N) return cudaErrorInvalidValue;
  const int wpc = 8;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((B + wpc - 1) / wpc); cfg.blockDim = dim3(wpc * 32); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, mpc_step_classify_kernel, d, B, mats, X, U, ref, f, ub, K, P, counts, lists, keepX, keepU, keepRef);
}
cudaError_t launch_mpc_apply_control(int B, int n, const double *x, const int *status, double *U, cudaStream_t stream) {
  mpc_apply_control_kernel<<<(B + 255) / 256, 256, 0, stream>>>(B, n, x, status, U);
  return cudaGetLastError();
}
cudaError_t launch_mpc_plant_step(int B, int nx, int per_instance, const double *Ad, const double *Bd, double *X,
                                  const double *U, cudaStream_t stream) {
  if (nx > kMaxNx) return cudaErrorInvalidValue;
  mpc_plant_step_kernel<<<(B + 127) / 128, 128, 0, stream>>>(B, nx, per_instance, Ad, Bd, X, U);
  return cudaGetLastError();
}

}  // namespace smpc
