// kernels.cuh -- launchers of the CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include "../../include/solvempc_b200.h"
#include "device_types.cuh"

namespace smpc {

// admm_shared_generic.cu
cudaError_t launch_admm_shared_generic(const SharedPlanDev &P, const BatchDev &Bt, const SettingsDev &S,
                                       cudaStream_t stream);
cudaError_t launch_fill(double *p, double v, size_t count, cudaStream_t stream);
// xspace != 0: the handle's kernel keeps the scaled iterate x̄ itself as its state (tile kernel, x-space variant), not xi = V^-1 x̄
cudaError_t launch_warm_start(const SharedPlanDev &P, int B, const double *x, const double *y, double *xi,
                              double *z, double *ys, cudaStream_t stream, bool xspace = false);

// admm_shared_small.cu : register-resident kernel for small QPs (n <= 16, m <= 32)
bool small_kernel_supports(int n, int m);
size_t small_pack_doubles();
int small_queue_ints();
int small_sched_classes();
// lists != nullptr: longest-expected-first order; classified: the class lists were already filled for this solve (fused
// into the MPC layer's step-vector kernel), so the stand-alone pre-pass is skipped
cudaError_t launch_admm_shared_small(const SmallPackDev &K, const SharedPlanDev &P, const BatchDev &Bt,
                                     const SettingsDev &S, int *queue, int *lists, bool classified, int num_sms, cudaStream_t stream);

// admm_shared_small_fused.cu : one-phase kernel for paired rows [G; -G] (mp <= 16); launch_admm_shared_small dispatches to it
bool small_fused_supports(const SmallPackDev &K, const SharedPlanDev &P);
cudaError_t launch_admm_shared_small_fused(const SmallPackDev &K, const SharedPlanDev &P, const BatchDev &Bt, const SettingsDev &S,
                                           int *queue, int *lists, int num_sms, cudaStream_t stream);

// DMMA variant of the small-QP kernel: 8 QPs per CTA of four warps (same packs, same queue / lists)
cudaError_t launch_admm_shared_small_mma(const SmallPackDev &K, const SharedPlanDev &P, const BatchDev &Bt,
                                         const SettingsDev &S, int *queue, int *lists, bool classified, int num_sms, cudaStream_t stream);

// admm_shared_tile.cu : DMMA tile kernel for mid-size QPs (8 or 16 QPs per CTA, iterates in shared memory)
bool tile_kernel_supports(int n, int m);
int tile_kernel_nb(int n, int m);                 // 8-slot blocks per tile that fit shared memory (0 = unsupported)
size_t tile_smem_bytes(int n, int m, int nb);
cudaError_t launch_admm_shared_tile(const TilePackDev &K, const SharedPlanDev &P, const BatchDev &Bt, const SettingsDev &S,
                                    int *queue, int nb, int num_sms, cudaStream_t stream);

// admm_instance.cu : per-instance regime (own P_i, A_i per QP; batched Cholesky in shared memory)
cudaError_t launch_ruiz_instance(const InstanceDataDev &I, int iters, cudaStream_t stream);
// prepare != 0: create-time pass of the register-operator kernel (fills I.S0, I.T, I.Minv0); no-op for other sizes
cudaError_t launch_admm_instance(const InstanceDataDev &I, const BatchDev &Bt, const SettingsDev &S, cudaStream_t stream, int prepare = 0);
bool instance_reg_supports(int n, int m);
cudaError_t launch_instance_pairs(const InstanceDataDev &I, int *flag, cudaStream_t stream);   // *flag &= all instances are [G; -G]
cudaError_t launch_warm_start_instance(const InstanceDataDev &I, const double *x, const double *y, double *xs, double *z,
                                       double *ys, cudaStream_t stream);
bool instance_kernel_supports(int n, int m);
// admm_instance_pair.cu : two warps per QP for [G; -G] instances, TMA-staged operands, persistent CTAs
bool instance_pair_supports(int n, int m);
size_t instance_pair_pack_doubles(int n);
cudaError_t launch_admm_instance_pair(const InstanceDataDev &I, const BatchDev &Bt, const SettingsDev &S, int num_sms, cudaStream_t stream,
                                      int prepare);

// polish.cu : OSQP's solution polishing after the ADMM kernel (opt-in), one warp per SOLVED instance
size_t polish_scratch_doubles(int n, int m, int num_sms);   // global workspace the kernel needs (0: shared memory suffices)
cudaError_t launch_polish(const PolishDataDev &P, const BatchDev &Bt, const SettingsDev &S, double delta, int refine,
                          int *status_polish, double *scratch, int num_sms, cudaStream_t stream);

// mpc_assembly.cu
struct MpcDims { int N, nx, n_state_rows; double Q, R, RD, u_limit; };
struct MpcMatsDev {   // per plant (index p): all row-major
  double *H, *Gbar, *Fx, *Fu, *Fr, *Sbar, *Ku, *W0, *Sx, *Su, *CAB;
  double *FrT;   // Fr transposed (FrT[j][i] = Fr(i,j)): the per-step gradient sweeps Fr by rows with lane = row, coalesced through FrT
};
cudaError_t launch_mpc_assemble(const MpcDims &d, int plants, const double *Ad, const double *Bd,
                                const double *Cd, const double *K, const MpcMatsDev &out, cudaStream_t stream);
// f = Fx X + Fu U + Fr ref ; ub = W0 + Sbar X + Ku U  for the whole batch (mats shared when stride 0)
cudaError_t launch_mpc_step_vectors(const MpcDims &d, int B, int per_instance, const MpcMatsDev &mats,
                                    const double *X, const double *U, const double *ref, double *f,
                                    double *ub, cudaStream_t stream);
// set_state from device buffers: X, U, ref (any may be NULL) -> the controller's own copies, one launch
cudaError_t launch_mpc_copy_state(int B, int nx, const double *X, const double *U, const double *ref, double *dX, double *dU,
                                  double *dref, cudaStream_t stream);
// step vectors of a shared plant with N <= 16 AND the small-QP kernels' scheduling pre-pass in one launch (classify.cuh):
// counts = queue + 1, lists as launch_admm_shared_small takes them; lower bounds are the plan's (P.l0)
cudaError_t launch_mpc_step_classify(const MpcDims &d, int B, const MpcMatsDev &mats, const double *X, const double *U,
                                     const double *ref, double *f, double *ub, const SmallPackDev &K, const SharedPlanDev &P,
                                     int *counts, int *lists, double *keepX, double *keepU, double *keepRef, cudaStream_t stream);
// (keepX / keepU / keepRef != nullptr: X, U, ref are the caller's buffers and are also copied into the controller's own)
// results to (pinned, device-mapped) host memory in one launch: U and the per-instance status (either may be NULL)
cudaError_t launch_mpc_export(int B, const double *U, const int *status, double *outU, int *outStatus, cudaStream_t stream);
// U += dU[0]
cudaError_t launch_mpc_apply_control(int B, int n, const double *x, const int *status, double *U, cudaStream_t stream);
// X <- Ad X + Bd U
cudaError_t launch_mpc_plant_step(int B, int nx, int per_instance, const double *Ad, const double *Bd,
                                  double *X, const double *U, cudaStream_t stream);

// closed-loop driver pieces: square-wave reference generator; U += dU[0], plant step, statistics, step counter
cudaError_t launch_mpc_square_ref(int B, double amplitude, int period, const int *phase, const int *step, double *ref,
                                  cudaStream_t stream);
cudaError_t launch_mpc_advance(int B, int n, int nx, int per_instance, const double *Ad, const double *Bd, const double *x,
                               const int *status, const int *iter, double *X, double *U, unsigned long long *stats, int *step,
                               cudaStream_t stream);

// mimo_assembly.cu : multi-input condensed MPC (BASELINE config 3)
struct MimoDims { int N, nx, nu; };
struct MimoMatsDev { double *H, *A, *ub, *Fx, *Fr, *Su, *Sx; };   // row-major; shapes in include/solvempc_b200.h
cudaError_t launch_mimo_assemble(const MimoDims &d, const double *Ad, const double *Bd, const double *Q, const double *R,
                                 const double *umin, const double *umax, double *AB, double *AP, const MimoMatsDev &out,
                                 cudaStream_t stream);
cudaError_t launch_mimo_step_vectors(const MimoDims &d, int B, const double *Fx, const double *Fr, const double *X0,
                                     const double *Xr, double *q, cudaStream_t stream);
cudaError_t launch_mimo_first_move(int B, int n, int nu, const double *x, const int *status, double *u0, cudaStream_t stream);

}  // namespace smpc
