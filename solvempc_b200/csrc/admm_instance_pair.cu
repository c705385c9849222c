// admm_instance_pair.cu -- per-instance regime (every QP has its OWN P_i, A_i: BASELINE config 4) for the reference's
// constraint form A = [G; -G] (two-sided limit as row pairs, src/ModelPredictiveControlAPI.cpp:335), n <= 32, m = 2 mp <= 64.
// Replaces what the reference reaches through solver.initSolver() / solver.solve() (cpp:64,102) for distinct P, A.
//
// One CTA of TWO warps owns one QP at a time (persistent CTAs, device ticket queue: no CTA waits for a slow neighbour):
//   x-warp (warp 0), lane i:  x_i, q_i and row i of M^-1          in registers
//   z-warp (warp 1), lane r:  z, y of rows r and r + mp, row r of K = G M^-1   in registers
//   both warps:               half a row of G' each (lane i: G(16 w .. 16 w + 15, i))
// One OSQP iteration (x-space, same steps and order as osqp_solve):
//   phase 1   rhs = sigma x - q + G' wd,           wd = (rho_vec z - y)_top - (rho_vec z - y)_bot      (two half dot products)
//   phase 2   x~ = M^-1 rhs (x-warp)   ||   z~_top = K rhs = G x~ (z-warp), z~_bot = -z~_top            (independent)
//   then relaxation, clip to [l, u], dual update in the z-warp's registers; x update in the x-warp's.
// Three CTA barriers of 64 threads per iteration; 96 DFMA + 24 LDS.128 per thread-pair-iteration; no operator is read from
// shared memory inside the iteration.
//
// Operands are staged by TMA: every instance has a prepared PACK in HBM (written once at create time by the same kernel
// with prepare = 1, as osqp_setup factors once): G' (padded rows of 34 doubles), M(rho)^-1 and K for the rho it was last
// factored with, and the rho-independent split M(rho) = S0 + rho T.  While a QP iterates, thread 0 has already drawn the
// NEXT ticket, issued cp.async.bulk (global -> the second G' stage in shared memory, completion on an mbarrier) and an L2
// prefetch of the rest of that pack, so a solve starts with its operands on chip.
// Refactorisation (every rho update, exactly when OSQP refactors): M = S0 + rho T, left-looking Cholesky in shared memory,
// lane c solves L L' v = e_c in registers (row c of M^-1), then the z-warp forms its rows of K = G M^-1.
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "device_types.cuh"
#include "kernels.cuh"

namespace smpc {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kPN = 32;    // max n and mp
constexpr int kPLD = 34;   // padded row stride in doubles: 16-byte row sweeps of a quarter warp hit 8 different bank groups

__device__ __forceinline__ double wmax(double v) {   // exact max over the warp of NON-NEGATIVE doubles (IEEE bit patterns order them)
  const unsigned hi = (unsigned)__double2hiint(v);
  const unsigned mh = __reduce_max_sync(kFull, hi);
  const unsigned lo = hi == mh ? (unsigned)__double2loint(v) : 0u;
  const unsigned ml = __reduce_max_sync(kFull, lo);
  return __hiloint2double((int)mh, (int)ml);
}
// OSQP's c_min(c_max(v, lo), hi) (a > b ? a : b, then a < b ? a : b: NaN -> lo) as two compare / select pairs: fmin(fmax())
// compiles to DSETP.MAX / MIN + selects + NaN fix-ups, almost twice the instructions
__device__ __forceinline__ double clamp_sel(double v, double lo, double hi) {
  double r;
  asm("{\n\t.reg .pred p, q;\n\tsetp.gt.f64 p, %1, %2;\n\tselp.f64 %0, %1, %2, p;\n\tsetp.lt.f64 q, %0, %3;\n\tselp.f64 %0, %0, %3, q;\n\t}"
      : "=&d"(r) : "d"(v), "d"(lo), "d"(hi));
  return r;
}
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// ---- TMA (1-D bulk copy) + mbarrier, raw PTX (SASS: UBLKCP.S.G, UBLKPF.L2, SYNCS.*)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// sum_{k < 32} reg[k] * vec[k]: the row in registers, vec (zero padded to 32) broadcast from shared memory with LDS.128
__device__ __forceinline__ double dot_reg32(const double (&reg)[kPN], const double *vec) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int k = 0; k < kPN; k += 4) {
    const double2 v0 = *reinterpret_cast<const double2 *>(vec + k), v1 = *reinterpret_cast<const double2 *>(vec + k + 2);
    a0 = fma(reg[k], v0.x, a0); a1 = fma(reg[k + 1], v0.y, a1); a2 = fma(reg[k + 2], v1.x, a2); a3 = fma(reg[k + 3], v1.y, a3);
  }
  return (a0 + a1) + (a2 + a3);
}
__device__ __forceinline__ double dot_reg16(const double (&reg)[16], const double *vec) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int k = 0; k < 16; k += 4) {
    const double2 v0 = *reinterpret_cast<const double2 *>(vec + k), v1 = *reinterpret_cast<const double2 *>(vec + k + 2);
    a0 = fma(reg[k], v0.x, a0); a1 = fma(reg[k + 1], v0.y, a1); a2 = fma(reg[k + 2], v1.x, a2); a3 = fma(reg[k + 3], v1.y, a3);
  }
  return (a0 + a1) + (a2 + a3);
}

// Reciprocal, quotient and square root without the library's out-of-line slow path (a CALL in the middle of the solve makes
// ptxas keep a quarter of the register-resident operator rows in local memory): hardware seed + Newton steps, accurate to
// the last bit or two for the normal, finite operands that occur here (scalings, rho, pivots of a positive definite matrix).
__device__ __forceinline__ double fast_rcp(double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0); r = fma(r, e, r);
  e = fma(-b, r, 1.0); r = fma(r, e, r);
  e = fma(-b, r, 1.0); r = fma(r, e, r);
  return r;
}
__device__ __forceinline__ double fast_div(double a, double b) {
  const double r = fast_rcp(b);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}
__device__ __forceinline__ double fast_sqrt(double d) {   // d > 0
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double g = d * y, h = 0.5 * y;
  double r = fma(-h, g, 0.5); g = fma(g, r, g); h = fma(h, r, h);
  r = fma(-h, g, 0.5); g = fma(g, r, g); h = fma(h, r, h);
  return fma(fma(-g, g, d), h, g);
}

__device__ __forceinline__ int row_class(double lo, double hi) {   // -1 free, 0 inequality, 1 equality (OSQP constr_type)
  return (lo < -kInfty * kMinScaling && hi > kInfty * kMinScaling) ? -1 : ((hi - lo < kRhoTolRow) ? 1 : 0);
}

// indices into the CTA-wide scalar exchange red[]
enum { R_SRP, R_SZ, R_SAX, R_URP, R_UZ, R_UAX, R_PIND, R_PILHS, R_SRD, R_SQ, R_SATY, R_SPX, R_URD, R_UQ, R_UATY, R_UPX, R_OB, R_DIND,
       R_DIQD, R_RARE0, R_RARE1, R_FLAG0, R_FLAG1, R_COUNT };

}  // namespace

__host__ __device__ inline int pair_tri2(int n) { return ((n * (n + 1) / 2) + 1) & ~1; }
// pack of one instance (doubles): G' | M^-1 | K (kPN rows of kPLD each) | S0 | T (packed lower triangles) | V' (kPN rows of kPLD,
// V'[k][i] = V(i, k)) | lambda (kPN) | flag (2): the generalised eigen-decomposition V' S0 V = I, V' T V = diag(lambda) of the
// rho-independent split, so that M(rho)^-1 = V diag(1 / (1 + rho lambda)) V' for every rho (flag = 1: valid)
__host__ __device__ inline size_t pair_pack_stride(int n) { return 4 * (size_t)kPN * kPLD + 2 * (size_t)pair_tri2(n) + kPN + 2; }
size_t instance_pair_pack_doubles(int n) { return pair_pack_stride(n); }
bool instance_pair_supports(int n, int m) { return n >= 1 && n <= kPN && m >= 2 && m % 2 == 0 && m / 2 <= kPN; }
static size_t pair_smem_bytes(int n) { return ((size_t)3 * n * kPLD + 21 * kPN) * sizeof(double) + 4 * sizeof(uint64_t); }

template <int CTAS_PER_SM>
__global__ void __launch_bounds__(64, CTAS_PER_SM) admm_instance_pair_kernel(InstanceDataDev I, BatchDev Bt, SettingsDev S, int *queue, int prepare) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const bool xw = w == 0;
  const int n = I.n, m = I.m, mp = m >> 1, B = Bt.B;
  const int tri = n * (n + 1) / 2, tri2 = pair_tri2(n);
  const size_t pack_stride = pair_pack_stride(n);
  const size_t oVT = 3 * (size_t)kPN * kPLD + 2 * (size_t)tri2, oLam = oVT + (size_t)kPN * kPLD, oFlag = oLam + kPN;
  double *const Gt0 = smem, *const Gt1 = smem + n * kPLD;
  double *Sc = smem + 2 * n * kPLD;
  double *wd = Sc + n * kPLD, *p1 = wd + kPN, *rhs = p1 + kPN, *xs = rhs + kPN, *ubt = xs + kPN, *ubb = ubt + kPN, *lbt = ubb + kPN,
         *lbb = lbt + kPN, *t0 = lbb + kPN, *t1 = t0 + kPN, *evt = t1 + kPN, *evb = evt + kPN, *dl0 = evb + kPN, *dl1 = dl0 + kPN,
         *dl2 = dl1 + kPN, *rvt = dl2 + kPN, *rvb = rvt + kPN, *rit = rvb + kPN, *rib = rit + kPN, *red = rib + kPN;
  // dl0 / dl1: delta_y (top, bottom), dl2: delta_x of the iteration before a check; rvt .. rib: rho_vec and 1 / rho_vec per row
  double *zer = red + kPN;                                    // kPN zeros (never written after the initial clear)
  uint64_t *mbar = reinterpret_cast<uint64_t *>(zer + kPN);   // [0], [1]: the two G' stages
  __shared__ int s_ticket;

  const double alpha = S.alpha, sigma = S.sigma;
  const bool unscale = !S.scaled_termination;
  const uint32_t gt_bytes = (uint32_t)(n * kPLD * sizeof(double));

  if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_fence_init(); }
  for (int e = tid; e < 21 * kPN; e += 64) wd[e] = 0.0;
  __syncthreads();
  // first ticket + its G' stage
  if (tid == 0) {
    const int t = atomicAdd(queue, 1);
    SMPC_DBG(t >= 0, "pair-kernel ticket");
    s_ticket = t;
    if (t < B && !prepare) { fence_proxy_async(); mbar_expect_tx(&mbar[0], gt_bytes); bulk_g2s(Gt0, I.pack + (size_t)t * pack_stride, gt_bytes, &mbar[0]); }
  }
  __syncthreads();
  int cur = s_ticket, stage = 0;
  uint32_t phase0 = 0u, phase1 = 0u;
#ifdef SMPC_PAIR_PROFILE
  long long pf[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pc = clock64();
  int pf_solves = 0, pf_iters = 0, pf_refac = 0;
#define PFI(i) { const long long tt = clock64(); pf[i] += tt - pc; pc = tt; }
#else
#define PFI(i)
#endif

  while (cur < B) {
    const int b = cur;
    __syncthreads();                       // everyone has read s_ticket / finished the previous solve
    if (tid == 0) {                        // draw the next ticket now: its operands travel while this QP is solved
      const int t = atomicAdd(queue, 1);
      SMPC_DBG(t > b, "pair-kernel ticket order");
      s_ticket = t;
      if (t < B && !prepare) {
        const double *pk = I.pack + (size_t)t * pack_stride;
        fence_proxy_async();                                                                       // the other stage was read with LDS until a moment ago
        mbar_expect_tx(&mbar[stage ^ 1], gt_bytes);
        bulk_g2s(stage ? Gt0 : Gt1, pk, gt_bytes, &mbar[stage ^ 1]);
        bulk_prefetch_l2(pk + kPN * kPLD, (uint32_t)((kPN + n) * kPLD * sizeof(double)));          // M^-1 and the first n rows of K
        bulk_prefetch_l2(pk + oVT, (uint32_t)((n * kPLD + 0) * sizeof(double)));                   // V' (rho updates)
        bulk_prefetch_l2(pk + oLam, (uint32_t)((kPN + 2) * sizeof(double)));                       // lambda, flag
        const uintptr_t pa = reinterpret_cast<uintptr_t>(I.P + (size_t)t * n * n) & ~(uintptr_t)15;
        bulk_prefetch_l2(reinterpret_cast<const void *>(pa), (uint32_t)((n * n * sizeof(double) + 15) & ~(size_t)15));   // P̄ (termination checks)
      }
    }
    double *Gt = stage ? Gt1 : Gt0;
    double *pack = I.pack + (size_t)b * pack_stride;
    const double *gP = I.P + (size_t)b * n * n;
    const double c = I.c[b], cinv = fast_rcp(c);
    const bool warm = S.warm_start && !Bt.fresh && !prepare;

    // ---- per-instance vectors into registers
    // four state registers per thread, by role.  x-warp (lane < n): x, q̄, D, -.  z-warp (lane < mp): z_top, z_bot, y_top, y_bot.
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#define X_ s0
#define Q_ s1
#define DV_ s2
#define ZT_ s0
#define ZB_ s1
#define YT_ s2
#define YB_ s3
    int cls_t = 0, cls_b = 0, flag_bad = 0, flag_diff = 0;
    if (tid < 32) { dl0[lane] = 0.0; dl1[lane] = 0.0; dl2[lane] = 0.0; }
    if (xw) {
      if (lane < n) {
        DV_ = I.D[(size_t)b * n + lane];
        Q_ = Bt.q ? c * (DV_ * Bt.q[(size_t)b * n + lane]) : 0.0;
        X_ = warm ? Bt.xi[(size_t)b * n + lane] : 0.0;
      }
    } else {
      double lt = -1.0, lb_ = -1.0, ut = 1.0, ub_ = 1.0, et = 1.0, eb = 1.0;
      if (lane < mp) {
        const size_t rt = (size_t)b * m + lane, rb = rt + mp;
        et = I.E[rt]; eb = I.E[rb];
        lt = et * (Bt.l ? Bt.l[rt] : I.l0[lane]); lb_ = eb * (Bt.l ? Bt.l[rb] : I.l0[lane + mp]);
        ut = et * (Bt.u ? Bt.u[rt] : I.u0[lane]); ub_ = eb * (Bt.u ? Bt.u[rb] : I.u0[lane + mp]);
        if (warm) { ZT_ = Bt.z[rt]; ZB_ = Bt.z[rb]; YT_ = Bt.y[rt]; YB_ = Bt.y[rb]; }
        flag_bad = (lt > ut) || (lb_ > ub_);
        cls_t = row_class(lt, ut); cls_b = row_class(lb_, ub_);
        flag_diff = (cls_t != row_class(et * I.l0[lane], et * I.u0[lane])) || (cls_b != row_class(eb * I.l0[lane + mp], eb * I.u0[lane + mp]));
      }
      lbt[lane] = lt; lbb[lane] = lb_; ubt[lane] = ut; ubb[lane] = ub_; evt[lane] = et; evb[lane] = eb;
      flag_bad = __any_sync(kFull, flag_bad); flag_diff = __any_sync(kFull, flag_diff);
      if (lane == 0) { red[R_FLAG0] = flag_bad ? 1.0 : 0.0; red[R_FLAG1] = flag_diff ? 1.0 : 0.0; }
    }
    double rho = (Bt.fresh || prepare) ? fmin(fmax(S.rho0, kRhoMin), kRhoMax) : Bt.rho[b];
    if (prepare) {   // build G' (padded) from the scaled A̅ of this instance; everything else of the pack follows below
      const double *gA = I.A + (size_t)b * m * n;
      for (int e = tid; e < n * kPLD; e += 64) Gt[e] = 0.0;
      __syncthreads();
      for (int e = tid; e < mp * n; e += 64) { const int r = e / n, i = e % n; Gt[i * kPLD + r] = gA[e]; }
      __syncthreads();
      for (int e = tid; e < n * kPLD; e += 64) pack[e] = Gt[e];
    } else {
      mbar_wait(&mbar[stage], stage ? phase1 : phase0);   // G' of this instance has landed (issued one solve ago)
      if (stage) phase1 ^= 1u; else phase0 ^= 1u;
    }
    __syncthreads();
    const bool bad_bounds = red[R_FLAG0] != 0.0;
    bool split_ok = red[R_FLAG1] == 0.0;     // the pack's S0 / T / M^-1 / K hold for the row classes of the SETUP bounds only
    static const bool eig_on = true;
    const bool eig_ok = eig_on && !prepare && pack[oFlag] == 1.0;   // the pack holds the pencil's eigen-decomposition

    // ---- operators into registers
    double gt[16], op[kPN];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const double2 v = lane < n ? *reinterpret_cast<const double2 *>(Gt + lane * kPLD + 16 * w + j) : make_double2(0.0, 0.0);
      gt[j] = v.x; gt[j + 1] = v.y;
    }
#pragma unroll
    for (int k = 0; k < kPN; ++k) op[k] = 0.0;

    // OSQP set_rho_vec / update_rho_vec for this solve's row classes (z-warp lanes own their rows' entries)
    auto set_rho = [&]() {
      if (!xw) {
        const double rho_eq = kRhoEqOverIneq * rho;
        const double vt = cls_t < 0 ? kRhoMin : (cls_t == 1 ? rho_eq : rho), vb = cls_b < 0 ? kRhoMin : (cls_b == 1 ? rho_eq : rho);
        rvt[lane] = vt; rvb[lane] = vb; rit[lane] = fast_rcp(vt); rib[lane] = fast_rcp(vb);
      }
    };
    set_rho();

    // (A̅' v)_lane for v = top - bottom given in t0[] (valid on the x-warp's lanes < n).  Two barriers.
    auto AT_dot_t0 = [&]() -> double {
      const double part = dot_reg16(gt, t0 + 16 * w);
      if (!xw) p1[lane] = part;
      __syncthreads();
      const double full = part + p1[lane];
      __syncthreads();
      return full;
    };

    // Factorisation and inverse of M = P̄ + sigma I + A̅' diag(rho_vec) A̅ by pivot-free Gauss-Jordan sweeps IN REGISTERS: thread (w, i)
    // holds M(i, 16 w .. 16 w + 15).  Sweep k uses the k-th LDL' pivot d_k (the pivots QDLDL / a Cholesky factorisation produce;
    // d_k > 0 for all k <=> M positive definite), broadcasts the pivot row and column through shared memory (double buffered:
    // one CTA barrier per sweep) and updates 16 entries per thread; after n sweeps the registers hold M^-1.  All 64 threads
    // work in every sweep and the dependent chain per sweep is one reciprocal + one FMA.  Then rows of M^-1 -> x-warp registers,
    // rows of K = G M^-1 -> z-warp registers.
    // mode 0: M = S0 + rho T from the pack; 1: assembled from G and this solve's rho_vec; 2 (prepare): also writes S0, T;
    // mode 3: no sweeps at all -- M(rho)^-1 = V diag(1 / (1 + rho lambda)) V' from the pack's eigen-decomposition (prepared at
    // create time by eigen_prepare below): 16 entries per thread, 30 independent FMAs each, no barrier inside.
    auto refactor = [&](int mode) -> bool {
      double *row = Sc + lane * kPLD + 16 * w;      // this thread's 16 entries of M (rows >= n are never touched)
      if (mode == 3) {
        const double *VT = pack + oVT, *lam = pack + oLam;
        // every 128-byte line of V' into L1 at once (64 lines at n = 30, one per thread): the k loop below would otherwise take
        // the L2 latency once per k (its loads of row k + 1 are not issued before the FMAs of row k)
        for (int e = tid * 16; e < n * kPLD; e += 64 * 16) asm volatile("prefetch.global.L1 [%0];" ::"l"(VT + e));
        if (tid < kPN) t1[tid] = tid < n ? fast_rcp(1.0 + rho * lam[tid]) : 0.0;
        __syncthreads();
        if (lane < n) {
          double acc[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) acc[jj] = 0.0;
          for (int k = 0; k < n; ++k) {
            const double vik = VT[k * kPLD + lane] * t1[k];
            const double *vr = VT + k * kPLD + 16 * w;
#pragma unroll
            for (int jj = 0; jj < 16; jj += 2) {
              const double2 v2 = *reinterpret_cast<const double2 *>(vr + jj);
              acc[jj] = fma(vik, v2.x, acc[jj]); acc[jj + 1] = fma(vik, v2.y, acc[jj + 1]);
            }
          }
#pragma unroll
          for (int jj = 0; jj < 16; jj += 2) *reinterpret_cast<double2 *>(row + jj) = make_double2(acc[jj], acc[jj + 1]);
        }
        __syncthreads();
        PFI(5)
      } else {
      if (mode == 0) {
        const double *S0 = pack + 3 * kPN * kPLD, *T = S0 + tri2;
        if (lane < n) {
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const int j = 16 * w + jj;
            double v = 0.0;
            if (j < n) {
              const int hi_ = lane > j ? lane : j, lo_ = lane > j ? j : lane, e = hi_ * (hi_ + 1) / 2 + lo_;
              v = fma(rho, T[e], S0[e]);
            }
            row[jj] = v;
          }
        }
      } else {
        // per-row weights of G_r' G_r: t0 = kappa_top + kappa_bot (times rho), t1 = the free rows' rho_min
        if (!xw) {
          if (mode == 2) {
            t0[lane] = lane < mp ? ((cls_t < 0 ? 0.0 : (cls_t == 1 ? kRhoEqOverIneq : 1.0)) + (cls_b < 0 ? 0.0 : (cls_b == 1 ? kRhoEqOverIneq : 1.0))) : 0.0;
            t1[lane] = lane < mp ? ((cls_t < 0 ? kRhoMin : 0.0) + (cls_b < 0 ? kRhoMin : 0.0)) : 0.0;
          } else {
            t0[lane] = lane < mp ? rvt[lane] + rvb[lane] : 0.0;
          }
        }
        __syncthreads();
        for (int e = tid; e < n * kPN; e += 64) {      // full symmetric M, columns >= n zero
          const int i = e >> 5, j = e & 31;
          double mij = 0.0;
          if (j < n) {
            const double *gi = Gt + i * kPLD, *gj = Gt + j * kPLD;
            double s0 = gP[i * n + j] + (i == j ? sigma : 0.0), tt = 0.0;
            if (mode == 2) {
              for (int r = 0; r < mp; ++r) { const double aa = gi[r] * gj[r]; s0 = fma(t1[r], aa, s0); tt = fma(t0[r], aa, tt); }
              if (j <= i) { pack[3 * kPN * kPLD + i * (i + 1) / 2 + j] = s0; pack[3 * kPN * kPLD + tri2 + i * (i + 1) / 2 + j] = tt; }
              mij = fma(rho, tt, s0);
            } else {
              for (int r = 0; r < mp; ++r) s0 = fma(t0[r] * gi[r], gj[r], s0);
              mij = s0;
            }
          }
          Sc[i * kPLD + j] = mij;
        }
      }
      __syncthreads();
      PFI(5)
      // Gauss-Jordan sweeps in place (shared memory); the scaled pivot row goes through a double-buffered broadcast row
      int ok = 1;
      const bool live = lane < n;
      for (int k = 0; k < n; ++k) {
        double *prow = ((k & 1) ? xs : rhs) + 16 * w;                      // this warp's half of the broadcast row
        const double d = Sc[k * kPLD + k];
        if (!(d > 0.0)) ok = 0;                                             // LDL' pivot: M is not positive definite
        const double pv = fast_rcp(d);
        // Nothing of M is written before the barrier: every thread still reads the pivot d and its column-k entry f.
        const double f = live ? Sc[lane * kPLD + k] : 0.0;
        if (lane == k) {                                                    // pivot row, scaled, into the broadcast buffer
#pragma unroll
          for (int jj = 0; jj < 16; jj += 2) {
            double2 r2 = *reinterpret_cast<const double2 *>(row + jj);
            r2.x *= pv; r2.y *= pv;
            *reinterpret_cast<double2 *>(prow + jj) = r2;
          }
        }
        __syncthreads();
        if (live) {                                                         // row_i <- row_i - f * (row_k / d); row_k <- row_k / d
          // the pivot row takes the same FMA with g = 1 on a row of zeros (1 * x + 0 = the scaled row, exactly): no per-entry select
          const bool piv = lane == k;
          const double g = piv ? 1.0 : -f;
          const double *src = piv ? zer : row;
#pragma unroll
          for (int jj = 0; jj < 16; jj += 2) {
            const double2 p2 = *reinterpret_cast<const double2 *>(prow + jj);
            double2 r2 = *reinterpret_cast<const double2 *>(src + jj);
            r2.x = fma(g, p2.x, r2.x);
            r2.y = fma(g, p2.y, r2.y);
            *reinterpret_cast<double2 *>(row + jj) = r2;
          }
          if (w == (k >> 4)) row[k & 15] = piv ? pv : -f * pv;              // column k: 1 / d on the pivot row, -f / d elsewhere
        }
        __syncthreads();
      }
      PFI(6)
      if (!ok) return false;                                               // (uniform: every thread saw the same pivots)
      }
      PFI(7)
      if (xw) {
#pragma unroll
        for (int k = 0; k < kPN; k += 2) {
          const double2 v = lane < n ? *reinterpret_cast<const double2 *>(Sc + lane * kPLD + k) : make_double2(0.0, 0.0);
          op[k] = v.x; op[k + 1] = v.y;
        }
      } else {   // K_r = sum_k G(r, k) M^-1(k, :)
#pragma unroll
        for (int k = 0; k < kPN; ++k) op[k] = 0.0;
        for (int k = 0; k < n; ++k) {
          const double g = lane < mp ? Gt[k * kPLD + lane] : 0.0;
          const double *row = Sc + k * kPLD;
#pragma unroll
          for (int c2 = 0; c2 < kPN; c2 += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(row + c2);
            op[c2] = fma(g, v.x, op[c2]); op[c2 + 1] = fma(g, v.y, op[c2 + 1]);
          }
        }
      }
      __syncthreads();
      PFI(9)
      return true;
    };

    // Create-time pass (warp 0, after S0 and T are in the pack): generalised eigen-decomposition of the pencil (T, S0) --
    // Cholesky S0 = L L', C = L^-1 T L^-T, cyclic Jacobi C = Q diag(lambda) Q', V = L^-T Q -- so that rho updates need no
    // refactorisation sweeps (refactor mode 3).  Work buffers: Sc (L), the idle G' stage (C), this instance's G' stage (Q: G' is
    // already in the pack and K has been formed).  Returns false when S0 is not positive definite or Jacobi does not settle.
    auto eigen_prepare = [&]() -> bool {
      double *Am = Sc, *Bm = stage ? Gt0 : Gt1, *Qm = Gt;
      const double *S0 = pack + 3 * kPN * kPLD, *T = S0 + tri2;
      for (int e = lane; e < n * n; e += 32) {
        const int i = e / n, j = e % n, hi_ = i > j ? i : j, lo_ = i > j ? j : i, t = hi_ * (hi_ + 1) / 2 + lo_;
        Am[i * kPLD + j] = S0[t]; Bm[i * kPLD + j] = T[t]; Qm[i * kPLD + j] = i == j ? 1.0 : 0.0;
      }
      __syncwarp();
      bool pd = true;
      for (int k = 0; k < n; ++k) {                       // Cholesky, lower triangle in place
        const double akk2 = Am[k * kPLD + k];
        if (!(akk2 > 0.0)) pd = false;
        const double akk = sqrt(akk2 > 0.0 ? akk2 : 1.0);
        __syncwarp();
        if (lane == k) Am[k * kPLD + k] = akk;
        else if (lane > k && lane < n) Am[lane * kPLD + k] /= akk;
        __syncwarp();
        if (lane > k && lane < n) {
          const double lik = Am[lane * kPLD + k];
          for (int j = k + 1; j <= lane; ++j) Am[lane * kPLD + j] -= lik * Am[j * kPLD + k];
        }
        __syncwarp();
      }
      if (!pd) return false;
      if (lane < n) {                                      // Y = L^-1 T (lane = column), in place
        for (int i = 0; i < n; ++i) {
          double sacc = Bm[i * kPLD + lane];
          for (int k = 0; k < i; ++k) sacc -= Am[i * kPLD + k] * Bm[k * kPLD + lane];
          Bm[i * kPLD + lane] = sacc / Am[i * kPLD + i];
        }
      }
      __syncwarp();
      if (lane < n) {                                      // C = Y L^-T (lane = row), in place
        for (int j = 0; j < n; ++j) {
          double sacc = Bm[lane * kPLD + j];
          for (int k = 0; k < j; ++k) sacc -= Am[j * kPLD + k] * Bm[lane * kPLD + k];
          Bm[lane * kPLD + j] = sacc / Am[j * kPLD + j];
        }
      }
      __syncwarp();
      for (int e = lane; e < n * n; e += 32) {             // symmetrise
        const int i = e / n, j = e % n;
        if (i < j) { const double v = 0.5 * (Bm[i * kPLD + j] + Bm[j * kPLD + i]); Bm[i * kPLD + j] = v; Bm[j * kPLD + i] = v; }
      }
      __syncwarp();
      bool settled = false;
      for (int sweep = 0; sweep < 30 && !settled; ++sweep) {
        double off = 0.0, dg = 0.0;
        if (lane < n)
          for (int j = 0; j < n; ++j) { const double v = Bm[lane * kPLD + j]; if (j == lane) dg += v * v; else off += v * v; }
        off = wsum(off); dg = wsum(dg);
        if (off <= 1e-30 * dg || off == 0.0) { settled = true; break; }
        for (int p_ = 0; p_ < n - 1; ++p_)
          for (int q_ = p_ + 1; q_ < n; ++q_) {
            const double apq = Bm[p_ * kPLD + q_];
            if (apq == 0.0) continue;                      // (uniform: every lane reads the same entry)
            const double app = Bm[p_ * kPLD + p_], aqq = Bm[q_ * kPLD + q_];
            const double theta = (aqq - app) / (2.0 * apq);
            const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double cs = 1.0 / sqrt(tt * tt + 1.0), sn = tt * cs;
            __syncwarp();
            if (lane < n) {
              const double ckp = Bm[lane * kPLD + p_], ckq = Bm[lane * kPLD + q_];
              const double qkp = Qm[lane * kPLD + p_], qkq = Qm[lane * kPLD + q_];
              Qm[lane * kPLD + p_] = cs * qkp - sn * qkq; Qm[lane * kPLD + q_] = sn * qkp + cs * qkq;
              if (lane != p_ && lane != q_) {
                const double np_ = cs * ckp - sn * ckq, nq_ = sn * ckp + cs * ckq;
                Bm[lane * kPLD + p_] = np_; Bm[lane * kPLD + q_] = nq_; Bm[p_ * kPLD + lane] = np_; Bm[q_ * kPLD + lane] = nq_;
              }
            }
            if (lane == 0) { Bm[p_ * kPLD + p_] = app - tt * apq; Bm[q_ * kPLD + q_] = aqq + tt * apq; Bm[p_ * kPLD + q_] = 0.0; Bm[q_ * kPLD + p_] = 0.0; }
            __syncwarp();
          }
      }
      if (!settled) return false;
      if (lane < n) {                                      // V = L^-T Q (lane = column), in place
        for (int i = n - 1; i >= 0; --i) {
          double sacc = Qm[i * kPLD + lane];
          for (int k = i + 1; k < n; ++k) sacc -= Am[k * kPLD + i] * Qm[k * kPLD + lane];
          Qm[i * kPLD + lane] = sacc / Am[i * kPLD + i];
        }
      }
      __syncwarp();
      double *VT = pack + oVT, *lam = pack + oLam;
      for (int e = lane; e < kPN * kPLD; e += 32) {
        const int k = e / kPLD, i = e % kPLD;
        VT[e] = (k < n && i < n) ? Qm[i * kPLD + k] : 0.0;
      }
      lam[lane] = lane < n ? fmax(Bm[lane * kPLD + lane], 0.0) : 0.0;
      return true;
    };

    int status = SMPC_UNSOLVED, iter = 0, rho_updates = 0;
    bool can_check = false, factor_ok = true, refactored = false;
    double F_pri = 0.0, F_dua = 0.0, F_obj = 0.0;

    // OSQP update_info: residuals, their norms, the objective -- and the first, cheap tests of both infeasibility checks
    auto update_info = [&]() {
      if (xw) xs[lane] = X_; else t0[lane] = YT_ - YB_;
      __syncthreads();
      // P̄ x: P̄ is stored full symmetric, so the column sweep is coalesced; each warp takes half of the sum (16 independent loads)
      double Px = 0.0, Ax = 0.0;
      {
        double pv[16];
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) { const int k = 16 * w + kk; pv[kk] = (lane < n && k < n) ? gP[k * n + lane] : 0.0; }
        double p0 = 0.0, p1_ = 0.0;
#pragma unroll
        for (int kk = 0; kk < 16; kk += 2) { p0 = fma(pv[kk], xs[16 * w + kk], p0); p1_ = fma(pv[kk + 1], xs[16 * w + kk + 1], p1_); }
        Px = p0 + p1_;
      }
      if (!xw && lane < mp) {
        double a0 = 0.0, a1 = 0.0;
        int k = 0;
        for (; k + 2 <= n; k += 2) { a0 = fma(Gt[k * kPLD + lane], xs[k], a0); a1 = fma(Gt[(k + 1) * kPLD + lane], xs[k + 1], a1); }
        if (k < n) a0 = fma(Gt[k * kPLD + lane], xs[k], a0);
        Ax = a0 + a1;
      }
      double Aty = dot_reg16(gt, t0 + 16 * w);
      if (!xw) { p1[lane] = Aty; t1[lane] = Px; }
      __syncthreads();
      if (xw) { Aty += p1[lane]; Px += t1[lane]; }
      if (xw) {
        double rd = 0.0, ob = 0.0, u_rd = 0.0, u_q = 0.0, u_Aty = 0.0, u_Px = 0.0, nd = 0.0, qd = 0.0;
        double a_q = 0.0, a_Aty = 0.0, a_Px = 0.0;
        if (lane < n) {
          const double x = X_, q = Q_, Dinv = fast_rcp(DV_), dx = dl2[lane];
          rd = fabs((q + Px) + Aty); a_q = fabs(q); a_Aty = fabs(Aty); a_Px = fabs(Px);
          u_rd = fabs(Dinv * ((q + Px) + Aty)); u_q = fabs(Dinv * q); u_Aty = fabs(Dinv * Aty); u_Px = fabs(Dinv * Px);
          ob = 0.5 * x * Px + q * x;
          nd = fabs(unscale ? DV_ * dx : dx); qd = q * dx;
        }
        rd = wmax(rd); a_q = wmax(a_q); a_Aty = wmax(a_Aty); a_Px = wmax(a_Px);
        u_rd = wmax(u_rd); u_q = wmax(u_q); u_Aty = wmax(u_Aty); u_Px = wmax(u_Px);
        ob = wsum(ob); nd = wmax(nd); qd = wsum(qd);
        if (lane == 0) {
          red[R_SRD] = rd; red[R_SQ] = a_q; red[R_SATY] = a_Aty; red[R_SPX] = a_Px;
          red[R_URD] = u_rd; red[R_UQ] = u_q; red[R_UATY] = u_Aty; red[R_UPX] = u_Px; red[R_OB] = ob; red[R_DIND] = nd; red[R_DIQD] = qd;
        }
      } else {
        double a_rp = 0.0, a_z = 0.0, a_Ax = 0.0, u_rp = 0.0, u_z = 0.0, u_Ax = 0.0, nd = 0.0, lhs = 0.0;
        if (lane < mp) {
          const double et = evt[lane], eb = evb[lane], eit = fast_rcp(et), eib = fast_rcp(eb);
          const double zt = ZT_, zb = ZB_;
          const double rpt = Ax - zt, rpb = -Ax - zb;
          a_rp = fmax(fabs(rpt), fabs(rpb)); a_z = fmax(fabs(zt), fabs(zb)); a_Ax = fabs(Ax);
          u_rp = fmax(fabs(eit * rpt), fabs(eib * rpb)); u_z = fmax(fabs(eit * zt), fabs(eib * zb)); u_Ax = fmax(fabs(eit * Ax), fabs(eib * Ax));
          // is_primal_infeasible, first part: project delta_y on the recession directions, its norm and u' dy+ + l' dy-
          const double lt = lbt[lane], lb_ = lbb[lane], ut = ubt[lane], ub_ = ubb[lane];
          double dt = dl0[lane], db = dl1[lane];
          const bool uti = ut > kInfty * kMinScaling, lti = lt < -kInfty * kMinScaling, ubi = ub_ > kInfty * kMinScaling, lbi = lb_ < -kInfty * kMinScaling;
          if (uti) dt = lti ? 0.0 : fmin(dt, 0.0); else if (lti) dt = fmax(dt, 0.0);
          if (ubi) db = lbi ? 0.0 : fmin(db, 0.0); else if (lbi) db = fmax(db, 0.0);
          dl0[lane] = dt; dl1[lane] = db;
          nd = fmax(fabs(unscale ? et * dt : dt), fabs(unscale ? eb * db : db));
          const double dpt = fmax(dt, 0.0), dmt = fmin(dt, 0.0), dpb = fmax(db, 0.0), dmb = fmin(db, 0.0);
          if (dpt != 0.0) lhs += ut * dpt;
          if (dmt != 0.0) lhs += lt * dmt;
          if (dpb != 0.0) lhs += ub_ * dpb;
          if (dmb != 0.0) lhs += lb_ * dmb;
        }
        a_rp = wmax(a_rp); a_z = wmax(a_z); a_Ax = wmax(a_Ax); u_rp = wmax(u_rp); u_z = wmax(u_z); u_Ax = wmax(u_Ax);
        nd = wmax(nd); lhs = wsum(lhs);
        if (lane == 0) {
          red[R_SRP] = a_rp; red[R_SZ] = a_z; red[R_SAX] = a_Ax; red[R_URP] = u_rp; red[R_UZ] = u_z; red[R_UAX] = u_Ax;
          red[R_PIND] = nd; red[R_PILHS] = lhs;
        }
      }
      __syncthreads();
      if (unscale) { F_pri = red[R_URP]; F_dua = cinv * red[R_URD]; F_obj = cinv * red[R_OB]; }
      else { F_pri = red[R_SRP]; F_dua = red[R_SRD]; F_obj = red[R_OB]; }
    };

    auto primal_infeasible = [&](double eps) -> bool {   // uniform over the CTA
      const double nd = red[R_PIND], lhs = red[R_PILHS];
      if (!(nd > eps)) return false;
      if (!(lhs < -eps * nd)) return false;
      if (!xw) t0[lane] = dl0[lane] - dl1[lane];
      __syncthreads();
      const double v = AT_dot_t0();
      if (xw) {
        const double na = wmax(lane < n ? fabs(unscale ? fast_rcp(DV_) * v : v) : 0.0);
        if (lane == 0) red[R_RARE0] = na;
      }
      __syncthreads();
      return red[R_RARE0] < eps * nd;
    };

    auto dual_infeasible = [&](double eps) -> bool {   // uniform over the CTA
      const double nd = red[R_DIND], qd = red[R_DIQD];
      const double cs = unscale ? c : 1.0;
      if (!(nd > eps)) return false;
      if (!(qd < -cs * eps * nd)) return false;
      if (xw) {
        double s = 0.0;
        if (lane < n) for (int k = 0; k < n; ++k) s = fma(gP[k * n + lane], dl2[k], s);
        const double np = wmax(lane < n ? fabs(unscale ? fast_rcp(DV_) * s : s) : 0.0);
        if (lane == 0) red[R_RARE0] = np;
      } else {
        int bad = 0;
        if (lane < mp) {
          double v = 0.0;
          for (int k = 0; k < n; ++k) v = fma(Gt[k * kPLD + lane], dl2[k], v);
          const double vt = unscale ? v * fast_rcp(evt[lane]) : v, vb = unscale ? -v * fast_rcp(evb[lane]) : -v;
          if (((ubt[lane] < kInfty * kMinScaling) && (vt > eps * nd)) || ((lbt[lane] > -kInfty * kMinScaling) && (vt < -eps * nd))) bad = 1;
          if (((ubb[lane] < kInfty * kMinScaling) && (vb > eps * nd)) || ((lbb[lane] > -kInfty * kMinScaling) && (vb < -eps * nd))) bad = 1;
        }
        bad = __any_sync(kFull, bad);
        if (lane == 0) red[R_RARE1] = bad ? 1.0 : 0.0;
      }
      __syncthreads();
      return (red[R_RARE0] < cs * eps * nd) && red[R_RARE1] == 0.0;
    };

    auto check_termination = [&](bool approx) -> bool {   // uniform over the CTA
      double ea = S.eps_abs, er = S.eps_rel, epi = S.eps_prim_inf, edi = S.eps_dual_inf;
      if (approx) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
      const double nEz = unscale ? red[R_UZ] : red[R_SZ], nEAx = unscale ? red[R_UAX] : red[R_SAX];
      const double nDq = unscale ? red[R_UQ] : red[R_SQ], nDAty = unscale ? red[R_UATY] : red[R_SATY], nDPx = unscale ? red[R_UPX] : red[R_SPX];
      bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
      if (F_pri < ea + er * fmax(nEz, nEAx)) prim_ok = true;
      else prim_inf = primal_infeasible(epi);
      if (F_dua < ea + er * (unscale ? cinv : 1.0) * fmax(fmax(nDq, nDAty), nDPx)) dual_ok = true;
      else dual_inf = dual_infeasible(edi);
      if (prim_ok && dual_ok) { status = approx ? SMPC_SOLVED_INACCURATE : SMPC_SOLVED; return true; }
      if (prim_inf) { status = approx ? SMPC_PRIMAL_INFEASIBLE_INACCURATE : SMPC_PRIMAL_INFEASIBLE; F_obj = kInfty; return true; }
      if (dual_inf) { status = approx ? SMPC_DUAL_INFEASIBLE_INACCURATE : SMPC_DUAL_INFEASIBLE; F_obj = -kInfty; return true; }
      return false;
    };

    // ---- factorisation for this solve's rho: the pack's (osqp_setup / the last refactorisation that was stored), or a new one
    // at the top of the first iteration (one call site for setup, rho updates and the prepare pass)
    bool need_factor = !bad_bounds;
    if (!prepare && !bad_bounds && split_ok && rho == I.pack_rho[b]) {
      need_factor = false;
      if (lane < (xw ? n : mp)) {
        const double *src = pack + (xw ? 1 : 2) * kPN * kPLD + lane * kPLD;
#pragma unroll
        for (int k = 0; k < kPN; k += 2) { const double2 v = *reinterpret_cast<const double2 *>(src + k); op[k] = v.x; op[k + 1] = v.y; }
      }
    }
    // wd for the first iteration
    if (!xw) wd[lane] = lane < mp ? (rvt[lane] * ZT_ - YT_) - (rvb[lane] * ZB_ - YB_) : 0.0;
    // Iterations run in stretches that end at the next event (termination check, rho adaptation, max_iter): the hot loop is a
    // plain counted loop with nothing but the iteration in it, so the operator rows stay in registers across it.
    const int kNever = 0x3fffffff;
    int to_check = S.check_every ? S.check_every : kNever, to_adapt = (S.adaptive_rho && S.rho_interval) ? S.rho_interval : kNever;
    iter = 0;
    PFI(0)
    while (!bad_bounds) {
      if (need_factor) {
        PFI(2)
        factor_ok = refactor(prepare ? 2 : (split_ok ? (eig_ok ? 3 : 0) : 1));
#ifdef SMPC_PAIR_PROFILE
        ++pf_refac;
#endif
        need_factor = false; refactored = true;
        if (prepare || !factor_ok) break;
      }
      const int run = min(min(to_check, to_adapt), S.max_iter - iter);
      for (int k = run; k > 0; --k) {
        __syncthreads();                                                                  // wd is visible
        const double part = dot_reg16(gt, wd + 16 * w);                                   // half of (G' wd)_lane
        if (!xw) p1[lane] = part;
        __syncthreads();
        if (xw) rhs[lane] = lane < n ? (sigma * X_ - Q_) + (part + p1[lane]) : 0.0;
        __syncthreads();
        const double acc = dot_reg32(op, rhs);                                            // x~_lane (x-warp) | (G x~)_lane (z-warp)
        const bool keep_deltas = k == 1;                                                  // delta_x, delta_y feed the infeasibility tests only
        if (xw) {
          const double xn = alpha * acc + (1.0 - alpha) * X_;
          if (keep_deltas) dl2[lane] = xn - X_;
          X_ = xn;
        } else if (lane < mp) {
          {
            const double zr = alpha * acc + (1.0 - alpha) * ZT_;
            const double zn = clamp_sel(zr + rit[lane] * YT_, lbt[lane], ubt[lane]);
            const double d = rvt[lane] * (zr - zn);
            ZT_ = zn; YT_ += d;
            if (keep_deltas) dl0[lane] = d;
          }
          {
            const double zr = alpha * (-acc) + (1.0 - alpha) * ZB_;
            const double zn = clamp_sel(zr + rib[lane] * YB_, lbb[lane], ubb[lane]);
            const double d = rvb[lane] * (zr - zn);
            ZB_ = zn; YB_ += d;
            if (keep_deltas) dl1[lane] = d;
          }
          wd[lane] = (rvt[lane] * ZT_ - YT_) - (rvb[lane] * ZB_ - YB_);
        }
      }
      PFI(2)
      iter += run; to_check -= run; to_adapt -= run;
      can_check = to_check == 0;
      const bool adapt = to_adapt == 0, last = iter >= S.max_iter;
      if (can_check) to_check = S.check_every;
      if (adapt) to_adapt = S.rho_interval;
      bool done = false;
      if (can_check || adapt || last) {
        update_info();
        // one call site for the termination test: pass 0 = osqp_solve's check inside the loop (and, at max_iter, the one behind
        // it), then rho adaptation, pass 1 (max_iter only) = the approximate test with 10 x tolerances
        for (int pass = 0; pass < 2 && !done; ++pass) {
          if (pass == 0 ? (can_check || last) : last) done = check_termination(pass == 1);
          if (pass == 0 && !done && adapt) {
            const double pr = fast_div(red[R_SRP], fmax(red[R_SZ], red[R_SAX]) + kDivTol);
            const double dr = fast_div(red[R_SRD], fmax(fmax(red[R_SQ], red[R_SATY]), red[R_SPX]) + kDivTol);
            const double ratio = fast_div(pr, dr + kDivTol);
            const double rn = fmin(fmax(rho * (ratio > 0.0 ? fast_sqrt(ratio) : 0.0), kRhoMin), kRhoMax);
            if (rn > rho * S.rho_tol || rn * S.rho_tol < rho) {
              rho = rn; ++rho_updates;
              set_rho();
              need_factor = !last;
              if (!xw && lane < mp) wd[lane] = (rvt[lane] * ZT_ - YT_) - (rvb[lane] * ZB_ - YB_);   // rho_vec changed
            }
          }
        }
        if (last && !done) { status = SMPC_MAX_ITER_REACHED; done = true; }
      }
      PFI(3)
      if (done) break;
    }
    if (prepare) {   // create-time pass: the pack now holds G', S0, T; add M(rho0)^-1 and K and leave
      const bool ok = !bad_bounds && factor_ok;
      if (ok && lane < (xw ? n : mp)) {
        double *dst = pack + (xw ? 1 : 2) * kPN * kPLD + lane * kPLD;
#pragma unroll
        for (int k = 0; k < kPN; k += 2) *reinterpret_cast<double2 *>(dst + k) = make_double2(op[k], op[k + 1]);
      }
      if (tid == 0) I.pack_rho[b] = ok ? rho : -1.0;   // -1: no usable factorisation in the pack
      __syncthreads();
      bool eig = false;
      if (ok && xw) eig = eigen_prepare();
      if (tid == 0) { pack[oFlag] = eig ? 1.0 : 0.0; pack[oFlag + 1] = 0.0; red[R_FLAG0] = eig ? 1.0 : 0.0; }
      __syncthreads();
      if (red[R_FLAG0] != 0.0) {
        // the pack's M(rho0)^-1 and K in the eigen form too, so that a later cold solve (refactor mode 3 at rho0) reproduces the
        // first one bit for bit; G' comes back from the pack (its stage served as a work buffer)
        for (int e = tid; e < n * kPLD; e += 64) Gt[e] = pack[e];
        __syncthreads();
        refactor(3);
        if (lane < (xw ? n : mp)) {
          double *dst = pack + (xw ? 1 : 2) * kPN * kPLD + lane * kPLD;
#pragma unroll
          for (int k = 0; k < kPN; k += 2) *reinterpret_cast<double2 *>(dst + k) = make_double2(op[k], op[k + 1]);
        }
        __syncthreads();
      }
      cur = s_ticket;
      continue;
    }
    if (bad_bounds || !factor_ok) iter = 0;

    const bool has_sol = !bad_bounds && factor_ok && !(status == SMPC_PRIMAL_INFEASIBLE || status == SMPC_PRIMAL_INFEASIBLE_INACCURATE ||
                                                       status == SMPC_DUAL_INFEASIBLE || status == SMPC_DUAL_INFEASIBLE_INACCURATE);
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    if (xw) {
      if (lane < n) {
        if (Bt.x_out) Bt.x_out[(size_t)b * n + lane] = has_sol ? DV_ * X_ : qnan;
        Bt.xi[(size_t)b * n + lane] = has_sol ? X_ : 0.0;
      }
    } else if (lane < mp) {
      const size_t rt = (size_t)b * m + lane, rb = rt + mp;
      if (Bt.y_out) { Bt.y_out[rt] = has_sol ? cinv * (evt[lane] * YT_) : qnan; Bt.y_out[rb] = has_sol ? cinv * (evb[lane] * YB_) : qnan; }
      Bt.z[rt] = has_sol ? ZT_ : 0.0; Bt.z[rb] = has_sol ? ZB_ : 0.0;
      Bt.y[rt] = has_sol ? YT_ : 0.0; Bt.y[rb] = has_sol ? YB_ : 0.0;
    }
    // OSQP keeps its factorisation between solves: store the one for the final rho (warm solves restart from it)
    if (refactored && factor_ok && split_ok && !Bt.fresh) {
      if (lane < (xw ? n : mp)) {
        double *dst = pack + (xw ? 1 : 2) * kPN * kPLD + lane * kPLD;
#pragma unroll
        for (int k = 0; k < kPN; k += 2) *reinterpret_cast<double2 *>(dst + k) = make_double2(op[k], op[k + 1]);
      }
      if (tid == 0) I.pack_rho[b] = rho;
    }
    if (tid == 0) {
      Bt.rho[b] = rho;
      Bt.status[b] = status; Bt.iter[b] = iter; Bt.rho_updates[b] = rho_updates;
      Bt.obj[b] = F_obj; Bt.pri_res[b] = F_pri; Bt.dua_res[b] = F_dua;
    }
    __syncthreads();
    cur = s_ticket;
    stage ^= 1;
    PFI(4)
#ifdef SMPC_PAIR_PROFILE
    ++pf_solves; pf_iters += iter;
#endif
  }
#ifdef SMPC_PAIR_PROFILE
  if (blockIdx.x == 0 && tid == 0 && !prepare)
    printf("pair profile (cycles, CTA 0): prologue %lld, refactor (%d): assembly %lld cholesky %lld Linv %lld Minv %lld K %lld, iterations %lld (%d), checks %lld, epilogue %lld; solves %d\n",
           pf[0], pf_refac, pf[5], pf[6], pf[7], pf[8], pf[9], pf[2], pf_iters, pf[3], pf[4], pf_solves);
#endif
  // the last CTA to leave re-arms the ticket counter for the next launch
  if (tid == 0) {
    __threadfence();
    const int done = atomicAdd(queue + 1, 1);
    if (done == (int)gridDim.x - 1) { queue[0] = 0; queue[1] = 0; __threadfence(); }
  }
}

template <int CTAS_PER_SM>
static cudaError_t launch_pair(const InstanceDataDev &I, const BatchDev &Bt, const SettingsDev &S, int num_sms, cudaStream_t stream, int prepare) {
  const size_t smem = pair_smem_bytes(I.n);
  cudaError_t e = cudaFuncSetAttribute(admm_instance_pair_kernel<CTAS_PER_SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device: every launch
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, admm_instance_pair_kernel<CTAS_PER_SM>, 64, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidValue;
  int grid = num_sms * per_sm;
  if (grid > Bt.B) grid = Bt.B;
  admm_instance_pair_kernel<CTAS_PER_SM><<<grid, 64, smem, stream>>>(I, Bt, S, I.queue, prepare);
  return cudaGetLastError();
}

cudaError_t launch_admm_instance_pair(const InstanceDataDev &I, const BatchDev &Bt, const SettingsDev &S, int num_sms, cudaStream_t stream,
                                      int prepare) {
  static const int ctas = getenv("SMPC_PAIR_CTAS") ? atoi(getenv("SMPC_PAIR_CTAS")) : 5;   // register budget per thread: 6 -> 168, 7 -> 144, 5 -> 200
  if (ctas == 7) return launch_pair<7>(I, Bt, S, num_sms, stream, prepare);
  if (ctas == 5) return launch_pair<5>(I, Bt, S, num_sms, stream, prepare);
  if (ctas == 4) return launch_pair<4>(I, Bt, S, num_sms, stream, prepare);
  return launch_pair<6>(I, Bt, S, num_sms, stream, prepare);
}

}  // namespace smpc
