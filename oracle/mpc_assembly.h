/* oracle/mpc_assembly.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see mpc_assembly.c). */
#ifndef ORACLE_MPC_ASSEMBLY_H
#define ORACLE_MPC_ASSEMBLY_H
#ifdef __cplusplus
extern "C" {
#endif

/* Builders of /root/reference/src/ModelPredictiveControlAPI.cpp:180-369 at run-time N, nx.
 * Outputs (row-major): H[N*N] Gbar[2N*N] Fx[N*nx] Fu[N] Fr[N*N] Sbar[2N*nx] Ku[2N] W0[2N]
 * Sx[N*nx] Su[N*N] CAB[N]. */
void orc_mpc_build(int N, int nx, const double *Ad, const double *Bd, const double *Cd,
                   const double *K, double Q, double R, double RD, int n_state_rows,
                   double u_limit, double *H, double *Gbar, double *Fx, double *Fu,
                   double *Fr, double *Sbar, double *Ku, double *W0, double *Sx,
                   double *Su, double *CAB);

/* f = Fx*X + Fu*U + Fr*ref' (cpp:374);  ub = W0 + Sbar*X + Ku*U (cpp:99) */
void orc_mpc_step_vectors(int N, int nx, const double *Fx, const double *Fu, const double *Fr,
                          const double *Sbar, const double *Ku, const double *W0,
                          const double *X, double U, const double *ref, double *f, double *ub);

double orc_mpc_lower_bound(void); /* -DBL_MAX, cpp:42 */

#ifdef __cplusplus
}
#endif
#endif
