/* split_shim.c -- entry points of the four split-field solvers (fdtdTM_get*,
 * fdtdTE_get*, nsFdtdTM_get*, nsFdtdTE_get*) implemented over the GPU engine.
 *
 * The reference keeps, per solver, five complex fields and eight dense coefficient
 * arrays that depend on the permittivity (fdtdTM.c:10-17, fdtdTE.c:10-17,
 * nsFdtdTM.c:10-18, nsFdtdTE.c:11-18).  The coefficient loops run here, on the host,
 * with the reference's expressions and libm calls (so the arrays are bit-identical);
 * the time stepping, source injection included, runs on the device.  Lifecycle as in
 * the reference: init() builds and uploads, update() is one asynchronous step, reset()
 * dumps |F|^2 on the validation circle (field_outputElliptic) and zeroes the fields,
 * finish() = reset() + free.
 */
#define _USE_MATH_DEFINES
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b200fdtd.h"
#include "host_internal.h"

#ifndef M_PI
#define M_PI 3.1415926535897932384626433832795
#endif

typedef struct SplitSolver {
  int kind;                       /* B200FDTD_TM / _TE / _NS_TM / _NS_TE          */
  b200fdtd_engine *engine;
  double *eps[3];                 /* TM: EZ, HX, HY      TE: EX, EY, HZ           */
  double *coef[8];
  double *src[2];
  dcomplex *mirror[5];            /* host mirrors of the five fields              */
  int mirror_reads[5];            /* refreshes so far; pinned in place at the first */
  int n_batch;                    /* angles of a batched run (mpifdtd_setAngleBatch), 0 = one simulation */
  int *batch_angles;
  int defer;                      /* update() collects step arguments, the engine replays them in chunks */
  int pending, pending_cap;
  b200fdtd_step_args *pending_args;
} SplitSolver;

static SplitSolver tm_plain = { .kind = B200FDTD_TM }, te_plain = { .kind = B200FDTD_TE };
static SplitSolver tm_ns = { .kind = B200FDTD_NS_TM }, te_ns = { .kind = B200FDTD_NS_TE };

static void die_on(int rc, const char *what)
{
  if (rc == B200FDTD_OK) return;
  printf("b200fdtd: %s failed (%d): %s\n", what, rc, b200fdtd_last_error());
  exit(2);
}

static int is_tm(const SplitSolver *s) { return s->kind == B200FDTD_TM || s->kind == B200FDTD_NS_TM; }

/* NS-FDTD helpers (nsFdtdTM.c:221-229) */
static double ns_beta(double alpha) { return tanh(alpha) / (1 + pow(tanh(alpha), 2)); }
static double ns_coef1(double beta) { return (1 - beta) / (1 + beta); }

/* per-cell factor of field_nsScatteredWaveNotUPML (field.c:168-174) */
static double ns_source_factor(double eps, double w_s, double k_s)
{
  double n = sqrt(eps / EPSILON_0_S);
  double u0 = sin(w_s * 0.5) / sin(k_s * 0.5);
  double u1 = sin(w_s * 0.5) / sin(n * k_s * 0.5);
  double _n = u0 / u1;
  return 1.0 / (_n * n) - 1.0;
}

/* ---- coefficient loops ----------------------------------------------------------
 * One libm-heavy expression set per cell (tanh, pow, sin, sqrt for the NS kinds), every cell
 * independent of the others: the builders take a row range [i_first, i_end) and run on all
 * host cores (16.7 M cells at 4096 x 4096 took 2-3 s single-threaded, four times the
 * 1000-step run itself). */
static void build_plain_tm(SplitSolver *s, int i_first, int i_end)                      /* fdtdTM.c:197-242 */
{
  double R = 1.0e-8, M = 2.0;
  const double sig_max = -(M + 1.0) * EPSILON_0_S * LIGHT_SPEED_S / 2.0 / N_PML * log(R);
  for (int i = i_first; i < i_end; i++)
    for (int j = 0; j < N_PY; j++) {
      int k = ind(i, j);
      double eps_ez = s->eps[0][k];
      double sig_ez_x = sig_max * field_sigmaX(i, j);
      double sig_ez_y = sig_max * field_sigmaY(i, j);
      double sig_hx_y = sig_max * field_sigmaY(i, j + 0.5);
      double sig_hx_yy = MU_0_S / EPSILON_0_S * sig_hx_y;
      double sig_hy_x = sig_max * field_sigmaX(i + 0.5, j);
      double sig_hy_xx = MU_0_S / EPSILON_0_S * sig_hy_x;
      s->coef[B200FDTD_STM_C_EZX][k]   = field_pmlCoef(eps_ez, sig_ez_x);
      s->coef[B200FDTD_STM_C_EZXLX][k] = field_pmlCoef_LXY(eps_ez, sig_ez_x);
      s->coef[B200FDTD_STM_C_EZY][k]   = field_pmlCoef(eps_ez, sig_ez_y);
      s->coef[B200FDTD_STM_C_EZYLY][k] = field_pmlCoef_LXY(eps_ez, sig_ez_y);
      s->coef[B200FDTD_STM_C_HX][k]    = field_pmlCoef(MU_0_S, sig_hx_yy);
      s->coef[B200FDTD_STM_C_HXLY][k]  = field_pmlCoef_LXY(MU_0_S, sig_hx_yy);
      s->coef[B200FDTD_STM_C_HY][k]    = field_pmlCoef(MU_0_S, sig_hy_xx);
      s->coef[B200FDTD_STM_C_HYLX][k]  = field_pmlCoef_LXY(MU_0_S, sig_hy_xx);
      s->src[0][k] = EPSILON_0_S / eps_ez - 1.0;               /* field.c:193 */
    }
}

static void build_plain_te(SplitSolver *s, int i_first, int i_end)                      /* fdtdTE.c:199-243 */
{
  double R = 1.0e-8, M = 2.0;
  const double sig_max = -(M + 1.0) * EPSILON_0_S * LIGHT_SPEED_S / 2.0 / N_PML * log(R);
  for (int i = i_first; i < i_end; i++)
    for (int j = 0; j < N_PY; j++) {
      int k = ind(i, j);
      double eps_ex = s->eps[0][k], eps_ey = s->eps[1][k];
      double sig_ex_y = sig_max * field_sigmaY(i + 0.5, j);
      double sig_ey_x = sig_max * field_sigmaX(i, j + 0.5);
      double sig_hz_x = sig_max * field_sigmaX(i + 0.5, j + 0.5);
      double sig_hz_xx = MU_0_S / EPSILON_0_S * sig_hz_x;
      double sig_hz_y = sig_max * field_sigmaY(i + 0.5, j + 0.5);
      double sig_hz_yy = MU_0_S / EPSILON_0_S * sig_hz_y;
      s->coef[B200FDTD_STE_C_EX][k]    = field_pmlCoef(eps_ex, sig_ex_y);
      s->coef[B200FDTD_STE_C_EXLY][k]  = field_pmlCoef_LXY(eps_ex, sig_ex_y);
      s->coef[B200FDTD_STE_C_EY][k]    = field_pmlCoef(eps_ey, sig_ey_x);
      s->coef[B200FDTD_STE_C_EYLX][k]  = field_pmlCoef_LXY(eps_ey, sig_ey_x);
      s->coef[B200FDTD_STE_C_HZX][k]   = field_pmlCoef(MU_0_S, sig_hz_xx);
      s->coef[B200FDTD_STE_C_HZXLX][k] = field_pmlCoef_LXY(MU_0_S, sig_hz_xx);
      s->coef[B200FDTD_STE_C_HZY][k]   = field_pmlCoef(MU_0_S, sig_hz_yy);
      s->coef[B200FDTD_STE_C_HZYLY][k] = field_pmlCoef_LXY(MU_0_S, sig_hz_yy);
      s->src[1][k] = EPSILON_0_S / eps_ey - 1.0;               /* source on Ey only, fdtdTE.c:285 */
    }
}

static void build_ns_tm(SplitSolver *s, int i_first, int i_end)                         /* nsFdtdTM.c:231-307 */
{
  const double R = 1.0e-8, M = 2.0;
  const double sig_max = -(M + 1.0) * EPSILON_0_S * C_0_S / N_PML * log(R);
  const double w_s = field_getOmega(), k_s = field_getK();
  for (int i = i_first; i < i_end; i++)
    for (int j = 0; j < N_PY; j++) {
      int k = field_index(i, j);
      double eps_ez = s->eps[0][k], eps_hx = s->eps[1][k], eps_hy = s->eps[2][k];
      double sig_ez_x = sig_max * field_sigmaX(i, j);
      double sig_ez_y = sig_max * field_sigmaY(i, j);
      double sig_hx_y = sig_max * field_sigmaY(i, j + 0.5);
      double sig_hy_x = sig_max * field_sigmaX(i + 0.5, j);
      double a_hx_y = sig_hx_y / (2 * EPSILON_0_S);
      double a_hy_x = sig_hy_x / (2 * EPSILON_0_S);
      double a_ez_x = sig_ez_x / (2 * EPSILON_0_S);
      double a_ez_y = sig_ez_y / (2 * EPSILON_0_S);
      double b_hx_y = ns_beta(a_hx_y), b_hy_x = ns_beta(a_hy_x);
      double b_ez_x = ns_beta(a_ez_x), b_ez_y = ns_beta(a_ez_y);

      double z_ez = sqrt(MU_0_S / eps_ez);
      double n_ez = sqrt(eps_ez / EPSILON_0_S);
      double k_ez_s = k_s * n_ez;
      double u_ez = sin(w_s * 0.5) / sin(k_ez_s * 0.5);
      s->coef[B200FDTD_STM_C_EZX][k]   = ns_coef1(b_ez_x);
      s->coef[B200FDTD_STM_C_EZXLX][k] = u_ez * z_ez / (1 + b_ez_x);
      s->coef[B200FDTD_STM_C_EZY][k]   = ns_coef1(b_ez_y);
      s->coef[B200FDTD_STM_C_EZYLY][k] = u_ez * z_ez / (1.0 + b_ez_x);   /* b_ez_x: upstream quirk, line 287 */

      double z_hx = sqrt(MU_0_S / eps_hx);
      double n_hx = sqrt(eps_hx / EPSILON_0_S);
      double k_hx_s = k_s * n_hx;
      double u_hx = sin(w_s * 0.5) / sin(k_hx_s * 0.5);
      s->coef[B200FDTD_STM_C_HX][k]   = ns_coef1(b_hx_y);
      s->coef[B200FDTD_STM_C_HXLY][k] = u_hx / z_hx / (1.0 + b_hx_y);

      double z_hy = sqrt(MU_0_S / eps_hy);
      double n_hy = sqrt(eps_hy / EPSILON_0_S);
      double k_hy_s = k_s * n_hy;
      double u_hy = sin(w_s * 0.5) / sin(k_hy_s * 0.5);
      s->coef[B200FDTD_STM_C_HY][k]   = ns_coef1(b_hy_x);
      s->coef[B200FDTD_STM_C_HYLX][k] = u_hy / z_hy / (1.0 + b_hy_x);

      s->src[0][k] = ns_source_factor(eps_ez, w_s, k_s);      /* on Ezy, nsFdtdTM.c:73 */
    }
}

static void build_ns_te(SplitSolver *s, int i_first, int i_end)                         /* nsFdtdTE.c:100-181: interior cells only */
{
  FieldInfo_S g = field_getFieldInfo_S();
  double R = 1.0e-8, M = 2.0;
  const double sig_max = -(M + 1.0) * EPSILON_0_S * C_0_S / g.N_PML * log(R);
  double w_s = field_getOmega(), k_s = field_getK();
  for (int i = (i_first < 1 ? 1 : i_first); i < (i_end > g.N_PX - 1 ? g.N_PX - 1 : i_end); i++)
    for (int j = 1; j < g.N_PY - 1; j++) {
      int k = field_index(i, j);
      double eps_ex = s->eps[0][k], eps_ey = s->eps[1][k], eps_hz = s->eps[2][k];
      double sig_ex_y = sig_max * field_sigmaY(i + 0.5, j);
      double sig_ey_x = sig_max * field_sigmaX(i, j + 0.5);
      double sig_hz_x = sig_max * field_sigmaX(i + 0.5, j + 0.5);
      double sig_hz_y = sig_max * field_sigmaY(i + 0.5, j + 0.5);
      double a_ex_y = sig_ex_y / (2 * eps_ex);
      double a_ey_x = sig_ey_x / (2 * eps_ey);
      double a_hz_x = sig_hz_x / (2 * eps_hz);
      double a_hz_y = sig_hz_y / (2 * eps_hz);
      double b_ex_y = ns_beta(a_ex_y), b_ey_x = ns_beta(a_ey_x);
      double b_hz_x = ns_beta(a_hz_x), b_hz_y = ns_beta(a_hz_y);

      double z_hz = sqrt(MU_0_S / eps_hz);
      double n_hz = sqrt(eps_hz / EPSILON_0_S);
      double k_hz_s = k_s * n_hz;
      double u_hz = sin(w_s * 0.5) / sin(k_hz_s * 0.5);
      s->coef[B200FDTD_STE_C_HZX][k]   = ns_coef1(b_hz_x);
      s->coef[B200FDTD_STE_C_HZXLX][k] = u_hz / z_hz / (1.0 + b_hz_x);
      s->coef[B200FDTD_STE_C_HZY][k]   = ns_coef1(b_hz_y);
      s->coef[B200FDTD_STE_C_HZYLY][k] = u_hz / z_hz / (1.0 + b_hz_y);

      double z_ex = sqrt(MU_0_S / eps_ex);
      double n_ex = sqrt(eps_ex / EPSILON_0_S);
      double k_ex_s = k_s * n_ex;
      double u_ex = sin(w_s * 0.5) / sin(k_ex_s * 0.5);
      s->coef[B200FDTD_STE_C_EX][k]   = ns_coef1(b_ex_y);
      s->coef[B200FDTD_STE_C_EXLY][k] = u_ex * z_ex / (1.0 + b_ex_y);

      double z_ey = sqrt(MU_0_S / eps_ey);
      double n_ey = sqrt(eps_ey / EPSILON_0_S);
      double k_ey_s = k_s * n_ey;
      double u_ey = sin(w_s * 0.5) / sin(k_ey_s * 0.5);
      s->coef[B200FDTD_STE_C_EY][k]   = ns_coef1(b_ey_x);
      s->coef[B200FDTD_STE_C_EYLX][k] = u_ey * z_ey / (1.0 + b_ey_x);

      s->src[0][k] = ns_source_factor(eps_ex, w_s, k_s);      /* on Ex, nsFdtdTE.c:247-248 */
      s->src[1][k] = ns_source_factor(eps_ey, w_s, k_s);      /* on Ey, nsFdtdTE.c:249-250 */
    }
}

#include <pthread.h>
#include <unistd.h>
typedef struct RowJob { void (*fn)(SplitSolver *, int, int); SplitSolver *s; int i0, i1; } RowJob;
static void *row_job(void *arg) { RowJob *j = (RowJob *)arg; j->fn(j->s, j->i0, j->i1); return NULL; }

static void build_rows_parallel(void (*fn)(SplitSolver *, int, int), SplitSolver *s)
{
  long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
  const char *cap = getenv("MPIFDTD_HOST_THREADS");
  if (cap != NULL && atoi(cap) > 0) ncpu = atoi(cap);
  int nthr = (int)(ncpu < 1 ? 1 : (ncpu > 64 ? 64 : ncpu));
  if (nthr > N_PX) nthr = N_PX > 0 ? N_PX : 1;
  pthread_t tid[64];
  RowJob job[64];
  int spawned[64];
  for (int t = 0; t < nthr; t++) {
    job[t] = (RowJob){ fn, s, (int)((long long)N_PX * t / nthr), (int)((long long)N_PX * (t + 1) / nthr) };
    spawned[t] = (t < nthr - 1) && pthread_create(&tid[t], NULL, row_job, &job[t]) == 0;
    if (!spawned[t]) row_job(&job[t]);          /* last chunk, or no thread to be had */
  }
  for (int t = 0; t < nthr; t++)
    if (spawned[t]) pthread_join(tid[t], NULL);
}

/* ---- lean form (kinds 0, 1, 6; include/b200fdtd.h "lean form") ---------------------
 * The same expressions as the dense builders above, evaluated once per row / column where
 * they do not depend on eps; what does depend on eps goes to the device as the eps map
 * itself (kinds 0, 1) or as one G array per curl coefficient (kind 6).
 * MPIFDTD_SPLIT_DENSE=1 keeps the dense form (A/B tests). */
static int lean_wanted(int kind)
{
  const char *v = getenv("MPIFDTD_SPLIT_DENSE");
  if (v != NULL && v[0] == '1') return 0;
  return kind == B200FDTD_TM || kind == B200FDTD_TE || kind == B200FDTD_NS_TM;
}

static void build_lean_tables(int kind, double *ti, double *tj)
{
  const double R = 1.0e-8, M = 2.0;
  for (int k = 0; k < B200FDTD_SPLIT_TABS * N_PX; k++) ti[k] = 1.0;
  for (int k = 0; k < B200FDTD_SPLIT_TABS * N_PY; k++) tj[k] = 1.0;
  if (kind == B200FDTD_TM) {                                   /* fdtdTM.c:204-238 */
    const double sig_max = -(M + 1.0) * EPSILON_0_S * LIGHT_SPEED_S / 2.0 / N_PML * log(R);
    for (int i = 0; i < N_PX; i++) {
      double sig_hy_x = sig_max * field_sigmaX(i + 0.5, 0);
      double sig_hy_xx = MU_0_S / EPSILON_0_S * sig_hy_x;
      ti[B200FDTD_LTM_I_SIG_EZ_X * N_PX + i] = sig_max * field_sigmaX(i, 0);
      ti[B200FDTD_LTM_I_C_HY * N_PX + i]     = field_pmlCoef(MU_0_S, sig_hy_xx);
      ti[B200FDTD_LTM_I_C_HYLX * N_PX + i]   = field_pmlCoef_LXY(MU_0_S, sig_hy_xx);
    }
    for (int j = 0; j < N_PY; j++) {
      double sig_hx_y = sig_max * field_sigmaY(0, j + 0.5);
      double sig_hx_yy = MU_0_S / EPSILON_0_S * sig_hx_y;
      tj[B200FDTD_LTM_J_SIG_EZ_Y * N_PY + j] = sig_max * field_sigmaY(0, j);
      tj[B200FDTD_LTM_J_C_HX * N_PY + j]     = field_pmlCoef(MU_0_S, sig_hx_yy);
      tj[B200FDTD_LTM_J_C_HXLY * N_PY + j]   = field_pmlCoef_LXY(MU_0_S, sig_hx_yy);
    }
  } else if (kind == B200FDTD_TE) {                            /* fdtdTE.c:203-240 */
    const double sig_max = -(M + 1.0) * EPSILON_0_S * LIGHT_SPEED_S / 2.0 / N_PML * log(R);
    for (int i = 0; i < N_PX; i++) {
      double sig_hz_x = sig_max * field_sigmaX(i + 0.5, 0.5);
      double sig_hz_xx = MU_0_S / EPSILON_0_S * sig_hz_x;
      ti[B200FDTD_LTE_I_SIG_EY_X * N_PX + i] = sig_max * field_sigmaX(i, 0.5);
      ti[B200FDTD_LTE_I_C_HZX * N_PX + i]    = field_pmlCoef(MU_0_S, sig_hz_xx);
      ti[B200FDTD_LTE_I_C_HZXLX * N_PX + i]  = field_pmlCoef_LXY(MU_0_S, sig_hz_xx);
    }
    for (int j = 0; j < N_PY; j++) {
      double sig_hz_y = sig_max * field_sigmaY(0.5, j + 0.5);
      double sig_hz_yy = MU_0_S / EPSILON_0_S * sig_hz_y;
      tj[B200FDTD_LTE_J_SIG_EX_Y * N_PY + j] = sig_max * field_sigmaY(0.5, j);
      tj[B200FDTD_LTE_J_C_HZY * N_PY + j]    = field_pmlCoef(MU_0_S, sig_hz_yy);
      tj[B200FDTD_LTE_J_C_HZYLY * N_PY + j]  = field_pmlCoef_LXY(MU_0_S, sig_hz_yy);
    }
  } else {                                                     /* NS TM, nsFdtdTM.c:231-305 */
    const double sig_max = -(M + 1.0) * EPSILON_0_S * C_0_S / N_PML * log(R);
    for (int i = 0; i < N_PX; i++) {
      double b_ez_x = ns_beta(sig_max * field_sigmaX(i, 0) / (2 * EPSILON_0_S));
      double b_hy_x = ns_beta(sig_max * field_sigmaX(i + 0.5, 0) / (2 * EPSILON_0_S));
      ti[B200FDTD_LNS_I_C_EZX * N_PX + i]  = ns_coef1(b_ez_x);
      ti[B200FDTD_LNS_I_DEN_EZ * N_PX + i] = 1 + b_ez_x;
      ti[B200FDTD_LNS_I_C_HY * N_PX + i]   = ns_coef1(b_hy_x);
      ti[B200FDTD_LNS_I_DEN_HY * N_PX + i] = 1.0 + b_hy_x;
    }
    for (int j = 0; j < N_PY; j++) {
      double b_ez_y = ns_beta(sig_max * field_sigmaY(0, j) / (2 * EPSILON_0_S));
      double b_hx_y = ns_beta(sig_max * field_sigmaY(0, j + 0.5) / (2 * EPSILON_0_S));
      tj[B200FDTD_LNS_J_C_EZY * N_PY + j]  = ns_coef1(b_ez_y);
      tj[B200FDTD_LNS_J_C_HX * N_PY + j]   = ns_coef1(b_hx_y);
      tj[B200FDTD_LNS_J_DEN_HX * N_PY + j] = 1.0 + b_hx_y;
    }
  }
}

/* kind 6: the eps-dependent numerators of the three curl coefficients and the source factor
 * (nsFdtdTM.c:279-305, field.c:168-174); written into the coef slots the device reads them from */
static void build_ns_tm_numerators(SplitSolver *s, int i_first, int i_end)
{
  const double w_s = field_getOmega(), k_s = field_getK();
  for (int i = i_first; i < i_end; i++)
    for (int j = 0; j < N_PY; j++) {
      int k = field_index(i, j);
      double eps_ez = s->eps[0][k], eps_hx = s->eps[1][k], eps_hy = s->eps[2][k];
      double z_ez = sqrt(MU_0_S / eps_ez);
      double n_ez = sqrt(eps_ez / EPSILON_0_S);
      double k_ez_s = k_s * n_ez;
      double u_ez = sin(w_s * 0.5) / sin(k_ez_s * 0.5);
      s->coef[B200FDTD_STM_C_EZXLX][k] = u_ez * z_ez;
      double z_hx = sqrt(MU_0_S / eps_hx);
      double n_hx = sqrt(eps_hx / EPSILON_0_S);
      double k_hx_s = k_s * n_hx;
      double u_hx = sin(w_s * 0.5) / sin(k_hx_s * 0.5);
      s->coef[B200FDTD_STM_C_HXLY][k] = u_hx / z_hx;
      double z_hy = sqrt(MU_0_S / eps_hy);
      double n_hy = sqrt(eps_hy / EPSILON_0_S);
      double k_hy_s = k_s * n_hy;
      double u_hy = sin(w_s * 0.5) / sin(k_hy_s * 0.5);
      s->coef[B200FDTD_STM_C_HYLX][k] = u_hy / z_hy;
      s->src[0][k] = ns_source_factor(eps_ez, w_s, k_s);
    }
}

/* ---- init ------------------------------------------------------------------------ */
static void free_host(SplitSolver *s)
{
  for (int m = 0; m < 3; m++) { free(s->eps[m]); s->eps[m] = NULL; }
  for (int m = 0; m < 8; m++) { free(s->coef[m]); s->coef[m] = NULL; }
  for (int m = 0; m < 2; m++) { free(s->src[m]); s->src[m] = NULL; }
}

/* host half of init(): permittivity maps, the eight coefficient arrays, source factors */
static void build_host(SplitSolver *s, int lean)
{
  FieldInfo_S g = field_getFieldInfo_S();
  const size_t n = (size_t)g.N_CELL;
  free_host(s);
  if (lean) {
    /* only what the lean kernels read: EPS_EZ (kind 0); EPS_EX, EPS_EY (kind 1); all three
     * maps, three numerator arrays and the source factor (kind 6) */
    if (s->kind == B200FDTD_TM) {
      s->eps[0] = newDouble(g.N_CELL);
      mpifdtd_fill_eps(s->eps[0], 0, 0, D_XY);
    } else if (s->kind == B200FDTD_TE) {
      s->eps[0] = newDouble(g.N_CELL);  s->eps[1] = newDouble(g.N_CELL);
      mpifdtd_fill_eps(s->eps[0], 0.5, 0, D_Y);
      mpifdtd_fill_eps(s->eps[1], 0, 0.5, D_X);
    } else {
      for (int m = 0; m < 3; m++) s->eps[m] = newDouble(g.N_CELL);
      mpifdtd_fill_eps(s->eps[0], 0, 0, D_XY);
      mpifdtd_fill_eps(s->eps[1], 0, 0.5, D_Y);
      mpifdtd_fill_eps(s->eps[2], 0.5, 0, D_X);
      s->coef[B200FDTD_STM_C_EZXLX] = newDouble(g.N_CELL);
      s->coef[B200FDTD_STM_C_HXLY] = newDouble(g.N_CELL);
      s->coef[B200FDTD_STM_C_HYLX] = newDouble(g.N_CELL);
      s->src[0] = newDouble(g.N_CELL);
      build_rows_parallel(build_ns_tm_numerators, s);
    }
    return;
  }
  for (int m = 0; m < 3; m++) s->eps[m] = newDouble(g.N_CELL);
  for (int m = 0; m < 8; m++) s->coef[m] = newDouble(g.N_CELL);
  for (int m = 0; m < 2; m++) s->src[m] = newDouble(g.N_CELL);

  /* permittivity maps (fdtdTM.c:209-211, fdtdTE.c:206-208, nsFdtdTM.c:241-243,
   * nsFdtdTE.c:117-119).  NS TE evaluates interior cells only; its ring stays 0. */
  if (is_tm(s)) {
    mpifdtd_fill_eps(s->eps[0], 0, 0, D_XY);
    mpifdtd_fill_eps(s->eps[1], 0, 0.5, D_Y);
    mpifdtd_fill_eps(s->eps[2], 0.5, 0, D_X);
  } else {
    mpifdtd_fill_eps(s->eps[0], 0.5, 0, D_Y);
    mpifdtd_fill_eps(s->eps[1], 0, 0.5, D_X);
    if (s->kind == B200FDTD_NS_TE) {
      mpifdtd_fill_eps(s->eps[2], 0.5, 0.5, D_XY);
      for (int m = 0; m < 3; m++)
        for (int i = 0; i < g.N_PX; i++)
          for (int j = 0; j < g.N_PY; j++)
            if (i == 0 || j == 0 || i == g.N_PX - 1 || j == g.N_PY - 1)
              s->eps[m][field_index(i, j)] = 0;
    } else {                                                  /* 0.5*(D_X + D_Y), fdtdTE.c:208 */
      double *tmp = newDouble(g.N_CELL);
      mpifdtd_fill_eps(s->eps[2], 0.5, 0.5, D_X);
      mpifdtd_fill_eps(tmp, 0.5, 0.5, D_Y);
      for (size_t k = 0; k < n; k++) s->eps[2][k] = 0.5 * (s->eps[2][k] + tmp[k]);
      free(tmp);
    }
  }
  switch (s->kind) {
  case B200FDTD_TM:    build_rows_parallel(build_plain_tm, s); break;
  case B200FDTD_TE:    build_rows_parallel(build_plain_te, s); break;
  case B200FDTD_NS_TM: build_rows_parallel(build_ns_tm, s);    break;
  default:             build_rows_parallel(build_ns_te, s);    break;
  }
}

/* NS TE (id 7): outside the absorbing frame sigma == 0, so beta == 0 and the four decay coefficients
 * of nsFdtdTE.c:155-177 are exactly 1.0 while C_HZXLX and C_HZYLY are the same u/z/(1 + 0).  The
 * largest centred rectangle where the arrays built above really say so -- checked cell by cell, not
 * assumed -- goes to the engine, whose kernels then read three coefficient arrays instead of eight
 * there (b200fdtd_set_split_interior).  MPIFDTD_SPLIT_DENSE=1 keeps the dense form everywhere. */
static void set_ns_te_interior(SplitSolver *s)
{
  FieldInfo_S g = field_getFieldInfo_S();
  const char *v = getenv("MPIFDTD_SPLIT_DENSE");
  if (v != NULL && v[0] == '1') return;
  for (int margin = g.N_PML + 1; margin < g.N_PML + 6; margin++) {
    const int i_lo = margin, i_hi = g.N_PX - 1 - margin, j_lo = margin, j_hi = g.N_PY - 1 - margin;
    if (i_hi - i_lo < 8 || j_hi - j_lo < 8) return;
    int ok = 1;
    for (int i = i_lo; i <= i_hi && ok; i++)
      for (int j = j_lo; j <= j_hi; j++) {
        const int k = field_index(i, j);
        if (s->coef[B200FDTD_STE_C_HZX][k] != 1.0 || s->coef[B200FDTD_STE_C_HZY][k] != 1.0 ||
            s->coef[B200FDTD_STE_C_EX][k] != 1.0 || s->coef[B200FDTD_STE_C_EY][k] != 1.0 ||
            s->coef[B200FDTD_STE_C_HZXLX][k] != s->coef[B200FDTD_STE_C_HZYLY][k]) { ok = 0; break; }
      }
    if (ok) {
      die_on(b200fdtd_set_split_interior(s->engine, i_lo, i_hi, j_lo, j_hi), "b200fdtd_set_split_interior");
      return;
    }
  }
}

static void fill_batch_cw(int kind, int angle_deg, b200fdtd_batch_cw *b);
static void solver_init(SplitSolver *s)
{
  FieldInfo_S g = field_getFieldInfo_S();
  const int lean = lean_wanted(s->kind);
  build_host(s, lean);

  b200fdtd_grid grid;
  memset(&grid, 0, sizeof grid);
  grid.kind = s->kind;
  grid.n_px = g.N_PX;  grid.n_py = g.N_PY;  grid.n_pml = g.N_PML;
  grid.j0 = 0;         grid.nj = g.N_PY;
  grid.i_lo = 1;       grid.i_hi = g.N_PX - 2;
  grid.j_lo = 1;       grid.j_hi = g.N_PY - 2;
  grid.device = -1;
  grid.mu0 = MU_0_S;
  /* an angle batch: main.c as shipped sweeps the incidence angle with NS_TE_2D, one
   * simulation after the other (main.c:152,183-211); the simulations share every coefficient array
   * and differ in the CW source only, so they run as planes of one engine, advanced together */
  {
    const int *req = NULL;
    const int n_req = mpifdtd_angle_batch_requested(&req);
    s->n_batch = n_req > 1 ? n_req : 0;
    free(s->batch_angles);  s->batch_angles = NULL;
    if (s->n_batch) {
      s->batch_angles = (int *)malloc(sizeof(int) * (size_t)s->n_batch);
      memcpy(s->batch_angles, req, sizeof(int) * (size_t)s->n_batch);
    }
  }
  grid.n_batch = s->n_batch;
  {
    const char *v = getenv("MPIFDTD_DEFER_STEPS"), *c = getenv("MPIFDTD_DEFER_CHUNK");
    s->defer = !(v != NULL && v[0] == '0');
    s->pending = 0;
    s->pending_cap = c != NULL && atoi(c) > 0 ? atoi(c) : 256;
    free(s->pending_args);
    s->pending_args = s->defer ? (b200fdtd_step_args *)malloc(sizeof(b200fdtd_step_args) * (size_t)s->pending_cap) : NULL;
  }
  die_on(b200fdtd_create(&grid, &s->engine), "b200fdtd_create");
  if (s->n_batch) {
    b200fdtd_batch_cw *rec = (b200fdtd_batch_cw *)calloc((size_t)s->n_batch, sizeof *rec);
    for (int k = 0; k < s->n_batch; k++) fill_batch_cw(s->kind, s->batch_angles[k], &rec[k]);
    die_on(b200fdtd_set_batch_cw(s->engine, rec), "b200fdtd_set_batch_cw");
    free(rec);
  }
  if (lean) {
    double *ti = (double *)malloc(sizeof(double) * B200FDTD_SPLIT_TABS * (size_t)g.N_PX);
    double *tj = (double *)malloc(sizeof(double) * B200FDTD_SPLIT_TABS * (size_t)g.N_PY);
    build_lean_tables(s->kind, ti, tj);
    die_on(b200fdtd_set_split_tables(s->engine, ti, tj), "b200fdtd_set_split_tables");
    free(ti); free(tj);
    if (s->kind == B200FDTD_NS_TM) {
      const int slots[3] = { B200FDTD_STM_C_EZXLX, B200FDTD_STM_C_HXLY, B200FDTD_STM_C_HYLX };
      for (int m = 0; m < 3; m++)
        die_on(b200fdtd_set_dense(s->engine, slots[m], s->coef[slots[m]]), "b200fdtd_set_dense");
      die_on(b200fdtd_set_dense(s->engine, B200FDTD_DENSE_SRC0, s->src[0]), "b200fdtd_set_dense(src0)");
    } else {
      die_on(b200fdtd_set_eps(s->engine, 0, s->eps[0]), "b200fdtd_set_eps");
      if (s->kind == B200FDTD_TE) die_on(b200fdtd_set_eps(s->engine, 1, s->eps[1]), "b200fdtd_set_eps");
    }
    return;
  }
  for (int m = 0; m < 8; m++)
    die_on(b200fdtd_set_dense(s->engine, m, s->coef[m]), "b200fdtd_set_dense");
  die_on(b200fdtd_set_dense(s->engine, B200FDTD_DENSE_SRC0, s->src[0]), "b200fdtd_set_dense(src0)");
  die_on(b200fdtd_set_dense(s->engine, B200FDTD_DENSE_SRC1, s->src[1]), "b200fdtd_set_dense(src1)");
  if (s->kind == B200FDTD_NS_TE) set_ns_te_interior(s);
  /* the pinned getter mirrors are allocated by the first getter call (solver_field) */
}

/* ---- update ------------------------------------------------------------------------ */
static void fill_cw(b200fdtd_cw *c, int ns, double gap_x, double gap_y, double dot)
{
  double time = field_getTime(), w_s = field_getOmega(), k_s = field_getK();
  double rad = field_getWaveAngle() * M_PI / 180.0;
  c->enabled = 1;
  c->two_term = 1;
  c->gap_x = gap_x;  c->gap_y = gap_y;
  c->ks_cos = cos(rad) * k_s;
  c->ks_sin = sin(rad) * k_s;
  if (ns) {                                  /* field.c:160,174 */
    c->scale = field_getRayCoef() * dot;
    c->phase_a = w_s * (time + 1.0);
    c->phase_b = w_s * (time);
  } else {                                   /* field.c:184,193 */
    c->scale = field_getRayCoef();
    c->phase_a = w_s * (time + 0.5);
    c->phase_b = w_s * (time - 0.5);
  }
}

void mpifdtd_split_step_args(int kind, b200fdtd_step_args *a)
{
  memset(a, 0, sizeof *a);
  a->time = field_getTime();
  a->ray_coef = field_getRayCoef();
  switch (kind) {
  case B200FDTD_TM:                                           /* fdtdTM.c:294 */
    fill_cw(&a->cw[0], 0, 0.0, 0.0, 1.0);
    break;
  case B200FDTD_TE:                                           /* fdtdTE.c:285 */
    fill_cw(&a->cw[1], 0, 0.0, 0.5, 1.0);
    break;
  case B200FDTD_NS_TM:                                        /* nsFdtdTM.c:73 */
    fill_cw(&a->cw[0], 1, 0, 0, 1.0);
    break;
  default: {                                                  /* nsFdtdTE.c:243-250 */
    WaveInfo_S w = field_getWaveInfo_S();
    double co = cos((w.Angle_deg + 90) * M_PI / 180.0);
    double si = sin((w.Angle_deg + 90) * M_PI / 180.0);
    if (co != 0.0) fill_cw(&a->cw[0], 1, 0, 0.5, co);
    if (si != 0.0) fill_cw(&a->cw[1], 1, 0, 0.5, si);
    break;
  }
  }
  if (kind == B200FDTD_NS_TM || kind == B200FDTD_NS_TE) {     /* nsFdtdTM.c:115-117 */
    double k_s = field_getK();
    double r = 1.0 / 6.0 + k_s * k_s / 180.0 - pow(k_s, 4) / 23040;
    a->ns_r2 = r / 2.0;
  }
}

/* Batched engine: step_args carries what the simulations share (gaps, phases, scale = ray_coef);
 * the angle-dependent rest sits in the per-simulation records below. */
static void split_step_args_batched(int kind, b200fdtd_step_args *a)
{
  mpifdtd_split_step_args(kind, a);
  memset(a->cw, 0, sizeof a->cw);
  switch (kind) {
  case B200FDTD_TM:    fill_cw(&a->cw[0], 0, 0.0, 0.0, 1.0); break;
  case B200FDTD_TE:    fill_cw(&a->cw[1], 0, 0.0, 0.5, 1.0); break;
  case B200FDTD_NS_TM: fill_cw(&a->cw[0], 1, 0, 0, 1.0);     break;
  default:             fill_cw(&a->cw[0], 1, 0, 0.5, 1.0);  fill_cw(&a->cw[1], 1, 0, 0.5, 1.0);  break;
  }
}

/* what fill_cw / mpifdtd_split_step_args take from the incidence angle, for one simulation */
static void fill_batch_cw(int kind, int angle_deg, b200fdtd_batch_cw *b)
{
  const double k_s = field_getK();
  const double rad = angle_deg * M_PI / 180.0;
  memset(b, 0, sizeof *b);
  b->ks_cos = cos(rad) * k_s;
  b->ks_sin = sin(rad) * k_s;
  b->dot[0] = b->dot[1] = 1.0;
  if (kind == B200FDTD_NS_TE) {                               /* nsFdtdTE.c:243-250 */
    double co = cos((angle_deg + 90) * M_PI / 180.0);
    double si = sin((angle_deg + 90) * M_PI / 180.0);
    b->dot[0] = co;  b->enabled[0] = co != 0.0;
    b->dot[1] = si;  b->enabled[1] = si != 0.0;
  } else {
    b->enabled[kind == B200FDTD_TE ? 1 : 0] = 1;
  }
}

/* update() is asynchronous anyway, so it only RECORDS the step's arguments (the CW source's phases
 * and ramp, which the host clock advances between calls); the pending steps go to the engine in
 * one b200fdtd_run_split_steps call -- a CUDA-graph replay, each kernel reading its own step's
 * record -- when somebody looks (a getter, reset/finish, the engine handle) or MPIFDTD_DEFER_CHUNK
 * (256) steps have piled up.  Bit-identical to stepping one by one; on small grids, where a step
 * is a few microseconds of device work, it removes the per-launch host cost.
 * MPIFDTD_DEFER_STEPS=0 hands every step over immediately. */
static void flush_pending(SplitSolver *s)
{
  if (s->engine == NULL || s->pending == 0) return;
  const int n = s->pending;
  s->pending = 0;
  die_on(b200fdtd_run_split_steps(s->engine, s->pending_args, n), "b200fdtd_run_split_steps");
}

void mpifdtd_split_flush_pending_steps(void)
{
  flush_pending(&tm_plain);  flush_pending(&te_plain);  flush_pending(&tm_ns);  flush_pending(&te_ns);
}

static void solver_update(SplitSolver *s)
{
  b200fdtd_step_args a;
  if (s->n_batch) split_step_args_batched(s->kind, &a);
  else            mpifdtd_split_step_args(s->kind, &a);
  if (!s->defer) {
    die_on(b200fdtd_step(s->engine, &a), "b200fdtd_step");
    return;
  }
  s->pending_args[s->pending++] = a;
  if (s->pending >= s->pending_cap) flush_pending(s);
}

/* ---- getters / reset / finish ------------------------------------------------------ */
static dcomplex *solver_field(SplitSolver *s, int slot)
{
  if (s->engine == NULL) return NULL;
  flush_pending(s);
  const size_t mirror_bytes = sizeof(dcomplex) * (size_t)field_getFieldInfo_S().N_CELL;
  if (s->mirror[slot] == NULL) {        /* only if somebody looks */
    die_on(b200fdtd_mirror_alloc((void **)&s->mirror[slot], mirror_bytes), "mirror_alloc");
    s->mirror_reads[slot] = 0;
  }
  if (++s->mirror_reads[slot] == 1) die_on(b200fdtd_mirror_pin(s->mirror[slot], mirror_bytes), "mirror_pin");
  die_on(b200fdtd_get_field(s->engine, slot, (double *)s->mirror[slot]), "b200fdtd_get_field");
  return s->mirror[slot];
}

static void solver_reset(SplitSolver *s)
{
  if (s->engine == NULL) return;
  flush_pending(s);
  /* validation-circle dump of the drawn field (fdtdTM.c:154-160, fdtdTE.c:273-278,
   * nsFdtdTM.c:159-164, nsFdtdTE.c:218-223) */
  static const char *const stem[] = { "tm_%dnm.txt", "te_%dnm.txt", NULL, NULL, NULL, NULL,
                                      "ns_tm_%dnm.txt", "ns_te_%dnm.txt" };
  FieldInfo phys = field_getFieldInfo();
  char name[128], per_angle[160];
  sprintf(name, stem[s->kind], phys.h_u_nm);
  const int slot = is_tm(s) ? B200FDTD_STM_EZ : B200FDTD_STE_EY;
  /* An angle batch writes the dump of every simulation in sweep order under the reference's name
   * -- which carries no angle, so upstream's sweep keeps the last one, and so does this -- and,
   * so that nothing is lost, also as "<angle>[deg]_<name>". */
  for (int k = 0; k < (s->n_batch ? s->n_batch : 1); k++) {
    if (s->n_batch) {
      die_on(b200fdtd_select_batch(s->engine, k), "b200fdtd_select_batch");
      sprintf(per_angle, "%d[deg]_%s", s->batch_angles[k], name);
      field_outputElliptic(per_angle, solver_field(s, slot));
    }
    field_outputElliptic(name, solver_field(s, slot));
  }
  if (s->n_batch) die_on(b200fdtd_select_batch(s->engine, 0), "b200fdtd_select_batch");
  die_on(b200fdtd_zero_state(s->engine), "b200fdtd_zero_state");
}

static void solver_finish(SplitSolver *s)
{
  if (s->engine == NULL) return;
  solver_reset(s);
  die_on(b200fdtd_destroy(s->engine), "b200fdtd_destroy");
  s->engine = NULL;
  free(s->batch_angles);  s->batch_angles = NULL;  s->n_batch = 0;
  free(s->pending_args);  s->pending_args = NULL;  s->pending = 0;
  free_host(s);
  for (int m = 0; m < 5; m++) {
    b200fdtd_mirror_free(s->mirror[m], s->mirror_reads[m] >= 1);
    s->mirror[m] = NULL;  s->mirror_reads[m] = 0;
  }
}

/* which simulation of an angle batch the getters show (mpifdtd_selectAngle) */
void mpifdtd_split_select_angle(int index)
{
  SplitSolver *all[] = { &tm_plain, &te_plain, &tm_ns, &te_ns };
  for (int n = 0; n < 4; n++)
    if (all[n]->engine != NULL && all[n]->n_batch > 1)
      die_on(b200fdtd_select_batch(all[n]->engine, index), "b200fdtd_select_batch");
}

/* test hooks: host-built dense arrays and the engine of a split solver */
void mpifdtd_split_prepare_host(int kind)
{
  SplitSolver *all[] = { &tm_plain, &te_plain, &tm_ns, &te_ns };
  for (int n = 0; n < 4; n++)
    if (all[n]->kind == kind) build_host(all[n], 0);        /* the dense arrays, whatever the solver runs */
}
/* test hooks for the lean form: the host arrays a lean init() builds (eps maps; kind 6: the
 * three numerator arrays in their coef slots + the source factor) and the 1-D tables */
void mpifdtd_split_prepare_host_lean(int kind)
{
  SplitSolver *all[] = { &tm_plain, &te_plain, &tm_ns, &te_ns };
  for (int n = 0; n < 4; n++)
    if (all[n]->kind == kind && kind != B200FDTD_NS_TE) build_host(all[n], 1);
}
void mpifdtd_split_lean_tables(int kind, double *tab_i, double *tab_j)
{
  build_lean_tables(kind, tab_i, tab_j);
}
const double *mpifdtd_split_dense(int kind, int slot)
{
  SplitSolver *all[] = { &tm_plain, &te_plain, &tm_ns, &te_ns };
  for (int n = 0; n < 4; n++)
    if (all[n]->kind == kind)
      return slot < 8 ? all[n]->coef[slot] : (slot < 10 ? all[n]->src[slot - 8] : all[n]->eps[slot - 10]);
  return NULL;
}
b200fdtd_engine *mpifdtd_split_engine(int kind)
{
  SplitSolver *all[] = { &tm_plain, &te_plain, &tm_ns, &te_ns };
  for (int n = 0; n < 4; n++)
    if (all[n]->kind == kind) { flush_pending(all[n]); return all[n]->engine; }   /* whoever takes the handle sees every update() */
  return NULL;
}

/* ---- exported entry points ----------------------------------------------------------- */
#define SPLIT_ENTRY_POINTS(PREFIX, OBJ)                                              \
  static void PREFIX##_update(void) { solver_update(&OBJ); }                         \
  static void PREFIX##_init(void)   { solver_init(&OBJ); }                           \
  static void PREFIX##_reset(void)  { solver_reset(&OBJ); }                          \
  static void PREFIX##_finish(void) { solver_finish(&OBJ); }                         \
  void (*PREFIX##_getUpdate(void))(void) { return PREFIX##_update; }                 \
  void (*PREFIX##_getInit(void))(void)   { return PREFIX##_init; }                   \
  void (*PREFIX##_getReset(void))(void)  { return PREFIX##_reset; }                  \
  void (*PREFIX##_getFinish(void))(void) { return PREFIX##_finish; }

SPLIT_ENTRY_POINTS(fdtdTM, tm_plain)
double complex *fdtdTM_getHx(void)  { return solver_field(&tm_plain, B200FDTD_STM_HX); }
double complex *fdtdTM_getHy(void)  { return solver_field(&tm_plain, B200FDTD_STM_HY); }
double complex *fdtdTM_getEz(void)  { return solver_field(&tm_plain, B200FDTD_STM_EZ); }
double complex *fdtdTM_getEzx(void) { return solver_field(&tm_plain, B200FDTD_STM_EZX); }
double complex *fdtdTM_getEzy(void) { return solver_field(&tm_plain, B200FDTD_STM_EZY); }
double *fdtdTM_getEps(void) { return tm_plain.eps[0]; }                    /* EPS_EZ */

SPLIT_ENTRY_POINTS(fdtdTE, te_plain)
double complex *fdtdTE_getEx(void)  { return solver_field(&te_plain, B200FDTD_STE_EX); }
double complex *fdtdTE_getEy(void)  { return solver_field(&te_plain, B200FDTD_STE_EY); }
double complex *fdtdTE_getHz(void)  { return solver_field(&te_plain, B200FDTD_STE_HZ); }
double complex *fdtdTE_getHzx(void) { return solver_field(&te_plain, B200FDTD_STE_HZX); }
double complex *fdtdTE_getHzy(void) { return solver_field(&te_plain, B200FDTD_STE_HZY); }
double *fdtdTE_getEps(void) { return te_plain.eps[1]; }                    /* EPS_EY, fdtdTE.c:63-66 */

SPLIT_ENTRY_POINTS(nsFdtdTM, tm_ns)
double complex *nsFdtdTM_getHx(void)  { return solver_field(&tm_ns, B200FDTD_STM_HX); }
double complex *nsFdtdTM_getHy(void)  { return solver_field(&tm_ns, B200FDTD_STM_HY); }
double complex *nsFdtdTM_getEz(void)  { return solver_field(&tm_ns, B200FDTD_STM_EZ); }
double complex *nsFdtdTM_getEzx(void) { return solver_field(&tm_ns, B200FDTD_STM_EZX); }
double complex *nsFdtdTM_getEzy(void) { return solver_field(&tm_ns, B200FDTD_STM_EZY); }
double *nsFdtdTM_getEps(void)  { return tm_ns.eps[0]; }
double *nsFdtdTM_getEpsX(void) { return tm_ns.eps[1]; }                    /* EPS_HX */
double *nsFdtdTM_getEpsY(void) { return tm_ns.eps[2]; }                    /* EPS_HY */
double *nsFdtdTM_getEpsZ(void) { return tm_ns.eps[0]; }                    /* EPS_EZ */

SPLIT_ENTRY_POINTS(nsFdtdTE, te_ns)
double complex *nsFdtdTE_getEx(void)  { return solver_field(&te_ns, B200FDTD_STE_EX); }
double complex *nsFdtdTE_getEy(void)  { return solver_field(&te_ns, B200FDTD_STE_EY); }
double complex *nsFdtdTE_getHz(void)  { return solver_field(&te_ns, B200FDTD_STE_HZ); }
double complex *nsFdtdTE_getHzx(void) { return solver_field(&te_ns, B200FDTD_STE_HZX); }
double complex *nsFdtdTE_getHzy(void) { return solver_field(&te_ns, B200FDTD_STE_HZY); }
double *nsFdtdTE_getEps(void)  { return te_ns.eps[1]; }                    /* EPS_EY */
double *nsFdtdTE_getEpsX(void) { return te_ns.eps[0]; }
double *nsFdtdTE_getEpsY(void) { return te_ns.eps[1]; }
double *nsFdtdTE_getEpsZ(void) { return te_ns.eps[2]; }

/* solver.h:7-20: the struct upstream declares and returns zero-filled (its assignments are
 * commented out, nsFdtdTM.c:45-61).  Kept for link compatibility, filled in for real. */
typedef struct Solver {
  void (*update)(void), (*finish)(void), (*init)(void), (*reset)(void);
  dcomplex *(*getDataX)(void), *(*getDataY)(void), *(*getDataZ)(void);
  dcomplex *(*getEpsX)(void), *(*getEpsY)(void), *(*getEpsZ)(void);
} Solver;
Solver *nsFdtdTM_getSolver(void)
{
  static Solver solver;
  solver.update = nsFdtdTM_update;  solver.finish = nsFdtdTM_finish;
  solver.init = nsFdtdTM_init;      solver.reset = nsFdtdTM_reset;
  solver.getDataX = nsFdtdTM_getHx; solver.getDataY = nsFdtdTM_getHy; solver.getDataZ = nsFdtdTM_getEz;
  return &solver;
}
Solver *nsFdtdTE_getSolver(void)
{
  static Solver solver;
  solver.update = nsFdtdTE_update;  solver.finish = nsFdtdTE_finish;
  solver.init = nsFdtdTE_init;      solver.reset = nsFdtdTE_reset;
  solver.getDataX = nsFdtdTE_getEx; solver.getDataY = nsFdtdTE_getEy; solver.getDataZ = nsFdtdTE_getHz;
  return &solver;
}
