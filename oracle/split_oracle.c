/* split_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of rennone/mpiFDTD's four split-field solvers, used by tests/ as the
 * CHECKER for the CUDA path of solver ids 0, 1, 6, 7.  Nothing under mpifdtd_b200/ links,
 * imports or calls it; the product has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks this file bit for bit against
 * oracle/_ref/libref.so (the reference compiled in place) -- coefficients and all five fields
 * after a few hundred steps -- and against the reference-recorded snapshots
 * tests/golden/split_kind*.npz.
 *
 * What is restated, with the reference lines each block follows:
 *   PML profile ................................ field.c:259-283
 *   field_pmlCoef / field_pmlCoef_LXY .......... field.c:288-295
 *   soft-start clock ........................... field.c:312-315
 *   id 0  Yee TM + Berenger PML ................ fdtdTM.c:197-242 (coefficients), 290-327 (update)
 *   id 1  Yee TE + Berenger PML ................ fdtdTE.c:199-243, 283-321
 *   id 6  NS-FDTD TM ........................... nsFdtdTM.c:231-307, 68-151
 *   id 7  NS-FDTD TE ........................... nsFdtdTE.c:100-181, 233-308
 *   CW scattered-field sources ................. field.c:155-196
 * Permittivity maps are inputs (three per solver, in the reference's order), pinned bit-exactly
 * against the reference by tests/test_split_cpu.py.  Layout is the reference's: k = i*N_PY + j,
 * dense coefficient arrays, one full-grid pass per sub-step.
 */
#define _USE_MATH_DEFINES
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.1415926535897932384626433832795
#endif

typedef double complex cplx;

#define C0 0.7071
static const double EPS0 = 1.0;
static const double MU0 = 1.0 / C0 / C0;

enum { KIND_TM = 0, KIND_TE = 1, KIND_NS_TM = 6, KIND_NS_TE = 7 };

typedef struct SplitOracle {
  int kind, npx, npy, npml, nx, ny, ncell;
  double k_s, omega_s, angle_deg, time, ray_coef;
  /* TM kinds: Ez, Ezx, Ezy, Hx, Hy      TE kinds: Hz, Hzx, Hzy, Ex, Ey */
  cplx *f[5];
  /* TM kinds: C_EZX, C_EZXLX, C_EZY, C_EZYLY, C_HX, C_HXLY, C_HY, C_HYLX
   * TE kinds: C_EX, C_EXLY, C_EY, C_EYLX, C_HZX, C_HZXLX, C_HZY, C_HZYLY */
  double *c[8];
  /* TM kinds: EPS_EZ, EPS_HX, EPS_HY     TE kinds: EPS_EX, EPS_EY, EPS_HZ */
  double *eps[3];
} SplitOracle;

static double profile(double u, int npml, int ninner, int ntotal)       /* field.c:259-283 */
{
  if (u < npml) return pow(1.0 * (npml - u) / npml, 2);
  if (u >= npml && u < ninner + npml) return 0;
  return pow(1.0 * (u - (ntotal - npml - 1)) / npml, 2);
}
static double sig_x(const SplitOracle *s, double x) { return profile(x, s->npml, s->nx, s->npx); }
static double sig_y(const SplitOracle *s, double y) { return profile(y, s->npml, s->ny, s->npy); }
static double pml_coef(double ep_mu, double sig) { return (1.0 - sig / ep_mu) / (1.0 + sig / ep_mu); }   /* field.c:288 */
static double pml_coef_lxy(double ep_mu, double sig) { return 1.0 / (ep_mu + sig); }                     /* field.c:292 */
static double ns_beta(double alpha) { return tanh(alpha) / (1 + pow(tanh(alpha), 2)); }                  /* nsFdtdTM.c:221 */
static double ns_coef1(double beta) { return (1 - beta) / (1 + beta); }                                  /* nsFdtdTM.c:226 */

/* ---- coefficient loops -------------------------------------------------------------- */
static void coef_tm(SplitOracle *s)                                     /* fdtdTM.c:204-241 */
{
  const double sig_max = -(2.0 + 1.0) * EPS0 * C0 / 2.0 / s->npml * log(1.0e-8);
  for (int i = 0; i < s->npx; i++)
    for (int j = 0; j < s->npy; j++) {
      const int k = i * s->npy + j;
      const double eps_ez = s->eps[0][k];
      const double sig_ez_x = sig_max * sig_x(s, i), sig_ez_y = sig_max * sig_y(s, j);
      const double sig_hx_yy = MU0 / EPS0 * (sig_max * sig_y(s, j + 0.5));
      const double sig_hy_xx = MU0 / EPS0 * (sig_max * sig_x(s, i + 0.5));
      s->c[0][k] = pml_coef(eps_ez, sig_ez_x);   s->c[1][k] = pml_coef_lxy(eps_ez, sig_ez_x);
      s->c[2][k] = pml_coef(eps_ez, sig_ez_y);   s->c[3][k] = pml_coef_lxy(eps_ez, sig_ez_y);
      s->c[4][k] = pml_coef(MU0, sig_hx_yy);     s->c[5][k] = pml_coef_lxy(MU0, sig_hx_yy);
      s->c[6][k] = pml_coef(MU0, sig_hy_xx);     s->c[7][k] = pml_coef_lxy(MU0, sig_hy_xx);
    }
}

static void coef_te(SplitOracle *s)                                     /* fdtdTE.c:203-240 */
{
  const double sig_max = -(2.0 + 1.0) * EPS0 * C0 / 2.0 / s->npml * log(1.0e-8);
  for (int i = 0; i < s->npx; i++)
    for (int j = 0; j < s->npy; j++) {
      const int k = i * s->npy + j;
      const double eps_ex = s->eps[0][k], eps_ey = s->eps[1][k];
      const double sig_ex_y = sig_max * sig_y(s, j), sig_ey_x = sig_max * sig_x(s, i);
      const double sig_hz_xx = MU0 / EPS0 * (sig_max * sig_x(s, i + 0.5));
      const double sig_hz_yy = MU0 / EPS0 * (sig_max * sig_y(s, j + 0.5));
      s->c[0][k] = pml_coef(eps_ex, sig_ex_y);   s->c[1][k] = pml_coef_lxy(eps_ex, sig_ex_y);
      s->c[2][k] = pml_coef(eps_ey, sig_ey_x);   s->c[3][k] = pml_coef_lxy(eps_ey, sig_ey_x);
      s->c[4][k] = pml_coef(MU0, sig_hz_xx);     s->c[5][k] = pml_coef_lxy(MU0, sig_hz_xx);
      s->c[6][k] = pml_coef(MU0, sig_hz_yy);     s->c[7][k] = pml_coef_lxy(MU0, sig_hz_yy);
    }
}

static double ns_u(const SplitOracle *s, double eps) { return sin(s->omega_s * 0.5) / sin(s->k_s * sqrt(eps / EPS0) * 0.5); }

static void coef_ns_tm(SplitOracle *s)                                  /* nsFdtdTM.c:231-305 */
{
  const double sig_max = -(2.0 + 1.0) * EPS0 * C0 / s->npml * log(1.0e-8);
  for (int i = 0; i < s->npx; i++)
    for (int j = 0; j < s->npy; j++) {
      const int k = i * s->npy + j;
      const double b_hx_y = ns_beta(sig_max * sig_y(s, j + 0.5) / (2 * EPS0));
      const double b_hy_x = ns_beta(sig_max * sig_x(s, i + 0.5) / (2 * EPS0));
      const double b_ez_x = ns_beta(sig_max * sig_x(s, i) / (2 * EPS0));
      const double b_ez_y = ns_beta(sig_max * sig_y(s, j) / (2 * EPS0));
      const double z_ez = sqrt(MU0 / s->eps[0][k]), u_ez = ns_u(s, s->eps[0][k]);
      s->c[0][k] = ns_coef1(b_ez_x);
      s->c[1][k] = u_ez * z_ez / (1 + b_ez_x);
      s->c[2][k] = ns_coef1(b_ez_y);
      s->c[3][k] = u_ez * z_ez / (1.0 + b_ez_x);                         /* b_ez_x, as nsFdtdTM.c:287 has it */
      const double z_hx = sqrt(MU0 / s->eps[1][k]), u_hx = ns_u(s, s->eps[1][k]);
      s->c[4][k] = ns_coef1(b_hx_y);
      s->c[5][k] = u_hx / z_hx / (1.0 + b_hx_y);
      const double z_hy = sqrt(MU0 / s->eps[2][k]), u_hy = ns_u(s, s->eps[2][k]);
      s->c[6][k] = ns_coef1(b_hy_x);
      s->c[7][k] = u_hy / z_hy / (1.0 + b_hy_x);
    }
}

static void coef_ns_te(SplitOracle *s)                                  /* nsFdtdTE.c:116-180: interior cells only */
{
  const double sig_max = -(2.0 + 1.0) * EPS0 * C0 / s->npml * log(1.0e-8);
  for (int i = 1; i < s->npx - 1; i++)
    for (int j = 1; j < s->npy - 1; j++) {
      const int k = i * s->npy + j;
      const double eps_ex = s->eps[0][k], eps_ey = s->eps[1][k], eps_hz = s->eps[2][k];
      const double b_ex_y = ns_beta(sig_max * sig_y(s, j) / (2 * eps_ex));
      const double b_ey_x = ns_beta(sig_max * sig_x(s, i) / (2 * eps_ey));
      const double b_hz_x = ns_beta(sig_max * sig_x(s, i + 0.5) / (2 * eps_hz));
      const double b_hz_y = ns_beta(sig_max * sig_y(s, j + 0.5) / (2 * eps_hz));
      const double z_hz = sqrt(MU0 / eps_hz), u_hz = ns_u(s, eps_hz);
      s->c[4][k] = ns_coef1(b_hz_x);   s->c[5][k] = u_hz / z_hz / (1.0 + b_hz_x);
      s->c[6][k] = ns_coef1(b_hz_y);   s->c[7][k] = u_hz / z_hz / (1.0 + b_hz_y);
      const double z_ex = sqrt(MU0 / eps_ex), u_ex = ns_u(s, eps_ex);
      s->c[0][k] = ns_coef1(b_ex_y);   s->c[1][k] = u_ex * z_ex / (1.0 + b_ex_y);
      const double z_ey = sqrt(MU0 / eps_ey), u_ey = ns_u(s, eps_ey);
      s->c[2][k] = ns_coef1(b_ey_x);   s->c[3][k] = u_ey * z_ey / (1.0 + b_ey_x);
    }
}

/* ---- lifetime -------------------------------------------------------------------------- */
SplitOracle *split_oracle_create(int kind, int npx, int npy, int npml, double lambda_cells, double angle_deg,
                                 const double *eps0, const double *eps1, const double *eps2)
{
  SplitOracle *s = (SplitOracle *)calloc(1, sizeof *s);
  s->kind = kind; s->npx = npx; s->npy = npy; s->npml = npml;
  s->nx = npx - 2 * npml; s->ny = npy - 2 * npml; s->ncell = npx * npy;
  s->k_s = 2 * M_PI / lambda_cells;                                      /* field.c:103-106 */
  s->omega_s = C0 * s->k_s;
  s->angle_deg = angle_deg;
  const double *maps[3] = { eps0, eps1, eps2 };
  for (int m = 0; m < 3; m++) {
    s->eps[m] = (double *)malloc(sizeof(double) * (size_t)s->ncell);
    memcpy(s->eps[m], maps[m], sizeof(double) * (size_t)s->ncell);
  }
  for (int m = 0; m < 5; m++) s->f[m] = (cplx *)calloc((size_t)s->ncell, sizeof(cplx));
  for (int m = 0; m < 8; m++) s->c[m] = (double *)calloc((size_t)s->ncell, sizeof(double));
  switch (kind) {
  case KIND_TM: coef_tm(s); break;
  case KIND_TE: coef_te(s); break;
  case KIND_NS_TM: coef_ns_tm(s); break;
  default: coef_ns_te(s); break;
  }
  return s;
}

void split_oracle_destroy(SplitOracle *s)
{
  if (!s) return;
  for (int m = 0; m < 3; m++) free(s->eps[m]);
  for (int m = 0; m < 5; m++) free(s->f[m]);
  for (int m = 0; m < 8; m++) free(s->c[m]);
  free(s);
}

cplx *split_oracle_field(SplitOracle *s, int slot) { return s->f[slot]; }
double *split_oracle_coef(SplitOracle *s, int slot) { return s->c[slot]; }

/* ---- sources (field.c:155-196) ---------------------------------------------------------- */
static void scattered_wave_not_upml(SplitOracle *s, cplx *p, const double *eps, double gap_x, double gap_y)
{
  const double rad = s->angle_deg * M_PI / 180.0;
  const double ks_cos = cos(rad) * s->k_s, ks_sin = sin(rad) * s->k_s;
  for (int i = 1; i < s->npx - 1; i++)
    for (int j = 1; j < s->npy - 1; j++) {
      const int k = i * s->npy + j;
      const double kr = (i + gap_x) * ks_cos + (j + gap_y) * ks_sin;
      p[k] += s->ray_coef * (EPS0 / eps[k] - 1.0) *
              (cexp(I * (kr - s->omega_s * (s->time + 0.5))) - cexp(I * (kr - s->omega_s * (s->time - 0.5))));
    }
}

static void ns_scattered_wave_not_upml(SplitOracle *s, cplx *p, const double *eps, double gap_x, double gap_y, double dot)
{
  const double ray = s->ray_coef * dot;
  const double rad = s->angle_deg * M_PI / 180.0;
  const double ks_cos = cos(rad) * s->k_s, ks_sin = sin(rad) * s->k_s;
  for (int i = 1; i < s->npx - 1; i++)
    for (int j = 1; j < s->npy - 1; j++) {
      const int k = i * s->npy + j;
      const double kr = (i + gap_x) * ks_cos + (j + gap_y) * ks_sin;
      const double n = sqrt(eps[k] / EPS0);
      const double u0 = sin(s->omega_s * 0.5) / sin(s->k_s * 0.5);
      const double u1 = sin(s->omega_s * 0.5) / sin(n * s->k_s * 0.5);
      const double _n = u0 / u1;
      p[k] += ray * (1.0 / (_n * n) - 1.0) *
              (cexp(I * (kr - s->omega_s * (s->time + 1.0))) - cexp(I * (kr - s->omega_s * (s->time))));
    }
}

/* ---- update loops ------------------------------------------------------------------------ */
#define FOR_INTERIOR(s, i, j) for (int i = 1; i < (s)->npx - 1; i++) for (int j = 1; j < (s)->npy - 1; j++)

static void step_tm(SplitOracle *s)                                     /* fdtdTM.c:290-327 */
{
  cplx *Ez = s->f[0], *Ezx = s->f[1], *Ezy = s->f[2], *Hx = s->f[3], *Hy = s->f[4];
  const int P = s->npy;
  FOR_INTERIOR(s, i, j) { const int k = i * P + j;
    Hx[k] = s->c[4][k] * Hx[k] - s->c[5][k] * (Ezx[k + 1] - Ezx[k] + Ezy[k + 1] - Ezy[k]); }
  FOR_INTERIOR(s, i, j) { const int k = i * P + j;
    Hy[k] = s->c[6][k] * Hy[k] + s->c[7][k] * (Ezx[k + P] - Ezx[k] + Ezy[k + P] - Ezy[k]); }
  FOR_INTERIOR(s, i, j) { const int k = i * P + j;
    Ezx[k] = s->c[0][k] * Ezx[k] + s->c[1][k] * (Hy[k] - Hy[k - P]); }
  FOR_INTERIOR(s, i, j) { const int k = i * P + j;
    Ezy[k] = s->c[2][k] * Ezy[k] - s->c[3][k] * (Hx[k] - Hx[k - 1]); }
  FOR_INTERIOR(s, i, j) { const int k = i * P + j; Ez[k] = Ezx[k] + Ezy[k]; }
  scattered_wave_not_upml(s, Ezx, s->eps[0], 0.0, 0.0);
}

static void step_te(SplitOracle *s)                                     /* fdtdTE.c:283-321 */
{
  cplx *Hz = s->f[0], *Hzx = s->f[1], *Hzy = s->f[2], *Ex = s->f[3], *Ey = s->f[4];
  const int P = s->npy;
  FOR_INTERIOR(s, i, j) { const int k = i * P + j;
    Ex[k] = s->c[0][k] * Ex[k] + s->c[1][k] * (Hzx[k] - Hzx[k - 1] + Hzy[k] - Hzy[k - 1]); }
  FOR_INTERIOR(s, i, j) { const int k = i * P + j;
    Ey[k] = s->c[2][k] * Ey[k] - s->c[3][k] * (Hzx[k] - Hzx[k - P] + Hzy[k] - Hzy[k - P]); }
  scattered_wave_not_upml(s, Ey, s->eps[1], 0.0, 0.5);
  FOR_INTERIOR(s, i, j) { const int k = i * P + j;
    Hzx[k] = s->c[4][k] * Hzx[k] - s->c[5][k] * (Ey[k + P] - Ey[k]); }
  FOR_INTERIOR(s, i, j) { const int k = i * P + j;
    Hzy[k] = s->c[6][k] * Hzy[k] + s->c[7][k] * (Ex[k + 1] - Ex[k]); }
  FOR_INTERIOR(s, i, j) { const int k = i * P + j; Hz[k] = Hzx[k] + Hzy[k]; }
}

static double ns_r2(const SplitOracle *s)                                /* nsFdtdTM.c:115-117 */
{
  const double r = 1.0 / 6.0 + s->k_s * s->k_s / 180.0 - pow(s->k_s, 4) / 23040;
  return r / 2.0;
}

static void step_ns_tm(SplitOracle *s)                                  /* nsFdtdTM.c:68-151 */
{
  cplx *Ez = s->f[0], *Ezx = s->f[1], *Ezy = s->f[2], *Hx = s->f[3], *Hy = s->f[4];
  const int dx = s->npy, dy = 1;
  const double r_2 = ns_r2(s);
  FOR_INTERIOR(s, i, j) { const int k = i * dx + j;
    cplx ns = r_2 * ((Ez[k + dy + dx] + Ez[k + dy - dx] - 2 * Ez[k + dy]) - (Ez[k + dx] + Ez[k - dx] - 2 * Ez[k]));
    Hx[k] = s->c[4][k] * Hx[k] - s->c[5][k] * (Ez[k + dy] - Ez[k] + ns); }
  FOR_INTERIOR(s, i, j) { const int k = i * dx + j;
    cplx ns = r_2 * ((Ez[k + dx + dy] + Ez[k + dx - dy] - 2 * Ez[k + dx]) - (Ez[k + dy] + Ez[k - dy] - 2 * Ez[k]));
    Hy[k] = s->c[6][k] * Hy[k] + s->c[7][k] * (Ez[k + dx] - Ez[k] + ns); }
  FOR_INTERIOR(s, i, j) { const int k = i * dx + j;
    Ezx[k] = s->c[0][k] * Ezx[k] + s->c[1][k] * (Hy[k] - Hy[k - dx]); }
  FOR_INTERIOR(s, i, j) { const int k = i * dx + j;
    Ezy[k] = s->c[2][k] * Ezy[k] - s->c[3][k] * (Hx[k] - Hx[k - dy]); }
  ns_scattered_wave_not_upml(s, Ezy, s->eps[0], 0, 0, 1.0);
  FOR_INTERIOR(s, i, j) { const int k = i * dx + j; Ez[k] = Ezx[k] + Ezy[k]; }
}

static void step_ns_te(SplitOracle *s)                                  /* nsFdtdTE.c:233-308 */
{
  cplx *Hz = s->f[0], *Hzx = s->f[1], *Hzy = s->f[2], *Ex = s->f[3], *Ey = s->f[4];
  const int dx = s->npy, dy = 1;
  const double r_2 = ns_r2(s);
  FOR_INTERIOR(s, i, j) { const int k = i * dx + j;
    Hzx[k] = s->c[4][k] * Hzx[k] - s->c[5][k] * (Ey[k + dx] - Ey[k]); }
  FOR_INTERIOR(s, i, j) { const int k = i * dx + j;
    Hzy[k] = s->c[6][k] * Hzy[k] + s->c[7][k] * (Ex[k + dy] - Ex[k]); }
  FOR_INTERIOR(s, i, j) { const int k = i * dx + j; Hz[k] = Hzx[k] + Hzy[k]; }
  FOR_INTERIOR(s, i, j) { const int k = i * dx + j;
    cplx ns = r_2 * ((Hz[k + dx] + Hz[k - dx] - 2 * Hz[k]) - (Hz[k - dy + dx] + Hz[k - dy - dx] - 2 * Hz[k - dy]));
    Ex[k] = s->c[0][k] * Ex[k] + s->c[1][k] * (Hz[k] - Hz[k - dy] + ns); }
  FOR_INTERIOR(s, i, j) { const int k = i * dx + j;
    cplx ns = r_2 * ((Hz[k + dy] + Hz[k - dy] - 2 * Hz[k]) - (Hz[k - dx + dy] + Hz[k - dx - dy] - 2 * Hz[k - dx]));
    Ey[k] = s->c[2][k] * Ey[k] - s->c[3][k] * (Hz[k] - Hz[k - dx] + ns); }
  const double co = cos((s->angle_deg + 90) * M_PI / 180.0), si = sin((s->angle_deg + 90) * M_PI / 180.0);
  if (co != 0.0) ns_scattered_wave_not_upml(s, Ex, s->eps[0], 0, 0.5, co);   /* gap (0, 0.5) on both, nsFdtdTE.c:247-250 */
  if (si != 0.0) ns_scattered_wave_not_upml(s, Ey, s->eps[1], 0, 0.5, si);
}

void split_oracle_step(SplitOracle *s, int n)
{
  for (int t = 0; t < n; t++) {
    switch (s->kind) {
    case KIND_TM: step_tm(s); break;
    case KIND_TE: step_te(s); break;
    case KIND_NS_TM: step_ns_tm(s); break;
    default: step_ns_te(s); break;
    }
    s->time += 1.0;                                                      /* field.c:312-315 */
    s->ray_coef = 1.0 - exp(-pow(0.01 * s->time, 2));
  }
}
