/* fdtd_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the hot path of rennone/mpiFDTD's serial UPML
 * solvers, used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg as the CHECKER for the CUDA path.  Nothing under mpifdtd_b200/ links,
 * imports or calls it; the product has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks this file against
 * (a) the golden vectors under tests/golden/ generated from the unmodified
 * reference by tests/golden/make_golden.py (serial UPML runs, and the MPI-variant
 * ids incl. the plane-wave source), and (b) oracle/_ref/libref.so (the
 * reference compiled in place) whenever that library is present.
 *
 * What is restated, with the reference lines each block follows:
 *   grid / wave / NTFF box ........ field.c:89-143
 *   PML profile ................... field.c:259-283
 *   TM coefficients ............... fdtdTM_upml.c:224-274
 *   TE coefficients ............... fdtdTE_upml.c:361-412
 *   TM update loops ............... fdtdTM_upml.c:54-66,155-219
 *   TE update loops ............... fdtdTE_upml.c:168-192,252-314
 *   Gaussian-pulse source ......... field.c:224-256
 *   soft-start clock .............. field.c:312-315
 *   opt-in sources ................ field.c:145-152,202-218; mpiTM_UPML.c:377-403
 *   MPI-variant solvers (ids 4/5) . mpiTM_UPML.c:196-217,337-374,430-520; mpiTE_UPML.c:250-400
 *                                   (stepping at one rank) and their own ntff():
 *                                   mpiTM_UPML.c:840-1037, mpiTE_UPML.c:601-790
 *   NTFF time-domain accumulation . ntffTM.c:279-371, ntffTE.c:57-157
 *   translate / FFT / spectrum .... ntffTM.c:161-232, ntffTE.c:20-55,160-195,
 *                                   cfft.c:104-179
 * Permittivity maps are inputs (the product's host C builds them; they are
 * pinned bit-exactly against the reference by tests/test_materials_cpu.py).
 *
 * Layout is the reference's: k = i*N_PY + j, dense coefficient arrays, one pass
 * per sub-step -- deliberately NOT the GPU engine's organisation, so the two
 * implementations share nothing but the equations.
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
static const double Z0 = 1.41422712488;

enum { KIND_TM = 2, KIND_TE = 3 };
enum { N_ANG = 360, N_FFT = 8192, LAM_FIRST = 380, LAM_LAST = 700 };
#define UW_GUARD 16   /* elements of slack before/after the U/W block: the reference
                         indexes m-1 = -2 at step 0 and m+1 = arraySize at the last
                         step without a bound check (ntffTM.c:285-287) */

typedef struct OracleSim {
  int kind, npx, npy, npml, nx, ny, ncell;
  int h_u_nm, steps, point_source;
  int source_form;              /* 0 the solver's own source; 1 CW instead (id 2); 2 plane-wave line source added (ids 2, 4) */
  double lambda_s, k_s, omega_s, angle_deg;
  double time, ray_coef;
  /* NTFF box */
  int top, bottom, left, right, cx, cy, array_size;
  double rf_per_c;
  /* fields: TM Ez,Jz,Dz,Hx,Mx,Bx,Hy,My,By | TE Ex,Jx,Dx,Ey,Jy,Dy,Hz,Mz,Bz */
  cplx *f[9];
  double *c[15];
  double *eps[2];
  cplx *uw_block, *uw[3];          /* TM Ux,Uy,Wz | TE Wx,Wy,Uz ; [360][array_size] each */
} OracleSim;

/* ---- field.c:259-283 ------------------------------------------------------ */
static double profile(double u, int npml, int ninner, int ntotal)
{
  if (u < npml) return pow(1.0 * (npml - u) / npml, 2);
  if (u >= npml && u < ninner + npml) return 0;
  return pow(1.0 * (u - (ntotal - npml - 1)) / npml, 2);
}
static double sig_x(const OracleSim *s, double x) { return profile(x, s->npml, s->nx, s->npx); }
static double sig_y(const OracleSim *s, double y) { return profile(y, s->npml, s->ny, s->npy); }

static void *zalloc(size_t n, size_t sz) { return calloc(n ? n : 1, sz); }

/* TM slots */
enum { EZ, JZ, DZ, HX, MX, BX, HY, MY, BY };
enum { C_JZ, C_JZHXHY, C_DZ, C_DZJZ1, C_DZJZ0, C_MX, C_MXEZ, C_BX, C_BXMX1, C_BXMX0,
       C_MY, C_MYEZ, C_BY, C_BYMY1, C_BYMY0 };
/* TE slots */
enum { EX, JX, DX, EY, JY, DY, HZ, MZ, BZ };
enum { C_JX, C_JXHZ, C_DX, C_DXJX1, C_DXJX0, C_JY, C_JYHZ, C_DY, C_DYJY1, C_DYJY0,
       C_MZ, C_MZEXEY, C_BZ, C_BZMZ1, C_BZMZ0 };

static void set_coefficients(OracleSim *s)
{
  const double R = 1.0e-8, M = 2.0, e = EPS0, sz = 0;
  if (s->kind == KIND_TM) {           /* fdtdTM_upml.c:230-271 */
    const double smax = -(M + 1.0) * EPS0 * C0 / 2.0 / s->npml / cos(M_PI / 3) * log(R);
    for (int i = 0; i < s->npx; i++)
      for (int j = 0; j < s->npy; j++) {
        int k = i * s->npy + j;
        double ezx = smax * sig_x(s, i), ezy = smax * sig_y(s, j);
        double hxx = smax * sig_x(s, i), hxy = smax * sig_y(s, j + 0.5);
        double hyx = smax * sig_x(s, i + 0.5), hyy = smax * sig_y(s, j);
        s->c[C_JZ][k] = (2 * e - ezx) / (2 * e + ezx);
        s->c[C_JZHXHY][k] = (2 * e) / (2 * e + ezx);
        s->c[C_DZ][k] = (2 * e - ezy) / (2 * e + ezy);
        s->c[C_DZJZ1][k] = (2 * e + sz) / (2 * e + ezy);
        s->c[C_DZJZ0][k] = (2 * e - sz) / (2 * e + ezy);
        s->c[C_MX][k] = (2 * e - hxy) / (2 * e + hxy);
        s->c[C_MXEZ][k] = (2 * e) / (2 * e + hxy);
        s->c[C_BX][k] = (2 * e - sz) / (2 * e + sz);
        s->c[C_BXMX1][k] = (2 * e + hxx) / (2 * e + sz);
        s->c[C_BXMX0][k] = (2 * e - hxx) / (2 * e + sz);
        s->c[C_MY][k] = (2 * e - sz) / (2 * e + sz);
        s->c[C_MYEZ][k] = (2 * e) / (2 * e + sz);
        s->c[C_BY][k] = (2 * e - hyx) / (2 * e + hyx);
        s->c[C_BYMY1][k] = (2 * e + hyy) / (2 * e + hyx);
        s->c[C_BYMY0][k] = (2 * e - hyy) / (2 * e + hyx);
      }
  } else {                            /* fdtdTE_upml.c:367-409 */
    const double smax = -(M + 1.0) * EPS0 * C0 / 2.0 / s->npml * log(R);
    for (int i = 0; i < s->npx; i++)
      for (int j = 0; j < s->npy; j++) {
        int k = i * s->npy + j;
        double exx = smax * sig_x(s, i + 0.5), exy = smax * sig_y(s, j);
        double eyx = smax * sig_x(s, i), eyy = smax * sig_y(s, j + 0.5);
        double hzx = smax * sig_x(s, i + 0.5), hzy = smax * sig_y(s, j + 0.5);
        s->c[C_JX][k] = (2 * e - exy) / (2 * e + exy);
        s->c[C_JXHZ][k] = (2 * e) / (2 * e + exy);
        s->c[C_DX][k] = (2 * e - sz) / (2 * e + sz);
        s->c[C_DXJX1][k] = (2 * e + exx) / (2 * e + sz);
        s->c[C_DXJX0][k] = (2 * e - exx) / (2 * e + sz);
        s->c[C_JY][k] = (2 * e - sz) / (2 * e + sz);
        s->c[C_JYHZ][k] = (2 * e) / (2 * e + sz);
        s->c[C_DY][k] = (2 * e - eyx) / (2 * e + eyx);
        s->c[C_DYJY1][k] = (2 * e + eyy) / (2 * e + eyx);
        s->c[C_DYJY0][k] = (2 * e - eyy) / (2 * e + eyx);
        s->c[C_MZ][k] = (2 * e - hzx) / (2 * e + hzx);
        s->c[C_MZEXEY][k] = (2 * e) / (2 * e + hzx);
        s->c[C_BZ][k] = (2 * e - hzy) / (2 * e + hzy);
        s->c[C_BZMZ1][k] = (2 * e + sz) / (2 * e + hzy);
        s->c[C_BZMZ0][k] = (2 * e - sz) / (2 * e + hzy);
      }
  }
}

/* eps0/eps1: TM EPS_EZ | TE EPS_EX, EPS_EY; arrays of npx*npy doubles (copied) */
OracleSim *oracle_create(int kind, int width_nm, int height_nm, int h_u_nm, int npml, int lambda_nm,
                         int angle_deg, int steps, const double *eps0, const double *eps1)
{
  OracleSim *s = (OracleSim *)zalloc(1, sizeof *s);
  s->kind = kind;
  s->h_u_nm = h_u_nm;
  s->npx = (double)width_nm / h_u_nm;          /* field.c:95-96, truncation */
  s->npy = (double)height_nm / h_u_nm;
  s->npml = npml;
  s->nx = s->npx - 2 * npml;
  s->ny = s->npy - 2 * npml;
  s->ncell = s->npx * s->npy;
  s->lambda_s = (double)lambda_nm / h_u_nm;    /* field.c:105-110 */
  s->k_s = 2 * M_PI / s->lambda_s;
  s->omega_s = C0 * s->k_s;
  s->angle_deg = angle_deg;
  s->steps = steps;
  s->cx = s->npx / 2;  s->cy = s->npy / 2;     /* field.c:133-142 */
  s->top = s->npy - npml - 5;  s->bottom = npml + 5;
  s->left = npml + 5;          s->right = s->npx - npml - 5;
  double len = (s->top - s->bottom) / 2;
  s->rf_per_c = len * 2;
  s->array_size = (double)steps + 2 * s->rf_per_c;
  for (int n = 0; n < 9; n++) s->f[n] = (cplx *)zalloc(s->ncell, sizeof(cplx));
  for (int n = 0; n < 15; n++) s->c[n] = (double *)zalloc(s->ncell, sizeof(double));
  s->eps[0] = (double *)zalloc(s->ncell, sizeof(double));
  s->eps[1] = (double *)zalloc(s->ncell, sizeof(double));
  memcpy(s->eps[0], eps0, sizeof(double) * s->ncell);
  if (eps1) memcpy(s->eps[1], eps1, sizeof(double) * s->ncell);
  set_coefficients(s);
  size_t per = (size_t)N_ANG * s->array_size;
  s->uw_block = (cplx *)zalloc(3 * (per + 2 * UW_GUARD), sizeof(cplx));
  for (int n = 0; n < 3; n++) s->uw[n] = s->uw_block + UW_GUARD + n * (per + 2 * UW_GUARD);
  return s;
}

void oracle_destroy(OracleSim *s)
{
  if (!s) return;
  for (int n = 0; n < 9; n++) free(s->f[n]);
  for (int n = 0; n < 15; n++) free(s->c[n]);
  free(s->eps[0]); free(s->eps[1]); free(s->uw_block);
  free(s);
}

void oracle_set_point_source(OracleSim *s, int on) { s->point_source = on; }
void oracle_set_source_form(OracleSim *s, int form) { s->source_form = form; }
void oracle_set_angle(OracleSim *s, int deg) { s->angle_deg = deg; }

/* ---- field.c:224-256 ------------------------------------------------------- */
static void pulse(OracleSim *s, cplx *p, const double *eps, double gx, double gy, double dot)
{
  double rad = s->angle_deg * M_PI / 180.0;
  double cpc = cos(rad) / C0, spc = sin(rad) / C0;
  const double bw = 50;
  const double peak = (s->npx / 2.0 + gx) * cpc + (s->npy / 2 + gy) * spc;
  const double t0 = -peak + 500;
  for (int i = 1; i < s->npx - 1; i++)
    for (int j = 1; j < s->npy - 1; j++) {
      int k = i * s->npy + j;
      if (EPS0 == eps[k]) continue;
      const double r = (i + gx) * cpc + (j + gy) * spc - (s->time - t0);
      const double g = exp(-pow(r / bw, 2));
      p[k] += dot * g * (EPS0 / eps[k] - 1) * cexp(I * r * s->omega_s);
    }
}

/* ---- ntffTM.c:279-288 -------------------------------------------------------- */
static inline void bin3(double t_plus_shift, cplx v, cplx *row)
{
  int m = floor(t_plus_shift + 0.5);
  double a = (0.5 + t_plus_shift - m);
  double b = 1.0 - a;
  double ab = a - b;
  row[m - 1] += v * b;
  row[m] += v * ab;
  row[m + 1] -= v * a;
}

/* ntffTM.c:293-371 and ntffTE.c:68-157; `stag` is TE's half-cell stagger */
static void ntff_accumulate(OracleSim *s)
{
  const int N = s->npy;
  double tE = s->time - 1, tH = s->time - 0.5;
  double lt = s->left - s->cx, rt = s->right - s->cx, bm = s->bottom - s->cy, tp = s->top - s->cy;
  int lb = s->left * N + s->bottom, ltk = s->left * N + s->top;
  int rb = s->right * N + s->bottom, rtk = s->right * N + s->top;
  const double to_rad = M_PI / 180.0;
  const int tm = (s->kind == KIND_TM);
  const double stag = tm ? 0.0 : 0.5;
  for (int ang = 0; ang < N_ANG; ang++) {
    double rad = ang * to_rad;
    double r1x = cos(rad) / C0, r1y = sin(rad) / C0;
    cplx *A0 = s->uw[0] + (size_t)ang * s->array_size;
    cplx *A1 = s->uw[1] + (size_t)ang * s->array_size;
    cplx *A2 = s->uw[2] + (size_t)ang * s->array_size;
    double sh;
    if (tm) {
      const cplx *Ez = s->f[EZ], *Hx = s->f[HX], *Hy = s->f[HY];
      sh = -(r1x * lt + r1y * bm) + s->rf_per_c;                      /* bottom */
      for (int k = lb; k < rb; k += N) { bin3(tE + sh, Ez[k], A0); bin3(tH + sh, 0.5 * (Hx[k] + Hx[k - 1]), A2); sh -= r1x; }
      sh = -(r1x * rt + r1y * bm) + s->rf_per_c;                      /* right  */
      for (int k = rb; k < rtk; k++) { bin3(tE + sh, Ez[k], A1); bin3(tH + sh, 0.5 * (Hy[k] + Hy[k - N]), A2); sh -= r1y; }
      sh = -(r1x * lt + r1y * tp) + s->rf_per_c;                      /* top    */
      for (int k = ltk; k < rtk; k += N) { bin3(tE + sh, -Ez[k], A0); bin3(tH + sh, -0.5 * (Hx[k] + Hx[k - 1]), A2); sh -= r1x; }
      sh = -(r1x * lt + r1y * bm) + s->rf_per_c;                      /* left   */
      for (int k = lb; k < ltk; k++) { bin3(tE + sh, -Ez[k], A1); bin3(tH + sh, -0.5 * (Hy[k] + Hy[k - N]), A2); sh -= r1y; }
    } else {
      const cplx *Ex = s->f[EX], *Ey = s->f[EY], *Hz = s->f[HZ];
      sh = -(r1x * (lt + stag) + r1y * bm) + s->rf_per_c;
      for (int k = lb; k < rb; k += N) { bin3(tE + sh, -Ex[k], A2); bin3(tH + sh, -0.5 * (Hz[k] + Hz[k - 1]), A0); sh -= r1x; }
      sh = -(r1x * rt + r1y * (bm + stag)) + s->rf_per_c;
      for (int k = rb; k < rtk; k++) { bin3(tE + sh, -Ey[k], A2); bin3(tH + sh, -0.5 * (Hz[k] + Hz[k - N]), A1); sh -= r1y; }
      sh = -(r1x * (lt + stag) + r1y * tp) + s->rf_per_c;
      for (int k = ltk; k < rtk; k += N) { bin3(tE + sh, Ex[k], A2); bin3(tH + sh, 0.5 * (Hz[k] + Hz[k - 1]), A0); sh -= r1x; }
      sh = -(r1x * lt + r1y * (bm + stag)) + s->rf_per_c;
      for (int k = lb; k < ltk; k++) { bin3(tE + sh, Ey[k], A2); bin3(tH + sh, 0.5 * (Hz[k] + Hz[k - N]), A1); sh -= r1y; }
    }
  }
}

/* ---- opt-in source forms (row a10), as oracle/refbuild/wrap_*.c compose them from the reference ----
 * field_scatteredWave (field.c:202-218; the commented alternative at fdtdTM_upml.c:62) */
static void cw_wave(OracleSim *s, cplx *p, const double *eps, double gx, double gy)
{
  double rad = s->angle_deg * M_PI / 180.0;
  double ks_cos = cos(rad) * s->k_s, ks_sin = sin(rad) * s->k_s;
  for (int i = 1; i < s->npx - 1; i++)
    for (int j = 1; j < s->npy - 1; j++) {
      int k = i * s->npy + j;
      double kr = (i + gx) * ks_cos + (j + gy) * ks_sin;
      p[k] += s->ray_coef * (EPS0 / eps[k] - 1.0) * cexp(I * (kr - s->omega_s * s->time));
    }
}
/* planeWave (mpiTM_UPML.c:377-403; commented call at :204): a line source on grid row x, columns
 * y_lo..y_hi */
static void plane_line(OracleSim *s, cplx *p, int x, int y_lo, int y_hi)
{
  const double rad = s->angle_deg * M_PI / 180;
  const double ks_cos = cos(rad) * s->k_s, ks_sin = sin(rad) * s->k_s;
  for (int y = y_lo; y <= y_hi; y++) {
    double kr = (x * ks_cos + y * ks_sin) - s->time;
    p[x * s->npy + y] += s->ray_coef * cexp(I * kr * s->omega_s);
  }
}

static void step_tm(OracleSim *s, int with_ntff)
{
  const int N = s->npy;
  cplx **f = s->f; double **c = s->c;
#define INTERIOR for (int i = 1; i < s->npx - 1; i++) for (int j = 1; j < s->npy - 1; j++)
  INTERIOR { int k = i * N + j; cplx o = f[MX][k];                    /* calcMB */
    f[MX][k] = c[C_MX][k] * f[MX][k] - c[C_MXEZ][k] * (f[EZ][k + 1] - f[EZ][k]);
    f[BX][k] = c[C_BX][k] * f[BX][k] + c[C_BXMX1][k] * f[MX][k] - c[C_BXMX0][k] * o; }
  INTERIOR { int k = i * N + j; cplx o = f[MY][k];
    f[MY][k] = c[C_MY][k] * f[MY][k] - c[C_MYEZ][k] * (-f[EZ][k + N] + f[EZ][k]);
    f[BY][k] = c[C_BY][k] * f[BY][k] + c[C_BYMY1][k] * f[MY][k] - c[C_BYMY0][k] * o; }
  INTERIOR { int k = i * N + j; f[HX][k] = f[BX][k] / MU0; }          /* calcH  */
  INTERIOR { int k = i * N + j; f[HY][k] = f[BY][k] / MU0; }
  INTERIOR { int k = i * N + j; cplx o = f[JZ][k];                    /* calcJD */
    f[JZ][k] = c[C_JZ][k] * f[JZ][k] + c[C_JZHXHY][k] * (+f[HY][k] - f[HY][k - N] - f[HX][k] + f[HX][k - 1]);
    f[DZ][k] = c[C_DZ][k] * f[DZ][k] + c[C_DZJZ1][k] * f[JZ][k] - c[C_DZJZ0][k] * o; }
  INTERIOR { int k = i * N + j; f[EZ][k] = f[DZ][k] / s->eps[0][k]; } /* calcE  */
  if (s->source_form == 1) cw_wave(s, f[EZ], s->eps[0], 0, 0);
  else pulse(s, f[EZ], s->eps[0], 0, 0, 1.0);
  if (s->source_form == 2) plane_line(s, f[EZ], s->left, 1, s->npy - 2);   /* serial grid: x = NTFF left edge */
  if (s->point_source)
    f[EZ][(s->npx / 2) * N + s->npy / 2] += s->ray_coef * cexp(I * s->omega_s * s->time);
  if (with_ntff) ntff_accumulate(s);
}

static void step_te(OracleSim *s, int with_ntff)
{
  const int N = s->npy;
  cplx **f = s->f; double **c = s->c;
  INTERIOR { int k = i * N + j; cplx o = f[MZ][k];                    /* calcMB */
    f[MZ][k] = c[C_MZ][k] * f[MZ][k] - c[C_MZEXEY][k] * (f[EY][k + N] - f[EY][k] - f[EX][k + 1] + f[EX][k]);
    f[BZ][k] = c[C_BZ][k] * f[BZ][k] + c[C_BZMZ1][k] * f[MZ][k] - c[C_BZMZ0][k] * o; }
  INTERIOR { int k = i * N + j; f[HZ][k] = f[BZ][k] / MU0; }          /* calcH  */
  INTERIOR { int k = i * N + j; cplx o = f[JX][k];                    /* calcJD */
    f[JX][k] = c[C_JX][k] * f[JX][k] + c[C_JXHZ][k] * (f[HZ][k] - f[HZ][k - 1]);
    f[DX][k] = c[C_DX][k] * f[DX][k] + c[C_DXJX1][k] * f[JX][k] - c[C_DXJX0][k] * o; }
  INTERIOR { int k = i * N + j; cplx o = f[JY][k];
    f[JY][k] = c[C_JY][k] * f[JY][k] + c[C_JYHZ][k] * (-f[HZ][k] + f[HZ][k - N]);
    f[DY][k] = c[C_DY][k] * f[DY][k] + c[C_DYJY1][k] * f[JY][k] - c[C_DYJY0][k] * o; }
  INTERIOR { int k = i * N + j; f[EX][k] = f[DX][k] / s->eps[0][k]; } /* calcE  */
  INTERIOR { int k = i * N + j; f[EY][k] = f[DY][k] / s->eps[1][k]; }
  double co = cos((s->angle_deg + 90) * M_PI / 180.0);                /* fdtdTE_upml.c:182-189 */
  double si = sin((s->angle_deg + 90) * M_PI / 180.0);
  if (co != 0.0) pulse(s, f[EX], s->eps[0], 0.5, 0.0, co);
  if (si != 0.0) pulse(s, f[EY], s->eps[1], 0.0, 0.5, si);
  if (s->point_source)
    f[EX][(s->npx / 2) * N + s->npy / 2] += s->ray_coef * cexp(I * s->omega_s * s->time);
  if (with_ntff) ntff_accumulate(s);
}

void oracle_step(OracleSim *s, int n, int with_ntff)
{
  for (int it = 0; it < n; it++) {
    if (s->kind == KIND_TM) step_tm(s, with_ntff); else step_te(s, with_ntff);
    s->time += 1.0;                                                   /* field.c:312-315 */
    s->ray_coef = 1.0 - exp(-pow(0.01 * s->time, 2));
  }
}

/* ---- the "MPI" solver variants (ids 4 / 5) at one rank ---------------------------------------
 * mpiTM_UPML.c:196-217 (update), 337-374 (scatteredWave), 430-520 (loops), 663-716 (coefficients);
 * mpiTE_UPML.c:250-304 (scatteredWave, update), 337-400 (loops), 489-536 (coefficients).
 * Differences from the serial solvers restated above: E phase first, then the source, then the
 * H phase; a continuous-wave source (TE: on Ey only); and the loops cover ALL npx x npy cells
 * (local 1..SUB_N_PX-2 <-> global 0..N-1) against a ring of ghost cells that stay zero at one
 * rank -- here: neighbours outside the grid read as 0.  Coefficients and permittivity positions
 * at global (x, y) are the serial solvers' at (i, j) (same sigma positions, same sig_max: with
 * cos(pi/3) for TM, without for TE), so set_coefficients() above serves both.  The reference keeps
 * (N+2) x (N+2) arrays; this restatement keeps N x N, i.e. the reference's arrays without the ring. */
static inline cplx cell(const OracleSim *s, const cplx *a, int i, int j)
{
  return (i < 0 || j < 0 || i >= s->npx || j >= s->npy) ? 0 : a[i * s->npy + j];
}

static void step_mpi_tm(OracleSim *s)
{
  const int N = s->npy;
  cplx **f = s->f; double **c = s->c;
#define ALL for (int i = 0; i < s->npx; i++) for (int j = 0; j < s->npy; j++)
  ALL { int k = i * N + j; cplx o = f[JZ][k];                         /* calcJD: left = i-1, bottom = j-1 */
    f[JZ][k] = c[C_JZ][k] * f[JZ][k] + c[C_JZHXHY][k] * (+f[HY][k] - cell(s, f[HY], i - 1, j) - f[HX][k] + cell(s, f[HX], i, j - 1));
    f[DZ][k] = c[C_DZ][k] * f[DZ][k] + c[C_DZJZ1][k] * f[JZ][k] - c[C_DZJZ0][k] * o; }
  ALL { int k = i * N + j; f[EZ][k] = f[DZ][k] / s->eps[0][k]; }      /* calcE */
  {                                                                  /* scatteredWave(Ez, EPS_EZ) */
    double rad = s->angle_deg * M_PI / 180;
    double ks_cos = cos(rad) * s->k_s, ks_sin = sin(rad) * s->k_s;
    ALL { int k = i * N + j;
      double kr = i * ks_cos + j * ks_sin;
      f[EZ][k] += s->ray_coef * (EPS0 / s->eps[0][k] - 1.0) * cexp(I * (kr - s->omega_s * s->time)); }
    if (s->source_form == 2)                                         /* planeWave(Ez, EPS_EZ): local i = left */
      plane_line(s, f[EZ], s->left - 1, 0, s->npy - 1);
  }
  ALL { int k = i * N + j; cplx o = f[MX][k];                         /* calcMB: top = j+1, right = i+1 */
    f[MX][k] = c[C_MX][k] * f[MX][k] - c[C_MXEZ][k] * (cell(s, f[EZ], i, j + 1) - f[EZ][k]);
    f[BX][k] = c[C_BX][k] * f[BX][k] + c[C_BXMX1][k] * f[MX][k] - c[C_BXMX0][k] * o; }
  ALL { int k = i * N + j; cplx o = f[MY][k];
    f[MY][k] = c[C_MY][k] * f[MY][k] - c[C_MYEZ][k] * (-cell(s, f[EZ], i + 1, j) + f[EZ][k]);
    f[BY][k] = c[C_BY][k] * f[BY][k] + c[C_BYMY1][k] * f[MY][k] - c[C_BYMY0][k] * o; }
  ALL { int k = i * N + j; f[HX][k] = f[BX][k] / MU0; }               /* calcH */
  ALL { int k = i * N + j; f[HY][k] = f[BY][k] / MU0; }
}

static void step_mpi_te(OracleSim *s)
{
  const int N = s->npy;
  cplx **f = s->f; double **c = s->c;
  ALL { int k = i * N + j; cplx o = f[JX][k];                         /* calcJD */
    f[JX][k] = c[C_JX][k] * f[JX][k] + c[C_JXHZ][k] * (f[HZ][k] - cell(s, f[HZ], i, j - 1));
    f[DX][k] = c[C_DX][k] * f[DX][k] + c[C_DXJX1][k] * f[JX][k] - c[C_DXJX0][k] * o; }
  ALL { int k = i * N + j; cplx o = f[JY][k];
    f[JY][k] = c[C_JY][k] * f[JY][k] + c[C_JYHZ][k] * (-f[HZ][k] + cell(s, f[HZ], i - 1, j));
    f[DY][k] = c[C_DY][k] * f[DY][k] + c[C_DYJY1][k] * f[JY][k] - c[C_DYJY0][k] * o; }
  ALL { int k = i * N + j; f[EX][k] = f[DX][k] / s->eps[0][k]; }      /* calcE */
  ALL { int k = i * N + j; f[EY][k] = f[DY][k] / s->eps[1][k]; }
  {                                                                  /* scatteredWave(Ey, EPS_EY) */
    double rad = s->angle_deg * M_PI / 180.0;
    double ks_cos = cos(rad) * s->k_s, ks_sin = sin(rad) * s->k_s;
    ALL { int k = i * N + j;
      double ikx = i * ks_cos + j * ks_sin;
      f[EY][k] += s->ray_coef * (EPS0 / s->eps[1][k] - 1) * (cos(ikx - s->omega_s * s->time) + I * sin(ikx - s->omega_s * s->time)); }
  }
  ALL { int k = i * N + j; cplx o = f[MZ][k];                         /* calcMB */
    f[MZ][k] = c[C_MZ][k] * f[MZ][k] - c[C_MZEXEY][k] * (cell(s, f[EY], i + 1, j) - f[EY][k] - cell(s, f[EX], i, j + 1) + f[EX][k]);
    f[BZ][k] = c[C_BZ][k] * f[BZ][k] + c[C_BZMZ1][k] * f[MZ][k] - c[C_BZMZ0][k] * o; }
  ALL { int k = i * N + j; f[HZ][k] = f[BZ][k] / MU0; }               /* calcH */
}

/* ---- the variants' own ntff(): mpiTM_UPML.c:840-1037, mpiTE_UPML.c:601-790 ------------------
 * Same three-tap binning as ntffTM_TimeCalc, but: the time shift is evaluated directly per cell (no
 * running subtraction), every tap is scaled by coef = 1/(4 pi C 1e6), and the NTFF box indices are
 * used as LOCAL indices of the (N+2)^2 arrays -- local i is global i-1, so the surface sits one cell
 * down-left of where the time shift says.  LOC() maps a local index to this file's N x N arrays. */
#define LOC(a, li, lj) cell(s, (a), (li) - 1, (lj) - 1)
#define TAPS(U, W, e, h) do {                                              \
    U[stp + m_e - 1] += (e) * b_e * coef;  W[stp + m_h - 1] += (h) * b_h * coef;  \
    U[stp + m_e]     += (e) * ab_e * coef; W[stp + m_h]     += (h) * ab_h * coef; \
    U[stp + m_e + 1] -= (e) * a_e * coef;  W[stp + m_h + 1] -= (h) * a_h * coef;  \
  } while (0)
static inline void ntff_coef(double time, double shift, int *m, double *a, double *b, double *ab)
{
  double t = time + shift;
  *m = floor(t + 0.5);
  *a = (0.5 + t - *m);
  *b = 1.0 - *a;
  *ab = *a - *b;
}
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

static void ntff_mpi(OracleSim *s)
{
  const double C = C0, R = 1.0e6;
  const double coef = 1.0 / (4 * M_PI * C * R);
  const int spx = s->npx + 2, spy = s->npy + 2;         /* SUB_N_PX, SUB_N_PY at one rank */
  const int cx = s->npx / 2, cy = s->npy / 2;
  const int tp = s->top, bm = s->bottom, rt = s->right, lt = s->left;
  const double timeE = s->time - 1, timeH = s->time - 0.5;
  cplx **f = s->f;
  int m_e, m_h;
  double a_e, b_e, ab_e, a_h, b_h, ab_h;
  for (int ang = 0; ang < 360; ang++) {
    double rad = ang * M_PI / 180.0;
    double r1x = cos(rad), r1y = sin(rad);
    const int stp = ang * s->array_size;
    if (s->kind == KIND_TM) {
      cplx *Ux = s->uw[0], *Uy = s->uw[1], *Wz = s->uw[2];
      if (0 < bm && bm < spy - 1)                       /* bottom: U_x += Ez, W_z += Hx */
        for (int i = imax(1, lt); i < imin(spx, rt); i++) {
          const double r2x = i - cx, r2y = bm - cy;
          const double shift = -(r1x * r2x + r1y * r2y) / C + s->rf_per_c;
          ntff_coef(timeE, shift, &m_e, &a_e, &b_e, &ab_e);
          ntff_coef(timeH, shift, &m_h, &a_h, &b_h, &ab_h);
          const cplx ez = LOC(f[EZ], i, bm);
          const cplx hx = 0.5 * (LOC(f[HX], i, bm) + LOC(f[HX], i, bm - 1));
          TAPS(Ux, Wz, ez, hx);
        }
      if (0 < rt && rt < spx - 1)                       /* right: U_y += Ez, W_z += Hy */
        for (int j = imax(1, bm); j < imin(spy, tp); j++) {
          const double r2x = rt - cx, r2y = j - cy;
          const double shift = -(r1x * r2x + r1y * r2y) / C + s->rf_per_c;
          ntff_coef(timeE, shift, &m_e, &a_e, &b_e, &ab_e);
          ntff_coef(timeH, shift, &m_h, &a_h, &b_h, &ab_h);
          const cplx ez = LOC(f[EZ], rt, j);
          const cplx hy = 0.5 * (LOC(f[HY], rt, j) + LOC(f[HY], rt - 1, j));
          TAPS(Uy, Wz, ez, hy);
        }
      if (0 < tp && tp < spy - 1)                       /* top */
        for (int i = imax(1, lt); i < imin(spx, rt); i++) {
          const double r2x = i - cx, r2y = tp - cy;
          const double shift = -(r1x * r2x + r1y * r2y) / C + s->rf_per_c;
          ntff_coef(timeE, shift, &m_e, &a_e, &b_e, &ab_e);
          ntff_coef(timeH, shift, &m_h, &a_h, &b_h, &ab_h);
          const cplx ez = -LOC(f[EZ], i, tp);
          const cplx hx = -0.5 * (LOC(f[HX], i, tp) + LOC(f[HX], i, tp - 1));
          TAPS(Ux, Wz, ez, hx);
        }
      if (0 < lt && lt < spx)                           /* left (the reference's own bound) */
        for (int j = imax(1, bm); j < imin(spy, tp); j++) {
          const double r2x = lt - cx, r2y = j - cy;
          const double shift = -(r1x * r2x + r1y * r2y) / C + s->rf_per_c;
          ntff_coef(timeE, shift, &m_e, &a_e, &b_e, &ab_e);
          ntff_coef(timeH, shift, &m_h, &a_h, &b_h, &ab_h);
          const cplx ez = -LOC(f[EZ], lt, j);
          const cplx hy = -0.5 * (LOC(f[HY], lt, j) + LOC(f[HY], lt - 1, j));
          TAPS(Uy, Wz, ez, hy);
        }
    } else {
      cplx *Wx = s->uw[0], *Wy = s->uw[1], *Uz = s->uw[2];
      if (0 < bm && bm < spy - 1)                       /* bottom: U_z -= Ex, W_x -= Hz */
        for (int i = imax(1, lt); i < imin(spx, rt); i++) {
          const double r2x = i - cx + 0.5, r2y = bm - cy;
          double shift = -(r1x * r2x + r1y * r2y) / C0 + s->rf_per_c;
          ntff_coef(timeE, shift, &m_e, &a_e, &b_e, &ab_e);
          ntff_coef(timeH, shift, &m_h, &a_h, &b_h, &ab_h);
          cplx ex = -LOC(f[EX], i, bm);
          cplx hz = -0.5 * (LOC(f[HZ], i, bm) + LOC(f[HZ], i, bm - 1));
          TAPS(Uz, Wx, ex, hz);
        }
      if (0 < rt && rt < spx - 1)                       /* right */
        for (int j = imax(1, bm); j < imin(spy, tp); j++) {
          double r2x = rt - cx, r2y = j - cy + 0.5;
          double shift = -(r1x * r2x + r1y * r2y) / C0 + s->rf_per_c;
          ntff_coef(timeE, shift, &m_e, &a_e, &b_e, &ab_e);
          ntff_coef(timeH, shift, &m_h, &a_h, &b_h, &ab_h);
          cplx ey = -LOC(f[EY], rt, j);
          cplx hz = -0.5 * (LOC(f[HZ], rt, j) + LOC(f[HZ], rt - 1, j));
          TAPS(Uz, Wy, ey, hz);
        }
      if (0 < tp && tp < spy - 1)                       /* top */
        for (int i = imax(1, lt); i < imin(spx, rt); i++) {
          const double r2x = i - cx + 0.5, r2y = tp - cy;
          double shift = -(r1x * r2x + r1y * r2y) / C0 + s->rf_per_c;
          ntff_coef(timeE, shift, &m_e, &a_e, &b_e, &ab_e);
          ntff_coef(timeH, shift, &m_h, &a_h, &b_h, &ab_h);
          cplx ex = LOC(f[EX], i, tp);
          cplx hz = 0.5 * (LOC(f[HZ], i, tp) + LOC(f[HZ], i, tp - 1));
          TAPS(Uz, Wx, ex, hz);
        }
      if (0 < lt && lt < spx)                           /* left */
        for (int j = imax(1, bm); j < imin(spy, tp); j++) {
          double r2x = lt - cx, r2y = j - cy + 0.5;
          double shift = -(r1x * r2x + r1y * r2y) / C0 + s->rf_per_c;
          ntff_coef(timeE, shift, &m_e, &a_e, &b_e, &ab_e);
          ntff_coef(timeH, shift, &m_h, &a_h, &b_h, &ab_h);
          cplx ey = LOC(f[EY], lt, j);
          cplx hz = 0.5 * (LOC(f[HZ], lt, j) + LOC(f[HZ], lt - 1, j));
          TAPS(Uz, Wy, ey, hz);
        }
    }
  }
}

/* n update() calls of solver id 4 (kind TM) / 5 (kind TE) on a sim made by oracle_create */
void oracle_step_mpi(OracleSim *s, int n, int with_ntff)
{
  for (int it = 0; it < n; it++) {
    if (s->kind == KIND_TM) step_mpi_tm(s); else step_mpi_te(s);
    if (with_ntff) ntff_mpi(s);
    s->time += 1.0;                                                   /* field.c:312-315 */
    s->ray_coef = 1.0 - exp(-pow(0.01 * s->time, 2));
  }
}

cplx *oracle_field(OracleSim *s, int slot) { return s->f[slot]; }
double *oracle_coef(OracleSim *s, int slot) { return s->c[slot]; }
cplx *oracle_uw(OracleSim *s, int slot) { return s->uw[slot]; }
int oracle_array_size(const OracleSim *s) { return s->array_size; }
int oracle_npx(const OracleSim *s) { return s->npx; }
int oracle_npy(const OracleSim *s) { return s->npy; }
double oracle_time(const OracleSim *s) { return s->time; }
void oracle_ntff_box(const OracleSim *s, int *out6) {
  out6[0] = s->top; out6[1] = s->bottom; out6[2] = s->left; out6[3] = s->right; out6[4] = s->cx; out6[5] = s->cy;
}

/* ---- ntffTM.c:72-158: one-shot frequency-domain surface integral (TM) -------------- */
void oracle_frequency_tm(OracleSim *s, cplx *result)
{
  const int N = s->npy;
  const cplx *Ez = s->f[EZ], *Hx = s->f[HX], *Hy = s->f[HY];
  double R0 = 1.0e6 * (500.0 / s->h_u_nm);                                   /* ntffTM.c:31 */
  double cx = s->cx, cy = s->cy, k_s = s->k_s;
  cplx coef = csqrt(I * k_s / (8 * M_PI * R0)) * cexp(I * k_s * R0);
  for (int ang = 0; ang < N_ANG; ang++) {
    double rad = ang * M_PI / 180.0;
    double rx = cos(rad), ry = sin(rad);
    cplx Nz = 0, Lx = 0, Ly = 0;
    for (int i = s->left; i < s->right; i++) {                               /* bottom */
      int k = i * N + s->bottom;
      double inner = rx * (i - cx) + ry * (s->bottom - cy);
      Nz += 0.5 * (Hx[k] + Hx[k - 1]) * cexp(I * k_s * inner);
      Lx += Ez[k] * cexp(I * k_s * inner);
    }
    for (int j = s->bottom; j < s->top; j++) {                               /* right */
      int k = s->right * N + j;
      double inner = rx * (s->right - cx) + ry * (j - cy);
      Nz += 0.5 * (Hy[k] + Hy[k - N]) * cexp(I * k_s * inner);
      Ly += Ez[k] * cexp(I * k_s * inner);
    }
    for (int i = s->left; i < s->right; i++) {                               /* top */
      int k = i * N + s->top;
      double inner = rx * (i - cx) + ry * (s->top - cy);
      Nz -= 0.5 * (Hx[k] + Hx[k - 1]) * cexp(I * k_s * inner);
      Lx -= Ez[k] * cexp(I * k_s * inner);
    }
    for (int j = s->bottom; j < s->top; j++) {                               /* left */
      int k = s->left * N + j;
      double inner = rx * (s->left - cx) + ry * (j - cy);
      Nz -= 0.5 * (Hy[k] + Hy[k - N]) * cexp(I * k_s * inner);
      Ly -= Ez[k] * cexp(I * k_s * inner);
    }
    cplx Lphi = -Lx * sin(rad) + Ly * cos(rad);
    result[ang] = coef * (Z0 * Nz + Lphi) * sqrt(s->h_u_nm);
  }
}

/* ---- cfft.c:104-179: radix-2 DIF, e^{+i...} twiddles, then bit reversal ----- */
void oracle_fft(cplx *a, int n)
{
  int iter = 0;
  for (int v = n; v >>= 1;) iter++;
  const double sign = -1.;
  int span = n;
  for (int it = 0; it < iter; it++) {
    int full = span;
    span = full / 2;
    double w = -M_PI / span;
    for (int k = 0; k < span; k++) {
      cplx ww = cexp(I * sign * w * k);
      for (int base = k; base + span < n; base += full) {
        cplx t = a[base] - a[base + span];
        a[base] = a[base] + a[base + span];
        a[base + span] = t * ww;
      }
    }
  }
  for (int i = 0, j = 0; i < n; i++) {            /* in-place bit reversal */
    if (i < j) { cplx t = a[i]; a[i] = a[j]; a[j] = t; }
    int bit = n >> 1;
    while (bit && (j & bit)) { j ^= bit; bit >>= 1; }
    j |= bit;
  }
}

/* ntffTM.c:161-232 / ntffTE.c:20-55,160-195: out[(lambda-380)*360 + ang] */
void oracle_far_field(OracleSim *s, double *out)
{
  const cplx coef = 1.0 / (4 * M_PI * C0) * csqrt(2 * M_PI * C0 / (I * s->omega_s));
  const int tm = (s->kind == KIND_TM);
  const double theta = 0, to_rad = M_PI / 180.0;
  cplx *series = (cplx *)zalloc(N_FFT, sizeof(cplx));
  for (int ang = 0; ang < N_ANG; ang++) {
    double phi = tm ? ang * to_rad : ang * M_PI / 180.0;
    double sx = cos(theta) * cos(phi), sy = cos(theta) * sin(phi), sz = -cos(theta);
    double px = -sin(phi), py = cos(phi);
    (void)sx; (void)sy;
    size_t k = (size_t)ang * s->array_size;
    memset(series, 0, sizeof(cplx) * N_FFT);
    for (int n = 0; n < s->steps; n++) {
      if (tm) {
        cplx WTH = 0 + 0 + s->uw[2][k + n] * sz;
        cplx UPH = s->uw[0][k + n] * px + s->uw[1][k + n] * py;
        series[n] = coef * (-Z0 * WTH - UPH);                         /* Eth */
      } else {
        cplx WPH = s->uw[0][k + n] * px + s->uw[1][k + n] * py;
        cplx UTH = 0 + 0 + s->uw[2][k + n] * sz;
        series[n] = coef * (-Z0 * WPH + UTH);                         /* Eph */
      }
    }
    oracle_fft(series, N_FFT);
    for (int lam = LAM_FIRST; lam <= LAM_LAST; lam++) {
      double p = C0 * s->h_u_nm * N_FFT / lam;
      int idx = floor(p);
      p = p - idx;
      double n0 = creal(series[idx]) * creal(series[idx]) + cimag(series[idx]) * cimag(series[idx]);
      double n1 = creal(series[idx + 1]) * creal(series[idx + 1]) + cimag(series[idx + 1]) * cimag(series[idx + 1]);
      out[(size_t)(lam - LAM_FIRST) * N_ANG + ang] = ((1 - p) * n0 + p * n1) / N_FFT;
    }
  }
  free(series);
}
