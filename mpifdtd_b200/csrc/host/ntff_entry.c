/* ntff_entry.c -- the reference's PUBLIC NTFF entry points (ntffTM.h:7-28, ntffTE.h:5-15) for a
 * maintainer who keeps a solver .c file of their own and links it against this library.
 *
 * Their arguments are the caller's HOST arrays.  What the reference does with them per step --
 * walk the closed surface for each of 360 directions and scatter every tangential sample into three
 * retarded-time bins (ntffTM.c:279-371, ntffTE.c:57-157), ~85 % of its step -- is restructured here
 * exactly as for the built-in solvers: TimeCalc only GATHERS the surface samples (O(perimeter) host
 * work, the same reads, signs and two-cell H averages) and hands them to an engine that has no
 * field arrays at all (B200FDTD_GRID_NTFF_ONLY); the 360-direction binning runs once, on the GPU,
 * when the accumulators are looked at.  Consequence, and the one visible difference: the caller's
 * U / W arrays are filled in when they are handed to TimeTranslate / TimeOutput (or to
 * mpifdtd_ntffSync), not after every TimeCalc.  Values: within 1e-13 of the reference's
 * (summation order), all arraySize bins including the row-spill quirk.
 *
 * ntffTM_Frequency is a one-shot sum over the surface of HOST fields (360 x perimeter cexp); it is
 * evaluated where the data is.  Not exported: ntffTE_Frequency (upstream's divides by a static R0 that
 * nothing initialises and indexes with field_subIndex, ntffTE.c:12,242-338; no caller), the *_Split /
 * ntff_TMTime_MPI sub-domain variants and ntffTM_FreqOutput (no callers either).
 */
#define _USE_MATH_DEFINES
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "b200fdtd.h"
#include "host_internal.h"

#ifndef M_PI
#define M_PI 3.1415926535897932384626433832795
#endif
#define N_ANGLES 360
#define R0 (1.0e6 * field_toCellUnit(500))      /* ntffTM.c:31 */

typedef struct Standalone {
  int tm;
  b200fdtd_engine *engine;
  int n_points, array_size, max_time;
  dcomplex *e_buf, *h_buf;          /* one step's samples, perimeter order */
  int dirty;                        /* samples pushed since the last projection */
} Standalone;

static Standalone st_tm = { .tm = 1 }, st_te = { .tm = 0 };

static void die_on(int rc, const char *what)
{
  if (rc == B200FDTD_OK) return;
  printf("b200fdtd: %s failed (%d): %s\n", what, rc, b200fdtd_last_error());
  exit(2);
}

static void standalone_free(Standalone *s)
{
  if (s->engine != NULL) die_on(b200fdtd_destroy(s->engine), "b200fdtd_destroy");
  s->engine = NULL;
  free(s->e_buf); free(s->h_buf);
  s->e_buf = s->h_buf = NULL;
}

/* ntffTM_init / ntffTE_init: upstream's only computes sub-domain bounds (ntffTM.c:27-66); here it
 * (re)creates the accumulation engine for the current field_init() state */
static void standalone_init(Standalone *s)
{
  FieldInfo_S g = field_getFieldInfo_S();
  NTFFInfo box = field_getNTFFInfo();
  standalone_free(s);
  b200fdtd_grid grid;
  memset(&grid, 0, sizeof grid);
  grid.kind = s->tm ? B200FDTD_TM_UPML : B200FDTD_TE_UPML;
  grid.n_px = g.N_PX;  grid.n_py = g.N_PY;  grid.n_pml = g.N_PML;
  grid.j0 = 0;         grid.nj = g.N_PY;
  grid.i_lo = 1;       grid.i_hi = g.N_PX - 2;
  grid.j_lo = 1;       grid.j_hi = g.N_PY - 2;
  grid.device = -1;
  grid.mu0 = MU_0_S;
  grid.flags = B200FDTD_GRID_NTFF_ONLY;
  die_on(b200fdtd_create(&grid, &s->engine), "b200fdtd_create");

  b200fdtd_ntff_plan plan;
  memset(&plan, 0, sizeof plan);
  plan.top = box.top; plan.bottom = box.bottom; plan.left = box.left; plan.right = box.right;
  plan.n_points = mpifdtd_ntff_point_count(&box);
  plan.n_local = plan.n_points;
  plan.max_time = (int)field_getMaxTime();
  plan.n_bins = box.arraySize;                      /* the caller's arrays are [360][arraySize] */
  plan.n_angles = N_ANGLES;
  plan.array_size = box.arraySize;
  double *shift = mpifdtd_ntff_time_shift(&box, N_ANGLES, s->tm ? 0.0 : 0.5, 0, g.N_PY);
  plan.time_shift = shift;
  if (plan.n_points < 1 || plan.max_time < 1) { printf("ntff init: no surface or no steps\n"); exit(2); }
  die_on(b200fdtd_set_ntff_plan(s->engine, &plan), "b200fdtd_set_ntff_plan");
  free(shift);
  s->n_points = plan.n_points;  s->array_size = box.arraySize;  s->max_time = plan.max_time;
  s->e_buf = (dcomplex *)malloc(sizeof(dcomplex) * (size_t)s->n_points);
  s->h_buf = (dcomplex *)malloc(sizeof(dcomplex) * (size_t)s->n_points);
  s->dirty = 0;
}

void ntffTM_init(void) { standalone_init(&st_tm); }
void ntffTE_init(void) { standalone_init(&st_te); }

static Standalone *ready(Standalone *s)
{
  if (s->engine == NULL) standalone_init(s);        /* upstream's solvers call init; a caller who did not still works */
  return s;
}

/* The surface sample of one step, from host arrays indexed k = i*N_PY + j, in the perimeter order
 * bottom, right, top, left (corner (right, top) excluded, edges half-open): e_a / h on the edges
 * along x, e_b / h on the edges along y; the H value is the two-cell average 0.5*(H[k] + H[k - 1])
 * resp. 0.5*(H[k] + H[k - N_PY]); `neg_first` negates bottom and right (TE), else top and left (TM).
 * ntffTM.c:326-369, ntffTE.c:100-155. */
static void gather(Standalone *s, const dcomplex *e_a, const dcomplex *e_b, const dcomplex *h_a, const dcomplex *h_b,
                   int neg_first)
{
  NTFFInfo box = field_getNTFFInfo();
  const int P = N_PY;
  int q = 0;
  for (int edge = 0; edge < 4; edge++) {
    const int along_x = (edge == 0 || edge == 2);
    const int negate = neg_first ? (edge < 2) : (edge >= 2);
    const int len = along_x ? box.right - box.left : box.top - box.bottom;
    for (int n = 0; n < len; n++, q++) {
      const int i = along_x ? box.left + n : (edge == 1 ? box.right : box.left);
      const int j = along_x ? (edge == 0 ? box.bottom : box.top) : box.bottom + n;
      const int k = i * P + j;
      dcomplex ev = along_x ? e_a[k] : e_b[k];
      dcomplex hv = along_x ? 0.5 * (h_a[k] + h_a[k - 1]) : 0.5 * (h_b[k] + h_b[k - P]);
      if (negate) { ev = -ev; hv = -hv; }
      s->e_buf[q] = ev;
      s->h_buf[q] = hv;
    }
  }
  const int t = (int)field_getTime();
  if (t < 0 || t >= s->max_time) return;            /* past maxTime upstream writes beyond what anyone reads */
  die_on(b200fdtd_ntff_push_samples(s->engine, t, (const double *)s->e_buf, (const double *)s->h_buf),
         "b200fdtd_ntff_push_samples");
  s->dirty = 1;
}

/* project what has been gathered and store it into the caller's accumulators */
static void sync_to(Standalone *s, dcomplex *a0, dcomplex *a1, dcomplex *a2)
{
  if (!s->dirty) return;
  dcomplex *dst[3] = { a0, a1, a2 };
  die_on(b200fdtd_ntff_project(s->engine), "b200fdtd_ntff_project");
  for (int m = 0; m < 3; m++)
    if (dst[m] != NULL) die_on(b200fdtd_ntff_get_uw(s->engine, m, (double *)dst[m]), "b200fdtd_ntff_get_uw");
  s->dirty = 0;
}

/* ---- TM (ntffTM.h) ------------------------------------------------------------------------------- */
void ntffTM_TimeCalc(dcomplex *Hx, dcomplex *Hy, dcomplex *Ez, dcomplex *Ux, dcomplex *Uy, dcomplex *Wz)
{
  (void)Ux; (void)Uy; (void)Wz;                     /* filled by the next TimeTranslate / TimeOutput / mpifdtd_ntffSync */
  gather(ready(&st_tm), Ez, Ez, Hx, Hy, 0);
}

/* E_theta, E_phi at theta = 0 from the accumulators (ntffTM.c:161-194, ntffTE.c:20-55): TM has the
 * electric current along z (W = (0, 0, a2)) and the magnetic one in the plane (U = (a0, a1, 0)), TE the
 * other way round.  The direction angle is rounded as each upstream file rounds it
 * (mpifdtd_ntff_direction_cosines); the "+ 0" terms of the upstream expressions are kept (they turn a
 * -0 into +0).  Only the first maxTime of arraySize bins, like upstream. */
static void translate(int tm, const dcomplex *a0, const dcomplex *a1, const dcomplex *a2, dcomplex *Eth, dcomplex *Eph)
{
  const double complex coef = mpifdtd_ntff_translate_coef(field_getOmega());
  const int n_time = (int)field_getMaxTime(), stride = field_getNTFFInfo().arraySize;
  double cos_phi[N_ANGLES], sin_phi[N_ANGLES];
  mpifdtd_ntff_direction_cosines(N_ANGLES, tm, cos_phi, sin_phi);
  const double theta = 0;
  for (int ang = 0; ang < N_ANGLES; ang++) {
    const double sx = cos(theta) * cos_phi[ang], sy = cos(theta) * sin_phi[ang], sz = -cos(theta);
    const double px = -sin_phi[ang], py = cos_phi[ang];
    for (int n = 0; n < n_time; n++) {
      const size_t k = (size_t)ang * stride + n;
      const double complex in_th = a0[k] * sx + a1[k] * sy + 0, in_ph = a0[k] * px + a1[k] * py;   /* the in-plane current */
      const double complex z_th = 0 + 0 + a2[k] * sz, z_ph = 0 + 0;                                 /* the one along z */
      const double complex WTH = tm ? z_th : in_th, WPH = tm ? z_ph : in_ph;
      const double complex UTH = tm ? in_th : z_th, UPH = tm ? in_ph : z_ph;
      Eth[k] = coef * (-Z_0_S * WTH - UPH);
      Eph[k] = coef * (-Z_0_S * WPH + UTH);
    }
  }
}

void ntffTM_TimeTranslate(dcomplex *Ux, dcomplex *Uy, dcomplex *Wz, dcomplex *Eth, dcomplex *Eph)
{
  sync_to(ready(&st_tm), Ux, Uy, Wz);
  translate(1, Ux, Uy, Wz, Eth, Eph);
}

/* translate + 8192-point FFT + wavelength interpolation on the GPU, the two files into cwd
 * (ntffTM.c:197-275, ntffTE.c:160-238; writers byte-identical, tests/test_formats_cpu.py) */
static void time_output(Standalone *s, dcomplex *a0, dcomplex *a1, dcomplex *a2)
{
  sync_to(s, a0, a1, a2);
  dcomplex *src[3] = { a0, a1, a2 };
  for (int m = 0; m < 3; m++)                       /* the arrays the caller hands in are what gets transformed */
    die_on(b200fdtd_ntff_set_uw(s->engine, m, (const double *)src[m]), "b200fdtd_ntff_set_uw");
  const int rows = LAMBDA_EN_NM - LAMBDA_ST_NM + 1;
  double *table = (double *)malloc(sizeof(double) * (size_t)rows * N_ANGLES);
  double **by_row = (double **)malloc(sizeof(double *) * (size_t)rows);
  for (int r = 0; r < rows; r++) by_row[r] = table + (size_t)r * N_ANGLES;
  mpifdtd_upml_far_field(s->engine, s->tm ? B200FDTD_TM_UPML : B200FDTD_TE_UPML, 0, table);
  char name[256], cwd[512];
  if (getcwd(cwd, sizeof cwd) == NULL) cwd[0] = '\0';
  sprintf(name, "%d[deg].txt", (int)field_getWaveAngle());
  ntff_outputEnormTxt(by_row, name);
  printf("saved %s/%s\n", cwd, name);
  sprintf(name, "%d[deg]_%dnm_%dnm_b.dat", (int)field_getWaveAngle(), LAMBDA_ST_NM, LAMBDA_EN_NM);
  ntff_outputEnormBin(by_row, name);
  printf("saved %s/%s\n", cwd, name);
  free(by_row); free(table);
}

void ntffTM_TimeOutput(dcomplex *Ux, dcomplex *Uy, dcomplex *Wz) { time_output(ready(&st_tm), Ux, Uy, Wz); }

/* ntffTM.c:72-158, one-shot: for each direction, Nz = sum over the surface of the tangential H, Lx /
 * Ly = that of Ez on the edges along x / along y, each times cexp(i k r^ . r2); top and left enter
 * negated; result coef * (Z0 Nz + L_phi) * sqrt(h_u).  One walk round the perimeter in upstream's edge
 * order (bottom, right, top, left), so the sums round the same way. */
void ntffTM_Frequency(dcomplex *Hx, dcomplex *Hy, dcomplex *Ez, dcomplex resultEz[360])
{
  const NTFFInfo box = field_getNTFFInfo();
  const int stride = field_getFieldInfo_S().N_PY;
  const double k_s = field_getK(), cx = box.cx, cy = box.cy;
  const double complex coef = csqrt(I * k_s / (8 * M_PI * R0)) * cexp(I * k_s * R0);
  for (int ang = 0; ang < N_ANGLES; ang++) {
    const double rad = ang * M_PI / 180.0, rx = cos(rad), ry = sin(rad);
    dcomplex Nz = 0, L[2] = { 0, 0 };                     /* L[0] = Lx (edges along x), L[1] = Ly */
    for (int edge = 0; edge < 4; edge++) {
      const int along_x = (edge == 0 || edge == 2), far_side = edge >= 2;
      const int len = along_x ? box.right - box.left : box.top - box.bottom;
      for (int n = 0; n < len; n++) {
        const int i = along_x ? box.left + n : (edge == 1 ? box.right : box.left);
        const int j = along_x ? (edge == 0 ? box.bottom : box.top) : box.bottom + n;
        const int k = i * stride + j;
        const dcomplex h = along_x ? 0.5 * (Hx[k] + Hx[k - 1]) : 0.5 * (Hy[k] + Hy[k - stride]);
        const dcomplex phase = cexp(I * k_s * (rx * (i - cx) + ry * (j - cy)));
        if (far_side) { Nz -= h * phase;  L[!along_x] -= Ez[k] * phase; }
        else          { Nz += h * phase;  L[!along_x] += Ez[k] * phase; }
      }
    }
    const double complex Lphi = -L[0] * sin(rad) + L[1] * cos(rad);
    resultEz[ang] = coef * (Z_0_S * Nz + Lphi) * sqrt(field_getFieldInfo().h_u_nm);
  }
}

/* ---- TE (ntffTE.h) ------------------------------------------------------------------------------- */
void ntffTE_TimeCalc(dcomplex *Ex, dcomplex *Ey, dcomplex *Hz, dcomplex *Wx, dcomplex *Wy, dcomplex *Uz)
{
  (void)Wx; (void)Wy; (void)Uz;
  gather(ready(&st_te), Ex, Ey, Hz, Hz, 1);
}

void ntffTE_TimeTranslate(dcomplex *Wx, dcomplex *Wy, dcomplex *Uz, dcomplex *Eth, dcomplex *Eph)
{
  sync_to(ready(&st_te), Wx, Wy, Uz);
  translate(0, Wx, Wy, Uz, Eth, Eph);
}

void ntffTE_TimeOutput(dcomplex *Wx, dcomplex *Wy, dcomplex *Uz) { time_output(ready(&st_te), Wx, Wy, Uz); }

/* Fill the caller's accumulators now (the arrays TimeCalc was given), without translating: for a
 * caller that reads U / W itself between TimeCalc and TimeOutput.  tm != 0: Ux, Uy, Wz; else Wx, Wy, Uz. */
void mpifdtd_ntffSync(int tm, dcomplex *a0, dcomplex *a1, dcomplex *a2)
{
  sync_to(ready(tm ? &st_tm : &st_te), a0, a1, a2);
}

/* release the accumulation engines (upstream's ntffTM_finish is empty, ntffTM.c:68-71) */
void ntffTM_finish(void) { standalone_free(&st_tm); }
void ntffTE_finish(void) { standalone_free(&st_te); }
