/* farfield.c -- host side of the NTFF path: the sampling plan handed to the GPU
 * engine, the constants of the far-field post-processing, and the on-disk
 * formats.
 *
 * What it restates (rennone/mpiFDTD): the per-direction time-shift recurrence of
 * ntffTM_TimeCalc / ntffTE_TimeCalc (ntffTM.c:316-369, ntffTE.c:90-155), the
 * translate coefficient and direction cosines of ntffT?_TimeTranslate
 * (ntffTM.c:161-179, ntffTE.c:20-40), the twiddle factors of the radix-2 FFT
 * (cfft.c:131-141) and the writers of ntff.c:6-33.  Everything here is evaluated
 * with the host libm so the constants match the reference bit for bit; all
 * summation over the surface history runs on the GPU.
 */
#define _USE_MATH_DEFINES
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "host_internal.h"

#ifndef M_PI
#define M_PI 3.1415926535897932384626433832795
#endif

int mpifdtd_ntff_point_count(const NTFFInfo *box)
{
  return 2 * (box->right - box->left) + 2 * (box->top - box->bottom);
}

/* Number of surface points whose column j lies in [j0, j0+nj). */
int mpifdtd_ntff_local_count(const NTFFInfo *box, int j0, int nj)
{
  const int j1 = j0 + nj;
  int n = 0;
  if (box->bottom >= j0 && box->bottom < j1) n += box->right - box->left;
  if (box->top >= j0 && box->top < j1) n += box->right - box->left;
  int lo = box->bottom > j0 ? box->bottom : j0, hi = box->top < j1 ? box->top : j1;
  if (hi > lo) n += 2 * (hi - lo);
  return n;
}

/* table[a*L + q] = the value `timeShift` holds when the reference visits the
 * q-th surface point owned by the slab j in [j0, j0+nj), for direction a.  Points
 * are ordered bottom, right, top, left.  Each edge starts from -(r1 . r2) + RFperC
 * at its first cell and is then decremented once per cell, exactly like the
 * running variable in the reference loops (so the accumulated rounding is the
 * same); cells outside the slab are walked but not stored.  stagger = 0 for TM,
 * 0.5 for TE, where the E samples sit half a cell along the edge
 * (ntffTE.c:102,116,130,145). */
double *mpifdtd_ntff_time_shift(const NTFFInfo *box, int n_angles, double stagger, int j0, int nj)
{
  const int nx = box->right - box->left, ny = box->top - box->bottom;
  const int L = mpifdtd_ntff_local_count(box, j0, nj);
  double *table = (double *)malloc(sizeof(double) * (size_t)n_angles * (size_t)(L > 0 ? L : 1));
  if (table == NULL) { printf("cannot allocate NTFF time-shift table\n"); exit(2); }

  const double lt_cx = box->left - box->cx,   rt_cx = box->right - box->cx;
  const double bm_cy = box->bottom - box->cy, tp_cy = box->top - box->cy;
  const double to_rad = M_PI / 180.0;
  const int edge_j[4] = { box->bottom, -1, box->top, -1 };     /* fixed j of the x-edges */

  for (int a = 0; a < n_angles; a++) {
    double rad = a * to_rad;
    double r1x = cos(rad) / C_0_S, r1y = sin(rad) / C_0_S;
    double *row = table + (size_t)a * L;
    /* edge start vectors r2 = first cell - centre */
    const double start[4][2] = { { lt_cx + stagger, bm_cy },      /* bottom: (l,b) -> (r,b) */
                                 { rt_cx, bm_cy + stagger },      /* right:  (r,b) -> (r,t) */
                                 { lt_cx + stagger, tp_cy },      /* top:    (l,t) -> (r,t) */
                                 { lt_cx, bm_cy + stagger } };    /* left:   (l,b) -> (l,t) */
    int q = 0;
    for (int edge = 0; edge < 4; edge++) {
      const int along_x = (edge == 0 || edge == 2);
      const int len = along_x ? nx : ny;
      const double step = along_x ? r1x : r1y;
      double shift = -(r1x * start[edge][0] + r1y * start[edge][1]) + box->RFperC;
      for (int n = 0; n < len; n++) {
        const int j = along_x ? edge_j[edge] : box->bottom + n;
        if (j >= j0 && j < j0 + nj)
          row[q++] = shift;
        shift -= step;
      }
    }
  }
  return table;
}

/* The MPI-variant solvers' ntff() (mpiTM_UPML.c:849-1037, mpiTE_UPML.c:602-794) evaluates
 * timeShift = -(r1x*r2x + r1y*r2y)/C + RFperC afresh at every surface point, with
 * r1 = (cos, sin) NOT pre-divided by C, and (TE) the half-cell stagger added to the integer
 * difference: r2x = i - cx + 0.5.  Same point order as above.  A y-slab keeps the points whose
 * SAMPLED cell column j + sample_dj lies in [j0, j0+nj) (these solvers read one cell down-left of
 * where r2 says); the whole surface with j0 = 0, nj = N_PY. */
int mpifdtd_ntff_local_count_shifted(const NTFFInfo *box, int sample_dj, int j0, int nj)
{
  NTFFInfo moved = *box;
  moved.bottom += sample_dj;
  moved.top += sample_dj;
  return mpifdtd_ntff_local_count(&moved, j0, nj);
}

double *mpifdtd_ntff_time_shift_direct(const NTFFInfo *box, int n_angles, double stagger, int sample_dj, int j0, int nj)
{
  const int nx = box->right - box->left, ny = box->top - box->bottom;
  const int P = mpifdtd_ntff_local_count_shifted(box, sample_dj, j0, nj);
  double *table = (double *)malloc(sizeof(double) * (size_t)n_angles * (size_t)(P > 0 ? P : 1));
  if (table == NULL) { printf("cannot allocate NTFF time-shift table\n"); exit(2); }
  const int cx = box->cx, cy = box->cy;          /* N_PX/2 - offsetX with offset 0 */
  for (int a = 0; a < n_angles; a++) {
    double rad = a * M_PI / 180.0;
    double r1x = cos(rad), r1y = sin(rad);
    double *row = table + (size_t)a * P;
    int q = 0;
    for (int edge = 0; edge < 4; edge++) {
      const int along_x = (edge == 0 || edge == 2);
      const int len = along_x ? nx : ny;
      for (int n = 0; n < len; n++) {
        double r2x, r2y;
        if (along_x) {
          const int i = box->left + n, j = (edge == 0) ? box->bottom : box->top;
          r2x = i - cx + stagger;  r2y = j - cy;
        } else {
          const int i = (edge == 1) ? box->right : box->left, j = box->bottom + n;
          r2x = i - cx;            r2y = j - cy + stagger;
        }
        const int j_point = along_x ? ((edge == 0) ? box->bottom : box->top) : box->bottom + n;
        if (j_point + sample_dj < j0 || j_point + sample_dj >= j0 + nj) continue;
        row[q++] = -(r1x * r2x + r1y * r2y) / C_0_S + box->RFperC;
      }
    }
  }
  return table;
}

/* 1/(4 pi C) * csqrt(2 pi C / (i omega))  (ntffTM.c:165, ntffTE.c:25) */
double complex mpifdtd_ntff_translate_coef(double omega)
{
  return 1.0 / (4 * M_PI * C_0_S) * csqrt(2 * M_PI * C_0_S / (I * omega));
}

/* cos(phi), sin(phi) per direction.  TM forms phi as ang*(pi/180) (ntffTM.c:169,173),
 * TE as ang*pi/180 (ntffTE.c:33): different roundings, both kept. */
void mpifdtd_ntff_direction_cosines(int n_angles, int is_tm, double *cos_phi, double *sin_phi)
{
  const double to_rad = M_PI / 180.0;
  for (int a = 0; a < n_angles; a++) {
    double phi = is_tm ? a * to_rad : a * M_PI / 180.0;
    cos_phi[a] = cos(phi);
    sin_phi[a] = sin(phi);
  }
}

/* Twiddles of the n-point decimation-in-frequency FFT, stage by stage (span
 * halves n/2, n/4, ..., 1): cexp(I*sign*w*k) with sign = -1, w = -pi/half
 * (cfft.c:131-141), i.e. the positive-exponent transform.  n-1 complex values. */
double complex *mpifdtd_fft_twiddles(int n)
{
  double complex *tw = (double complex *)malloc(sizeof(double complex) * (size_t)(n > 1 ? n - 1 : 1));
  if (tw == NULL) { printf("cannot allocate FFT twiddles\n"); exit(2); }
  const double sign = -1.;
  size_t at = 0;
  for (int half = n / 2; half >= 1; half /= 2) {
    double w = -M_PI / half;
    for (int k = 0; k < half; k++)
      tw[at++] = cexp(I * sign * w * k);
  }
  return tw;
}

/* ---- one-shot frequency-domain far field (ntffTM_Frequency, ntffTM.c:72-158) ---------
 * The surface sums run on the GPU over the engine's current fields; the closing algebra
 * (polar component, coef, sqrt(h_u)) is the reference's, evaluated here with host libm.
 * Serves the TM-type solvers (ids 0, 2, 4, 6).  result[360] like resultEz. */
#include "b200fdtd.h"
extern b200fdtd_engine *mpifdtd_upml_engine(int kind);
extern b200fdtd_engine *mpifdtd_split_engine(int kind);

int mpifdtd_ntffFrequency(int solver_id, double complex result[360])
{
  b200fdtd_engine *engine = (solver_id == TM_UPML_2D || solver_id == MPI_TM_UPML_2D)
                                ? mpifdtd_upml_engine(solver_id) : mpifdtd_split_engine(solver_id);
  if (engine == NULL) { printf("ntffFrequency: solver %d is not initialised\n", solver_id); exit(2); }
  NTFFInfo box = field_getNTFFInfo();
  FieldInfo phys = field_getFieldInfo();
  double k_s = field_getK();
  double R0 = 1.0e6 * field_toCellUnit(500);                  /* ntffTM.c:31 */
  double complex coef = csqrt(I * k_s / (8 * M_PI * R0)) * cexp(I * k_s * R0);
  double cos_a[360], sin_a[360];
  for (int ang = 0; ang < 360; ang++) {
    double rad = ang * M_PI / 180.0;
    cos_a[ang] = cos(rad);
    sin_a[ang] = sin(rad);
  }
  b200fdtd_freq_args fa = { box.top, box.bottom, box.left, box.right, box.cx, box.cy, 360, 0, k_s, cos_a, sin_a };
  double complex sums[3][360];
  int rc = b200fdtd_ntff_frequency(engine, &fa, (double *)sums);
  if (rc != B200FDTD_OK) { printf("b200fdtd: ntff_frequency failed (%d): %s\n", rc, b200fdtd_last_error()); exit(2); }
  for (int ang = 0; ang < 360; ang++) {
    double complex Nz = sums[0][ang], Lx = sums[1][ang], Ly = sums[2][ang];
    double complex Lphi = -Lx * sin_a[ang] + Ly * cos_a[ang];
    result[ang] = coef * (Z_0_S * Nz + Lphi) * sqrt(phys.h_u_nm);
  }
  return 0;
}

/* ---- on-disk formats (ntff.c:6-33) ----------------------------------------
 * text: one line per wavelength, "<nm> " then 360 values printed with "%lf.20 "
 * (upstream's format string: six decimals followed by the literal ".20");
 * binary: 321 rows of 360 float64, row = wavelength 380..700 nm. */
/* The text twin is 321 x 360 printf conversions (~21 ms single-threaded), which is what an
 * angle sweep on the GPU ends up waiting for; the rows are formatted by a few threads with the
 * same conversion ("%lf.20 ") into private buffers and written out in order, so the file is
 * byte-identical to the sequential fprintf loop. */
#include <pthread.h>
#include <string.h>
typedef struct TxtJob { double **e_norm; int row0, row1; char *buf; size_t len; } TxtJob;

static void *format_rows(void *arg)
{
  TxtJob *job = (TxtJob *)arg;
  const size_t cap = (size_t)(job->row1 - job->row0) * (16 + 360 * 40) + 64;   /* "%lf.20 " of |x| < 1e21 is <= 34 chars */
  char *p = job->buf = (char *)malloc(cap);
  if (p == NULL) return NULL;
  for (int row = job->row0; row < job->row1; row++) {
    p += sprintf(p, "%d ", LAMBDA_ST_NM + row);
    for (int ang = 0; ang < 360; ang++) {
      const double v = job->e_norm[row][ang];
      if (!(v > -1e21 && v < 1e21)) {            /* huge / inf / nan: keep the bounded path safe */
        p += snprintf(p, 400, "%lf.20 ", v);
        if ((size_t)(p - job->buf) + 512 > cap) break;
        continue;
      }
      p += sprintf(p, "%lf.20 ", v);
    }
    *p++ = '\n';
  }
  job->len = (size_t)(p - job->buf);
  return NULL;
}

void ntff_outputEnormTxt(double **e_norm, const char *file_name)
{
  FILE *fp = FileOpen(file_name, "w");
  enum { N_THREADS = 8 };
  const int rows = LAMBDA_EN_NM - LAMBDA_ST_NM + 1;
  TxtJob jobs[N_THREADS];
  pthread_t tid[N_THREADS];
  int started[N_THREADS];
  for (int t = 0; t < N_THREADS; t++) {
    jobs[t].e_norm = e_norm;
    jobs[t].row0 = rows * t / N_THREADS;
    jobs[t].row1 = rows * (t + 1) / N_THREADS;
    jobs[t].buf = NULL; jobs[t].len = 0;
    started[t] = pthread_create(&tid[t], NULL, format_rows, &jobs[t]) == 0;
    if (!started[t]) format_rows(&jobs[t]);        /* no thread to be had: format inline */
  }
  for (int t = 0; t < N_THREADS; t++) {
    if (started[t]) pthread_join(tid[t], NULL);
    if (jobs[t].buf == NULL) { printf("cannot allocate the text buffer for %s\n", file_name); exit(2); }
    fwrite(jobs[t].buf, 1, jobs[t].len, fp);
    free(jobs[t].buf);
  }
  fclose(fp);
}

void ntff_outputEnormBin(double **e_norm, const char *file_name)
{
  FILE *fp = FileOpen(file_name, "wb");
  for (int row = 0; row <= LAMBDA_EN_NM - LAMBDA_ST_NM; row++) {
    if (fwrite(e_norm[row], sizeof(double), 360, fp) < 360) {
      printf("error in write binary %s\n", file_name);
      exit(2);
    }
  }
  fclose(fp);
}
