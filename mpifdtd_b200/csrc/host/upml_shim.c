/* upml_shim.c -- the serial UPML solvers' entry points (fdtdTM_upml_get*,
 * fdtdTE_upml_get*) implemented over the GPU engine of include/b200fdtd.h.
 *
 * Lifecycle contract kept from rennone/mpiFDTD (fdtdTM_upml.c:54-135,
 * fdtdTE_upml.c:60-192):
 *   init()   after field_init + models_initModel: allocate, build eps maps and
 *            coefficients, prepare the NTFF          (allocateMemories,
 *            setCoefficient, ntffT?_init)
 *   update() one time step; the host advances time afterwards (simulator_calc)
 *   reset()  write "<angle>[deg].txt" and "<angle>[deg]_380nm_700nm_b.dat" into
 *            cwd, then zero all state
 *   finish() reset() + free
 *   getters  borrowed host pointers, k = i*N_PY + j, valid from init to finish
 *
 * What moved: the nine complex fields live in GPU memory; the 15 dense
 * coefficient arrays of the reference are twelve 1-D tables; update() is an
 * asynchronous launch; getters refresh a pinned host mirror on demand.
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

#define MPIFDTD_MAX_ANGLE_BATCH 4096
#define MPIFDTD_MAX_SLABS 16

typedef struct UpmlSolver {
  int kind;                       /* B200FDTD_TM_UPML or B200FDTD_TE_UPML          */
  b200fdtd_engine *engine;        /* slab 0: the whole grid unless n_slabs > 1      */
  b200fdtd_engine *slab[MPIFDTD_MAX_SLABS];   /* y-slabs, one engine each (slab[0] == engine) */
  int n_slabs;
  double *eps[3];                 /* host maps: TM EZ,HX,HY (HX/HY lazily) | TE EX,EY,HZ */
  dcomplex *mirror[3];            /* host mirrors for the X, Y, Z getters (lazy)    */
  int mirror_reads[3];            /* refreshes so far; pinned in place at the first   */
  size_t mirror_cells;
  int n_cell;
  int point_source;               /* opt-in, see mpifdtd_enablePointSource         */
  int source_form;                /* opt-in, see mpifdtd_setSourceForm             */
  int pending;                    /* update() calls not yet handed to the engine    */
  double pending_from;            /* field_getTime() of the first of them           */
  int defer;                      /* this solver instance may defer (see solver_update) */
  int n_batch;                    /* > 1: angle batch, see mpifdtd_setAngleBatch    */
  int batch_angles[MPIFDTD_MAX_ANGLE_BATCH];
  double *eps_ringed;             /* MPI-variant ids: eps map inside its ghost ring */
} UpmlSolver;

static UpmlSolver tm_solver = { .kind = B200FDTD_TM_UPML };
static UpmlSolver te_solver = { .kind = B200FDTD_TE_UPML };
/* The "MPI" solver ids (mpiTM_UPML.c, mpiTE_UPML.c) run here as rank 0 of 1: same
 * recurrences and coefficients, but E phase first, a CW source, every one of the N x N
 * cells updated against a zero ghost ring, and (N+2) x (N+2) arrays behind the getters
 * (mpiTM_UPML.c:196-217, 337-374, 674-716, 737-743).  Their per-step ntff() runs through the
 * same sample + deferred-projection kernels with its own plan (see solver_init); multi-GPU
 * decomposition is the y-slab path of mpifdtd_b200/slab.py instead. */
static UpmlSolver mpi_tm_solver = { .kind = B200FDTD_MPI_TM_UPML };
static UpmlSolver mpi_te_solver = { .kind = B200FDTD_MPI_TE_UPML };

static int is_mpi_kind(int kind) { return kind == B200FDTD_MPI_TM_UPML || kind == B200FDTD_MPI_TE_UPML; }
static int is_tm_kind(int kind)  { return kind == B200FDTD_TM_UPML || kind == B200FDTD_MPI_TM_UPML; }
static int point_source_requested;
static int source_form_requested;      /* MPIFDTD_SRC_* */
static int precision_requested;        /* B200FDTD_F64 / B200FDTD_F32 */
static int slabs_requested;            /* y-slabs (GPUs) of the next init(), 0 = ask MPIFDTD_DEVICES, else 1 */
static int batch_requested;            /* number of angles of the next init(), 0 = unbatched */
static int batch_angles_requested[MPIFDTD_MAX_ANGLE_BATCH];

static void fill_batch_source(int kind, double angle_deg, b200fdtd_batch_source *b);

/* MPIFDTD_TIMING=1: where init() and the far-field output spend their wall time (stderr) */
#include <time.h>
static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static void lap(const char *what, double *t)
{
  if (getenv("MPIFDTD_TIMING") == NULL) return;
  double n = now_s();
  fprintf(stderr, "[timing] %-28s %8.3f ms\n", what, 1e3 * (n - *t));
  *t = n;
}

static void die_on(int rc, const char *what)
{
  if (rc == B200FDTD_OK) return;
  printf("b200fdtd: %s failed (%d): %s\n", what, rc, b200fdtd_last_error());
  exit(2);
}

/* Opt-in CW point source at the domain centre (the reference's field_pointLight,
 * field.c:145-152, has no caller).  Exists so the NoModel configuration, whose
 * scattered-field sources are identically zero, has something to propagate. */
void mpifdtd_enablePointSource(int on) { point_source_requested = on; }

/* Opt-in source forms the reference carries but does not call:
 *   MPIFDTD_SRC_CW     serial TM UPML: field_scatteredWave(Ez, EPS_EZ, 0, 0) INSTEAD of the
 *                      pulse -- the commented alternative at fdtdTM_upml.c:62 (field.c:202-218)
 *   MPIFDTD_SRC_PLANE  planeWave(Ez, EPS_EZ) (mpiTM_UPML.c:377-403) IN ADDITION to the
 *                      solver's own source, as the commented call at mpiTM_UPML.c:204 would:
 *                      a line source on the grid row of the NTFF box's left edge.  Id 4 keeps
 *                      the local/global shift (local i = left is global left-1, all N columns);
 *                      the serial TM UPML solver, which has no ghost ring, drives global row
 *                      `left`, columns 1..N_PY-2.
 * Read at init(), like the point source. */
void mpifdtd_setSourceForm(int form) { source_form_requested = form; }

/* Optional single-precision path of the UPML solvers (ids 2-5): complex64 fields, f32
 * permittivity and coefficients on the GPU; getters still hand out double complex mirrors and
 * the NTFF history / far field stay double.  Read at init(); anything but 0/1 is exit(2). */
void mpifdtd_setPrecision(int precision)
{
  if (precision != B200FDTD_F64 && precision != B200FDTD_F32) {
    printf("mpifdtd_setPrecision: unknown precision %d\n", precision);
    exit(2);
  }
  precision_requested = precision;
}

/* Multi-GPU mode of the serial UPML solvers (ids 2, 3), host code staying C: the next init()
 * cuts the grid into n y-slabs, one engine per slab, slab g on CUDA device g modulo the number
 * of visible devices (so a one-GPU box can still run -- and test -- several slabs).  ONE process,
 * ONE host thread: update() issues the step on every engine, the slabs exchange their one-column
 * halos as direct stores into the neighbour's ghost columns over NVLink (CUDA peer access) and
 * order themselves with device-side flags; reset()/finish() sum the slabs' NTFF partial sums
 * onto slab 0 and write the files; the getters gather the slabs into the [N_PX][N_PY] mirror.
 * Replaces init_mpi + Connection_ISend_IRecvH/E of the MPI solvers (mpiTM_UPML.c:196-217,
 * 252-334, 718-748) and adds the far-field reduce they never do.
 * n <= 1 (default): one engine.  Environment MPIFDTD_DEVICES=n does the same for an unmodified
 * main.c.  Not combined with an angle batch (a batch already fills the GPU). */
void mpifdtd_setDevices(int n)
{
  if (n < 0 || n > MPIFDTD_MAX_SLABS) {
    printf("mpifdtd_setDevices: %d slabs (max %d)\n", n, MPIFDTD_MAX_SLABS);
    exit(2);
  }
  slabs_requested = n;
}

static int slabs_for_next_init(void)
{
  int n = slabs_requested;
  if (n == 0) {
    const char *v = getenv("MPIFDTD_DEVICES");
    if (v != NULL) n = atoi(v);
  }
  if (n > MPIFDTD_MAX_SLABS) n = MPIFDTD_MAX_SLABS;
  return n > 1 ? n : 1;
}

/* contiguous, near-equal column ranges; the first slabs take the remainder (slab.py agrees) */
static void slab_columns(int n_py, int n_slabs, int g, int *j0, int *nj)
{
  const int base = n_py / n_slabs, extra = n_py % n_slabs;
  *j0 = g * base + (g < extra ? g : extra);
  *nj = base + (g < extra ? 1 : 0);
}

/* Angle batch.  The reference sweeps incidence angles one simulation at a
 * time -- one angle per MPI rank, reset() + field_setWaveAngle() in between (main.c:114-138,
 * 183-211).  The simulations of a sweep share grid, permittivity and coefficients and differ
 * only in the source, so here the next init() of a serial UPML solver (ids 2, 3) can take ALL
 * angles at once: one batched engine, every update() advances every angle in the same two
 * kernel launches, and reset()/finish() write each angle's "<ang>[deg]..." files.
 * mpifdtd_selectAngle(k) picks the simulation the getters show.  n = 0 switches it off. */
void mpifdtd_setAngleBatch(const int *angles_deg, int n)
{
  if (n < 0 || n > MPIFDTD_MAX_ANGLE_BATCH || (n > 0 && angles_deg == NULL)) {
    printf("mpifdtd_setAngleBatch: bad batch of %d angles (max %d)\n", n, MPIFDTD_MAX_ANGLE_BATCH);
    exit(2);
  }
  batch_requested = n;
  for (int k = 0; k < n; k++) batch_angles_requested[k] = angles_deg[k];
}

/* the batch the next init() takes (split_shim.c reads it for ids 0, 1, 6, 7) */
int mpifdtd_angle_batch_requested(const int **angles_deg)
{
  if (angles_deg != NULL) *angles_deg = batch_angles_requested;
  return batch_requested;
}

/* ---- coefficient tables ------------------------------------------------------
 * Same expressions as setCoefficient (fdtdTM_upml.c:230-271, fdtdTE_upml.c:367-409)
 * evaluated once per row / column instead of once per cell.  sigma_z = 0 and
 * eps = EPSILON_0_S in every coefficient, which is what makes them separable. */
static void build_tables(const UpmlSolver *s, double *ti, double *tj)
{
  FieldInfo_S g = field_getFieldInfo_S();
  const double R = 1.0e-8, M = 2.0;
  const double eps = EPSILON_0_S, sig_z = 0;
  if (is_tm_kind(s->kind)) {
    const double sig_max = -(M + 1.0) * EPSILON_0_S * C_0_S / 2.0 / N_PML / cos(M_PI / 3) * log(R);
    for (int i = 0; i < g.N_PX; i++) {
      double sig_ez_x = sig_max * field_sigmaX(i, 0);
      double sig_hx_x = sig_max * field_sigmaX(i, 0.5);
      double sig_hy_x = sig_max * field_sigmaX(i + 0.5, 0);
      ti[B200FDTD_TMI_C_JZ * g.N_PX + i]     = (2 * eps - sig_ez_x) / (2 * eps + sig_ez_x);
      ti[B200FDTD_TMI_C_JZHXHY * g.N_PX + i] = (2 * eps) / (2 * eps + sig_ez_x);
      ti[B200FDTD_TMI_C_BXMX1 * g.N_PX + i]  = (2 * eps + sig_hx_x) / (2 * eps + sig_z);
      ti[B200FDTD_TMI_C_BXMX0 * g.N_PX + i]  = (2 * eps - sig_hx_x) / (2 * eps + sig_z);
      ti[B200FDTD_TMI_C_BY * g.N_PX + i]     = (2 * eps - sig_hy_x) / (2 * eps + sig_hy_x);
      ti[B200FDTD_TMI_DEN_BYMY * g.N_PX + i] = (2 * eps + sig_hy_x);
    }
    for (int j = 0; j < g.N_PY; j++) {
      double sig_ez_y = sig_max * field_sigmaY(0, j);
      double sig_hx_y = sig_max * field_sigmaY(0, j + 0.5);
      double sig_hy_y = sig_max * field_sigmaY(0.5, j);
      tj[B200FDTD_TMJ_C_DZ * g.N_PY + j]      = (2 * eps - sig_ez_y) / (2 * eps + sig_ez_y);
      tj[B200FDTD_TMJ_C_DZJZ * g.N_PY + j]    = (2 * eps + sig_z) / (2 * eps + sig_ez_y);
      tj[B200FDTD_TMJ_C_MX * g.N_PY + j]      = (2 * eps - sig_hx_y) / (2 * eps + sig_hx_y);
      tj[B200FDTD_TMJ_C_MXEZ * g.N_PY + j]    = (2 * eps) / (2 * eps + sig_hx_y);
      tj[B200FDTD_TMJ_NUM_BYMY1 * g.N_PY + j] = (2 * eps + sig_hy_y);
      tj[B200FDTD_TMJ_NUM_BYMY0 * g.N_PY + j] = (2 * eps - sig_hy_y);
    }
  } else {
    /* no cos(pi/3) factor for TE (fdtdTE_upml.c:369) */
    const double sig_max = -(M + 1.0) * EPSILON_0_S * LIGHT_SPEED_S / 2.0 / N_PML * log(R);
    for (int i = 0; i < g.N_PX; i++) {
      double sig_ex_x = sig_max * field_sigmaX(i + 0.5, 0);
      double sig_ey_x = sig_max * field_sigmaX(i, 0.5);
      double sig_hz_x = sig_max * field_sigmaX(i + 0.5, 0.5);
      ti[B200FDTD_TEI_C_DXJX1 * g.N_PX + i]  = (2 * eps + sig_ex_x) / (2 * eps + sig_z);
      ti[B200FDTD_TEI_C_DXJX0 * g.N_PX + i]  = (2 * eps - sig_ex_x) / (2 * eps + sig_z);
      ti[B200FDTD_TEI_C_DY * g.N_PX + i]     = (2 * eps - sig_ey_x) / (2 * eps + sig_ey_x);
      ti[B200FDTD_TEI_DEN_DYJY * g.N_PX + i] = (2 * eps + sig_ey_x);
      ti[B200FDTD_TEI_C_MZ * g.N_PX + i]     = (2 * eps - sig_hz_x) / (2 * eps + sig_hz_x);
      ti[B200FDTD_TEI_C_MZEXEY * g.N_PX + i] = (2 * eps) / (2 * eps + sig_hz_x);
    }
    for (int j = 0; j < g.N_PY; j++) {
      double sig_ex_y = sig_max * field_sigmaY(0.5, j);
      double sig_ey_y = sig_max * field_sigmaY(0, j + 0.5);
      double sig_hz_y = sig_max * field_sigmaY(0.5, j + 0.5);
      tj[B200FDTD_TEJ_C_JX * g.N_PY + j]      = (2 * eps - sig_ex_y) / (2 * eps + sig_ex_y);
      tj[B200FDTD_TEJ_C_JXHZ * g.N_PY + j]    = (2 * eps) / (2 * eps + sig_ex_y);
      tj[B200FDTD_TEJ_NUM_DYJY1 * g.N_PY + j] = (2 * eps + sig_ey_y);
      tj[B200FDTD_TEJ_NUM_DYJY0 * g.N_PY + j] = (2 * eps - sig_ey_y);
      tj[B200FDTD_TEJ_C_BZ * g.N_PY + j]      = (2 * eps - sig_hz_y) / (2 * eps + sig_hz_y);
      tj[B200FDTD_TEJ_C_BZMZ * g.N_PY + j]    = (2 * eps + sig_z) / (2 * eps + sig_hz_y);
    }
  }
}

void mpifdtd_upml_tables(int kind, double *tab_i, double *tab_j)
{
  UpmlSolver probe = { .kind = kind };
  build_tables(&probe, tab_i, tab_j);
}

/* Exposed for the parity tests: the 1-D tables expanded back to the reference's
 * dense coefficient (name as in fdtdTM_upml.c:30-35 / fdtdTE_upml.c:27-32). */
int mpifdtd_upml_dense_coefficient(int kind, const char *name, double *dst)
{
  FieldInfo_S g = field_getFieldInfo_S();
  UpmlSolver probe = { .kind = kind };
  double *ti = (double *)malloc(sizeof(double) * B200FDTD_UPML_TABS * g.N_PX);
  double *tj = (double *)malloc(sizeof(double) * B200FDTD_UPML_TABS * g.N_PY);
  build_tables(&probe, ti, tj);
  /* each dense array = I[i] (by-i), J[j] (by-j), constant 1, or J[j] / I[i] */
  struct { const char *name; int kind, by_i, by_j; } map[] = {
    { "C_JZ", B200FDTD_TM_UPML, B200FDTD_TMI_C_JZ, -1 },
    { "C_JZHXHY", B200FDTD_TM_UPML, B200FDTD_TMI_C_JZHXHY, -1 },
    { "C_BXMX1", B200FDTD_TM_UPML, B200FDTD_TMI_C_BXMX1, -1 },
    { "C_BXMX0", B200FDTD_TM_UPML, B200FDTD_TMI_C_BXMX0, -1 },
    { "C_BY", B200FDTD_TM_UPML, B200FDTD_TMI_C_BY, -1 },
    { "C_DZ", B200FDTD_TM_UPML, -1, B200FDTD_TMJ_C_DZ },
    { "C_DZJZ1", B200FDTD_TM_UPML, -1, B200FDTD_TMJ_C_DZJZ },
    { "C_DZJZ0", B200FDTD_TM_UPML, -1, B200FDTD_TMJ_C_DZJZ },
    { "C_MX", B200FDTD_TM_UPML, -1, B200FDTD_TMJ_C_MX },
    { "C_MXEZ", B200FDTD_TM_UPML, -1, B200FDTD_TMJ_C_MXEZ },
    { "C_BYMY1", B200FDTD_TM_UPML, B200FDTD_TMI_DEN_BYMY, B200FDTD_TMJ_NUM_BYMY1 },
    { "C_BYMY0", B200FDTD_TM_UPML, B200FDTD_TMI_DEN_BYMY, B200FDTD_TMJ_NUM_BYMY0 },
    { "C_BX", B200FDTD_TM_UPML, -1, -1 }, { "C_MY", B200FDTD_TM_UPML, -1, -1 },
    { "C_MYEZ", B200FDTD_TM_UPML, -1, -1 },
    { "C_DXJX1", B200FDTD_TE_UPML, B200FDTD_TEI_C_DXJX1, -1 },
    { "C_DXJX0", B200FDTD_TE_UPML, B200FDTD_TEI_C_DXJX0, -1 },
    { "C_DY", B200FDTD_TE_UPML, B200FDTD_TEI_C_DY, -1 },
    { "C_MZ", B200FDTD_TE_UPML, B200FDTD_TEI_C_MZ, -1 },
    { "C_MZEXEY", B200FDTD_TE_UPML, B200FDTD_TEI_C_MZEXEY, -1 },
    { "C_JX", B200FDTD_TE_UPML, -1, B200FDTD_TEJ_C_JX },
    { "C_JXHZ", B200FDTD_TE_UPML, -1, B200FDTD_TEJ_C_JXHZ },
    { "C_BZ", B200FDTD_TE_UPML, -1, B200FDTD_TEJ_C_BZ },
    { "C_BZMZ1", B200FDTD_TE_UPML, -1, B200FDTD_TEJ_C_BZMZ },
    { "C_BZMZ0", B200FDTD_TE_UPML, -1, B200FDTD_TEJ_C_BZMZ },
    { "C_DYJY1", B200FDTD_TE_UPML, B200FDTD_TEI_DEN_DYJY, B200FDTD_TEJ_NUM_DYJY1 },
    { "C_DYJY0", B200FDTD_TE_UPML, B200FDTD_TEI_DEN_DYJY, B200FDTD_TEJ_NUM_DYJY0 },
    { "C_DX", B200FDTD_TE_UPML, -1, -1 }, { "C_JY", B200FDTD_TE_UPML, -1, -1 },
    { "C_JYHZ", B200FDTD_TE_UPML, -1, -1 },
  };
  int found = 0;
  for (size_t m = 0; m < sizeof map / sizeof map[0] && !found; m++) {
    if (map[m].kind != kind || strcmp(map[m].name, name) != 0) continue;
    found = 1;
    for (int i = 0; i < g.N_PX; i++)
      for (int j = 0; j < g.N_PY; j++) {
        double v = 1.0;
        if (map[m].by_i >= 0 && map[m].by_j >= 0)
          v = tj[map[m].by_j * g.N_PY + j] / ti[map[m].by_i * g.N_PX + i];
        else if (map[m].by_i >= 0) v = ti[map[m].by_i * g.N_PX + i];
        else if (map[m].by_j >= 0) v = tj[map[m].by_j * g.N_PY + j];
        dst[(size_t)i * g.N_PY + j] = v;
      }
  }
  free(ti); free(tj);
  return found ? 0 : -1;
}

/* ---- init -------------------------------------------------------------------- */
static void solver_init(UpmlSolver *s)
{
  FieldInfo_S g = field_getFieldInfo_S();
  NTFFInfo box = field_getNTFFInfo();
  const int tm = is_tm_kind(s->kind);
  const int mpi = is_mpi_kind(s->kind);
  const int ring = mpi ? 1 : 0;                               /* ghost ring of the getters' arrays */
  const size_t sub_cells = (size_t)(g.N_PX + 2 * ring) * (size_t)(g.N_PY + 2 * ring);
  s->n_cell = g.N_CELL;
  s->point_source = point_source_requested;
  s->source_form = source_form_requested;
  s->pending = 0;
  {
    const char *v = getenv("MPIFDTD_DEFER_STEPS");
    s->defer = !mpi && !s->point_source && s->source_form == MPIFDTD_SRC_DEFAULT && !(v != NULL && v[0] == '0');
  }
  s->n_batch = 0;
  if (batch_requested > 0 && !mpi) {
    s->n_batch = batch_requested;
    memcpy(s->batch_angles, batch_angles_requested, sizeof(int) * (size_t)batch_requested);
  }
  s->n_slabs = s->n_batch > 1 ? 1 : slabs_for_next_init();
  if (s->n_slabs > g.N_PY / 4) s->n_slabs = g.N_PY / 4 > 0 ? g.N_PY / 4 : 1;
  if (s->n_slabs > 1) s->defer = 0;     /* multi-step replay is a single-engine feature */

  b200fdtd_grid grid;
  memset(&grid, 0, sizeof grid);
  grid.kind = s->kind;
  grid.n_px = g.N_PX;  grid.n_py = g.N_PY;  grid.n_pml = g.N_PML;
  grid.j0 = 0;         grid.nj = g.N_PY;
  grid.i_lo = 1;       grid.i_hi = g.N_PX - 2;      /* fdtdTM_upml.c:158-159 */
  grid.j_lo = 1;       grid.j_hi = g.N_PY - 2;
  if (mpi) {                                        /* local 1..SUB_N_PX-2 <-> global 0..N-1 */
    grid.i_lo = 0;     grid.i_hi = g.N_PX - 1;
    grid.j_lo = 0;     grid.j_hi = g.N_PY - 1;
  }
  grid.device = -1;
  grid.precision = precision_requested;
  grid.mu0 = MU_0_S;
  grid.n_batch = s->n_batch;
  double t_lap = now_s();
  int n_dev = 1;
  if (s->n_slabs > 1) die_on(b200fdtd_device_count(&n_dev), "b200fdtd_device_count");
  for (int k = 0; k < s->n_slabs; k++) {
    if (s->n_slabs > 1) {
      int j0, nj;
      slab_columns(g.N_PY, s->n_slabs, k, &j0, &nj);
      grid.j0 = j0;  grid.nj = nj;  grid.device = k % n_dev;
    }
    die_on(b200fdtd_create(&grid, &s->slab[k]), "b200fdtd_create");
  }
  s->engine = s->slab[0];
  for (int k = 0; k + 1 < s->n_slabs; k++) {        /* neighbours along y */
    die_on(b200fdtd_peer_attach_engine(s->slab[k], 1, s->slab[k + 1]), "b200fdtd_peer_attach_engine");
    die_on(b200fdtd_peer_attach_engine(s->slab[k + 1], 0, s->slab[k]), "b200fdtd_peer_attach_engine");
  }
  lap("init: engine create", &t_lap);
  if (s->n_batch > 1) {
    b200fdtd_batch_source *src = (b200fdtd_batch_source *)malloc(sizeof *src * (size_t)s->n_batch);
    for (int k = 0; k < s->n_batch; k++) fill_batch_source(s->kind, (double)s->batch_angles[k], &src[k]);
    die_on(b200fdtd_set_batch_sources(s->engine, src), "b200fdtd_set_batch_sources");
    free(src);
  }

  /* permittivity maps.  TM: EPS_EZ at (i,j) area-averaged; EPS_HX/EPS_HY are
   * computed upstream but never read (fdtdTM_upml.c:237-239), so they are not
   * built.  TE: EPS_EX at (i+1/2, j) averaged along y, EPS_EY at (i, j+1/2)
   * along x (fdtdTE_upml.c:374-375); EPS_HZ likewise unused. */
  const int n_eps = tm ? 1 : 2;
  for (int m = 0; m < n_eps; m++)
    die_on(b200fdtd_mirror_alloc((void **)&s->eps[m], sizeof(double) * (size_t)g.N_CELL), "mirror_alloc(eps)");   /* uploaded once: not worth pinning */
  if (tm) {
    mpifdtd_fill_eps(s->eps[0], 0, 0, D_XY);
  } else {
    mpifdtd_fill_eps(s->eps[0], 0.5, 0, D_Y);
    mpifdtd_fill_eps(s->eps[1], 0, 0.5, D_X);
  }
  lap("init: eps maps (host)", &t_lap);
  /* upload: as 16-bit indices into the table of the map's distinct values where there are at most
   * 65536 of them (2 instead of 8 bytes per cell over PCIe; MPIFDTD_EPS_DENSE=1 or a richer map: the
   * dense doubles); each engine takes its columns */
  {
    const char *dense_env = getenv("MPIFDTD_EPS_DENSE");
    const int want_palette = !(dense_env != NULL && dense_env[0] == '1');
    uint16_t *index = want_palette ? (uint16_t *)malloc(sizeof(uint16_t) * (size_t)g.N_CELL) : NULL;
    double *table = want_palette ? (double *)malloc(sizeof(double) * 65536) : NULL;
    for (int m = 0; m < n_eps; m++) {
      const int n_values = (index != NULL && table != NULL) ? mpifdtd_eps_palette(s->eps[m], (size_t)g.N_CELL, index, table) : -1;
      for (int k = 0; k < s->n_slabs; k++) {
        if (n_values > 0) {
          int j0 = 0, nj = g.N_PY;
          if (s->n_slabs > 1) slab_columns(g.N_PY, s->n_slabs, k, &j0, &nj);
          die_on(b200fdtd_set_eps_palette(s->slab[k], m, index + j0, g.N_PY, table, n_values), "b200fdtd_set_eps_palette");
        } else {
          die_on(b200fdtd_set_eps(s->slab[k], m, s->eps[m]), "b200fdtd_set_eps");
        }
      }
    }
    free(index); free(table);
  }
  lap("init: eps upload", &t_lap);

  double *ti = (double *)malloc(sizeof(double) * B200FDTD_UPML_TABS * g.N_PX);
  double *tj = (double *)malloc(sizeof(double) * B200FDTD_UPML_TABS * g.N_PY);
  build_tables(s, ti, tj);
  for (int k = 0; k < s->n_slabs; k++)
    die_on(b200fdtd_set_upml_tables(s->slab[k], ti, tj), "b200fdtd_set_upml_tables");
  free(ti); free(tj);

  s->mirror_cells = sub_cells;          /* the pinned mirrors are allocated by the first getter call */
  if (mpi) {                                        /* EPS array as the getter shows it: with the ring */
    const int m = tm ? 0 : 1;                       /* TM: EPS_EZ, TE: EPS_EY (mpiTE_UPML.c:103-106) */
    s->eps_ringed = newDouble((int)sub_cells);
    for (int i = 0; i < g.N_PX; i++)
      memcpy(s->eps_ringed + (size_t)(i + 1) * (g.N_PY + 2) + 1, s->eps[m] + (size_t)i * g.N_PY,
             sizeof(double) * (size_t)g.N_PY);
    /* ntff() of these solvers (mpiTM_UPML.c:849-1037, mpiTE_UPML.c:602-794): the serial
     * binning with three differences, all carried by the plan -- timeShift evaluated
     * directly per point, every tap multiplied by coef = 1/(4 pi C R), R = 1e6, and the box
     * indices used as LOCAL indices (local i is global i-1, so the cells read sit one cell
     * down-left of where r2 says). */
    b200fdtd_ntff_plan mp;
    memset(&mp, 0, sizeof mp);
    mp.top = box.top; mp.bottom = box.bottom; mp.left = box.left; mp.right = box.right;
    mp.n_points = mpifdtd_ntff_point_count(&box);
    mp.max_time = (int)field_getMaxTime();
    mp.n_bins = getenv("MPIFDTD_NTFF_FULL_BINS") ? box.arraySize : mp.max_time;
    mp.n_angles = N_ANGLES;
    mp.array_size = box.arraySize;
    mp.tap_scale = 1.0 / (4 * M_PI * C_0_S * 1.0e6);
    mp.sample_di = -1; mp.sample_dj = -1;
    /* Several slabs: ONE surface over the whole grid (the single-rank meaning of the reference's
     * ntff(); with more ranks upstream integrates every rank's own sub-grid over the same local
     * indices and never sums the pieces), each slab sampling the cells it owns. */
    for (int k = 0; k < s->n_slabs; k++) {
      int j0 = 0, nj = g.N_PY;
      if (s->n_slabs > 1) slab_columns(g.N_PY, s->n_slabs, k, &j0, &nj);
      mp.n_local = mpifdtd_ntff_local_count_shifted(&box, mp.sample_dj, j0, nj);
      double *direct = mpifdtd_ntff_time_shift_direct(&box, N_ANGLES, tm ? 0.0 : 0.5, mp.sample_dj, j0, nj);
      mp.time_shift = direct;
      if (mp.n_points > 0 && mp.max_time > 0)
        die_on(b200fdtd_set_ntff_plan(s->slab[k], &mp), "b200fdtd_set_ntff_plan");
      free(direct);
    }
    return;
  }

  /* ntffT?_init: surface, history length = stepNum, bins kept = the part of
   * arraySize the far field reads (ntffTM.c:181: i < maxTime) */
  b200fdtd_ntff_plan plan;
  memset(&plan, 0, sizeof plan);
  plan.top = box.top; plan.bottom = box.bottom; plan.left = box.left; plan.right = box.right;
  plan.n_points = mpifdtd_ntff_point_count(&box);
  plan.n_local = plan.n_points;
  plan.max_time = (int)field_getMaxTime();
  plan.n_bins = getenv("MPIFDTD_NTFF_FULL_BINS") ? box.arraySize : plan.max_time;
  plan.n_angles = N_ANGLES;
  plan.array_size = box.arraySize;
  lap("init: tables", &t_lap);
  for (int k = 0; k < s->n_slabs; k++) {            /* each slab samples the surface points it owns */
    int j0 = 0, nj = g.N_PY;
    if (s->n_slabs > 1) slab_columns(g.N_PY, s->n_slabs, k, &j0, &nj);
    plan.n_local = mpifdtd_ntff_local_count(&box, j0, nj);
    double *shift = mpifdtd_ntff_time_shift(&box, N_ANGLES, tm ? 0.0 : 0.5, j0, nj);
    plan.time_shift = shift;
    if (plan.n_points > 0 && plan.max_time > 0)
      die_on(b200fdtd_set_ntff_plan(s->slab[k], &plan), "b200fdtd_set_ntff_plan");
    free(shift);
  }
  lap("init: ntff plans", &t_lap);
}

/* ---- update ------------------------------------------------------------------ */
static double fill_pulse_at(b200fdtd_pulse *p, double angle_deg, double gap_x, double gap_y, double dot)
{
  FieldInfo_S g = field_getFieldInfo_S();
  double rad = angle_deg * M_PI / 180.0;                      /* field.c:228-241 */
  double cos_per_c = cos(rad) / C_0_S, sin_per_c = sin(rad) / C_0_S;
  const double center_peak = (g.N_PX / 2.0 + gap_x) * cos_per_c + (g.N_PY / 2 + gap_y) * sin_per_c;
  const double t0 = -center_peak + 500;
  p->enabled = 1;
  p->gap_x = gap_x;  p->gap_y = gap_y;  p->dot = dot;
  p->cos_per_c = cos_per_c;  p->sin_per_c = sin_per_c;
  p->time_minus_t0 = field_getTime() - t0;
  p->omega = field_getOmega();
  p->beam_width = 50;
  return t0;
}

static void fill_pulse(b200fdtd_pulse *p, double gap_x, double gap_y, double dot)
{
  fill_pulse_at(p, field_getWaveAngle(), gap_x, gap_y, dot);
}

/* The pulse(s) update() fires for one incidence angle (fdtdTM_upml.c:63, fdtdTE_upml.c:182-189),
 * as the per-simulation record of a batched engine. */
static void fill_batch_source(int kind, double angle_deg, b200fdtd_batch_source *b)
{
  memset(b, 0, sizeof *b);
  if (kind == B200FDTD_TM_UPML) {
    b->t0[0] = fill_pulse_at(&b->pulse[0], angle_deg, 0, 0, 1.0);
  } else {
    double co = cos((angle_deg + 90) * M_PI / 180.0);
    double si = sin((angle_deg + 90) * M_PI / 180.0);
    if (co != 0.0) b->t0[0] = fill_pulse_at(&b->pulse[0], angle_deg, 0.5, 0.0, co);
    if (si != 0.0) b->t0[1] = fill_pulse_at(&b->pulse[1], angle_deg, 0.0, 0.5, si);
  }
}

/* test hook: the per-simulation pulse record of one incidence angle (b200fdtd_set_batch_sources) */
void mpifdtd_upml_batch_source(int kind, double angle_deg, b200fdtd_batch_source *out)
{
  fill_batch_source(kind, angle_deg, out);
}

/* Everything update() reads from the host's grid/time state, packed for the
 * engine.  Also used by slab (multi-GPU) drivers, which call the engine phases
 * themselves and then field_nextStep(). */
static void fill_plane_wave(int kind, b200fdtd_line_source *l)
{
  NTFFInfo box = field_getNTFFInfo();
  const double k_s = field_getK();
  const double rad = field_getWaveAngle() * M_PI / 180;        /* mpiTM_UPML.c:383 */
  l->enabled = 1;
  if (is_mpi_kind(kind)) { l->i = box.left - 1;  l->j_lo = 0;  l->j_hi = N_PY - 1; }
  else                   { l->i = box.left;      l->j_lo = 1;  l->j_hi = N_PY - 2; }
  l->scale = field_getRayCoef();
  l->ks_cos = cos(rad) * k_s;  l->ks_sin = sin(rad) * k_s;
  l->time = field_getTime();
  l->omega = field_getOmega();
}

void mpifdtd_upml_step_args_form(int kind, int point_source, int form, b200fdtd_step_args *a);
void mpifdtd_upml_step_args(int kind, int point_source, b200fdtd_step_args *a)
{
  mpifdtd_upml_step_args_form(kind, point_source, MPIFDTD_SRC_DEFAULT, a);
}

void mpifdtd_upml_step_args_form(int kind, int point_source, int form, b200fdtd_step_args *a)
{
  memset(a, 0, sizeof *a);
  a->time = field_getTime();
  a->ray_coef = field_getRayCoef();
  if (form == MPIFDTD_SRC_PLANE && is_tm_kind(kind)) fill_plane_wave(kind, &a->line);
  if (form == MPIFDTD_SRC_CW && kind == B200FDTD_TM_UPML) {
    /* field_scatteredWave (field.c:202-218): p += ray_coef*(eps0/eps - 1)*cexp(I*(kr - w t)) */
    b200fdtd_cw *c = &a->cw[0];
    double k_s = field_getK();
    double rad = field_getWaveAngle() * M_PI / 180.0;
    c->enabled = 1;
    c->scale = field_getRayCoef();
    c->ks_cos = cos(rad) * k_s;
    c->ks_sin = sin(rad) * k_s;
    c->phase_a = field_getOmega() * field_getTime();
    return;
  }
  if (is_mpi_kind(kind)) {
    /* CW scattered wave, integer global coordinates, on Ez (TM) or on Ey only (TE):
     * mpiTM_UPML.c:337-374, mpiTE_UPML.c:250-281 */
    b200fdtd_cw *c = &a->cw[kind == B200FDTD_MPI_TM_UPML ? 0 : 1];
    double k_s = field_getK();
    double rad = field_getWaveAngle() * M_PI / 180;
    c->enabled = 1;
    c->two_term = 0;
    c->scale = field_getRayCoef();
    c->ks_cos = cos(rad) * k_s;
    c->ks_sin = sin(rad) * k_s;
    c->phase_a = field_getOmega() * field_getTime();
    return;
  }
  if (kind == B200FDTD_TM_UPML) {
    fill_pulse(&a->pulse[0], 0, 0, 1.0);                      /* fdtdTM_upml.c:63 */
  } else {
    /* polarisation 90 degrees from the wave vector (fdtdTE_upml.c:182-189); both
     * components fire at 0 degrees because cos(90 deg) != 0 in floating point */
    WaveInfo_S w = field_getWaveInfo_S();
    double co = cos((w.Angle_deg + 90) * M_PI / 180.0);
    double si = sin((w.Angle_deg + 90) * M_PI / 180.0);
    if (co != 0.0) fill_pulse(&a->pulse[0], 0.5, 0.0, co);
    if (si != 0.0) fill_pulse(&a->pulse[1], 0.0, 0.5, si);
  }
  if (point_source) {
    dcomplex v = field_pointLight();
    a->point.enabled = 1;
    a->point.i = N_PX / 2;  a->point.j = N_PY / 2;
    a->point.re = creal(v); a->point.im = cimag(v);
  }
}

/* update() is asynchronous anyway, so for the serial solvers with their default pulse source
 * it only COUNTS: the pending steps go to the engine in one b200fdtd_run_steps call -- a CUDA
 * graph replay with the time read from a device-side clock -- when somebody looks (a getter,
 * reset/finish, the engine handle) or MPIFDTD_DEFER_CHUNK (256) steps have piled up.  Results
 * are bit-identical to stepping one by one; small grids, where a step is a few microseconds of
 * device work, run 2-3x faster.  MPIFDTD_DEFER_STEPS=0 hands every step over immediately. */
static int defer_chunk(void)
{
  const char *v = getenv("MPIFDTD_DEFER_CHUNK");
  int n = v != NULL ? atoi(v) : 256;
  return n > 0 ? n : 256;
}

static void flush_pending(UpmlSolver *s)
{
  if (s->engine == NULL || s->pending == 0) return;
  if (s->n_batch <= 1) {                 /* the pulse record of the current incidence angle */
    b200fdtd_batch_source one;
    fill_batch_source(s->kind, field_getWaveAngle(), &one);
    die_on(b200fdtd_set_batch_sources(s->engine, &one), "b200fdtd_set_batch_sources");
  }
  const int n = s->pending;
  s->pending = 0;
  die_on(b200fdtd_run_steps(s->engine, s->pending_from, n), "b200fdtd_run_steps");
}

void mpifdtd_flush_pending_steps(void)
{
  flush_pending(&tm_solver);
  flush_pending(&te_solver);
  mpifdtd_split_flush_pending_steps();
}

static void solver_update(UpmlSolver *s)
{
  if (s->defer) {
    if (s->pending == 0) s->pending_from = field_getTime();
    if (++s->pending >= defer_chunk()) flush_pending(s);
    return;
  }
  b200fdtd_step_args a;
  mpifdtd_upml_step_args_form(s->kind, s->point_source, s->source_form, &a);
  /* asynchronous launches: with several slabs the engines order themselves across GPUs through
   * the peer-halo flags, the host never waits */
  for (int k = 0; k < s->n_slabs; k++)
    die_on(b200fdtd_step(s->slab[k], &a), "b200fdtd_step");
}

/* ---- reset / finish ------------------------------------------------------------ */
/* ntffT?_TimeOutput (ntffTM.c:197-275, ntffTE.c:160-238): project the recorded
 * surface history, translate + FFT + interpolate on the GPU, return the
 * 321 x 360 table.  Exposed so slab drivers can call it on the reduced U/W. */
void mpifdtd_upml_far_field(b200fdtd_engine *engine, int kind, int project, double *table)
{
  const int tm = (kind == B200FDTD_TM_UPML);
  double cos_phi[N_ANGLES], sin_phi[N_ANGLES];
  mpifdtd_ntff_direction_cosines(N_ANGLES, tm, cos_phi, sin_phi);
  double complex coef = mpifdtd_ntff_translate_coef(field_getOmega());
  double complex *tw = mpifdtd_fft_twiddles(NTFF_NUM);
  FieldInfo phys = field_getFieldInfo();

  b200fdtd_spectrum_args sa;
  memset(&sa, 0, sizeof sa);
  sa.coef_re = creal(coef);  sa.coef_im = cimag(coef);
  sa.z0 = Z_0_S;
  sa.cos_phi = cos_phi;  sa.sin_phi = sin_phi;
  sa.n_fft = NTFF_NUM;
  sa.lambda_first_nm = LAMBDA_ST_NM;  sa.lambda_last_nm = LAMBDA_EN_NM;
  sa.c_hu_nfft = C_0_S * phys.h_u_nm * NTFF_NUM;              /* ntffTM.c:227 */
  sa.twiddle = (const double *)tw;
  if (project) {
    double t_lap = now_s();
    die_on(b200fdtd_ntff_project(engine), "b200fdtd_ntff_project");
    if (getenv("MPIFDTD_TIMING")) { die_on(b200fdtd_sync(engine), "b200fdtd_sync"); lap("finish: ntff projection", &t_lap); }
  }
  die_on(b200fdtd_ntff_spectrum(engine, &sa, table), "b200fdtd_ntff_spectrum");
  free(tw);
}

static void write_far_field(UpmlSolver *s)
{
  if (field_getMaxTime() < 1) return;
  const int rows = LAMBDA_EN_NM - LAMBDA_ST_NM + 1;
  double *table = (double *)malloc(sizeof(double) * (size_t)rows * N_ANGLES);
  double **by_row = (double **)malloc(sizeof(double *) * (size_t)rows);
  for (int r = 0; r < rows; r++) by_row[r] = table + (size_t)r * N_ANGLES;
  char name[256], cwd[512];
  if (getcwd(cwd, sizeof cwd) == NULL) cwd[0] = '\0';

  /* an angle batch writes one pair of files per simulation, named by its own angle; the
   * projection kernel turns every simulation's history into U/W in one launch */
  const int n = s->n_batch > 1 ? s->n_batch : 1;
  flush_pending(s);
  double t_lap = now_s(), t_gpu = 0, t_txt = 0, t_bin = 0;
  for (int k = 0; k < s->n_slabs; k++) die_on(b200fdtd_sync(s->slab[k]), "b200fdtd_sync");
  lap("finish: drain the stepping", &t_lap);
  if (s->n_slabs > 1) {                 /* partial sums of the slabs -> slab 0, in slab order */
    for (int k = 0; k < s->n_slabs; k++) die_on(b200fdtd_ntff_project(s->slab[k]), "b200fdtd_ntff_project");
    for (int k = 1; k < s->n_slabs; k++) die_on(b200fdtd_ntff_add_uw(s->slab[0], s->slab[k]), "b200fdtd_ntff_add_uw");
    lap("finish: ntff projection + sum over slabs", &t_lap);
  }
  for (int k = 0; k < n; k++) {
    const int angle = s->n_batch > 1 ? s->batch_angles[k] : (int)field_getWaveAngle();
    if (s->n_batch > 1) die_on(b200fdtd_select_batch(s->engine, k), "b200fdtd_select_batch");
    double t0 = now_s();
    mpifdtd_upml_far_field(s->engine, s->kind, k == 0 && s->n_slabs == 1, table);
    double t1 = now_s();
    sprintf(name, "%d[deg].txt", angle);
    ntff_outputEnormTxt(by_row, name);
    printf("saved %s/%s\n", cwd, name);
    double t2 = now_s();
    sprintf(name, "%d[deg]_%dnm_%dnm_b.dat", angle, LAMBDA_ST_NM, LAMBDA_EN_NM);
    ntff_outputEnormBin(by_row, name);
    printf("saved %s/%s\n", cwd, name);
    t_gpu += t1 - t0; t_txt += t2 - t1; t_bin += now_s() - t2;
  }
  if (getenv("MPIFDTD_TIMING"))
    fprintf(stderr, "[timing] finish: project+spectrum %.3f ms, .txt %.3f ms, .dat %.3f ms (%d simulation(s))\n",
            1e3 * t_gpu, 1e3 * t_txt, 1e3 * t_bin, n);
  if (s->n_batch > 1) die_on(b200fdtd_select_batch(s->engine, 0), "b200fdtd_select_batch");
  free(by_row); free(table);
}

/* which simulation of an angle batch the getters (and mpifdtd_upml_far_field) refer to */
void mpifdtd_selectAngle(int index)
{
  UpmlSolver *all[2] = { &tm_solver, &te_solver };
  for (int m = 0; m < 2; m++)
    if (all[m]->engine != NULL && all[m]->n_batch > 1)
      die_on(b200fdtd_select_batch(all[m]->engine, index), "b200fdtd_select_batch");
  mpifdtd_split_select_angle(index);
}

/* ntffOutput + ntffSaveData of the MPI TE solver (mpiTE_UPML.c:795-878): translate the first
 * maxTime bins of Wx, Wy, Uz into E_theta / E_phi at theta = 0 and print them, one line per
 * direction, "%.20lf " per value, into MPI_TE_UPML/E{ph,th}_{r,i}.txt (upstream requires the
 * directory to exist and exit(2)s otherwise; here it is created when missing).  The accumulation ran
 * on the GPU; this is 360 x maxTime values of closing algebra and formatting.  (Upstream then
 * walks its debug arrays, which are NULL without -DDEBUG, and crashes; that is not kept.) */
static FILE *open_in_mpi_te_dir(const char *file_name)
{
  char name[256];
  sprintf(name, "MPI_TE_UPML/%s", file_name);
  makeDirectory("MPI_TE_UPML");     /* upstream expects it to exist; a missing one costs the run */
  FILE *fp = fopen(name, "w");
  if (fp == NULL) { printf("cannot open file %s \n", name); exit(2); }
  return fp;
}

static void save_mpi_ntff_series(const char *stem, const dcomplex *data, int max_time)
{
  char real_file[256], imag_file[256];
  sprintf(real_file, "%s_r.txt", stem);
  sprintf(imag_file, "%s_i.txt", stem);
  FILE *fr = open_in_mpi_te_dir(real_file), *fi = open_in_mpi_te_dir(imag_file);
  for (int ang = 0; ang < N_ANGLES; ang++) {
    for (int i = 0; i < max_time; i++) {
      fprintf(fr, "%.20lf ", creal(data[(size_t)ang * max_time + i]));
      fprintf(fi, "%.20lf ", cimag(data[(size_t)ang * max_time + i]));
    }
    fprintf(fr, "\n");
    fprintf(fi, "\n");
  }
  fclose(fr); fclose(fi);
  printf("saved at %s & %s\n", real_file, imag_file);
}

/* E_theta, E_phi [360][maxTime] of the MPI TE solver from the GPU's U/W (exposed for tests) */
int mpifdtd_mpi_te_far_series(dcomplex *eth, dcomplex *eph)
{
  UpmlSolver *s = &mpi_te_solver;
  const int max_time = (int)field_getMaxTime();
  if (s->engine == NULL || max_time < 1) return -1;
  const int n_bins = getenv("MPIFDTD_NTFF_FULL_BINS") ? field_getNTFFInfo().arraySize : max_time;
  const size_t count = (size_t)N_ANGLES * (size_t)n_bins;
  dcomplex *uw[3];
  for (int k = 0; k < s->n_slabs; k++) die_on(b200fdtd_sync(s->slab[k]), "b200fdtd_sync");
  for (int k = 0; k < s->n_slabs; k++) die_on(b200fdtd_ntff_project(s->slab[k]), "b200fdtd_ntff_project");
  for (int k = 1; k < s->n_slabs; k++)   /* partial sums of the slabs -> slab 0, in slab order */
    die_on(b200fdtd_ntff_add_uw(s->slab[0], s->slab[k]), "b200fdtd_ntff_add_uw");
  for (int m = 0; m < 3; m++) {
    uw[m] = (dcomplex *)malloc(sizeof(dcomplex) * count);
    die_on(b200fdtd_ntff_get_uw(s->engine, m, (double *)uw[m]), "b200fdtd_ntff_get_uw");
  }
  const double w_s = field_getOmega();
  const double complex coef = csqrt(2 * M_PI * C_0_S / (I * w_s));        /* mpiTE_UPML.c:844 */
  const double theta = 0;
  for (int ang = 0; ang < N_ANGLES; ang++) {
    double phi = ang * M_PI / 180.0;
    double sx = cos(theta) * cos(phi), sy = cos(theta) * sin(phi), sz = -cos(theta);
    double px = -sin(phi), py = cos(phi);
    const dcomplex *Wx = uw[0] + (size_t)ang * n_bins, *Wy = uw[1] + (size_t)ang * n_bins;
    const dcomplex *Uz = uw[2] + (size_t)ang * n_bins;
    for (int i = 0; i < max_time; i++) {
      double complex WTH = Wx[i] * sx + Wy[i] * sy + 0;
      double complex WPH = Wx[i] * px + Wy[i] * py;
      double complex UTH = 0 + 0 + Uz[i] * sz;
      double complex UPH = 0 + 0;
      eth[(size_t)ang * max_time + i] = coef * (-Z_0_S * WTH - UPH);
      eph[(size_t)ang * max_time + i] = coef * (-Z_0_S * WPH + UTH);
    }
  }
  for (int m = 0; m < 3; m++) free(uw[m]);
  return 0;
}

static void write_mpi_te_far_field(void)
{
  const int max_time = (int)field_getMaxTime();
  if (max_time < 1) return;
  const size_t count = (size_t)N_ANGLES * (size_t)max_time;
  dcomplex *eth = (dcomplex *)malloc(sizeof(dcomplex) * count);
  dcomplex *eph = (dcomplex *)malloc(sizeof(dcomplex) * count);
  if (mpifdtd_mpi_te_far_series(eth, eph) == 0) {
    save_mpi_ntff_series("Eph", eph, max_time);
    save_mpi_ntff_series("Eth", eth, max_time);
  }
  free(eph); free(eth);
}

static void solver_reset(UpmlSolver *s)
{
  if (s->engine == NULL) return;
  flush_pending(s);
  if (!is_mpi_kind(s->kind))                        /* mpiTM_UPML.c:219-232: reset only zeroes */
    write_far_field(s);
  /* nobody may be in mid-step while a neighbour's flags are zeroed */
  for (int k = 0; k < s->n_slabs; k++) die_on(b200fdtd_sync(s->slab[k]), "b200fdtd_sync");
  for (int k = 0; k < s->n_slabs; k++) die_on(b200fdtd_zero_state(s->slab[k]), "b200fdtd_zero_state");
}

static void solver_finish(UpmlSolver *s)
{
  if (s->engine == NULL) return;
  /* mpiTE_UPML.c:304-317: finish() = ntffOutput + free (no reset); the TM twin's ntffOutput
   * call is commented out (mpiTM_UPML.c:240), so id 4 writes nothing */
  if (s->kind == B200FDTD_MPI_TE_UPML) write_mpi_te_far_field();
  solver_reset(s);
  for (int k = 0; k < s->n_slabs; k++) {
    die_on(b200fdtd_destroy(s->slab[k]), "b200fdtd_destroy");
    s->slab[k] = NULL;
  }
  s->engine = NULL;
  s->n_slabs = 0;
  free(s->eps_ringed); s->eps_ringed = NULL;
  for (int m = 0; m < 3; m++) {
    b200fdtd_mirror_free(s->eps[m], 0); s->eps[m] = NULL;
    b200fdtd_mirror_free(s->mirror[m], s->mirror_reads[m] >= 1);
    s->mirror[m] = NULL;  s->mirror_reads[m] = 0;
  }
}

static dcomplex *solver_field(UpmlSolver *s, int mirror, int slot)
{
  if (s->engine == NULL) return NULL;                         /* upstream returns its NULL static */
  flush_pending(s);
  if (s->mirror[mirror] == NULL) {      /* only if somebody looks */
    die_on(b200fdtd_mirror_alloc((void **)&s->mirror[mirror], sizeof(dcomplex) * s->mirror_cells), "mirror_alloc");
    s->mirror_reads[mirror] = 0;
  }
  /* a huge-page mapping, pinned in place the first time somebody looks (b200fdtd_mirror_alloc) */
  if (++s->mirror_reads[mirror] == 1)
    die_on(b200fdtd_mirror_pin(s->mirror[mirror], sizeof(dcomplex) * s->mirror_cells), "mirror_pin");
  if (is_mpi_kind(s->kind)) {                                 /* (N+2) x (N+2) with a zero ring */
    FieldInfo_S g = field_getFieldInfo_S();
    for (int k = 0; k < s->n_slabs; k++) {
      int j0 = 0, nj = g.N_PY;
      if (s->n_slabs > 1) slab_columns(g.N_PY, s->n_slabs, k, &j0, &nj);
      double *first = (double *)(s->mirror[mirror] + (size_t)(g.N_PY + 2) + 1 + j0);
      die_on(b200fdtd_get_field_ld(s->slab[k], slot, first, g.N_PY + 2), "b200fdtd_get_field_ld");
    }
    return s->mirror[mirror];
  }
  for (int k = 0; k < s->n_slabs; k++)               /* every slab lands in its own columns of the mirror */
    die_on(b200fdtd_get_field(s->slab[k], slot, (double *)s->mirror[mirror]), "b200fdtd_get_field");
  return s->mirror[mirror];
}

/* ---- the exported entry points -------------------------------------------------- */
static void tm_update(void) { solver_update(&tm_solver); }
static void tm_init(void)   { solver_init(&tm_solver); }
static void tm_reset(void)  { solver_reset(&tm_solver); }
static void tm_finish(void) { solver_finish(&tm_solver); }
void (*fdtdTM_upml_getUpdate(void))(void) { return tm_update; }
void (*fdtdTM_upml_getInit(void))(void)   { return tm_init; }
void (*fdtdTM_upml_getReset(void))(void)  { return tm_reset; }
void (*fdtdTM_upml_getFinish(void))(void) { return tm_finish; }
double complex *fdtdTM_upml_getHx(void) { return solver_field(&tm_solver, 0, B200FDTD_TM_HX); }
double complex *fdtdTM_upml_getHy(void) { return solver_field(&tm_solver, 1, B200FDTD_TM_HY); }
double complex *fdtdTM_upml_getEz(void) { return solver_field(&tm_solver, 2, B200FDTD_TM_EZ); }
double *fdtdTM_upml_getEps(void) { return tm_solver.eps[0]; }              /* EPS_EZ */

static void te_update(void) { solver_update(&te_solver); }
static void te_init(void)   { solver_init(&te_solver); }
static void te_reset(void)  { solver_reset(&te_solver); }
static void te_finish(void) { solver_finish(&te_solver); }
void (*fdtdTE_upml_getUpdate(void))(void) { return te_update; }
void (*fdtdTE_upml_getInit(void))(void)   { return te_init; }
void (*fdtdTE_upml_getReset(void))(void)  { return te_reset; }
void (*fdtdTE_upml_getFinish(void))(void) { return te_finish; }
double complex *fdtdTE_upml_getEx(void) { return solver_field(&te_solver, 0, B200FDTD_TE_EX); }
double complex *fdtdTE_upml_getEy(void) { return solver_field(&te_solver, 1, B200FDTD_TE_EY); }
double complex *fdtdTE_upml_getHz(void) { return solver_field(&te_solver, 2, B200FDTD_TE_HZ); }
double *fdtdTE_upml_getEps(void) { return te_solver.eps[0]; }              /* EPS_EX, fdtdTE_upml.c:93-96 */

/* ---- the MPI-variant ids (mpiTM_UPML.h:5-18, mpiTE_UPML.h) -------------------------- */
static void mtm_update(void) { solver_update(&mpi_tm_solver); }
static void mtm_init(void)   { solver_init(&mpi_tm_solver); }
static void mtm_reset(void)  { solver_reset(&mpi_tm_solver); }
static void mtm_finish(void) { solver_finish(&mpi_tm_solver); }
void (*mpi_fdtdTM_upml_getUpdate(void))(void) { return mtm_update; }
void (*mpi_fdtdTM_upml_getInit(void))(void)   { return mtm_init; }
void (*mpi_fdtdTM_upml_getReset(void))(void)  { return mtm_reset; }
void (*mpi_fdtdTM_upml_getFinish(void))(void) { return mtm_finish; }
double complex *mpi_fdtdTM_upml_getHx(void) { return solver_field(&mpi_tm_solver, 0, B200FDTD_TM_HX); }
double complex *mpi_fdtdTM_upml_getHy(void) { return solver_field(&mpi_tm_solver, 1, B200FDTD_TM_HY); }
double complex *mpi_fdtdTM_upml_getEz(void) { return solver_field(&mpi_tm_solver, 2, B200FDTD_TM_EZ); }
double *mpi_fdtdTM_upml_getEps(void) { return mpi_tm_solver.eps_ringed; }
int mpi_fdtdTM_upml_getSubNx(void)    { return N_PX; }
int mpi_fdtdTM_upml_getSubNy(void)    { return N_PY; }
int mpi_fdtdTM_upml_getSubNpx(void)   { return N_PX + 2; }
int mpi_fdtdTM_upml_getSubNpy(void)   { return N_PY + 2; }
int mpi_fdtdTM_upml_getSubNcell(void) { return (N_PX + 2) * (N_PY + 2); }

static void mte_update(void) { solver_update(&mpi_te_solver); }
static void mte_init(void)   { solver_init(&mpi_te_solver); }
static void mte_reset(void)  { solver_reset(&mpi_te_solver); }
static void mte_finish(void) { solver_finish(&mpi_te_solver); }
void (*mpi_fdtdTE_upml_getUpdate(void))(void) { return mte_update; }
void (*mpi_fdtdTE_upml_getInit(void))(void)   { return mte_init; }
void (*mpi_fdtdTE_upml_getReset(void))(void)  { return mte_reset; }
void (*mpi_fdtdTE_upml_getFinish(void))(void) { return mte_finish; }
double complex *mpi_fdtdTE_upml_getEx(void) { return solver_field(&mpi_te_solver, 0, B200FDTD_TE_EX); }
double complex *mpi_fdtdTE_upml_getEy(void) { return solver_field(&mpi_te_solver, 1, B200FDTD_TE_EY); }
double complex *mpi_fdtdTE_upml_getHz(void) { return solver_field(&mpi_te_solver, 2, B200FDTD_TE_HZ); }
double *mpi_fdtdTE_upml_getEps(void) { return mpi_te_solver.eps_ringed; }
int mpi_fdtdTE_upml_getSubNx(void)    { return N_PX; }
int mpi_fdtdTE_upml_getSubNy(void)    { return N_PY; }
int mpi_fdtdTE_upml_getSubNpx(void)   { return N_PX + 2; }
int mpi_fdtdTE_upml_getSubNpy(void)   { return N_PY + 2; }
int mpi_fdtdTE_upml_getSubNcell(void) { return (N_PX + 2) * (N_PY + 2); }

static UpmlSolver *solver_of_kind(int kind)
{
  switch (kind) {
  case B200FDTD_TM_UPML:     return &tm_solver;
  case B200FDTD_TE_UPML:     return &te_solver;
  case B200FDTD_MPI_TM_UPML: return &mpi_tm_solver;
  case B200FDTD_MPI_TE_UPML: return &mpi_te_solver;
  default:                   return NULL;
  }
}

/* engine handle of the active serial UPML solver, for harnesses that want device
 * timers or the U/W arrays (not part of the reference surface) */
/* slab g of the active serial UPML solver (NULL past the end), and how many there are */
b200fdtd_engine *mpifdtd_upml_slab_engine(int kind, int g)
{
  UpmlSolver *s = solver_of_kind(kind);
  if (s == NULL || g < 0 || g >= s->n_slabs) return NULL;
  flush_pending(s);
  return s->slab[g];
}
int mpifdtd_upml_slab_count(int kind)
{
  UpmlSolver *s = solver_of_kind(kind);
  return s != NULL ? s->n_slabs : 0;
}

b200fdtd_engine *mpifdtd_upml_engine(int kind)
{
  switch (kind) {                          /* whoever takes the handle sees every update() so far */
  case B200FDTD_TM_UPML:     flush_pending(&tm_solver); return tm_solver.engine;
  case B200FDTD_TE_UPML:     flush_pending(&te_solver); return te_solver.engine;
  case B200FDTD_MPI_TM_UPML: return mpi_tm_solver.engine;
  default:                   return mpi_te_solver.engine;
  }
}
