/* plugin_table.c -- the solver table behind simulator_* (simulator.c:19-279 of
 * rennone/mpiFDTD): nine function pointers filled from the selected solver's
 * entry points, plus the lifecycle calls the driver makes (main.c:150-213):
 *
 *   simulator_setSolver(id) -> simulator_init(FieldInfo) -> simulator_calc()*
 *     -> simulator_reset() | simulator_finish()
 *
 * A solver is described by one row of `solver_rows`; selecting it copies the row
 * into the active table.  All eight solver ids of simulator.h:8-18 are served; an unknown
 * id keeps the reference's error convention: a message and exit(2).
 */
#include <stdio.h>
#include <stdlib.h>
#include <sys/time.h>
#include "host_internal.h"

typedef void (*step_fn)(void);
typedef double complex *(*field_fn)(void);
typedef double *(*eps_fn)(void);

typedef struct SolverRow {
  const char *banner;          /* printed on selection, as upstream does      */
  const char *dir;             /* output sub-directory (simulator.c:31-173)   */
  step_fn (*get_update)(void), (*get_init)(void), (*get_finish)(void), (*get_reset)(void);
  field_fn data_x, data_y, data_z;
  int draw;                    /* which of x/y/z the viewer paints: 1 = y, 2 = z */
  eps_fn eps;
} SolverRow;

#define ROW(P, A, B, Cc, banner, dir, draw) \
  { banner, dir, P##_getUpdate, P##_getInit, P##_getFinish, P##_getReset, \
    P##_get##A, P##_get##B, P##_get##Cc, draw, P##_getEps }

/* TM solvers draw Ez (getDataZ), TE solvers draw Ey (getDataY): simulator.c:43,62 */
static const SolverRow solver_rows[] = {
  [TM_2D]      = ROW(fdtdTM, Hx, Hy, Ez, "TM mode \n", "TM", 2),
  [TE_2D]      = ROW(fdtdTE, Ex, Ey, Hz, "TE mode \n", "TE", 1),
  [TM_UPML_2D] = ROW(fdtdTM_upml, Hx, Hy, Ez, "TM UPML mode \n", "TM_UPML", 2),
  [TE_UPML_2D] = ROW(fdtdTE_upml, Ex, Ey, Hz, "TE UPML mode \n", "TE_UPML", 1),
  [MPI_TM_UPML_2D] = ROW(mpi_fdtdTM_upml, Hx, Hy, Ez, "MPI TM UPML mode \n", "MPI_TM_UPML", 2),
  [MPI_TE_UPML_2D] = ROW(mpi_fdtdTE_upml, Ex, Ey, Hz, "MPI TE UPML mode \n", "MPI_TE_UPML", 1),
  [NS_TM_2D]   = ROW(nsFdtdTM, Hx, Hy, Ez, "NS TM mode \n", "NS_TM", 2),
  [NS_TE_2D]   = ROW(nsFdtdTE, Ex, Ey, Hz, "NS TE mode \n", "NS_TE", 1),
};

static struct {
  step_fn update, init, finish, reset;
  field_fn data_x, data_y, data_z, draw;
  eps_fn eps;
  const char *dir;
  struct timeval started;
} active = { .dir = "" };

void simulator_setSolver(enum SOLVER id)                      /* simulator.c:175-224 */
{
  const int n_rows = (int)(sizeof solver_rows / sizeof solver_rows[0]);
  if ((int)id < 0 || (int)id >= n_rows || solver_rows[id].get_update == NULL) {
    if ((int)id >= 0 && (int)id <= NS_TE_2D)
      printf("error, solver %d has no GPU kernels in this build (plugin_table.c)\n", (int)id);
    else
      printf("error, not implement simulator (simulator.c)\n");
    exit(2);
  }
  const SolverRow *row = &solver_rows[id];
  active.update = row->get_update();
  active.init   = row->get_init();
  active.finish = row->get_finish();
  active.reset  = row->get_reset();
  active.data_x = row->data_x;
  active.data_y = row->data_y;
  active.data_z = row->data_z;
  active.draw   = row->draw == 1 ? row->data_y : row->data_z;
  active.eps    = row->eps;
  active.dir    = row->dir;
  printf("%s", row->banner);
}

void simulator_moveDirectory(void)                            /* simulator.c:209-213 */
{
  makeDirectory(active.dir);
  moveDirectory(active.dir);
}

void simulator_calc(void)                                     /* simulator.c:215-219 */
{
  active.update();
  field_nextStep();
}

void simulator_init(FieldInfo field_info)                     /* simulator.c:226-234 */
{
  field_init(field_info);
  models_initModel();
  active.init();
  gettimeofday(&active.started, NULL);
}

void simulator_solverInit(void)                               /* simulator.c:236-241 */
{
  makeDirectory(active.dir);
  moveDirectory(active.dir);
  active.init();
}

void simulator_reset(void)                                    /* simulator.c:243-250 */
{
  printf("simulator_reset \n");
  field_reset();
  active.reset();
  gettimeofday(&active.started, NULL);
}

void simulator_changeModelAndRestart(void) { moveDirectory(".."); }   /* simulator.c:252-255 */

void simulator_finish(void)                                   /* simulator.c:257-264 */
{
  struct timeval now;
  printf("simulator_finish at %d step \n", (int)field_getTime());
  gettimeofday(&now, NULL);
  printf("time = %lf \n",
         now.tv_sec - active.started.tv_sec + (now.tv_usec - active.started.tv_usec) * 1e-6);
  active.finish();
}

double complex *simulator_getDrawingData(void) { return active.draw(); }     /* simulator.c:266 */
bool simulator_isFinish(void) { return field_isFinish(); }                   /* simulator.c:270 */
double *simulator_getEps(void) { return active.eps(); }                      /* simulator.c:276 */
