/* materials.c -- models_* dispatcher (models.c:13-160 of rennone/mpiFDTD) and a
 * multi-threaded permittivity-map builder on top of it.
 *
 * The permittivity callbacks stay host C on purpose: epsilon maps must be
 * bit-exact against the reference, which means the same libm calls in the same
 * order.  They run once per structure at init; the per-step path never calls
 * them.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include "materials_internal.h"
#include "host_internal.h"

static const MaterialModel *active_model;
static material_eps_fn active_eps;

/* MPIFDTD_ENABLE_CONCENTRIC=1 opts in to the concentric-circle model, which the
 * reference ships disabled (models.c:82-90 prints and exits). */
static int concentric_enabled(void)
{
  const char *v = getenv("MPIFDTD_ENABLE_CONCENTRIC");
  return v != NULL && v[0] == '1';
}

void models_setModel(enum MODEL model)
{
  static const MaterialModel *const table[] = {
    [NO_MODEL] = &material_vacuum,        [MIE_CYLINDER] = &material_mie_cylinder,
    [LAYER] = &material_multilayer,       [MORPHO_SCALE] = &material_morpho,
    [CONCENTRIC_CIRCLE] = &material_concentric,
    [ZIGZAG] = &material_zigzag,          [TRACE_IMAGE] = &material_trace_image };
  if ((int)model < 0 || (int)model > TRACE_IMAGE)
    return;                                   /* upstream's switch has no default */
  if (model == CONCENTRIC_CIRCLE && !concentric_enabled()) {
    printf("not implemented concentricCircle Model");
    exit(2);
  }
  active_model = table[model];
  active_eps = active_model->select();
}

double models_eps(double x, double y, enum MODE mode)
{
  switch (mode) {
  case D_X: return active_eps(x, y, 1, 0);
  case D_Y: return active_eps(x, y, 0, 1);
  default:  return active_eps(x, y, 1, 1);   /* D_XY and anything else */
  }
}

bool models_isFinish(void)  { return active_model->advance(); }
void models_initModel(void) { active_model->prepare(); }
void models_needSize(int *x_nm, int *y_nm) { active_model->need_size(x_nm, y_nm); }

void models_moveDirectory(void)
{
  makeDirectory(active_model->dir);
  moveDirectory(active_model->dir);
  active_model->enter_dirs();
}

/* ---- permittivity map builder ------------------------------------------------
 * dst[i*N_PY + j] = models_eps(i + xoff, j + yoff, mode) for i in [i0, i1).
 * Rows are independent and every callback is a pure function of static model
 * parameters after models_initModel(), so rows are spread over host threads. */
typedef struct { double *dst; double xoff, yoff; enum MODE mode; int i0, i1, j0, nj; } EpsJob;

static void *eps_rows(void *arg)
{
  const EpsJob *job = (const EpsJob *)arg;
  for (int i = job->i0; i < job->i1; i++) {
    double *row = job->dst + (size_t)i * job->nj;
    for (int j = 0; j < job->nj; j++)
      row[j] = models_eps(i + job->xoff, (job->j0 + j) + job->yoff, job->mode);
  }
  return NULL;
}

void mpifdtd_fill_eps(double *dst, double xoff, double yoff, enum MODE mode)
{
  mpifdtd_fill_eps_slab(dst, xoff, yoff, mode, 0, field_getFieldInfo_S().N_PY);
}

void mpifdtd_fill_eps_slab(double *dst, double xoff, double yoff, enum MODE mode, int j0, int nj)
{
  FieldInfo_S g = field_getFieldInfo_S();
  long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
  const char *cap = getenv("MPIFDTD_HOST_THREADS");
  if (cap != NULL && atoi(cap) > 0) ncpu = atoi(cap);
  int nthr = (int)MPIFDTD_MIN(MPIFDTD_MAX(ncpu, 1), 64);
  if (nthr > g.N_PX) nthr = g.N_PX > 0 ? g.N_PX : 1;
  pthread_t tid[64];
  EpsJob job[64];
  for (int t = 0; t < nthr; t++) {
    job[t] = (EpsJob){ dst, xoff, yoff, mode,
                       (int)((long long)g.N_PX * t / nthr),
                       (int)((long long)g.N_PX * (t + 1) / nthr), j0, nj };
    if (t == nthr - 1 || pthread_create(&tid[t], NULL, eps_rows, &job[t]) != 0) {
      /* last chunk (or a failed spawn) runs on the calling thread */
      eps_rows(&job[t]);
      tid[t] = pthread_self();
    }
  }
  for (int t = 0; t < nthr - 1; t++)
    if (!pthread_equal(tid[t], pthread_self()))
      pthread_join(tid[t], NULL);
}

/* ---- permittivity map as a palette ----------------------------------------------------------------
 * A map holds few distinct values: vacuum, the materials, and the area-averaged cells along the
 * material boundaries.  index[k] = position of map[k] in `table` (distinct bit patterns in order of
 * first appearance, table[0] = the first cell's value); returns the number of distinct values, or -1
 * when there are more than 65536 (the caller then uploads the dense map).  Open addressing on the
 * 64-bit pattern; runs of equal neighbours -- almost every cell -- skip the table. */
#include <stdint.h>
#include <string.h>
int mpifdtd_eps_palette(const double *map, size_t n, uint16_t *index, double *table)
{
  enum { SLOTS = 1 << 18 };                       /* 4 x the largest palette */
  int32_t *slot = (int32_t *)malloc(sizeof(int32_t) * SLOTS);
  if (slot == NULL) return -1;
  memset(slot, 0xff, sizeof(int32_t) * SLOTS);
  int n_values = 0;
  uint64_t last_bits = 0;
  int last_index = -1;
  for (size_t k = 0; k < n; k++) {
    uint64_t bits;
    memcpy(&bits, &map[k], sizeof bits);
    if (last_index >= 0 && bits == last_bits) { index[k] = (uint16_t)last_index; continue; }
    uint64_t h = bits * 0x9e3779b97f4a7c15ull;
    uint32_t at = (uint32_t)(h >> 46) & (SLOTS - 1);
    for (;;) {
      if (slot[at] < 0) {
        if (n_values == 65536) { free(slot); return -1; }
        slot[at] = n_values;
        table[n_values++] = map[k];
        break;
      }
      uint64_t have;
      memcpy(&have, &table[slot[at]], sizeof have);
      if (have == bits) break;
      at = (at + 1) & (SLOTS - 1);
    }
    last_bits = bits;
    last_index = slot[at];
    index[k] = (uint16_t)last_index;
  }
  free(slot);
  return n_values;
}
