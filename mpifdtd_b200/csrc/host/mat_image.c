/* mat_image.c -- permittivity from a traced refractive-index image.
 *
 * Behavioural restatement of traceImageModel.c:8-159 of rennone/mpiFDTD.  The
 * image file holds "width height scale_px" followed by height rows of width
 * refractive indices; it is read when the model is selected, like upstream
 * (traceImageModel_EPS calls readImage).  Upstream opens "traceImage1.txt"
 * although the shipped file is "traceImage.txt"; the name is kept (quirk), with
 * MPIFDTD_TRACE_IMAGE as an opt-in override of the path.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "materials_internal.h"

#define TI_FILE_STEM "traceImage1"
enum { TI_SCALE_FIRST_NM = 300, TI_SCALE_LAST_NM = 1000, TI_SCALE_STEP_NM = 100 };

static struct {
  int scale_nm, scale_px, width_px, height_px;
  double nm_per_px, px_per_cell, width, height, left, top;
  double *eps_px;                       /* [x*height_px + y] */
} ti = { .scale_nm = TI_SCALE_FIRST_NM };

/* bilinear lookup in pixel space (traceImageModel.c:26-43); the clamp mixes int
 * and double through the reference's min/max macros */
static double ti_lookup(double x, double y)
{
  double xp = MPIFDTD_MIN(ti.width_px - 1, MPIFDTD_MAX(0, x * ti.px_per_cell));
  double yp = MPIFDTD_MIN(ti.height_px - 1, MPIFDTD_MAX(0, y * ti.px_per_cell));
  if (xp == ti.width_px - 1 || yp == ti.height_px - 1)
    return EPSILON_0_S;
  int i = floor(xp), j = floor(yp);
  double fx = xp - i, fy = yp - j;
  const double *q = ti.eps_px + (i * ti.height_px + j);
  return q[0] * (1.0 - fx) * (1.0 - fy) + q[ti.height_px] * fx * (1.0 - fy)
       + q[1] * (1.0 - fx) * fy         + q[ti.height_px + 1] * fx * fy;
}

/* traceImageModel.c:45-76: image y axis points down, so y is mirrored about the top edge */
static double ti_eps(double x_in, double y_in, int col, int row)
{
  double x = x_in - ti.left;
  double y = -y_in + ti.top;
  if (x < -0.5 || x >= ti.width + 0.5 || y < -0.5 || y > ti.height + 0.5)
    return EPSILON_0_S;

  double acc = 0;
  FOR_SUBCELL(u) {
    FOR_SUBCELL(v) {
      double sx = x + col * u / SUBCELL_SPLIT;
      double sy = y + row * v / SUBCELL_SPLIT;
      if (sx < 0 || sx >= ti.width - 1 || sy < 0 || sy >= ti.height - 1)
        acc += EPSILON_0_S;
      else
        acc += ti_lookup(sx, sy);
    }
  }
  acc = acc / SUBCELL_SPLIT / SUBCELL_SPLIT;
  if (acc < EPSILON_0_S)
    printf("%lf \n", acc);
  return acc;
}

static void ti_read(void)                                   /* :78-104 */
{
  const char *path = getenv("MPIFDTD_TRACE_IMAGE");
  if (path == NULL) path = TI_FILE_STEM ".txt";
  FILE *fp = fopen(path, "r");
  if (fp == NULL) {
    printf("cannot find traceImage.txt of morphoScaleModel\n");
    exit(2);
  }
  if (fscanf(fp, "%d %d %d", &ti.width_px, &ti.height_px, &ti.scale_px) != 3) {
    printf("cannot find traceImage.txt of morphoScaleModel\n");
    exit(2);
  }
  ti.nm_per_px = 1.0 * ti.scale_nm / ti.scale_px;
  free(ti.eps_px);
  ti.eps_px = newDouble(ti.width_px * ti.height_px);
  for (int y = 0; y < ti.height_px; y++) {
    for (int x = 0; x < ti.width_px; x++) {
      double n = 1.0;
      if (fscanf(fp, "%lf ", &n) != 1) n = 1.0;
      if (n < 1.0)
        printf("%lf \n", n);
      ti.eps_px[x * ti.height_px + y] = n * n * EPSILON_0_S;
    }
  }
  fclose(fp);
  printf("%d %d %.2lf", ti.width_px, ti.height_px, ti.nm_per_px);
}
static material_eps_fn ti_select(void) { ti_read(); return ti_eps; }

static void ti_prepare(void)                                /* :145-159 */
{
  ti.px_per_cell = 1.0 / field_toCellUnit(ti.nm_per_px);
  ti.width = field_toCellUnit(ti.width_px * ti.nm_per_px);
  ti.height = field_toCellUnit(ti.height_px * ti.nm_per_px);
  FieldInfo_S g = field_getFieldInfo_S();
  ti.left = g.N_PX / 2 - ti.width / 2;
  ti.top = g.N_PY / 2 + ti.height / 2;
  printf("px=(%d,%d), %lf\n", ti.width_px, ti.height_px, ti.nm_per_px);
  printf("s=(%lf,%lf)\n", ti.width, ti.height);
  printf("lp=(%lf,%lf)\n", ti.left, ti.top);
  printf("%d %d\n", g.N_X, g.N_Y);
}
static void ti_size(int *x_nm, int *y_nm)                   /* :139-143 */
{
  *x_nm = ceil(ti.width_px * ti.nm_per_px);
  *y_nm = ceil(ti.height_px * ti.nm_per_px);
}
static bool ti_advance(void)                                /* :106-115 */
{
  ti.scale_nm += TI_SCALE_STEP_NM;
  if (ti.scale_nm > TI_SCALE_LAST_NM)
    return true;
  ti.nm_per_px = ti.scale_nm / ti.scale_px;                 /* integer division upstream */
  return false;
}
static void ti_dirs(void)                                   /* :128-137 */
{
  char name[512];
  makeAndMoveDirectory(TI_FILE_STEM);
  sprintf(name, "scale_width%d", ti.scale_nm);
  makeAndMoveDirectory(name);
}
const MaterialModel material_trace_image = { "TraceImageModel", ti_select, ti_prepare,
                                             ti_size, ti_advance, ti_dirs };
