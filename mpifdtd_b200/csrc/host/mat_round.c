/* mat_round.c -- vacuum, Mie cylinder and concentric-cylinder permittivity models.
 *
 * Behavioural restatement of noModel.c, circleModel.c and
 * concentricCircleModel.c of rennone/mpiFDTD.  The arithmetic (operand order,
 * libm calls, comparison direction) follows the cited lines because epsilon maps
 * are compared bit-for-bit against the reference.
 */
#include <math.h>
#include <stdio.h>
#include "materials_internal.h"

/* ======================= vacuum (noModel.c:5-45) ========================== */
static double vacuum_eps(double x, double y, int col, int row)
{
  (void)x; (void)y; (void)col; (void)row;
  return EPSILON_0_S;
}
static material_eps_fn vacuum_select(void) { return vacuum_eps; }
static void vacuum_prepare(void) {}
static void vacuum_size(int *x_nm, int *y_nm) { *x_nm = 1000; *y_nm = 1000; }
static bool vacuum_advance(void) { return true; }
static void vacuum_dirs(void) { makeAndMoveDirectory("NoModel"); }
const MaterialModel material_vacuum = { "NoModel", vacuum_select, vacuum_prepare,
                                        vacuum_size, vacuum_advance, vacuum_dirs };

/* ================= Mie cylinder (circleModel.c:6-88) ====================== */
enum { MIE_RADIUS_FIRST_NM = 500, MIE_RADIUS_LAST_NM = 500, MIE_RADIUS_STEP_NM = 100 };
#define MIE_INDEX 1.6

static struct {
  int radius_nm;
  double radius, eps_in, cx, cy;     /* cell units, set by mie_prepare */
} mie = { .radius_nm = MIE_RADIUS_FIRST_NM };

/* circleModel.c:27-58 */
static double mie_eps(double x, double y, int col, int row)
{
  /* the PML band is vacuum */
  if (x < N_PML || y < N_PML || x > N_X + N_PML || y > N_Y + N_PML)
    return EPSILON_0_S;

  double dx = x - mie.cx, dy = y - mie.cy;
  double d2 = dx * dx + dy * dy;
  if (d2 >= (mie.radius + 1) * (mie.radius + 1))     /* a full cell outside */
    return EPSILON_0_S;
  if (d2 <= (mie.radius - 1) * (mie.radius - 1))     /* a full cell inside  */
    return mie.eps_in;

  /* rim cell: volume fraction over the 10 x 10 sub-cell lattice */
  double inside = 0;
  FOR_SUBCELL(u) {
    FOR_SUBCELL(v) {
      if (pow(dx + col * u / SUBCELL_SPLIT, 2.0) + pow(dy + row * v / SUBCELL_SPLIT, 2.0)
          <= mie.radius * mie.radius)
        inside += 1;
    }
  }
  inside /= SUBCELL_SPLIT * SUBCELL_SPLIT;
  return mie.eps_in * inside + EPSILON_0_S * (1 - inside);
}
static material_eps_fn mie_select(void) { return mie_eps; }

static void mie_prepare(void)                          /* circleModel.c:74-81 */
{
  FieldInfo_S g = field_getFieldInfo_S();
  mie.radius = field_toCellUnit(mie.radius_nm);
  mie.cx = g.N_PX / 2;                                 /* integer halves */
  mie.cy = g.N_PY / 2;
  mie.eps_in = MIE_INDEX * MIE_INDEX * EPSILON_0_S;
}
static void mie_size(int *x_nm, int *y_nm)             /* circleModel.c:83-88 */
{
  *x_nm = 2.3 * mie.radius_nm * 2;
  *y_nm = 2.3 * mie.radius_nm * 2;
}
static bool mie_advance(void)                          /* circleModel.c:60-64 */
{
  mie.radius_nm += MIE_RADIUS_STEP_NM;
  return mie.radius_nm > MIE_RADIUS_LAST_NM;
}
static void mie_dirs(void)
{
  char name[512];
  sprintf(name, "radius_%dnm", mie.radius_nm);
  makeDirectory(name);
  moveDirectory(name);
}
const MaterialModel material_mie_cylinder = { "MieCylinderModel", mie_select, mie_prepare,
                                              mie_size, mie_advance, mie_dirs };

/* ======== concentric cylinders (concentricCircleModel.c:10-104) ===========
 * Si core (r = 500 nm, n = 3.882) inside an SiO2 shell (r = 1000 nm, n = 1.457).
 * Upstream disables this model in models.c:82-90 and has no init/needSize; here
 * it is reachable only with MPIFDTD_ENABLE_CONCENTRIC=1, radii are resolved in
 * prepare() (after field_init, so the cell size is known) and need_size follows
 * the Mie model's margin rule for the outer radius. */
static struct { double radius[2], index[2], eps[2]; } ring;

/* concentricCircleModel.c:18-68.  The "fully inside the shell" early-out at
 * lines 38-40 can never fire (>= (R+1)^2 and <= (R-1)^2 at once); it is kept as
 * written because it is part of the reference's observable arithmetic. */
static double ring_eps_at(double cx, double cy, double x, double y, int col, int row)
{
  double dx = x - cx, dy = y - cy;
  double d2 = dx * dx + dy * dy;
  double r0 = ring.radius[0], r1 = ring.radius[1];

  if (d2 >= (r1 + 1) * (r1 + 1)) return EPSILON_0_S;
  if (d2 <= (r0 - 1) * (r0 - 1)) return ring.eps[0];
  if (d2 >= (r1 + 1) * (r1 + 1) && d2 <= (r1 - 1) * (r1 - 1)) return ring.eps[1];

  /* 32 x 32 sub-cell lattice, offsets (-15.5 ... 15.5)/32 */
  double part[2] = { 0, 0 };
  for (double u = -16 + 0.5; u < 16; u += 1) {
    for (double v = -16 + 0.5; v < 16; v += 1) {
      double s2 = pow(dx + col * u / 32.0, 2.0) + pow(dy + row * v / 32.0, 2.0);
      if (s2 < r0 * r0) {
        part[0] += 1;
      } else if (s2 == r1 * r1) {
        part[1] += 0.5;
        part[0] += 0.5;
      } else if (s2 < r1 * r1) {
        part[1] += 1;
      }
    }
  }
  part[0] /= 32.0 * 32.0;
  part[1] /= 32.0 * 32.0;
  return part[0] * ring.eps[0] + part[1] * ring.eps[1] + EPSILON_0_S * (1 - part[0] - part[1]);
}

static double ring_eps(double x, double y, int col, int row)   /* :70-80 */
{
  FieldInfo_S g = field_getFieldInfo_S();
  if (x < g.N_PML || y < g.N_PML || x > g.N_X + g.N_PML || y > g.N_Y + g.N_PML)
    return EPSILON_0_S;
  return ring_eps_at(g.N_PX / 2, g.N_PY / 2, x, y, col, row);
}
static void ring_prepare(void)                                  /* :82-94 */
{
  ring.radius[0] = field_toCellUnit(500);
  ring.radius[1] = field_toCellUnit(1000);
  ring.index[0] = 3.882;
  ring.index[1] = 1.457;
  ring.eps[0] = ring.index[0] * ring.index[0] * EPSILON_0_S;
  ring.eps[1] = ring.index[1] * ring.index[1] * EPSILON_0_S;
}
static material_eps_fn ring_select(void) { return ring_eps; }
static void ring_size(int *x_nm, int *y_nm) { *x_nm = 2.3 * 1000 * 2; *y_nm = 2.3 * 1000 * 2; }
static bool ring_advance(void) { return true; }                 /* :96-99 */
static void ring_dirs(void) {}                                  /* :101-104 */
const MaterialModel material_concentric = { "ConcentricCircleModel", ring_select, ring_prepare,
                                            ring_size, ring_advance, ring_dirs };
