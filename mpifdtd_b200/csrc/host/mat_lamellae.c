/* mat_lamellae.c -- layered-ridge permittivity models: the two-material lamella
 * stack, the Morpho-butterfly scale ridge and the zig-zag slab stack.
 *
 * Behavioural restatement of multiLayerModel.c, morphoScaleModel.c and
 * zigzagModel.c of rennone/mpiFDTD (cited per function).  Compile-time switches
 * of the reference (ASYMMETRY, USE_GAP, UNIAXIAL, CURVE, RANDOMNESS) are kept as
 * named constants with the reference's default values, and the expressions they
 * gate are kept whole so the floating-point result is the reference's.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "materials_internal.h"

#ifndef M_PI
#define M_PI 3.1415926535897932384626433832795
#endif

/* ===================================================================== *
 *  Two-material lamella stack  (multiLayerModel.c:7-348)
 * ===================================================================== */
#define ML_N0 1.0
#define ML_N1 1.56
#define ML_UNIAXIAL 0
#define ML_N0_X 1.0
#define ML_N1_X 1.1
#define ML_ASYMMETRY 0
#define ML_USE_GAP 0
#define ML_GAP_STEP_NM 0
#define ML_CURVE 0.0
enum { ML_WIDTH_NM = 300,
       ML_THICK0_FIRST = 100, ML_THICK0_STEP = 10,
       ML_THICK1_FIRST = 90, ML_THICK1_LAST = 150, ML_THICK1_STEP = 10,
       ML_LAYERS_FIRST = 4, ML_LAYERS_LAST = 10, ML_LAYERS_STEP = 4,
       ML_BRANCH_FIRST = 0, ML_BRANCH_LAST = 50, ML_BRANCH_STEP = 50 };
#define ML_EDGE_FIRST 0.0
#define ML_EDGE_LAST 1.0
#define ML_EDGE_STEP 0.5

static struct {
  int width_nm[2], thick_nm[2], branch_nm, layers, gap_nm;
  double edge_rate;
  /* cell units, resolved in ml_prepare */
  double width[2], thick[2], branch, gap, eps[2], eps_x[2], bow[2];
} ml = { .width_nm = { ML_WIDTH_NM, ML_WIDTH_NM }, .thick_nm = { ML_THICK0_FIRST, ML_THICK1_FIRST },
         .branch_nm = ML_BRANCH_FIRST, .layers = ML_LAYERS_FIRST, .gap_nm = 0,
         .edge_rate = ML_EDGE_FIRST };

/* lamella width at height sy, tapered towards the tip and bowed by a parabola
 * across the lamella's own thickness (multiLayerModel.c:87-103) */
static double ml_width_at(double sx, double sy, double wid, double hei, double mod_y, int k)
{
  double p = 1 - sy / hei;
  double tapered = (wid + ml.branch) * (p + (1 - p) * ml.edge_rate);
  double dh = k == 0 ? mod_y : mod_y - ml.thick[0];
  double c = k == 0 ? ml.bow[0] : ml.bow[1];
  if (ML_ASYMMETRY && sx < 0)
    dh = (k == 1 ? mod_y : mod_y - ml.thick[1]);
  return c * pow((dh - ml.thick[k] / 2), 2) + tapered;
}

/* multiLayerModel.c:107-185 */
static double ml_eps(double x, double y, int col, int row)
{
  FieldInfo_S g = field_getFieldInfo_S();
  double width = MPIFDTD_MAX(ml.width[0], ml.width[1]);
  double period = ml.thick[0] + ml.thick[1];
  double height = period * ml.layers + ml.gap;

  /* stack bottom sits height/2 below the centre; both origins are truncated to int */
  int oy = g.N_PY / 2 - height / 2;
  int ox = g.N_PX / 2;
  double lx = x - ox, ly = y - oy;

  if (fabs(lx) > (width / 2 + 0.5) || ly < -0.5 || ly > height + 0.5)
    return EPSILON_0_S;

  height = period * ml.layers;
  double part[2] = { 0, 0 };
  FOR_SUBCELL(u) {
    FOR_SUBCELL(v) {
      double sx = lx + col * u / SUBCELL_SPLIT;
      double sy = ly + row * v / SUBCELL_SPLIT;
      if (sx < 0 && ML_USE_GAP)
        sy -= ml.gap;
      if (sy < 0 || sy > height)
        continue;

      double p = 1 - sy / height;
      if (fabs(sx) < ml.branch * (p + (1 - p) * ml.edge_rate)) {   /* central stem */
        part[1] += 1;
        continue;
      }

      double mod_y = sy - floor(sy / period) * period;
      if (mod_y == ml.thick[0]) {            /* exactly on the interface: half each */
        part[0] += 0.5 * (fabs(sx) < ml.width[0] / 2);
        part[1] += 0.5 * (fabs(sx) < ml.width[1] / 2);
        continue;
      }

      int k;
      if (sx < 0 && ML_ASYMMETRY)
        k = (mod_y < ml.thick[1]);
      else
        k = (mod_y > ml.thick[0]);

      double wid = ml_width_at(sx, sy, ml.width[k], height, mod_y, k);
      if (fabs(sx) < wid / 2)
        part[k] += 1;
    }
  }
  part[0] /= SUBCELL_SPLIT * SUBCELL_SPLIT;
  part[1] /= SUBCELL_SPLIT * SUBCELL_SPLIT;
  if (ML_UNIAXIAL && col == 0 && row == 1)
    return EPSILON_0_S * (1 - part[0] - part[1]) + ml.eps_x[0] * part[0] + ml.eps_x[1] * part[1];
  return EPSILON_0_S * (1 - part[0] - part[1]) + ml.eps[0] * part[0] + ml.eps[1] * part[1];
}
static material_eps_fn ml_select(void) { return ml_eps; }

static void ml_prepare(void)                              /* multiLayerModel.c:331-348 */
{
  for (int k = 0; k < 2; k++) {
    ml.width[k] = field_toCellUnit(ml.width_nm[k]);
    ml.thick[k] = field_toCellUnit(ml.thick_nm[k]);
  }
  ml.gap = field_toCellUnit(ml.gap_nm);
  ml.eps[0] = ML_N0 * ML_N0 * EPSILON_0_S;
  ml.eps[1] = ML_N1 * ML_N1 * EPSILON_0_S;
  ml.eps_x[0] = ML_N0_X * ML_N0_X * EPSILON_0_S;
  ml.eps_x[1] = ML_N1_X * ML_N1_X * EPSILON_0_S;
  ml.branch = field_toCellUnit(ml.branch_nm);
  ml.bow[0] = -4 * ml.width[0] * ML_CURVE / ml.thick[0] / ml.thick[0];
  ml.bow[1] = -4 * ml.width[1] * ML_CURVE / ml.thick[1] / ml.thick[1];
}

static void ml_size(int *x_nm, int *y_nm)                 /* multiLayerModel.c:277-283 */
{
  *x_nm = MPIFDTD_MAX(ml.width_nm[0], ml.width_nm[1]) + ml.branch_nm;
  *y_nm = (ml.thick_nm[0] + ml.thick_nm[1]) * ml.layers
        + (ML_USE_GAP ? ml.thick_nm[0] + ml.thick_nm[0] : 0);
}

/* parameter sweep, innermost first: gap, thick0, thick1, edge rate, stem, layers
 * (nextStructure1, multiLayerModel.c:235-269) */
static bool ml_advance(void)
{
  ml.gap_nm += ML_GAP_STEP_NM;
  if (ml.gap_nm >= (ml.thick_nm[0] + ml.thick_nm[1]) || !ML_USE_GAP) {
    ml.gap_nm = ML_USE_GAP ? ML_GAP_STEP_NM : 0;
    ml.thick_nm[0] += ML_THICK0_STEP;
    if (ml.thick_nm[0] > ml.thick_nm[1] + 50) {
      ml.thick_nm[1] += ML_THICK1_STEP;
      ml.thick_nm[0] = ml.thick_nm[1] + 10;
      if (ml.thick_nm[1] > ML_THICK1_LAST) {
        ml.thick_nm[1] = ML_THICK1_FIRST;
        ml.edge_rate += ML_EDGE_STEP;
        if (ml.edge_rate > ML_EDGE_LAST) {
          ml.edge_rate = ML_EDGE_FIRST;
          ml.branch_nm += ML_BRANCH_STEP;
          if (ml.branch_nm > ML_BRANCH_LAST) {
            ml.branch_nm = ML_BRANCH_FIRST;
            ml.layers += ML_LAYERS_STEP;
            if (ml.layers > ML_LAYERS_LAST) {
              printf("there are no models which hasn't been simulated yet\n");
              return true;
            }
          }
        }
      }
    }
  }
  return false;
}

static void ml_dirs(void)                                 /* multiLayerModel.c:285-329 */
{
  char name[512];
  makeAndMoveDirectory(ML_ASYMMETRY ? "asymmetry" : "symmetry");
  if (ML_UNIAXIAL)
    sprintf(name, "uniaxial_n0y%.2lf_n0x%.2lf_n1x%.2lf_n1x%.2lf", ML_N0, ML_N0_X, ML_N1, ML_N1_X);
  else
    sprintf(name, "n_%.2lf_%.2lf", ML_N0, ML_N1);
  makeAndMoveDirectory(name);
  sprintf(name, "curve_%.2lf", ML_CURVE);                   makeAndMoveDirectory(name);
  sprintf(name, "width%d_%d", ml.width_nm[0], ml.width_nm[0]); makeAndMoveDirectory(name);
  sprintf(name, "thick%d_%d", ml.thick_nm[0], ml.thick_nm[1]); makeAndMoveDirectory(name);
  sprintf(name, "gap%d", ml.gap_nm);                        makeAndMoveDirectory(name);
  sprintf(name, "layer%d", ml.layers);                      makeAndMoveDirectory(name);
  sprintf(name, "edge%.1lf", ml.edge_rate);                 makeAndMoveDirectory(name);
  sprintf(name, "branch%d", ml.branch_nm);                  makeAndMoveDirectory(name);
}
const MaterialModel material_multilayer = { "MultiLayerModel", ml_select, ml_prepare,
                                            ml_size, ml_advance, ml_dirs };

/* ===================================================================== *
 *  Morpho scale ridge  (morphoScaleModel.c:10-440)
 * ===================================================================== */
#define MS_SIDE_LEFT 0
#define MS_SIDE_RIGHT 1
#define MS_UNIAXIAL 0
#define MS_N0 1.56
#define MS_N0_X 1.1
#define MS_ASYMMETRY 0
#define MS_USE_GAP 0
#define MS_GAP_STEP_NM 10
#define MS_CURVE 0.0
#define MS_RANDOMNESS 0
#define MS_RANDOM_SEED 0
enum { MS_WIDTH_NM = 300,
       MS_THICK0_FIRST = 90, MS_THICK0_STEP = 10,
       MS_THICK1_FIRST = 90, MS_THICK1_LAST = 160, MS_THICK1_STEP = 10,
       MS_LAYERS_FIRST = 6, MS_LAYERS_LAST = 10, MS_LAYERS_STEP = 2,
       MS_BRANCH_FIRST = 0, MS_BRANCH_LAST = 50, MS_BRANCH_STEP = 25 };
#define MS_EDGE_FIRST 0.0
#define MS_EDGE_LAST 1.0
#define MS_EDGE_STEP 0.5

/* one lamella = the strip between two parallel lines y = a x + b1 and y = a x + b2 */
typedef struct { double a, b1, b2; int id; } Strip;

static struct {
  int width_nm, thick_nm[2], layers, branch_nm, gap_nm;
  double edge_rate;
  double width, thick[2], branch, gap, height, ox, oy, eps, eps_x;
  Strip strip[2][MS_LAYERS_LAST];
  double tilt[2][MS_LAYERS_LAST];
} ms = { .width_nm = MS_WIDTH_NM, .thick_nm = { MS_THICK0_FIRST, MS_THICK1_FIRST },
         .layers = MS_LAYERS_FIRST, .branch_nm = MS_BRANCH_FIRST, .gap_nm = 0,
         .edge_rate = MS_EDGE_FIRST };

static bool strip_holds(const Strip *s, double x, double y)        /* :101-104 */
{
  return (s->a * x + s->b1 <= y && y <= s->a * x + s->b2);
}

/* half-width of lamella s on the line through (sx, sy) parallel to it (:107-125) */
static double ms_half_width(double sx, double sy, const Strip *s)
{
  double b = (sy - sx * s->a);
  int side = sx < 0 ? MS_SIDE_LEFT : MS_SIDE_RIGHT;
  double rad = ms.tilt[side][s->id];
  double shift = (side == MS_SIDE_LEFT ? -1 : 1) * (0.5 * (s->b2 + s->b1) - b) * sin(rad);

  double p = 1 - b / ms.height;
  double taper = (p + (1 - p) * ms.edge_rate);
  double yy = (b - s->b1) / (s->b2 - s->b1) - 0.5;
  double half = ms.width / 2.0;
  double c0 = 4.0 * half * MS_CURVE;
  return taper * (half * (1 - MS_CURVE)) + c0 * (0.25 - yy * yy) + shift;
}

/* which lamella (if any) a cell-sized rectangle touches; side chosen by the
 * sign of its left edge (:128-145) */
static void ms_find_strip(double lft, double rht, double btm, double top, Strip *out)
{
  out->id = -1;
  int side = lft < 0 ? MS_SIDE_LEFT : MS_SIDE_RIGHT;
  for (int n = 0; n < ms.layers; n++) {
    out->a = ms.strip[side][n].a;
    out->b1 = ms.strip[side][n].b1;
    out->b2 = ms.strip[side][n].b2;
    if (strip_holds(out, lft, btm) || strip_holds(out, rht, btm) ||
        strip_holds(out, lft, top) || strip_holds(out, rht, top)) {
      out->id = n;
      return;
    }
  }
}

/* morphoScaleModel.c:147-253 */
static double ms_eps(double x, double y, int col, int row)
{
  double lx = x - ms.ox, ly = y - ms.oy;

  /* lamellae rotate about their centres, so the reach is the diagonal */
  if (fabs(lx) > 0.5 * sqrt(pow(ms.width, 2) + pow(ms.thick[0], 2)) + 0.5)
    return EPSILON_0_S;

  Strip near[2];
  near[0].id = near[1].id = -1;
  if (fabs(lx) > 0.5) {
    if (lx < 0)
      ms_find_strip(lx - 0.5, lx + 0.5, ly - ms.gap - 0.5, ly - ms.gap + 0.5,
                    lx < 0 ? &near[MS_SIDE_LEFT] : &near[MS_SIDE_RIGHT]);
    else
      ms_find_strip(lx - 0.5, lx + 0.5, ly - 0.5, ly + 0.5,
                    lx < 0 ? &near[MS_SIDE_LEFT] : &near[MS_SIDE_RIGHT]);
  } else {
    ms_find_strip(lx - 0.5, 0, ly - ms.gap - 0.5, ly - ms.gap + 0.5, &near[MS_SIDE_LEFT]);
    ms_find_strip(0, lx + 0.5, ly - 0.5, ly + 0.5, &near[MS_SIDE_RIGHT]);
  }

  double filled = 0;
  FOR_SUBCELL(u) {
    FOR_SUBCELL(v) {
      double sx = lx + col * u / SUBCELL_SPLIT;
      double sy = ly + row * v / SUBCELL_SPLIT;
      if (sx < 0)
        sy -= ms.gap;

      if (sy >= 0 && sy <= ms.height) {                    /* central stem */
        double p = 1 - sy / ms.height;
        double taper = (p + (1 - p) * ms.edge_rate);
        if (fabs(sx) < ms.branch / 2.0 * taper) {
          filled += 1;
          continue;
        }
      }

      int side = sx < 0 ? MS_SIDE_LEFT : MS_SIDE_RIGHT;
      const Strip *hit = &near[side];
      if (hit->id < 0)
        continue;
      hit = &ms.strip[side][hit->id];

      if (strip_holds(hit, sx, sy)) {
        double reach2 = pow(sx, 2) * (1 + pow(hit->a, 2));
        double half = ms_half_width(sx, sy, hit);
        if (reach2 < pow(half, 2)) {
          filled += 1;
          continue;
        }
      }
      /* otherwise try the lamella just below and just above */
      for (int d = -1; d < 2; d += 2) {
        int id = hit->id + d;
        if (id < 0 || id >= ms.layers)
          continue;
        const Strip *nb = &ms.strip[side][id];
        if (strip_holds(nb, sx, sy)) {
          double reach2 = pow(sx, 2) * (1 + pow(nb->a, 2));
          double half = ms_half_width(sx, sy, nb);
          if (reach2 < pow(half, 2)) {
            filled += 1;
            break;
          }
        }
      }
    }
  }
  filled /= SUBCELL_SPLIT * SUBCELL_SPLIT;
  if (MS_UNIAXIAL && col == 0 && row == 1)
    return EPSILON_0_S * (1 - filled) + ms.eps_x * filled;
  return EPSILON_0_S * (1 - filled) + ms.eps * filled;
}
static material_eps_fn ms_select(void) { return ms_eps; }

static void ms_prepare(void)                                /* :379-422 */
{
  ms.width = field_toCellUnit(ms.width_nm);
  ms.thick[0] = field_toCellUnit(ms.thick_nm[0]);
  ms.thick[1] = field_toCellUnit(ms.thick_nm[1]);
  ms.branch = field_toCellUnit(ms.branch_nm);
  ms.gap = field_toCellUnit(ms.gap_nm);
  ms.height = (ms.thick[0] + ms.thick[1]) * ms.layers + (ms.thick[0] + ms.thick[1]);
  ms.eps = MS_N0 * MS_N0 * EPSILON_0_S;
  ms.eps_x = MS_N0_X * MS_N0_X * EPSILON_0_S;

  FieldInfo_S g = field_getFieldInfo_S();
  ms.oy = g.N_PY / 2 - ms.height / 2;
  ms.ox = g.N_PX / 2;

  /* same generator, seed and draw order as upstream: right side first */
  srand(MS_RANDOM_SEED);
  double to_rad = M_PI / 180.0;
  for (int n = 0; n < ms.layers; n++) {
    ms.tilt[MS_SIDE_RIGHT][n] = (rand() % (MS_RANDOMNESS + 1) - MS_RANDOMNESS / 2) * to_rad;
    ms.tilt[MS_SIDE_LEFT][n]  = (rand() % (MS_RANDOMNESS + 1) - MS_RANDOMNESS / 2) * to_rad;
  }

  double period = ms.thick[0] + ms.thick[1];
  double half_thick = 0.5 * ms.thick[0];
  double base[2];
  base[MS_SIDE_LEFT]  = MS_ASYMMETRY ? ms.thick[1] + half_thick : half_thick;
  base[MS_SIDE_RIGHT] = half_thick;
  for (int n = 0; n < ms.layers; n++) {
    for (int side = MS_SIDE_LEFT; side <= MS_SIDE_RIGHT; side++) {
      Strip *s = &ms.strip[side][n];
      s->a  = tan(ms.tilt[side][n]);
      s->b1 = period * n + base[side] - half_thick / cos(ms.tilt[side][n]);
      s->b2 = period * n + base[side] + half_thick / cos(ms.tilt[side][n]);
      s->id = n;
    }
  }
}

static void ms_size(int *x_nm, int *y_nm)                   /* :435-440 */
{
  *x_nm = ms.width_nm + ms.branch_nm;
  *y_nm = (ms.thick_nm[0] + ms.thick_nm[1]) * ms.layers
        + (MS_USE_GAP ? ms.thick_nm[0] + ms.thick_nm[1] : 0);
}

static bool ms_advance(void)                                /* nextStructure, :256-285 */
{
  ms.gap_nm += MS_GAP_STEP_NM;
  if (ms.gap_nm >= ms.thick_nm[0] + ms.thick_nm[1] || !MS_USE_GAP) {
    ms.gap_nm = MS_USE_GAP ? MS_GAP_STEP_NM : 0;
    ms.thick_nm[0] += MS_THICK0_STEP;
    ms.thick_nm[1] += MS_THICK1_STEP;
    if (ms.thick_nm[1] > MS_THICK1_LAST) {
      ms.thick_nm[0] = MS_THICK0_FIRST;
      ms.thick_nm[1] = MS_THICK1_FIRST;
      ms.edge_rate += MS_EDGE_STEP;
      if (ms.edge_rate > MS_EDGE_LAST) {
        ms.edge_rate = MS_EDGE_FIRST;
        ms.branch_nm += MS_BRANCH_STEP;
        if (ms.branch_nm > MS_BRANCH_LAST) {
          ms.layers += MS_LAYERS_STEP;
          ms.branch_nm = MS_BRANCH_FIRST;
          if (ms.layers > MS_LAYERS_LAST) {
            printf("there are no models which hasn't been simulated yet\n");
            return true;
          }
        }
      }
    }
  }
  return false;
}

static void ms_dirs(void)                                   /* :327-377 */
{
  char name[512];
  makeAndMoveDirectory(MS_ASYMMETRY ? "asymmetry" : "symmetry");
  if (MS_UNIAXIAL)
    sprintf(name, "uniaxial_ny%.2lf_nx%.2lf", MS_N0, MS_N0_X);
  else
    sprintf(name, "n_%.2lf", MS_N0);
  makeAndMoveDirectory(name);
  sprintf(name, "curve_%.2lf", MS_CURVE);                   makeAndMoveDirectory(name);
  sprintf(name, "width%d", ms.width_nm);                    makeAndMoveDirectory(name);
  sprintf(name, "thick%d_%d", ms.thick_nm[0], ms.thick_nm[1]); makeAndMoveDirectory(name);
  sprintf(name, "gap%d", ms.gap_nm);                        makeAndMoveDirectory(name);
  sprintf(name, "layer%d", ms.layers);                      makeAndMoveDirectory(name);
  sprintf(name, "edge%.1lf", ms.edge_rate);                 makeAndMoveDirectory(name);
  sprintf(name, "branch%d", ms.branch_nm);                  makeAndMoveDirectory(name);
  sprintf(name, "random%d_%d", MS_RANDOMNESS, MS_RANDOM_SEED); makeAndMoveDirectory(name);
}
const MaterialModel material_morpho = { "MorphoScaleModel", ms_select, ms_prepare,
                                        ms_size, ms_advance, ms_dirs };

/* ===================================================================== *
 *  Zig-zag slab stack  (zigzagModel.c:7-160)
 * ===================================================================== */
#define ZZ_N0 1.56
enum { ZZ_WIDTH_FIRST = 300, ZZ_WIDTH_LAST = 300, ZZ_WIDTH_STEP = 10,
       ZZ_THICK_FIRST = 80, ZZ_THICK_LAST = 150, ZZ_THICK_STEP = 10,
       ZZ_LAYERS_FIRST = 5, ZZ_LAYERS_LAST = 11, ZZ_LAYERS_STEP = 2,
       ZZ_DEG_FIRST = 10, ZZ_DEG_LAST = 80, ZZ_DEG_STEP = 10 };

static struct {
  int width_nm, thick_nm, layers, degree;
  double rad, width, thick, height, ox, oy, spill_x, spill_y, eps;
} zz = { .width_nm = ZZ_WIDTH_FIRST, .thick_nm = ZZ_THICK_FIRST,
         .layers = ZZ_LAYERS_FIRST, .degree = ZZ_DEG_FIRST };

/* zigzagModel.c:42-78: a sub-sample is inside when its squared distance to the
 * centre line of its slab is within thickness^2 */
static double zz_eps(double x_in, double y_in, int col, int row)
{
  double x = x_in - zz.ox, y = y_in - zz.oy;
  if (fabs(x) > zz.width / 2.0 + zz.spill_x + 0.5 || y < -zz.spill_y - 0.5 ||
      y > zz.height + zz.spill_y + 0.5)
    return EPSILON_0_S;

  double filled = 0;
  FOR_SUBCELL(u) {
    FOR_SUBCELL(v) {
      double sx = x + col * u / SUBCELL_SPLIT;
      double sy = y + row * v / SUBCELL_SPLIT;
      if (fabs(sx) > zz.width / 2 + zz.spill_x || sy < -zz.spill_y || sy > zz.height + zz.spill_y)
        continue;

      /* slab index; the clamp mixes int and double exactly like function.h's macros */
      int k = MPIFDTD_MIN(zz.layers - 1,
                          MPIFDTD_MAX(0, sy / (zz.width * sin(zz.rad) + cos(zz.rad) * zz.thick / 2)));
      double px = -zz.width * cos(zz.rad) / 2.0;
      double py = floor((k + 1) / 2) * 2 * (zz.width * sin(zz.rad)) + k * cos(zz.rad) * zz.thick / 2;
      double vx = (1 - ((k & 1) << 1)) * cos(zz.rad);     /* direction flips every slab */
      double vy = sin(zz.rad);

      double dist2 = pow(sx - px, 2) + pow(sy - py, 2) - pow((sx - px) * vx + (sy - py) * vy, 2);
      if (dist2 <= zz.thick * zz.thick)
        filled += 1;
    }
  }
  filled /= SUBCELL_SPLIT * SUBCELL_SPLIT;
  return EPSILON_0_S * (1 - filled) + zz.eps * filled;
}
static material_eps_fn zz_select(void) { return zz_eps; }

static void zz_prepare(void)                                /* zigzagModel.c:146-160 */
{
  zz.rad = zz.degree * M_PI / 180.0;
  zz.width = field_toCellUnit(zz.width_nm);
  zz.thick = field_toCellUnit(zz.thick_nm);
  zz.height = (zz.width * sin(zz.rad)) * zz.layers + zz.layers * zz.thick / 2 * cos(zz.rad);
  zz.spill_x = zz.thick / 2.0 * sin(zz.rad);
  zz.spill_y = zz.thick / 2.0 * cos(zz.rad);
  FieldInfo_S g = field_getFieldInfo_S();
  zz.ox = g.N_PX / 2.0;
  zz.oy = (g.N_PY - zz.height) / 2.0;
  zz.eps = ZZ_N0 * ZZ_N0 * EPSILON_0_S;
}

static void zz_size(int *x_nm, int *y_nm)                   /* zigzagModel.c:137-144 */
{
  /* always sized for the steepest angle so images of different angles line up */
  double steepest = ZZ_DEG_LAST * M_PI / 180.0;
  *x_nm = cos(steepest) * zz.width_nm + zz.thick_nm * sin(steepest);
  *y_nm = sin(steepest) * zz.width_nm * zz.layers + (1 + zz.layers) * zz.thick_nm * cos(steepest);
}

static bool zz_advance(void)                                /* zigzagModel.c:80-104 */
{
  zz.degree += ZZ_DEG_STEP;
  if (zz.degree > ZZ_DEG_LAST) {
    zz.degree = ZZ_DEG_FIRST;
    zz.thick_nm += ZZ_THICK_STEP;
    if (zz.thick_nm > ZZ_THICK_LAST) {
      zz.thick_nm = ZZ_THICK_FIRST;
      zz.width_nm += ZZ_WIDTH_STEP;
      if (zz.width_nm > ZZ_WIDTH_LAST) {
        zz.width_nm = ZZ_WIDTH_FIRST;
        zz.layers += ZZ_LAYERS_STEP;
        if (zz.layers > ZZ_LAYERS_LAST) {
          printf("there are no models which hasn't been simulated yet\n");
          return true;
        }
      }
    }
  }
  return false;
}

static void zz_dirs(void)                                   /* zigzagModel.c:116-135 */
{
  char name[512];
  sprintf(name, "n_%.2lf", ZZ_N0);                          makeAndMoveDirectory(name);
  sprintf(name, "width%d_thick%d_layer%d", zz.width_nm, zz.thick_nm, zz.layers);
  makeAndMoveDirectory(name);
  sprintf(name, "degree_%d", zz.degree);                    makeAndMoveDirectory(name);
}
const MaterialModel material_zigzag = { "ZigZagModel", zz_select, zz_prepare,
                                        zz_size, zz_advance, zz_dirs };
