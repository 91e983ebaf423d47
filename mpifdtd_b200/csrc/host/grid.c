/* grid.c -- grid, time and wave state behind the reference's field_* services.
 *
 * Host C99.  Restates WHAT field.c of rennone/mpiFDTD computes (cited per
 * function) so that every integer the solvers index with and every double they
 * feed into coefficient tables is bit-identical to the reference; the loops that
 * consumed these values on the CPU (field.c:155-256, the source injectors) live
 * in the CUDA engine instead.
 */
#define _USE_MATH_DEFINES
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mpifdtd_plugin.h"

extern void mpifdtd_flush_pending_steps(void);     /* upml_shim.c: deferred update() calls */

#ifndef M_PI
#define M_PI 3.1415926535897932384626433832795
#endif

int N_X, N_Y, N_CELL, N_PML, N_PX, N_PY;

static struct {
  FieldInfo      phys;      /* what the caller asked for (nm, degrees)   */
  FieldInfo_S    cells;     /* the same in cell units                    */
  WaveInfo_S     wave;
  NTFFInfo       farfield;
  SubFieldInfo_S sub;       /* stays zero: field.c:112 never splits      */
  double now;               /* step counter as a double (field.c:21)     */
  double last;              /* stepNum                                   */
  double ramp;              /* soft-start factor (field.c:22)            */
  double lambda_cells;
} G;

/* -- unit conversion (field.c:76-81) -------------------------------------- */
double field_toCellUnit(const double nm)      { return nm / G.phys.h_u_nm; }
double field_toPhisycalUnit(const double cell){ return cell * G.phys.h_u_nm; }

/* -- field.c:89-143 -------------------------------------------------------- */
void field_init(FieldInfo req)
{
  G.phys = req;

  /* double -> int truncation, exactly as the assignments at field.c:95-96 */
  G.cells.N_PX  = field_toCellUnit(G.phys.width_nm);
  G.cells.N_PY  = field_toCellUnit(G.phys.height_nm);
  G.cells.N_PML = G.phys.pml;
  G.cells.N_X   = G.cells.N_PX - 2 * G.cells.N_PML;
  G.cells.N_Y   = G.cells.N_PY - 2 * G.cells.N_PML;
  /* wraps for grids beyond 2^31 cells (slab runs never index with it) */
  G.cells.N_CELL = (int)((unsigned)G.cells.N_PY * (unsigned)G.cells.N_PX);
  G.cells.DX = G.cells.N_PY;
  G.cells.DY = 1;

  G.wave.Lambda_s  = field_toCellUnit(G.phys.lambda_nm);
  G.wave.T_s       = G.wave.Lambda_s / C_0_S;
  G.wave.K_s       = 2 * M_PI / G.wave.Lambda_s;
  G.wave.K_0_s     = 2 * M_PI / G.wave.Lambda_s;
  G.wave.Omega_s   = C_0_S * G.wave.K_s;
  G.wave.Angle_deg = G.phys.angle_deg;
  G.lambda_cells   = G.wave.Lambda_s;

  N_PX = G.cells.N_PX;  N_PY = G.cells.N_PY;  N_PML = G.cells.N_PML;
  N_X = G.cells.N_X;    N_Y = G.cells.N_Y;    N_CELL = G.cells.N_CELL;

  G.now = 0;
  G.last = G.phys.stepNum;
  G.ramp = 0;
  memset(&G.sub, 0, sizeof G.sub);

  /* closed NTFF surface 5 cells inside the PML (field.c:132-142).  The radius
   * term comes from the y extent only, halved in integer arithmetic. */
  NTFFInfo *f = &G.farfield;
  f->cx = N_PX / 2;
  f->cy = N_PY / 2;
  f->top    = N_PY - N_PML - 5;
  f->bottom = N_PML + 5;
  f->left   = N_PML + 5;
  f->right  = N_PX - N_PML - 5;
  double half = (f->top - f->bottom) / 2;      /* int / int, then widened */
  f->RFperC = half * 2;
  f->arraySize = G.last + 2 * f->RFperC;
}

void field_reset(void) { G.now = 0; G.ramp = 0; }                 /* field.c:83  */

void field_nextStep(void)                                         /* field.c:312 */
{
  G.now += 1.0;
  G.ramp = 1.0 - exp(-pow(0.01 * G.now, 2));
}

bool field_isFinish(void) { return G.now >= G.last; }             /* field.c:317 */

void field_setWaveAngle(int deg)                                  /* field.c:39  */
{
  mpifdtd_flush_pending_steps();    /* deferred update() calls were made under the old angle */
  G.phys.angle_deg = deg;
  G.wave.Angle_deg = deg;
}

/* -- getters (field.c:44-74) ----------------------------------------------- */
double field_getT(void)         { return G.wave.T_s; }
double field_getK(void)         { return G.wave.K_s; }
double field_getRayCoef(void)   { return G.ramp; }
double field_getOmega(void)     { return G.wave.Omega_s; }
double field_getLambda(void)    { return G.lambda_cells; }
double field_getWaveAngle(void) { return G.wave.Angle_deg; }
double field_getTime(void)      { return G.now; }
double field_getMaxTime(void)   { return G.last; }
NTFFInfo field_getNTFFInfo(void)             { return G.farfield; }
WaveInfo_S field_getWaveInfo_S(void)         { return G.wave; }
SubFieldInfo_S field_getSubFieldInfo_S(void) { return G.sub; }
FieldInfo_S field_getFieldInfo_S(void)       { return G.cells; }
FieldInfo field_getFieldInfo(void)           { return G.phys; }
int field_getOffsetX(void)  { return G.sub.OFFSET_X; }
int field_getOffsetY(void)  { return G.sub.OFFSET_Y; }
int field_getSubNx(void)    { return G.sub.SUB_N_X; }
int field_getSubNy(void)    { return G.sub.SUB_N_Y; }
int field_getSubNpx(void)   { return G.sub.SUB_N_PX; }
int field_getSubNpy(void)   { return G.sub.SUB_N_PY; }
int field_getSubNcell(void) { return G.sub.SUB_N_CELL; }
int field_subIndex(int i, int j) { return i * G.sub.SUB_N_PY + j; }
int field_index(int i, int j)    { return i * G.cells.N_PY + j; }
int ind(const int i, const int j){ return i * N_PY + j; }

/* -- PML conductivity profile, polynomial order 2 (field.c:259-283) --------
 * Three zones per axis: the low-side layer, the PML-free interior, the
 * high-side layer.  pow() is called exactly like the reference does, so the
 * value matches the reference's libm result bit for bit. */
static double pml_profile(double u, int n_pml, int n_inner, int n_total)
{
  const int order = 2;
  if (u < n_pml)
    return pow(1.0 * (n_pml - u) / n_pml, order);
  if (n_pml <= u && u < (n_inner + n_pml))
    return 0;
  return pow(1.0 * (u - (n_total - n_pml - 1)) / n_pml, order);
}
double field_sigmaX(double x, double y) { (void)y; return pml_profile(x, N_PML, N_X, N_PX); }
double field_sigmaY(double x, double y) { (void)x; return pml_profile(y, N_PML, N_Y, N_PY); }

/* -- split-field PML coefficient shapes, dt = 1 (field.c:288-295) ---------- */
double field_pmlCoef(double ep_mu, double sig)     { return (1.0 - sig / ep_mu) / (1.0 + sig / ep_mu); }
double field_pmlCoef_LXY(double ep_mu, double sig) { return 1.0 / (ep_mu + sig); }

/* -- NS-PML beta terms (field.c:298-309) ------------------------------------ */
double field_ns_beta(double alpha, double alpha_aster)
{
  double ta = tanh(alpha), tb = tanh(alpha_aster);
  return ta / (1 + ta * tb);
}
double field_ns_beta_aster(double alpha, double alpha_aster)
{
  double ta = tanh(alpha), tb = tanh(alpha_aster);
  return tb / (1 + ta * tb);
}

/* -- soft-started CW point source value (field.c:145-152) ------------------- */
dcomplex field_pointLight(void)
{
  return G.ramp * cexp(I * G.wave.Omega_s * G.now);
}

/* -- text dumps of host mirrors (field.c:322-388) --------------------------- */
static FILE *open_or_die(const char *name)
{
  FILE *fp = fopen(name, "w");
  if (fp == NULL) { printf("cannot open file %s \n", name); exit(2); }
  return fp;
}

void field_outputElliptic(const char *fileName, double complex *data)
{
  printf("output start\n");
  FILE *fp = open_or_die(fileName);
  /* |F|^2 sampled bilinearly on a circle of radius 1.2 lambda, 180 -> 0 degrees */
  for (int ang = 180; ang >= 0; ang--) {
    double rad = ang * M_PI / 180.0;
    double x = 1.2 * G.lambda_cells * cos(rad) + N_PX / 2.0;
    double y = 1.2 * G.lambda_cells * sin(rad) + N_PY / 2.0;
    fprintf(fp, "%d %lf \n", 180 - ang, cnorm(cbilinear(data, x, y, N_PX, N_PY)));
  }
  fclose(fp);
  printf("output to %s end\n", fileName);
}

void field_outputAllDataComplex(const char *fileName, double complex *data)
{
  printf("output all data start\n");
  FILE *fp = open_or_die(fileName);
  for (int k = 0; k < N_PX * N_PY; k++)
    fprintf(fp, "%lf \n", cnorm(data[k]));
  fclose(fp);
  printf("output all data to %s end\n", fileName);
}

void field_outputAllDataDouble(const char *fileName, double *data)
{
  printf("output all data double start\n");
  FILE *fp = open_or_die(fileName);
  for (int k = 0; k < N_PX * N_PY; k++)
    if (data[k] != 1.0)
      fprintf(fp, "%lf \n", data[k]);
  fclose(fp);
  printf("output all data to %s end\n", fileName);
}
