#include "fdtdTM_upml.c"
#include "hook.h"
HOOK_BEGIN(refhook_tm_upml)
HOOK(Ez) HOOK(Jz) HOOK(Dz) HOOK(Hx) HOOK(Mx) HOOK(Bx) HOOK(Hy) HOOK(My) HOOK(By)
HOOK(Ux) HOOK(Uy) HOOK(Wz)
HOOK(C_JZ) HOOK(C_MX) HOOK(C_MY) HOOK(C_JZHXHY) HOOK(C_MXEZ) HOOK(C_MYEZ)
HOOK(C_DZ) HOOK(C_BX) HOOK(C_BY) HOOK(C_DZJZ0) HOOK(C_DZJZ1)
HOOK(C_BXMX0) HOOK(C_BXMX1) HOOK(C_BYMY0) HOOK(C_BYMY1)
HOOK(EPS_EZ) HOOK(EPS_HX) HOOK(EPS_HY)
HOOK_END
/* stencil + source only (the reference's update() minus its NTFF call), used to
 * time the CPU stencil apples-to-apples with the GPU kernels */
void refhook_tm_upml_update_no_ntff(void)
{
  calcMB(); calcH(); calcJD(); calcE();
  field_scatteredPulse(Ez, EPS_EZ, 0, 0, 1.0);
}
/* opt-in point source for the NoModel configuration: update() with
 * field_pointLight() (field.c:145-152, no caller in the reference) added to
 * Ez at the domain centre before the NTFF sampling */
void refhook_tm_upml_update_point_source(void)
{
  calcMB(); calcH(); calcJD(); calcE();
  field_scatteredPulse(Ez, EPS_EZ, 0, 0, 1.0);
  Ez[field_index(N_PX/2, N_PY/2)] += field_pointLight();
  ntffTM_TimeCalc(Hx,Hy,Ez,Ux,Uy,Wz);
}
/* update() with the alternative source the reference keeps commented out at
 * fdtdTM_upml.c:62: field_scatteredWave (field.c:202-218) instead of the pulse */
void refhook_tm_upml_update_cw(void)
{
  calcMB(); calcH(); calcJD(); calcE();
  field_scatteredWave(Ez, EPS_EZ, 0, 0);
  ntffTM_TimeCalc(Hx,Hy,Ez,Ux,Uy,Wz);
}
/* opt-in plane-wave line source for the NoModel configuration: the expression of
 * planeWave (mpiTM_UPML.c:377-403, no caller) on the serial grid -- row i = NTFF left
 * edge, columns 1..N_PY-2, x = i, y = j -- added to Ez after the pulse */
void refhook_tm_upml_update_plane_wave(void)
{
  calcMB(); calcH(); calcJD(); calcE();
  field_scatteredPulse(Ez, EPS_EZ, 0, 0, 1.0);
  {
    const double time = field_getTime();
    const double w_s  = field_getOmega();
    const double ray_coef = field_getRayCoef();
    const double k_s = field_getK();
    const double rad = field_getWaveAngle()*M_PI/180;
    const double ks_cos = cos(rad)*k_s, ks_sin = sin(rad)*k_s;
    NTFFInfo nInfo = field_getNTFFInfo();
    const int x = nInfo.left;
    for (int y = 1; y < N_PY-1; y++) {
      double kr = (x*ks_cos + y*ks_sin) - time;
      Ez[field_index(x, y)] += ray_coef*cexp(I*kr*w_s);
    }
  }
  ntffTM_TimeCalc(Hx,Hy,Ez,Ux,Uy,Wz);
}
