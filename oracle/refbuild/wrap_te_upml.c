#include "fdtdTE_upml.c"
#include "hook.h"
HOOK_BEGIN(refhook_te_upml)
HOOK(Ex) HOOK(Jx) HOOK(Dx) HOOK(Ey) HOOK(Jy) HOOK(Dy) HOOK(Hz) HOOK(Mz) HOOK(Bz)
HOOK(Wx) HOOK(Wy) HOOK(Uz)
HOOK(C_JX) HOOK(C_JY) HOOK(C_MZ) HOOK(C_JXHZ) HOOK(C_JYHZ) HOOK(C_MZEXEY)
HOOK(C_DX) HOOK(C_DY) HOOK(C_BZ) HOOK(C_DXJX0) HOOK(C_DXJX1)
HOOK(C_DYJY0) HOOK(C_DYJY1) HOOK(C_BZMZ0) HOOK(C_BZMZ1)
HOOK(EPS_EX) HOOK(EPS_EY) HOOK(EPS_HZ)
HOOK_END
void refhook_te_upml_update_no_ntff(void)
{
  calcMB(); calcH(); calcJD(); calcE();
  WaveInfo_S wInfo = field_getWaveInfo_S();
  double co = cos( (wInfo.Angle_deg+90) * M_PI/ 180.0);
  double si = sin( (wInfo.Angle_deg+90) * M_PI/ 180.0);
  if(co != 0.0) field_scatteredPulse(Ex, EPS_EX, 0.5, 0.0, co);
  if(si != 0.0) field_scatteredPulse(Ey, EPS_EY, 0.0, 0.5, si);
}
