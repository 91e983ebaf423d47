#include "mpiTM_UPML.c"
#include "hook.h"
HOOK_BEGIN(refhook_mpi_tm_upml)
HOOK(Ez) HOOK(Jz) HOOK(Dz) HOOK(Hx) HOOK(Mx) HOOK(Bx) HOOK(Hy) HOOK(My) HOOK(By)
HOOK(Ux) HOOK(Uy) HOOK(Wz)
HOOK(C_JZ) HOOK(C_MX) HOOK(C_MY) HOOK(C_JZHXHY) HOOK(C_MXEZ) HOOK(C_MYEZ)
HOOK(C_DZ) HOOK(C_BX) HOOK(C_BY) HOOK(C_DZJZ0) HOOK(C_DZJZ1)
HOOK(C_BXMX0) HOOK(C_BXMX1) HOOK(C_BYMY0) HOOK(C_BYMY1)
HOOK(EPS_EZ) HOOK(EPS_HX) HOOK(EPS_HY)
HOOK_END
/* update() (mpiTM_UPML.c:196-217) with the commented planeWave call at line 204 enabled */
void refhook_mpi_tm_upml_update_plane_wave(void)
{
  calcJD(); calcE();
  scatteredWave(Ez, EPS_EZ);
  planeWave(Ez, EPS_EZ);
  Connection_ISend_IRecvE();
  calcMB(); calcH();
  Connection_ISend_IRecvH();
  ntff();
}
