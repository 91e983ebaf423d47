#include "mpiTE_UPML.c"
#include "hook.h"
HOOK_BEGIN(refhook_mpi_te_upml)
HOOK(Ex) HOOK(Jx) HOOK(Dx) HOOK(Ey) HOOK(Jy) HOOK(Dy) HOOK(Hz) HOOK(Mz) HOOK(Bz)
HOOK(Wx) HOOK(Wy) HOOK(Uz)
HOOK(C_JX) HOOK(C_JY) HOOK(C_MZ) HOOK(C_JXHZ) HOOK(C_JYHZ) HOOK(C_MZEXEY)
HOOK(C_DX) HOOK(C_DY) HOOK(C_BZ) HOOK(C_DXJX0) HOOK(C_DXJX1)
HOOK(C_DYJY0) HOOK(C_DYJY1) HOOK(C_BZMZ0) HOOK(C_BZMZ1)
HOOK(EPS_EX) HOOK(EPS_EY) HOOK(EPS_HZ)
HOOK_END

/* finish() of this solver writes MPI_TE_UPML/E{ph,th}_{r,i}.txt through ntffOutput() and then
 * dereferences its debug arrays, which are NULL unless built with -DDEBUG; it also never
 * closes the files.  This hook gives the debug pointers zeroed storage, runs the unmodified
 * ntffOutput() and flushes, so tests can read the reference's own files. */
void refhook_mpi_te_upml_ntff_output(void)
{
  NTFFInfo info = field_getNTFFInfo();
  for (int i = 0; i < 4; i++) {
    if (!debug_U[i]) debug_U[i] = (double complex *)calloc((size_t)360 * info.arraySize, sizeof(double complex));
    if (!debug_W[i]) debug_W[i] = (double complex *)calloc((size_t)360 * info.arraySize, sizeof(double complex));
  }
  ntffOutput();
  fflush(NULL);
}
