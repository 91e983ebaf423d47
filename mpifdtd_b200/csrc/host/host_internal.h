/* host_internal.h -- host-side helpers shared between the plugin translation
 * units (not part of the exported surface, but visible for tests through the
 * shared library). */
#ifndef MPIFDTD_HOST_INTERNAL_H
#define MPIFDTD_HOST_INTERNAL_H
#include "mpifdtd_plugin.h"

/* dst[i*N_PY+j] = models_eps(i+xoff, j+yoff, mode) over the whole grid, threaded */
extern void mpifdtd_fill_eps(double *dst, double xoff, double yoff, enum MODE mode);


/* farfield.c: NTFF sampling plan and post-processing constants (host libm) */
extern int mpifdtd_ntff_point_count(const NTFFInfo *box);
extern double *mpifdtd_ntff_time_shift(const NTFFInfo *box, int n_angles, double stagger);
extern double complex mpifdtd_ntff_translate_coef(double omega);
extern void mpifdtd_ntff_direction_cosines(int n_angles, int is_tm, double *cos_phi, double *sin_phi);
extern double complex *mpifdtd_fft_twiddles(int n);

/* upml_shim.c */
extern void mpifdtd_enablePointSource(int on);
extern int mpifdtd_upml_dense_coefficient(int kind, const char *name, double *dst);

#endif
