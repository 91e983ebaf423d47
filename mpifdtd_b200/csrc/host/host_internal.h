/* host_internal.h -- host-side helpers shared between the plugin translation
 * units (not part of the exported surface, but visible for tests through the
 * shared library). */
#ifndef MPIFDTD_HOST_INTERNAL_H
#define MPIFDTD_HOST_INTERNAL_H
#include "mpifdtd_plugin.h"

/* dst[i*N_PY+j] = models_eps(i+xoff, j+yoff, mode) over the whole grid, threaded */
extern void mpifdtd_fill_eps(double *dst, double xoff, double yoff, enum MODE mode);

#endif
