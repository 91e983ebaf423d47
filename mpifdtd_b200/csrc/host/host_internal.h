/* host_internal.h -- host-side helpers shared between the plugin translation
 * units (not part of the exported surface, but visible for tests through the
 * shared library). */
#ifndef MPIFDTD_HOST_INTERNAL_H
#define MPIFDTD_HOST_INTERNAL_H
#include "mpifdtd_plugin.h"

/* dst[i*N_PY+j] = models_eps(i+xoff, j+yoff, mode) over the whole grid, threaded */
extern void mpifdtd_fill_eps(double *dst, double xoff, double yoff, enum MODE mode);
/* slab form: dst[i*nj + (j-j0)] for j in [j0, j0+nj) */
extern void mpifdtd_fill_eps_slab(double *dst, double xoff, double yoff, enum MODE mode, int j0, int nj);


/* map[n] -> 16-bit indices into a table of its distinct values (table: room for 65536 doubles);
 * returns the number of distinct values, -1 if there are more than 65536 */
#include <stddef.h>
#include <stdint.h>
extern int mpifdtd_eps_palette(const double *map, size_t n, uint16_t *index, double *table);

/* farfield.c: NTFF sampling plan and post-processing constants (host libm) */
extern int mpifdtd_ntff_point_count(const NTFFInfo *box);
extern int mpifdtd_ntff_local_count(const NTFFInfo *box, int j0, int nj);
extern double *mpifdtd_ntff_time_shift(const NTFFInfo *box, int n_angles, double stagger, int j0, int nj);
extern int mpifdtd_angle_batch_requested(const int **angles_deg);
extern void mpifdtd_split_select_angle(int index);
extern int mpifdtd_ntff_local_count_shifted(const NTFFInfo *box, int sample_dj, int j0, int nj);
extern double *mpifdtd_ntff_time_shift_direct(const NTFFInfo *box, int n_angles, double stagger, int sample_dj, int j0, int nj);
extern double complex mpifdtd_ntff_translate_coef(double omega);
extern void mpifdtd_ntff_direction_cosines(int n_angles, int is_tm, double *cos_phi, double *sin_phi);
extern double complex *mpifdtd_fft_twiddles(int n);

/* upml_shim.c */
/* hands every update() the serial UPML solvers have only counted so far to the engine */
extern void mpifdtd_flush_pending_steps(void);
extern void mpifdtd_split_flush_pending_steps(void);
#include "b200fdtd.h"
extern void mpifdtd_upml_step_args(int kind, int point_source, b200fdtd_step_args *a);
extern void mpifdtd_upml_far_field(b200fdtd_engine *engine, int kind, int project, double *table);
extern void mpifdtd_upml_step_args_form(int kind, int point_source, int form, b200fdtd_step_args *a);
/* E_theta / E_phi [360][maxTime] of the MPI TE solver (mpiTE_UPML.c:840-878) */
extern int mpifdtd_mpi_te_far_series(dcomplex *eth, dcomplex *eph);
extern int mpifdtd_upml_dense_coefficient(int kind, const char *name, double *dst);
/* the twelve 1-D tables of include/b200fdtd.h for the current field_init() state */
extern void mpifdtd_upml_tables(int kind, double *tab_i, double *tab_j);


/* split_shim.c */
extern void mpifdtd_split_step_args(int kind, b200fdtd_step_args *a);
extern void mpifdtd_split_prepare_host(int kind);
extern const double *mpifdtd_split_dense(int kind, int slot);
extern void mpifdtd_split_prepare_host_lean(int kind);
extern void mpifdtd_split_lean_tables(int kind, double *tab_i, double *tab_j);
extern b200fdtd_engine *mpifdtd_split_engine(int kind);

#endif
