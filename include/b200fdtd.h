/* b200fdtd.h -- C ABI of the B200 (sm_100a) FDTD time-stepping engine.
 *
 * This is the drop-in boundary for the hot path of rennone/mpiFDTD: everything
 * the reference does per time step inside a solver's update() -- the UPML E/H
 * leapfrog with its J/D and M/B recurrences, source injection, and the NTFF
 * surface accumulation -- plus the once-per-run far-field post-processing.
 * Plain pointers and sizes only; no CUDA or torch types.  The host-side C shims
 * in mpifdtd_b200/csrc/host (symbols fdtdTM_upml_getUpdate etc., see
 * mpifdtd_plugin.h) call these functions; other hosts (ctypes, cgo, JNI) can
 * bind them directly.
 *
 * All functions return B200FDTD_OK (0) or a B200FDTD_ERR_* code;
 * b200fdtd_last_error() returns a message for the calling thread's last
 * failure.  There is no CPU fallback: without a CUDA device every entry point
 * that would compute returns B200FDTD_ERR_NODEVICE.
 *
 * Index convention: host arrays are the reference's, k = i*N_PY + j (x slow, y
 * fast; field.c:70-72).  Complex values are interleaved (re, im) doubles, the
 * memory layout of C99 double complex.
 */
#ifndef B200FDTD_H
#define B200FDTD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200FDTD_ABI_VERSION 7

enum {
  B200FDTD_OK = 0,
  B200FDTD_ERR_ARG = 1,       /* bad argument / unsupported combination       */
  B200FDTD_ERR_CUDA = 2,      /* a CUDA runtime call failed                    */
  B200FDTD_ERR_NOMEM = 3,     /* host or device allocation failed              */
  B200FDTD_ERR_STATE = 4,     /* call order violated (e.g. step before tables) */
  B200FDTD_ERR_NODEVICE = 5   /* no usable CUDA device                         */
};

/* Solver kinds; numeric values are enum SOLVER of simulator.h:8-18. */
enum {
  B200FDTD_TM = 0, B200FDTD_TE = 1,
  B200FDTD_TM_UPML = 2, B200FDTD_TE_UPML = 3,
  B200FDTD_MPI_TM_UPML = 4, B200FDTD_MPI_TE_UPML = 5,
  B200FDTD_NS_TM = 6, B200FDTD_NS_TE = 7
};

/* Field slots.  UPML kinds keep 9 complex arrays (fdtdTM_upml.c:15-25,
 * fdtdTE_upml.c:15-25): */
enum { B200FDTD_TM_EZ = 0, B200FDTD_TM_JZ, B200FDTD_TM_DZ, B200FDTD_TM_HX, B200FDTD_TM_MX,
       B200FDTD_TM_BX, B200FDTD_TM_HY, B200FDTD_TM_MY, B200FDTD_TM_BY };
enum { B200FDTD_TE_EX = 0, B200FDTD_TE_JX, B200FDTD_TE_DX, B200FDTD_TE_EY, B200FDTD_TE_JY,
       B200FDTD_TE_DY, B200FDTD_TE_HZ, B200FDTD_TE_MZ, B200FDTD_TE_BZ };
#define B200FDTD_MAX_FIELDS 9
enum { B200FDTD_F64 = 0, B200FDTD_F32 = 1 };
/* Split-field kinds (plain Berenger-PML Yee 0/1 and NS-FDTD 6/7) keep 5 complex arrays
 * (fdtdTM.c:10-14, fdtdTE.c:10-14, nsFdtdTM.c:10-14, nsFdtdTE.c:11-15): */
enum { B200FDTD_STM_EZ = 0, B200FDTD_STM_EZX, B200FDTD_STM_EZY, B200FDTD_STM_HX, B200FDTD_STM_HY };
enum { B200FDTD_STE_HZ = 0, B200FDTD_STE_HZX, B200FDTD_STE_HZY, B200FDTD_STE_EX, B200FDTD_STE_EY };
/* Their coefficients depend on the permittivity and are therefore genuinely 2-D; the host
 * builds them with the reference's expressions and uploads them as dense arrays
 * (b200fdtd_set_dense).  Slots 0-7 in the order the reference declares them: */
enum { B200FDTD_STM_C_EZX = 0, B200FDTD_STM_C_EZXLX, B200FDTD_STM_C_EZY, B200FDTD_STM_C_EZYLY,
       B200FDTD_STM_C_HX, B200FDTD_STM_C_HXLY, B200FDTD_STM_C_HY, B200FDTD_STM_C_HYLX };
enum { B200FDTD_STE_C_EX = 0, B200FDTD_STE_C_EXLY, B200FDTD_STE_C_EY, B200FDTD_STE_C_EYLX,
       B200FDTD_STE_C_HZX, B200FDTD_STE_C_HZXLX, B200FDTD_STE_C_HZY, B200FDTD_STE_C_HZYLY };
/* slots 8, 9: per-cell factor of the scattered-field CW source on target 0 / 1, i.e.
 * (eps0/eps - 1) (field.c:193) or (1/(_n*n) - 1) for NS-FDTD (field.c:168-174) */
#define B200FDTD_DENSE_SRC0 8
#define B200FDTD_DENSE_SRC1 9
#define B200FDTD_MAX_DENSE 10

/* "Lean" form of the split-field kinds 0, 1 and 6 (b200fdtd_set_split_tables): most of the
 * eight coefficients are separable after all.  Berenger (fdtdTM.c:197-242, fdtdTE.c:199-243):
 * the H coefficients use MU_0_S and a sigma that depends on i only or j only -> 1-D tables; the
 * E coefficients are field_pmlCoef(eps, sigma) / field_pmlCoef_LXY(eps, sigma) of the cell's
 * eps and a 1-D sigma, which the kernel evaluates itself with the reference's operations
 * (exactly 1 and 1/eps outside the PML, three IEEE divisions inside).  NS-FDTD TM
 * (nsFdtdTM.c:231-307): the decay coefficients coef1(beta) are 1-D; the curl coefficients are
 * (u*z)(eps) / (1 + beta(sigma)) -> one dense per-cell array G (host libm) divided by a 1-D
 * table in the kernel.  Per cell-update this reads eps (8 B, kinds 0/1) or three G arrays +
 * the source factor (32 B, kind 6) instead of nine dense doubles (72 B).  Bit-identical to the
 * dense form.  NS-FDTD TE (kind 7) keeps the dense form: its beta depends on eps inside the PML
 * (nsFdtdTE.c:130-137) through tanh, which only the host libm reproduces.
 * tab_i[slot*n_px + i], tab_j[slot*n_py + j]: */
enum { /* kind 0, by i */ B200FDTD_LTM_I_SIG_EZ_X = 0, B200FDTD_LTM_I_C_HY, B200FDTD_LTM_I_C_HYLX,
       /* kind 0, by j */ B200FDTD_LTM_J_SIG_EZ_Y = 0, B200FDTD_LTM_J_C_HX, B200FDTD_LTM_J_C_HXLY };
enum { /* kind 1, by i */ B200FDTD_LTE_I_SIG_EY_X = 0, B200FDTD_LTE_I_C_HZX, B200FDTD_LTE_I_C_HZXLX,
       /* kind 1, by j */ B200FDTD_LTE_J_SIG_EX_Y = 0, B200FDTD_LTE_J_C_HZY, B200FDTD_LTE_J_C_HZYLY };
enum { /* kind 6, by i */ B200FDTD_LNS_I_C_EZX = 0, B200FDTD_LNS_I_DEN_EZ /* 1 + b_ez_x */, B200FDTD_LNS_I_C_HY,
       B200FDTD_LNS_I_DEN_HY /* 1 + b_hy_x */,
       /* kind 6, by j */ B200FDTD_LNS_J_C_EZY = 0, B200FDTD_LNS_J_C_HX, B200FDTD_LNS_J_DEN_HX /* 1 + b_hx_y */ };
/* kind 6 lean: dense slots C_EZXLX, C_HXLY, C_HYLX hold G_EZ = u_ez*z_ez, G_HX = u_hx/z_hx,
 * G_HY = u_hy/z_hy; slot 8 the source factor; the other dense slots are not uploaded */
#define B200FDTD_SPLIT_TABS 4

/* 1-D coefficient tables of the UPML kinds.  The reference stores 15 dense
 * N_CELL arrays per solver (fdtdTM_upml.c:30-35); because every one of them is
 * built with eps = EPSILON_0_S (fdtdTM_upml.c:253, fdtdTE_upml.c:390) each is a
 * function of i only, of j only, exactly 1, or a j-term divided by an i-term.
 * The host passes the i-terms as tab_i[slot*n_px + i] and the j-terms as
 * tab_j[slot*n_py + j], computed with the reference's own expressions. */
enum { /* TM, by i */ B200FDTD_TMI_C_JZ = 0, B200FDTD_TMI_C_JZHXHY, B200FDTD_TMI_C_BXMX1,
       B200FDTD_TMI_C_BXMX0, B200FDTD_TMI_C_BY, B200FDTD_TMI_DEN_BYMY,   /* 2eps + sig_hy_x */
       /* TM, by j */ B200FDTD_TMJ_C_DZ = 0, B200FDTD_TMJ_C_DZJZ, B200FDTD_TMJ_C_MX,
       B200FDTD_TMJ_C_MXEZ, B200FDTD_TMJ_NUM_BYMY1 /* 2eps + sig_hy_y */,
       B200FDTD_TMJ_NUM_BYMY0 /* 2eps - sig_hy_y */ };
enum { /* TE, by i */ B200FDTD_TEI_C_DXJX1 = 0, B200FDTD_TEI_C_DXJX0, B200FDTD_TEI_C_DY,
       B200FDTD_TEI_DEN_DYJY /* 2eps + sig_ey_x */, B200FDTD_TEI_C_MZ, B200FDTD_TEI_C_MZEXEY,
       /* TE, by j */ B200FDTD_TEJ_C_JX = 0, B200FDTD_TEJ_C_JXHZ, B200FDTD_TEJ_NUM_DYJY1,
       B200FDTD_TEJ_NUM_DYJY0, B200FDTD_TEJ_C_BZ, B200FDTD_TEJ_C_BZMZ };
#define B200FDTD_UPML_TABS 6

typedef struct b200fdtd_engine b200fdtd_engine;   /* opaque */

/* Geometry of one engine = one y-slab of the global grid on one GPU.
 * Replaces field_init's FieldInfo_S (field.c:94-121) plus, for multi-GPU runs,
 * the MPI sub-domain of mpiTM_UPML.c:718-748 (here a 1-D split along y). */
typedef struct b200fdtd_grid {
  int32_t kind;             /* B200FDTD_* solver kind                              */
  int32_t n_px, n_py;       /* global cells incl. PML                              */
  int32_t n_pml;
  int32_t j0, nj;           /* this engine owns global columns j in [j0, j0+nj)    */
  int32_t i_lo, i_hi;       /* updated cells, inclusive, global; the serial solvers */
  int32_t j_lo, j_hi;       /*   use 1 .. N-2 (fdtdTM_upml.c:158-159)              */
  int32_t device;           /* CUDA device ordinal, or -1 for the current device   */
  int32_t precision;        /* B200FDTD_F64 (0, the reference's arithmetic) or      */
                            /* B200FDTD_F32: complex64 fields, f32 eps/coefficients  */
                            /* (UPML kinds; own tolerance)            */
  double mu0;               /* MU_0_S, passed so the divisor is the host's value   */
  int32_t n_batch;          /* independent simulations sharing this grid, eps and   */
                            /* coefficients and differing only in their source      */
                            /* (incidence angle): the angle sweep of main.c:114-211  */
                            /* as ONE engine.  0 or 1 = a single simulation.  Serial */
                            /* UPML kinds (2, 3), whole grid on one engine.          */
  int32_t flags;            /* B200FDTD_GRID_* bits, 0 by default                     */
} b200fdtd_grid;
/* An engine without field arrays: only the NTFF machinery (history, projection, spectrum), fed with
 * surface samples the caller gathered from its OWN host-side fields (b200fdtd_ntff_push_samples) --
 * what the exported ntffTM_* / ntffTE_* entry points run on.  Stepping such an engine is an error. */
#define B200FDTD_GRID_NTFF_ONLY 1

/* Scattered-field Gaussian pulse, field_scatteredPulse (field.c:224-256):
 * p[k] += dot*exp(-(r/beam_width)^2)*(eps0/eps[k]-1)*cexp(i*r*omega) on cells with
 * eps != eps0, r = (i+gap_x)*cos_per_c + (j+gap_y)*sin_per_c - time_minus_t0. */
typedef struct b200fdtd_pulse {
  int32_t enabled, reserved;
  double gap_x, gap_y, dot;
  double cos_per_c, sin_per_c;     /* cos(rad)/C_0_S, sin(rad)/C_0_S from host libm  */
  double time_minus_t0;            /* (time - t0), t0 = -center_peak + 500            */
  double omega, beam_width;
} b200fdtd_pulse;

/* Per-simulation source of a batched engine (b200fdtd_grid.n_batch > 1): the pulse
 * parameters that depend on the incidence angle.  pulse[m].time_minus_t0 is ignored; the
 * kernels form it as step_args.time - t0[m], the very subtraction the host does for a single
 * simulation (field.c:241,251).  Uploaded once per sweep with b200fdtd_set_batch_sources. */
typedef struct b200fdtd_batch_source {
  b200fdtd_pulse pulse[2];
  double t0[2];
} b200fdtd_batch_source;

/* Per-simulation part of the continuous-wave source of a batched split-field engine (kinds 0, 1,
 * 6, 7): what of b200fdtd_cw depends on the incidence angle.  step_args.cw[m] then carries what the
 * simulations share -- gaps, phases, two_term, and scale = ray_coef WITHOUT the polarisation
 * factor; the kernels form scale * dot[m] (the host's own multiplication, field.c:160) and take
 * enabled[m], ks_cos, ks_sin from here.  Uploaded once per sweep with b200fdtd_set_batch_cw. */
typedef struct b200fdtd_batch_cw {
  double ks_cos, ks_sin;           /* cos(rad)*k_s, sin(rad)*k_s from host libm        */
  double dot[2];                   /* 1.0, or cos / sin(angle + 90 deg) for kind 7     */
  int32_t enabled[2];
} b200fdtd_batch_cw;

/* Opt-in soft-started point source for the NoModel configuration:
 * value field_pointLight() (field.c:145-152) added to E-slot 0 at (i, j). */
typedef struct b200fdtd_point_source {
  int32_t enabled, i, j, reserved;
  double re, im;
} b200fdtd_point_source;

/* Scattered-field continuous wave.  p[k] += scale * factor[k] * (cexp(i(kr - phase_a))
 * [- cexp(i(kr - phase_b)) when two_term]), kr = (i+gap_x)*ks_cos + (j+gap_y)*ks_sin.
 *   field_scatteredWaveNotUPML   (field.c:179-196): two_term, phases w(t+1/2), w(t-1/2)
 *   field_nsScatteredWaveNotUPML (field.c:155-177): two_term, phases w(t+1),   w t
 *   scatteredWave of the MPI solvers (mpiTM_UPML.c:337-374): one term, phase w t, gaps 0
 * scale = ray_coef (* dot).  factor[k] comes from dense slot 8/9 for the split-field kinds
 * and is (eps0/eps[k] - 1) evaluated in the kernel for the UPML kinds. */
typedef struct b200fdtd_cw {
  int32_t enabled, two_term;
  double gap_x, gap_y, scale;
  double ks_cos, ks_sin;
  double phase_a, phase_b;
} b200fdtd_cw;

/* planeWave of mpiTM_UPML.c:377-403 (no caller upstream: the call at mpiTM_UPML.c:204 is
 * commented out; opt-in here): on the cells (i, j_lo..j_hi) of one grid row,
 *   p += scale * cexp(I * ((i*ks_cos + j*ks_sin) - time) * omega)
 * with i, j GLOBAL cell indices.  TM kinds, target Ez. */
typedef struct b200fdtd_line_source {
  int32_t enabled, i, j_lo, j_hi;
  double scale;                    /* field_getRayCoef()                              */
  double ks_cos, ks_sin;           /* cos(rad)*k_s, sin(rad)*k_s                       */
  double time, omega;
} b200fdtd_line_source;

/* Everything one update() call depends on that the host owns (field.c:44-51). */
typedef struct b200fdtd_step_args {
  double time;                     /* field_getTime() BEFORE field_nextStep()         */
  double ray_coef;                 /* field_getRayCoef()                              */
  b200fdtd_pulse pulse[2];         /* TM: [0] on Ez.  TE: [0] on Ex, [1] on Ey        */
  b200fdtd_point_source point;
  b200fdtd_cw cw[2];               /* CW sources: target 0 / 1 of the solver kind     */
  double ns_r2;                    /* NS-FDTD: r/2 of the 9-point operator             */
                                   /* (nsFdtdTM.c:115-117), 0 otherwise               */
  b200fdtd_line_source line;       /* opt-in plane-wave line source (TM UPML kinds)    */
} b200fdtd_step_args;

/* Closed NTFF surface (NTFFInfo, field.h:62-68) and its sampling plan.
 * Perimeter points are ordered bottom (i = left..right-1), right (j =
 * bottom..top-1), top (i = left..right-1), left (j = bottom..top-1), the loop
 * order of ntffTM_TimeCalc (ntffTM.c:326-369).  time_shift[a*n_points + p] is the
 * reference's running timeShift for angle a at point p, built on the host by the
 * same repeated subtraction (ntffTM.c:329-334). */
typedef struct b200fdtd_ntff_plan {
  int32_t top, bottom, left, right;
  int32_t n_points;                /* 2(right-left) + 2(top-bottom), whole surface   */
  int32_t n_local;                 /* points whose column j lies in this engine's     */
                                   /* slab, in the same order (= n_points for 1 slab) */
  int32_t max_time;                /* steps sampled = length of each point's history  */
  int32_t n_bins;                  /* bins kept per angle in U/W (>= max_time to      */
                                   /*   mirror arraySize; max_time suffices for the   */
                                   /*   far field, ntffTM.c:181)                      */
  int32_t n_angles;                /* 360                                             */
  int32_t array_size;              /* NTFFInfo.arraySize: the reference keeps U/W as  */
                                   /* one [360][arraySize] block, and its last-step   */
                                   /* taps at index == arraySize land in the NEXT     */
                                   /* direction's bin 0 (ntffTM.c:285-287 has no      */
                                   /* bound check).  The projection reproduces that   */
                                   /* spill; 0 disables it.                           */
  const double *time_shift;        /* host, [n_angles][n_local]                       */
  /* ntff() of the MPI-variant solvers (mpiTM_UPML.c:849-1037, mpiTE_UPML.c:602-794):  */
  double tap_scale;                /* every tap is (v*w)*tap_scale; they multiply by   */
                                   /* coef = 1/(4 pi C 1e6) per step.  0 or 1: no scale */
  int32_t sample_di, sample_dj;    /* surface point (i, j) reads the fields at global   */
                                   /* cell (i+di, j+dj): those solvers use box indices  */
                                   /* as LOCAL indices, and local i <-> global i-1      */
} b200fdtd_ntff_plan;

/* Far-field post-processing, ntffT?_TimeTranslate + TimeOutput
 * (ntffTM.c:161-232, ntffTE.c:20-55,160-195). */
typedef struct b200fdtd_spectrum_args {
  double coef_re, coef_im;         /* 1/(4 pi C) * csqrt(2 pi C/(i omega))            */
  double z0;                       /* Z_0_S                                           */
  const double *cos_phi, *sin_phi; /* host [n_angles], cos/sin(ang*pi/180)            */
  int32_t n_fft;                   /* NTFF_NUM = 8192                                 */
  int32_t lambda_first_nm, lambda_last_nm;   /* 380, 700                             */
  int32_t reserved;
  double c_hu_nfft;                /* C_0_S * h_u_nm * NTFF_NUM                       */
  const double *twiddle;           /* host, interleaved complex, stage-major: for     */
                                   /* half = n/2, n/4, .. 1: cexp(i*pi/half*k), k<half */
} b200fdtd_spectrum_args;

/* ---- lifetime ----------------------------------------------------------- */
int b200fdtd_device_count(int *count);
const char *b200fdtd_last_error(void);
int b200fdtd_abi_version(void);
/* sizeof() of the ABI structs as this library was compiled, for foreign-language bindings to
 * check their mirror declarations: which = 0 grid, 1 step_args, 2 ntff_plan, 3 spectrum_args,
 * 4 freq_args, 5 batch_source; -1 for an unknown index. */
int b200fdtd_struct_size(int32_t which);

int b200fdtd_create(const b200fdtd_grid *grid, b200fdtd_engine **out);   /* allocateMemories */
int b200fdtd_destroy(b200fdtd_engine *e);                                /* freeMemories     */

/* Pinned host memory for mirrors the getters hand out (cudaHostAlloc). */
int b200fdtd_host_alloc(void **ptr, uint64_t bytes);
int b200fdtd_host_free(void *ptr);
/* Getter mirrors: page-aligned pageable memory that is pinned in place (cudaHostRegister, the pointer
 * does not change) once the caller has refreshed it a few times; see engine.cu. */
int b200fdtd_mirror_alloc(void **ptr, uint64_t bytes);
int b200fdtd_mirror_pin(void *ptr, uint64_t bytes);
int b200fdtd_mirror_free(void *ptr, int32_t pinned);

/* ---- init-time uploads --------------------------------------------------- */
/* setCoefficient (fdtdTM_upml.c:224-274, fdtdTE_upml.c:361-412) */
int b200fdtd_set_upml_tables(b200fdtd_engine *e, const double *tab_i, const double *tab_j);
/* eps_slot: TM 0 = EPS_EZ; TE 0 = EPS_EX, 1 = EPS_EY.  host_eps is the full
 * [n_px][n_py] map; the engine takes its slab. */
int b200fdtd_set_eps(b200fdtd_engine *e, int32_t eps_slot, const double *host_eps);
/* the same from a slab-shaped map [n_px][nj] (what a rank of a multi-GPU run builds) */
int b200fdtd_set_eps_slab(b200fdtd_engine *e, int32_t eps_slot, const double *slab_eps);
/* The same map as a palette: a permittivity map holds few distinct values (vacuum, the materials,
 * the area-averaged boundary cells), so the host can ship 16-bit indices into a table of at most
 * 65536 doubles -- 2 instead of 8 bytes per cell over PCIe -- and the device expands them into the
 * dense map the kernels read (north_star: "per-cell material indices resolved from the model at
 * init").  index_first points at (i = 0, first owned column), rows are ld elements apart (ld = n_py
 * for a whole-grid map, nj for a slab-shaped one).  Bit-identical to b200fdtd_set_eps. */
int b200fdtd_set_eps_palette(b200fdtd_engine *e, int32_t eps_slot, const uint16_t *index_first, int64_t ld,
                             const double *table, int32_t n_values);
int b200fdtd_set_ntff_plan(b200fdtd_engine *e, const b200fdtd_ntff_plan *plan);   /* ntffTM_init */
/* batched engines: the n_batch per-simulation sources (host array) */
int b200fdtd_set_batch_sources(b200fdtd_engine *e, const b200fdtd_batch_source *sources);
/* batched split-field engines: the n_batch per-simulation CW records (host array) */
int b200fdtd_set_batch_cw(b200fdtd_engine *e, const b200fdtd_batch_cw *sources);
/* batched engines: the simulation the state-access and NTFF read-out calls below refer to
 * (get/set_field*, ntff_get_uw, ntff_spectrum, ntff_frequency); default 0 */
int b200fdtd_select_batch(b200fdtd_engine *e, int32_t index);

/* lean form of the split-field kinds 0, 1, 6 (see B200FDTD_LTM_* above); eps through
 * b200fdtd_set_eps (kind 0: slot 0 = EPS_EZ; kind 1: slot 0 = EPS_EX, 1 = EPS_EY) */
int b200fdtd_set_split_tables(b200fdtd_engine *e, const double *tab_i, const double *tab_j);
/* dense per-cell array of a split-field kind: host [n_px][n_py] map (B200FDTD_ST?_C_* or
 * B200FDTD_DENSE_SRC?); replaces the coefficient loops of fdtdTM.c:197-242,
 * nsFdtdTM.c:231-307 etc. */
int b200fdtd_set_dense(b200fdtd_engine *e, int32_t slot, const double *host_map);

/* NS-FDTD TE (kind 7) keeps eight dense coefficient arrays because its PML terms depend on the
 * permittivity through tanh (nsFdtdTE.c:100-181).  Outside the absorbing frame sigma == 0, the four
 * decay coefficients are exactly 1.0 and C_HZXLX == C_HZYLY bit for bit, so there the kernels read
 * three arrays instead of eight (232 instead of 272 B per cell-update) and produce the same bits
 * (1.0 * x == x).  The caller names that rectangle (global i / j, inclusive; lo > hi switches the
 * form off) after checking it on its host arrays; thread blocks that are not wholly inside it take the
 * dense path of the same kernel. */
int b200fdtd_set_split_interior(b200fdtd_engine *e, int32_t i_lo, int32_t i_hi, int32_t j_lo, int32_t j_hi);

/* ---- the hot path -------------------------------------------------------- */
/* One update() (fdtdTM_upml.c:54-66 / fdtdTE_upml.c:168-192): H phase, E phase
 * with source, NTFF surface sample.  Asynchronous on the engine's stream. */
int b200fdtd_step(b200fdtd_engine *e, const b200fdtd_step_args *args);
/* The same, split so a multi-GPU driver can exchange halos between phases.  With peer halos
 * attached AND an NTFF plan set, b200fdtd_phase_sample completes the step: it publishes the
 * E-phase flag that b200fdtd_phase_e holds back until the surface has been sampled (the lower
 * neighbour's next H phase overwrites a ghost column the sample reads).  b200fdtd_step runs
 * the same protocol itself whenever peers are attached, whatever form the step takes. */
int b200fdtd_phase_h(b200fdtd_engine *e, const b200fdtd_step_args *args);
int b200fdtd_phase_e(b200fdtd_engine *e, const b200fdtd_step_args *args);
int b200fdtd_phase_sample(b200fdtd_engine *e, const b200fdtd_step_args *args);
/* H and E phase of the serial UPML kinds in ONE pass (what b200fdtd_step launches when
 * b200fdtd_get_step_form says 3 or 4): edge pre-pass + the TMA-staged marching kernel; an error if
 * the engine / source is not served by it (see B200FDTD_OPT_FUSED). */
int b200fdtd_phase_fused(b200fdtd_engine *e, const b200fdtd_step_args *args);
int b200fdtd_sync(b200fdtd_engine *e);
/* n_steps consecutive update() calls starting at time0, replayed from a CUDA graph: the step's
 * only time dependence -- the pulse's (time - t0) and the NTFF sample index -- is read from a
 * device-side clock that a one-thread kernel advances after every step, so the launch sequence
 * [H phase, E phase + source, surface sample, clock] is identical for every step and a chunk of
 * it is captured once and replayed.  Removes the per-launch host cost that dominates small
 * grids (a 256 x 256 step is ~3 us of device work).  Serial UPML kinds (2, 3) driven by the
 * pulse sources of b200fdtd_set_batch_sources (also accepted for n_batch = 1); no point / CW /
 * line source, no peer halos.  Bit-identical to n_steps b200fdtd_step calls. */
int b200fdtd_run_steps(b200fdtd_engine *e, double time0, int32_t n_steps);
/* The same for the split-field kinds (0, 1, 6, 7), whose CW source changes with the step (phases
 * w(t +- 1/2) or w(t+1), w t and the ramp ray_coef): args[s] is what b200fdtd_step would have been
 * given at step s.  The CW records of a chunk go to a device table in one copy and every kernel of
 * the captured chunk reads its own step's record, so the graph is the same for every chunk.  ns_r2
 * is taken from args[0].  Bit-identical to n_steps b200fdtd_step calls. */
int b200fdtd_run_split_steps(b200fdtd_engine *e, const b200fdtd_step_args *args, int32_t n_steps);

/* Halo columns for the y-slab split (replaces Connection_ISend_IRecvH/E,
 * mpiTM_UPML.c:252-296).  which: 0 = after the H phase (TM Hx / TE Hz, my last
 * owned column -> upper neighbour's low ghost), 1 = after the E phase (TM Ez / TE
 * Ex, my first owned column -> lower neighbour's high ghost).  dev_buf is a
 * device pointer to n_px complex values. */
int b200fdtd_halo_pack(b200fdtd_engine *e, int32_t which, void *dev_buf);
int b200fdtd_halo_unpack(b200fdtd_engine *e, int32_t which, const void *dev_buf);
/* Direct peer (NVLink) halos between the y-slab engines of different processes on one node.
 * Each engine exports a 256-byte blob (CUDA IPC handles of its E array, H array and flag
 * words + its nj); the driver moves blobs between ranks (any transport) and attaches the
 * lower (which_neighbour = 0) and/or upper (1) neighbour's blob.  From then on
 * b200fdtd_phase_h / _phase_e store their halo column straight into the neighbour's ghost
 * column and order themselves across GPUs with device-side flag words -- no halo_pack /
 * halo_unpack, no NCCL call and no host synchronisation per step. */
#define B200FDTD_PEER_BLOB_BYTES 256
int b200fdtd_peer_export(b200fdtd_engine *e, void *blob);
int b200fdtd_peer_attach(b200fdtd_engine *e, int32_t which_neighbour, const void *blob);
/* The same between two engines of ONE process (several devices driven by one host thread -- the
 * C plugin's multi-GPU mode -- or several slabs on one device): plain pointers, CUDA peer access
 * enabled on demand; no IPC, no NCCL.  Attach both directions: lower.attach(1, upper) and
 * upper.attach(0, lower). */
int b200fdtd_peer_attach_engine(b200fdtd_engine *e, int32_t which_neighbour, b200fdtd_engine *neighbour);
/* Launch everything on this CUDA stream (a cudaStream_t) from now on. */
int b200fdtd_set_stream(b200fdtd_engine *e, void *cuda_stream);

/* ---- state access (synchronous) ------------------------------------------ */
/* getters: device slab -> host [n_px][n_py] complex array, columns [j0, j0+nj) */
int b200fdtd_get_field(b200fdtd_engine *e, int32_t slot, double *host_complex);
int b200fdtd_set_field(b200fdtd_engine *e, int32_t slot, const double *host_complex);
/* general form: element (i = 0, j = j0) goes to host_first_cell, rows are ld_complex complex
 * values apart (e.g. the (N+2)-wide ghost-ring mirrors of the MPI-variant solvers) */
int b200fdtd_get_field_ld(b200fdtd_engine *e, int32_t slot, double *host_first_cell, int64_t ld_complex);
/* slab-shaped variant: host array is [n_px][nj] complex (what one rank mirrors) */
int b200fdtd_get_field_slab(b200fdtd_engine *e, int32_t slot, double *slab_complex);
/* Order-independent 64-bit digest of the owned cells of a field plane (every 64-bit word mixed
 * with its GLOBAL cell position, summed mod 2^64): equal digests <=> equal bits in equal cells,
 * and the digests of the y-slabs of a split run add up (mod 2^64) to the single-slab digest.
 * Lets a multi-GPU run or a stress loop prove bit-identity without moving the planes. */
int b200fdtd_field_digest(b200fdtd_engine *e, int32_t slot, uint64_t *digest);
int b200fdtd_zero_state(b200fdtd_engine *e);        /* the memsets of reset(), fdtdTM_upml.c:98-113 */

/* ---- NTFF ---------------------------------------------------------------- */
/* Turn the recorded surface history into U/W[3][n_angles][n_bins]
 * (ntffTM_TimeCalc's accumulation, ntffTM.c:279-371, deferred).  Slots: TM
 * Ux,Uy,Wz; TE Wx,Wy,Uz. */
int b200fdtd_ntff_project(b200fdtd_engine *e);
int b200fdtd_ntff_get_uw(b200fdtd_engine *e, int32_t slot, double *host_complex);
/* dst.U/W += src.U/W after both have been projected: the end-of-run sum over the y-slabs of one
 * process (peer copy + one kernel; the reference's MPI solvers never reduce) */
int b200fdtd_ntff_add_uw(b200fdtd_engine *dst, b200fdtd_engine *src);
/* The surface sample of step t as the caller gathered it (n_local complex E values and n_local
 * complex H values in the perimeter order of b200fdtd_set_ntff_plan, signs and the two-cell H average
 * already applied: what ntff_sample_kernel records).  Replaces the per-step field reads of
 * ntffTM_TimeCalc / ntffTE_TimeCalc for hosts that keep their fields on the CPU. */
int b200fdtd_ntff_push_samples(b200fdtd_engine *e, int32_t t, const double *e_complex, const double *h_complex);
/* overwrite one U/W slot ([n_angles][n_bins] complex) from the host */
int b200fdtd_ntff_set_uw(b200fdtd_engine *e, int32_t slot, const double *host_complex);
/* device pointer + element count of the whole U/W block, for an NCCL reduce */
int b200fdtd_ntff_uw_device(b200fdtd_engine *e, void **dev_ptr, uint64_t *n_doubles);
/* out[(lambda-lambda_first)*n_angles + ang], the table ntff_outputEnormBin writes */
int b200fdtd_ntff_spectrum(b200fdtd_engine *e, const b200fdtd_spectrum_args *args, double *out);

/* One-shot frequency-domain surface integral, ntffTM_Frequency (ntffTM.c:72-158): for each
 * direction a, Nz, Lx, Ly = sum over the surface of (tangential H or E) * cexp(i k r^.r2).
 * Works on the current device fields of any TM-type kind (UPML, Berenger, NS).  The caller
 * finishes with coef*(Z0*Nz + Lphi)*sqrt(h_u) on the host (ntffTM.c:82,155-156).
 * out: [3][n_angles] complex (Nz, Lx, Ly). */
typedef struct b200fdtd_freq_args {
  int32_t top, bottom, left, right, cx, cy;
  int32_t n_angles, reserved;
  double k;                        /* field_getK()                                     */
  const double *cos_a, *sin_a;     /* host [n_angles]: cos/sin(ang*pi/180), host libm  */
} b200fdtd_freq_args;
int b200fdtd_ntff_frequency(b200fdtd_engine *e, const b200fdtd_freq_args *args, double *out);

/* ---- tuning switches ------------------------------------------------------ */
enum {
  B200FDTD_OPT_FUSED = 1,     /* the one-pass H+E step (fused_kernels.cu) of the serial UPML kinds (2 TM,
                                 3 TE; one unbatched double-precision slab, alone or with peer halos):
                                 TM 232 instead of 264 B per cell-update, TE 272 instead of 288,
                                 bit-identical to the two-kernel step; with OPT_LEAN_INTERIOR TM 136
                                 instead of 168, TE 176 instead of 192, bit-identical to the two-kernel
                                 lean step.  1 wherever it can run, 0 never, 2 (default) on grids of
                                 >= 2^22 updated cells                                           */
  B200FDTD_OPT_STORE_H = 2,   /* 1: the one-pass kernel also writes H every step (+32 / +16 B/cell);
                                 0 (default): H is derived from B on demand (getters, NTFF,
                                 halo) -- Hx == Bx/mu0 holds after every H phase                 */
  B200FDTD_OPT_BAND_ROWS = 3, /* rows a CTA marches per band in the one-pass kernel (default 32)  */
  B200FDTD_OPT_FUSED_SHAPE = 4, /* launch shape of the one-pass kernel, consumer warps x row buffers:
                                 20 (default) 8 x 4, 21 8 x 6, 22 16 x 3 (TM only), 23 4 x 8, 24 8 x 3 */
  /* 5, 6: retired (the pipelined persistent step of round 1) */
  B200FDTD_OPT_F32_PAIRS = 7,  /* single-precision engines: 1 (default) two cells per thread with
                                  128-bit accesses, 0 the one-cell-per-thread kernels; identical bits */
  B200FDTD_OPT_UNIT_SPLIT = 9,  /* UPML kinds, double precision, two-kernel step.  Inside the frame-free
                                  rectangle (b200fdtd_upml_interior) every coefficient is exactly 1.0
                                  and 1.0 * x == x, so dedicated kernels evaluate the reference's
                                  expressions there without table reads and multiplications -- same
                                  bits, fewer registers, more loads in flight; the frame goes through
                                  the full kernels.  0: one kernel per phase over the whole grid,
                                  1: split whenever the rectangle exists, 2 (default): split when the
                                  rectangle holds >= 2^20 cells and >= 3/4 of the updated cells.     */
  B200FDTD_OPT_DERIVED_E = 10,  /* one-pass step.  1 (default): in "vacuum row-strips" -- a row of a CTA tile
                                  whose cells all have eps == 1 (and none is an NTFF sample cell) -- the E
                                  arrays are derived state: E = D/1.0 holds the bits of D, so the pass takes
                                  the old E from D, stores no E and stages no eps there: TM 232 -> 192 B per
                                  cell-update, TE 272 -> 192, lean 136 / 176 -> 96.  Bit-identical; the E
                                  arrays are brought up to date whenever something outside the pass reads
                                  them.  0: E read and written everywhere.                            */
  B200FDTD_OPT_LEAN_INTERIOR = 8 /* UPML kinds, two-kernel step.  1: cells outside the absorbing frame
                                  -- where every UPML coefficient of fdtdTM_upml.c:253-271 /
                                  fdtdTE_upml.c:384-403 is exactly 1 -- advance B and D directly
                                  (B' = B - curl E, D' = D + curl H) and skip the M / J recurrences,
                                  which cancel there: 168 instead of 264 B per cell-update (TM), 192
                                  instead of 288 (TE).  Mathematically the same update with fewer
                                  roundings, so NOT bit-identical to the reference: fields agree to
                                  ~1e-13 relative (tests: <= 1e-12, far field <= 1e-10).  The M / J
                                  arrays then only mean something inside the frame.  0 (default): the
                                  reference's arithmetic in every cell.                            */
};
int b200fdtd_set_option(b200fdtd_engine *e, int32_t option, int32_t value);
/* cells per step that lie in vacuum row-strips of the one-pass step (B200FDTD_OPT_DERIVED_E) with the
 * current launch shape, permittivity maps and NTFF plan; 0 when the step does not use them */
int b200fdtd_onepass_vacuum_cells(b200fdtd_engine *e, uint64_t *cells);

/* The region OPT_LEAN_INTERIOR may treat as frame-free, from the 1-D coefficient tables alone
 * (host arithmetic, no device needed): the longest run of i (rows) and of j (columns) around the
 * grid centre whose table entries all hold their sigma == 0 values.  kind: a UPML kind (2-5);
 * tab_i / tab_j as for b200fdtd_set_upml_tables ([B200FDTD_UPML_TABS][n_px] / [..][n_py]).
 * out = { i_lo, i_hi, j_lo, j_hi } inclusive, an empty range as lo > hi. */
int b200fdtd_upml_interior(int32_t kind, const double *tab_i, int32_t n_px, const double *tab_j, int32_t n_py,
                           int32_t out[4]);
/* the rectangle this engine's split forms use (global i / j of this slab's share; without row
 * r_lo / column c_lo, which stay with the frame kernels); {1,0,1,0} when no split form is active */
int b200fdtd_get_lean_extent(b200fdtd_engine *e, int32_t out[4]);
/* Host-only (no device needed): the launch geometry of the split forms for an updated rectangle
 * and an interior rectangle inside it, both { r_lo, r_hi, c_lo, c_hi } inclusive.  rects receives up
 * to 5 x 7 ints { r_lo, r_hi, c_lo, c_hi, bw_log2, nbx, blk_end }: first the interior launch, then the
 * pieces of the frame launch (block numbering restarts; a block covers 2^bw_log2 columns x
 * 256 >> bw_log2 rows).  For tests: together they must tile the updated rectangle exactly. */
int b200fdtd_split_geometry(const int32_t updated[4], const int32_t interior[4], int32_t *rects, int32_t *n_rects);
/* which form b200fdtd_step launches right now: 0 one full kernel per phase, 1 unit-coefficient
 * interior kernel + frame (bit-identical to 0), 2 lean interior + frame, 3 the one-pass step
 * (bit-identical to 0; phase_h / phase_e then still launch form 0 or 1), 4 the one-pass step in
 * the lean form (bit-identical to 2) */
int b200fdtd_get_step_form(b200fdtd_engine *e, int32_t *form);

/* Device self-test: the kernels replace `x / d` (d loop-invariant, e.g. MU_0_S) by a
 * reciprocal-multiply with an FMA correction that is claimed to be the identical,
 * correctly rounded quotient.  Counts disagreements with IEEE division over `samples`
 * pseudo-random operands (raw bit patterns and field-like magnitudes); must be 0. */
int b200fdtd_selftest_division(double divisor, uint64_t samples, uint64_t *mismatches);
/* divisor == 0: the per-cell division by the permittivity instead -- every sample its own divisor
 * (random significands, short decimals, all-ones significands, 2^-4 .. 2^8), through the reciprocal
 * + two correction steps the material cells use (div_eps, upml_common.cuh); must be 0 as well. */

/* ---- introspection -------------------------------------------------------- */
/* kernels launched by this engine since creation (bench.py's gpu_launches) */
int b200fdtd_launch_count(b200fdtd_engine *e, uint64_t *count);
int b200fdtd_device_bytes(b200fdtd_engine *e, uint64_t *bytes);
/* free / total memory of a CUDA device (-1 = current), for sizing angle batches */
int b200fdtd_mem_info(int32_t device, uint64_t *free_bytes, uint64_t *total_bytes);
/* CUDA-event timing on the engine's stream: elapsed ms between the two marks */
int b200fdtd_timer_start(b200fdtd_engine *e);
int b200fdtd_timer_stop(b200fdtd_engine *e, float *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* B200FDTD_H */
