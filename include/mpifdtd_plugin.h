/* mpifdtd_plugin.h -- the host-side plugin surface libmpifdtd_b200.so exports.
 *
 * Every symbol below keeps the name, argument meaning and error behaviour
 * (printf + exit(2)) of the reference interface it replaces, so the reference's
 * own driver (main.c) and viewer (drawer.c) link against this library instead of
 * the reference's simulator/field/models/solver objects.  Citations are
 * file:line into rennone/mpiFDTD.  Host code is C99; the time-stepping itself
 * runs on the GPU behind include/b200fdtd.h.
 *
 * A translation unit may include either this header or the reference's own
 * headers (simulator.h, field.h, models.h, ...): the declarations are ABI
 * identical.
 */
#ifndef MPIFDTD_PLUGIN_H
#define MPIFDTD_PLUGIN_H

#include <complex.h>
#include <stdio.h>

#ifdef __cplusplus
#error "C99 host interface (uses double complex); bind the engine through b200fdtd.h from C++"
#endif

#ifndef bool                 /* bool.h:1-18 of the reference: bool is int */
#define bool int
#define true 1
#define false 0
#endif

typedef double complex dcomplex;          /* myComplex.h:6 */

/* ---- units and constants (field.h:70-75) ------------------------------- */
#define C_0_S 0.7071                      /* deliberately not 1/sqrt(2) */
static const double LIGHT_SPEED_S = 0.7071;
static const double EPSILON_0_S = 1.0;
static const double MU_0_S = 1.0 / C_0_S / C_0_S;
static const double Z_0_S = 1.41422712488;

/* ---- grid / wave / far-field descriptors (field.h:20-68) ---------------- */
typedef struct FieldInfo {
  int width_nm, height_nm;   /* physical size of the region            */
  int h_u_nm;                /* cell edge                               */
  int pml;                   /* PML thickness in cells                  */
  int lambda_nm;             /* wavelength                              */
  int angle_deg;             /* incidence angle                         */
  int stepNum;               /* number of time steps                    */
} FieldInfo;

typedef struct FieldInfo_S {
  int N_X, N_Y;              /* cells without PML */
  int N_PX, N_PY;            /* cells with PML    */
  int N_CELL;
  int N_PML;
  int DX, DY;                /* index strides: DX = N_PY, DY = 1 */
} FieldInfo_S;

typedef struct SubFieldInfo_S {
  int OFFSET_X, OFFSET_Y;
  int SUB_N_X, SUB_N_Y;
  int SUB_N_PX, SUB_N_PY;
  int SUB_N_CELL;
  int Rank;
  int RtRank, LtRank, TpRank, BmRank;
} SubFieldInfo_S;

typedef struct WaveInfo_S {
  double Lambda_s, T_s, Omega_s, K_s, K_0_s, Angle_deg;
} WaveInfo_S;

typedef struct NFFInfo {
  int top, bottom, left, right;
  int cx, cy;
  double RFperC;
  int arraySize;
} NTFFInfo;

/* globals kept for source compatibility (field.h:77-82) */
extern int N_X, N_Y, N_CELL, N_PML, N_PX, N_PY;

/* ---- grid / time services (field.h:102-163) ----------------------------- */
extern void field_init(FieldInfo field_info);                 /* field.c:89  */
extern void field_reset(void);                                /* field.c:83  */
extern void field_nextStep(void);                             /* field.c:312 */
extern bool field_isFinish(void);                             /* field.c:317 */
extern void field_setWaveAngle(int deg);                      /* field.c:39  */
extern double field_sigmaX(double x, double y);               /* field.c:259 */
extern double field_sigmaY(double x, double y);               /* field.c:272 */
extern double field_pmlCoef(double ep_mu, double sig);        /* field.c:288 */
extern double field_pmlCoef_LXY(double ep_mu, double sig);    /* field.c:292 */
extern double field_ns_beta(double alpha, double alpha_aster);        /* field.c:298 */
extern double field_ns_beta_aster(double alpha, double alpha_aster);  /* field.c:304 */
extern double field_toCellUnit(const double);                 /* field.c:76  */
extern double field_toPhisycalUnit(const double);             /* field.c:79  */
extern double field_getT(void);
extern double field_getK(void);
extern double field_getRayCoef(void);
extern double field_getOmega(void);
extern double field_getLambda(void);
extern double field_getWaveAngle(void);
extern double field_getTime(void);
extern double field_getMaxTime(void);
extern NTFFInfo field_getNTFFInfo(void);
extern WaveInfo_S field_getWaveInfo_S(void);
extern SubFieldInfo_S field_getSubFieldInfo_S(void);
extern FieldInfo_S field_getFieldInfo_S(void);
extern FieldInfo field_getFieldInfo(void);
extern int field_getOffsetX(void), field_getOffsetY(void);
extern int field_getSubNx(void), field_getSubNy(void);
extern int field_getSubNpx(void), field_getSubNpy(void), field_getSubNcell(void);
extern int ind(const int, const int);                         /* field.c:74 */
extern int field_index(int i, int j);                         /* field.c:70 */
extern int field_subIndex(int i, int j);                      /* field.c:66 */
extern dcomplex field_pointLight(void);                       /* field.c:145 */
extern void field_outputElliptic(const char *fileName, double complex *data);      /* field.c:322 */
extern void field_outputAllDataComplex(const char *fileName, double complex *data);/* field.c:346 */
extern void field_outputAllDataDouble(const char *fileName, double *data);         /* field.c:368 */
/* field_scattered*(), the host-array source injectors of field.h:143-149, are NOT
 * exported: source injection is fused into the GPU E-phase kernel. */

/* ---- material models (models.h:5-31) ------------------------------------ */
enum MODEL { NO_MODEL, MIE_CYLINDER, LAYER, MORPHO_SCALE, CONCENTRIC_CIRCLE, ZIGZAG, TRACE_IMAGE };
enum MODE { D_X, D_Y, D_XY };
extern void models_setModel(enum MODEL model);                /* models.c:105 */
extern double models_eps(double x, double y, enum MODE mode); /* models.c:132 */
extern bool models_isFinish(void);                            /* models.c:92  */
extern void models_moveDirectory(void);                       /* models.c:98  */
extern void models_needSize(int *x_nm, int *y_nm);            /* models.c:152 */
extern void models_initModel(void);                           /* models.c:157 */

/* ---- the solver table (simulator.h:8-31, simulator.c:19-27) ------------- */
enum SOLVER { TM_2D, TE_2D, TM_UPML_2D, TE_UPML_2D, MPI_TM_UPML_2D, MPI_TE_UPML_2D,
              NS_TM_2D, NS_TE_2D };
extern void simulator_setSolver(enum SOLVER solverType);      /* simulator.c:221 */
extern void simulator_init(FieldInfo field_info);             /* simulator.c:226 */
extern void simulator_calc(void);                             /* simulator.c:215 */
extern bool simulator_isFinish(void);                         /* simulator.c:270 */
extern void simulator_finish(void);                           /* simulator.c:257 */
extern void simulator_reset(void);                            /* simulator.c:243 */
extern void simulator_solverInit(void);                       /* simulator.c:236 */
extern void simulator_changeModelAndRestart(void);            /* simulator.c:252 */
extern double complex *simulator_getDrawingData(void);        /* simulator.c:266 */
extern double *simulator_getEps(void);                        /* simulator.c:276 */
extern void simulator_moveDirectory(void);                    /* simulator.c:209 */

/* per-solver entry points the table is filled from (fdtdTM_upml.h:5-14 etc.).
 * Getters return borrowed host pointers, index k = i*N_PY + j, valid from init
 * until finish; each call refreshes the host mirror from device memory. */
#define MPIFDTD_DECLARE_SOLVER(P, A, B, Cc)                   \
  extern void (*P##_getUpdate(void))(void);                  \
  extern void (*P##_getFinish(void))(void);                  \
  extern void (*P##_getReset(void))(void);                   \
  extern void (*P##_getInit(void))(void);                    \
  extern double complex *P##_get##A(void);                   \
  extern double complex *P##_get##B(void);                   \
  extern double complex *P##_get##Cc(void);                  \
  extern double *P##_getEps(void);
MPIFDTD_DECLARE_SOLVER(fdtdTM_upml, Hx, Hy, Ez)               /* fdtdTM_upml.h:5-14 */
MPIFDTD_DECLARE_SOLVER(fdtdTE_upml, Ex, Ey, Hz)               /* fdtdTE_upml.h:5-14 */
MPIFDTD_DECLARE_SOLVER(mpi_fdtdTM_upml, Hx, Hy, Ez)           /* mpiTM_UPML.h:5-12  */
MPIFDTD_DECLARE_SOLVER(mpi_fdtdTE_upml, Ex, Ey, Hz)           /* mpiTE_UPML.h       */
MPIFDTD_DECLARE_SOLVER(fdtdTM, Hx, Hy, Ez)                    /* fdtdTM.h:5-17      */
MPIFDTD_DECLARE_SOLVER(fdtdTE, Ex, Ey, Hz)                    /* fdtdTE.h:5-17      */
MPIFDTD_DECLARE_SOLVER(nsFdtdTM, Hx, Hy, Ez)                  /* nsFdtdTM.h:5-19    */
MPIFDTD_DECLARE_SOLVER(nsFdtdTE, Ex, Ey, Hz)                  /* nsFdtdTE.h:5-19    */

/* sub-domain sizes of the MPI-variant solvers (mpiTM_UPML.h:14-18); one rank owns the
 * whole grid here, so they report N and N+2 (ghost ring) */
extern int mpi_fdtdTM_upml_getSubNx(void), mpi_fdtdTM_upml_getSubNy(void), mpi_fdtdTM_upml_getSubNpx(void);
extern int mpi_fdtdTM_upml_getSubNpy(void), mpi_fdtdTM_upml_getSubNcell(void);
extern int mpi_fdtdTE_upml_getSubNx(void), mpi_fdtdTE_upml_getSubNy(void), mpi_fdtdTE_upml_getSubNpx(void);
extern int mpi_fdtdTE_upml_getSubNpy(void), mpi_fdtdTE_upml_getSubNcell(void);

/* extra getters of the split-field solvers (fdtdTM.h:10-11, fdtdTE.h:12-13,
 * nsFdtdTM.h:11-19, nsFdtdTE.h:11-19) */
extern double complex *fdtdTM_getEzx(void), *fdtdTM_getEzy(void);
extern double complex *fdtdTE_getHzx(void), *fdtdTE_getHzy(void);
extern double complex *nsFdtdTM_getEzx(void), *nsFdtdTM_getEzy(void);
extern double complex *nsFdtdTE_getHzx(void), *nsFdtdTE_getHzy(void);
extern double *nsFdtdTM_getEpsX(void), *nsFdtdTM_getEpsY(void), *nsFdtdTM_getEpsZ(void);
extern double *nsFdtdTE_getEpsX(void), *nsFdtdTE_getEpsY(void), *nsFdtdTE_getEpsZ(void);
struct Solver;                                                /* solver.h:7-20 */
extern struct Solver *nsFdtdTM_getSolver(void);
extern struct Solver *nsFdtdTE_getSolver(void);

/* ---- far-field output format (ntff.h:4-9) ------------------------------- */
#define LAMBDA_ST_NM 380
#define LAMBDA_EN_NM 700
#define NTFF_NUM 8192
extern void ntff_outputEnormTxt(double **e_norm, const char *file_name);  /* ntff.c:6  */
extern void ntff_outputEnormBin(double **e_norm, const char *file_name);  /* ntff.c:21 */

/* One-shot frequency-domain far field of the active TM-type solver (ids 0, 2, 4, 6):
 * the surface integral of ntffTM_Frequency (ntffTM.c:72-158) evaluated on the GPU over the
 * solver's current fields.  result[360] as the reference's resultEz. */
extern int mpifdtd_ntffFrequency(int solver_id, double complex result[360]);

/* ---- opt-in source forms (extensions; nothing of the reference calls them) ------------
 * The reference carries source injectors it never calls.  These switches wire them in; they
 * are read by the next init().  Default: the sources update() hard-wires.
 *   mpifdtd_enablePointSource(1): field_pointLight() (field.c:145-152) added to Ez (TM) / Ex
 *     (TE) at the grid centre -- gives the NoModel configuration something to propagate.
 *   mpifdtd_setSourceForm(MPIFDTD_SRC_CW): serial TM UPML uses field_scatteredWave
 *     (field.c:202-218) instead of the pulse, the commented line at fdtdTM_upml.c:62.
 *   mpifdtd_setSourceForm(MPIFDTD_SRC_PLANE): planeWave (mpiTM_UPML.c:377-403, commented call
 *     at mpiTM_UPML.c:204) added on Ez for solver ids 2 and 4. */
enum { MPIFDTD_SRC_DEFAULT = 0, MPIFDTD_SRC_CW = 1, MPIFDTD_SRC_PLANE = 2 };
extern void mpifdtd_enablePointSource(int on);
extern void mpifdtd_setSourceForm(int form);
/* Optional single-precision path of the UPML solvers (ids 2-5), read by the next init():
 * 0 = double (default, the reference's arithmetic), 1 = float fields/eps/coefficients on the
 * GPU.  Getters and far-field files keep their double formats. */
extern void mpifdtd_setPrecision(int precision);

/* ---- angle sweeps as one batched GPU job (extension; replaces the one-angle-per-rank loop
 * of main.c:114-138,183-211) ------------------------------------------------------------
 * mpifdtd_setAngleBatch(angles, n): the next init() of a serial UPML solver (ids 2, 3) runs
 * all n incidence angles at once in one batched engine (they share grid, permittivity and
 * coefficients); every simulator_calc() advances all of them, reset()/finish() write each
 * angle's "<ang>[deg].txt" / "<ang>[deg]_380nm_700nm_b.dat" into cwd.  n = 0: off.
 * mpifdtd_selectAngle(k): the simulation the getters show (default 0).
 * mpifdtd_runAngleSweep: the whole batch loop for the current model and solver -- angles
 * start, start+delta, .. <= end in chunks of at most max_batch simulations (0 = as many as
 * fit), each chunk init -> stepNum x calc -> finish.  Returns the number of simulations run. */
extern void mpifdtd_setAngleBatch(const int *angles_deg, int n);
extern void mpifdtd_selectAngle(int index);
extern int mpifdtd_runAngleSweep(FieldInfo field_info, int start_deg, int end_deg, int delta_deg, int max_batch);

/* ---- the reference's public NTFF entry points (ntffTM.h:7-28, ntffTE.h:5-15), for a solver file of
 * the maintainer's own that keeps its fields on the host.  Same signatures.  TimeCalc gathers the
 * surface samples of the step from the host arrays (the reference's reads, signs and H averages) and
 * appends them to a GPU-side history; the 360-direction binning of ntffTM.c:279-371 runs on the GPU
 * when the accumulators are needed: TimeTranslate / TimeOutput (and mpifdtd_ntffSync) first store
 * them into the arrays they are given -- all arraySize bins, within 1e-13 of the reference's -- so
 * U / W are current from then on, not after every TimeCalc.  TimeOutput writes "<ang>[deg].txt" and
 * "<ang>[deg]_380nm_700nm_b.dat" into cwd like upstream.  ntff?_init (re)creates the accumulators for
 * the current field_init() state; ntff?_finish releases them. */
extern void ntffTM_init(void);
extern void ntffTM_finish(void);
extern void ntffTM_TimeCalc(dcomplex *Hx, dcomplex *Hy, dcomplex *Ez, dcomplex *Ux, dcomplex *Uy, dcomplex *Wz);
extern void ntffTM_TimeTranslate(dcomplex *Ux, dcomplex *Uy, dcomplex *Wz, dcomplex *Eth, dcomplex *Eph);
extern void ntffTM_TimeOutput(dcomplex *Ux, dcomplex *Uy, dcomplex *Wz);
extern void ntffTM_Frequency(dcomplex *Hx, dcomplex *Hy, dcomplex *Ez, dcomplex resEz[360]);
extern void ntffTE_init(void);
extern void ntffTE_finish(void);
extern void ntffTE_TimeCalc(dcomplex *Ex, dcomplex *Ey, dcomplex *Hz, dcomplex *Wx, dcomplex *Wy, dcomplex *Uz);
extern void ntffTE_TimeTranslate(dcomplex *Wx, dcomplex *Wy, dcomplex *Uz, dcomplex *Eth, dcomplex *Eph);
extern void ntffTE_TimeOutput(dcomplex *Wx, dcomplex *Wy, dcomplex *Uz);
/* store the accumulators now (tm != 0: Ux, Uy, Wz; else Wx, Wy, Uz; NULL skips one) */
extern void mpifdtd_ntffSync(int tm, dcomplex *a0, dcomplex *a1, dcomplex *a2);

/* ---- multi-GPU mode of the serial UPML solvers, host code staying C (extension; replaces
 * init_mpi + the per-step halo Sendrecv of mpiTM_UPML.c:196-217,252-334,718-748) ---------------
 * mpifdtd_setDevices(n): the next init() of solver ids 2 / 3 cuts the grid into n y-slabs, one
 * engine per slab, slab g on CUDA device g modulo the visible devices; one process, one host
 * thread; halos are direct NVLink peer stores ordered by device flags; reset()/finish() sum the
 * slabs' NTFF partial sums and write the same files; getters gather the slabs.  n <= 1: one
 * engine (default).  Environment MPIFDTD_DEVICES=n does the same for an unmodified main.c.
 * mpifdtd_upml_slab_count / _slab_engine: how many slabs the active solver (kind 2 / 3) runs and
 * the engine handle of slab g (harness use: device timers, digests). */
extern void mpifdtd_setDevices(int n);
extern int mpifdtd_upml_slab_count(int kind);
extern struct b200fdtd_engine *mpifdtd_upml_slab_engine(int kind, int g);

/* ---- config.txt (parser.h:5, configSample.txt:6-22, main.c:319-366) ------ */
extern bool parser_nextLine(FILE *fp, char buf[]);            /* parser.c:3 */
typedef struct MpifdtdConfig {
  FieldInfo field_info;
  int startAngle, endAngle, deltaAngle;
  int ModelType, SolverType;
} MpifdtdConfig;
/* Reads the 11 values of configSample.txt in order (width, height, h_u, pml,
 * lambda, step, start/end/delta angle, model id, solver id).  Returns 0 on
 * success; on a missing file or a short file prints and exit(2)s like the
 * reference's readConfig (main.c:319-366, commented out upstream). */
extern int mpifdtd_readConfig(const char *path, MpifdtdConfig *out);
/* initConfigFromText (main.c:368-394): rank 0 reads config.txt, prints the FieldSetting banner and
 * sends the struct -- sizeof(MpifdtdConfig)/sizeof(int) ints, as upstream's MPI_Send of MPI_INTs --
 * to every other rank, which receive it from rank 0.  The transport is the caller's (thin wrappers
 * of MPI_Send / MPI_Recv, or any other message layer); with send == recv == NULL a built-in one
 * for ranks on one node: a file under /dev/shm keyed by MPIFDTD_JOB_ID (default: the launcher's
 * pid), published atomically by rank 0 and polled by the others (MPIFDTD_BCAST_TIMEOUT_S, default
 * 120).  Callbacks return 0 on success.  Errors: message + exit(2).  mpifdtd_configBroadcastDone()
 * removes the published file (rank 0, once everybody has read). */
typedef int (*mpifdtd_send_ints)(const int *buf, int count, int dest, void *ctx);
typedef int (*mpifdtd_recv_ints)(int *buf, int count, int src, void *ctx);
extern int mpifdtd_initConfigFromText(const char *path, int rank, int n_ranks, MpifdtdConfig *cfg,
                                      mpifdtd_send_ints send, mpifdtd_recv_ints recv, void *ctx);
extern void mpifdtd_configBroadcastDone(void);

/* ---- small utilities main.c / drawer.c pull in (function.h, myComplex.h) -- */
extern double *newDouble(int size);
extern void freeDouble(double *array);
extern dcomplex *newDComplex(int size);
extern void freeDComplex(dcomplex *array);
extern double cnorm(dcomplex c);                              /* squared magnitude */
extern double complex cbilinear(dcomplex *p, double x, double y, int width, int height);
extern double dbilinear(double *p, double x, double y, int width, int height);
extern FILE *openFile(const char *file_name);
extern FILE *FileOpen(const char *file_name, const char *mode);
extern bool makeDirectory(const char *);
extern void moveDirectory(const char *);
extern void makeAndMoveDirectory(const char *);

#endif /* MPIFDTD_PLUGIN_H */
