// engine.h -- internal state of one b200fdtd engine (one y-slab on one GPU).
//
// Device layout.  Every complex field is a pitched 2-D array of double2 with the
// reference's orientation (x = i slow, y = j fast; field.c:70-72) so host
// mirrors need no transpose:
//
//   element(i, j) = base[(i + 1) * pitch + (j - j0) + JOFF]
//
// Row -1 and row n_px are ghost rows (always zero for the serial solvers; the
// "MPI" solver kinds update all N x N cells against them, mpiTM_UPML.c:737-743).
// Column offset JOFF = 8 puts the first owned column on a 128-byte boundary and
// leaves the low ghost column at JOFF-1; the high ghost column is JOFF + nj.
// pitch is a multiple of 8 elements (128 B).  eps arrays share the layout with
// double elements.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "b200fdtd.h"

#define B200_JOFF 8

struct NtffPoint {          // one perimeter sample owned by this slab
  long long k;              // element offset of (i, j) in the pitched layout
  int edge;                 // 0 bottom, 1 right, 2 top, 3 left
  int p_global;             // index into the reference's perimeter order
};

struct NtffState {
  bool ready;
  int top, bottom, left, right;
  int n_points_global;
  int n_local;              // perimeter points whose column lies in this slab
  int max_time, n_bins, n_angles, array_size;
  double tap_scale;         // per-tap factor of the MPI-variant ntff() (1 = none)
  NtffPoint *pts;           // device [n_local]
  double *ts;               // device [n_angles][n_local]
  double2 *hist_e, *hist_h; // device [n_batch][n_local][max_time]
  double2 *uw;              // device [n_batch][3][n_angles][n_bins]
  int steps_recorded;
  // device scratch of the spectrum step, kept between calls (a sweep calls it once per angle)
  double *sp_cos, *sp_sin, *sp_out;
  double2 *sp_tw;
  int sp_n_fft, sp_n_lam;
};

struct FusedState {          // side buffers of the one-pass step (fused_kernels.cu)
  bool ready;
  int n_strips, n_bands, band_h;
  double2 *col_e, *col_h, *row_e, *row_h;   // old E / new B at strip and band edges
  double2 *ghost_e;          // y-slab with an upper neighbour: old Ez of the high ghost column (see fused_kernels.cu)
  // "vacuum row-strips" (fused_kernels.cu): bit i of vac[band][cta strip] = every cell of row r0+i of that tile
  // has eps == 1, is updated, and is not an NTFF sample cell -> E == D there and the E arrays are not kept
  unsigned long long *vac;
  int vac_strip_w, vac_band_h, vac_edges;   // geometry the masks were built for
  unsigned vac_eps_epoch, vac_ntff_epoch;   // ... and the uploads they reflect
  bool vac_built;
  unsigned long long vac_cells;             // cells in flagged row-strips
};

// Peer (NVLink) halo state of a y-slab engine: the neighbours' field arrays and flag words,
// opened through CUDA IPC.  flags[0]: "the lower neighbour's H-phase halo of step n has
// landed in my low ghost column" (value n+1); flags[1]: the same for the upper neighbour's
// E-phase halo in my high ghost column.
struct PeerState {
  bool attached[2];                  // [0] lower neighbour, [1] upper neighbour
  unsigned long long *flags;         // device, 2 words, owned
  double2 *up_h, *down_e;            // neighbours' H / E arrays (peer pointers), or nullptr
  unsigned long long *up_flag;       // the upper neighbour's flags[0]
  unsigned long long *down_flag;     // the lower neighbour's flags[1]
  void *opened[6];                   // IPC mappings to close
  int up_pitch, down_pitch, down_nj;
  unsigned long long pending_down;   // E-phase signal held back until the NTFF sample of the step has run (0 = none)
};

struct b200fdtd_engine {
  b200fdtd_grid g;
  int device;
  cudaStream_t stream;
  bool own_stream;
  int pitch, rows;
  size_t plane;             // rows * pitch elements (one simulation)
  int n_batch;              // simulations stacked in every field array: element offset b * plane
  int sel;                  // simulation the getters / NTFF read-out refer to
  b200fdtd_batch_source *batch_src;   // device [n_batch], or nullptr
  bool have_batch_src;
  b200fdtd_batch_cw *batch_cw;        // device [n_batch]: batched split-field engines, or nullptr
  bool have_batch_cw;
  int n_fields;
  bool fp32;                // optional single-precision path: the arrays below then hold
  size_t csize, rsize;      //   float2 / float elements (csize = 8, rsize = 4) behind the same pointers
  double2 *field[B200FDTD_MAX_FIELDS];
  double *eps[2];
  double *dense[B200FDTD_MAX_DENSE];   // split-field kinds: coefficients + source factors
  bool have_dense[B200FDTD_MAX_DENSE];
  double *tab_i;            // device [B200FDTD_UPML_TABS][rows], indexed by row = i + 1
  double *tab_j;            // device [B200FDTD_UPML_TABS][pitch], indexed by in-row offset
  bool have_tabs, have_eps[2];
  bool split_lean;          // split-field kinds 0/1/6: 1-D tables + eps / G arrays instead of 8 dense arrays
  // kind 7 (NS TE): rectangle (layout coordinates, inclusive; empty as lo > hi) where the four decay
  // coefficients are exactly 1.0 and C_HZXLX == C_HZYLY: three coefficient arrays instead of eight there
  int split_in_r_lo, split_in_r_hi, split_in_c_lo, split_in_c_hi;
  NtffState ntff;
  FusedState fused;
  PeerState peer;
  // multi-step replay (b200fdtd_run_steps): device clock {time} and the cached graph of a chunk
  double *clock_dev;        // device: [0] = time of the step being computed
  bool clock_mode;          // launchers build views that read the time from clock_dev
  void *graph_exec;         // cudaGraphExec_t of `graph_steps` steps, or nullptr
  int graph_steps;
  uint64_t graph_launches;  // kernels one replay of the cached graph launches
  b200fdtd_cw *cw_tab;      // device [kSplitChunk][2]: per-step CW records of a replayed chunk (split kinds)
  b200fdtd_cw *cw_stage[2]; // pinned host twins, used alternately
  void *cw_stage_done[2];   // cudaEvent_t: the copy out of cw_stage[b] has run
  const b200fdtd_cw *split_cw_step;   // while capturing: the record pair of the step being launched
  unsigned graph_epoch, graph_built_epoch;   // bumped by anything that changes what a step launches
  bool f32_pairs;           // single precision: two cells per thread (default on)
  bool use_fused;           // b200fdtd_step runs the one-pass kernel (serial UPML kinds) wherever it can
  bool fused_auto;          // ... or on large grids only (default)
  bool store_h;             // the fused kernel also writes Hx/Hy (264 instead of 232 B/cell)
  bool h_stale;             // Hx/Hy arrays lag Bx/By (fused step without store_h)
  bool derived_e;           // option: the one-pass step may skip the E arrays in vacuum row-strips
  bool e_consistent;        // E == D/eps (+ pulse) in every updated cell: true from rest and after any full step
                            //   without point / line source; false after b200fdtd_set_field
  bool e_stale;             // the E arrays lag D in the flagged row-strips (b200_refresh_e brings them up to date)
  unsigned eps_epoch, ntff_epoch;   // bumped by eps uploads / NTFF plans
  int fused_variant;        // launch shape of the fused kernel (tuning)
  int unit_split;           // frame-free rectangle through the unit-coefficient kernels: 0 off, 1 on, 2 auto
  bool lean_interior;       // opt-in: cells outside the absorbing frame skip the M / J recurrences
  int lean_r_lo, lean_r_hi, lean_c_lo, lean_c_hi;   // that region (layout coordinates), from the tables
  uint64_t launches;
  uint64_t dev_bytes;
  cudaEvent_t ev0, ev1;
  // update extents in layout coordinates (row r = i + 1, column c = j - j0 + JOFF), inclusive
  int r_lo, r_hi, c_lo, c_hi;
};

// error plumbing (engine.cu)
int b200_fail(int code, const char *fmt, ...);
#define B200_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t err__ = (call);                                                      \
    if (err__ != cudaSuccess)                                                        \
      return b200_fail(B200FDTD_ERR_CUDA, "%s failed: %s (%s:%d)", #call,            \
                       cudaGetErrorString(err__), __FILE__, __LINE__);               \
  } while (0)

// launchers (upml_kernels.cu)
int b200_launch_upml_h(b200fdtd_engine *e, const b200fdtd_step_args *a);
int b200_step_form(const b200fdtd_engine *e);
int b200_split_geometry(const int updated[4], const int interior[4], int out[5][7], int *n_out);   // 0 one kernel per phase, 1 unit-coefficient interior + frame, 2 lean interior + frame
int b200_launch_upml_e(b200fdtd_engine *e, const b200fdtd_step_args *a);
int b200_launch_halo(b200fdtd_engine *e, int which, void *buf, bool pack);
int b200_peer_wait(b200fdtd_engine *e, int which, unsigned long long value);
int b200_peer_signal(b200fdtd_engine *e, unsigned long long *peer_flag, unsigned long long value);
int b200_selftest_division(double divisor, unsigned long long samples, unsigned long long *mismatches);

// single-precision support (upml_kernels.cu): widen / narrow between the float2 device
// arrays and double2 staging planes of the same pitched shape, typed halo column copies
int b200_widen_plane(b200fdtd_engine *e, const void *src_c64, double2 *dst, size_t count);
int b200_narrow_region(b200fdtd_engine *e, const double2 *src_plane, void *dst_c64);
int b200_narrow_real_region(b200fdtd_engine *e, const double *src_region, size_t ld, float *dst_plane);
int b200_fill_float(b200fdtd_engine *e, float *dst, size_t n, float value);
int b200_derive_h_f32(b200fdtd_engine *e, int b_slot, int h_slot);

// launchers (split_kernels.cu)
int b200_launch_split_step(b200fdtd_engine *e, const b200fdtd_step_args *a);

// launchers (fused_kernels.cu)
int b200_launch_upml_fused(b200fdtd_engine *e, const b200fdtd_step_args *a);
bool b200_want_fused(const b200fdtd_engine *e, const b200fdtd_step_args *a);
int b200_launch_fused_edge(b200fdtd_engine *e, const b200fdtd_step_args *a);
int b200_refresh_h(b200fdtd_engine *e);
int b200_refresh_e(b200fdtd_engine *e);
bool b200_fused_derives_e(const b200fdtd_engine *e);
int b200_fused_prepare(b200fdtd_engine *e);
void b200_fused_release(b200fdtd_engine *e);

// launchers (ntff_kernels.cu)
int b200_launch_ntff_sample(b200fdtd_engine *e, const b200fdtd_step_args *a);
int b200_launch_clock_advance(b200fdtd_engine *e);
int b200_launch_ntff_project(b200fdtd_engine *e);
int b200_run_ntff_spectrum(b200fdtd_engine *e, const b200fdtd_spectrum_args *s, double *out);
int b200_run_ntff_frequency(b200fdtd_engine *e, const b200fdtd_freq_args *a, double *out);
