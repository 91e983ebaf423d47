// upml_common.cuh -- pieces shared by the UPML kernels (two-kernel and fused forms):
// the kernel-side view of an engine, complex helpers with the reference's
// component-wise semantics, and the Gaussian-pulse source term.
#pragma once
#include "engine.h"

namespace upml {

constexpr int kBlock = 256;

struct UpmlView {
  double2 *f[B200FDTD_MAX_FIELDS];
  const double *eps0, *eps1;
  const double *ti, *tj;
  int pitch, rows;
  int r_lo, r_hi, c_lo, c_hi;
  int nbx;                      // thread blocks per row (two-kernel form)
  int j_base;                   // global j = j_base + c
  double mu0;
  b200fdtd_pulse pulse[2];
  long long point_k;            // layout offset of the opt-in point source, or -1
  double point_re, point_im;
};

__device__ __forceinline__ double2 operator+(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 operator-(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 operator-(double2 a) { return make_double2(-a.x, -a.y); }
__device__ __forceinline__ double2 operator*(double r, double2 z) { return make_double2(r * z.x, r * z.y); }
__device__ __forceinline__ double2 operator/(double2 z, double r) { return make_double2(z.x / r, z.y / r); }

// field_scatteredPulse (field.c:243-254) for one cell; i, j are GLOBAL indices.
__device__ __forceinline__ double2 pulse_term(const b200fdtd_pulse &s, int i, int j, double eps)
{
  const double r = ((i + s.gap_x) * s.cos_per_c + (j + s.gap_y) * s.sin_per_c) - s.time_minus_t0;
  const double q = r / s.beam_width;
  const double gauss = exp(-(q * q));
  const double amp = s.dot * gauss * (1.0 / eps - 1);
  double sn, cs;
  sincos(r * s.omega, &sn, &cs);
  return make_double2(amp * cs, amp * sn);
}

inline bool is_tm(int kind) { return kind == B200FDTD_TM_UPML || kind == B200FDTD_MPI_TM_UPML; }

inline UpmlView make_view(const b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  UpmlView v;
  for (int s = 0; s < B200FDTD_MAX_FIELDS; s++) v.f[s] = e->field[s];
  v.eps0 = e->eps[0];
  v.eps1 = e->eps[1];
  v.ti = e->tab_i;
  v.tj = e->tab_j;
  v.pitch = e->pitch;
  v.rows = e->rows;
  v.r_lo = e->r_lo;
  v.r_hi = e->r_hi;
  v.c_lo = e->c_lo;
  v.c_hi = e->c_hi;
  v.nbx = (e->c_hi - e->c_lo + 1 + kBlock - 1) / kBlock;
  v.j_base = e->g.j0 - B200_JOFF;
  v.mu0 = e->g.mu0;
  v.pulse[0] = a->pulse[0];
  v.pulse[1] = a->pulse[1];
  v.point_k = -1;
  v.point_re = a->point.re;
  v.point_im = a->point.im;
  if (a->point.enabled) {
    const int pj = a->point.j - e->g.j0;
    if (pj >= 0 && pj < e->g.nj && a->point.i >= 0 && a->point.i < e->g.n_px)
      v.point_k = (long long)(a->point.i + 1) * e->pitch + pj + B200_JOFF;
  }
  return v;
}

}  // namespace upml
