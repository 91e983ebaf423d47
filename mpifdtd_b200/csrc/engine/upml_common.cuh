// upml_common.cuh -- pieces shared by the UPML kernels (two-kernel and fused forms):
// the kernel-side view of an engine, complex helpers with the reference's
// component-wise semantics, and the Gaussian-pulse source term.
#pragma once
#include <cstring>
#include "engine.h"

namespace upml {

constexpr int kBlock = 256;

// complex vector type of a real type: double -> double2 (the reference's arithmetic),
// float -> float2 (the optional single-precision path, its own tolerance)
template <typename T> struct Cx;
template <> struct Cx<double> { using type = double2; };
template <> struct Cx<float>  { using type = float2; };

template <typename T> struct ConstDivisorT { T d, r; };   // a loop-invariant divisor and RN(1/d), see div_const()
using ConstDivisor = ConstDivisorT<double>;

// one rectangle of a multi-rectangle launch, layout coordinates, inclusive
struct LaunchRect {
  int r_lo, r_hi, c_lo, c_hi;
  int bw_log2;                  // a block covers 2^bw_log2 columns x (kBlock >> bw_log2) rows
  int nbx;                      // blocks per block-row
  unsigned blk_end;             // cumulative block count up to and including this rectangle
};

template <typename T>
struct UpmlViewT {
  using C = typename Cx<T>::type;
  C *f[B200FDTD_MAX_FIELDS];
  const T *eps0, *eps1;
  const T *ti, *tj;
  int pitch, rows;
  size_t plane;                 // elements per simulation; blockIdx.y selects the simulation of a batch
  const b200fdtd_batch_source *batch;   // per-simulation sources of a batched engine, or nullptr
  double time;                  // step_args.time (batched pulses form time - t0 themselves)
  const double *time_ptr;       // multi-step replay: the time lives in a device-side clock instead
  int r_lo, r_hi, c_lo, c_hi;
  int nbx;                      // thread blocks per row (two-kernel form)
  int j_base;                   // global j = j_base + c
  ConstDivisorT<T> mu0;         // MU_0_S and its rounded reciprocal
  b200fdtd_pulse pulse[2];
  b200fdtd_cw cw[2];            // CW source of the MPI-variant kinds (mpiTM_UPML.c:337-374)
  // direct halo stores into the neighbour slabs' ghost columns over NVLink (peer memory):
  C *peer_up_h;                 // upper neighbour's H array (Hx / Hz), or nullptr
  C *peer_down_e;               // lower neighbour's E array (Ez / Ex), or nullptr
  int peer_up_pitch, peer_down_pitch, peer_down_col;
  int c_first, c_last;          // first / last owned column in layout coordinates
  b200fdtd_line_source line;    // opt-in planeWave line source (mpiTM_UPML.c:377-403)
  long long point_k;            // layout offset of the opt-in point source, or -1
  double point_re, point_im;
  // launches over a table of rectangles (lean interior form: the frame-free rectangle, or the
  // up to four pieces of the frame around it); see locate_rect() in upml_kernels.cu
  LaunchRect rect[4];
};
using UpmlView = UpmlViewT<double>;

__device__ __forceinline__ double2 operator+(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 operator-(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 operator-(double2 a) { return make_double2(-a.x, -a.y); }
__device__ __forceinline__ double2 operator*(double r, double2 z) { return make_double2(r * z.x, r * z.y); }
__device__ __forceinline__ double2 operator/(double2 z, double r) { return make_double2(z.x / r, z.y / r); }

__device__ __forceinline__ float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 operator-(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 operator-(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 operator*(float r, float2 z) { return make_float2(r * z.x, r * z.y); }
// a source term is always evaluated in double (it touches material cells only) and then
// rounded once to the field type
__device__ __forceinline__ double2 to_field(double2 z, double2) { return z; }
__device__ __forceinline__ float2 to_field(double2 z, float2) { return make_float2((float)z.x, (float)z.y); }
__device__ __forceinline__ double2 add_source(double2 f, double2 s) { return f + s; }
__device__ __forceinline__ float2 add_source(float2 f, double2 s) { return f + make_float2((float)s.x, (float)s.y); }

// ---- exact division shortcuts ---------------------------------------------------------
// The reference divides by MU_0_S four times and by eps twice per TM cell, and two of its
// coefficients are quotients (fdtdTM_upml.c:175,209,216,270-271).  IEEE double division
// costs ~50 SASS instructions on the GPU; these helpers return the SAME correctly rounded
// quotient for a fraction of that, so results stay bit-identical to `x / d`:
//   * x / d with a loop-invariant d:  q = x*r, q' = fma(fma(-q, d, x), r, q) with
//     r = RN(1/d) is the correctly rounded quotient (Markstein 1990) as long as nothing
//     over/underflows; outside a conservative normal range, and never for x == 0, the
//     plain division is used.  b200fdtd_selftest_division() checks the claim on the device
//     against `/` over random bit patterns.
//   * x / 1.0 == x and d / d == 1.0 exactly, so vacuum cells and non-PML coefficients skip
//     the division altogether.
// The full IEEE division lives behind a real call so the compiler cannot hoist its ~100
// instructions out of the (rare) branches below and execute them speculatively.
static __device__ __noinline__ double ieee_div(double x, double d) { return x / d; }

__device__ __forceinline__ double div_const(double x, const ConstDivisor c)
{
  // Markstein: q = RN(x*r); rem = x - q*d exactly (FMA); RN(q + rem*r) == RN(x/d).
  // rem == 0 means q is already the exact quotient (this also keeps the sign of a zero).
  const double q = x * c.r;
  const double rem = fma(-q, c.d, x);
  double res = (rem == 0.0) ? q : fma(rem, c.r, q);
  // The proof needs x, q and rem to be normal numbers.  One integer test on the exponent
  // field: biased exponent in [93, 1953] (|x| roughly 1e-280 .. 1e280), or x == +-0.
  const unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu;
  const bool exotic = (hi - 0x05d00000u) >= (0x7a100000u - 0x05d00000u);
  if (exotic && x != 0.0) res = ieee_div(x, c.d);       // subnormal-ish, huge, inf, nan
  return res;
}
__device__ __forceinline__ double2 div_const(double2 z, const ConstDivisor c)
{
  return make_double2(div_const(z.x, c), div_const(z.y, c));
}
// num / den where both are usually the same number (2*eps outside the PML)
__device__ __forceinline__ double quotient_or_one(double num, double den)
{
  double q = 1.0;
  if (num != den) q = ieee_div(num, den);
  return q;
}
// x / d for an ARBITRARY divisor (a cell's permittivity), exactly, given y = RN(1/d) -- which the
// caller forms once per cell with the correctly rounded reciprocal __drcp_rn and shares between the
// two components and the source term's 1/eps.  Two Markstein steps: q0 = RN(x*y) is within 1.5 ulp
// of x/d, the first correction makes it faithful (its error before rounding is ~1e-16 ulp), and from
// a faithful quotient the second is the correctly rounded one (Markstein 1990, Cornea et al.); a
// zero residual means the quotient is already exact (and keeps the sign of a zero).  Operands whose
// products could leave the normal range take the IEEE path, as in div_const.  Checked against `/`
// on the device over 2^31 random (x, d) pairs: b200fdtd_selftest_division with divisor 0.
__device__ __forceinline__ double div_exact(double x, double d, double y)
{
  const double q0 = x * y;
  const double r0 = fma(-q0, d, x);
  double res = q0;
  if (r0 != 0.0) {
    const double q1 = fma(r0, y, q0);
    const double r1 = fma(-q1, d, x);
    res = (r1 == 0.0) ? q1 : fma(r1, y, q1);
  }
  const unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu;
  const bool exotic = (hi - 0x05d00000u) >= (0x7a100000u - 0x05d00000u);
  if (exotic && x != 0.0) res = ieee_div(x, d);
  return res;
}
// z / eps where eps is usually exactly 1 (vacuum); material cells: permittivities are O(1..100),
// anything outside [2^-20, 2^20] goes through the IEEE division
__device__ __forceinline__ double2 div_eps(double2 z, double eps, double inv_eps)
{
  return make_double2(div_exact(z.x, eps, inv_eps), div_exact(z.y, eps, inv_eps));
}
__device__ __forceinline__ bool eps_is_plain(double eps) { return eps > 9.5367431640625e-07 && eps < 1048576.0; }
__device__ __forceinline__ double2 div_eps(double2 z, double eps)
{
  if (eps != 1.0) {
    if (eps_is_plain(eps)) z = div_eps(z, eps, __drcp_rn(eps));
    else z = make_double2(ieee_div(z.x, eps), ieee_div(z.y, eps));
  }
  return z;
}

// Single-precision path: no bit-exactness contract, so a division by a loop-invariant is one
// multiplication by the rounded reciprocal and the rest is plain float arithmetic.
__device__ __forceinline__ float2 div_const(float2 z, const ConstDivisorT<float> c)
{
  return make_float2(z.x * c.r, z.y * c.r);
}
__device__ __forceinline__ float quotient_or_one(float num, float den) { return num == den ? 1.0f : num / den; }
__device__ __forceinline__ float2 div_eps(float2 z, float eps)
{
  if (eps != 1.0f) z = make_float2(z.x / eps, z.y / eps);
  return z;
}

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

// Adding the pulse to a field component, with the shortcut the pulse's own shape offers:
// exp(-(r/w)^2) is EXACTLY +0 once |r| > 28.3 w (the true value lies below half the smallest
// subnormal), the term is then a signed zero, and x + (+-0) == x for every x except x == -0.  So far
// from the pulse -- most material cells at any given step -- the exp, the sincos and the 1/eps are
// skipped and the bits stay what field.c:248-253 produces; a component that IS -0 takes the full path.
__device__ __forceinline__ bool neg_zero(double v) { return __double_as_longlong(v) == (long long)0x8000000000000000ull; }
__device__ __forceinline__ bool neg_zero(float v) { return __float_as_int(v) == (int)0x80000000u; }
template <typename C>
__device__ __forceinline__ C pulse_add(C e, const b200fdtd_pulse &s, int i, int j, double eps)
{
  const double r = ((i + s.gap_x) * s.cos_per_c + (j + s.gap_y) * s.sin_per_c) - s.time_minus_t0;
  if (fabs(r) > 28.3 * s.beam_width && !neg_zero(e.x) && !neg_zero(e.y)) return e;
  return add_source(e, pulse_term(s, i, j, eps));
}

// The pulse of source slot m for this block's simulation: the step's own parameters, or -- in a
// batched engine -- the per-simulation record with time - t0 formed here (field.c:241,251).
template <typename T>
__device__ __forceinline__ b200fdtd_pulse pulse_of(const UpmlViewT<T> &v, int m)
{
  if (v.batch == nullptr) return v.pulse[m];
  b200fdtd_pulse p = v.batch[blockIdx.y].pulse[m];
  const double time = v.time_ptr != nullptr ? *v.time_ptr : v.time;
  p.time_minus_t0 = time - v.batch[blockIdx.y].t0[m];
  return p;
}
template <typename T>
__device__ __forceinline__ bool pulse_on(const UpmlViewT<T> &v, int m)
{
  return v.batch == nullptr ? v.pulse[m].enabled != 0 : v.batch[blockIdx.y].pulse[m].enabled != 0;
}

// scatteredWave of the MPI solvers (mpiTM_UPML.c:366-370, mpiTE_UPML.c:270-279):
// p += ray_coef*(eps0/eps - 1)*cexp(i(kr - w t)); i, j are GLOBAL indices.
__device__ __forceinline__ double2 cw_eps_term(const b200fdtd_cw &s, int i, int j, double eps)
{
  const double kr = (i + s.gap_x) * s.ks_cos + (j + s.gap_y) * s.ks_sin;
  double sn, cs;
  sincos(kr - s.phase_a, &sn, &cs);
  const double amp = s.scale * (1.0 / eps - 1.0);
  return make_double2(amp * cs, amp * sn);
}

// planeWave (mpiTM_UPML.c:396-400): kr = (x*ks_cos + y*ks_sin) - time; p += ray_coef*cexp(I*kr*w_s)
__device__ __forceinline__ double2 line_term(const b200fdtd_line_source &s, int i, int j)
{
  const double kr = (i * s.ks_cos + j * s.ks_sin) - s.time;
  double sn, cs;
  sincos(kr * s.omega, &sn, &cs);
  return make_double2(s.scale * cs, s.scale * sn);
}

inline bool is_tm(int kind) { return kind == B200FDTD_TM_UPML || kind == B200FDTD_MPI_TM_UPML; }

template <typename T>
inline UpmlViewT<T> make_view_t(const b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  using C = typename Cx<T>::type;
  UpmlViewT<T> v;
  for (int s = 0; s < B200FDTD_MAX_FIELDS; s++) v.f[s] = (C *)e->field[s];
  v.eps0 = (const T *)e->eps[0];
  v.eps1 = (const T *)e->eps[1];
  v.ti = (const T *)e->tab_i;
  v.tj = (const T *)e->tab_j;
  v.pitch = e->pitch;
  v.rows = e->rows;
  v.plane = e->plane;
  v.batch = (e->n_batch > 1 || e->clock_mode) ? e->batch_src : nullptr;
  v.time = a->time;
  v.time_ptr = e->clock_mode ? e->clock_dev : nullptr;
  v.r_lo = e->r_lo;
  v.r_hi = e->r_hi;
  v.c_lo = e->c_lo;
  v.c_hi = e->c_hi;
  v.nbx = (e->c_hi - e->c_lo + 1 + kBlock - 1) / kBlock;
  v.j_base = e->g.j0 - B200_JOFF;
  v.mu0.d = (T)e->g.mu0;
  v.mu0.r = (T)(1.0 / e->g.mu0);
  v.pulse[0] = a->pulse[0];
  v.pulse[1] = a->pulse[1];
  v.cw[0] = a->cw[0];
  v.cw[1] = a->cw[1];
  v.peer_up_h = (C *)e->peer.up_h;
  v.peer_down_e = (C *)e->peer.down_e;
  v.peer_up_pitch = e->peer.up_pitch;
  v.peer_down_pitch = e->peer.down_pitch;
  v.peer_down_col = B200_JOFF + e->peer.down_nj;   // the lower neighbour's high ghost column
  v.c_first = B200_JOFF;
  v.c_last = B200_JOFF + e->g.nj - 1;
  v.line = a->line;
  v.point_k = -1;
  v.point_re = a->point.re;
  v.point_im = a->point.im;
  if (a->point.enabled) {
    const int pj = a->point.j - e->g.j0;
    if (pj >= 0 && pj < e->g.nj && a->point.i >= 0 && a->point.i < e->g.n_px)
      v.point_k = (long long)(a->point.i + 1) * e->pitch + pj + B200_JOFF;
  }
  memset(v.rect, 0, sizeof v.rect);
  return v;
}
inline UpmlView make_view(const b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  return make_view_t<double>(e, a);
}

}  // namespace upml
