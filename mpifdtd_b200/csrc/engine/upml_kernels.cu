// upml_kernels.cu -- UPML leapfrog kernels for sm_100a (TM and TE).
//
// What they replace (rennone/mpiFDTD):
//   H phase  = calcMB + calcH   fdtdTM_upml.c:181-219 / fdtdTE_upml.c:293-314
//   E phase  = calcJD + calcE   fdtdTM_upml.c:155-178 / fdtdTE_upml.c:252-291
//              + field_scatteredPulse                     field.c:224-256
// The reference makes 7 separate sweeps per step over 9 complex fields and 15
// dense coefficient arrays (416 B/cell); here a step is two streaming kernels
// that touch each field once and read coefficients from 1-D tables (296 B/cell
// for TM: H phase reads Ez,Mx,Bx,My,By and writes Mx,Bx,My,By,Hx,Hy; E phase reads
// Hx,Hy,Jz,Dz,eps and writes Jz,Dz,Ez).
//
// Arithmetic contract: every expression keeps the reference's operand order and
// is compiled with -fmad=false, so each add/mul/div is the same IEEE-754 double
// operation gcc emits for the C99 source (real x complex is component-wise, as
// gcc lowers it).  One thread updates one cell; a warp covers 32 consecutive j
// = 512 contiguous bytes per field, every access a 128-bit LDG/STG.
#include <cstring>
#include <type_traits>
#include "upml_common.cuh"

#ifndef B200_H_MIN_BLOCKS
#define B200_H_MIN_BLOCKS 1
#endif
#ifndef B200_E_MIN_BLOCKS
#define B200_E_MIN_BLOCKS 5   /* <= 51 registers: measured 23.4 -> 24.4 Gcell/s at 16384^2 */
#endif
#ifndef B200_TE_H_MIN_BLOCKS
#define B200_TE_H_MIN_BLOCKS 1
#endif
#ifndef B200_TE_E_MIN_BLOCKS
#define B200_TE_E_MIN_BLOCKS 4   /* measured 20.9 -> 22.7 Gcell/s at 16384^2 (1: 20.9, 5: 20.8, 6: 19.9) */
#endif

namespace {

using namespace upml;

// Cell owned by this thread, or false when past the row end.
// k0 indexes what all simulations of a batch share (eps, the point source); k = k0 + the
// simulation's plane offset (blockIdx.y, 0 for an unbatched engine) indexes the fields.
template <typename T>
__device__ __forceinline__ bool locate(const UpmlViewT<T> &v, int &r, int &c, size_t &k, size_t &k0)
{
  const long long b = blockIdx.x;
  const int rb = (int)(b / v.nbx);
  const int cb = (int)(b - (long long)rb * v.nbx);
  r = v.r_lo + rb;
  c = v.c_lo + cb * kBlock + (int)threadIdx.x;
  k0 = (size_t)r * (size_t)v.pitch + (size_t)c;
  k = k0 + (size_t)blockIdx.y * v.plane;
  return c <= v.c_hi;
}

// The same for a launch over a table of up to four rectangles (the absorbing frame around the
// lean interior, or the interior itself): blocks are numbered rectangle after rectangle; a block
// covers 2^bw_log2 columns x (kBlock >> bw_log2) rows, so the narrow side strips of the frame do
// not waste 245 of 256 threads.  Which rectangle: three compares on launch constants.
template <typename T>
__device__ __forceinline__ bool locate_rect(const UpmlViewT<T> &v, int &r, int &c, size_t &k, size_t &k0)
{
  unsigned b = blockIdx.x;
  LaunchRect R = v.rect[0];
  if (b >= v.rect[2].blk_end)      { R = v.rect[3]; b -= v.rect[2].blk_end; }
  else if (b >= v.rect[1].blk_end) { R = v.rect[2]; b -= v.rect[1].blk_end; }
  else if (b >= v.rect[0].blk_end) { R = v.rect[1]; b -= v.rect[0].blk_end; }
  const int rb = (int)(b / (unsigned)R.nbx);
  const int cb = (int)(b - (unsigned)rb * (unsigned)R.nbx);
  const int tx = (int)threadIdx.x & ((1 << R.bw_log2) - 1), ty = (int)threadIdx.x >> R.bw_log2;
  r = R.r_lo + rb * (kBlock >> R.bw_log2) + ty;
  c = R.c_lo + (cb << R.bw_log2) + tx;
  k0 = (size_t)r * (size_t)v.pitch + (size_t)c;
  k = k0 + (size_t)blockIdx.y * v.plane;
  return r <= R.r_hi && c <= R.c_hi;
}

// ---- the arithmetic of one cell, shared by the one-cell-per-thread kernels and the
// two-cells-per-thread single-precision kernels.  Expression order is the
// reference's (see the citations); compiled with -fmad=false.
template <typename T, typename C = typename Cx<T>::type>
__device__ __forceinline__ void tm_h_math(C ez, C ez_j1, C ez_i1, C mx_old, C bx_old, C my_old, C by_old, T c_mx,
                                          T c_mxez, T num1, T num0, T c_bx1, T c_bx0, T c_by, T den, C &mx, C &bx,
                                          C &my, C &by)
{
  // fdtdTM_upml.c:187-189 (C_BX == 1 exactly)
  mx = c_mx * mx_old - c_mxez * (ez_j1 - ez);
  bx = (bx_old + c_bx1 * mx) - c_bx0 * mx_old;
  // fdtdTM_upml.c:196-198 (C_MY == C_MYEZ == 1 exactly)
  my = my_old - ((-ez_i1) + ez);
  const T c_by1 = quotient_or_one(num1, den), c_by0 = quotient_or_one(num0, den);   // fdtdTM_upml.c:270-271
  by = (c_by * by_old + c_by1 * my) - c_by0 * my_old;
}

// the source terms the reference adds to Ez after calcE (shared by the full and the lean E phase)
template <typename T, typename C = typename Cx<T>::type>
__device__ __forceinline__ void tm_e_sources(const UpmlViewT<T> &v, int r, int c, size_t k0, T eps, C &ez)
{
  if (eps != (T)1 && pulse_on(v, 0))       // field.c:248
    ez = pulse_add(ez, pulse_of(v, 0), r - 1, v.j_base + c, (double)eps);
  if (v.cw[0].enabled && eps != (T)1)          // mpiTM_UPML.c:370
    ez = add_source(ez, cw_eps_term(v.cw[0], r - 1, v.j_base + c, (double)eps));
  if ((long long)k0 == v.point_k)
    ez = add_source(ez, make_double2(v.point_re, v.point_im));
  if (v.line.enabled && r - 1 == v.line.i) {  // block-uniform: one grid row
    const int j = v.j_base + c;
    if (j >= v.line.j_lo && j <= v.line.j_hi) ez = add_source(ez, line_term(v.line, r - 1, j));
  }
}

template <typename T, typename C>
__device__ __forceinline__ void te_e_sources(const UpmlViewT<T> &v, int r, int c, size_t k0, T eps_x, T eps_y, C &ex,
                                             C &ey);

// Rare work of the lean E kernels, kept out of line so the streaming path stays within 32
// registers (8 blocks/SM): material cells (eps != 1: the division and the source terms with their
// exp / sincos) and the opt-in point / line sources.  `v` is the kernel's __grid_constant__
// parameter, so passing its address copies nothing.  (The same trick does nothing for the full
// kernels: forcing them to 6-8 blocks/SM spills the coefficient arithmetic, measured slower.)
template <typename T, typename C = typename Cx<T>::type>
static __device__ __noinline__ C tm_e_material(const UpmlViewT<T> *v, int r, int c, size_t k0, T eps, C dz)
{
  C ez = div_eps(dz, eps);
  tm_e_sources<T>(*v, r, c, k0, eps, ez);
  return ez;
}
template <typename T, typename C = typename Cx<T>::type>
static __device__ __noinline__ void te_e_material(const UpmlViewT<T> *v, int r, int c, size_t k0, T eps_x, T eps_y,
                                                     C dx, C dy, C *ex, C *ey)
{
  C x = div_eps(dx, eps_x), y = div_eps(dy, eps_y);
  te_e_sources<T, C>(*v, r, c, k0, eps_x, eps_y, x, y);
  *ex = x;
  *ey = y;
}

template <typename T, typename C = typename Cx<T>::type>
__device__ __forceinline__ void tm_e_math(const UpmlViewT<T> &v, int r, int c, size_t k0, C hy, C hy_i0, C hx, C hx_j0,
                                          C jz_old, C dz_old, T eps, T c_jz, T c_jzh, T c_dz, T c_dzjz, C &jz, C &dz,
                                          C &ez)
{
  // fdtdTM_upml.c:161-163 (C_DZJZ1 == C_DZJZ0 because sigma_z = 0)
  jz = c_jz * jz_old + c_jzh * (((hy - hy_i0) - hx) + hx_j0);
  dz = (c_dz * dz_old + c_dzjz * jz) - c_dzjz * jz_old;
  ez = div_eps(dz, eps);              // fdtdTM_upml.c:175
  tm_e_sources<T>(v, r, c, k0, eps, ez);
}

template <typename T, typename C = typename Cx<T>::type>
__device__ __forceinline__ void te_h_math(C ey_i1, C ey, C ex_j1, C ex, C mz_old, C bz_old, T c_mz, T c_mze, T c_bz,
                                          T c_bzmz, C &mz, C &bz)
{
  // fdtdTE_upml.c:299-301 (C_BZMZ1 == C_BZMZ0 because sigma_z = 0)
  mz = c_mz * mz_old - c_mze * (((ey_i1 - ey) - ex_j1) + ex);
  bz = (c_bz * bz_old + c_bzmz * mz) - c_bzmz * mz_old;
}

// the source terms the reference adds to Ex / Ey after calcE (shared by the full and the lean E phase)
template <typename T, typename C>
__device__ __forceinline__ void te_e_sources(const UpmlViewT<T> &v, int r, int c, size_t k0, T eps_x, T eps_y, C &ex,
                                             C &ey)
{
  const int i = r - 1, j = v.j_base + c;
  if (eps_x != (T)1 && pulse_on(v, 0))     // fdtdTE_upml.c:186-187
    ex = pulse_add(ex, pulse_of(v, 0), i, j, (double)eps_x);
  if (eps_y != (T)1 && pulse_on(v, 1))     // fdtdTE_upml.c:188-189
    ey = pulse_add(ey, pulse_of(v, 1), i, j, (double)eps_y);
  if (v.cw[0].enabled && eps_x != (T)1) ex = add_source(ex, cw_eps_term(v.cw[0], i, j, (double)eps_x));
  if (v.cw[1].enabled && eps_y != (T)1) ey = add_source(ey, cw_eps_term(v.cw[1], i, j, (double)eps_y));   // mpiTE_UPML.c:278
  if ((long long)k0 == v.point_k)
    ex = add_source(ex, make_double2(v.point_re, v.point_im));
}

template <typename T, typename C = typename Cx<T>::type>
__device__ __forceinline__ void te_e_math(const UpmlViewT<T> &v, int r, int c, size_t k0, C hz, C hz_j0, C hz_i0,
                                          C jx_old, C dx_old, C jy_old, C dy_old, T eps_x, T eps_y, T c_jx, T c_jxhz,
                                          T num1, T num0, T c_dx1, T c_dx0, T c_dy, T den, C &jx, C &dx, C &jy, C &dy,
                                          C &ex, C &ey)
{
  // fdtdTE_upml.c:259-262 (C_DX == 1)
  jx = c_jx * jx_old + c_jxhz * (hz - hz_j0);
  dx = (dx_old + c_dx1 * jx) - c_dx0 * jx_old;
  // fdtdTE_upml.c:269-271 (C_JY == C_JYHZ == 1)
  jy = jy_old + ((-hz) + hz_i0);
  const T c_dy1 = quotient_or_one(num1, den), c_dy0 = quotient_or_one(num0, den);   // fdtdTE_upml.c:402-403
  dy = (c_dy * dy_old + c_dy1 * jy) - c_dy0 * jy_old;

  ex = div_eps(dx, eps_x);            // fdtdTE_upml.c:283
  ey = div_eps(dy, eps_y);            // fdtdTE_upml.c:289
  te_e_sources<T, C>(v, r, c, k0, eps_x, eps_y, ex, ey);
}

// ------------------------------------------------------------------ TM -----
// slots: 0 Ez 1 Jz 2 Dz 3 Hx 4 Mx 5 Bx 6 Hy 7 My 8 By
// STORE_H = false: Hx/Hy are not written; the E phase recomputes them from Bx/By
// (Hx == Bx/mu0 exactly, fdtdTM_upml.c:209), which removes one 32 B/cell write and turns
// the E phase's H reads into B reads: 264 instead of 296 B per cell-update, in place.
template <typename T, bool STORE_H>
__device__ __forceinline__ void tm_upml_h_cell(const UpmlViewT<T> &v, int r, int c, size_t k, size_t k0)
{
  using C = typename Cx<T>::type;
  (void)k0;
  const C *__restrict__ Ez = v.f[B200FDTD_TM_EZ];

  const C ez = Ez[k];
  const C ez_j1 = Ez[k + 1];            // Ez(i, j+1)
  const C ez_i1 = Ez[k + v.pitch];      // Ez(i+1, j)
  const C mx_old = v.f[B200FDTD_TM_MX][k];
  const C bx_old = v.f[B200FDTD_TM_BX][k];
  const C my_old = v.f[B200FDTD_TM_MY][k];
  const C by_old = v.f[B200FDTD_TM_BY][k];

  const T c_mx   = v.tj[B200FDTD_TMJ_C_MX * v.pitch + c];
  const T c_mxez = v.tj[B200FDTD_TMJ_C_MXEZ * v.pitch + c];
  const T num1   = v.tj[B200FDTD_TMJ_NUM_BYMY1 * v.pitch + c];
  const T num0   = v.tj[B200FDTD_TMJ_NUM_BYMY0 * v.pitch + c];
  const T c_bx1  = v.ti[B200FDTD_TMI_C_BXMX1 * v.rows + r];
  const T c_bx0  = v.ti[B200FDTD_TMI_C_BXMX0 * v.rows + r];
  const T c_by   = v.ti[B200FDTD_TMI_C_BY * v.rows + r];
  const T den    = v.ti[B200FDTD_TMI_DEN_BYMY * v.rows + r];

  C mx, bx, my, by;
  tm_h_math<T>(ez, ez_j1, ez_i1, mx_old, bx_old, my_old, by_old, c_mx, c_mxez, num1, num0, c_bx1, c_bx0, c_by, den,
               mx, bx, my, by);

  v.f[B200FDTD_TM_MX][k] = mx;
  v.f[B200FDTD_TM_BX][k] = bx;
  v.f[B200FDTD_TM_MY][k] = my;
  v.f[B200FDTD_TM_BY][k] = by;
  if (STORE_H) {
    v.f[B200FDTD_TM_HX][k] = div_const(bx, v.mu0);   // fdtdTM_upml.c:209
    v.f[B200FDTD_TM_HY][k] = div_const(by, v.mu0);   // fdtdTM_upml.c:216
  }
  // y-slab halo: my top owned column of Hx is the upper neighbour's low ghost column
  if (v.peer_up_h != nullptr && c == v.c_last)
    v.peer_up_h[(size_t)r * v.peer_up_pitch + (B200_JOFF - 1)] = div_const(bx, v.mu0);
}

template <typename T, bool STORE_H, bool RECTS = false>
__global__ void __launch_bounds__(kBlock, B200_H_MIN_BLOCKS) tm_upml_h_kernel(const __grid_constant__ UpmlViewT<T> v)
{
  int r, c; size_t k, k0;
  if (!(RECTS ? locate_rect(v, r, c, k, k0) : locate(v, r, c, k, k0))) return;
  tm_upml_h_cell<T, STORE_H>(v, r, c, k, k0);
}

// FROM_B = true: H is formed on the fly as B/mu0.  Cells just outside the updated range
// (the ring, or a neighbour slab's halo column) are not derived state: there the H array
// itself is read, exactly like the STORE_H form does.
template <typename T, bool FROM_B>
__device__ __forceinline__ void tm_upml_e_cell(const UpmlViewT<T> &v, int r, int c, size_t k, size_t k0)
{
  using C = typename Cx<T>::type;
  C hy, hy_i0, hx, hx_j0;
  if (FROM_B) {
    const C *__restrict__ Bx = v.f[B200FDTD_TM_BX];
    const C *__restrict__ By = v.f[B200FDTD_TM_BY];
    // all four loads are issued unconditionally; the (block-uniform, rare) edge cases
    // then replace the derived value by the stored one
    const C by = By[k], bx = Bx[k], by_i0 = By[k - v.pitch],
            bx_j0 = Bx[k - 1];
    hy = div_const(by, v.mu0);
    hx = div_const(bx, v.mu0);
    hy_i0 = div_const(by_i0, v.mu0);
    hx_j0 = div_const(bx_j0, v.mu0);
    if (r == v.r_lo) hy_i0 = v.f[B200FDTD_TM_HY][k - v.pitch];
    if (c == v.c_lo) hx_j0 = v.f[B200FDTD_TM_HX][k - 1];
  } else {
    const C *__restrict__ Hx = v.f[B200FDTD_TM_HX];
    const C *__restrict__ Hy = v.f[B200FDTD_TM_HY];
    hy = Hy[k];
    hy_i0 = Hy[k - v.pitch];      // Hy(i-1, j)
    hx = Hx[k];
    hx_j0 = Hx[k - 1];            // Hx(i, j-1)
  }
  const C jz_old = v.f[B200FDTD_TM_JZ][k];
  const C dz_old = v.f[B200FDTD_TM_DZ][k];
  const T eps = v.eps0[k0];

  const T c_jz   = v.ti[B200FDTD_TMI_C_JZ * v.rows + r];
  const T c_jzh  = v.ti[B200FDTD_TMI_C_JZHXHY * v.rows + r];
  const T c_dz   = v.tj[B200FDTD_TMJ_C_DZ * v.pitch + c];
  const T c_dzjz = v.tj[B200FDTD_TMJ_C_DZJZ * v.pitch + c];

  C jz, dz, ez;
  tm_e_math<T>(v, r, c, k0, hy, hy_i0, hx, hx_j0, jz_old, dz_old, eps, c_jz, c_jzh, c_dz, c_dzjz, jz, dz, ez);

  v.f[B200FDTD_TM_JZ][k] = jz;
  v.f[B200FDTD_TM_DZ][k] = dz;
  v.f[B200FDTD_TM_EZ][k] = ez;
  // y-slab halo: my bottom owned column of Ez is the lower neighbour's high ghost column
  if (v.peer_down_e != nullptr && c == v.c_first)
    v.peer_down_e[(size_t)r * v.peer_down_pitch + v.peer_down_col] = ez;
}

template <typename T, bool FROM_B, bool RECTS = false>
__global__ void __launch_bounds__(kBlock, B200_E_MIN_BLOCKS) tm_upml_e_kernel(const __grid_constant__ UpmlViewT<T> v)
{
  int r, c; size_t k, k0;
  if (!(RECTS ? locate_rect(v, r, c, k, k0) : locate(v, r, c, k, k0))) return;
  tm_upml_e_cell<T, FROM_B>(v, r, c, k, k0);
}

// ------------------------------------------------------------------ TE -----
// slots: 0 Ex 1 Jx 2 Dx 3 Ey 4 Jy 5 Dy 6 Hz 7 Mz 8 Bz
template <typename T, bool STORE_H>
__device__ __forceinline__ void te_upml_h_cell(const UpmlViewT<T> &v, int r, int c, size_t k, size_t k0)
{
  using C = typename Cx<T>::type;
  (void)k0;
  const C *__restrict__ Ex = v.f[B200FDTD_TE_EX];
  const C *__restrict__ Ey = v.f[B200FDTD_TE_EY];

  const C ey_i1 = Ey[k + v.pitch];
  const C ey = Ey[k];
  const C ex_j1 = Ex[k + 1];
  const C ex = Ex[k];
  const C mz_old = v.f[B200FDTD_TE_MZ][k];
  const C bz_old = v.f[B200FDTD_TE_BZ][k];

  const T c_mz   = v.ti[B200FDTD_TEI_C_MZ * v.rows + r];
  const T c_mze  = v.ti[B200FDTD_TEI_C_MZEXEY * v.rows + r];
  const T c_bz   = v.tj[B200FDTD_TEJ_C_BZ * v.pitch + c];
  const T c_bzmz = v.tj[B200FDTD_TEJ_C_BZMZ * v.pitch + c];

  C mz, bz;
  te_h_math<T>(ey_i1, ey, ex_j1, ex, mz_old, bz_old, c_mz, c_mze, c_bz, c_bzmz, mz, bz);

  v.f[B200FDTD_TE_MZ][k] = mz;
  v.f[B200FDTD_TE_BZ][k] = bz;
  if (STORE_H) v.f[B200FDTD_TE_HZ][k] = div_const(bz, v.mu0);   // fdtdTE_upml.c:312
  if (v.peer_up_h != nullptr && c == v.c_last)
    v.peer_up_h[(size_t)r * v.peer_up_pitch + (B200_JOFF - 1)] = div_const(bz, v.mu0);
}

template <typename T, bool STORE_H, bool RECTS = false>
__global__ void __launch_bounds__(kBlock, B200_TE_H_MIN_BLOCKS) te_upml_h_kernel(const __grid_constant__ UpmlViewT<T> v)
{
  int r, c; size_t k, k0;
  if (!(RECTS ? locate_rect(v, r, c, k, k0) : locate(v, r, c, k, k0))) return;
  te_upml_h_cell<T, STORE_H>(v, r, c, k, k0);
}

template <typename T, bool FROM_B>
__device__ __forceinline__ void te_upml_e_cell(const UpmlViewT<T> &v, int r, int c, size_t k, size_t k0)
{
  using C = typename Cx<T>::type;
  const C *__restrict__ Hz = v.f[B200FDTD_TE_HZ];
  C hz, hz_j0, hz_i0;
  if (FROM_B) {                               // Hz == Bz/mu0 (fdtdTE_upml.c:312), formed on the fly
    const C *__restrict__ Bz = v.f[B200FDTD_TE_BZ];
    const C bz = Bz[k], bz_j0 = Bz[k - 1], bz_i0 = Bz[k - v.pitch];
    hz = div_const(bz, v.mu0);
    hz_j0 = div_const(bz_j0, v.mu0);
    hz_i0 = div_const(bz_i0, v.mu0);
    if (c == v.c_lo) hz_j0 = Hz[k - 1];       // ring / halo column: only the H array holds it
    if (r == v.r_lo) hz_i0 = Hz[k - v.pitch];
  } else {
    hz = Hz[k];
    hz_j0 = Hz[k - 1];
    hz_i0 = Hz[k - v.pitch];
  }
  const C jx_old = v.f[B200FDTD_TE_JX][k];
  const C dx_old = v.f[B200FDTD_TE_DX][k];
  const C jy_old = v.f[B200FDTD_TE_JY][k];
  const C dy_old = v.f[B200FDTD_TE_DY][k];
  const T eps_x = v.eps0[k0], eps_y = v.eps1[k0];

  const T c_jx   = v.tj[B200FDTD_TEJ_C_JX * v.pitch + c];
  const T c_jxhz = v.tj[B200FDTD_TEJ_C_JXHZ * v.pitch + c];
  const T num1   = v.tj[B200FDTD_TEJ_NUM_DYJY1 * v.pitch + c];
  const T num0   = v.tj[B200FDTD_TEJ_NUM_DYJY0 * v.pitch + c];
  const T c_dx1  = v.ti[B200FDTD_TEI_C_DXJX1 * v.rows + r];
  const T c_dx0  = v.ti[B200FDTD_TEI_C_DXJX0 * v.rows + r];
  const T c_dy   = v.ti[B200FDTD_TEI_C_DY * v.rows + r];
  const T den    = v.ti[B200FDTD_TEI_DEN_DYJY * v.rows + r];

  C jx, dx, jy, dy, ex, ey;
  te_e_math<T>(v, r, c, k0, hz, hz_j0, hz_i0, jx_old, dx_old, jy_old, dy_old, eps_x, eps_y, c_jx, c_jxhz, num1, num0,
               c_dx1, c_dx0, c_dy, den, jx, dx, jy, dy, ex, ey);

  v.f[B200FDTD_TE_JX][k] = jx;
  v.f[B200FDTD_TE_DX][k] = dx;
  v.f[B200FDTD_TE_JY][k] = jy;
  v.f[B200FDTD_TE_DY][k] = dy;
  v.f[B200FDTD_TE_EX][k] = ex;
  v.f[B200FDTD_TE_EY][k] = ey;
  if (v.peer_down_e != nullptr && c == v.c_first)
    v.peer_down_e[(size_t)r * v.peer_down_pitch + v.peer_down_col] = ex;
}

template <typename T, bool FROM_B, bool RECTS = false>
__global__ void __launch_bounds__(kBlock, B200_TE_E_MIN_BLOCKS) te_upml_e_kernel(const __grid_constant__ UpmlViewT<T> v)
{
  int r, c; size_t k, k0;
  if (!(RECTS ? locate_rect(v, r, c, k, k0) : locate(v, r, c, k, k0))) return;
  te_upml_e_cell<T, FROM_B>(v, r, c, k, k0);
}

// ------------------------------------------------------------------ unit-coefficient interior -----
// The DEFAULT form on large grids.  Inside the frame-free rectangle every coefficient the full
// kernels would fetch is exactly 1.0, and 1.0 * x == x for every double, so the reference's
// expressions can be evaluated there without the twelve table reads, the two quotients and the
// multiplications -- with the identical IEEE results, operation for operation:
//   Mx' = Mx - (Ez(j+1) - Ez),  Bx' = (Bx + Mx') - Mx,  Jz' = Jz + curl H,  Dz' = (Dz + Jz') - Jz ...
// Same arrays, same 264 / 288 B per cell-update, bit-identical to the one-kernel-per-phase form
// (tests/test_gpu_unit.py); what it buys is registers (more blocks per SM, more loads in flight).
// The frame goes through the full kernels (RECTS = true), exactly as in the lean form below.
// blocks/SM measured at 16384^2 (round 1, A/B builds): TM H 4/5/6/7/8 -> 5.60/5.60/5.58/5.67/5.67 ms,
// TM E 4/5/6/7 -> 5.41/4.93/5.00/5.03 ms, TE E 4/5/6/7/8 -> 8.65/8.68/8.60/10.0/10.0 ms
#ifndef B200_UNIT_H_MIN_BLOCKS
#define B200_UNIT_H_MIN_BLOCKS 6
#endif
#ifndef B200_UNIT_E_MIN_BLOCKS
#define B200_UNIT_E_MIN_BLOCKS 5
#endif
#ifndef B200_UNIT_TE_H_MIN_BLOCKS
#define B200_UNIT_TE_H_MIN_BLOCKS 6
#endif
#ifndef B200_UNIT_TE_E_MIN_BLOCKS
#define B200_UNIT_TE_E_MIN_BLOCKS 6
#endif

template <typename T, bool STORE_H>
__global__ void __launch_bounds__(kBlock, B200_UNIT_H_MIN_BLOCKS) tm_unit_h_kernel(const __grid_constant__ UpmlViewT<T> v)
{
  using C = typename Cx<T>::type;
  int r, c; size_t k, k0;
  if (!locate_rect(v, r, c, k, k0)) return;
  const C *__restrict__ Ez = v.f[B200FDTD_TM_EZ];
  const C ez = Ez[k], ez_j1 = Ez[k + 1], ez_i1 = Ez[k + v.pitch];
  const C mx_old = v.f[B200FDTD_TM_MX][k], bx_old = v.f[B200FDTD_TM_BX][k];
  const C my_old = v.f[B200FDTD_TM_MY][k], by_old = v.f[B200FDTD_TM_BY][k];
  const C mx = mx_old - (ez_j1 - ez);              // fdtdTM_upml.c:188 with C_MX = C_MXEZ = 1
  const C bx = (bx_old + mx) - mx_old;             // :189 with C_BX = C_BXMX1 = C_BXMX0 = 1
  const C my = my_old - ((-ez_i1) + ez);           // :197
  const C by = (by_old + my) - my_old;             // :198 with C_BY = C_BYMY1 = C_BYMY0 = 1
  v.f[B200FDTD_TM_MX][k] = mx;
  v.f[B200FDTD_TM_BX][k] = bx;
  v.f[B200FDTD_TM_MY][k] = my;
  v.f[B200FDTD_TM_BY][k] = by;
  if (STORE_H) {
    v.f[B200FDTD_TM_HX][k] = div_const(bx, v.mu0);
    v.f[B200FDTD_TM_HY][k] = div_const(by, v.mu0);
  }
  if (v.peer_up_h != nullptr && c == v.c_last)
    v.peer_up_h[(size_t)r * v.peer_up_pitch + (B200_JOFF - 1)] = div_const(bx, v.mu0);
}

template <typename T, bool FROM_B>
__global__ void __launch_bounds__(kBlock, B200_UNIT_E_MIN_BLOCKS) tm_unit_e_kernel(const __grid_constant__ UpmlViewT<T> v)
{
  using C = typename Cx<T>::type;
  int r, c; size_t k, k0;
  if (!locate_rect(v, r, c, k, k0)) return;
  C hy, hy_i0, hx, hx_j0;
  if (FROM_B) {       // the rectangle never touches row r_lo / column c_lo: every neighbour is a kept B value
    const C *__restrict__ Bx = v.f[B200FDTD_TM_BX];
    const C *__restrict__ By = v.f[B200FDTD_TM_BY];
    const C by = By[k], bx = Bx[k], by_i0 = By[k - v.pitch], bx_j0 = Bx[k - 1];
    hy = div_const(by, v.mu0);
    hx = div_const(bx, v.mu0);
    hy_i0 = div_const(by_i0, v.mu0);
    hx_j0 = div_const(bx_j0, v.mu0);
  } else {
    const C *__restrict__ Hx = v.f[B200FDTD_TM_HX];
    const C *__restrict__ Hy = v.f[B200FDTD_TM_HY];
    hy = Hy[k]; hy_i0 = Hy[k - v.pitch]; hx = Hx[k]; hx_j0 = Hx[k - 1];
  }
  const C jz_old = v.f[B200FDTD_TM_JZ][k];
  const C dz_old = v.f[B200FDTD_TM_DZ][k];
  const T eps = v.eps0[k0];
  const C jz = jz_old + (((hy - hy_i0) - hx) + hx_j0);   // fdtdTM_upml.c:162 with C_JZ = C_JZHXHY = 1
  const C dz = (dz_old + jz) - jz_old;                   // :163 with C_DZ = C_DZJZ1 = C_DZJZ0 = 1
  v.f[B200FDTD_TM_JZ][k] = jz;
  v.f[B200FDTD_TM_DZ][k] = dz;
  C ez = dz;                                             // :175 with eps == 1: x / 1.0 == x
  if (eps != (T)1 || (long long)k0 == v.point_k || (v.line.enabled && r - 1 == v.line.i))
    ez = tm_e_material<T>(&v, r, c, k0, eps, dz);
  v.f[B200FDTD_TM_EZ][k] = ez;
  if (v.peer_down_e != nullptr && c == v.c_first)
    v.peer_down_e[(size_t)r * v.peer_down_pitch + v.peer_down_col] = ez;
}

template <typename T, bool STORE_H>
__global__ void __launch_bounds__(kBlock, B200_UNIT_TE_H_MIN_BLOCKS) te_unit_h_kernel(const __grid_constant__ UpmlViewT<T> v)
{
  using C = typename Cx<T>::type;
  int r, c; size_t k, k0;
  if (!locate_rect(v, r, c, k, k0)) return;
  const C *__restrict__ Ex = v.f[B200FDTD_TE_EX];
  const C *__restrict__ Ey = v.f[B200FDTD_TE_EY];
  const C ey_i1 = Ey[k + v.pitch], ey = Ey[k], ex_j1 = Ex[k + 1], ex = Ex[k];
  const C mz_old = v.f[B200FDTD_TE_MZ][k], bz_old = v.f[B200FDTD_TE_BZ][k];
  const C mz = mz_old - (((ey_i1 - ey) - ex_j1) + ex);   // fdtdTE_upml.c:300 with C_MZ = C_MZEXEY = 1
  const C bz = (bz_old + mz) - mz_old;                   // :301 with C_BZ = C_BZMZ1 = C_BZMZ0 = 1
  v.f[B200FDTD_TE_MZ][k] = mz;
  v.f[B200FDTD_TE_BZ][k] = bz;
  if (STORE_H) v.f[B200FDTD_TE_HZ][k] = div_const(bz, v.mu0);
  if (v.peer_up_h != nullptr && c == v.c_last)
    v.peer_up_h[(size_t)r * v.peer_up_pitch + (B200_JOFF - 1)] = div_const(bz, v.mu0);
}

template <typename T, bool FROM_B>
__global__ void __launch_bounds__(kBlock, B200_UNIT_TE_E_MIN_BLOCKS) te_unit_e_kernel(const __grid_constant__ UpmlViewT<T> v)
{
  using C = typename Cx<T>::type;
  int r, c; size_t k, k0;
  if (!locate_rect(v, r, c, k, k0)) return;
  C hz, hz_j0, hz_i0;
  if (FROM_B) {
    const C *__restrict__ Bz = v.f[B200FDTD_TE_BZ];
    hz = div_const(Bz[k], v.mu0);
    hz_j0 = div_const(Bz[k - 1], v.mu0);
    hz_i0 = div_const(Bz[k - v.pitch], v.mu0);
  } else {
    const C *__restrict__ Hz = v.f[B200FDTD_TE_HZ];
    hz = Hz[k]; hz_j0 = Hz[k - 1]; hz_i0 = Hz[k - v.pitch];
  }
  const C jx_old = v.f[B200FDTD_TE_JX][k], dx_old = v.f[B200FDTD_TE_DX][k];
  const C jy_old = v.f[B200FDTD_TE_JY][k], dy_old = v.f[B200FDTD_TE_DY][k];
  const T eps_x = v.eps0[k0], eps_y = v.eps1[k0];
  const C jx = jx_old + (hz - hz_j0);                    // fdtdTE_upml.c:260 with C_JX = C_JXHZ = 1
  const C dx = (dx_old + jx) - jx_old;                   // :261
  const C jy = jy_old + ((-hz) + hz_i0);                 // :270
  const C dy = (dy_old + jy) - jy_old;                   // :271 with C_DY = C_DYJY1 = C_DYJY0 = 1
  v.f[B200FDTD_TE_JX][k] = jx;
  v.f[B200FDTD_TE_DX][k] = dx;
  v.f[B200FDTD_TE_JY][k] = jy;
  v.f[B200FDTD_TE_DY][k] = dy;
  C ex = dx, ey = dy;
  if (eps_x != (T)1 || eps_y != (T)1 || (long long)k0 == v.point_k)
    te_e_material<T>(&v, r, c, k0, eps_x, eps_y, dx, dy, &ex, &ey);
  v.f[B200FDTD_TE_EX][k] = ex;
  v.f[B200FDTD_TE_EY][k] = ey;
  if (v.peer_down_e != nullptr && c == v.c_first)
    v.peer_down_e[(size_t)r * v.peer_down_pitch + v.peer_down_col] = ex;
}

// ------------------------------------------------------------------ lean interior -----
// Opt-in (B200FDTD_OPT_LEAN_INTERIOR).  Outside the absorbing frame every UPML coefficient of
// fdtdTM_upml.c:253-271 / fdtdTE_upml.c:384-403 is exactly 1, and the recurrences collapse:
//   Mx' = Mx - d,  Bx' = (Bx + Mx') - Mx   ==>  Bx' = Bx - d          (d = Ez(j+1) - Ez)
//   Jz' = Jz + q,  Dz' = (Dz + Jz') - Jz   ==>  Dz' = Dz + q          (q = curl H = curl B / mu0)
// These kernels run over that rectangle only and neither read nor write M / J: 80 + 88 = 168 B
// per TM cell-update instead of 144 + 120 = 264 (TE: 64 + 128 = 192 instead of 288); the frame
// keeps the reference's arithmetic through the kernels above (RECTS = true).  Same mathematics,
// fewer roundings -- curl H is formed as RN(1/mu0) * curl B, one multiplication instead of four
// divisions -- so this form carries a tolerance (fields <= 1e-12 of the reference) instead of the
// bit-for-bit contract.  Only differences of M / J enter the update out here, so the arrays may
// go stale and the form can be switched between steps.  Few registers, no division: 8 blocks/SM.
#ifndef B200_LEAN_MIN_BLOCKS
#define B200_LEAN_MIN_BLOCKS 8
#endif

template <typename T, bool STORE_H>
__global__ void __launch_bounds__(kBlock, B200_LEAN_MIN_BLOCKS) tm_lean_h_kernel(const __grid_constant__ UpmlViewT<T> v)
{
  using C = typename Cx<T>::type;
  int r, c; size_t k, k0;
  if (!locate_rect(v, r, c, k, k0)) return;
  const C *__restrict__ Ez = v.f[B200FDTD_TM_EZ];
  const C ez = Ez[k], ez_j1 = Ez[k + 1], ez_i1 = Ez[k + v.pitch];
  const C bx = v.f[B200FDTD_TM_BX][k] - (ez_j1 - ez);
  const C by = v.f[B200FDTD_TM_BY][k] - ((-ez_i1) + ez);
  v.f[B200FDTD_TM_BX][k] = bx;
  v.f[B200FDTD_TM_BY][k] = by;
  if (STORE_H) {
    v.f[B200FDTD_TM_HX][k] = div_const(bx, v.mu0);
    v.f[B200FDTD_TM_HY][k] = div_const(by, v.mu0);
  }
  if (v.peer_up_h != nullptr && c == v.c_last)
    v.peer_up_h[(size_t)r * v.peer_up_pitch + (B200_JOFF - 1)] = div_const(bx, v.mu0);
}

// (The lean rectangle never touches row r_lo or column c_lo -- see b200fdtd_set_upml_tables -- so
// every neighbour read below is a B value this engine keeps up to date.)
template <typename T, bool FROM_B>
__global__ void __launch_bounds__(kBlock, B200_LEAN_MIN_BLOCKS) tm_lean_e_kernel(const __grid_constant__ UpmlViewT<T> v)
{
  using C = typename Cx<T>::type;
  int r, c; size_t k, k0;
  if (!locate_rect(v, r, c, k, k0)) return;
  C curl;
  if (FROM_B) {
    const C *__restrict__ Bx = v.f[B200FDTD_TM_BX];
    const C *__restrict__ By = v.f[B200FDTD_TM_BY];
    curl = v.mu0.r * (((By[k] - By[k - v.pitch]) - Bx[k]) + Bx[k - 1]);
  } else {
    const C *__restrict__ Hx = v.f[B200FDTD_TM_HX];
    const C *__restrict__ Hy = v.f[B200FDTD_TM_HY];
    curl = ((Hy[k] - Hy[k - v.pitch]) - Hx[k]) + Hx[k - 1];
  }
  const T eps = v.eps0[k0];
  const C dz = v.f[B200FDTD_TM_DZ][k] + curl;
  v.f[B200FDTD_TM_DZ][k] = dz;
  C ez = dz;
  if (eps != (T)1 || (long long)k0 == v.point_k || (v.line.enabled && r - 1 == v.line.i))
    ez = tm_e_material<T>(&v, r, c, k0, eps, dz);
  v.f[B200FDTD_TM_EZ][k] = ez;
  if (v.peer_down_e != nullptr && c == v.c_first)
    v.peer_down_e[(size_t)r * v.peer_down_pitch + v.peer_down_col] = ez;
}

template <typename T, bool STORE_H>
__global__ void __launch_bounds__(kBlock, B200_LEAN_MIN_BLOCKS) te_lean_h_kernel(const __grid_constant__ UpmlViewT<T> v)
{
  using C = typename Cx<T>::type;
  int r, c; size_t k, k0;
  if (!locate_rect(v, r, c, k, k0)) return;
  const C *__restrict__ Ex = v.f[B200FDTD_TE_EX];
  const C *__restrict__ Ey = v.f[B200FDTD_TE_EY];
  const C ey_i1 = Ey[k + v.pitch], ey = Ey[k], ex_j1 = Ex[k + 1], ex = Ex[k];
  const C bz = v.f[B200FDTD_TE_BZ][k] - (((ey_i1 - ey) - ex_j1) + ex);
  v.f[B200FDTD_TE_BZ][k] = bz;
  if (STORE_H) v.f[B200FDTD_TE_HZ][k] = div_const(bz, v.mu0);
  if (v.peer_up_h != nullptr && c == v.c_last)
    v.peer_up_h[(size_t)r * v.peer_up_pitch + (B200_JOFF - 1)] = div_const(bz, v.mu0);
}

template <typename T, bool FROM_B>
__global__ void __launch_bounds__(kBlock, B200_LEAN_MIN_BLOCKS) te_lean_e_kernel(const __grid_constant__ UpmlViewT<T> v)
{
  using C = typename Cx<T>::type;
  int r, c; size_t k, k0;
  if (!locate_rect(v, r, c, k, k0)) return;
  C curl_x, curl_y;                     // Hz - Hz(j-1) and -Hz + Hz(i-1)
  if (FROM_B) {
    const C *__restrict__ Bz = v.f[B200FDTD_TE_BZ];
    const C bz = Bz[k];
    curl_x = v.mu0.r * (bz - Bz[k - 1]);
    curl_y = v.mu0.r * ((-bz) + Bz[k - v.pitch]);
  } else {
    const C *__restrict__ Hz = v.f[B200FDTD_TE_HZ];
    const C hz = Hz[k];
    curl_x = hz - Hz[k - 1];
    curl_y = (-hz) + Hz[k - v.pitch];
  }
  const T eps_x = v.eps0[k0], eps_y = v.eps1[k0];
  const C dx = v.f[B200FDTD_TE_DX][k] + curl_x;
  const C dy = v.f[B200FDTD_TE_DY][k] + curl_y;
  v.f[B200FDTD_TE_DX][k] = dx;
  v.f[B200FDTD_TE_DY][k] = dy;
  C ex = dx, ey = dy;
  if (eps_x != (T)1 || eps_y != (T)1 || (long long)k0 == v.point_k)
    te_e_material<T>(&v, r, c, k0, eps_x, eps_y, dx, dy, &ex, &ey);
  v.f[B200FDTD_TE_EX][k] = ex;
  v.f[B200FDTD_TE_EY][k] = ey;
  if (v.peer_down_e != nullptr && c == v.c_first)
    v.peer_down_e[(size_t)r * v.peer_down_pitch + v.peer_down_col] = ex;
}

#include "upml_pairs_f32.cuh"

// One halo column <-> a contiguous buffer of n_px complex values (always double complex on
// the wire, whatever the engine's precision).
// divisor.d != 0: `field` holds B and the packed value is H = B/divisor (H arrays not kept).
template <typename T>
__global__ void halo_column_kernel(typename Cx<T>::type *field, double2 *buf, int pitch, int col, int n_px,
                                   int pack, ConstDivisorT<T> divisor, int derive)
{
  using C = typename Cx<T>::type;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_px) return;
  const size_t k = (size_t)(i + 1) * pitch + col;
  if (pack) {
    C v = field[k];
    if (derive) v = div_const(v, divisor);
    buf[i] = make_double2((double)v.x, (double)v.y);
  } else {
    const double2 w = buf[i];
    C v; v.x = (T)w.x; v.y = (T)w.y;
    field[k] = v;
  }
}

// ---- single-precision support: widen / narrow between float device arrays and double staging
__global__ void widen_plane_kernel(const float2 *__restrict__ src, double2 *dst, size_t n)
{
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    const float2 v = src[k];
    dst[k] = make_double2((double)v.x, (double)v.y);
  }
}
// owned cells only (rows 1..n_px, columns JOFF..JOFF+nj-1): ghosts keep what they hold
__global__ void narrow_region_kernel(const double2 *__restrict__ src, float2 *dst, int pitch, int n_px, int nj)
{
  const size_t n = (size_t)n_px * nj;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const size_t k = (size_t)(1 + t / nj) * pitch + B200_JOFF + t % nj;
    const double2 v = src[k];
    dst[k] = make_float2((float)v.x, (float)v.y);
  }
}
__global__ void narrow_real_region_kernel(const double *__restrict__ src, size_t ld, float *dst, int pitch,
                                          int n_px, int nj)
{
  const size_t n = (size_t)n_px * nj;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const size_t i = t / nj, c = t % nj;
    dst[(size_t)(1 + i) * pitch + B200_JOFF + c] = (float)src[i * ld + c];
  }
}
__global__ void fill_float_kernel(float *dst, size_t n, float value)
{
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    dst[k] = value;
}
// H = B/mu0 exactly as the float STORE_H kernels form it (one multiplication by RN(1/mu0))
__global__ void derive_h_f32_kernel(const float2 *__restrict__ b, float2 *h, int pitch, int r_lo, int n_rows,
                                    int c_lo, int n_cols, ConstDivisorT<float> mu0)
{
  const size_t n = (size_t)n_rows * n_cols;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const size_t k = (size_t)(r_lo + t / n_cols) * pitch + c_lo + t % n_cols;
    h[k] = div_const(b[k], mu0);
  }
}

}  // namespace

// ---- cross-GPU ordering without the host: one flag word per direction in each engine's
// memory.  After a phase whose kernel stored halo values into a neighbour, a one-thread
// kernel publishes the step number into that neighbour's flag (system-scope release);
// before a phase that reads a ghost column, a one-thread kernel spins on the local flag
// (system-scope acquire) until the producing step has been published.  Stream order ties
// both to the phase kernels, so no NCCL call, host barrier or event is needed per step.
namespace {
__global__ void peer_signal_kernel(unsigned long long *peer_flag, unsigned long long value)
{
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flag), "l"(value) : "memory");
}
__global__ void peer_wait_kernel(const unsigned long long *my_flag, unsigned long long value)
{
  unsigned long long seen;
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(my_flag) : "memory");
  } while (seen < value);
}
}  // namespace

int b200_peer_wait(b200fdtd_engine *e, int which, unsigned long long value)
{
  peer_wait_kernel<<<1, 1, 0, e->stream>>>(e->peer.flags + which, value);
  e->launches++;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}

int b200_peer_signal(b200fdtd_engine *e, unsigned long long *peer_flag, unsigned long long value)
{
  peer_signal_kernel<<<1, 1, 0, e->stream>>>(peer_flag, value);
  e->launches++;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}

// Device self-test of div_const / quotient_or_one / div_eps against plain IEEE division.
namespace {
__global__ void division_selftest_kernel(unsigned long long seed, unsigned long long per_thread, double d,
                                         unsigned long long *mismatches)
{
  ConstDivisor c; c.d = d; c.r = 1.0 / d;
  const bool any_divisor = d == 0.0;     // every sample its own divisor, through div_eps (permittivity-like values)
  unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
  unsigned long long bad = 0;
  for (unsigned long long n = 0; n < per_thread; n++) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;                 // xorshift64
    // alternate between raw bit patterns (all exponents, subnormals, inf/nan) and field-like magnitudes
    double v = (n & 1) ? __longlong_as_double((long long)x)
                       : ((double)(long long)x) * ((n & 2) ? 1.0e-19 : 1.0e-31);
    if (any_divisor) {
      unsigned long long y = x * 0x2545F4914F6CDD1Dull;
      // a random 52-bit significand in [1, 2), scaled into [2^-4, 2^8): 1.0 itself, all-ones
      // significands and short decimals like 2.56 are among the patterns
      double dd = __longlong_as_double((long long)((y >> 12) | 0x3ff0000000000000ull));
      if ((n & 12) == 4) dd = 1.0 + (double)((y >> 40) % 2000) * 0.01;
      if ((n & 12) == 8) dd = __longlong_as_double((long long)(0x3fffffffffffffffull - ((y >> 58) & 7)));
      dd = ldexp(dd, (int)((y >> 3) % 12) - 4);
      const double2 got2 = div_eps(make_double2(v, -v), dd);
      const double want = v / dd, want2 = (-v) / dd;
      const bool ok = ((__double_as_longlong(want) == __double_as_longlong(got2.x)) || (want != want && got2.x != got2.x)) &&
                      ((__double_as_longlong(want2) == __double_as_longlong(got2.y)) || (want2 != want2 && got2.y != got2.y));
      if (!ok) bad++;
      continue;
    }
    const double want = v / d, got = div_const(v, c);
    const bool same = (__double_as_longlong(want) == __double_as_longlong(got)) || (want != want && got != got);
    if (!same) bad++;
  }
  if (bad) atomicAdd(mismatches, bad);
}
}  // namespace

int b200_selftest_division(double divisor, unsigned long long samples, unsigned long long *mismatches)
{
  unsigned long long *d_bad = nullptr;
  B200_CUDA(cudaMalloc(&d_bad, sizeof *d_bad));
  B200_CUDA(cudaMemset(d_bad, 0, sizeof *d_bad));
  const unsigned blocks = 592, threads = 256;
  const unsigned long long per_thread = samples / ((unsigned long long)blocks * threads) + 1;
  division_selftest_kernel<<<blocks, threads>>>(0x1234567ull, per_thread, divisor, d_bad);
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpy(mismatches, d_bad, sizeof *d_bad, cudaMemcpyDeviceToHost));
  cudaFree(d_bad);
  return B200FDTD_OK;
}

// single precision, default forms: two cells per thread (upml_pairs_f32.cuh)
static PairGeom pair_geom(const b200fdtd_engine *e)
{
  PairGeom g;
  g.nbx2 = (e->c_hi - (e->c_lo & ~1) + 1 + 2 * kBlock - 1) / (2 * kBlock);
  return g;
}

// ---- launch geometry of the lean interior form -------------------------------------------
static bool have_interior(const b200fdtd_engine *e)
{
  return e->lean_r_hi >= e->lean_r_lo && e->lean_c_hi >= e->lean_c_lo;
}
static bool lean_active(const b200fdtd_engine *e) { return e->lean_interior && have_interior(e); }

// Unit-coefficient interior kernels (bit-identical to the full ones): double precision, and by
// default only where the rectangle is most of a large grid -- on small grids a step is launch-
// bound and one kernel per phase is the better deal.
static bool unit_active(const b200fdtd_engine *e)
{
  if (e->fp32 || e->lean_interior || !have_interior(e) || e->unit_split == 0) return false;
  if (e->unit_split == 1) return true;
  const double inner = (double)(e->lean_r_hi - e->lean_r_lo + 1) * (e->lean_c_hi - e->lean_c_lo + 1);
  const double all = (double)(e->r_hi - e->r_lo + 1) * (e->c_hi - e->c_lo + 1);
  return inner >= 1048576.0 && inner >= 0.75 * all;
}

// blocks of one rectangle; narrow rectangles get narrow, tall blocks
static unsigned add_rect(LaunchRect *rect, int &n, unsigned blk_end, int r_lo, int r_hi, int c_lo, int c_hi)
{
  if (r_hi < r_lo || c_hi < c_lo) return blk_end;
  const int w = c_hi - c_lo + 1, h = r_hi - r_lo + 1;
  int lg = 8;                                   // kBlock = 256 columns x 1 row
  while (lg > 4 && (1 << (lg - 1)) >= w) lg--;  // down to 16 columns x 16 rows
  LaunchRect &R = rect[n++];
  R.r_lo = r_lo; R.r_hi = r_hi; R.c_lo = c_lo; R.c_hi = c_hi; R.bw_log2 = lg;
  R.nbx = (w + (1 << lg) - 1) >> lg;
  const int rows_per_blk = kBlock >> lg;
  R.blk_end = blk_end + (unsigned)((long long)R.nbx * ((h + rows_per_blk - 1) / rows_per_blk));
  return R.blk_end;
}

static void pad_rects(LaunchRect *rect, int n, unsigned blk_end)
{
  for (int q = n; q < 4; q++) {                 // unused entries: empty, never selected
    rect[q] = rect[n ? n - 1 : 0];
    rect[q].blk_end = blk_end;
    rect[q].r_hi = rect[q].r_lo - 1;
  }
}

template <typename T>
static unsigned interior_rect(const b200fdtd_engine *e, UpmlViewT<T> &v)
{
  int n = 0;
  const unsigned end = add_rect(v.rect, n, 0, e->lean_r_lo, e->lean_r_hi, e->lean_c_lo, e->lean_c_hi);
  pad_rects(v.rect, n, end);
  return end;
}

// the updated cells outside the lean rectangle: rows below and above it over the full width, then
// the two side strips beside it
template <typename T>
static unsigned frame_rects(const b200fdtd_engine *e, UpmlViewT<T> &v)
{
  int n = 0;
  unsigned end = 0;
  end = add_rect(v.rect, n, end, e->r_lo, e->lean_r_lo - 1, e->c_lo, e->c_hi);
  end = add_rect(v.rect, n, end, e->lean_r_hi + 1, e->r_hi, e->c_lo, e->c_hi);
  end = add_rect(v.rect, n, end, e->lean_r_lo, e->lean_r_hi, e->c_lo, e->lean_c_lo - 1);
  end = add_rect(v.rect, n, end, e->lean_r_lo, e->lean_r_hi, e->lean_c_hi + 1, e->c_hi);
  if (n == 0) { memset(v.rect, 0, sizeof v.rect); return 0; }
  pad_rects(v.rect, n, end);
  return end;
}

template <typename T>
static int launch_h(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  const UpmlViewT<T> v = make_view_t<T>(e, a);
  if constexpr (std::is_same<T, float>::value) {
    if (e->f32_pairs && !e->store_h && !e->lean_interior) {
      const PairGeom g = pair_geom(e);
      const dim3 grid((unsigned)((long long)g.nbx2 * (e->r_hi - e->r_lo + 1)), (unsigned)e->n_batch);
      if (is_tm(e->g.kind)) tm_upml_h_pair_kernel<<<grid, kBlock, 0, e->stream>>>(v, g);
      else                  te_upml_h_pair_kernel<<<grid, kBlock, 0, e->stream>>>(v, g);
      e->h_stale = true;
      e->launches++;
      B200_CUDA(cudaGetLastError());
      return B200FDTD_OK;
    }
  }
  const long long nblk = (long long)v.nbx * (e->r_hi - e->r_lo + 1);
  const dim3 grid((unsigned)nblk, (unsigned)e->n_batch);
  if (lean_active(e) || unit_active(e)) {
    // the frame-free rectangle through the lean / unit-coefficient kernel, the frame around it
    // through the full one
    const bool lean = lean_active(e);
    UpmlViewT<T> vi = v, vf = v;
    const unsigned nb_i = interior_rect(e, vi), nb_f = frame_rects(e, vf);
    const dim3 gi(nb_i, (unsigned)e->n_batch), gf(nb_f, (unsigned)e->n_batch);
#define INTERIOR_H(KERNEL)                                                        \
    do {                                                                          \
      if (e->store_h) KERNEL<T, true><<<gi, kBlock, 0, e->stream>>>(vi);          \
      else            KERNEL<T, false><<<gi, kBlock, 0, e->stream>>>(vi);         \
    } while (0)
    if (is_tm(e->g.kind)) {
      if (lean) INTERIOR_H(tm_lean_h_kernel); else INTERIOR_H(tm_unit_h_kernel);
      if (nb_f) {
        if (e->store_h) tm_upml_h_kernel<T, true, true><<<gf, kBlock, 0, e->stream>>>(vf);
        else            tm_upml_h_kernel<T, false, true><<<gf, kBlock, 0, e->stream>>>(vf);
      }
    } else {
      if (lean) INTERIOR_H(te_lean_h_kernel); else INTERIOR_H(te_unit_h_kernel);
      if (nb_f) {
        if (e->store_h) te_upml_h_kernel<T, true, true><<<gf, kBlock, 0, e->stream>>>(vf);
        else            te_upml_h_kernel<T, false, true><<<gf, kBlock, 0, e->stream>>>(vf);
      }
    }
#undef INTERIOR_H
    if (nb_f) e->launches++;
  } else if (is_tm(e->g.kind)) {
    if (e->store_h) tm_upml_h_kernel<T, true><<<grid, kBlock, 0, e->stream>>>(v);
    else            tm_upml_h_kernel<T, false><<<grid, kBlock, 0, e->stream>>>(v);
  } else {
    if (e->store_h) te_upml_h_kernel<T, true><<<grid, kBlock, 0, e->stream>>>(v);
    else            te_upml_h_kernel<T, false><<<grid, kBlock, 0, e->stream>>>(v);
  }
  e->h_stale = !e->store_h;
  e->launches++;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}

template <typename T>
static int launch_e(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  const UpmlViewT<T> v = make_view_t<T>(e, a);
  if constexpr (std::is_same<T, float>::value) {
    if (e->f32_pairs && e->h_stale && !e->lean_interior) {
      const PairGeom g = pair_geom(e);
      const dim3 grid((unsigned)((long long)g.nbx2 * (e->r_hi - e->r_lo + 1)), (unsigned)e->n_batch);
      if (is_tm(e->g.kind)) tm_upml_e_pair_kernel<<<grid, kBlock, 0, e->stream>>>(v, g);
      else                  te_upml_e_pair_kernel<<<grid, kBlock, 0, e->stream>>>(v, g);
      e->launches++;
      B200_CUDA(cudaGetLastError());
      return B200FDTD_OK;
    }
  }
  const long long nblk = (long long)v.nbx * (e->r_hi - e->r_lo + 1);
  const dim3 grid((unsigned)nblk, (unsigned)e->n_batch);
  if (lean_active(e) || unit_active(e)) {
    const bool lean = lean_active(e);
    UpmlViewT<T> vi = v, vf = v;
    const unsigned nb_i = interior_rect(e, vi), nb_f = frame_rects(e, vf);
    const dim3 gi(nb_i, (unsigned)e->n_batch), gf(nb_f, (unsigned)e->n_batch);
#define INTERIOR_E(KERNEL)                                                        \
    do {                                                                          \
      if (e->h_stale) KERNEL<T, true><<<gi, kBlock, 0, e->stream>>>(vi);          \
      else            KERNEL<T, false><<<gi, kBlock, 0, e->stream>>>(vi);         \
    } while (0)
    if (is_tm(e->g.kind)) {
      if (lean) INTERIOR_E(tm_lean_e_kernel); else INTERIOR_E(tm_unit_e_kernel);
      if (nb_f) {
        if (e->h_stale) tm_upml_e_kernel<T, true, true><<<gf, kBlock, 0, e->stream>>>(vf);
        else            tm_upml_e_kernel<T, false, true><<<gf, kBlock, 0, e->stream>>>(vf);
      }
    } else {
      if (lean) INTERIOR_E(te_lean_e_kernel); else INTERIOR_E(te_unit_e_kernel);
      if (nb_f) {
        if (e->h_stale) te_upml_e_kernel<T, true, true><<<gf, kBlock, 0, e->stream>>>(vf);
        else            te_upml_e_kernel<T, false, true><<<gf, kBlock, 0, e->stream>>>(vf);
      }
    }
#undef INTERIOR_E
    if (nb_f) e->launches++;
  } else if (is_tm(e->g.kind)) {
    if (e->h_stale) tm_upml_e_kernel<T, true><<<grid, kBlock, 0, e->stream>>>(v);
    else            tm_upml_e_kernel<T, false><<<grid, kBlock, 0, e->stream>>>(v);
  } else {
    if (e->h_stale) te_upml_e_kernel<T, true><<<grid, kBlock, 0, e->stream>>>(v);
    else            te_upml_e_kernel<T, false><<<grid, kBlock, 0, e->stream>>>(v);
  }
  e->launches++;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}

// Host-only view of the launch geometry of the split forms, for tests: rectangle 0 is the interior
// launch, the others the frame launch (block numbering restarts there).  7 ints per rectangle:
// r_lo, r_hi, c_lo, c_hi, bw_log2, nbx, blk_end.
int b200_split_geometry(const int updated[4], const int interior[4], int out[5][7], int *n_out)
{
  b200fdtd_engine e;
  memset(&e, 0, sizeof e);
  e.r_lo = updated[0]; e.r_hi = updated[1]; e.c_lo = updated[2]; e.c_hi = updated[3];
  e.lean_r_lo = interior[0]; e.lean_r_hi = interior[1]; e.lean_c_lo = interior[2]; e.lean_c_hi = interior[3];
  UpmlViewT<double> vi, vf;
  memset(&vi, 0, sizeof vi); memset(&vf, 0, sizeof vf);
  const unsigned nb_i = interior_rect(&e, vi), nb_f = frame_rects(&e, vf);
  int n = 0;
  auto put = [&](const LaunchRect &R) {
    const int row[7] = { R.r_lo, R.r_hi, R.c_lo, R.c_hi, R.bw_log2, R.nbx, (int)R.blk_end };
    memcpy(out[n++], row, sizeof row);
  };
  if (nb_i) put(vi.rect[0]);
  for (int q = 0; q < 4 && nb_f; q++)
    if (vf.rect[q].r_hi >= vf.rect[q].r_lo && (q == 0 || vf.rect[q].blk_end > vf.rect[q - 1].blk_end)) put(vf.rect[q]);
  *n_out = n;
  return B200FDTD_OK;
}

int b200_step_form(const b200fdtd_engine *e)
{
  return lean_active(e) ? 2 : unit_active(e) ? 1 : 0;
}

int b200_launch_upml_h(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  if (e->r_hi < e->r_lo || e->c_hi < e->c_lo) return B200FDTD_OK;   // slab owns no updated cell
  { int rc = b200_refresh_e(e); if (rc) return rc; }                // a one-pass step may have left E derived
  return e->fp32 ? launch_h<float>(e, a) : launch_h<double>(e, a);
}

int b200_launch_upml_e(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  if (e->r_hi < e->r_lo || e->c_hi < e->c_lo) return B200FDTD_OK;
  int rc = b200_refresh_e(e); if (rc) return rc;
  rc = e->fp32 ? launch_e<float>(e, a) : launch_e<double>(e, a);
  // every updated cell now holds E = D/eps + the pulse (material cells only); the opt-in sources also act in
  // vacuum cells (point, line) or add a signed zero there (CW), after which E == D is not given
  if (!rc) e->e_consistent = !(a->line.enabled || a->point.enabled || a->cw[0].enabled || a->cw[1].enabled);
  return rc;
}

template <typename T>
static void launch_halo_t(b200fdtd_engine *e, int slot, void *buf, int col, bool pack, bool derive)
{
  const int n = e->g.n_px;
  ConstDivisorT<T> d; d.d = (T)e->g.mu0; d.r = (T)(1.0 / e->g.mu0);
  halo_column_kernel<T><<<(n + 255) / 256, 256, 0, e->stream>>>((typename Cx<T>::type *)e->field[slot], (double2 *)buf,
                                                               e->pitch, col, n, pack ? 1 : 0, d, derive ? 1 : 0);
}

int b200_launch_halo(b200fdtd_engine *e, int which, void *buf, bool pack)
{
  int slot, col;
  if (which == 0) {           // after the H phase: Hx (TM) / Hz (TE), last owned -> low ghost
    slot = is_tm(e->g.kind) ? (int)B200FDTD_TM_HX : (int)B200FDTD_TE_HZ;
    col = pack ? B200_JOFF + e->g.nj - 1 : B200_JOFF - 1;
  } else {                    // after the E phase: Ez (TM) / Ex (TE), first owned -> high ghost
    slot = is_tm(e->g.kind) ? (int)B200FDTD_TM_EZ : (int)B200FDTD_TE_EX;
    col = pack ? B200_JOFF : B200_JOFF + e->g.nj;
  }
  if (which == 1 && pack) { int rc = b200_refresh_e(e); if (rc) return rc; }   // the E arrays themselves, not D
  bool derive = false;
  if (which == 0 && pack && e->h_stale) {     // H column = B column / mu0 (H arrays not kept)
    slot = is_tm(e->g.kind) ? (int)B200FDTD_TM_BX : (int)B200FDTD_TE_BZ;
    derive = true;
  }
  if (e->fp32) launch_halo_t<float>(e, slot, buf, col, pack, derive);
  else         launch_halo_t<double>(e, slot, buf, col, pack, derive);
  e->launches++;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}

// ---- single-precision support launchers ---------------------------------------------------
int b200_widen_plane(b200fdtd_engine *e, const void *src_c64, double2 *dst, size_t count)
{
  widen_plane_kernel<<<1184, 256, 0, e->stream>>>((const float2 *)src_c64, dst, count);
  e->launches++;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}

int b200_narrow_region(b200fdtd_engine *e, const double2 *src_plane, void *dst_c64)
{
  narrow_region_kernel<<<1184, 256, 0, e->stream>>>(src_plane, (float2 *)dst_c64, e->pitch, e->g.n_px, e->g.nj);
  e->launches++;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}

int b200_narrow_real_region(b200fdtd_engine *e, const double *src_region, size_t ld, float *dst_plane)
{
  narrow_real_region_kernel<<<1184, 256, 0, e->stream>>>(src_region, ld, dst_plane, e->pitch, e->g.n_px, e->g.nj);
  e->launches++;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}

int b200_fill_float(b200fdtd_engine *e, float *dst, size_t n, float value)
{
  fill_float_kernel<<<1184, 256, 0, e->stream>>>(dst, n, value);
  e->launches++;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}

int b200_derive_h_f32(b200fdtd_engine *e, int b_slot, int h_slot)
{
  ConstDivisorT<float> d; d.d = (float)e->g.mu0; d.r = (float)(1.0 / e->g.mu0);
  for (int b = 0; b < e->n_batch; b++) {
    const size_t off = (size_t)b * e->plane;
    derive_h_f32_kernel<<<1184, 256, 0, e->stream>>>((const float2 *)e->field[b_slot] + off,
                                                     (float2 *)e->field[h_slot] + off, e->pitch, e->r_lo,
                                                     e->r_hi - e->r_lo + 1, e->c_lo, e->c_hi - e->c_lo + 1, d);
    e->launches++;
  }
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}
