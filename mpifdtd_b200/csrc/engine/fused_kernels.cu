// fused_kernels.cu -- one-pass UPML step for sm_100a: H phase and E phase fused.
//
// The two-kernel step (upml_kernels.cu) writes Hx/Hy in the H phase and reads them
// back in the E phase.  Here one kernel does both so H never round-trips through
// HBM: per cell it reads Ez,Mx,Bx,My,By,Jz,Dz (112 B) + eps (8 B) and writes
// Mx,Bx,My,By,Jz,Dz,Ez (112 B) [+ Hx,Hy (32 B) when the H arrays are kept]:
// 232 B (264 B) per cell-update against 296 B for the two-kernel form.
//
// Decomposition ("warp-strip marching").  E(i,j) needs the NEW Hy(i-1,j) and
// Hx(i,j-1) (fdtdTM_upml.c:162), i.e. a one-cell low-side skirt of H, and all state
// is updated in place -- a tile that recomputed its skirt would race with the tile
// that owns it.  So:
//   * a warp owns a strip of 32 columns (lane <-> column j, 512 contiguous bytes per
//     field per row, 128-bit accesses) and a band of rows, and marches down the band;
//     Hy(i-1,j) is carried in registers from the previous row, Hx(i,j-1) comes from
//     the neighbouring lane by shuffle, Ez(i,j+1) likewise, Ez(i+1,j) is the next
//     row's Ez(i,j) and is loaded exactly once;
//   * the values a warp would need from ANOTHER warp -- lane 0's Hx(i,j-1), lane 31's
//     old Ez(i,j+1), the first row's Hy(i-1,j), the last row's old Ez(i+1,j) -- are
//     produced by a small pre-pass from the OLD state into side buffers before the
//     main kernel starts (columns at strip edges, rows at band edges).  The pre-pass
//     evaluates the same expressions, so the values are bit-identical to what the
//     owning warp computes.
// After the pre-pass every warp is independent: no barriers, no shared memory, no
// inter-block ordering, in-place update.  Pre-pass + side-buffer traffic is ~3 % of
// the step.  Arithmetic is the two-kernel form's, expression for expression
// (-fmad=false), so both forms produce identical bits (tests/test_gpu_fused.py).
#include <cstring>
#include "upml_common.cuh"

namespace {

using namespace upml;

struct FusedView {
  UpmlView u;
  int n_strips, n_bands, band_h;
  int strip_w, n_edges;         // the column pre-pass serves strips of strip_w columns (32: every warp
                                //   strip; 32*WARPS: only CTA strips, TMA form); n_edges = ceil(cols / strip_w)
  double2 *col_e, *col_h;       // [n_edges + 1][rows]: old E at a strip's first column,
                                //   new H at the column just below it
  double2 *row_e, *row_h;       // [n_bands + 1][pitch]: old E at a band's first row,
                                //   new H at the row just above it
  int unit_r_lo, unit_r_hi, unit_c_lo, unit_c_hi;   // frame-free rectangle (all coefficients exactly 1), or empty
  double2 *ghost_e;             // [rows]: y-slab with an upper neighbour: the OLD Ez of my high ghost column,
                                //   captured by the edge kernel before the neighbour may overwrite it; else nullptr
};

__device__ __forceinline__ double2 shfl_down1(double2 v)
{
  return make_double2(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1));
}
__device__ __forceinline__ double2 shfl_up1(double2 v)
{
  return make_double2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1));
}

// The pulse of source slot 0 as the two-kernel form sees it (pulse_of in upml_common.cuh with
// blockIdx.y == 0): the step's own parameters, or -- multi-step replay -- the engine's source
// record with time - t0 formed from the device clock.
__device__ __forceinline__ b200fdtd_pulse fused_pulse(const UpmlView &v)
{
  if (v.batch == nullptr) return v.pulse[0];
  b200fdtd_pulse p = v.batch[0].pulse[0];
  const double time = v.time_ptr != nullptr ? *v.time_ptr : v.time;
  p.time_minus_t0 = time - v.batch[0].t0[0];
  return p;
}

// ---- TM: the H-phase arithmetic for one cell (fdtdTM_upml.c:187-216) ------------
struct TmH { double2 mx, bx, my, by, hx, hy; };
struct TmColCoef { double c_mx, c_mxez, num1, num0; };          // by j (this lane's column)
struct TmRowCoef { double c_bx1, c_bx0, c_by, den; };           // by i (this row)

__device__ __forceinline__ TmColCoef tm_col_coef(const UpmlView &v, int c)
{
  TmColCoef k;
  k.c_mx   = v.tj[B200FDTD_TMJ_C_MX * v.pitch + c];
  k.c_mxez = v.tj[B200FDTD_TMJ_C_MXEZ * v.pitch + c];
  k.num1   = v.tj[B200FDTD_TMJ_NUM_BYMY1 * v.pitch + c];
  k.num0   = v.tj[B200FDTD_TMJ_NUM_BYMY0 * v.pitch + c];
  return k;
}
__device__ __forceinline__ TmRowCoef tm_row_coef(const UpmlView &v, int r)
{
  TmRowCoef k;
  k.c_bx1 = v.ti[B200FDTD_TMI_C_BXMX1 * v.rows + r];
  k.c_bx0 = v.ti[B200FDTD_TMI_C_BXMX0 * v.rows + r];
  k.c_by  = v.ti[B200FDTD_TMI_C_BY * v.rows + r];
  k.den   = v.ti[B200FDTD_TMI_DEN_BYMY * v.rows + r];
  return k;
}

__device__ __forceinline__ TmH tm_h_cell(const UpmlView &v, const TmColCoef &cc, const TmRowCoef &rc,
                                         double2 ez, double2 ez_j1, double2 ez_i1, double2 mx_old,
                                         double2 bx_old, double2 my_old, double2 by_old)
{
  TmH o;
  o.mx = cc.c_mx * mx_old - cc.c_mxez * (ez_j1 - ez);
  o.bx = (bx_old + rc.c_bx1 * o.mx) - rc.c_bx0 * mx_old;
  o.my = my_old - ((-ez_i1) + ez);
  const double c_by1 = quotient_or_one(cc.num1, rc.den), c_by0 = quotient_or_one(cc.num0, rc.den);
  o.by = (rc.c_by * by_old + c_by1 * o.my) - c_by0 * my_old;
  o.hx = div_const(o.bx, v.mu0);
  o.hy = div_const(o.by, v.mu0);
  return o;
}

// Pre-pass over strip edges: one thread per (strip edge s, row r).
__global__ void tm_prepass_cols_kernel(const FusedView f)
{
  const UpmlView &v = f.u;
  const int n_rows = v.r_hi - v.r_lo + 1;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)(f.n_edges + 1) * n_rows) return;
  const int s = (int)(t / n_rows);
  const int r = v.r_lo + (int)(t - (long long)s * n_rows);
  const int c0 = v.c_lo + f.strip_w * s;                // first column of strip s
  const size_t out = (size_t)s * v.rows + r;
  if (c0 > v.c_hi + 1) return;                          // past the ragged end: nobody reads it
  const size_t k0 = (size_t)r * v.pitch + c0;
  const double2 ez0 = v.f[B200FDTD_TM_EZ][k0];
  f.col_e[out] = ez0;                                   // old Ez(r, c0) for strip s-1's lane 31
  if (s == 0) {
    // column c_lo-1 is never updated by this engine: the ring (H == 0), or the low ghost
    // column a neighbour slab's halo landed in.  The Hx array keeps it either way.
    f.col_h[out] = v.f[B200FDTD_TM_HX][k0 - 1];
  } else {
    const size_t k = k0 - 1;                            // cell (r, c0-1), owned by strip s-1
    const double2 zero = make_double2(0, 0);
    const TmH h = tm_h_cell(v, tm_col_coef(v, c0 - 1), tm_row_coef(v, r), v.f[B200FDTD_TM_EZ][k], ez0,
                            zero, v.f[B200FDTD_TM_MX][k], v.f[B200FDTD_TM_BX][k], zero, zero);
    f.col_h[out] = h.hx;                                // new Hx(r, c0-1)
  }
}

// Pre-pass over band edges: one thread per (band edge b, column c).
__global__ void tm_prepass_rows_kernel(const FusedView f)
{
  const UpmlView &v = f.u;
  const int n_cols = v.c_hi - v.c_lo + 1;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)(f.n_bands + 1) * n_cols) return;
  const int b = (int)(t / n_cols);
  const int c = v.c_lo + (int)(t - (long long)b * n_cols);
  int r0 = v.r_lo + b * f.band_h;                       // first row of band b
  if (r0 > v.r_hi + 1) r0 = v.r_hi + 1;
  const size_t out = (size_t)b * v.pitch + c;
  const size_t k0 = (size_t)r0 * v.pitch + c;
  const double2 ez0 = v.f[B200FDTD_TM_EZ][k0];
  f.row_e[out] = ez0;                                   // old Ez(r0, c) for band b-1's last row
  if (b == 0) {
    f.row_h[out] = v.f[B200FDTD_TM_HY][k0 - v.pitch];   // row r_lo-1 is never updated (ring / ghost row)
  } else {
    const size_t k = k0 - v.pitch;                      // cell (r0-1, c), owned by band b-1
    const double2 zero = make_double2(0, 0);
    const TmH h = tm_h_cell(v, tm_col_coef(v, c), tm_row_coef(v, r0 - 1), v.f[B200FDTD_TM_EZ][k], zero, ez0,
                            zero, zero, v.f[B200FDTD_TM_MY][k], v.f[B200FDTD_TM_BY][k]);
    f.row_h[out] = h.hy;                                // new Hy(r0-1, c)
  }
}

// Operands of one row of one lane, kept in registers one row ahead of the arithmetic so
// that two rows of loads (16 x 512 B per warp) are in flight while a row is computed.
struct TmRowIn {
  double2 ez_below;          // old Ez(r+1, c)
  double2 mx, bx, my, by, jz, dz;
  double eps;
  double2 edge_e, edge_h;    // lane 31: old Ez(r, c0+32); lane 0: new Hx(r, c0-1)
};

template <bool STORE_H, int WARPS, bool LOCKSTEP>
__global__ void __launch_bounds__(32 * WARPS) tm_upml_fused_kernel(const FusedView f)
{
  const UpmlView &v = f.u;
  const int lane = threadIdx.x & 31;
  const int strip = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const bool idle = strip >= f.n_strips;                // whole warp idles together
  if (idle && !LOCKSTEP) return;
  const int band = blockIdx.y;
  const int c = v.c_lo + 32 * strip + lane;
  const bool active = !idle && c <= v.c_hi;             // ragged last strip
  const bool sees_e = !idle && c <= v.c_hi + 1;         // one extra lane feeds Ez(i, j+1)
  const int r0 = v.r_lo + band * f.band_h;
  int r1 = r0 + f.band_h;
  if (r1 > v.r_hi + 1) r1 = v.r_hi + 1;

  double2 *Ez = v.f[B200FDTD_TM_EZ];
  const double2 zero = make_double2(0, 0);
  const b200fdtd_pulse pulse = fused_pulse(v);
  const double2 *col_e_next = f.col_e + (size_t)(strip + 1) * v.rows;   // old Ez(r, c0 + 32)
  const double2 *col_h_mine = f.col_h + (size_t)strip * v.rows;         // new Hx(r, c0 - 1)
  const double2 *row_e_next = f.row_e + (size_t)(band + 1) * v.pitch;   // old Ez(r1, c)

  // per-lane column coefficients stay in registers for the whole march
  TmColCoef cc = { 1.0, 1.0, 2.0, 2.0 };
  double c_dz = 1, c_dzjz = 1;
  if (active) {
    cc = tm_col_coef(v, c);
    c_dz = v.tj[B200FDTD_TMJ_C_DZ * v.pitch + c];
    c_dzjz = v.tj[B200FDTD_TMJ_C_DZJZ * v.pitch + c];
  }

  auto load_row = [&](int r, size_t k) {
    TmRowIn in;
    in.ez_below = zero; in.mx = in.bx = in.my = in.by = in.jz = in.dz = zero;
    in.eps = 1.0; in.edge_e = in.edge_h = zero;
    if (sees_e) in.ez_below = (r + 1 < r1) ? Ez[k + v.pitch] : row_e_next[c];
    if (active) {
      in.mx = v.f[B200FDTD_TM_MX][k];
      in.bx = v.f[B200FDTD_TM_BX][k];
      in.my = v.f[B200FDTD_TM_MY][k];
      in.by = v.f[B200FDTD_TM_BY][k];
      in.jz = v.f[B200FDTD_TM_JZ][k];
      in.dz = v.f[B200FDTD_TM_DZ][k];
      in.eps = v.eps0[k];
    }
    if (lane == 31 && !idle) in.edge_e = col_e_next[r];
    if (lane == 0 && !idle) in.edge_h = col_h_mine[r];
    return in;
  };

  size_t k = (size_t)r0 * v.pitch + c;
  double2 ez_cur = sees_e ? Ez[k] : zero;
  double2 hy_prev = active ? f.row_h[(size_t)band * v.pitch + c] : zero;
  TmRowIn cur = load_row(r0, k);

  for (int r = r0; r < r1; r++, k += v.pitch) {
    // keep the warps of a block on the same row so each field is touched in one
    // contiguous 32*WARPS*16-byte run at a time (DRAM page locality)
    if (LOCKSTEP) __syncthreads();
    // next row's operands first: they travel while this row is computed and stored
    TmRowIn nxt;
    if (r + 1 < r1) nxt = load_row(r + 1, k + v.pitch);
    const TmRowCoef rc = tm_row_coef(v, r);
    const double c_jz  = v.ti[B200FDTD_TMI_C_JZ * v.rows + r];
    const double c_jzh = v.ti[B200FDTD_TMI_C_JZHXHY * v.rows + r];

    double2 ez_right = shfl_down1(ez_cur);              // old Ez(r, c+1)
    if (lane == 31) ez_right = cur.edge_e;

    // ---- H phase ------------------------------------------------------------------
    TmH h;
    h.hx = zero; h.hy = zero;
    if (active) h = tm_h_cell(v, cc, rc, ez_cur, ez_right, cur.ez_below, cur.mx, cur.bx, cur.my, cur.by);
    double2 hx_left = shfl_up1(h.hx);                   // new Hx(r, c-1)
    if (lane == 0) hx_left = cur.edge_h;

    // ---- E phase (fdtdTM_upml.c:161-175) + source (field.c:248-253) -----------------
    if (active) {
      const double2 jz = c_jz * cur.jz + c_jzh * (((h.hy - hy_prev) - h.hx) + hx_left);
      const double2 dz = (c_dz * cur.dz + c_dzjz * jz) - c_dzjz * cur.jz;
      double2 ez = div_eps(dz, cur.eps);
      if (pulse.enabled && cur.eps != 1.0)
        ez = ez + pulse_term(pulse, r - 1, v.j_base + c, cur.eps);
      if ((long long)k == v.point_k)
        ez = ez + make_double2(v.point_re, v.point_im);

      v.f[B200FDTD_TM_MX][k] = h.mx;
      v.f[B200FDTD_TM_BX][k] = h.bx;
      v.f[B200FDTD_TM_MY][k] = h.my;
      v.f[B200FDTD_TM_BY][k] = h.by;
      v.f[B200FDTD_TM_JZ][k] = jz;
      v.f[B200FDTD_TM_DZ][k] = dz;
      Ez[k] = ez;
      if (STORE_H) {
        v.f[B200FDTD_TM_HX][k] = h.hx;
        v.f[B200FDTD_TM_HY][k] = h.hy;
      }
    }
    hy_prev = h.hy;
    ez_cur = cur.ez_below;
    cur = nxt;
  }
}

// ---- the same march with the operands staged through shared memory ------------------
// cp.async (LDGSTS) copies each lane's own 16-byte operands of the next STAGES-1 rows into a
// per-warp ring in shared memory, so several rows of loads are in flight per warp without
// holding them in registers (the register-prefetch form above needs 141 registers and
// leaves 12 warps per SM).  Every lane reads back only what it copied itself: no barrier,
// no cross-lane hazard; cp.async.wait_group orders a lane's own copies.
struct __align__(16) TmStage {
  double2 ez_below[32], mx[32], bx[32], my[32], by[32], jz[32], dz[32];
  double2 edge_e, edge_h;        // lane 31's old Ez(r, c0+32), lane 0's new Hx(r, c0-1)
  double eps[32];
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool STORE_H, int WARPS, int STAGES>
__global__ void __launch_bounds__(32 * WARPS) tm_upml_fused_async_kernel(const FusedView f)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const UpmlView &v = f.u;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int strip = blockIdx.x * WARPS + warp;
  if (strip >= f.n_strips) return;                      // whole warp leaves together (no block barrier used)
  TmStage *ring = reinterpret_cast<TmStage *>(smem_raw) + (size_t)warp * STAGES;
  const int band = blockIdx.y;
  const int c = v.c_lo + 32 * strip + lane;
  const bool active = c <= v.c_hi;
  const bool sees_e = c <= v.c_hi + 1;
  const int r0 = v.r_lo + band * f.band_h;
  int r1 = r0 + f.band_h;
  if (r1 > v.r_hi + 1) r1 = v.r_hi + 1;

  double2 *Ez = v.f[B200FDTD_TM_EZ];
  const double2 zero = make_double2(0, 0);
  const b200fdtd_pulse pulse = fused_pulse(v);
  const double2 *col_e_next = f.col_e + (size_t)(strip + 1) * v.rows;
  const double2 *col_h_mine = f.col_h + (size_t)strip * v.rows;
  const double2 *row_e_next = f.row_e + (size_t)(band + 1) * v.pitch;

  TmColCoef cc = { 1.0, 1.0, 2.0, 2.0 };
  double c_dz = 1, c_dzjz = 1;
  if (active) {
    cc = tm_col_coef(v, c);
    c_dz = v.tj[B200FDTD_TMJ_C_DZ * v.pitch + c];
    c_dzjz = v.tj[B200FDTD_TMJ_C_DZJZ * v.pitch + c];
  }

  // issue the copies of row r into its ring slot (each lane: its own elements)
  auto issue_row = [&](int r) {
    if (r < r1) {
      TmStage &st = ring[(r - r0) % STAGES];
      const size_t k = (size_t)r * v.pitch + c;
      if (sees_e) cp_async16(&st.ez_below[lane], (r + 1 < r1) ? &Ez[k + v.pitch] : &row_e_next[c]);
      if (active) {
        cp_async16(&st.mx[lane], &v.f[B200FDTD_TM_MX][k]);
        cp_async16(&st.bx[lane], &v.f[B200FDTD_TM_BX][k]);
        cp_async16(&st.my[lane], &v.f[B200FDTD_TM_MY][k]);
        cp_async16(&st.by[lane], &v.f[B200FDTD_TM_BY][k]);
        cp_async16(&st.jz[lane], &v.f[B200FDTD_TM_JZ][k]);
        cp_async16(&st.dz[lane], &v.f[B200FDTD_TM_DZ][k]);
        cp_async8(&st.eps[lane], &v.eps0[k]);
      }
      if (lane == 31) cp_async16(&st.edge_e, &col_e_next[r]);
      if (lane == 0) cp_async16(&st.edge_h, &col_h_mine[r]);
    }
    cp_async_commit();                                  // one group per row, empty past the band end
  };

  size_t k = (size_t)r0 * v.pitch + c;
  double2 ez_cur = sees_e ? Ez[k] : zero;
  double2 hy_prev = active ? f.row_h[(size_t)band * v.pitch + c] : zero;
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) issue_row(r0 + s);

  for (int r = r0; r < r1; r++, k += v.pitch) {
    issue_row(r + STAGES - 1);
    cp_async_wait<STAGES - 1>();                        // this lane's copies of row r have landed
    const TmStage &st = ring[(r - r0) % STAGES];
    const TmRowCoef rc = tm_row_coef(v, r);
    const double c_jz  = v.ti[B200FDTD_TMI_C_JZ * v.rows + r];
    const double c_jzh = v.ti[B200FDTD_TMI_C_JZHXHY * v.rows + r];

    const double2 ez_below = sees_e ? st.ez_below[lane] : zero;
    double2 ez_right = shfl_down1(ez_cur);
    if (lane == 31) ez_right = st.edge_e;

    TmH h;
    h.hx = zero; h.hy = zero;
    double2 jz_old = zero, dz_old = zero;
    double eps = 1.0;
    if (active) {
      h = tm_h_cell(v, cc, rc, ez_cur, ez_right, ez_below, st.mx[lane], st.bx[lane], st.my[lane], st.by[lane]);
      jz_old = st.jz[lane];
      dz_old = st.dz[lane];
      eps = st.eps[lane];
    }
    double2 hx_left = shfl_up1(h.hx);
    if (lane == 0) hx_left = st.edge_h;

    if (active) {
      const double2 jz = c_jz * jz_old + c_jzh * (((h.hy - hy_prev) - h.hx) + hx_left);
      const double2 dz = (c_dz * dz_old + c_dzjz * jz) - c_dzjz * jz_old;
      double2 ez = div_eps(dz, eps);
      if (pulse.enabled && eps != 1.0)
        ez = ez + pulse_term(pulse, r - 1, v.j_base + c, eps);
      if ((long long)k == v.point_k)
        ez = ez + make_double2(v.point_re, v.point_im);
      v.f[B200FDTD_TM_MX][k] = h.mx;
      v.f[B200FDTD_TM_BX][k] = h.bx;
      v.f[B200FDTD_TM_MY][k] = h.my;
      v.f[B200FDTD_TM_BY][k] = h.by;
      v.f[B200FDTD_TM_JZ][k] = jz;
      v.f[B200FDTD_TM_DZ][k] = dz;
      Ez[k] = ez;
      if (STORE_H) {
        v.f[B200FDTD_TM_HX][k] = h.hx;
        v.f[B200FDTD_TM_HY][k] = h.hy;
      }
    }
    hy_prev = h.hy;
    ez_cur = ez_below;
  }
  cp_async_wait<0>();
}

// ---- y-slab halos of the one-pass step (peer stores over NVLink) ----------------------------
// The two-kernel step ships my last column's new Hx upward inside the H kernel and reads the upper
// neighbour's Ez from the ghost column while no one writes it.  In one pass both need care: the
// upper neighbour must have my Hx before ITS pass starts, and once it has been told so it may
// overwrite my Ez ghost column at any time.  So a step opens with this kernel over my last owned
// column: it evaluates the x-half of that column's H phase from the old state (the expressions of
// fdtdTM_upml.c:188-189,209, i.e. the bits the pass itself will produce), stores Hx into the
// neighbour's low ghost column, and copies the old ghost Ez into a side column the pass reads
// instead of the live ghost.  b200fdtd_step orders it with the device-side flags (engine.cu).
__global__ void tm_fused_edge_kernel(const FusedView f)
{
  const UpmlView &v = f.u;
  const int r = v.r_lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (r > v.r_hi) return;
  const int c = v.c_hi;                                   // == c_last: the slab has an upper neighbour
  const size_t k = (size_t)r * v.pitch + c;
  const double2 ez = v.f[B200FDTD_TM_EZ][k], ez_ghost = v.f[B200FDTD_TM_EZ][k + 1];
  const double2 mx_old = v.f[B200FDTD_TM_MX][k], bx_old = v.f[B200FDTD_TM_BX][k];
  const double c_mx = v.tj[B200FDTD_TMJ_C_MX * v.pitch + c], c_mxez = v.tj[B200FDTD_TMJ_C_MXEZ * v.pitch + c];
  const double c_bx1 = v.ti[B200FDTD_TMI_C_BXMX1 * v.rows + r], c_bx0 = v.ti[B200FDTD_TMI_C_BXMX0 * v.rows + r];
  const double2 mx = c_mx * mx_old - c_mxez * (ez_ghost - ez);
  const double2 bx = (bx_old + c_bx1 * mx) - c_bx0 * mx_old;
  v.peer_up_h[(size_t)r * v.peer_up_pitch + (B200_JOFF - 1)] = div_const(bx, v.mu0);
  f.ghost_e[r] = ez_ghost;
}

// ---- the same march with the operands staged by TMA -----------------------------------
// Blackwell form of the one-pass step.  A CTA owns a strip of 32*WARPS columns and a band of rows.
// One producer warp streams the band through a ring of STAGES row buffers in shared memory with
// bulk asynchronous copies (cp.async.bulk, completion counted on an mbarrier per stage): per row
// the strip's segments of Ez(r+1), Mx, Bx, My, By, Jz, Dz (4 KB each at 256 columns) and eps.
// WARPS consumer warps march down the band exactly like tm_upml_fused_kernel -- a lane owns a
// column, Hy(i-1,j) is carried in registers, Hx(i,j-1) / Ez(i,j+1) come from the neighbouring lane,
// warp-edge values from the pre-pass side buffers -- but read their operands from the ring, so
// the bytes in flight (STAGES x 30 KB per SM) no longer depend on registers or occupancy: one CTA
// per SM, no spills, no block barrier; a warp hands a row buffer back with one mbarrier arrive.
// Tiles inside the frame-free rectangle use the unit-coefficient expressions (1.0 * x == x), the
// others the full ones: bit-identical to the two-kernel step either way.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int W>
struct __align__(128) TmaRow {
  double2 ez[W], mx[W], bx[W], my[W], by[W], jz[W], dz[W];
  double eps[W + 2];             // starts at an even column so the copy is 16-byte aligned
};

template <bool STORE_H, int WARPS, int STAGES, int MINB = 1>
__global__ void __launch_bounds__(32 * (WARPS + 1), MINB) tm_upml_fused_tma_kernel(const __grid_constant__ FusedView f)
{
  constexpr int W = 32 * WARPS;
  extern __shared__ __align__(128) unsigned char tma_smem[];
  TmaRow<W> *ring = reinterpret_cast<TmaRow<W> *>(tma_smem);
  double2 *ez_first = reinterpret_cast<double2 *>(tma_smem + sizeof(TmaRow<W>) * STAGES);   // Ez(r0, strip)
  unsigned long long *full = reinterpret_cast<unsigned long long *>(ez_first + W);
  unsigned long long *empty = full + STAGES;
  unsigned long long *first_bar = empty + STAGES;

  const UpmlView &v = f.u;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int band = blockIdx.y;
  const int c0 = v.c_lo + W * (int)blockIdx.x;          // first column of this CTA's strip
  const int r0 = v.r_lo + band * f.band_h;
  int r1 = r0 + f.band_h;
  if (r1 > v.r_hi + 1) r1 = v.r_hi + 1;
  int live_warps = f.n_strips - (int)blockIdx.x * WARPS; // 32-column strips that exist in this CTA
  if (live_warps > WARPS) live_warps = WARPS;
  const int eps_off = c0 & 1;
  int wcopy = v.pitch - c0;                               // never read past the row's pitch
  if (wcopy > W) wcopy = W;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], (unsigned)live_warps); }
    mbar_init(first_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double2 *Ez = v.f[B200FDTD_TM_EZ];
  if (warp == WARPS) {
    // ---- producer: one lane streams the band, STAGES rows ahead of the slowest consumer ----
    if (lane != 0) return;
    const int n_eps = (wcopy + eps_off + 1) & ~1;
    const unsigned tx = (unsigned)(7 * wcopy * sizeof(double2) + n_eps * sizeof(double));
    const double2 *row_e_next = f.row_e + (size_t)(band + 1) * v.pitch;
    // the band's first Ez row: nobody has written it yet (only this CTA's consumers will, and they
    // wait for this copy), so every consumer sees the OLD values of its own and its neighbours' cells
    mbar_expect_tx(first_bar, (unsigned)(wcopy * sizeof(double2)));
    bulk_g2s(ez_first, &Ez[(size_t)r0 * v.pitch + c0], wcopy * sizeof(double2), first_bar);
    int s = 0; unsigned ph = 0;
    for (int r = r0; r < r1; r++) {
      if (r - r0 >= STAGES) mbar_wait(&empty[s], ph ^ 1u);   // every consumer has handed the slot back
      TmaRow<W> &st = ring[s];
      const size_t k = (size_t)r * v.pitch + c0;
      mbar_expect_tx(&full[s], tx);
      bulk_g2s(st.ez, (r + 1 < r1) ? &Ez[k + v.pitch] : &row_e_next[c0], wcopy * sizeof(double2), &full[s]);
      bulk_g2s(st.mx, &v.f[B200FDTD_TM_MX][k], wcopy * sizeof(double2), &full[s]);
      bulk_g2s(st.bx, &v.f[B200FDTD_TM_BX][k], wcopy * sizeof(double2), &full[s]);
      bulk_g2s(st.my, &v.f[B200FDTD_TM_MY][k], wcopy * sizeof(double2), &full[s]);
      bulk_g2s(st.by, &v.f[B200FDTD_TM_BY][k], wcopy * sizeof(double2), &full[s]);
      bulk_g2s(st.jz, &v.f[B200FDTD_TM_JZ][k], wcopy * sizeof(double2), &full[s]);
      bulk_g2s(st.dz, &v.f[B200FDTD_TM_DZ][k], wcopy * sizeof(double2), &full[s]);
      bulk_g2s(st.eps, &v.eps0[k - eps_off], n_eps * sizeof(double), &full[s]);
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
    return;
  }

  // ---- consumers ------------------------------------------------------------------------
  if (warp >= live_warps) return;
  const int t = 32 * warp + lane;
  const int c = c0 + t;
  const bool active = c <= v.c_hi;
  const bool sees_e = c <= v.c_hi + 1;
  const double2 zero = make_double2(0, 0);
  // Values from OTHER CTAs come from the pre-pass side buffers (CTA-strip granularity); values from
  // other warps of this CTA come from the ring: old Ez of the neighbouring columns directly, the new
  // Hx(r, c-1) a warp's lane 0 needs by evaluating the left neighbour's x-half itself (same
  // expressions on the same old values as the owning lane, so the same bits).
  const bool cta_left = t == 0;                           // lane 0 of warp 0
  const bool cta_right = t == W - 1;                      // lane 31 of the last warp of a full strip
  const double2 *col_e_next = f.col_e + (size_t)(blockIdx.x + 1) * v.rows;   // old Ez(r, c0 + W)
  const double2 *col_h_mine = f.col_h + (size_t)blockIdx.x * v.rows;         // new Hx(r, c0 - 1)
  // this warp's tile inside the frame-free rectangle: every coefficient is exactly 1.0
  const bool unit = r0 >= f.unit_r_lo && r1 - 1 <= f.unit_r_hi && c0 + 32 * warp >= f.unit_c_lo &&
                    c0 + 32 * warp + 31 <= f.unit_c_hi;

  TmColCoef cc = { 1.0, 1.0, 2.0, 2.0 };
  double c_dz = 1, c_dzjz = 1;
  if (active && !unit) {
    cc = tm_col_coef(v, c);
    c_dz = v.tj[B200FDTD_TMJ_C_DZ * v.pitch + c];
    c_dzjz = v.tj[B200FDTD_TMJ_C_DZJZ * v.pitch + c];
  }
  // lane 0 of warps 1..: the x-half coefficients of column c-1 (full expressions, whatever the
  // neighbouring tile uses: with unit coefficients they give the same bits)
  const bool inner_left = lane == 0 && t > 0;
  double l_c_mx = 1, l_c_mxez = 1;
  if (inner_left) {
    l_c_mx = v.tj[B200FDTD_TMJ_C_MX * v.pitch + c - 1];
    l_c_mxez = v.tj[B200FDTD_TMJ_C_MXEZ * v.pitch + c - 1];
  }

  const b200fdtd_pulse pulse = fused_pulse(v);
  size_t k = (size_t)r0 * v.pitch + c;
  double2 hy_prev = active ? f.row_h[(size_t)band * v.pitch + c] : zero;
  double2 edge_e = zero, edge_h = zero;                   // CTA-edge lanes only, one row ahead
  if (cta_right) edge_e = col_e_next[r0];
  if (cta_left) edge_h = col_h_mine[r0];

  mbar_wait(first_bar, 0);
  double2 ez_cur = (sees_e && t < wcopy) ? ez_first[t] : zero;
  double2 ez_nb = zero;                                   // lane 31: old Ez(r, c+1); lane 0: old Ez(r, c-1)
  if (lane == 31 && !cta_right && t + 1 < wcopy) ez_nb = ez_first[t + 1];
  if (inner_left) ez_nb = ez_first[t - 1];

  int s = 0; unsigned ph = 0;
  for (int r = r0; r < r1; r++, k += v.pitch) {
    double2 edge_e_nxt = zero, edge_h_nxt = zero;
    if (r + 1 < r1) {
      if (cta_right) edge_e_nxt = col_e_next[r + 1];
      if (cta_left) edge_h_nxt = col_h_mine[r + 1];
    }
    mbar_wait(&full[s], ph);                              // this row's operands have landed
    const TmaRow<W> &st = ring[s];
    const double2 ez_below = (sees_e && t < wcopy) ? st.ez[t] : zero;
    double2 ez_nb_nxt = zero;
    if (lane == 31 && !cta_right && t + 1 < wcopy) ez_nb_nxt = st.ez[t + 1];
    double2 mx_old = zero, bx_old = zero, my_old = zero, by_old = zero, jz_old = zero, dz_old = zero;
    double2 l_mx = zero, l_bx = zero;
    double eps = 1.0;
    if (active) {
      mx_old = st.mx[t]; bx_old = st.bx[t]; my_old = st.my[t]; by_old = st.by[t];
      jz_old = st.jz[t]; dz_old = st.dz[t];
      eps = st.eps[t + eps_off];
    }
    if (inner_left) { ez_nb_nxt = st.ez[t - 1]; l_mx = st.mx[t - 1]; l_bx = st.bx[t - 1]; }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);                // the slot may be refilled
    if (++s == STAGES) { s = 0; ph ^= 1u; }

    double2 ez_right = shfl_down1(ez_cur);                // old Ez(r, c+1)
    if (lane == 31) ez_right = cta_right ? edge_e : ez_nb;
    if (f.ghost_e != nullptr && c == v.c_hi) ez_right = f.ghost_e[r];   // never the live ghost column (see the edge kernel)

    const TmRowCoef rc = tm_row_coef(v, r);               // warp-uniform, L1-resident
    TmH h;
    h.hx = zero; h.hy = zero; h.mx = h.bx = h.my = h.by = zero;
    if (active) {
      if (unit) {
        h.mx = mx_old - (ez_right - ez_cur);
        h.bx = (bx_old + h.mx) - mx_old;
        h.my = my_old - ((-ez_below) + ez_cur);
        h.by = (by_old + h.my) - my_old;
        h.hx = div_const(h.bx, v.mu0);
        h.hy = div_const(h.by, v.mu0);
      } else {
        h = tm_h_cell(v, cc, rc, ez_cur, ez_right, ez_below, mx_old, bx_old, my_old, by_old);
      }
    }
    double2 hx_left = shfl_up1(h.hx);                     // new Hx(r, c-1)
    if (cta_left) hx_left = edge_h;
    if (inner_left) {
      // fdtdTM_upml.c:188-189,209 for cell (r, c-1), as its owner evaluates them
      const double2 mx = l_c_mx * l_mx - l_c_mxez * (ez_cur - ez_nb);
      const double2 bx = (l_bx + rc.c_bx1 * mx) - rc.c_bx0 * l_mx;
      hx_left = div_const(bx, v.mu0);
    }

    if (active) {
      double2 jz, dz;
      if (unit) {
        jz = jz_old + (((h.hy - hy_prev) - h.hx) + hx_left);
        dz = (dz_old + jz) - jz_old;
      } else {
        const double c_jz  = v.ti[B200FDTD_TMI_C_JZ * v.rows + r];
        const double c_jzh = v.ti[B200FDTD_TMI_C_JZHXHY * v.rows + r];
        jz = c_jz * jz_old + c_jzh * (((h.hy - hy_prev) - h.hx) + hx_left);
        dz = (c_dz * dz_old + c_dzjz * jz) - c_dzjz * jz_old;
      }
      double2 ez = div_eps(dz, eps);
      if (pulse.enabled && eps != 1.0)
        ez = ez + pulse_term(pulse, r - 1, v.j_base + c, eps);
      if ((long long)k == v.point_k)
        ez = ez + make_double2(v.point_re, v.point_im);
      v.f[B200FDTD_TM_MX][k] = h.mx;
      v.f[B200FDTD_TM_BX][k] = h.bx;
      v.f[B200FDTD_TM_MY][k] = h.my;
      v.f[B200FDTD_TM_BY][k] = h.by;
      v.f[B200FDTD_TM_JZ][k] = jz;
      v.f[B200FDTD_TM_DZ][k] = dz;
      Ez[k] = ez;
      // y-slab halo: my first owned column of Ez is the lower neighbour's high ghost column
      if (v.peer_down_e != nullptr && c == v.c_first)
        v.peer_down_e[(size_t)r * v.peer_down_pitch + v.peer_down_col] = ez;
      if (STORE_H) {
        v.f[B200FDTD_TM_HX][k] = h.hx;
        v.f[B200FDTD_TM_HY][k] = h.hy;
      }
    }
    hy_prev = h.hy;
    ez_cur = ez_below;
    ez_nb = ez_nb_nxt;
    edge_e = edge_e_nxt;
    edge_h = edge_h_nxt;
  }
}

// H = B / mu0 over the whole plane: refreshes the H arrays when the fused kernel
// ran without storing them (the identity Hx == Bx/mu0 holds after every H phase).
// Only updated cells are touched: the ring / ghost cells of H are not derived state.
__global__ void derive_h_kernel(const double2 *__restrict__ b, double2 *h, int pitch, int r_lo, int n_rows,
                                int c_lo, int n_cols, double mu0)
{
  const size_t n = (size_t)n_rows * n_cols;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const size_t k = (size_t)(r_lo + t / n_cols) * pitch + c_lo + t % n_cols;
    h[k] = b[k] / mu0;
  }
}

}  // namespace

// The one-pass TM step (fused_kernels.cu): an unbatched double-precision slab (alone, or with peer
// halos), default pulse / point sources only; by default on grids of >= 2^22 updated cells, where
// its 232 B per cell-update beat the two kernels' 264 (on small grids a step is launch-bound and
// the pre-pass launches cost more than the bytes save).
bool b200_want_fused(const b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  if (e->g.kind != B200FDTD_TM_UPML || e->fp32 || e->n_batch > 1 || e->lean_interior) return false;
  if ((e->peer.attached[0] || e->peer.attached[1]) && !(e->fused_variant >= 20 && e->fused_variant <= 30))
    return false;                                       // only the TMA-staged form speaks the peer-halo protocol
  if (a != nullptr && (a->line.enabled || a->cw[0].enabled)) return false;
  if (e->use_fused) return true;
  if (!e->fused_auto) return false;
  return (double)(e->r_hi - e->r_lo + 1) * (double)(e->c_hi - e->c_lo + 1) >= 4194304.0;
}

int b200_fused_prepare(b200fdtd_engine *e)
{
  FusedState &fs = e->fused;
  if (e->peer.attached[1] && fs.ghost_e == nullptr) {   // (a neighbour may be attached after the first prepare)
    cudaError_t err = cudaMalloc((void **)&fs.ghost_e, (size_t)e->rows * sizeof(double2));
    if (err != cudaSuccess) return b200_fail(B200FDTD_ERR_NOMEM, "fused ghost column: %s", cudaGetErrorString(err));
    B200_CUDA(cudaMemsetAsync(fs.ghost_e, 0, (size_t)e->rows * sizeof(double2), e->stream));
    e->dev_bytes += (size_t)e->rows * sizeof(double2);
  }
  if (fs.ready) return B200FDTD_OK;
  const int n_cols = e->c_hi - e->c_lo + 1, n_rows = e->r_hi - e->r_lo + 1;
  if (n_cols < 1 || n_rows < 1) { fs.ready = true; fs.n_strips = fs.n_bands = 0; return B200FDTD_OK; }
  fs.n_strips = (n_cols + 31) / 32;
  if (fs.band_h <= 0) fs.band_h = (e->fused_variant >= 20 && e->fused_variant <= 30) ? 32 : 256;
  fs.n_bands = (n_rows + fs.band_h - 1) / fs.band_h;
  const size_t col_n = (size_t)(fs.n_strips + 1) * e->rows, row_n = (size_t)(fs.n_bands + 1) * e->pitch;
  void **ptrs[4] = { (void **)&fs.col_e, (void **)&fs.col_h, (void **)&fs.row_e, (void **)&fs.row_h };
  const size_t sizes[4] = { col_n, col_n, row_n, row_n };
  for (int n = 0; n < 4; n++) {
    cudaError_t err = cudaMalloc(ptrs[n], sizes[n] * sizeof(double2));
    if (err != cudaSuccess) return b200_fail(B200FDTD_ERR_NOMEM, "fused side buffers: %s", cudaGetErrorString(err));
    B200_CUDA(cudaMemsetAsync(*ptrs[n], 0, sizes[n] * sizeof(double2), e->stream));
    e->dev_bytes += sizes[n] * sizeof(double2);
  }
  fs.ready = true;
  return B200FDTD_OK;
}

void b200_fused_release(b200fdtd_engine *e)
{
  FusedState &fs = e->fused;
  cudaFree(fs.col_e); cudaFree(fs.col_h); cudaFree(fs.row_e); cudaFree(fs.row_h); cudaFree(fs.ghost_e);
  const int keep_band = fs.band_h;
  memset(&fs, 0, sizeof fs);
  fs.band_h = keep_band;
}

int b200_launch_upml_fused(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  if (e->fp32 || e->n_batch > 1)
    return b200_fail(B200FDTD_ERR_ARG, "the fused step serves unbatched double-precision engines");
  if (!is_tm(e->g.kind)) return b200_fail(B200FDTD_ERR_STATE, "fused step: TM only in this build");
  int rc = b200_fused_prepare(e);
  if (rc) return rc;
  FusedState &fs = e->fused;
  if (fs.n_strips == 0) return B200FDTD_OK;
  FusedView f;
  f.u = make_view(e, a);
  f.n_strips = fs.n_strips; f.n_bands = fs.n_bands; f.band_h = fs.band_h;
  const int variant = e->fused_variant;                 // tuning knob: launch shape / staging of the main kernel
  // TMA form: warps of a CTA serve each other from shared memory, only CTA strips need the pre-pass
  static const int tma_warps[] = { 8, 8, 16, 4, 8, 12, 8, 6, 5, 4, 8 };     // variants 20..30
  f.strip_w = (variant >= 20 && variant <= 30) ? 32 * tma_warps[variant - 20] : 32;
  f.n_edges = (e->c_hi - e->c_lo + 1 + f.strip_w - 1) / f.strip_w;
  f.col_e = fs.col_e; f.col_h = fs.col_h; f.row_e = fs.row_e; f.row_h = fs.row_h;
  f.unit_r_lo = e->lean_r_lo; f.unit_r_hi = e->lean_r_hi;       // empty (1, 0) when the tables have no such region
  f.unit_c_lo = e->lean_c_lo; f.unit_c_hi = e->lean_c_hi;
  f.ghost_e = fs.ghost_e;

  const int n_cols = e->c_hi - e->c_lo + 1, n_rows = e->r_hi - e->r_lo + 1;
  const long long n_col_items = (long long)(f.n_edges + 1) * n_rows;
  const long long n_row_items = (long long)(fs.n_bands + 1) * n_cols;
  tm_prepass_cols_kernel<<<(unsigned)((n_col_items + 255) / 256), 256, 0, e->stream>>>(f);
  tm_prepass_rows_kernel<<<(unsigned)((n_row_items + 255) / 256), 256, 0, e->stream>>>(f);
#define FUSED_LAUNCH(W, L)                                                                     \
  do {                                                                                         \
    dim3 grid((fs.n_strips + (W) - 1) / (W), fs.n_bands);                                      \
    if (e->store_h) tm_upml_fused_kernel<true, W, L><<<grid, 32 * (W), 0, e->stream>>>(f);      \
    else            tm_upml_fused_kernel<false, W, L><<<grid, 32 * (W), 0, e->stream>>>(f);     \
  } while (0)
#define FUSED_ASYNC_LAUNCH(W, S)                                                                \
  do {                                                                                         \
    dim3 grid((fs.n_strips + (W) - 1) / (W), fs.n_bands);                                      \
    const size_t smem = sizeof(TmStage) * (W) * (S);                                           \
    if (e->store_h) {                                                                          \
      B200_CUDA(cudaFuncSetAttribute(tm_upml_fused_async_kernel<true, W, S>,                   \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      tm_upml_fused_async_kernel<true, W, S><<<grid, 32 * (W), smem, e->stream>>>(f);          \
    } else {                                                                                   \
      B200_CUDA(cudaFuncSetAttribute(tm_upml_fused_async_kernel<false, W, S>,                  \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      tm_upml_fused_async_kernel<false, W, S><<<grid, 32 * (W), smem, e->stream>>>(f);         \
    }                                                                                          \
  } while (0)
#define FUSED_TMA_LAUNCH(W, S, MB)                                                              \
  do {                                                                                         \
    dim3 grid((fs.n_strips + (W) - 1) / (W), fs.n_bands);                                      \
    const size_t smem = sizeof(TmaRow<32 * (W)>) * (S) + sizeof(double2) * 32 * (W) +          \
                        sizeof(unsigned long long) * (2 * (S) + 1);                            \
    if (e->store_h) {                                                                          \
      B200_CUDA(cudaFuncSetAttribute(tm_upml_fused_tma_kernel<true, W, S, MB>,                 \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      tm_upml_fused_tma_kernel<true, W, S, MB><<<grid, 32 * ((W) + 1), smem, e->stream>>>(f);  \
    } else {                                                                                   \
      B200_CUDA(cudaFuncSetAttribute(tm_upml_fused_tma_kernel<false, W, S, MB>,                \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      tm_upml_fused_tma_kernel<false, W, S, MB><<<grid, 32 * ((W) + 1), smem, e->stream>>>(f); \
    }                                                                                          \
  } while (0)
  switch (variant) {
  case 20: FUSED_TMA_LAUNCH(8, 4, 1); break;
  case 21: FUSED_TMA_LAUNCH(8, 6, 1); break;
  case 22: FUSED_TMA_LAUNCH(16, 3, 1); break;
  case 23: FUSED_TMA_LAUNCH(4, 8, 1); break;
  case 24: FUSED_TMA_LAUNCH(8, 3, 1); break;
  case 25: FUSED_TMA_LAUNCH(12, 4, 1); break;
  case 26: FUSED_TMA_LAUNCH(8, 3, 2); break;      // two CTAs per SM: 16 consumer warps
  case 27: FUSED_TMA_LAUNCH(6, 4, 2); break;
  case 28: FUSED_TMA_LAUNCH(5, 3, 3); break;
  case 29: FUSED_TMA_LAUNCH(4, 3, 4); break;
  case 30: FUSED_TMA_LAUNCH(8, 2, 3); break;
  case 10: FUSED_ASYNC_LAUNCH(4, 3); break;
  case 11: FUSED_ASYNC_LAUNCH(4, 4); break;
  case 12: FUSED_ASYNC_LAUNCH(2, 4); break;
  case 13: FUSED_ASYNC_LAUNCH(8, 3); break;
  case 14: FUSED_ASYNC_LAUNCH(4, 2); break;
  case 15: FUSED_ASYNC_LAUNCH(2, 6); break;
  case 1: FUSED_LAUNCH(8, false); break;
  case 2: FUSED_LAUNCH(8, true); break;
  case 3: FUSED_LAUNCH(4, true); break;
  case 4: FUSED_LAUNCH(2, false); break;
  case 5: FUSED_LAUNCH(16, true); break;
  default: FUSED_LAUNCH(4, false); break;
  }
#undef FUSED_LAUNCH
#undef FUSED_ASYNC_LAUNCH
#undef FUSED_TMA_LAUNCH
  e->launches += 3;
  e->h_stale = !e->store_h;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}

// First kernel of a one-pass step on a slab with an upper neighbour (see tm_fused_edge_kernel).
int b200_launch_fused_edge(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  int rc = b200_fused_prepare(e);
  if (rc) return rc;
  if (e->fused.ghost_e == nullptr || e->peer.up_h == nullptr)
    return b200_fail(B200FDTD_ERR_STATE, "fused edge kernel without an upper neighbour");
  FusedView f;
  memset(&f, 0, sizeof f);
  f.u = make_view(e, a);
  f.ghost_e = e->fused.ghost_e;
  const int n_rows = e->r_hi - e->r_lo + 1;
  tm_fused_edge_kernel<<<(n_rows + 127) / 128, 128, 0, e->stream>>>(f);
  e->launches++;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}

// Bring Hx/Hy (TM) up to date from Bx/By if the fused kernel skipped storing them.
int b200_refresh_h(b200fdtd_engine *e)
{
  if (!e->h_stale) return B200FDTD_OK;
  const int n_rows = e->r_hi - e->r_lo + 1, n_cols = e->c_hi - e->c_lo + 1;
  if (e->fp32) {
    int rc;
    if (!is_tm(e->g.kind)) rc = b200_derive_h_f32(e, B200FDTD_TE_BZ, B200FDTD_TE_HZ);
    else {
      rc = b200_derive_h_f32(e, B200FDTD_TM_BX, B200FDTD_TM_HX);
      if (!rc) rc = b200_derive_h_f32(e, B200FDTD_TM_BY, B200FDTD_TM_HY);
    }
    if (!rc) e->h_stale = false;
    return rc;
  }
  for (int b = 0; b < e->n_batch; b++) {          // every simulation of a batched engine
    const size_t off = (size_t)b * e->plane;
    if (!is_tm(e->g.kind)) {
      derive_h_kernel<<<1184, 256, 0, e->stream>>>(e->field[B200FDTD_TE_BZ] + off, e->field[B200FDTD_TE_HZ] + off,
                                                   e->pitch, e->r_lo, n_rows, e->c_lo, n_cols, e->g.mu0);
      e->launches += 1;
    } else {
      derive_h_kernel<<<1184, 256, 0, e->stream>>>(e->field[B200FDTD_TM_BX] + off, e->field[B200FDTD_TM_HX] + off,
                                                   e->pitch, e->r_lo, n_rows, e->c_lo, n_cols, e->g.mu0);
      derive_h_kernel<<<1184, 256, 0, e->stream>>>(e->field[B200FDTD_TM_BY] + off, e->field[B200FDTD_TM_HY] + off,
                                                   e->pitch, e->r_lo, n_rows, e->c_lo, n_cols, e->g.mu0);
      e->launches += 2;
    }
  }
  e->h_stale = false;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}
