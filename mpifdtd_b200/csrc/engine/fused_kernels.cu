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
  double2 *col_e, *col_h;       // [n_strips + 1][rows]: old E at a strip's first column,
                                //   new H at the column just below it
  double2 *row_e, *row_h;       // [n_bands + 1][pitch]: old E at a band's first row,
                                //   new H at the row just above it
};

__device__ __forceinline__ double2 shfl_down1(double2 v)
{
  return make_double2(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1));
}
__device__ __forceinline__ double2 shfl_up1(double2 v)
{
  return make_double2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1));
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
  if (t >= (long long)(f.n_strips + 1) * n_rows) return;
  const int s = (int)(t / n_rows);
  const int r = v.r_lo + (int)(t - (long long)s * n_rows);
  const int c0 = v.c_lo + 32 * s;                       // first column of strip s
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
      if (v.pulse[0].enabled && cur.eps != 1.0)
        ez = ez + pulse_term(v.pulse[0], r - 1, v.j_base + c, cur.eps);
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
      if (v.pulse[0].enabled && eps != 1.0)
        ez = ez + pulse_term(v.pulse[0], r - 1, v.j_base + c, eps);
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

int b200_fused_prepare(b200fdtd_engine *e)
{
  FusedState &fs = e->fused;
  if (fs.ready) return B200FDTD_OK;
  const int n_cols = e->c_hi - e->c_lo + 1, n_rows = e->r_hi - e->r_lo + 1;
  if (n_cols < 1 || n_rows < 1) { fs.ready = true; fs.n_strips = fs.n_bands = 0; return B200FDTD_OK; }
  fs.n_strips = (n_cols + 31) / 32;
  if (fs.band_h <= 0) fs.band_h = 256;
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
  cudaFree(fs.col_e); cudaFree(fs.col_h); cudaFree(fs.row_e); cudaFree(fs.row_h);
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
  f.col_e = fs.col_e; f.col_h = fs.col_h; f.row_e = fs.row_e; f.row_h = fs.row_h;

  const int n_cols = e->c_hi - e->c_lo + 1, n_rows = e->r_hi - e->r_lo + 1;
  const long long n_col_items = (long long)(fs.n_strips + 1) * n_rows;
  const long long n_row_items = (long long)(fs.n_bands + 1) * n_cols;
  tm_prepass_cols_kernel<<<(unsigned)((n_col_items + 255) / 256), 256, 0, e->stream>>>(f);
  tm_prepass_rows_kernel<<<(unsigned)((n_row_items + 255) / 256), 256, 0, e->stream>>>(f);
  const int variant = e->fused_variant;                 // tuning knob: warps per block / lockstep
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
  switch (variant) {
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
  e->launches += 3;
  e->h_stale = !e->store_h;
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
