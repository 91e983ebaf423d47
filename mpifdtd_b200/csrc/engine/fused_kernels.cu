// fused_kernels.cu -- the one-pass UPML step for sm_100a: H phase and E phase in ONE kernel.
//
// The two-kernel step (upml_kernels.cu) writes B (and H) in the H phase and reads them back in
// the E phase.  Here one kernel does both, so nothing round-trips through HBM.  Per cell-update:
//   TM exact   reads Ez,Mx,Bx,My,By,Jz,Dz + eps (120 B), writes Mx,Bx,My,By,Jz,Dz,Ez (112 B) = 232 B
//              (two kernels: 264; the reference's own layout: 416)
//   TE exact   reads Ex,Ey,Mz,Bz,Jx,Dx,Jy,Dy + 2 eps (144 B), writes Mz,Bz,Jx,Dx,Jy,Dy,Ex,Ey (128 B)
//              = 272 B (two kernels: 288)
//   TM lean    reads Ez,Bx,By,Dz + eps (72 B), writes Bx,By,Dz,Ez (64 B) = 136 B   (two kernels: 168)
//   TE lean    reads Ex,Ey,Bz,Dx,Dy + 2 eps (96 B), writes Bz,Dx,Dy,Ex,Ey (80 B) = 176 B (two: 192)
//   (in vacuum row-strips -- see below -- the E arrays and eps do not move: 192 B exact, 96 B lean, TM and TE)
// "exact" evaluates the reference's expressions operation for operation (-fmad=false) and is
// bit-identical to the two-kernel step; "lean" (B200FDTD_OPT_LEAN_INTERIOR) is the tolerance form
// of upml_kernels.cu -- cells outside the absorbing frame advance B and D directly -- and is
// bit-identical to the two-kernel LEAN step.
//
// Decomposition.  E(i,j) needs the NEW H(i-1,j) and H(i,j-1) (fdtdTM_upml.c:162,
// fdtdTE_upml.c:260,270): a one-cell low-side skirt of H, and all state is updated in place, so a
// tile that recomputed its skirt would race with the tile that owns it.  Therefore:
//   * a CTA owns a strip of 32*WARPS columns and a band of rows and marches down the band: a lane
//     owns a column, B(i-1,j) of the previous row is carried in registers, B(i,j-1) and the old
//     E(i,j+1) come from the neighbouring lane by shuffle, the old E(i+1,j) is the next row's own
//     E and is loaded exactly once;
//   * ONE producer warp streams the band through a ring of STAGES row buffers in shared memory
//     with bulk asynchronous copies (cp.async.bulk = 1-D TMA; per row the strip's segments of every
//     array the tile reads, completion counted in bytes on an mbarrier per stage); WARPS consumer
//     warps wait on the row's `full` barrier, take their operands from the ring, hand the buffer
//     back with one arrive on its `empty` barrier and compute.  The bytes in flight (STAGES x ~30 KB
//     per SM) do not depend on registers or occupancy: one CTA per SM, no block barrier;
//   * values from ANOTHER WARP of the CTA come from the ring: the old E of the neighbouring
//     columns directly, the new B(i,j-1) a warp's lane 0 needs by evaluating the left neighbour's
//     H phase itself from the staged old values (same expressions on the same bits);
//   * values from ANOTHER CTA -- old E just past the strip / band, new B just before it -- are
//     produced by a small pre-pass from the OLD state into side buffers before the main kernel
//     starts (same expressions, same bits), which makes the in-place update race-free without any
//     inter-CTA ordering.  Pre-pass + side-buffer traffic is ~2 % of the step.
// Side buffers hold B, not H: an exact cell forms H = B/mu0 from it with the same correctly
// rounded division its owner uses, a lean cell uses B itself (curl H = RN(1/mu0) * curl B).
// The one exception is the first strip / first band: left of column c_lo and above row r_lo lies
// the ring or a neighbour slab's halo column, which only the H arrays hold.
#include <cstring>
#include "upml_common.cuh"

namespace {

using namespace upml;

struct OnePassView {
  UpmlView u;
  int n_strips;                 // 32-column warp strips over the updated columns
  int n_bands, band_h;
  int strip_w, n_edges;         // CTA strip width (32 * WARPS) and ceil(cols / strip_w)
  double2 *col_e, *col_b;       // [n_edges + 1][rows]: old E(j-dir) at a strip's first column; new B at the column
                                //   just below it (strip 0: the H array's value there)
  double2 *row_e, *row_b;       // [n_bands + 1][pitch]: old E(i-dir) at a band's first row; new B at the row just
                                //   above it (band 0: the H array's value there)
  int in_r_lo, in_r_hi, in_c_lo, in_c_hi;   // frame-free rectangle (all coefficients exactly 1), or empty
  int lean;                     // cells of that rectangle advance B / D directly (tolerance form)
  double2 *ghost_e;             // [rows]: y-slab with an upper neighbour: the OLD E of my high ghost column,
                                //   captured by the edge kernel before the neighbour may overwrite it; else nullptr
  const unsigned long long *vac;  // [n_bands][n_edges] row masks of the vacuum row-strips (below), or nullptr
};

// ---- vacuum row-strips: the E arrays are derived state there ------------------------------------------
// E = D/eps (+ the pulse, in material cells only: field.c:243) is formed from D every step, so in a cell
// with eps == 1 the E array holds the bits of D: x/1.0 == x.  A "vacuum row-strip" is one row of a CTA
// tile all of whose cells are updated cells with eps == 1 (and none an NTFF sample cell, so the sample
// kernel can keep reading the E arrays).  There the pass takes the old E from D -- the producer copies
// the D row where it would have copied the E row -- and does not store E nor stage eps: TM 232 -> 192 B
// per cell-update, TE 272 -> 192, lean 136 / 176 -> 96 (three complex fields in, three out).  Same
// bits: nothing is computed differently.  Whoever reads an E array outside the pass (getters, digests,
// the two-kernel forms, halo packing) calls b200_refresh_e first, which copies D over E in the
// flagged row-strips.  Only while E == D/eps actually holds (engine.h: e_consistent).
__device__ __forceinline__ bool vac_cell(const OnePassView &f, int r, int c)
{
  const UpmlView &v = f.u;
  if (f.vac == nullptr || r < v.r_lo || r > v.r_hi || c < v.c_lo || c > v.c_hi) return false;
  const int rr = r - v.r_lo;
  const unsigned long long m = f.vac[(size_t)(rr / f.band_h) * f.n_edges + (c - v.c_lo) / f.strip_w];
  return (m >> (rr % f.band_h)) & 1ull;
}
// the OLD value of an E component at (r, c): from its D array inside a vacuum row-strip
__device__ __forceinline__ double2 e_old_at(const OnePassView &f, int e_slot, int d_slot, int r, int c)
{
  const size_t k = (size_t)r * f.u.pitch + c;
  return vac_cell(f, r, c) ? f.u.f[d_slot][k] : f.u.f[e_slot][k];
}

__device__ __forceinline__ bool in_rect(const OnePassView &f, int r, int c)
{
  return r >= f.in_r_lo && r <= f.in_r_hi && c >= f.in_c_lo && c <= f.in_c_hi;
}

__device__ __forceinline__ double2 shfl_down1(double2 v)
{
  return make_double2(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1));
}
__device__ __forceinline__ double2 shfl_up1(double2 v)
{
  return make_double2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1));
}

// The pulse of source slot m as the two-kernel form sees it (pulse_of in upml_common.cuh with
// blockIdx.y == 0): the step's own parameters, or -- multi-step replay -- the engine's source
// record with time - t0 formed from the device clock.
__device__ __forceinline__ b200fdtd_pulse onepass_pulse(const UpmlView &v, int m)
{
  if (v.batch == nullptr) return v.pulse[m];
  b200fdtd_pulse p = v.batch[0].pulse[m];
  const double time = v.time_ptr != nullptr ? *v.time_ptr : v.time;
  p.time_minus_t0 = time - v.batch[0].t0[m];
  return p;
}

// ---- TM: x-half and y-half of the H phase for one cell (fdtdTM_upml.c:187-198) -------------------
// full expressions; with unit coefficients they give the bits of the unit-coefficient forms
__device__ __forceinline__ double2 tm_bx_full(const UpmlView &v, int r, int c, double2 ez, double2 ez_j1,
                                              double2 mx_old, double2 bx_old, double2 *mx_out)
{
  const double c_mx = v.tj[B200FDTD_TMJ_C_MX * v.pitch + c], c_mxez = v.tj[B200FDTD_TMJ_C_MXEZ * v.pitch + c];
  const double c_bx1 = v.ti[B200FDTD_TMI_C_BXMX1 * v.rows + r], c_bx0 = v.ti[B200FDTD_TMI_C_BXMX0 * v.rows + r];
  const double2 mx = c_mx * mx_old - c_mxez * (ez_j1 - ez);
  *mx_out = mx;
  return (bx_old + c_bx1 * mx) - c_bx0 * mx_old;
}
__device__ __forceinline__ double2 tm_by_full(const UpmlView &v, int r, int c, double2 ez, double2 ez_i1,
                                              double2 my_old, double2 by_old, double2 *my_out)
{
  const double num1 = v.tj[B200FDTD_TMJ_NUM_BYMY1 * v.pitch + c], num0 = v.tj[B200FDTD_TMJ_NUM_BYMY0 * v.pitch + c];
  const double c_by = v.ti[B200FDTD_TMI_C_BY * v.rows + r], den = v.ti[B200FDTD_TMI_DEN_BYMY * v.rows + r];
  const double2 my = my_old - ((-ez_i1) + ez);
  *my_out = my;
  const double c_by1 = quotient_or_one(num1, den), c_by0 = quotient_or_one(num0, den);
  return (c_by * by_old + c_by1 * my) - c_by0 * my_old;
}

// ---- TE: the H phase for one cell (fdtdTE_upml.c:299-301) -------------------------------------------
__device__ __forceinline__ double2 te_bz_full(const UpmlView &v, int r, int c, double2 ey_i1, double2 ey, double2 ex_j1,
                                              double2 ex, double2 mz_old, double2 bz_old, double2 *mz_out)
{
  const double c_mz = v.ti[B200FDTD_TEI_C_MZ * v.rows + r], c_mze = v.ti[B200FDTD_TEI_C_MZEXEY * v.rows + r];
  const double c_bz = v.tj[B200FDTD_TEJ_C_BZ * v.pitch + c], c_bzmz = v.tj[B200FDTD_TEJ_C_BZMZ * v.pitch + c];
  const double2 mz = c_mz * mz_old - c_mze * (((ey_i1 - ey) - ex_j1) + ex);
  *mz_out = mz;
  return (c_bz * bz_old + c_bzmz * mz) - c_bzmz * mz_old;
}

// ---- pre-pass over strip edges: one thread per (strip edge s, row r) --------------------------------
template <bool TM>
__global__ void onepass_prepass_cols_kernel(const OnePassView f)
{
  const UpmlView &v = f.u;
  const int n_rows = v.r_hi - v.r_lo + 1;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)(f.n_edges + 1) * n_rows) return;
  const int s = (int)(t / n_rows);
  const int r = v.r_lo + (int)(t - (long long)s * n_rows);
  const int c0 = v.c_lo + f.strip_w * s;                // first column of strip s
  if (c0 > v.c_hi + 1) return;                          // past the ragged end: nobody reads it
  const size_t out = (size_t)s * v.rows + r;
  const size_t k0 = (size_t)r * v.pitch + c0;
  constexpr int EJ = TM ? (int)B200FDTD_TM_EZ : (int)B200FDTD_TE_EX;           // the E component differenced along j
  constexpr int DJ = TM ? (int)B200FDTD_TM_DZ : (int)B200FDTD_TE_DX;           // ... and the D it is formed from
  // (a slab with an upper neighbour: the live ghost column may already hold the NEXT step's values)
  const double2 e0 = (f.ghost_e != nullptr && c0 == v.c_hi + 1) ? f.ghost_e[r] : e_old_at(f, EJ, DJ, r, c0);
  f.col_e[out] = e0;                                    // old E(r, c0) for strip s-1's last lane
  if (s == 0) {
    // column c_lo-1 is never updated by this engine: the ring (H == 0), or the low ghost column a
    // neighbour slab's halo landed in.  Only the H array holds it.
    f.col_b[out] = v.f[TM ? (int)B200FDTD_TM_HX : (int)B200FDTD_TE_HZ][k0 - 1];
    return;
  }
  const size_t k = k0 - 1;                              // cell (r, c0-1), owned by strip s-1
  const bool lean = f.lean && in_rect(f, r, c0 - 1);
  double2 m_unused;
  if (TM) {
    const double2 ez = e_old_at(f, EJ, DJ, r, c0 - 1), bx_old = v.f[B200FDTD_TM_BX][k];
    f.col_b[out] = lean ? bx_old - (e0 - ez)
                        : tm_bx_full(v, r, c0 - 1, ez, e0, v.f[B200FDTD_TM_MX][k], bx_old, &m_unused);
  } else {
    const double2 ey_i1 = e_old_at(f, B200FDTD_TE_EY, B200FDTD_TE_DY, r + 1, c0 - 1);
    const double2 ey = e_old_at(f, B200FDTD_TE_EY, B200FDTD_TE_DY, r, c0 - 1);
    const double2 ex = e_old_at(f, EJ, DJ, r, c0 - 1), bz_old = v.f[B200FDTD_TE_BZ][k];
    f.col_b[out] = lean ? bz_old - (((ey_i1 - ey) - e0) + ex)
                        : te_bz_full(v, r, c0 - 1, ey_i1, ey, e0, ex, v.f[B200FDTD_TE_MZ][k], bz_old, &m_unused);
  }
}

// ---- pre-pass over band edges: one thread per (band edge b, column c) -------------------------------
template <bool TM>
__global__ void onepass_prepass_rows_kernel(const OnePassView f)
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
  constexpr int EI = TM ? (int)B200FDTD_TM_EZ : (int)B200FDTD_TE_EY;           // the E component differenced along i
  constexpr int DI = TM ? (int)B200FDTD_TM_DZ : (int)B200FDTD_TE_DY;
  const double2 e0 = e_old_at(f, EI, DI, r0, c);
  f.row_e[out] = e0;                                    // old E(r0, c) for band b-1's last row
  if (b == 0) {
    f.row_b[out] = v.f[TM ? (int)B200FDTD_TM_HY : (int)B200FDTD_TE_HZ][k0 - v.pitch];   // row r_lo-1: ring / ghost row
    return;
  }
  const size_t k = k0 - v.pitch;                        // cell (r0-1, c), owned by band b-1
  const bool lean = f.lean && in_rect(f, r0 - 1, c);
  double2 m_unused;
  if (TM) {
    const double2 ez = e_old_at(f, EI, DI, r0 - 1, c), by_old = v.f[B200FDTD_TM_BY][k];
    f.row_b[out] = lean ? by_old - ((-e0) + ez)
                        : tm_by_full(v, r0 - 1, c, ez, e0, v.f[B200FDTD_TM_MY][k], by_old, &m_unused);
  } else {
    const double2 ex_j1 = (f.ghost_e != nullptr && c == v.c_hi) ? f.ghost_e[r0 - 1]
                                                                : e_old_at(f, B200FDTD_TE_EX, B200FDTD_TE_DX, r0 - 1, c + 1);
    const double2 ey = e_old_at(f, EI, DI, r0 - 1, c);
    const double2 ex = e_old_at(f, B200FDTD_TE_EX, B200FDTD_TE_DX, r0 - 1, c), bz_old = v.f[B200FDTD_TE_BZ][k];
    f.row_b[out] = lean ? bz_old - (((e0 - ey) - ex_j1) + ex)
                        : te_bz_full(v, r0 - 1, c, e0, ey, ex_j1, ex, v.f[B200FDTD_TE_MZ][k], bz_old, &m_unused);
  }
}

// ---- y-slab halos of the one-pass step (peer stores over NVLink) -------------------------------------
// The two-kernel step ships my last column's new H upward inside the H kernel and reads the upper
// neighbour's E from the ghost column while no one writes it.  In one pass both need care: the
// upper neighbour must have my H before ITS pass starts, and once it has been told so it may
// overwrite my E ghost column at any time.  So a step opens with this kernel over my last owned
// column: it evaluates that column's H phase from the old state (the bits the pass itself will
// produce), stores H into the neighbour's low ghost column, and copies the old ghost E into a side
// column that the pre-pass and the pass read instead of the live ghost.  b200fdtd_step orders it
// with the device-side flags (engine.cu).
template <bool TM>
__global__ void onepass_edge_kernel(const OnePassView f)
{
  const UpmlView &v = f.u;
  const int r = v.r_lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (r > v.r_hi) return;
  const int c = v.c_hi;                                   // == c_last: the slab has an upper neighbour
  const size_t k = (size_t)r * v.pitch + c;
  constexpr int EJ = TM ? (int)B200FDTD_TM_EZ : (int)B200FDTD_TE_EX;
  constexpr int DJ = TM ? (int)B200FDTD_TM_DZ : (int)B200FDTD_TE_DX;
  const double2 e = e_old_at(f, EJ, DJ, r, c), e_ghost = v.f[EJ][k + 1];      // the ghost column is never derived
  const bool lean = f.lean && in_rect(f, r, c);
  double2 m_unused, b;
  if (TM) {
    const double2 bx_old = v.f[B200FDTD_TM_BX][k];
    b = lean ? bx_old - (e_ghost - e) : tm_bx_full(v, r, c, e, e_ghost, v.f[B200FDTD_TM_MX][k], bx_old, &m_unused);
  } else {
    const double2 ey_i1 = e_old_at(f, B200FDTD_TE_EY, B200FDTD_TE_DY, r + 1, c);
    const double2 ey = e_old_at(f, B200FDTD_TE_EY, B200FDTD_TE_DY, r, c), bz_old = v.f[B200FDTD_TE_BZ][k];
    b = lean ? bz_old - (((ey_i1 - ey) - e_ghost) + e)
             : te_bz_full(v, r, c, ey_i1, ey, e_ghost, e, v.f[B200FDTD_TE_MZ][k], bz_old, &m_unused);
  }
  v.peer_up_h[(size_t)r * v.peer_up_pitch + (B200_JOFF - 1)] = div_const(b, v.mu0);
  f.ghost_e[r] = e_ghost;
}

// ---- mbarrier / bulk-copy primitives ------------------------------------------------------------------
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
// A consumer warp hands a ring slot back.  Its ld.shared reads of the slot are GENERIC-proxy
// accesses, the refill is an ASYNC-proxy write (cp.async.bulk), and nothing in `ld.shared ...;
// mbarrier.arrive` makes the arrive wait for the loads: ptxas issues SYNCS.ARRIVE right behind the
// last LDS without a scoreboard wait.  When this warp's arrival is the one that completes the
// phase, the producer refills at once and its first bulk copy can land in the array the warp
// loaded last while that load is still in flight.  Observed exactly so (profiles/r02_onepass_race.md):
// 16/32-byte-granular chunks of the NEXT refill's Ez in the columns of the slowest warp -- the
// frame-strip warp next to unit-coefficient warps --, 64 of 150 fresh-engine runs of the 16-warp x
// 3-buffer shape, and the one-off mismatch of round 1.  The cross-proxy fence (SASS: MEMBAR.ALL.CTA
// + FENCE.VIEW.ASYNC.S) makes every lane's loads complete first; the warp barrier orders the lanes;
// one lane arrives.  0 of 350 runs since.  (Forcing completion with a register dependency -- an
// st.shared of a fold of everything loaded -- cures it too, but measured 1-3 % slower.)
__device__ __forceinline__ void release_slot(unsigned long long *empty_bar, int lane)
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) mbar_arrive(empty_bar);
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// one row of a strip in shared memory
template <int W>
struct __align__(128) TmRow {
  double2 ez[W], bx[W], by[W], dz[W];      // every tile
  double2 mx[W], my[W], jz[W];             // not staged for tiles that lie inside the lean rectangle
  double eps[W + 2];                       // starts at an even column so the copy is 16-byte aligned
};
template <int W>
struct __align__(128) TeRow {
  double2 ey[W], ex[W], bz[W], dx[W], dy[W];
  double2 mz[W], jx[W], jy[W];
  double epx[W + 2], epy[W + 2];
};

// CTA geometry shared by the TM and TE kernels
struct Tile {
  int lane, warp, c0, r0, r1, live_warps, eps_off, wcopy, n_eps;
  bool lean_tile;
  unsigned long long vac;       // bit i: row r0 + i of this tile is a vacuum row-strip
};
template <bool LEAN, int WARPS>
__device__ __forceinline__ Tile make_tile(const OnePassView &f)
{
  constexpr int W = 32 * WARPS;
  const UpmlView &v = f.u;
  Tile t;
  t.lane = threadIdx.x & 31;
  t.warp = threadIdx.x >> 5;
  t.c0 = v.c_lo + W * (int)blockIdx.x;                  // first column of this CTA's strip
  t.r0 = v.r_lo + (int)blockIdx.y * f.band_h;
  t.r1 = t.r0 + f.band_h;
  if (t.r1 > v.r_hi + 1) t.r1 = v.r_hi + 1;
  t.live_warps = f.n_strips - (int)blockIdx.x * WARPS;  // 32-column strips that exist in this CTA
  if (t.live_warps > WARPS) t.live_warps = WARPS;
  t.eps_off = t.c0 & 1;
  t.wcopy = v.pitch - t.c0;                             // never read past the row's pitch
  if (t.wcopy > W) t.wcopy = W;
  t.n_eps = (t.wcopy + t.eps_off + 1) & ~1;
  int c_end = t.c0 + W - 1;
  if (c_end > v.c_hi) c_end = v.c_hi;
  t.lean_tile = LEAN && t.r0 >= f.in_r_lo && t.r1 - 1 <= f.in_r_hi && t.c0 >= f.in_c_lo && c_end <= f.in_c_hi;
  t.vac = f.vac != nullptr ? f.vac[(size_t)blockIdx.y * f.n_edges + blockIdx.x] : 0ull;
  return t;
}

// =========================================================================================== TM =====
// The consumer side of the TM kernel, compiled three times (MODE): the hot loops -- a warp tile inside
// the frame-free rectangle in the exact form (ROW_UNIT), a CTA tile inside it in the lean form
// (ROW_LEAN) -- carry none of the frame's table reads, quotients and per-cell case distinctions;
// everything else (frame and mixed tiles) runs ROW_GENERAL.  Same expressions, same bits.
enum { ROW_GENERAL = 0, ROW_UNIT = 1, ROW_LEAN = 2 };

// rare work of the hot loops, out of line: material cells (division by eps, the pulse's exp / sincos)
// and the opt-in point source
// true: pulse m adds a signed zero (or nothing) to cell (i, j) this step -- see pulse_add()
__device__ __forceinline__ bool pulse_is_far(const UpmlView &v, int m, int i, int j)
{
  const b200fdtd_pulse *p = &v.pulse[m];
  double tm0 = p->time_minus_t0;
  if (v.batch != nullptr) {
    p = &v.batch[0].pulse[m];
    tm0 = (v.time_ptr != nullptr ? *v.time_ptr : v.time) - v.batch[0].t0[m];
  }
  if (!p->enabled) return true;
  const double r = ((i + p->gap_x) * p->cos_per_c + (j + p->gap_y) * p->sin_per_c) - tm0;
  return fabs(r) > 28.3 * p->beam_width;
}
static __device__ __noinline__ double2 tm_pulse_cell(const UpmlView *v, int r, int c, double eps, double2 ez)
{
  const b200fdtd_pulse pulse = onepass_pulse(*v, 0);
  if (pulse.enabled) ez = pulse_add(ez, pulse, r - 1, v->j_base + c, eps);
  return ez;
}
// material cell of a hot loop: the division inline (a reciprocal and six FMAs per component pair),
// the pulse's exp / sincos out of line and only where the pulse is
// A lane marches down one column, where consecutive cells mostly share their permittivity: the
// reciprocal of the last one is kept (inv == 0: a permittivity outside the plain range, IEEE path).
struct EpsCache { double eps, inv; };
__device__ __forceinline__ double2 div_eps_cached(double2 z, double eps, EpsCache &cache)
{
  if (eps != cache.eps) {
    cache.eps = eps;
    cache.inv = eps_is_plain(eps) ? __drcp_rn(eps) : 0.0;
  }
  if (cache.inv != 0.0) return div_eps(z, eps, cache.inv);
  return make_double2(ieee_div(z.x, eps), ieee_div(z.y, eps));
}
__device__ __forceinline__ double2 tm_material_cell(const UpmlView *v, int r, int c, size_t k, double eps, double2 dz,
                                                    EpsCache &cache)
{
  double2 ez = dz;
  if (eps != 1.0) {
    ez = div_eps_cached(dz, eps, cache);
    if (!pulse_is_far(*v, 0, r - 1, v->j_base + c) || neg_zero(ez.x) || neg_zero(ez.y))
      ez = tm_pulse_cell(v, r, c, eps, ez);
  }
  if ((long long)k == v->point_k) ez = ez + make_double2(v->point_re, v->point_im);
  return ez;
}

template <int MODE, bool LEAN, bool STORE_H, int W, int STAGES>
__device__ __forceinline__ void tm_consume(const OnePassView &f, const Tile &T, TmRow<W> *ring, const double2 *ez_first,
                                           unsigned long long *full, unsigned long long *empty,
                                           unsigned long long *first_bar)
{
  constexpr bool lean_tile = MODE == ROW_LEAN;
  constexpr bool unit = MODE == ROW_UNIT;
  const UpmlView &v = f.u;
  const int lane = T.lane, warp = T.warp, c0 = T.c0, r0 = T.r0, r1 = T.r1, wcopy = T.wcopy;
  double2 *Ez = v.f[B200FDTD_TM_EZ];
  const int t = 32 * warp + lane;
  const int c = c0 + t;
  const bool active = c <= v.c_hi;
  const bool sees_e = c <= v.c_hi + 1;                    // one extra lane feeds Ez(i, j+1)
  const double2 zero = make_double2(0, 0);
  const bool cta_left = t == 0;                           // lane 0 of warp 0
  const bool cta_right = t == W - 1;                      // lane 31 of the last warp of a full strip
  const bool inner_left = lane == 0 && t > 0;             // lane 0 of the other warps
  const bool first_strip = blockIdx.x == 0, first_band = blockIdx.y == 0;
  const double2 *col_e_next = f.col_e + (size_t)(blockIdx.x + 1) * v.rows;   // old Ez(r, c0 + W)
  const double2 *col_b_mine = f.col_b + (size_t)blockIdx.x * v.rows;         // new Bx(r, c0 - 1)
  // lean form: which of this lane's / its left neighbour's cells advance B and D directly
  const bool col_in = LEAN && c >= f.in_c_lo && c <= f.in_c_hi;
  const bool left_col_in = LEAN && c - 1 >= f.in_c_lo && c - 1 <= f.in_c_hi;

  double c_dz = 1, c_dzjz = 1;
  if (active && !unit && !lean_tile) {
    c_dz = v.tj[B200FDTD_TMJ_C_DZ * v.pitch + c];
    c_dzjz = v.tj[B200FDTD_TMJ_C_DZJZ * v.pitch + c];
  }

  b200fdtd_pulse pulse;
  if (MODE == ROW_GENERAL) pulse = onepass_pulse(v, 0);
  EpsCache eps_cache = { 1.0, 1.0 };
  size_t k = (size_t)r0 * v.pitch + c;
  // new B of the previous row (by_prev) and, for exact cells, its quotient by mu0 (hy_prev)
  double2 by_prev = zero, hy_prev = zero;
  if (active) {
    const double2 x = f.row_b[(size_t)blockIdx.y * v.pitch + c];
    if (first_band) hy_prev = x;                          // the Hy array's own value at row r_lo - 1
    else { by_prev = x; if (!lean_tile) hy_prev = div_const(x, v.mu0); }
  }
  double2 edge_e = zero, edge_b = zero;                   // CTA-edge lanes only, one row ahead
  if (cta_right) edge_e = col_e_next[r0];
  if (cta_left) edge_b = col_b_mine[r0];

  mbar_wait(first_bar, 0);
  double2 ez_cur = (sees_e && t < wcopy) ? ez_first[t] : zero;
  double2 ez_nb = zero;                                   // lane 31: old Ez(r, c+1); lane 0: old Ez(r, c-1)
  if (lane == 31 && !cta_right && t + 1 < wcopy) ez_nb = ez_first[t + 1];
  if (inner_left) ez_nb = ez_first[t - 1];

  int s = 0; unsigned ph = 0;
  for (int r = r0; r < r1; r++, k += v.pitch) {
    double2 edge_e_nxt = zero, edge_b_nxt = zero;
    if (r + 1 < r1) {
      if (cta_right) edge_e_nxt = col_e_next[r + 1];
      if (cta_left) edge_b_nxt = col_b_mine[r + 1];
    }
    const bool vac_row = (T.vac >> (r - r0)) & 1ull;      // E == D in this row of the tile: E is not kept
    mbar_wait(&full[s], ph);                              // this row's operands have landed
    const TmRow<W> &st = ring[s];
    const double2 ez_below = (sees_e && t < wcopy) ? st.ez[t] : zero;
    double2 ez_nb_nxt = zero;
    if (lane == 31 && !cta_right && t + 1 < wcopy) ez_nb_nxt = st.ez[t + 1];
    double2 mx_old = zero, bx_old = zero, my_old = zero, by_old = zero, jz_old = zero, dz_old = zero;
    double2 l_mx = zero, l_bx = zero;
    double eps = 1.0;
    if (active) {
      bx_old = st.bx[t]; by_old = st.by[t]; dz_old = st.dz[t];
      if (!lean_tile) { mx_old = st.mx[t]; my_old = st.my[t]; jz_old = st.jz[t]; }
      if (!vac_row) eps = st.eps[t + T.eps_off];          // not staged for a vacuum row-strip: eps == 1
    }
    if (inner_left) {
      ez_nb_nxt = st.ez[t - 1]; l_bx = st.bx[t - 1];
      if (!lean_tile) l_mx = st.mx[t - 1];
    }
    release_slot(&empty[s], lane);                        // the slot may be refilled
    if (++s == STAGES) { s = 0; ph ^= 1u; }

    double2 ez_right = shfl_down1(ez_cur);                // old Ez(r, c+1)
    if (lane == 31) ez_right = cta_right ? edge_e : ez_nb;
    if (f.ghost_e != nullptr && c == v.c_hi) ez_right = f.ghost_e[r];   // never the live ghost column (see the edge kernel)

    const bool row_in = LEAN && r >= f.in_r_lo && r <= f.in_r_hi;
    const bool lean_cell = lean_tile || (row_in && col_in);

    // ---- H phase (fdtdTM_upml.c:187-216) -------------------------------------------------
    double2 mx = zero, bx = zero, my = zero, by = zero, hx = zero, hy = zero;
    if (active) {
      if (LEAN && lean_cell) {                            // Bx' = Bx - d(Ez)/dj, By' = By + d(Ez)/di
        bx = bx_old - (ez_right - ez_cur);
        by = by_old - ((-ez_below) + ez_cur);
      } else if (unit) {
        mx = mx_old - (ez_right - ez_cur);
        bx = (bx_old + mx) - mx_old;
        my = my_old - ((-ez_below) + ez_cur);
        by = (by_old + my) - my_old;
      } else {
        bx = tm_bx_full(v, r, c, ez_cur, ez_right, mx_old, bx_old, &mx);
        by = tm_by_full(v, r, c, ez_cur, ez_below, my_old, by_old, &my);
      }
      if (!lean_tile || STORE_H) {
        hx = div_const(bx, v.mu0);
        hy = div_const(by, v.mu0);
      }
    }
    // new Bx(r, c-1) and its quotient
    double2 bx_left = zero, hx_left = zero;
    if (LEAN) bx_left = shfl_up1(bx);
    if (!lean_tile) hx_left = shfl_up1(hx);
    if (cta_left) {
      if (first_strip) hx_left = edge_b;                  // the Hx array's own value at column c_lo - 1
      else { bx_left = edge_b; if (!lean_tile) hx_left = div_const(edge_b, v.mu0); }
    }
    if (inner_left) {
      // fdtdTM_upml.c:188-189 for cell (r, c-1), as its owner evaluates them
      double2 m_unused;
      if (LEAN && (lean_tile || (row_in && left_col_in))) bx_left = l_bx - (ez_cur - ez_nb);
      else if (unit && c - 1 >= f.in_c_lo) {              // the neighbour is a unit-coefficient cell too
        const double2 l_m = l_mx - (ez_cur - ez_nb);
        bx_left = (l_bx + l_m) - l_mx;
      }
      else bx_left = tm_bx_full(v, r, c - 1, ez_nb, ez_cur, l_mx, l_bx, &m_unused);
      if (!lean_tile) hx_left = div_const(bx_left, v.mu0);
    }

    // ---- E phase (fdtdTM_upml.c:161-175) + source (field.c:248-253) ------------------------
    if (active) {
      double2 jz = zero, dz;
      if (LEAN && lean_cell) {                            // Dz' = Dz + curl H, curl H = RN(1/mu0) * curl B
        dz = dz_old + v.mu0.r * (((by - by_prev) - bx) + bx_left);
      } else if (unit) {
        jz = jz_old + (((hy - hy_prev) - hx) + hx_left);
        dz = (dz_old + jz) - jz_old;
      } else {
        const double c_jz  = v.ti[B200FDTD_TMI_C_JZ * v.rows + r];
        const double c_jzh = v.ti[B200FDTD_TMI_C_JZHXHY * v.rows + r];
        jz = c_jz * jz_old + c_jzh * (((hy - hy_prev) - hx) + hx_left);
        dz = (c_dz * dz_old + c_dzjz * jz) - c_dzjz * jz_old;
      }
      double2 ez = dz;
      if (MODE == ROW_GENERAL) {
        ez = div_eps(dz, eps);
        if (pulse.enabled && eps != 1.0)
          ez = pulse_add(ez, pulse, r - 1, v.j_base + c, eps);
        if ((long long)k == v.point_k)
          ez = ez + make_double2(v.point_re, v.point_im);
      } else if (eps != 1.0 || (long long)k == v.point_k) {
        ez = tm_material_cell(&v, r, c, k, eps, dz, eps_cache);
      }
      if (!(LEAN && lean_cell)) {
        v.f[B200FDTD_TM_MX][k] = mx;
        v.f[B200FDTD_TM_MY][k] = my;
        v.f[B200FDTD_TM_JZ][k] = jz;
      }
      v.f[B200FDTD_TM_BX][k] = bx;
      v.f[B200FDTD_TM_BY][k] = by;
      v.f[B200FDTD_TM_DZ][k] = dz;
      if (!vac_row) Ez[k] = ez;
      // y-slab halo: my first owned column of Ez is the lower neighbour's high ghost column
      if (v.peer_down_e != nullptr && c == v.c_first)
        v.peer_down_e[(size_t)r * v.peer_down_pitch + v.peer_down_col] = ez;
      if (STORE_H) {
        v.f[B200FDTD_TM_HX][k] = hx;
        v.f[B200FDTD_TM_HY][k] = hy;
      }
    }
    by_prev = by;
    hy_prev = hy;
    ez_cur = ez_below;
    ez_nb = ez_nb_nxt;
    edge_e = edge_e_nxt;
    edge_b = edge_b_nxt;
  }
}


template <bool LEAN, bool STORE_H, int WARPS, int STAGES>
__global__ void __launch_bounds__(32 * (WARPS + 1), 1) tm_onepass_kernel(const __grid_constant__ OnePassView f)
{
  constexpr int W = 32 * WARPS;
  extern __shared__ __align__(128) unsigned char op_smem[];
  TmRow<W> *ring = reinterpret_cast<TmRow<W> *>(op_smem);
  double2 *ez_first = reinterpret_cast<double2 *>(op_smem + sizeof(TmRow<W>) * STAGES);   // Ez(r0, strip)
  unsigned long long *full = reinterpret_cast<unsigned long long *>(ez_first + W);
  unsigned long long *empty = full + STAGES;
  unsigned long long *first_bar = empty + STAGES;

  const UpmlView &v = f.u;
  const Tile T = make_tile<LEAN, WARPS>(f);
  const int lane = T.lane, warp = T.warp, c0 = T.c0, r0 = T.r0, r1 = T.r1, wcopy = T.wcopy;
  const bool lean_tile = T.lean_tile;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], (unsigned)T.live_warps); }
    mbar_init(first_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double2 *Ez = v.f[B200FDTD_TM_EZ];
  if (warp == WARPS) {
    // ---- producer: one lane streams the band, STAGES rows ahead of the slowest consumer ----
    if (lane != 0) return;
    const unsigned seg = (unsigned)(wcopy * sizeof(double2));
    const unsigned eps_bytes = (unsigned)(T.n_eps * sizeof(double));
    const unsigned tx = (lean_tile ? 4u : 7u) * seg;
    const double2 *row_e_next = f.row_e + (size_t)(blockIdx.y + 1) * v.pitch;
    const double2 *Dz = v.f[B200FDTD_TM_DZ];
    // the band's first Ez row: nobody has written it yet (only this CTA's consumers will, and they
    // wait for this copy), so every consumer sees the OLD values of its own and its neighbours' cells
    // (a vacuum row-strip: the old Ez IS the old Dz, and only Dz is kept)
    mbar_expect_tx(first_bar, seg);
    bulk_g2s(ez_first, ((T.vac & 1ull) ? Dz : Ez) + (size_t)r0 * v.pitch + c0, seg, first_bar);
    int s = 0; unsigned ph = 0;
    for (int r = r0; r < r1; r++) {
      if (r - r0 >= STAGES) mbar_wait(&empty[s], ph ^ 1u);   // every consumer has handed the slot back
      TmRow<W> &st = ring[s];
      const size_t k = (size_t)r * v.pitch + c0;
      const bool vac_row = (T.vac >> (r - r0)) & 1ull, vac_below = (T.vac >> (r + 1 - r0)) & 1ull;
      mbar_expect_tx(&full[s], tx + (vac_row ? 0u : eps_bytes));
      bulk_g2s(st.ez, (r + 1 < r1) ? (vac_below ? Dz : Ez) + k + v.pitch : &row_e_next[c0], seg, &full[s]);
      bulk_g2s(st.bx, &v.f[B200FDTD_TM_BX][k], seg, &full[s]);
      bulk_g2s(st.by, &v.f[B200FDTD_TM_BY][k], seg, &full[s]);
      bulk_g2s(st.dz, &v.f[B200FDTD_TM_DZ][k], seg, &full[s]);
      if (!lean_tile) {
        bulk_g2s(st.mx, &v.f[B200FDTD_TM_MX][k], seg, &full[s]);
        bulk_g2s(st.my, &v.f[B200FDTD_TM_MY][k], seg, &full[s]);
        bulk_g2s(st.jz, &v.f[B200FDTD_TM_JZ][k], seg, &full[s]);
      }
      if (!vac_row) bulk_g2s(st.eps, &v.eps0[k - T.eps_off], eps_bytes, &full[s]);
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
    return;
  }

  // ---- consumers ------------------------------------------------------------------------
  if (warp >= T.live_warps) return;
  if constexpr (LEAN) {
    if (lean_tile) tm_consume<ROW_LEAN, true, STORE_H, W, STAGES>(f, T, ring, ez_first, full, empty, first_bar);
    else           tm_consume<ROW_GENERAL, true, STORE_H, W, STAGES>(f, T, ring, ez_first, full, empty, first_bar);
  } else {
    // exact form: this warp's tile inside the frame-free rectangle -> unit-coefficient expressions
    const bool unit = r0 >= f.in_r_lo && r1 - 1 <= f.in_r_hi && c0 + 32 * warp >= f.in_c_lo &&
                      c0 + 32 * warp + 31 <= f.in_c_hi;
    if (unit) tm_consume<ROW_UNIT, false, STORE_H, W, STAGES>(f, T, ring, ez_first, full, empty, first_bar);
    else      tm_consume<ROW_GENERAL, false, STORE_H, W, STAGES>(f, T, ring, ez_first, full, empty, first_bar);
  }
}

// =========================================================================================== TE =====
// slots: 0 Ex 1 Jx 2 Dx 3 Ey 4 Jy 5 Dy 6 Hz 7 Mz 8 Bz.  Hz(i,j) needs Ey(i+1,j) (next row: carried like
// TM's Ez) and Ex(i,j+1) (right lane: staged per row); Ex(i,j) needs the new Hz(i,j-1) (left lane),
// Ey(i,j) the new Hz(i-1,j) (previous row, carried).  fdtdTE_upml.c:252-314.
static __device__ __noinline__ double2 te_pulse_cell(const UpmlView *v, int m, int r, int c, double eps, double2 e)
{
  const b200fdtd_pulse pulse = onepass_pulse(*v, m);
  if (pulse.enabled) e = pulse_add(e, pulse, r - 1, v->j_base + c, eps);
  return e;
}
__device__ __forceinline__ void te_material_cell(const UpmlView *v, int r, int c, size_t k, double eps_x, double eps_y,
                                                 double2 dx, double2 dy, double2 *ex_out, double2 *ey_out,
                                                 EpsCache &cache_x, EpsCache &cache_y)
{
  double2 ex = dx, ey = dy;
  if (eps_x != 1.0) {
    ex = div_eps_cached(dx, eps_x, cache_x);
    if (!pulse_is_far(*v, 0, r - 1, v->j_base + c) || neg_zero(ex.x) || neg_zero(ex.y))
      ex = te_pulse_cell(v, 0, r, c, eps_x, ex);
  }
  if (eps_y != 1.0) {
    ey = div_eps_cached(dy, eps_y, cache_y);
    if (!pulse_is_far(*v, 1, r - 1, v->j_base + c) || neg_zero(ey.x) || neg_zero(ey.y))
      ey = te_pulse_cell(v, 1, r, c, eps_y, ey);
  }
  if ((long long)k == v->point_k) ex = ex + make_double2(v->point_re, v->point_im);
  *ex_out = ex;
  *ey_out = ey;
}

// consumer side of the TE kernel, compiled per MODE like tm_consume
template <int MODE, bool LEAN, bool STORE_H, int W, int STAGES>
__device__ __forceinline__ void te_consume(const OnePassView &f, const Tile &T, TeRow<W> *ring, const double2 *ey_first,
                                           unsigned long long *full, unsigned long long *empty,
                                           unsigned long long *first_bar)
{
  constexpr bool lean_tile = MODE == ROW_LEAN;
  constexpr bool unit = MODE == ROW_UNIT;
  const UpmlView &v = f.u;
  const int lane = T.lane, warp = T.warp, c0 = T.c0, r0 = T.r0, r1 = T.r1, wcopy = T.wcopy;
  double2 *Ex = v.f[B200FDTD_TE_EX], *Ey = v.f[B200FDTD_TE_EY];
  const int t = 32 * warp + lane;
  const int c = c0 + t;
  const bool active = c <= v.c_hi;
  const bool sees_e = c <= v.c_hi + 1;
  const double2 zero = make_double2(0, 0);
  const bool cta_left = t == 0, cta_right = t == W - 1, inner_left = lane == 0 && t > 0;
  const bool first_strip = blockIdx.x == 0, first_band = blockIdx.y == 0;
  const double2 *col_e_next = f.col_e + (size_t)(blockIdx.x + 1) * v.rows;   // old Ex(r, c0 + W)
  const double2 *col_b_mine = f.col_b + (size_t)blockIdx.x * v.rows;         // new Bz(r, c0 - 1)
  const bool col_in = LEAN && c >= f.in_c_lo && c <= f.in_c_hi;
  const bool left_col_in = LEAN && c - 1 >= f.in_c_lo && c - 1 <= f.in_c_hi;

  // per-lane column coefficients of the E phase (fdtdTE_upml.c:384-403), registers for the whole march
  double c_jx = 1, c_jxhz = 1, num1 = 2, num0 = 2;
  if (active && !unit && !lean_tile) {
    c_jx = v.tj[B200FDTD_TEJ_C_JX * v.pitch + c];
    c_jxhz = v.tj[B200FDTD_TEJ_C_JXHZ * v.pitch + c];
    num1 = v.tj[B200FDTD_TEJ_NUM_DYJY1 * v.pitch + c];
    num0 = v.tj[B200FDTD_TEJ_NUM_DYJY0 * v.pitch + c];
  }

  b200fdtd_pulse pulse_x, pulse_y;
  if (MODE == ROW_GENERAL) { pulse_x = onepass_pulse(v, 0); pulse_y = onepass_pulse(v, 1); }
  EpsCache eps_cache_x = { 1.0, 1.0 }, eps_cache_y = { 1.0, 1.0 };
  size_t k = (size_t)r0 * v.pitch + c;
  double2 bz_prev = zero, hz_prev = zero;                 // new Bz(r-1, c) and its quotient by mu0
  if (active) {
    const double2 x = f.row_b[(size_t)blockIdx.y * v.pitch + c];
    if (first_band) hz_prev = x;
    else { bz_prev = x; if (!lean_tile) hz_prev = div_const(x, v.mu0); }
  }
  double2 edge_e = zero, edge_b = zero;
  if (cta_right) edge_e = col_e_next[r0];
  if (cta_left) edge_b = col_b_mine[r0];

  mbar_wait(first_bar, 0);
  double2 ey_cur = (active && t < wcopy) ? ey_first[t] : zero;
  double2 ey_nb = inner_left ? ey_first[t - 1] : zero;    // lane 0: old Ey(r, c-1)

  int s = 0; unsigned ph = 0;
  for (int r = r0; r < r1; r++, k += v.pitch) {
    double2 edge_e_nxt = zero, edge_b_nxt = zero;
    if (r + 1 < r1) {
      if (cta_right) edge_e_nxt = col_e_next[r + 1];
      if (cta_left) edge_b_nxt = col_b_mine[r + 1];
    }
    const bool vac_row = (T.vac >> (r - r0)) & 1ull;      // Ex == Dx, Ey == Dy in this row of the tile: E is not kept
    mbar_wait(&full[s], ph);
    const TeRow<W> &st = ring[s];
    double2 ey_below = zero, ex_old = zero, ex_nb = zero;
    if (active && t < wcopy) ey_below = st.ey[t];
    if (sees_e && t < wcopy) ex_old = st.ex[t];           // the extra lane past c_hi feeds Ex(i, j+1)
    if (lane == 31 && !cta_right && t + 1 < wcopy) ex_nb = st.ex[t + 1];
    double2 mz_old = zero, bz_old = zero, jx_old = zero, dx_old = zero, jy_old = zero, dy_old = zero;
    double2 l_ey_below = zero, l_ex = zero, l_mz = zero, l_bz = zero;
    double eps_x = 1.0, eps_y = 1.0;
    if (active) {
      bz_old = st.bz[t]; dx_old = st.dx[t]; dy_old = st.dy[t];
      if (!lean_tile) { mz_old = st.mz[t]; jx_old = st.jx[t]; jy_old = st.jy[t]; }
      if (!vac_row) { eps_x = st.epx[t + T.eps_off]; eps_y = st.epy[t + T.eps_off]; }
    }
    if (inner_left) {
      l_ey_below = st.ey[t - 1]; l_ex = st.ex[t - 1]; l_bz = st.bz[t - 1];
      if (!lean_tile) l_mz = st.mz[t - 1];
    }
    release_slot(&empty[s], lane);
    if (++s == STAGES) { s = 0; ph ^= 1u; }

    double2 ex_right = shfl_down1(ex_old);                // old Ex(r, c+1)
    if (lane == 31) ex_right = cta_right ? edge_e : ex_nb;
    if (f.ghost_e != nullptr && c == v.c_hi) ex_right = f.ghost_e[r];

    const bool row_in = LEAN && r >= f.in_r_lo && r <= f.in_r_hi;
    const bool lean_cell = lean_tile || (row_in && col_in);

    // ---- H phase (fdtdTE_upml.c:299-312) ------------------------------------------------
    double2 mz = zero, bz = zero, hz = zero;
    if (active) {
      if (LEAN && lean_cell) {
        bz = bz_old - (((ey_below - ey_cur) - ex_right) + ex_old);
      } else if (unit) {
        mz = mz_old - (((ey_below - ey_cur) - ex_right) + ex_old);
        bz = (bz_old + mz) - mz_old;
      } else {
        bz = te_bz_full(v, r, c, ey_below, ey_cur, ex_right, ex_old, mz_old, bz_old, &mz);
      }
      if (!lean_tile || STORE_H) hz = div_const(bz, v.mu0);
    }
    double2 bz_left = zero, hz_left = zero;               // new Bz(r, c-1) and its quotient
    if (LEAN) bz_left = shfl_up1(bz);
    if (!lean_tile) hz_left = shfl_up1(hz);
    if (cta_left) {
      if (first_strip) hz_left = edge_b;                  // the Hz array's own value at column c_lo - 1
      else { bz_left = edge_b; if (!lean_tile) hz_left = div_const(edge_b, v.mu0); }
    }
    if (inner_left) {
      double2 m_unused;
      if (LEAN && (lean_tile || (row_in && left_col_in))) bz_left = l_bz - (((l_ey_below - ey_nb) - ex_old) + l_ex);
      else if (unit && c - 1 >= f.in_c_lo) {              // the neighbour is a unit-coefficient cell too
        const double2 l_m = l_mz - (((l_ey_below - ey_nb) - ex_old) + l_ex);
        bz_left = (l_bz + l_m) - l_mz;
      }
      else bz_left = te_bz_full(v, r, c - 1, l_ey_below, ey_nb, ex_old, l_ex, l_mz, l_bz, &m_unused);
      if (!lean_tile) hz_left = div_const(bz_left, v.mu0);
    }

    // ---- E phase (fdtdTE_upml.c:259-289) + sources (fdtdTE_upml.c:186-189) ----------------------
    if (active) {
      double2 jx = zero, jy = zero, dx, dy;
      if (LEAN && lean_cell) {
        dx = dx_old + v.mu0.r * (bz - bz_left);
        dy = dy_old + v.mu0.r * ((-bz) + bz_prev);
      } else if (unit) {
        jx = jx_old + (hz - hz_left);
        dx = (dx_old + jx) - jx_old;
        jy = jy_old + ((-hz) + hz_prev);
        dy = (dy_old + jy) - jy_old;
      } else {
        const double c_dx1 = v.ti[B200FDTD_TEI_C_DXJX1 * v.rows + r], c_dx0 = v.ti[B200FDTD_TEI_C_DXJX0 * v.rows + r];
        const double c_dy = v.ti[B200FDTD_TEI_C_DY * v.rows + r], den = v.ti[B200FDTD_TEI_DEN_DYJY * v.rows + r];
        jx = c_jx * jx_old + c_jxhz * (hz - hz_left);
        dx = (dx_old + c_dx1 * jx) - c_dx0 * jx_old;
        jy = jy_old + ((-hz) + hz_prev);
        const double c_dy1 = quotient_or_one(num1, den), c_dy0 = quotient_or_one(num0, den);
        dy = (c_dy * dy_old + c_dy1 * jy) - c_dy0 * jy_old;
      }
      double2 ex = dx, ey = dy;
      if (MODE == ROW_GENERAL) {
        ex = div_eps(dx, eps_x); ey = div_eps(dy, eps_y);
        if (pulse_x.enabled && eps_x != 1.0) ex = pulse_add(ex, pulse_x, r - 1, v.j_base + c, eps_x);
        if (pulse_y.enabled && eps_y != 1.0) ey = pulse_add(ey, pulse_y, r - 1, v.j_base + c, eps_y);
        if ((long long)k == v.point_k) ex = ex + make_double2(v.point_re, v.point_im);
      } else if (eps_x != 1.0 || eps_y != 1.0 || (long long)k == v.point_k) {
        te_material_cell(&v, r, c, k, eps_x, eps_y, dx, dy, &ex, &ey, eps_cache_x, eps_cache_y);
      }
      if (!(LEAN && lean_cell)) {
        v.f[B200FDTD_TE_MZ][k] = mz;
        v.f[B200FDTD_TE_JX][k] = jx;
        v.f[B200FDTD_TE_JY][k] = jy;
      }
      v.f[B200FDTD_TE_BZ][k] = bz;
      v.f[B200FDTD_TE_DX][k] = dx;
      v.f[B200FDTD_TE_DY][k] = dy;
      if (!vac_row) {
        Ex[k] = ex;
        Ey[k] = ey;
      }
      if (v.peer_down_e != nullptr && c == v.c_first)
        v.peer_down_e[(size_t)r * v.peer_down_pitch + v.peer_down_col] = ex;
      if (STORE_H) v.f[B200FDTD_TE_HZ][k] = hz;
    }
    bz_prev = bz;
    hz_prev = hz;
    ey_cur = ey_below;
    ey_nb = l_ey_below;
    edge_e = edge_e_nxt;
    edge_b = edge_b_nxt;
  }
}

template <bool LEAN, bool STORE_H, int WARPS, int STAGES>
__global__ void __launch_bounds__(32 * (WARPS + 1), 1) te_onepass_kernel(const __grid_constant__ OnePassView f)
{
  constexpr int W = 32 * WARPS;
  extern __shared__ __align__(128) unsigned char op_smem[];
  TeRow<W> *ring = reinterpret_cast<TeRow<W> *>(op_smem);
  double2 *ey_first = reinterpret_cast<double2 *>(op_smem + sizeof(TeRow<W>) * STAGES);   // Ey(r0, strip)
  unsigned long long *full = reinterpret_cast<unsigned long long *>(ey_first + W);
  unsigned long long *empty = full + STAGES;
  unsigned long long *first_bar = empty + STAGES;

  const UpmlView &v = f.u;
  const Tile T = make_tile<LEAN, WARPS>(f);
  const int lane = T.lane, warp = T.warp, c0 = T.c0, r0 = T.r0, r1 = T.r1, wcopy = T.wcopy;
  const bool lean_tile = T.lean_tile;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], (unsigned)T.live_warps); }
    mbar_init(first_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double2 *Ex = v.f[B200FDTD_TE_EX], *Ey = v.f[B200FDTD_TE_EY];
  if (warp == WARPS) {
    if (lane != 0) return;
    const unsigned seg = (unsigned)(wcopy * sizeof(double2));
    const unsigned eps_bytes = (unsigned)(T.n_eps * sizeof(double));
    const unsigned tx = (lean_tile ? 5u : 8u) * seg;
    const double2 *row_e_next = f.row_e + (size_t)(blockIdx.y + 1) * v.pitch;
    const double2 *Dx = v.f[B200FDTD_TE_DX], *Dy = v.f[B200FDTD_TE_DY];
    mbar_expect_tx(first_bar, seg);
    bulk_g2s(ey_first, ((T.vac & 1ull) ? Dy : Ey) + (size_t)r0 * v.pitch + c0, seg, first_bar);
    int s = 0; unsigned ph = 0;
    for (int r = r0; r < r1; r++) {
      if (r - r0 >= STAGES) mbar_wait(&empty[s], ph ^ 1u);
      TeRow<W> &st = ring[s];
      const size_t k = (size_t)r * v.pitch + c0;
      // a vacuum row-strip keeps no E: the old Ex / Ey are the old Dx / Dy, and eps is not staged
      const bool vac_row = (T.vac >> (r - r0)) & 1ull, vac_below = (T.vac >> (r + 1 - r0)) & 1ull;
      mbar_expect_tx(&full[s], tx + (vac_row ? 0u : 2u * eps_bytes));
      bulk_g2s(st.ey, (r + 1 < r1) ? (vac_below ? Dy : Ey) + k + v.pitch : &row_e_next[c0], seg, &full[s]);
      bulk_g2s(st.ex, (vac_row ? Dx : Ex) + k, seg, &full[s]);
      bulk_g2s(st.bz, &v.f[B200FDTD_TE_BZ][k], seg, &full[s]);
      bulk_g2s(st.dx, &v.f[B200FDTD_TE_DX][k], seg, &full[s]);
      bulk_g2s(st.dy, &v.f[B200FDTD_TE_DY][k], seg, &full[s]);
      if (!lean_tile) {
        bulk_g2s(st.mz, &v.f[B200FDTD_TE_MZ][k], seg, &full[s]);
        bulk_g2s(st.jx, &v.f[B200FDTD_TE_JX][k], seg, &full[s]);
        bulk_g2s(st.jy, &v.f[B200FDTD_TE_JY][k], seg, &full[s]);
      }
      if (!vac_row) {
        bulk_g2s(st.epx, &v.eps0[k - T.eps_off], eps_bytes, &full[s]);
        bulk_g2s(st.epy, &v.eps1[k - T.eps_off], eps_bytes, &full[s]);
      }
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
    return;
  }

  if (warp >= T.live_warps) return;
  if constexpr (LEAN) {
    if (lean_tile) te_consume<ROW_LEAN, true, STORE_H, W, STAGES>(f, T, ring, ey_first, full, empty, first_bar);
    else           te_consume<ROW_GENERAL, true, STORE_H, W, STAGES>(f, T, ring, ey_first, full, empty, first_bar);
  } else {
    const bool unit = r0 >= f.in_r_lo && r1 - 1 <= f.in_r_hi && c0 + 32 * warp >= f.in_c_lo &&
                      c0 + 32 * warp + 31 <= f.in_c_hi;
    if (unit) te_consume<ROW_UNIT, false, STORE_H, W, STAGES>(f, T, ring, ey_first, full, empty, first_bar);
    else      te_consume<ROW_GENERAL, false, STORE_H, W, STAGES>(f, T, ring, ey_first, full, empty, first_bar);
  }
}

// H = B / mu0 over the whole plane: refreshes the H arrays when the one-pass kernel ran without
// storing them (the identity Hx == Bx/mu0 holds after every H phase).  Only updated cells are
// touched: the ring / ghost cells of H are not derived state.
__global__ void derive_h_kernel(const double2 *__restrict__ b, double2 *h, int pitch, int r_lo, int n_rows,
                                int c_lo, int n_cols, double mu0)
{
  const size_t n = (size_t)n_rows * n_cols;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const size_t k = (size_t)(r_lo + t / n_cols) * pitch + c_lo + t % n_cols;
    h[k] = b[k] / mu0;
  }
}

// ---- vacuum row-strip masks (see the top of the file) ---------------------------------------------------
// one warp per (updated row, CTA strip): the strip must be a full one of updated columns and every eps 1.0
__global__ void vac_build_kernel(unsigned long long *mask, const double *__restrict__ eps0, const double *__restrict__ eps1,
                                 int pitch, int r_lo, int n_rows, int c_lo, int c_hi, int strip_w, int n_edges, int band_h)
{
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)n_rows * n_edges) return;
  const int rr = (int)(w / n_edges), s = (int)(w - (long long)rr * n_edges);
  const int c0 = c_lo + s * strip_w;
  bool ok = c0 + strip_w - 1 <= c_hi;
  if (ok) {
    const size_t k = (size_t)(r_lo + rr) * pitch + c0;
    for (int t = lane; t < strip_w; t += 32)
      if (eps0[k + t] != 1.0 || (eps1 != nullptr && eps1[k + t] != 1.0)) ok = false;
  }
  ok = __all_sync(0xffffffffu, ok);
  if (ok && lane == 0) atomicOr(&mask[(size_t)(rr / band_h) * n_edges + s], 1ull << (rr % band_h));
}
// the rows the NTFF sample kernel reads E from keep their E arrays
__global__ void vac_clear_samples_kernel(unsigned long long *mask, const NtffPoint *__restrict__ pts, int n_local, int pitch,
                                         int r_lo, int r_hi, int c_lo, int c_hi, int strip_w, int n_edges, int band_h)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_local) return;
  const int r = (int)(pts[p].k / pitch), c = (int)(pts[p].k % pitch);
  if (r < r_lo || r > r_hi || c < c_lo || c > c_hi) return;
  const int rr = r - r_lo;
  atomicAnd(&mask[(size_t)(rr / band_h) * n_edges + (c - c_lo) / strip_w], ~(1ull << (rr % band_h)));
}
__global__ void vac_count_kernel(const unsigned long long *mask, int n, unsigned long long *rows)
{
  unsigned long long mine = 0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) mine += __popcll(mask[t]);
  if (mine) atomicAdd(rows, mine);
}
// E := D in the flagged row-strips (the identity the pass relied on)
__global__ void derive_e_kernel(const unsigned long long *__restrict__ mask, const double2 *__restrict__ d, double2 *e,
                                int pitch, int r_lo, int n_rows, int c_lo, int strip_w, int n_edges, int band_h)
{
  const size_t n = (size_t)n_rows * n_edges * strip_w;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const int rr = (int)(t / ((size_t)n_edges * strip_w));
    const int cc = (int)(t - (size_t)rr * n_edges * strip_w);
    if (!((mask[(size_t)(rr / band_h) * n_edges + cc / strip_w] >> (rr % band_h)) & 1ull)) continue;
    const size_t k = (size_t)(r_lo + rr) * pitch + c_lo + cc;       // a flagged strip is a full one: in range
    e[k] = d[k];
  }
}

// launch shapes: consumer warps x row buffers
struct Shape { int warps, stages; };
__host__ Shape shape_of(int variant)
{
  switch (variant) {
  case 21: return { 8, 6 };
  case 22: return { 16, 3 };
  case 23: return { 4, 8 };
  case 24: return { 8, 3 };
  default: return { 8, 4 };       // 20
  }
}

template <bool TM, bool LEAN, bool STORE_H, int WARPS, int STAGES>
cudaError_t launch_main(const OnePassView &f, dim3 grid, cudaStream_t stream)
{
  constexpr int W = 32 * WARPS;
  constexpr size_t row = TM ? sizeof(TmRow<W>) : sizeof(TeRow<W>);
  constexpr size_t smem = row * STAGES + sizeof(double2) * W + sizeof(unsigned long long) * (2 * STAGES + 1);
  if constexpr (smem > 232448) {
    return cudaErrorInvalidConfiguration;
  } else {
    cudaError_t err;
    if constexpr (TM) {
      err = cudaFuncSetAttribute(tm_onepass_kernel<LEAN, STORE_H, WARPS, STAGES>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err != cudaSuccess) return err;
      tm_onepass_kernel<LEAN, STORE_H, WARPS, STAGES><<<grid, 32 * (WARPS + 1), smem, stream>>>(f);
    } else {
      err = cudaFuncSetAttribute(te_onepass_kernel<LEAN, STORE_H, WARPS, STAGES>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err != cudaSuccess) return err;
      te_onepass_kernel<LEAN, STORE_H, WARPS, STAGES><<<grid, 32 * (WARPS + 1), smem, stream>>>(f);
    }
    return cudaGetLastError();
  }
}

template <bool TM, int WARPS, int STAGES>
cudaError_t launch_shape(const OnePassView &f, dim3 grid, cudaStream_t stream, bool lean, bool store_h)
{
  if (lean) return store_h ? launch_main<TM, true, true, WARPS, STAGES>(f, grid, stream)
                           : launch_main<TM, true, false, WARPS, STAGES>(f, grid, stream);
  return store_h ? launch_main<TM, false, true, WARPS, STAGES>(f, grid, stream)
                 : launch_main<TM, false, false, WARPS, STAGES>(f, grid, stream);
}

template <bool TM>
cudaError_t launch_variant(int variant, const OnePassView &f, dim3 grid, cudaStream_t stream, bool lean, bool store_h)
{
  switch (variant) {
  case 21: return launch_shape<TM, 8, 6>(f, grid, stream, lean, store_h);
  case 22: return launch_shape<TM, 16, 3>(f, grid, stream, lean, store_h);
  case 23: return launch_shape<TM, 4, 8>(f, grid, stream, lean, store_h);
  case 24: return launch_shape<TM, 8, 3>(f, grid, stream, lean, store_h);
  default: return launch_shape<TM, 8, 4>(f, grid, stream, lean, store_h);
  }
}

bool lean_form(const b200fdtd_engine *e)
{
  return e->lean_interior && e->lean_r_hi >= e->lean_r_lo && e->lean_c_hi >= e->lean_c_lo;
}

// may this step treat the E arrays as derived state in the vacuum row-strips?
bool use_vac(const b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  const FusedState &fs = e->fused;
  return e->derived_e && e->e_consistent && fs.vac_built && fs.vac != nullptr && fs.vac_cells > 0 &&
         !(a != nullptr && a->point.enabled);
}

void fill_view(const b200fdtd_engine *e, const b200fdtd_step_args *a, OnePassView &f)
{
  const FusedState &fs = e->fused;
  memset(&f, 0, sizeof f);
  f.u = make_view(e, a);
  f.n_strips = fs.n_strips; f.n_bands = fs.n_bands; f.band_h = fs.band_h;
  f.strip_w = 32 * shape_of(e->fused_variant).warps;
  f.n_edges = (e->c_hi - e->c_lo + 1 + f.strip_w - 1) / f.strip_w;
  f.col_e = fs.col_e; f.col_b = fs.col_h; f.row_e = fs.row_e; f.row_b = fs.row_h;
  f.in_r_lo = e->lean_r_lo; f.in_r_hi = e->lean_r_hi;       // empty (1, 0) when the tables have no such region
  f.in_c_lo = e->lean_c_lo; f.in_c_hi = e->lean_c_hi;
  f.lean = lean_form(e) ? 1 : 0;
  f.ghost_e = fs.ghost_e;
  f.vac = use_vac(e, a) ? fs.vac : nullptr;
}

}  // namespace

// The one-pass step: an unbatched double-precision slab (alone, or with peer halos) of the serial
// UPML kinds (2 TM, 3 TE), default pulse / point sources only; by default on grids of >= 2^22
// updated cells (on small grids a step is launch-bound and the pre-pass launches cost more than
// the bytes save).
bool b200_want_fused(const b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  if ((e->g.kind != B200FDTD_TM_UPML && e->g.kind != B200FDTD_TE_UPML) || e->fp32 || e->n_batch > 1) return false;
  if (a != nullptr && (a->line.enabled || a->cw[0].enabled || a->cw[1].enabled)) return false;
  if (e->use_fused) return true;
  if (!e->fused_auto) return false;
  return (double)(e->r_hi - e->r_lo + 1) * (double)(e->c_hi - e->c_lo + 1) >= 4194304.0;
}

// Row masks of the vacuum row-strips for the current launch shape, eps maps and NTFF plan.  Not while a
// stream capture is running (b200fdtd_run_steps prepares before it captures): the keys cannot change there.
static int build_vac(b200fdtd_engine *e)
{
  FusedState &fs = e->fused;
  const int strip_w = 32 * shape_of(e->fused_variant).warps;
  const int n_cols = e->c_hi - e->c_lo + 1, n_rows = e->r_hi - e->r_lo + 1;
  const int n_edges = (n_cols + strip_w - 1) / strip_w;
  if (fs.vac_built && fs.vac_strip_w == strip_w && fs.vac_band_h == fs.band_h && fs.vac_edges == n_edges &&
      fs.vac_eps_epoch == e->eps_epoch && fs.vac_ntff_epoch == e->ntff_epoch)
    return B200FDTD_OK;
  cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(e->stream, &capturing);
  if (capturing != cudaStreamCaptureStatusNone) return B200FDTD_OK;      // keep what there is
  // whatever the old masks let the pass skip is brought up to date first
  int rc = b200_refresh_e(e); if (rc) return rc;
  cudaFree(fs.vac); fs.vac = nullptr;
  fs.vac_built = true;
  fs.vac_strip_w = strip_w; fs.vac_band_h = fs.band_h; fs.vac_edges = n_edges;
  fs.vac_eps_epoch = e->eps_epoch; fs.vac_ntff_epoch = e->ntff_epoch;
  fs.vac_cells = 0;
  const bool tm = is_tm(e->g.kind);
  if (!e->derived_e || fs.band_h > 63 || n_edges < 1 || fs.n_bands < 1 || !e->have_eps[0] || (!tm && !e->have_eps[1]))
    return B200FDTD_OK;
  const size_t n_masks = (size_t)fs.n_bands * n_edges;
  unsigned long long *count = nullptr;
  cudaError_t err = cudaMalloc((void **)&fs.vac, (n_masks + 1) * sizeof(unsigned long long));
  if (err != cudaSuccess) return b200_fail(B200FDTD_ERR_NOMEM, "one-pass row masks: %s", cudaGetErrorString(err));
  count = fs.vac + n_masks;
  B200_CUDA(cudaMemsetAsync(fs.vac, 0, (n_masks + 1) * sizeof(unsigned long long), e->stream));
  const long long n_warps = (long long)n_rows * n_edges;
  vac_build_kernel<<<(unsigned)((n_warps * 32 + 255) / 256), 256, 0, e->stream>>>(
      fs.vac, e->eps[0], tm ? nullptr : e->eps[1], e->pitch, e->r_lo, n_rows, e->c_lo, e->c_hi, strip_w, n_edges, fs.band_h);
  if (e->ntff.ready && e->ntff.n_local > 0)
    vac_clear_samples_kernel<<<(e->ntff.n_local + 127) / 128, 128, 0, e->stream>>>(
        fs.vac, e->ntff.pts, e->ntff.n_local, e->pitch, e->r_lo, e->r_hi, e->c_lo, e->c_hi, strip_w, n_edges, fs.band_h);
  vac_count_kernel<<<148, 256, 0, e->stream>>>(fs.vac, (int)n_masks, count);
  e->launches += 3;
  unsigned long long rows = 0;
  B200_CUDA(cudaMemcpyAsync(&rows, count, sizeof rows, cudaMemcpyDeviceToHost, e->stream));
  B200_CUDA(cudaStreamSynchronize(e->stream));
  fs.vac_cells = rows * (unsigned long long)strip_w;
  e->graph_epoch++;
  return B200FDTD_OK;
}

int b200_fused_prepare(b200fdtd_engine *e)
{
  FusedState &fs = e->fused;
  if (e->peer.attached[1] && fs.ghost_e == nullptr) {   // (a neighbour may be attached after the first prepare)
    cudaError_t err = cudaMalloc((void **)&fs.ghost_e, (size_t)e->rows * sizeof(double2));
    if (err != cudaSuccess) return b200_fail(B200FDTD_ERR_NOMEM, "one-pass ghost column: %s", cudaGetErrorString(err));
    B200_CUDA(cudaMemsetAsync(fs.ghost_e, 0, (size_t)e->rows * sizeof(double2), e->stream));
    e->dev_bytes += (size_t)e->rows * sizeof(double2);
  }
  if (fs.ready) return build_vac(e);
  const int n_cols = e->c_hi - e->c_lo + 1, n_rows = e->r_hi - e->r_lo + 1;
  if (n_cols < 1 || n_rows < 1) { fs.ready = true; fs.n_strips = fs.n_bands = 0; return B200FDTD_OK; }
  fs.n_strips = (n_cols + 31) / 32;
  if (fs.band_h <= 0) fs.band_h = 32;
  fs.n_bands = (n_rows + fs.band_h - 1) / fs.band_h;
  const size_t col_n = (size_t)(fs.n_strips + 1) * e->rows, row_n = (size_t)(fs.n_bands + 1) * e->pitch;
  void **ptrs[4] = { (void **)&fs.col_e, (void **)&fs.col_h, (void **)&fs.row_e, (void **)&fs.row_h };
  const size_t sizes[4] = { col_n, col_n, row_n, row_n };
  for (int n = 0; n < 4; n++) {
    cudaError_t err = cudaMalloc(ptrs[n], sizes[n] * sizeof(double2));
    if (err != cudaSuccess) return b200_fail(B200FDTD_ERR_NOMEM, "one-pass side buffers: %s", cudaGetErrorString(err));
    B200_CUDA(cudaMemsetAsync(*ptrs[n], 0, sizes[n] * sizeof(double2), e->stream));
    e->dev_bytes += sizes[n] * sizeof(double2);
  }
  fs.ready = true;
  return build_vac(e);
}

// Does a step of this engine (pulse sources) leave the E arrays behind D in the vacuum row-strips?  For callers
// that replay captured steps, where the launch functions -- which keep e_stale -- do not run.
bool b200_fused_derives_e(const b200fdtd_engine *e)
{
  return b200_want_fused(e, nullptr) && use_vac(e, nullptr);
}

int b200fdtd_onepass_vacuum_cells(b200fdtd_engine *e, uint64_t *cells)
{
  if (!e || !cells) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  *cells = 0;
  if (!b200_want_fused(e, nullptr)) return B200FDTD_OK;
  cudaSetDevice(e->device);
  int rc = b200_fused_prepare(e); if (rc) return rc;
  if (e->derived_e && e->fused.vac != nullptr) *cells = e->fused.vac_cells;
  return B200FDTD_OK;
}

void b200_fused_release(b200fdtd_engine *e)
{
  FusedState &fs = e->fused;
  b200_refresh_e(e);            // the masks go away: nothing may stay derived
  cudaFree(fs.col_e); cudaFree(fs.col_h); cudaFree(fs.row_e); cudaFree(fs.row_h); cudaFree(fs.ghost_e); cudaFree(fs.vac);
  const int keep_band = fs.band_h;
  memset(&fs, 0, sizeof fs);
  fs.band_h = keep_band;
}

int b200_launch_upml_fused(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  if (e->fp32 || e->n_batch > 1)
    return b200_fail(B200FDTD_ERR_ARG, "the one-pass step serves unbatched double-precision engines");
  if (e->g.kind != B200FDTD_TM_UPML && e->g.kind != B200FDTD_TE_UPML)
    return b200_fail(B200FDTD_ERR_STATE, "the one-pass step serves the serial UPML kinds (2, 3)");
  int rc = b200_fused_prepare(e);
  if (rc) return rc;
  FusedState &fs = e->fused;
  if (fs.n_strips == 0) return B200FDTD_OK;
  const bool tm = is_tm(e->g.kind);
  if (!use_vac(e, a)) { rc = b200_refresh_e(e); if (rc) return rc; }    // this pass reads the E arrays everywhere
  OnePassView f;
  fill_view(e, a, f);
  const Shape sh = shape_of(e->fused_variant);

  const int n_cols = e->c_hi - e->c_lo + 1, n_rows = e->r_hi - e->r_lo + 1;
  const long long n_col_items = (long long)(f.n_edges + 1) * n_rows;
  const long long n_row_items = (long long)(fs.n_bands + 1) * n_cols;
  if (tm) {
    onepass_prepass_cols_kernel<true><<<(unsigned)((n_col_items + 255) / 256), 256, 0, e->stream>>>(f);
    onepass_prepass_rows_kernel<true><<<(unsigned)((n_row_items + 255) / 256), 256, 0, e->stream>>>(f);
  } else {
    onepass_prepass_cols_kernel<false><<<(unsigned)((n_col_items + 255) / 256), 256, 0, e->stream>>>(f);
    onepass_prepass_rows_kernel<false><<<(unsigned)((n_row_items + 255) / 256), 256, 0, e->stream>>>(f);
  }
  const dim3 grid((fs.n_strips + sh.warps - 1) / sh.warps, fs.n_bands);
  // the default shape keeps 4 row buffers in flight; the TM lean form stages 18 KB per row instead of 30 and
  // runs ~1.5 % faster with 6 (same strip width, so the pre-pass geometry is unchanged)
  const int variant = (e->fused_variant == 20 && tm && f.lean) ? 21 : e->fused_variant;
  const cudaError_t err = tm ? launch_variant<true>(variant, f, grid, e->stream, f.lean != 0, e->store_h)
                             : launch_variant<false>(variant, f, grid, e->stream, f.lean != 0, e->store_h);
  if (err != cudaSuccess)
    return b200_fail(B200FDTD_ERR_CUDA, "one-pass kernel (shape %d): %s", e->fused_variant, cudaGetErrorString(err));
  e->launches += 3;
  e->h_stale = !e->store_h;
  if (f.vac != nullptr) e->e_stale = true;
  // E == D/eps (+ pulse) holds again in every updated cell, unless the point source touched one
  e->e_consistent = !a->point.enabled;
  return B200FDTD_OK;
}

// First kernel of a one-pass step on a slab with an upper neighbour (see onepass_edge_kernel).
int b200_launch_fused_edge(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  int rc = b200_fused_prepare(e);
  if (rc) return rc;
  if (e->fused.ghost_e == nullptr || e->peer.up_h == nullptr)
    return b200_fail(B200FDTD_ERR_STATE, "one-pass edge kernel without an upper neighbour");
  if (!use_vac(e, a)) { rc = b200_refresh_e(e); if (rc) return rc; }    // this step reads the E arrays everywhere
  OnePassView f;
  fill_view(e, a, f);
  const int n_rows = e->r_hi - e->r_lo + 1;
  if (is_tm(e->g.kind)) onepass_edge_kernel<true><<<(n_rows + 127) / 128, 128, 0, e->stream>>>(f);
  else                  onepass_edge_kernel<false><<<(n_rows + 127) / 128, 128, 0, e->stream>>>(f);
  e->launches++;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}

// Bring the E arrays up to date in the vacuum row-strips, where the one-pass step does not store them.
int b200_refresh_e(b200fdtd_engine *e)
{
  if (!e->e_stale) return B200FDTD_OK;
  const FusedState &fs = e->fused;
  if (fs.vac == nullptr) { e->e_stale = false; return B200FDTD_OK; }
  const int n_rows = e->r_hi - e->r_lo + 1;
  const int pairs[3][2] = { { B200FDTD_TM_DZ, B200FDTD_TM_EZ }, { B200FDTD_TE_DX, B200FDTD_TE_EX }, { B200FDTD_TE_DY, B200FDTD_TE_EY } };
  for (int n = is_tm(e->g.kind) ? 0 : 1; n < (is_tm(e->g.kind) ? 1 : 3); n++) {
    derive_e_kernel<<<1184, 256, 0, e->stream>>>(fs.vac, e->field[pairs[n][0]], e->field[pairs[n][1]], e->pitch, e->r_lo,
                                                 n_rows, e->c_lo, fs.vac_strip_w, fs.vac_edges, fs.vac_band_h);
    e->launches++;
  }
  e->e_stale = false;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}

// Bring Hx/Hy (TM) / Hz (TE) up to date from B if the last step did not store them.
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
