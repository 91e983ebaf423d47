// split_kernels.cu -- split-field solvers for sm_100a: plain Yee with Berenger PML
// (solver ids 0/1) and non-standard FDTD (ids 6/7).
//
// What they replace (rennone/mpiFDTD):
//   TM     calcH, calcE                  fdtdTM.c:299-327      + field_scatteredWaveNotUPML field.c:179-196
//   TE     calcE, calcH                  fdtdTE.c:290-321      + the same source on Ey
//   NS TM  calcH (9-point NS operator),  nsFdtdTM.c:91-151     + field_nsScatteredWaveNotUPML field.c:155-177
//          calcE, Ez = Ezx + Ezy         nsFdtdTM.c:68-80
//   NS TE  calcH, Hz = Hzx + Hzy, calcE  nsFdtdTE.c:233-308    + the NS source on Ex and Ey
// Five complex fields; eight per-cell coefficients that depend on the permittivity (dense
// arrays built on the host with the reference's expressions, bit-identical) plus the
// per-cell source factor.  Each solver step is two streaming kernels, one thread per
// cell, 128-bit accesses; arithmetic keeps the reference's operand order (-fmad=false).
#include "upml_common.cuh"

// Resident blocks per SM the kernels are compiled for (register cap 65536 / (256 * n)).  A/B at 4096^2
// (scripts/gpu_jobs/r02_j22.sh): the TE pair and the NS-FDTD TM pair gain 4-7 % at 6 blocks (id 1 27.7 ->
// 29.6, id 6 25.7 -> 26.7 Gcell-updates/s, id 7 unchanged), the Berenger TM pair loses 5 %; 8 blocks spill.
#define B200_SPLIT_TE_MIN_BLOCKS 6
#define B200_SPLIT_TM_MIN_BLOCKS(NS) ((NS) ? 6 : 1)

namespace {

using namespace upml;

struct SplitView {
  double2 *f[5];
  const double *c[B200FDTD_MAX_DENSE];
  const double *eps0, *eps1;            // lean form, kinds 0 / 1
  const double *ti, *tj;                // lean form: 1-D tables [B200FDTD_SPLIT_TABS][rows | pitch]
  int rows;
  int pitch;
  int r_lo, c_lo, c_hi, nbx;
  int in_r_lo, in_r_hi, in_c_lo, in_c_hi;   // kind 7: the rectangle of b200fdtd_set_split_interior (empty: lo > hi)
  int j_base;
  b200fdtd_cw cw[2];
  double ns_r2;
  size_t plane;                         // batched engines: elements per simulation (blockIdx.y selects it)
  const b200fdtd_batch_cw *batch;       // batched engines: per-simulation part of the CW source
  const b200fdtd_cw *cw_step;           // replayed chunk (b200fdtd_run_split_steps): this step's two records, else nullptr
};

__device__ __forceinline__ bool locate(const SplitView &v, int &r, int &c, size_t &k)
{
  const long long b = blockIdx.x;
  const int rb = (int)(b / v.nbx);
  const int cb = (int)(b - (long long)rb * v.nbx);
  r = v.r_lo + rb;
  c = v.c_lo + cb * kBlock + (int)threadIdx.x;
  k = (size_t)r * (size_t)v.pitch + (size_t)c;
  return c <= v.c_hi;
}

// kind 7: does this whole thread block (kBlock consecutive columns of row r) lie inside the rectangle
// where the decay coefficients are exactly 1.0 and the two curl coefficients coincide?  Block-uniform.
__device__ __forceinline__ bool block_in_interior(const SplitView &v, int r, int c)
{
  const int c_first = c - (int)threadIdx.x;
  return r >= v.in_r_lo && r <= v.in_r_hi && c_first >= v.in_c_lo && c_first + kBlock - 1 <= v.in_c_hi;
}

// one CW source target; `factor` is the host-built per-cell (eps0/eps - 1)-type term
__device__ __forceinline__ double2 cw_term(const b200fdtd_cw &s, int i, int j, double factor)
{
  const double kr = (i + s.gap_x) * s.ks_cos + (j + s.gap_y) * s.ks_sin;
  double sa, ca;
  sincos(kr - s.phase_a, &sa, &ca);
  double2 wave = make_double2(ca, sa);
  if (s.two_term) {
    double sb, cb;
    sincos(kr - s.phase_b, &sb, &cb);
    wave = make_double2(ca - cb, sa - sb);
  }
  const double amp = s.scale * factor;
  return make_double2(amp * wave.x, amp * wave.y);
}

// The CW record of this thread block's simulation.  Batched engines (BATCH): the angle-dependent
// members come from the per-simulation record, scale = ray_coef * dot as the host forms it.
template <bool BATCH>
__device__ __forceinline__ bool cw_on(const SplitView &v, int m)
{
  if (BATCH) return v.batch[blockIdx.y].enabled[m] != 0;
  return (v.cw_step != nullptr ? v.cw_step[m].enabled : v.cw[m].enabled) != 0;
}
template <bool BATCH>
__device__ __forceinline__ b200fdtd_cw cw_of(const SplitView &v, int m)
{
  b200fdtd_cw s = v.cw_step != nullptr ? v.cw_step[m] : v.cw[m];
  if (BATCH) {
    const b200fdtd_batch_cw b = v.batch[blockIdx.y];
    s.ks_cos = b.ks_cos;
    s.ks_sin = b.ks_sin;
    s.scale = s.scale * b.dot[m];
  }
  return s;
}

__device__ __forceinline__ double2 twice(double2 z) { return make_double2(2 * z.x, 2 * z.y); }

// ---- lean form: coefficients evaluated in the kernel with the reference's operations -------
// field_pmlCoef(eps, sig) = (1.0 - sig/eps)/(1.0 + sig/eps) and field_pmlCoef_LXY(eps, sig) =
// 1.0/(eps + sig) (field.c:288-295).  Outside the PML sig == 0 and they are exactly 1.0 and
// 1.0/eps (inv_eps, shared by both directions and by the source factor eps0/eps - 1); inside,
// three IEEE divisions -- rows (sig_x) are block-uniform, columns (sig_y) diverge only at the
// row ends.
struct PmlPair { double c, l; };
__device__ __forceinline__ PmlPair pml_pair(double eps, double sig, double inv_eps)
{
  PmlPair p;
  p.c = 1.0;
  p.l = inv_eps;
  if (sig != 0.0) {
    const double q = ieee_div(sig, eps);
    p.c = ieee_div(1.0 - q, 1.0 + q);
    p.l = ieee_div(1.0, eps + sig);
  }
  return p;
}
__device__ __forceinline__ double one_over(double eps) { return eps == 1.0 ? 1.0 : ieee_div(1.0, eps); }
// g / den where den == 1.0 outside the PML (NS-FDTD: (u*z) / (1 + beta), nsFdtdTM.c:285-305)
__device__ __forceinline__ double over_den(double g, double den) { return den == 1.0 ? g : ieee_div(g, den); }

// ---------------------------------------------------------------- TM family ------
#define F(SLOT) (v.f[SLOT] + sim)

// slots: 0 Ez 1 Ezx 2 Ezy 3 Hx 4 Hy
// LEAN: see b200fdtd.h "lean form".  Kind 0 (Berenger): the four H coefficients are 1-D tables.
// Kind 6 (NS): decay coefficients 1-D, curl coefficients G[k] / DEN (1-D).
template <bool NS, bool LEAN, bool BATCH>
__global__ void __launch_bounds__(kBlock, B200_SPLIT_TM_MIN_BLOCKS(NS)) split_tm_h_kernel(const SplitView v)
{
  int r, c; size_t k;
  if (!locate(v, r, c, k)) return;
  const size_t sim = BATCH ? (size_t)blockIdx.y * v.plane : 0;   // field arrays only: coefficients are shared
  const int P = v.pitch;
  double2 dy_term, dx_term;
  if (NS) {                                             // nsFdtdTM.c:127-150
    const double2 *__restrict__ Ez = F(B200FDTD_STM_EZ);
    const double2 e = Ez[k], e_j = Ez[k + 1], e_i = Ez[k + P], e_ij = Ez[k + 1 + P];
    const double2 e_jm = Ez[k + 1 - P], e_im = Ez[k - P], e_ijm = Ez[k + P - 1], e_jmm = Ez[k - 1];
    const double2 ns_x = v.ns_r2 * (((e_ij + e_jm) - twice(e_j)) - ((e_i + e_im) - twice(e)));
    const double2 ns_y = v.ns_r2 * (((e_ij + e_ijm) - twice(e_i)) - ((e_j + e_jmm) - twice(e)));
    dy_term = (e_j - e) + ns_x;
    dx_term = (e_i - e) + ns_y;
  } else {                                              // fdtdTM.c:315-327
    const double2 *__restrict__ Ezx = F(B200FDTD_STM_EZX);
    const double2 *__restrict__ Ezy = F(B200FDTD_STM_EZY);
    const double2 zx = Ezx[k], zy = Ezy[k];
    dy_term = ((Ezx[k + 1] - zx) + Ezy[k + 1]) - zy;
    dx_term = ((Ezx[k + P] - zx) + Ezy[k + P]) - zy;
  }
  double c_hx, c_hxly, c_hy, c_hylx;
  if (!LEAN) {
    c_hx = v.c[B200FDTD_STM_C_HX][k];  c_hxly = v.c[B200FDTD_STM_C_HXLY][k];
    c_hy = v.c[B200FDTD_STM_C_HY][k];  c_hylx = v.c[B200FDTD_STM_C_HYLX][k];
  } else if (NS) {
    c_hx = v.tj[B200FDTD_LNS_J_C_HX * v.pitch + c];
    c_hy = v.ti[B200FDTD_LNS_I_C_HY * v.rows + r];
    c_hxly = over_den(v.c[B200FDTD_STM_C_HXLY][k], v.tj[B200FDTD_LNS_J_DEN_HX * v.pitch + c]);
    c_hylx = over_den(v.c[B200FDTD_STM_C_HYLX][k], v.ti[B200FDTD_LNS_I_DEN_HY * v.rows + r]);
  } else {
    c_hx = v.tj[B200FDTD_LTM_J_C_HX * v.pitch + c];  c_hxly = v.tj[B200FDTD_LTM_J_C_HXLY * v.pitch + c];
    c_hy = v.ti[B200FDTD_LTM_I_C_HY * v.rows + r];   c_hylx = v.ti[B200FDTD_LTM_I_C_HYLX * v.rows + r];
  }
  F(B200FDTD_STM_HX)[k] = c_hx * F(B200FDTD_STM_HX)[k] - c_hxly * dy_term;
  F(B200FDTD_STM_HY)[k] = c_hy * F(B200FDTD_STM_HY)[k] + c_hylx * dx_term;
}

template <bool NS, bool LEAN, bool BATCH>
__global__ void __launch_bounds__(kBlock, B200_SPLIT_TM_MIN_BLOCKS(NS)) split_tm_e_kernel(const SplitView v)
{
  int r, c; size_t k;
  if (!locate(v, r, c, k)) return;
  const size_t sim = BATCH ? (size_t)blockIdx.y * v.plane : 0;   // field arrays only: coefficients are shared
  const double2 *__restrict__ Hx = F(B200FDTD_STM_HX);
  const double2 *__restrict__ Hy = F(B200FDTD_STM_HY);
  double c_ezx, c_ezxlx, c_ezy, c_ezyly, factor;
  if (!LEAN) {
    c_ezx = v.c[B200FDTD_STM_C_EZX][k];  c_ezxlx = v.c[B200FDTD_STM_C_EZXLX][k];
    c_ezy = v.c[B200FDTD_STM_C_EZY][k];  c_ezyly = v.c[B200FDTD_STM_C_EZYLY][k];
    factor = v.c[B200FDTD_DENSE_SRC0][k];
  } else if (NS) {        // nsFdtdTM.c:283-287: both curl coefficients are (u*z)/(1 + b_ez_x)
    c_ezx = v.ti[B200FDTD_LNS_I_C_EZX * v.rows + r];
    c_ezy = v.tj[B200FDTD_LNS_J_C_EZY * v.pitch + c];
    c_ezxlx = c_ezyly = over_den(v.c[B200FDTD_STM_C_EZXLX][k], v.ti[B200FDTD_LNS_I_DEN_EZ * v.rows + r]);
    factor = v.c[B200FDTD_DENSE_SRC0][k];
  } else {                // fdtdTM.c:230-234 evaluated here from eps and the 1-D sigmas
    const double eps = v.eps0[k];
    const double inv_eps = one_over(eps);
    const PmlPair px = pml_pair(eps, v.ti[B200FDTD_LTM_I_SIG_EZ_X * v.rows + r], inv_eps);
    const PmlPair py = pml_pair(eps, v.tj[B200FDTD_LTM_J_SIG_EZ_Y * v.pitch + c], inv_eps);
    c_ezx = px.c;  c_ezxlx = px.l;  c_ezy = py.c;  c_ezyly = py.l;
    factor = inv_eps - 1.0;                        // EPSILON_0_S / eps - 1.0 (field.c:193)
  }
  // fdtdTM.c:302-308 / nsFdtdTM.c:95-108
  double2 ezx = c_ezx * F(B200FDTD_STM_EZX)[k] + c_ezxlx * (Hy[k] - Hy[k - v.pitch]);
  double2 ezy = c_ezy * F(B200FDTD_STM_EZY)[k] - c_ezyly * (Hx[k] - Hx[k - 1]);
  double2 ez;
  if (NS) {          // source on Ezy, then Ez = Ezx + Ezy (nsFdtdTM.c:73-79)
    if (factor != 0.0 && cw_on<BATCH>(v, 0)) ezy = ezy + cw_term(cw_of<BATCH>(v, 0), r - 1, v.j_base + c, factor);
    ez = ezx + ezy;
  } else {           // Ez = Ezx + Ezy first, then the source on Ezx (fdtdTM.c:290-297,310-312)
    ez = ezx + ezy;
    if (factor != 0.0 && cw_on<BATCH>(v, 0)) ezx = ezx + cw_term(cw_of<BATCH>(v, 0), r - 1, v.j_base + c, factor);
  }
  F(B200FDTD_STM_EZX)[k] = ezx;
  F(B200FDTD_STM_EZY)[k] = ezy;
  F(B200FDTD_STM_EZ)[k] = ez;
}

// ---------------------------------------------------------------- TE family ------
// slots: 0 Hz 1 Hzx 2 Hzy 3 Ex 4 Ey
template <bool NS, bool LEAN, bool INTERIOR, bool BATCH>          // LEAN: kind 1 only (NS TE keeps its dense arrays)
__global__ void __launch_bounds__(kBlock, B200_SPLIT_TE_MIN_BLOCKS) split_te_e_kernel(const SplitView v)
{
  int r, c; size_t k;
  if (!locate(v, r, c, k)) return;
  const size_t sim = BATCH ? (size_t)blockIdx.y * v.plane : 0;   // field arrays only: coefficients are shared
  const int P = v.pitch;
  double2 dy_term, dx_term;
  if (NS) {                                             // nsFdtdTE.c:270-289
    const double2 *__restrict__ Hz = F(B200FDTD_STE_HZ);
    const double2 h = Hz[k], h_jm = Hz[k - 1], h_im = Hz[k - P];
    const double2 h_i = Hz[k + P], h_j = Hz[k + 1];
    const double2 ns_x = v.ns_r2 * (((h_i + h_im) - twice(h)) - ((Hz[k - 1 + P] + Hz[k - 1 - P]) - twice(h_jm)));
    const double2 ns_y = v.ns_r2 * (((h_j + h_jm) - twice(h)) - ((Hz[k - P + 1] + Hz[k - P - 1]) - twice(h_im)));
    dy_term = (h - h_jm) + ns_x;
    dx_term = (h - h_im) + ns_y;
  } else {                                              // fdtdTE.c:293-301
    const double2 *__restrict__ Hzx = F(B200FDTD_STE_HZX);
    const double2 *__restrict__ Hzy = F(B200FDTD_STE_HZY);
    const double2 zx = Hzx[k], zy = Hzy[k];
    dy_term = ((zx - Hzx[k - 1]) + zy) - Hzy[k - 1];
    dx_term = ((zx - Hzx[k - P]) + zy) - Hzy[k - P];
  }
  double c_ex, c_exly, c_ey, c_eylx, fx, fy;
  const bool inside = INTERIOR && block_in_interior(v, r, c);
  if (inside) {           // C_EX == C_EY == 1.0 here: 1.0 * x == x
    c_ex = 1.0;  c_exly = v.c[B200FDTD_STE_C_EXLY][k];
    c_ey = 1.0;  c_eylx = v.c[B200FDTD_STE_C_EYLX][k];
    fx = v.c[B200FDTD_DENSE_SRC0][k];  fy = v.c[B200FDTD_DENSE_SRC1][k];
  } else if (!LEAN) {
    c_ex = v.c[B200FDTD_STE_C_EX][k];  c_exly = v.c[B200FDTD_STE_C_EXLY][k];
    c_ey = v.c[B200FDTD_STE_C_EY][k];  c_eylx = v.c[B200FDTD_STE_C_EYLX][k];
    fx = v.c[B200FDTD_DENSE_SRC0][k];  fy = v.c[B200FDTD_DENSE_SRC1][k];
  } else {                // fdtdTE.c:229-233 evaluated here; the source acts on Ey only (fdtdTE.c:285)
    const double eps_x = v.eps0[k], eps_y = v.eps1[k];
    const double inv_y = one_over(eps_y);
    const PmlPair px = pml_pair(eps_x, v.tj[B200FDTD_LTE_J_SIG_EX_Y * v.pitch + c], one_over(eps_x));
    const PmlPair py = pml_pair(eps_y, v.ti[B200FDTD_LTE_I_SIG_EY_X * v.rows + r], inv_y);
    c_ex = px.c;  c_exly = px.l;  c_ey = py.c;  c_eylx = py.l;
    fx = 0.0;  fy = inv_y - 1.0;
  }
  double2 ex = (inside ? F(B200FDTD_STE_EX)[k] : c_ex * F(B200FDTD_STE_EX)[k]) + c_exly * dy_term;
  double2 ey = (inside ? F(B200FDTD_STE_EY)[k] : c_ey * F(B200FDTD_STE_EY)[k]) - c_eylx * dx_term;
  const int i = r - 1, j = v.j_base + c;
  if (fx != 0.0 && cw_on<BATCH>(v, 0)) ex = ex + cw_term(cw_of<BATCH>(v, 0), i, j, fx);     // nsFdtdTE.c:247-248
  if (fy != 0.0 && cw_on<BATCH>(v, 1)) ey = ey + cw_term(cw_of<BATCH>(v, 1), i, j, fy);     // fdtdTE.c:285, nsFdtdTE.c:249-250
  F(B200FDTD_STE_EX)[k] = ex;
  F(B200FDTD_STE_EY)[k] = ey;
}

// INTERIOR (NS TE, kind 7): thread blocks inside the rectangle of b200fdtd_set_split_interior -- decay
// coefficients exactly 1.0, the two curl coefficients equal -- read three arrays fewer: same bits
template <bool LEAN, bool INTERIOR, bool BATCH>
__global__ void __launch_bounds__(kBlock, B200_SPLIT_TE_MIN_BLOCKS) split_te_h_kernel(const SplitView v)
{
  int r, c; size_t k;
  if (!locate(v, r, c, k)) return;
  const size_t sim = BATCH ? (size_t)blockIdx.y * v.plane : 0;   // field arrays only: coefficients are shared
  const double2 *__restrict__ Ex = F(B200FDTD_STE_EX);
  const double2 *__restrict__ Ey = F(B200FDTD_STE_EY);
  if (INTERIOR && block_in_interior(v, r, c)) {
    const double g = v.c[B200FDTD_STE_C_HZXLX][k];
    const double2 hzx = F(B200FDTD_STE_HZX)[k] - g * (Ey[k + v.pitch] - Ey[k]);
    const double2 hzy = F(B200FDTD_STE_HZY)[k] + g * (Ex[k + 1] - Ex[k]);
    F(B200FDTD_STE_HZX)[k] = hzx;
    F(B200FDTD_STE_HZY)[k] = hzy;
    F(B200FDTD_STE_HZ)[k] = hzx + hzy;
    return;
  }
  double c_hzx, c_hzxlx, c_hzy, c_hzyly;
  if (!LEAN) {
    c_hzx = v.c[B200FDTD_STE_C_HZX][k];  c_hzxlx = v.c[B200FDTD_STE_C_HZXLX][k];
    c_hzy = v.c[B200FDTD_STE_C_HZY][k];  c_hzyly = v.c[B200FDTD_STE_C_HZYLY][k];
  } else {                // fdtdTE.c:236-240: MU_0_S and a 1-D sigma each
    c_hzx = v.ti[B200FDTD_LTE_I_C_HZX * v.rows + r];   c_hzxlx = v.ti[B200FDTD_LTE_I_C_HZXLX * v.rows + r];
    c_hzy = v.tj[B200FDTD_LTE_J_C_HZY * v.pitch + c];  c_hzyly = v.tj[B200FDTD_LTE_J_C_HZYLY * v.pitch + c];
  }
  // fdtdTE.c:308-320 / nsFdtdTE.c:293-307,237-240
  const double2 hzx = c_hzx * F(B200FDTD_STE_HZX)[k] - c_hzxlx * (Ey[k + v.pitch] - Ey[k]);
  const double2 hzy = c_hzy * F(B200FDTD_STE_HZY)[k] + c_hzyly * (Ex[k + 1] - Ex[k]);
  F(B200FDTD_STE_HZX)[k] = hzx;
  F(B200FDTD_STE_HZY)[k] = hzy;
  F(B200FDTD_STE_HZ)[k] = hzx + hzy;
}

#undef F
}  // namespace

int b200_launch_split_step(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  if (e->r_hi < e->r_lo || e->c_hi < e->c_lo) return B200FDTD_OK;
  SplitView v;
  for (int s = 0; s < 5; s++) v.f[s] = e->field[s];
  for (int s = 0; s < B200FDTD_MAX_DENSE; s++) v.c[s] = e->dense[s];
  v.pitch = e->pitch;
  v.r_lo = e->r_lo; v.c_lo = e->c_lo; v.c_hi = e->c_hi;
  v.nbx = (e->c_hi - e->c_lo + 1 + kBlock - 1) / kBlock;
  v.j_base = e->g.j0 - B200_JOFF;
  v.cw[0] = a->cw[0]; v.cw[1] = a->cw[1];
  v.ns_r2 = a->ns_r2;
  v.eps0 = e->eps[0]; v.eps1 = e->eps[1];
  v.ti = e->tab_i; v.tj = e->tab_j;
  v.rows = e->rows;
  v.in_r_lo = e->split_in_r_lo; v.in_r_hi = e->split_in_r_hi;
  v.in_c_lo = e->split_in_c_lo; v.in_c_hi = e->split_in_c_hi;
  v.plane = e->plane;
  v.batch = e->n_batch > 1 ? e->batch_cw : nullptr;
  v.cw_step = e->split_cw_step;
  const bool batch = e->n_batch > 1;
  const dim3 nblk((unsigned)((long long)v.nbx * (e->r_hi - e->r_lo + 1)), (unsigned)e->n_batch);
  cudaStream_t st = e->stream;
  const bool lean = e->split_lean;
  // KERNEL<..., false> for one simulation, <..., true> for an angle batch (blockIdx.y = simulation)
#define SPLIT_LAUNCH(KERNEL, ...)                                                             \
  do {                                                                                        \
    if (batch) KERNEL<__VA_ARGS__, true><<<nblk, kBlock, 0, st>>>(v);                         \
    else       KERNEL<__VA_ARGS__, false><<<nblk, kBlock, 0, st>>>(v);                        \
  } while (0)
  switch (e->g.kind) {
  case B200FDTD_TM:        // fdtdTM.c:290-297: calcH, calcE, source
    if (lean) { SPLIT_LAUNCH(split_tm_h_kernel, false, true);  SPLIT_LAUNCH(split_tm_e_kernel, false, true); }
    else      { SPLIT_LAUNCH(split_tm_h_kernel, false, false); SPLIT_LAUNCH(split_tm_e_kernel, false, false); }
    break;
  case B200FDTD_TE:        // fdtdTE.c:283-287: calcE, source, calcH
    if (lean) { SPLIT_LAUNCH(split_te_e_kernel, false, true, false);  SPLIT_LAUNCH(split_te_h_kernel, true, false); }
    else      { SPLIT_LAUNCH(split_te_e_kernel, false, false, false); SPLIT_LAUNCH(split_te_h_kernel, false, false); }
    break;
  case B200FDTD_NS_TM:     // nsFdtdTM.c:68-80: calcH, calcE, source, Ez = Ezx + Ezy
    if (lean) { SPLIT_LAUNCH(split_tm_h_kernel, true, true);  SPLIT_LAUNCH(split_tm_e_kernel, true, true); }
    else      { SPLIT_LAUNCH(split_tm_h_kernel, true, false); SPLIT_LAUNCH(split_tm_e_kernel, true, false); }
    break;
  case B200FDTD_NS_TE:     // nsFdtdTE.c:233-251: calcH, Hz = Hzx + Hzy, calcE, sources
    if (v.in_r_hi >= v.in_r_lo && v.in_c_hi >= v.in_c_lo) {      // blocks inside the frame-free rectangle: 3 arrays
      SPLIT_LAUNCH(split_te_h_kernel, false, true);
      SPLIT_LAUNCH(split_te_e_kernel, true, false, true);
    } else {
      SPLIT_LAUNCH(split_te_h_kernel, false, false);
      SPLIT_LAUNCH(split_te_e_kernel, true, false, false);
    }
    break;
  default:
    return b200_fail(B200FDTD_ERR_STATE, "not a split-field kind: %d", e->g.kind);
  }
#undef SPLIT_LAUNCH
  e->launches += 2;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}
