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

namespace {

using namespace upml;

struct SplitView {
  double2 *f[5];
  const double *c[B200FDTD_MAX_DENSE];
  int pitch;
  int r_lo, c_lo, c_hi, nbx;
  int j_base;
  b200fdtd_cw cw[2];
  double ns_r2;
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

__device__ __forceinline__ double2 twice(double2 z) { return make_double2(2 * z.x, 2 * z.y); }

// ---------------------------------------------------------------- TM family ------
// slots: 0 Ez 1 Ezx 2 Ezy 3 Hx 4 Hy
template <bool NS>
__global__ void __launch_bounds__(kBlock) split_tm_h_kernel(const SplitView v)
{
  int r, c; size_t k;
  if (!locate(v, r, c, k)) return;
  const int P = v.pitch;
  double2 dy_term, dx_term;
  if (NS) {                                             // nsFdtdTM.c:127-150
    const double2 *__restrict__ Ez = v.f[B200FDTD_STM_EZ];
    const double2 e = Ez[k], e_j = Ez[k + 1], e_i = Ez[k + P], e_ij = Ez[k + 1 + P];
    const double2 e_jm = Ez[k + 1 - P], e_im = Ez[k - P], e_ijm = Ez[k + P - 1], e_jmm = Ez[k - 1];
    const double2 ns_x = v.ns_r2 * (((e_ij + e_jm) - twice(e_j)) - ((e_i + e_im) - twice(e)));
    const double2 ns_y = v.ns_r2 * (((e_ij + e_ijm) - twice(e_i)) - ((e_j + e_jmm) - twice(e)));
    dy_term = (e_j - e) + ns_x;
    dx_term = (e_i - e) + ns_y;
  } else {                                              // fdtdTM.c:315-327
    const double2 *__restrict__ Ezx = v.f[B200FDTD_STM_EZX];
    const double2 *__restrict__ Ezy = v.f[B200FDTD_STM_EZY];
    const double2 zx = Ezx[k], zy = Ezy[k];
    dy_term = ((Ezx[k + 1] - zx) + Ezy[k + 1]) - zy;
    dx_term = ((Ezx[k + P] - zx) + Ezy[k + P]) - zy;
  }
  v.f[B200FDTD_STM_HX][k] = v.c[B200FDTD_STM_C_HX][k] * v.f[B200FDTD_STM_HX][k] - v.c[B200FDTD_STM_C_HXLY][k] * dy_term;
  v.f[B200FDTD_STM_HY][k] = v.c[B200FDTD_STM_C_HY][k] * v.f[B200FDTD_STM_HY][k] + v.c[B200FDTD_STM_C_HYLX][k] * dx_term;
}

template <bool NS>
__global__ void __launch_bounds__(kBlock) split_tm_e_kernel(const SplitView v)
{
  int r, c; size_t k;
  if (!locate(v, r, c, k)) return;
  const double2 *__restrict__ Hx = v.f[B200FDTD_STM_HX];
  const double2 *__restrict__ Hy = v.f[B200FDTD_STM_HY];
  // fdtdTM.c:302-308 / nsFdtdTM.c:95-108
  double2 ezx = v.c[B200FDTD_STM_C_EZX][k] * v.f[B200FDTD_STM_EZX][k]
              + v.c[B200FDTD_STM_C_EZXLX][k] * (Hy[k] - Hy[k - v.pitch]);
  double2 ezy = v.c[B200FDTD_STM_C_EZY][k] * v.f[B200FDTD_STM_EZY][k]
              - v.c[B200FDTD_STM_C_EZYLY][k] * (Hx[k] - Hx[k - 1]);
  const double factor = v.c[B200FDTD_DENSE_SRC0][k];
  double2 ez;
  if (NS) {          // source on Ezy, then Ez = Ezx + Ezy (nsFdtdTM.c:73-79)
    if (v.cw[0].enabled && factor != 0.0) ezy = ezy + cw_term(v.cw[0], r - 1, v.j_base + c, factor);
    ez = ezx + ezy;
  } else {           // Ez = Ezx + Ezy first, then the source on Ezx (fdtdTM.c:290-297,310-312)
    ez = ezx + ezy;
    if (v.cw[0].enabled && factor != 0.0) ezx = ezx + cw_term(v.cw[0], r - 1, v.j_base + c, factor);
  }
  v.f[B200FDTD_STM_EZX][k] = ezx;
  v.f[B200FDTD_STM_EZY][k] = ezy;
  v.f[B200FDTD_STM_EZ][k] = ez;
}

// ---------------------------------------------------------------- TE family ------
// slots: 0 Hz 1 Hzx 2 Hzy 3 Ex 4 Ey
template <bool NS>
__global__ void __launch_bounds__(kBlock) split_te_e_kernel(const SplitView v)
{
  int r, c; size_t k;
  if (!locate(v, r, c, k)) return;
  const int P = v.pitch;
  double2 dy_term, dx_term;
  if (NS) {                                             // nsFdtdTE.c:270-289
    const double2 *__restrict__ Hz = v.f[B200FDTD_STE_HZ];
    const double2 h = Hz[k], h_jm = Hz[k - 1], h_im = Hz[k - P];
    const double2 h_i = Hz[k + P], h_j = Hz[k + 1];
    const double2 ns_x = v.ns_r2 * (((h_i + h_im) - twice(h)) - ((Hz[k - 1 + P] + Hz[k - 1 - P]) - twice(h_jm)));
    const double2 ns_y = v.ns_r2 * (((h_j + h_jm) - twice(h)) - ((Hz[k - P + 1] + Hz[k - P - 1]) - twice(h_im)));
    dy_term = (h - h_jm) + ns_x;
    dx_term = (h - h_im) + ns_y;
  } else {                                              // fdtdTE.c:293-301
    const double2 *__restrict__ Hzx = v.f[B200FDTD_STE_HZX];
    const double2 *__restrict__ Hzy = v.f[B200FDTD_STE_HZY];
    const double2 zx = Hzx[k], zy = Hzy[k];
    dy_term = ((zx - Hzx[k - 1]) + zy) - Hzy[k - 1];
    dx_term = ((zx - Hzx[k - P]) + zy) - Hzy[k - P];
  }
  double2 ex = v.c[B200FDTD_STE_C_EX][k] * v.f[B200FDTD_STE_EX][k] + v.c[B200FDTD_STE_C_EXLY][k] * dy_term;
  double2 ey = v.c[B200FDTD_STE_C_EY][k] * v.f[B200FDTD_STE_EY][k] - v.c[B200FDTD_STE_C_EYLX][k] * dx_term;
  const int i = r - 1, j = v.j_base + c;
  const double fx = v.c[B200FDTD_DENSE_SRC0][k], fy = v.c[B200FDTD_DENSE_SRC1][k];
  if (v.cw[0].enabled && fx != 0.0) ex = ex + cw_term(v.cw[0], i, j, fx);     // nsFdtdTE.c:247-248
  if (v.cw[1].enabled && fy != 0.0) ey = ey + cw_term(v.cw[1], i, j, fy);     // fdtdTE.c:285, nsFdtdTE.c:249-250
  v.f[B200FDTD_STE_EX][k] = ex;
  v.f[B200FDTD_STE_EY][k] = ey;
}

__global__ void __launch_bounds__(kBlock) split_te_h_kernel(const SplitView v)
{
  int r, c; size_t k;
  if (!locate(v, r, c, k)) return;
  const double2 *__restrict__ Ex = v.f[B200FDTD_STE_EX];
  const double2 *__restrict__ Ey = v.f[B200FDTD_STE_EY];
  // fdtdTE.c:308-320 / nsFdtdTE.c:293-307,237-240
  const double2 hzx = v.c[B200FDTD_STE_C_HZX][k] * v.f[B200FDTD_STE_HZX][k]
                    - v.c[B200FDTD_STE_C_HZXLX][k] * (Ey[k + v.pitch] - Ey[k]);
  const double2 hzy = v.c[B200FDTD_STE_C_HZY][k] * v.f[B200FDTD_STE_HZY][k]
                    + v.c[B200FDTD_STE_C_HZYLY][k] * (Ex[k + 1] - Ex[k]);
  v.f[B200FDTD_STE_HZX][k] = hzx;
  v.f[B200FDTD_STE_HZY][k] = hzy;
  v.f[B200FDTD_STE_HZ][k] = hzx + hzy;
}

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
  const unsigned nblk = (unsigned)((long long)v.nbx * (e->r_hi - e->r_lo + 1));
  cudaStream_t st = e->stream;
  switch (e->g.kind) {
  case B200FDTD_TM:        // fdtdTM.c:290-297: calcH, calcE, source
    split_tm_h_kernel<false><<<nblk, kBlock, 0, st>>>(v);
    split_tm_e_kernel<false><<<nblk, kBlock, 0, st>>>(v);
    break;
  case B200FDTD_TE:        // fdtdTE.c:283-287: calcE, source, calcH
    split_te_e_kernel<false><<<nblk, kBlock, 0, st>>>(v);
    split_te_h_kernel<<<nblk, kBlock, 0, st>>>(v);
    break;
  case B200FDTD_NS_TM:     // nsFdtdTM.c:68-80: calcH, calcE, source, Ez = Ezx + Ezy
    split_tm_h_kernel<true><<<nblk, kBlock, 0, st>>>(v);
    split_tm_e_kernel<true><<<nblk, kBlock, 0, st>>>(v);
    break;
  case B200FDTD_NS_TE:     // nsFdtdTE.c:233-251: calcH, Hz = Hzx + Hzy, calcE, sources
    split_te_h_kernel<<<nblk, kBlock, 0, st>>>(v);
    split_te_e_kernel<true><<<nblk, kBlock, 0, st>>>(v);
    break;
  default:
    return b200_fail(B200FDTD_ERR_STATE, "not a split-field kind: %d", e->g.kind);
  }
  e->launches += 2;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}
