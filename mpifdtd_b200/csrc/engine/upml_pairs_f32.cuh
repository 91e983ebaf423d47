// upml_pairs_f32.cuh -- two cells per thread for the single-precision UPML kernels.
//
// A complex64 cell is 8 bytes, so the one-cell-per-thread kernels issue 64-bit accesses and keep
// half as many bytes in flight per thread as the double kernels do: ncu shows them latency-bound
// (long-scoreboard stalls, 59-71 % DRAM throughput) at the same ~170 instructions per thread.
// Here a thread owns the aligned pair of columns (c0, c0+1), c0 even: every field access is a
// 128-bit LDG/STG again and the per-cell index/table overhead halves.  The arithmetic is the
// shared t?_?_math() of upml_kernels.cu, so results are bit-identical to the one-cell kernels.
// Forms covered: the default ones (H phase without H stores, E phase from B); anything else
// (store_h, the E-first MPI variants before their first H phase) takes the one-cell kernels.
// Included inside the anonymous namespace of upml_kernels.cu.
#pragma once

struct PairGeom { int nbx2; };   // thread blocks per row, each covering 2*kBlock columns

__device__ __forceinline__ float2 lo(float4 q) { return make_float2(q.x, q.y); }
__device__ __forceinline__ float2 hi(float4 q) { return make_float2(q.z, q.w); }
__device__ __forceinline__ float4 pack(float2 a, float2 b) { return make_float4(a.x, a.y, b.x, b.y); }
__device__ __forceinline__ float4 ld4(const float2 *p) { return *reinterpret_cast<const float4 *>(p); }

// store the pair, or only the cell(s) inside the updated range
__device__ __forceinline__ void st_pair(float2 *p, float2 a, float2 b, bool va, bool vb)
{
  if (va && vb) *reinterpret_cast<float4 *>(p) = pack(a, b);
  else if (va) p[0] = a;
  else if (vb) p[1] = b;
}

__device__ __forceinline__ bool locate_pair(const UpmlViewT<float> &v, const PairGeom g, int &r, int &c0, size_t &k,
                                            size_t &k0, bool &va, bool &vb)
{
  const long long b = blockIdx.x;
  const int rb = (int)(b / g.nbx2);
  const int cb = (int)(b - (long long)rb * g.nbx2);
  r = v.r_lo + rb;
  c0 = (v.c_lo & ~1) + 2 * (cb * kBlock + (int)threadIdx.x);
  k0 = (size_t)r * (size_t)v.pitch + (size_t)c0;
  k = k0 + (size_t)blockIdx.y * v.plane;
  va = c0 >= v.c_lo && c0 <= v.c_hi;
  vb = c0 + 1 >= v.c_lo && c0 + 1 <= v.c_hi;
  return va || vb;
}

// ------------------------------------------------------------------ TM -----
__global__ void __launch_bounds__(kBlock, 4) tm_upml_h_pair_kernel(const __grid_constant__ UpmlViewT<float> v, const PairGeom g)
{
  int r, c0; size_t k, k0; bool va, vb;
  if (!locate_pair(v, g, r, c0, k, k0, va, vb)) return;
  const float2 *__restrict__ Ez = v.f[B200FDTD_TM_EZ];
  const float4 ez = ld4(Ez + k), ez_i1 = ld4(Ez + k + v.pitch);
  const float2 ez_j2 = Ez[k + 2];
  const float4 mx_old = ld4(v.f[B200FDTD_TM_MX] + k), bx_old = ld4(v.f[B200FDTD_TM_BX] + k);
  const float4 my_old = ld4(v.f[B200FDTD_TM_MY] + k), by_old = ld4(v.f[B200FDTD_TM_BY] + k);
  const float2 c_mx   = *reinterpret_cast<const float2 *>(v.tj + B200FDTD_TMJ_C_MX * v.pitch + c0);
  const float2 c_mxez = *reinterpret_cast<const float2 *>(v.tj + B200FDTD_TMJ_C_MXEZ * v.pitch + c0);
  const float2 num1   = *reinterpret_cast<const float2 *>(v.tj + B200FDTD_TMJ_NUM_BYMY1 * v.pitch + c0);
  const float2 num0   = *reinterpret_cast<const float2 *>(v.tj + B200FDTD_TMJ_NUM_BYMY0 * v.pitch + c0);
  const float c_bx1 = v.ti[B200FDTD_TMI_C_BXMX1 * v.rows + r], c_bx0 = v.ti[B200FDTD_TMI_C_BXMX0 * v.rows + r];
  const float c_by = v.ti[B200FDTD_TMI_C_BY * v.rows + r], den = v.ti[B200FDTD_TMI_DEN_BYMY * v.rows + r];

  float2 mx_a, bx_a, my_a, by_a, mx_b, bx_b, my_b, by_b;
  tm_h_math<float>(lo(ez), hi(ez), lo(ez_i1), lo(mx_old), lo(bx_old), lo(my_old), lo(by_old), c_mx.x, c_mxez.x, num1.x,
                   num0.x, c_bx1, c_bx0, c_by, den, mx_a, bx_a, my_a, by_a);
  tm_h_math<float>(hi(ez), ez_j2, hi(ez_i1), hi(mx_old), hi(bx_old), hi(my_old), hi(by_old), c_mx.y, c_mxez.y, num1.y,
                   num0.y, c_bx1, c_bx0, c_by, den, mx_b, bx_b, my_b, by_b);
  st_pair(v.f[B200FDTD_TM_MX] + k, mx_a, mx_b, va, vb);
  st_pair(v.f[B200FDTD_TM_BX] + k, bx_a, bx_b, va, vb);
  st_pair(v.f[B200FDTD_TM_MY] + k, my_a, my_b, va, vb);
  st_pair(v.f[B200FDTD_TM_BY] + k, by_a, by_b, va, vb);
  if (v.peer_up_h != nullptr) {           // y-slab halo: top owned column of Hx -> upper neighbour's low ghost
    if (va && c0 == v.c_last) v.peer_up_h[(size_t)r * v.peer_up_pitch + (B200_JOFF - 1)] = div_const(bx_a, v.mu0);
    if (vb && c0 + 1 == v.c_last) v.peer_up_h[(size_t)r * v.peer_up_pitch + (B200_JOFF - 1)] = div_const(bx_b, v.mu0);
  }
}

__global__ void __launch_bounds__(kBlock, 4) tm_upml_e_pair_kernel(const __grid_constant__ UpmlViewT<float> v, const PairGeom g)
{
  int r, c0; size_t k, k0; bool va, vb;
  if (!locate_pair(v, g, r, c0, k, k0, va, vb)) return;
  const float2 *__restrict__ Bx = v.f[B200FDTD_TM_BX];
  const float2 *__restrict__ By = v.f[B200FDTD_TM_BY];
  const float4 by = ld4(By + k), bx = ld4(Bx + k), by_i0 = ld4(By + k - v.pitch);
  const float2 bx_m = Bx[k - 1];
  const float4 jz_old = ld4(v.f[B200FDTD_TM_JZ] + k), dz_old = ld4(v.f[B200FDTD_TM_DZ] + k);
  const float2 eps = *reinterpret_cast<const float2 *>(v.eps0 + k0);
  const float c_jz = v.ti[B200FDTD_TMI_C_JZ * v.rows + r], c_jzh = v.ti[B200FDTD_TMI_C_JZHXHY * v.rows + r];
  const float2 c_dz   = *reinterpret_cast<const float2 *>(v.tj + B200FDTD_TMJ_C_DZ * v.pitch + c0);
  const float2 c_dzjz = *reinterpret_cast<const float2 *>(v.tj + B200FDTD_TMJ_C_DZJZ * v.pitch + c0);

  // H = B/mu0 on the fly; outside the updated range (ring, neighbour's halo column) only the H arrays hold it
  const float2 hy_a = div_const(lo(by), v.mu0), hy_b = div_const(hi(by), v.mu0);
  const float2 hx_a = div_const(lo(bx), v.mu0), hx_b = div_const(hi(bx), v.mu0);
  float2 hy_i0_a = div_const(lo(by_i0), v.mu0), hy_i0_b = div_const(hi(by_i0), v.mu0);
  float2 hx_j0_a = div_const(bx_m, v.mu0), hx_j0_b = hx_a;
  if (r == v.r_lo) {
    const float4 h = ld4(v.f[B200FDTD_TM_HY] + k - v.pitch);
    hy_i0_a = lo(h); hy_i0_b = hi(h);
  }
  if (c0 == v.c_lo) hx_j0_a = v.f[B200FDTD_TM_HX][k - 1];
  if (c0 + 1 == v.c_lo) hx_j0_b = v.f[B200FDTD_TM_HX][k];

  float2 jz_a, dz_a, ez_a, jz_b, dz_b, ez_b;
  tm_e_math<float>(v, r, c0, k0, hy_a, hy_i0_a, hx_a, hx_j0_a, lo(jz_old), lo(dz_old), eps.x, c_jz, c_jzh, c_dz.x,
                   c_dzjz.x, jz_a, dz_a, ez_a);
  tm_e_math<float>(v, r, c0 + 1, k0 + 1, hy_b, hy_i0_b, hx_b, hx_j0_b, hi(jz_old), hi(dz_old), eps.y, c_jz, c_jzh,
                   c_dz.y, c_dzjz.y, jz_b, dz_b, ez_b);
  st_pair(v.f[B200FDTD_TM_JZ] + k, jz_a, jz_b, va, vb);
  st_pair(v.f[B200FDTD_TM_DZ] + k, dz_a, dz_b, va, vb);
  st_pair(v.f[B200FDTD_TM_EZ] + k, ez_a, ez_b, va, vb);
  if (v.peer_down_e != nullptr) {         // bottom owned column of Ez -> lower neighbour's high ghost
    if (va && c0 == v.c_first) v.peer_down_e[(size_t)r * v.peer_down_pitch + v.peer_down_col] = ez_a;
    if (vb && c0 + 1 == v.c_first) v.peer_down_e[(size_t)r * v.peer_down_pitch + v.peer_down_col] = ez_b;
  }
}

// ------------------------------------------------------------------ TE -----
__global__ void __launch_bounds__(kBlock, 4) te_upml_h_pair_kernel(const __grid_constant__ UpmlViewT<float> v, const PairGeom g)
{
  int r, c0; size_t k, k0; bool va, vb;
  if (!locate_pair(v, g, r, c0, k, k0, va, vb)) return;
  const float2 *__restrict__ Ex = v.f[B200FDTD_TE_EX];
  const float2 *__restrict__ Ey = v.f[B200FDTD_TE_EY];
  const float4 ey_i1 = ld4(Ey + k + v.pitch), ey = ld4(Ey + k), ex = ld4(Ex + k);
  const float2 ex_j2 = Ex[k + 2];
  const float4 mz_old = ld4(v.f[B200FDTD_TE_MZ] + k), bz_old = ld4(v.f[B200FDTD_TE_BZ] + k);
  const float c_mz = v.ti[B200FDTD_TEI_C_MZ * v.rows + r], c_mze = v.ti[B200FDTD_TEI_C_MZEXEY * v.rows + r];
  const float2 c_bz   = *reinterpret_cast<const float2 *>(v.tj + B200FDTD_TEJ_C_BZ * v.pitch + c0);
  const float2 c_bzmz = *reinterpret_cast<const float2 *>(v.tj + B200FDTD_TEJ_C_BZMZ * v.pitch + c0);

  float2 mz_a, bz_a, mz_b, bz_b;
  te_h_math<float>(lo(ey_i1), lo(ey), hi(ex), lo(ex), lo(mz_old), lo(bz_old), c_mz, c_mze, c_bz.x, c_bzmz.x, mz_a, bz_a);
  te_h_math<float>(hi(ey_i1), hi(ey), ex_j2, hi(ex), hi(mz_old), hi(bz_old), c_mz, c_mze, c_bz.y, c_bzmz.y, mz_b, bz_b);
  st_pair(v.f[B200FDTD_TE_MZ] + k, mz_a, mz_b, va, vb);
  st_pair(v.f[B200FDTD_TE_BZ] + k, bz_a, bz_b, va, vb);
  if (v.peer_up_h != nullptr) {
    if (va && c0 == v.c_last) v.peer_up_h[(size_t)r * v.peer_up_pitch + (B200_JOFF - 1)] = div_const(bz_a, v.mu0);
    if (vb && c0 + 1 == v.c_last) v.peer_up_h[(size_t)r * v.peer_up_pitch + (B200_JOFF - 1)] = div_const(bz_b, v.mu0);
  }
}

#ifndef B200_TE_E_PAIR_MIN_BLOCKS
#define B200_TE_E_PAIR_MIN_BLOCKS 3     /* 4 blocks/SM (64 registers) spills */
#endif
__global__ void __launch_bounds__(kBlock, B200_TE_E_PAIR_MIN_BLOCKS) te_upml_e_pair_kernel(const __grid_constant__ UpmlViewT<float> v, const PairGeom g)
{
  int r, c0; size_t k, k0; bool va, vb;
  if (!locate_pair(v, g, r, c0, k, k0, va, vb)) return;
  const float2 *__restrict__ Hz = v.f[B200FDTD_TE_HZ];
  const float2 *__restrict__ Bz = v.f[B200FDTD_TE_BZ];
  const float4 bz = ld4(Bz + k), bz_i0 = ld4(Bz + k - v.pitch);
  const float2 bz_m = Bz[k - 1];
  const float4 jx_old = ld4(v.f[B200FDTD_TE_JX] + k), dx_old = ld4(v.f[B200FDTD_TE_DX] + k);
  const float4 jy_old = ld4(v.f[B200FDTD_TE_JY] + k), dy_old = ld4(v.f[B200FDTD_TE_DY] + k);
  const float2 eps_x = *reinterpret_cast<const float2 *>(v.eps0 + k0);
  const float2 eps_y = *reinterpret_cast<const float2 *>(v.eps1 + k0);
  const float2 c_jx   = *reinterpret_cast<const float2 *>(v.tj + B200FDTD_TEJ_C_JX * v.pitch + c0);
  const float2 c_jxhz = *reinterpret_cast<const float2 *>(v.tj + B200FDTD_TEJ_C_JXHZ * v.pitch + c0);
  const float2 num1   = *reinterpret_cast<const float2 *>(v.tj + B200FDTD_TEJ_NUM_DYJY1 * v.pitch + c0);
  const float2 num0   = *reinterpret_cast<const float2 *>(v.tj + B200FDTD_TEJ_NUM_DYJY0 * v.pitch + c0);
  const float c_dx1 = v.ti[B200FDTD_TEI_C_DXJX1 * v.rows + r], c_dx0 = v.ti[B200FDTD_TEI_C_DXJX0 * v.rows + r];
  const float c_dy = v.ti[B200FDTD_TEI_C_DY * v.rows + r], den = v.ti[B200FDTD_TEI_DEN_DYJY * v.rows + r];

  const float2 hz_a = div_const(lo(bz), v.mu0), hz_b = div_const(hi(bz), v.mu0);
  float2 hz_i0_a = div_const(lo(bz_i0), v.mu0), hz_i0_b = div_const(hi(bz_i0), v.mu0);
  float2 hz_j0_a = div_const(bz_m, v.mu0), hz_j0_b = hz_a;
  if (c0 == v.c_lo) hz_j0_a = Hz[k - 1];          // ring / halo column: only the H array holds it
  if (c0 + 1 == v.c_lo) hz_j0_b = Hz[k];
  if (r == v.r_lo) {
    const float4 h = ld4(Hz + k - v.pitch);
    hz_i0_a = lo(h); hz_i0_b = hi(h);
  }

  float2 jx_a, dx_a, jy_a, dy_a, ex_a, ey_a, jx_b, dx_b, jy_b, dy_b, ex_b, ey_b;
  te_e_math<float>(v, r, c0, k0, hz_a, hz_j0_a, hz_i0_a, lo(jx_old), lo(dx_old), lo(jy_old), lo(dy_old), eps_x.x, eps_y.x,
                   c_jx.x, c_jxhz.x, num1.x, num0.x, c_dx1, c_dx0, c_dy, den, jx_a, dx_a, jy_a, dy_a, ex_a, ey_a);
  te_e_math<float>(v, r, c0 + 1, k0 + 1, hz_b, hz_j0_b, hz_i0_b, hi(jx_old), hi(dx_old), hi(jy_old), hi(dy_old), eps_x.y,
                   eps_y.y, c_jx.y, c_jxhz.y, num1.y, num0.y, c_dx1, c_dx0, c_dy, den, jx_b, dx_b, jy_b, dy_b, ex_b, ey_b);
  st_pair(v.f[B200FDTD_TE_JX] + k, jx_a, jx_b, va, vb);
  st_pair(v.f[B200FDTD_TE_DX] + k, dx_a, dx_b, va, vb);
  st_pair(v.f[B200FDTD_TE_JY] + k, jy_a, jy_b, va, vb);
  st_pair(v.f[B200FDTD_TE_DY] + k, dy_a, dy_b, va, vb);
  st_pair(v.f[B200FDTD_TE_EX] + k, ex_a, ex_b, va, vb);
  st_pair(v.f[B200FDTD_TE_EY] + k, ey_a, ey_b, va, vb);
  if (v.peer_down_e != nullptr) {
    if (va && c0 == v.c_first) v.peer_down_e[(size_t)r * v.peer_down_pitch + v.peer_down_col] = ex_a;
    if (vb && c0 + 1 == v.c_first) v.peer_down_e[(size_t)r * v.peer_down_pitch + v.peer_down_col] = ex_b;
  }
}
