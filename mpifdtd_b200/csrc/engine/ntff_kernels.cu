// ntff_kernels.cu -- near-to-far-field transform on the GPU.
//
// The reference accumulates the time-domain far field every step:
// ntffTM_TimeCalc / ntffTE_TimeCalc (ntffTM.c:293-371, ntffTE.c:68-157) walk the
// closed surface for each of 360 directions and scatter every tangential sample
// into three retarded-time bins (calc(), ntffTM.c:279-288).  That is 360 x P x 6
// colliding read-modify-writes per step and ~85 % of the CPU step.
//
// Here the per-step work is only a surface SAMPLE (P points, two complex values
// each) appended to a history buffer; the 360-direction binning is deferred to
// one PROJECTION kernel that, for every (direction, bin), gathers the samples
// that the reference would have scattered there.  For a point with time shift
// ts a sample taken at step t lands in bins m-1, m, m+1 with
//   T = (t - 1) + ts  (E)  or  (t - 1/2) + ts  (H),  m = floor(T + 1/2),
//   a = 1/2 + T - m,  b = 1 - a,  weights (+b, a-b, -a),
// and m, a, b are recomputed here per (direction, point, step) with exactly those
// double-precision expressions, so only the order of the additions differs from
// the CPU.  Collision-free, deterministic, no atomics.
//
// Post-processing (ntffT?_TimeTranslate, the 8192-point cfft and the wavelength
// interpolation, ntffTM.c:161-232, cfft.c:104-179) also runs here, once per run.
#include "engine.h"

namespace {

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 rmul(double r, double2 z) { return make_double2(r * z.x, r * z.y); }
__device__ __forceinline__ double2 cneg(double2 a) { return make_double2(-a.x, -a.y); }
// gcc's complex x complex product without -ffast-math: (ac - bd) + i(ad + bc)
__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// ---- per-step surface sample --------------------------------------------------
// TM (ntffTM.c:326-369): bottom/top sample Ez and the j-averaged Hx, right/left
// sample Ez and the i-averaged Hy; top and left are negated.
// TE (ntffTE.c:100-155): bottom/top sample Ex and j-averaged Hz, right/left Ey and
// i-averaged Hz; here bottom and right are the negated ones.
template <typename C> __device__ __forceinline__ double2 widen(C v) { return make_double2((double)v.x, (double)v.y); }
// B/mu0 as the engine's H phase forms it: IEEE division in double, one multiplication by
// RN(1/mu0) in the single-precision path (upml_common.cuh div_const)
__device__ __forceinline__ double2 h_from_b(double2 b, double mu0) { return make_double2(b.x / mu0, b.y / mu0); }
__device__ __forceinline__ double2 h_from_b(float2 b, double mu0)
{
  const float r = (float)(1.0 / mu0);
  return make_double2((double)(b.x * r), (double)(b.y * r));
}

// C = double2, or float2 for single-precision engines (the history is double either way)
template <typename C>
__global__ void ntff_sample_kernel(const NtffPoint *__restrict__ pts, int n_local, int is_tm,
                                   const C *__restrict__ e_a,   // TM Ez   | TE Ex
                                   const C *__restrict__ e_b,   // TM Ez   | TE Ey
                                   const C *__restrict__ h_a,   // TM Hx   | TE Hz
                                   const C *__restrict__ h_b,   // TM Hy   | TE Hz
                                   int pitch, double2 *hist_e, double2 *hist_h, int max_time, int t,
                                   // when the H arrays are not kept up to date (H == B/mu0 is formed on
                                   // demand): b_a/b_b = Bx/By (TM), divisor = mu0, and c_lo = first updated
                                   // column -- left of it lies the ring / a halo column, which only the H
                                   // array holds.  b_a == nullptr: read H directly.
                                   const C *__restrict__ b_a, const C *__restrict__ b_b,
                                   double h_divisor, int c_lo, size_t plane,
                                   const double *__restrict__ clock)     // multi-step replay: t = (int)*clock
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_local) return;
  if (clock != nullptr) {
    t = (int)*clock;
    if (t < 0 || t >= max_time) return;
  }
  const NtffPoint pt = pts[p];
  // blockIdx.y = simulation of a batched engine (0 otherwise): shift every array to its plane
  const size_t off = (size_t)blockIdx.y * plane;
  e_a += off; e_b += off; h_a += off; h_b += off;
  if (b_a != nullptr) { b_a += off; b_b += off; }
  hist_e += (size_t)blockIdx.y * n_local * max_time;
  hist_h += (size_t)blockIdx.y * n_local * max_time;
  const bool along_x = (pt.edge == 0 || pt.edge == 2);      // bottom / top edges
  double2 ev = widen(along_x ? e_a[pt.k] : e_b[pt.k]);
  double2 h0, h1;
  if (b_a == nullptr) {
    h0 = widen(along_x ? h_a[pt.k] : h_b[pt.k]);
    h1 = widen(along_x ? h_a[pt.k - 1] : h_b[pt.k - pitch]);
  } else {                                              // same quotients the H phase would have stored
    h0 = h_from_b(along_x ? b_a[pt.k] : b_b[pt.k], h_divisor);
    const bool left_is_outside = along_x && (int)(pt.k % pitch) - 1 < c_lo;
    if (left_is_outside) h1 = widen(h_a[pt.k - 1]);
    else                 h1 = h_from_b(along_x ? b_a[pt.k - 1] : b_b[pt.k - pitch], h_divisor);
  }
  double2 hv = rmul(0.5, cadd(h0, h1));
  const bool negate = is_tm ? (pt.edge >= 2) : (pt.edge < 2);
  if (negate) { ev = cneg(ev); hv = cneg(hv); }
  hist_e[(size_t)p * max_time + t] = ev;
  hist_h[(size_t)p * max_time + t] = hv;
}

// ---- deferred projection --------------------------------------------------------
// One thread owns one (direction, bin); the block walks the surface points.  For
// every point the three steps whose taps can reach this bin are t0-1, t0, t0+1
// around the nominal centre; when the shift sits within 1e-9 of a rounding
// knife-edge (frac(ts + 1/2) or frac(ts) ~ 0) five steps are examined so that no
// tap the reference would have written is missed.
constexpr int kProjBlock = 128;
constexpr int kSpillBins = 8;

__device__ __forceinline__ void gather(double2 &acc, const double2 *__restrict__ series, int max_time,
                                       int q, double ts, double lag, int t_center, int reach, double scale)
{
  for (int dt = -reach; dt <= reach; dt++) {
    const int t = t_center + dt;
    if (t < 0 || t >= max_time) continue;
    const double T = ((double)t - lag) + ts;           // timeE / timeH + timeShift
    const int m = (int)floor(T + 0.5);
    const int tap = q - m;
    if (tap < -1 || tap > 1) continue;
    const double a = (0.5 + T) - m;
    const double b = 1.0 - a;
    const double w = tap < 0 ? b : (tap == 0 ? a - b : -a);
    // scale = 1 (exact identity) for the serial solvers; the MPI-variant ntff() multiplies
    // every tap by coef after the weight: Ux[..] += ez*b_e*coef (mpiTM_UPML.c:901-906)
    acc = cadd(acc, rmul(scale, rmul(w, series[t])));
  }
}

__global__ void __launch_bounds__(kProjBlock)
ntff_project_kernel(const NtffPoint *__restrict__ pts, const double *__restrict__ ts_tab, int n_local,
                    const double2 *__restrict__ hist_e, const double2 *__restrict__ hist_h,
                    int max_time, int steps, int n_bins, int n_angles, int is_tm, int array_size,
                    double tap_scale, double2 *uw)
{
  __shared__ double s_ts[kProjBlock];
  __shared__ int s_edge[kProjBlock];
  const int ang = blockIdx.y;
  // blockIdx.z = simulation of a batched engine: its own history and U/W block, same plan
  hist_e += (size_t)blockIdx.z * n_local * max_time;
  hist_h += (size_t)blockIdx.z * n_local * max_time;
  uw += (size_t)blockIdx.z * 3 * n_angles * n_bins;
  const int q_block = blockIdx.x * kProjBlock;
  const int q = q_block + (int)threadIdx.x;
  double2 acc[3] = { make_double2(0, 0), make_double2(0, 0), make_double2(0, 0) };
  const int t_limit = steps < max_time ? steps : max_time;

  // The surface is walked in chunks of kProjBlock points: every thread tests one
  // point of the chunk against this block's bin window, so the (common) case of a
  // point whose retarded time falls outside the recorded steps costs one load per
  // 128 points instead of one loop trip each.
  for (int p0 = 0; p0 < n_local; p0 += kProjBlock) {
    const int pm = p0 + (int)threadIdx.x;
    double my_ts = 0.0;
    int my_edge = -1;                                  // -1: nothing to gather from this point
    if (pm < n_local) {
      my_ts = ts_tab[(size_t)ang * n_local + pm];
      const int se = (int)floor(my_ts + 0.5) - 1, sh = (int)floor(my_ts);
      // steps that can reach bins q_block .. q_block+127 (two extra each side for knife-edges)
      const int t_hi = q_block + kProjBlock - 1 - (se < sh ? se : sh) + 2;
      const int t_lo = q_block - (se > sh ? se : sh) - 2;
      if (t_hi >= 0 && t_lo < t_limit) my_edge = pts[pm].edge;
    }
    const int any = __syncthreads_or(my_edge >= 0);
    if (!any) continue;
    s_ts[threadIdx.x] = my_ts;
    s_edge[threadIdx.x] = my_edge;
    __syncthreads();
    const int chunk = (n_local - p0) < kProjBlock ? (n_local - p0) : kProjBlock;
    for (int c = 0; c < chunk; c++) {
      const int edge = s_edge[c];
      if (edge < 0) continue;
      const int p = p0 + c;
      const double ts = s_ts[c];
      const double fl_e = floor(ts + 0.5), fl_h = floor(ts);
      // nominal centres: E has m = (t-1) + floor(ts+1/2), H has m = t + floor(ts)
      const int shift_e = (int)fl_e - 1, shift_h = (int)fl_h;
      const double fe = (ts + 0.5) - fl_e, fh = ts - fl_h;
      const int reach_e = (fe < 1e-9 || fe > 1.0 - 1e-9) ? 2 : 1;
      const int reach_h = (fh < 1e-9 || fh > 1.0 - 1e-9) ? 2 : 1;
      const bool along_x = (edge == 0 || edge == 2);
      // TM: E -> Ux (0) on bottom/top, Uy (1) on right/left; H -> Wz (2)
      // TE: E -> Uz (2);  H -> Wx (0) on bottom/top, Wy (1) on right/left
      const int slot_e = is_tm ? (along_x ? 0 : 1) : 2;
      const int slot_h = is_tm ? 2 : (along_x ? 0 : 1);
      if (q < n_bins) {
        gather(acc[slot_e], hist_e + (size_t)p * max_time, t_limit, q, ts, 1.0, q - shift_e, reach_e, tap_scale);
        gather(acc[slot_h], hist_h + (size_t)p * max_time, t_limit, q, ts, 0.5, q - shift_h, reach_h, tap_scale);
      }
    }
    __syncthreads();
  }
  // Row spill of the reference's flat [360][arraySize] storage: a tap at index
  // arraySize + q' of direction ang-1 is physically bin q' of direction ang
  // (calc() has no bound check, ntffTM.c:285-287).  Only the last steps of the
  // points with the largest shift get there, so this touches a few low bins.
  if (array_size > 0 && ang > 0 && q_block == 0) {          // uniform per block
    const int qv = q + array_size;
    for (int p0 = 0; p0 < n_local; p0 += kProjBlock) {
      const int pm = p0 + (int)threadIdx.x;
      double my_ts = 0.0;
      int my_edge = -1;
      if (pm < n_local) {
        my_ts = ts_tab[(size_t)(ang - 1) * n_local + pm];
        if (my_ts + (double)t_limit + 2.0 >= (double)array_size) my_edge = pts[pm].edge;
      }
      const int any = __syncthreads_or(my_edge >= 0);
      if (!any) continue;
      s_ts[threadIdx.x] = my_ts;
      s_edge[threadIdx.x] = my_edge;
      __syncthreads();
      const int chunk = (n_local - p0) < kProjBlock ? (n_local - p0) : kProjBlock;
      if (q < kSpillBins && q < n_bins) {
        for (int c = 0; c < chunk; c++) {
          const int edge = s_edge[c];
          if (edge < 0) continue;
          const int p = p0 + c;
          const double ts = s_ts[c];
          const bool along_x = (edge == 0 || edge == 2);
          const int slot_e = is_tm ? (along_x ? 0 : 1) : 2;
          const int slot_h = is_tm ? 2 : (along_x ? 0 : 1);
          gather(acc[slot_e], hist_e + (size_t)p * max_time, t_limit, qv, ts, 1.0,
                 qv - ((int)floor(ts + 0.5) - 1), 2, tap_scale);
          gather(acc[slot_h], hist_h + (size_t)p * max_time, t_limit, qv, ts, 0.5,
                 qv - (int)floor(ts), 2, tap_scale);
        }
      }
      __syncthreads();
    }
  }
  // The same storage quirk at the other end: during the first steps a point whose shift is
  // ~0 or slightly negative (C_0_S = 0.7071 makes |r|/C exceed RFperC by 2e-5 at the box
  // corners) has m = floor(T + 1/2) <= 0, and its taps at index -1, -2 of direction ang+1 are
  // physically bins arraySize-1, arraySize-2 of direction ang.  They carry exact zeros unless
  // a source sits on the surface (the opt-in plane wave does); visible only when all
  // arraySize bins are kept.
  if (array_size > 0 && ang + 1 < n_angles && q_block + kProjBlock > array_size - kSpillBins &&
      q_block < array_size) {                                   // uniform per block
    const int qv = q - array_size;
    for (int p0 = 0; p0 < n_local; p0 += kProjBlock) {
      const int pm = p0 + (int)threadIdx.x;
      double my_ts = 0.0;
      int my_edge = -1;
      if (pm < n_local) {
        my_ts = ts_tab[(size_t)(ang + 1) * n_local + pm];
        if (my_ts < 2.0) my_edge = pts[pm].edge;
      }
      const int any = __syncthreads_or(my_edge >= 0);
      if (!any) continue;
      s_ts[threadIdx.x] = my_ts;
      s_edge[threadIdx.x] = my_edge;
      __syncthreads();
      const int chunk = (n_local - p0) < kProjBlock ? (n_local - p0) : kProjBlock;
      if (qv < 0 && qv >= -kSpillBins && q < n_bins) {
        for (int c = 0; c < chunk; c++) {
          const int edge = s_edge[c];
          if (edge < 0) continue;
          const int p = p0 + c;
          const double ts = s_ts[c];
          const bool along_x = (edge == 0 || edge == 2);
          const int slot_e = is_tm ? (along_x ? 0 : 1) : 2;
          const int slot_h = is_tm ? 2 : (along_x ? 0 : 1);
          gather(acc[slot_e], hist_e + (size_t)p * max_time, t_limit, qv, ts, 1.0,
                 qv - ((int)floor(ts + 0.5) - 1), 2, tap_scale);
          gather(acc[slot_h], hist_h + (size_t)p * max_time, t_limit, qv, ts, 0.5,
                 qv - (int)floor(ts), 2, tap_scale);
        }
      }
      __syncthreads();
    }
  }
  if (q < n_bins)
    for (int s = 0; s < 3; s++)
      uw[((size_t)s * n_angles + ang) * n_bins + q] = acc[s];
}

// ---- translate + FFT + wavelength interpolation ---------------------------------
// One block per direction; the 8192-point series lives in shared memory (128 KB).
// Radix-2 decimation in frequency with the positive-exponent twiddles of
// cfft.c:139-150, then bit reversal, then the two-point interpolation of
// ntffTM.c:224-231.  Twiddles come from the host (glibc cexp), so every butterfly
// is the same arithmetic as the reference's.
__global__ void ntff_spectrum_kernel(const double2 *__restrict__ uw, int n_bins, int n_angles,
                                     int max_time, int is_tm, double2 coef, double z0,
                                     const double *__restrict__ cos_phi,
                                     const double *__restrict__ sin_phi,
                                     const double2 *__restrict__ twiddle, int n_fft, int log2n,
                                     int lam0, int lam1, double c_hu_nfft, double *out)
{
  extern __shared__ double2 series[];
  const int ang = blockIdx.x;
  const double2 *A0 = uw + ((size_t)0 * n_angles + ang) * n_bins;
  const double2 *A1 = uw + ((size_t)1 * n_angles + ang) * n_bins;
  const double2 *A2 = uw + ((size_t)2 * n_angles + ang) * n_bins;
  // theta = 0: sx = cos(phi), sy = sin(phi), sz = -1, px = -sin(phi), py = cos(phi)
  const double cp = cos_phi[ang], sp = sin_phi[ang];
  const double cth = 1.0;                                    // cos(theta), theta = 0
  const double sx = cth * cp, sy = cth * sp, sz = -cth, px = -sp, py = cp;

  for (int n = threadIdx.x; n < n_fft; n += blockDim.x) {
    double2 val = make_double2(0, 0);
    if (n < max_time) {
      if (is_tm) {
        // ntffTM.c:183-187: Eth = coef * (-Z0*(Wz*sz) - (Ux*px + Uy*py))
        const double2 wth = rmul(sz, A2[n]);
        const double2 uph = cadd(rmul(px, A0[n]), rmul(py, A1[n]));
        val = cmul(coef, csub(rmul(-z0, wth), uph));
      } else {
        // ntffTE.c:45-49: Eph = coef * (-Z0*(Wx*px + Wy*py) + Uz*sz)
        const double2 wph = cadd(rmul(px, A0[n]), rmul(py, A1[n]));
        const double2 uth = rmul(sz, A2[n]);
        val = cmul(coef, cadd(rmul(-z0, wph), uth));
      }
    }
    series[n] = val;
  }
  (void)sx; (void)sy;
  __syncthreads();

  // stages: half = n/2, n/4, ..., 1 ; butterfly (p, q = p + half), twiddle index k = p mod span
  const double2 *tw = twiddle;
  for (int half = n_fft >> 1; half >= 1; half >>= 1) {
    for (int b = threadIdx.x; b < (n_fft >> 1); b += blockDim.x) {
      const int k = b % half;
      const int base = (b / half) * (half << 1) + k;
      const double2 lo = series[base], hi = series[base + half];
      const double2 diff = csub(lo, hi);
      series[base] = cadd(lo, hi);
      series[base + half] = cmul(diff, tw[k]);
    }
    tw += half;
    __syncthreads();
  }

  // interpolation reads X[index], X[index+1] in natural order = bit-reversed positions
  for (int lam = lam0 + (int)threadIdx.x; lam <= lam1; lam += blockDim.x) {
    double pos = c_hu_nfft / lam;
    const int index = (int)floor(pos);
    pos = pos - index;
    const unsigned r0 = __brev((unsigned)index) >> (32 - log2n);
    const unsigned r1 = __brev((unsigned)(index + 1)) >> (32 - log2n);
    const double2 x0 = series[r0], x1 = series[r1];
    const double n0 = x0.x * x0.x + x0.y * x0.y;
    const double n1 = x1.x * x1.x + x1.y * x1.y;
    out[(size_t)(lam - lam0) * n_angles + ang] = ((1 - pos) * n0 + pos * n1) / n_fft;
  }
}

// ---- one-shot frequency-domain NTFF (ntffTM.c:72-158) -----------------------------------
// One block per direction.  Threads stride over the surface points in the reference's
// order (bottom, right, top, left); partial sums are combined by a fixed-shape tree, so the
// result is deterministic.  Per point: phase = cexp(i*k*(rx*r2x + ry*r2y)), Nz += H_t*phase,
// Lx or Ly += Ez*phase, with top/left subtracted.
constexpr int kFreqBlock = 256;

__global__ void __launch_bounds__(kFreqBlock)
ntff_frequency_kernel(const double2 *__restrict__ Ez, const double2 *__restrict__ Hx,
                      const double2 *__restrict__ Hy, int pitch, int j0,
                      int top, int bottom, int left, int right, int cx, int cy, double k,
                      const double *__restrict__ cos_a, const double *__restrict__ sin_a,
                      int n_angles, double2 *out)
{
  __shared__ double2 red[3][kFreqBlock];
  const int ang = blockIdx.x;
  const double rx = cos_a[ang], ry = sin_a[ang];
  const int nx = right - left, ny = top - bottom, P = 2 * nx + 2 * ny;
  double2 nz = make_double2(0, 0), lx = make_double2(0, 0), ly = make_double2(0, 0);
  for (int p = threadIdx.x; p < P; p += kFreqBlock) {
    int edge, i, j;
    if (p < nx)               { edge = 0; i = left + p;               j = bottom; }
    else if (p < nx + ny)     { edge = 1; i = right;                  j = bottom + (p - nx); }
    else if (p < 2 * nx + ny) { edge = 2; i = left + (p - nx - ny);   j = top; }
    else                      { edge = 3; i = left;                   j = bottom + (p - 2 * nx - ny); }
    const size_t kk = (size_t)(i + 1) * pitch + (j - j0) + B200_JOFF;
    const double r2x = i - cx, r2y = j - cy;
    const double inner = rx * r2x + ry * r2y;
    double sn, cs;
    sincos(k * inner, &sn, &cs);
    const double2 phase = make_double2(cs, sn);
    const double2 ez = Ez[kk];
    const bool along_x = (edge == 0 || edge == 2);
    const double2 ht = along_x ? rmul(0.5, cadd(Hx[kk], Hx[kk - 1])) : rmul(0.5, cadd(Hy[kk], Hy[kk - pitch]));
    const double2 hp = cmul(ht, phase), ep = cmul(ez, phase);
    if (edge < 2) { nz = cadd(nz, hp); if (along_x) lx = cadd(lx, ep); else ly = cadd(ly, ep); }
    else          { nz = csub(nz, hp); if (along_x) lx = csub(lx, ep); else ly = csub(ly, ep); }
  }
  red[0][threadIdx.x] = nz; red[1][threadIdx.x] = lx; red[2][threadIdx.x] = ly;
  __syncthreads();
  for (int half = kFreqBlock / 2; half >= 1; half >>= 1) {
    if ((int)threadIdx.x < half)
      for (int s = 0; s < 3; s++)
        red[s][threadIdx.x] = cadd(red[s][threadIdx.x], red[s][threadIdx.x + half]);
    __syncthreads();
  }
  if (threadIdx.x < 3) out[(size_t)threadIdx.x * n_angles + ang] = red[threadIdx.x][0];
}

bool is_tm(int kind)
{
  return kind == B200FDTD_TM_UPML || kind == B200FDTD_MPI_TM_UPML || kind == B200FDTD_TM ||
         kind == B200FDTD_NS_TM;
}

}  // namespace

int b200_launch_ntff_sample(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  NtffState &n = e->ntff;
  if (!n.ready) return B200FDTD_OK;
  const int t = (int)a->time;
  if (!e->clock_mode && (t < 0 || t >= n.max_time))
    return b200_fail(B200FDTD_ERR_ARG, "NTFF sample at step %d outside [0, %d)", t, n.max_time);
  if (n.n_local > 0) {
    const bool tm = is_tm(e->g.kind);
    const int s_ea = tm ? (int)B200FDTD_TM_EZ : (int)B200FDTD_TE_EX, s_eb = tm ? (int)B200FDTD_TM_EZ : (int)B200FDTD_TE_EY;
    const int s_ha = tm ? (int)B200FDTD_TM_HX : (int)B200FDTD_TE_HZ, s_hb = tm ? (int)B200FDTD_TM_HY : (int)B200FDTD_TE_HZ;
    const int s_ba = tm ? (int)B200FDTD_TM_BX : (int)B200FDTD_TE_BZ, s_bb = tm ? (int)B200FDTD_TM_BY : (int)B200FDTD_TE_BZ;
    const bool from_b = e->h_stale;                     // H arrays not kept: sample B/mu0
    const dim3 blocks((n.n_local + 127) / 128, e->n_batch);
    if (e->fp32) {
      const float2 *const *f = (const float2 *const *)e->field;
      ntff_sample_kernel<float2><<<blocks, 128, 0, e->stream>>>(
          n.pts, n.n_local, tm ? 1 : 0, f[s_ea], f[s_eb], f[s_ha], f[s_hb], e->pitch, n.hist_e, n.hist_h,
          n.max_time, t, from_b ? f[s_ba] : nullptr, from_b ? f[s_bb] : nullptr, e->g.mu0, e->c_lo, e->plane,
          e->clock_mode ? e->clock_dev : nullptr);
    } else {
      double2 *const *f = e->field;
      ntff_sample_kernel<double2><<<blocks, 128, 0, e->stream>>>(
          n.pts, n.n_local, tm ? 1 : 0, f[s_ea], f[s_eb], f[s_ha], f[s_hb], e->pitch, n.hist_e, n.hist_h,
          n.max_time, t, from_b ? f[s_ba] : nullptr, from_b ? f[s_bb] : nullptr, e->g.mu0, e->c_lo, e->plane,
          e->clock_mode ? e->clock_dev : nullptr);
    }
    e->launches++;
    B200_CUDA(cudaGetLastError());
  }
  if (!e->clock_mode && t + 1 > n.steps_recorded) n.steps_recorded = t + 1;
  return B200FDTD_OK;
}

namespace {
__global__ void clock_advance_kernel(double *clock) { clock[0] += 1.0; }   // field_nextStep: time += 1 (field.c:313)
}

int b200_launch_clock_advance(b200fdtd_engine *e)
{
  clock_advance_kernel<<<1, 1, 0, e->stream>>>(e->clock_dev);
  e->launches++;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}

int b200_launch_ntff_project(b200fdtd_engine *e)
{
  NtffState &n = e->ntff;
  if (!n.ready) return b200_fail(B200FDTD_ERR_STATE, "ntff_project before set_ntff_plan");
  dim3 grid((n.n_bins + kProjBlock - 1) / kProjBlock, n.n_angles, e->n_batch);
  ntff_project_kernel<<<grid, kProjBlock, 0, e->stream>>>(
      n.pts, n.ts, n.n_local, n.hist_e, n.hist_h, n.max_time, n.steps_recorded, n.n_bins,
      n.n_angles, is_tm(e->g.kind) ? 1 : 0, n.array_size, n.tap_scale, n.uw);
  e->launches++;
  B200_CUDA(cudaGetLastError());
  return B200FDTD_OK;
}

int b200_run_ntff_frequency(b200fdtd_engine *e, const b200fdtd_freq_args *a, double *out)
{
  if (!is_tm(e->g.kind))
    return b200_fail(B200FDTD_ERR_ARG, "frequency NTFF serves the TM-type kinds (reference: ntffTM_Frequency)");
  if (e->g.nj != e->g.n_py)
    return b200_fail(B200FDTD_ERR_ARG, "frequency NTFF needs the whole grid on one engine");
  if (e->fp32)
    return b200_fail(B200FDTD_ERR_ARG, "frequency NTFF is built for double-precision engines");
  if (a->left < 1 || a->bottom < 1 || a->right >= e->g.n_px || a->top >= e->g.n_py ||
      a->right <= a->left || a->top <= a->bottom || a->n_angles < 1)
    return b200_fail(B200FDTD_ERR_ARG, "bad NTFF box");
  int rc = b200_refresh_h(e);
  if (rc) return rc;
  rc = b200_refresh_e(e);
  if (rc) return rc;
  const bool upml = (e->g.kind == B200FDTD_TM_UPML || e->g.kind == B200FDTD_MPI_TM_UPML);
  const size_t sel = (size_t)e->sel * e->plane;
  const double2 *Ez = e->field[0] + sel;
  const double2 *Hx = e->field[upml ? (int)B200FDTD_TM_HX : (int)B200FDTD_STM_HX] + sel;
  const double2 *Hy = e->field[upml ? (int)B200FDTD_TM_HY : (int)B200FDTD_STM_HY] + sel;
  double *d_cos = nullptr, *d_sin = nullptr;
  double2 *d_out = nullptr;
  B200_CUDA(cudaMalloc(&d_cos, sizeof(double) * a->n_angles));
  B200_CUDA(cudaMalloc(&d_sin, sizeof(double) * a->n_angles));
  B200_CUDA(cudaMalloc(&d_out, sizeof(double2) * 3 * a->n_angles));
  B200_CUDA(cudaMemcpyAsync(d_cos, a->cos_a, sizeof(double) * a->n_angles, cudaMemcpyHostToDevice, e->stream));
  B200_CUDA(cudaMemcpyAsync(d_sin, a->sin_a, sizeof(double) * a->n_angles, cudaMemcpyHostToDevice, e->stream));
  ntff_frequency_kernel<<<a->n_angles, kFreqBlock, 0, e->stream>>>(Ez, Hx, Hy, e->pitch, e->g.j0, a->top, a->bottom,
                                                                  a->left, a->right, a->cx, a->cy, a->k, d_cos,
                                                                  d_sin, a->n_angles, d_out);
  e->launches++;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double2) * 3 * a->n_angles, cudaMemcpyDeviceToHost, e->stream));
  B200_CUDA(cudaStreamSynchronize(e->stream));
  cudaFree(d_cos); cudaFree(d_sin); cudaFree(d_out);
  return B200FDTD_OK;
}

int b200_run_ntff_spectrum(b200fdtd_engine *e, const b200fdtd_spectrum_args *s, double *out)
{
  NtffState &n = e->ntff;
  if (!n.ready) return b200_fail(B200FDTD_ERR_STATE, "ntff_spectrum before set_ntff_plan");
  int log2n = 0;
  while ((1 << log2n) < s->n_fft) log2n++;
  if ((1 << log2n) != s->n_fft || s->n_fft < 2 || n.max_time > s->n_fft)
    return b200_fail(B200FDTD_ERR_ARG, "n_fft=%d must be a power of two >= max_time=%d", s->n_fft,
                     n.max_time);
  const int n_lam = s->lambda_last_nm - s->lambda_first_nm + 1;
  if (n_lam <= 0) return b200_fail(B200FDTD_ERR_ARG, "empty wavelength range");

  const size_t tw_count = (size_t)s->n_fft - 1;
  if (n.sp_n_fft != s->n_fft || n.sp_n_lam != n_lam) {          // (re)allocate the scratch once
    cudaFree(n.sp_cos); cudaFree(n.sp_sin); cudaFree(n.sp_out); cudaFree(n.sp_tw);
    n.sp_cos = n.sp_sin = n.sp_out = nullptr; n.sp_tw = nullptr; n.sp_n_fft = 0;
    B200_CUDA(cudaMalloc(&n.sp_cos, sizeof(double) * n.n_angles));
    B200_CUDA(cudaMalloc(&n.sp_sin, sizeof(double) * n.n_angles));
    B200_CUDA(cudaMalloc(&n.sp_tw, sizeof(double2) * tw_count));
    B200_CUDA(cudaMalloc(&n.sp_out, sizeof(double) * (size_t)n_lam * n.n_angles));
    n.sp_n_fft = s->n_fft; n.sp_n_lam = n_lam;
  }
  double *d_cos = n.sp_cos, *d_sin = n.sp_sin, *d_out = n.sp_out;
  double2 *d_tw = n.sp_tw;
  B200_CUDA(cudaMemcpyAsync(d_cos, s->cos_phi, sizeof(double) * n.n_angles, cudaMemcpyHostToDevice, e->stream));
  B200_CUDA(cudaMemcpyAsync(d_sin, s->sin_phi, sizeof(double) * n.n_angles, cudaMemcpyHostToDevice, e->stream));
  B200_CUDA(cudaMemcpyAsync(d_tw, s->twiddle, sizeof(double2) * tw_count, cudaMemcpyHostToDevice, e->stream));

  const size_t smem = sizeof(double2) * (size_t)s->n_fft;
  B200_CUDA(cudaFuncSetAttribute(ntff_spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ntff_spectrum_kernel<<<n.n_angles, 1024, smem, e->stream>>>(
      n.uw + (size_t)e->sel * 3 * n.n_angles * n.n_bins, n.n_bins, n.n_angles, n.max_time, is_tm(e->g.kind) ? 1 : 0,
      make_double2(s->coef_re, s->coef_im), s->z0, d_cos, d_sin, d_tw, s->n_fft, log2n,
      s->lambda_first_nm, s->lambda_last_nm, s->c_hu_nfft, d_out);
  e->launches++;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)n_lam * n.n_angles, cudaMemcpyDeviceToHost, e->stream));
  B200_CUDA(cudaStreamSynchronize(e->stream));
  return B200FDTD_OK;
}
