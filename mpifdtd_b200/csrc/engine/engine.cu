// engine.cu -- lifetime, uploads, state access and the C ABI of include/b200fdtd.h.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <sys/mman.h>
#include <vector>
#include "engine.h"

static thread_local char g_last_error[512] = "";

int b200_fail(int code, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof g_last_error, fmt, ap);
  va_end(ap);
  return code;
}

namespace {

bool kind_is_split(int kind)
{
  return kind == B200FDTD_TM || kind == B200FDTD_TE || kind == B200FDTD_NS_TM || kind == B200FDTD_NS_TE;
}

bool kind_is_upml(int kind)
{
  return kind == B200FDTD_TM_UPML || kind == B200FDTD_TE_UPML || kind == B200FDTD_MPI_TM_UPML ||
         kind == B200FDTD_MPI_TE_UPML;
}

int select_device(b200fdtd_engine *e)
{
  B200_CUDA(cudaSetDevice(e->device));
  return B200FDTD_OK;
}

int dev_alloc_zero(b200fdtd_engine *e, void **ptr, size_t bytes)
{
  cudaError_t err = cudaMalloc(ptr, bytes ? bytes : 16);
  if (err != cudaSuccess)
    return b200_fail(B200FDTD_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(err));
  B200_CUDA(cudaMemsetAsync(*ptr, 0, bytes ? bytes : 16, e->stream));
  e->dev_bytes += bytes;
  return B200FDTD_OK;
}

__global__ void fill_double_kernel(double *dst, size_t n, double value)
{
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    dst[k] = value;
}

// Order-independent 64-bit digest of the owned cells of one field plane: every 64-bit word is
// mixed with its GLOBAL position (i * n_py + j, word index), the mixes are summed mod 2^64.  Two
// planes have the same digest iff (up to 2^-64) they hold the same bits in the same cells, and the
// digests of the y-slabs of a split run add up to the digest of the single-slab plane.
__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

template <int WORDS>      // 64-bit words per element: 2 (double2) or 1 (float2)
__global__ void digest_kernel(const unsigned long long *plane, int pitch, int n_px, int nj, long long n_py,
                              long long j0, unsigned long long *out)
{
  unsigned long long acc = 0;
  const size_t n = (size_t)n_px * nj;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const size_t i = t / nj, c = t - i * nj;
    const unsigned long long *cell = plane + ((i + 1) * pitch + c + B200_JOFF) * WORDS;
    const unsigned long long g = ((unsigned long long)i * n_py + j0 + c) * WORDS;
#pragma unroll
    for (int w = 0; w < WORDS; w++) acc += mix64(cell[w] ^ mix64(g + w + 0x9e3779b97f4a7c15ull));
  }
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

__global__ void axpy_double_kernel(double *dst, const double *src, size_t n)
{
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    dst[k] += src[k];
}

// eps[k] = table[index[k]] over rows [i0, i0 + n_rows) of the owned cells (ghosts and padding keep
// the vacuum 1.0); `index` holds just those rows
template <typename T>
__global__ void expand_palette_kernel(const unsigned short *__restrict__ index, const double *__restrict__ table,
                                      int n_values, T *eps, int pitch, int i0, int n_rows, int nj)
{
  const size_t n = (size_t)n_rows * nj;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const size_t i = t / nj, c = t - i * nj;
    const int v = index[t];
    eps[(i0 + i + 1) * pitch + B200_JOFF + c] = (T)table[v < n_values ? v : 0];
  }
}

void free_ntff(b200fdtd_engine *e)
{
  NtffState &n = e->ntff;
  cudaFree(n.pts); cudaFree(n.ts); cudaFree(n.hist_e); cudaFree(n.hist_h); cudaFree(n.uw);
  cudaFree(n.sp_cos); cudaFree(n.sp_sin); cudaFree(n.sp_out); cudaFree(n.sp_tw);
  memset(&n, 0, sizeof n);
}

}  // namespace

extern "C" {

static void peer_release(b200fdtd_engine *e);

const char *b200fdtd_last_error(void) { return g_last_error; }
int b200fdtd_abi_version(void) { return B200FDTD_ABI_VERSION; }

int b200fdtd_struct_size(int32_t which)
{
  switch (which) {
  case 0: return (int)sizeof(b200fdtd_grid);
  case 1: return (int)sizeof(b200fdtd_step_args);
  case 2: return (int)sizeof(b200fdtd_ntff_plan);
  case 3: return (int)sizeof(b200fdtd_spectrum_args);
  case 4: return (int)sizeof(b200fdtd_freq_args);
  case 5: return (int)sizeof(b200fdtd_batch_source);
  case 6: return (int)sizeof(b200fdtd_batch_cw);
  default: return -1;
  }
}

int b200fdtd_device_count(int *count)
{
  if (!count) return b200_fail(B200FDTD_ERR_ARG, "count is NULL");
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess) {
    *count = 0;
    return b200_fail(B200FDTD_ERR_NODEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(err));
  }
  *count = n;
  return B200FDTD_OK;
}

int b200fdtd_host_alloc(void **ptr, uint64_t bytes)
{
  if (!ptr) return b200_fail(B200FDTD_ERR_ARG, "ptr is NULL");
  cudaError_t err = cudaHostAlloc(ptr, bytes ? bytes : 16, cudaHostAllocDefault);
  if (err != cudaSuccess)
    return b200_fail(err == cudaErrorMemoryAllocation ? B200FDTD_ERR_NOMEM : B200FDTD_ERR_NODEVICE,
                     "cudaHostAlloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(err));
  memset(*ptr, 0, bytes);
  return B200FDTD_OK;
}

int b200fdtd_host_free(void *ptr)
{
  if (ptr) B200_CUDA(cudaFreeHost(ptr));
  return B200FDTD_OK;
}

// Getter mirrors.  Pinning 256 MB of 4 KB pages costs ~100 ms, a one-off download of it through
// pageable memory about as much.  A mirror is an anonymous mapping aligned to 2 MB with transparent
// huge pages requested (zero-filled by the kernel, 128 faults instead of 65536), and it is pinned IN
// PLACE (cudaHostRegister: same pointer, so the borrowed pointers callers hold stay valid).
namespace {
struct MirrorRec { void *ptr; size_t bytes; bool pinned; };
MirrorRec g_mirrors[256];
MirrorRec *find_mirror(void *ptr)
{
  for (MirrorRec &m : g_mirrors)
    if (m.ptr == ptr) return &m;
  return nullptr;
}
}  // namespace

int b200fdtd_mirror_alloc(void **ptr, uint64_t bytes)
{
  if (!ptr) return b200_fail(B200FDTD_ERR_ARG, "ptr is NULL");
  MirrorRec *rec = find_mirror(nullptr);
  if (!rec) return b200_fail(B200FDTD_ERR_NOMEM, "too many host mirrors");
  const size_t huge = 2u << 20;
  const size_t len = ((bytes ? bytes : 1) + huge - 1) / huge * huge;
  // over-map by one huge page so the start can be aligned; the unused head and tail are unmapped again
  char *raw = (char *)mmap(nullptr, len + huge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (raw == (char *)MAP_FAILED) return b200_fail(B200FDTD_ERR_NOMEM, "host mirror of %llu bytes", (unsigned long long)bytes);
  char *p = (char *)(((uintptr_t)raw + huge - 1) / huge * huge);
  if (p > raw) munmap(raw, (size_t)(p - raw));
  if (p + len < raw + len + huge) munmap(p + len, (size_t)(raw + len + huge - (p + len)));
  madvise(p, len, MADV_HUGEPAGE);
  rec->ptr = p; rec->bytes = len; rec->pinned = false;
  *ptr = p;
  return B200FDTD_OK;
}

int b200fdtd_mirror_pin(void *ptr, uint64_t bytes)
{
  MirrorRec *rec = ptr ? find_mirror(ptr) : nullptr;
  if (!rec) return b200_fail(B200FDTD_ERR_ARG, "not a mirror");
  (void)bytes;
  if (rec->pinned) return B200FDTD_OK;
  cudaError_t err = cudaHostRegister(ptr, rec->bytes, cudaHostRegisterDefault);
  if (err != cudaSuccess) { cudaGetLastError(); return b200_fail(B200FDTD_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(err)); }
  rec->pinned = true;
  return B200FDTD_OK;
}

int b200fdtd_mirror_free(void *ptr, int32_t pinned)
{
  (void)pinned;
  MirrorRec *rec = ptr ? find_mirror(ptr) : nullptr;
  if (!rec) return B200FDTD_OK;
  if (rec->pinned && cudaHostUnregister(ptr) != cudaSuccess) cudaGetLastError();
  munmap(ptr, rec->bytes);
  rec->ptr = nullptr; rec->bytes = 0; rec->pinned = false;
  return B200FDTD_OK;
}

int b200fdtd_create(const b200fdtd_grid *grid, b200fdtd_engine **out)
{
  if (!grid || !out) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  *out = nullptr;
  if (!kind_is_upml(grid->kind) && !kind_is_split(grid->kind))
    return b200_fail(B200FDTD_ERR_ARG, "solver kind %d is not served by this engine build", grid->kind);
  if (grid->n_px < 3 || grid->n_py < 3 || grid->nj < 1 || grid->j0 < 0 ||
      grid->j0 + grid->nj > grid->n_py || grid->n_pml < 0)
    return b200_fail(B200FDTD_ERR_ARG, "bad grid %dx%d slab [%d,+%d)", grid->n_px, grid->n_py,
                     grid->j0, grid->nj);
  if (grid->i_lo < 0 || grid->i_hi >= grid->n_px || grid->j_lo < 0 || grid->j_hi >= grid->n_py)
    return b200_fail(B200FDTD_ERR_ARG, "update extents outside the grid");
  if (grid->precision != B200FDTD_F64 && grid->precision != B200FDTD_F32)
    return b200_fail(B200FDTD_ERR_ARG, "unknown precision %d", grid->precision);
  if (grid->precision == B200FDTD_F32 && !kind_is_upml(grid->kind))
    return b200_fail(B200FDTD_ERR_ARG, "the single-precision path serves the UPML kinds (2-5)");
  const int n_batch = grid->n_batch > 1 ? grid->n_batch : 1;
  if (n_batch > 1 && ((grid->kind != B200FDTD_TM_UPML && grid->kind != B200FDTD_TE_UPML && !kind_is_split(grid->kind)) ||
                      grid->j0 != 0 || grid->nj != grid->n_py || n_batch > 65535))
    return b200_fail(B200FDTD_ERR_ARG, "batched engines serve the serial UPML kinds (2, 3) and the split-field "
                                       "kinds (0, 1, 6, 7) with the whole grid on one engine, n_batch <= 65535");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return b200_fail(B200FDTD_ERR_NODEVICE, "no CUDA device: the FDTD engine has no CPU path");

  b200fdtd_engine *e = new (std::nothrow) b200fdtd_engine();
  if (!e) return b200_fail(B200FDTD_ERR_NOMEM, "host allocation failed");
  memset(e, 0, sizeof *e);
  e->g = *grid;
  if (grid->device >= 0) e->device = grid->device;
  else if (cudaGetDevice(&e->device) != cudaSuccess) e->device = 0;
  int rc = select_device(e);
  if (rc) { delete e; return rc; }

  cudaError_t err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
  if (err != cudaSuccess) { delete e; return b200_fail(B200FDTD_ERR_CUDA, "stream: %s", cudaGetErrorString(err)); }
  e->own_stream = true;
  cudaEventCreate(&e->ev0);
  cudaEventCreate(&e->ev1);

  e->rows = grid->n_px + 2;
  e->pitch = ((B200_JOFF + grid->nj + 1) + 7) / 8 * 8;
  e->plane = (size_t)e->rows * e->pitch;
  e->n_fields = kind_is_split(grid->kind) ? 5 : 9;
  const bool ntff_only = (grid->flags & B200FDTD_GRID_NTFF_ONLY) != 0;
  if (ntff_only) e->n_fields = 0;                 // history, projection and spectrum only: no field arrays
  e->n_batch = n_batch;
  e->fp32 = grid->precision == B200FDTD_F32;
  e->csize = e->fp32 ? sizeof(float2) : sizeof(double2);
  e->rsize = e->fp32 ? sizeof(float) : sizeof(double);
  e->use_fused = false;     // the one-pass kernel: everywhere it can run (B200FDTD_OPT_FUSED = 1) ...
  e->fused_auto = true;     // ... or, by default, on large single-slab TM grids (2 = auto)
  e->fused_variant = 20;    // 8 consumer warps (256 columns) x 4 row buffers
  e->store_h = false;
  e->h_stale = false;
  e->derived_e = true;      // the one-pass step keeps no E arrays in vacuum row-strips (fused_kernels.cu)
  if (const char *v = getenv("B200FDTD_DERIVED_E")) e->derived_e = atoi(v) != 0;
  e->e_consistent = true;   // at rest E == D == 0
  e->e_stale = false;
  if (const char *v = getenv("B200FDTD_FUSED")) {
    e->use_fused = atoi(v) == 1 && (grid->kind == B200FDTD_TM_UPML || grid->kind == B200FDTD_TE_UPML) && !e->fp32 &&
                   n_batch == 1;
    e->fused_auto = atoi(v) == 2;
  }
  if (const char *v = getenv("B200FDTD_STORE_H")) e->store_h = atoi(v) != 0;
  e->f32_pairs = true;
  if (const char *v = getenv("B200FDTD_F32_PAIRS")) e->f32_pairs = atoi(v) != 0;
  if (const char *v = getenv("B200FDTD_FUSED_SHAPE")) e->fused_variant = atoi(v);
  if (const char *v = getenv("B200FDTD_BAND_ROWS")) e->fused.band_h = atoi(v);
  e->unit_split = 2;
  if (const char *v = getenv("B200FDTD_UNIT_SPLIT")) e->unit_split = atoi(v);
  e->lean_interior = false;
  if (const char *v = getenv("B200FDTD_LEAN_INTERIOR")) e->lean_interior = atoi(v) != 0 && kind_is_upml(grid->kind);
  e->lean_r_lo = e->lean_c_lo = 1;
  e->lean_r_hi = e->lean_c_hi = 0;
  e->split_in_r_lo = e->split_in_c_lo = 1;
  e->split_in_r_hi = e->split_in_c_hi = 0;

  // update extents clipped to this slab, in layout coordinates
  const int jl = grid->j_lo > grid->j0 ? grid->j_lo : grid->j0;
  const int jh = grid->j_hi < grid->j0 + grid->nj - 1 ? grid->j_hi : grid->j0 + grid->nj - 1;
  e->r_lo = grid->i_lo + 1;
  e->r_hi = grid->i_hi + 1;
  e->c_lo = jl - grid->j0 + B200_JOFF;
  e->c_hi = jh - grid->j0 + B200_JOFF;

  for (int s = 0; s < e->n_fields && !rc; s++)
    rc = dev_alloc_zero(e, (void **)&e->field[s], e->plane * e->csize * (size_t)n_batch);
  if (ntff_only) {
    if (!kind_is_upml(grid->kind) || n_batch > 1) rc = b200_fail(B200FDTD_ERR_ARG, "NTFF-only engines: an unbatched UPML kind");
  } else if (kind_is_split(grid->kind)) {
    // dense coefficient arrays and eps maps are allocated by the first set_dense / set_eps
    // call for their slot: the lean form needs only a few of them
    rc = dev_alloc_zero(e, (void **)&e->tab_i, sizeof(double) * B200FDTD_SPLIT_TABS * e->rows);
    if (!rc) rc = dev_alloc_zero(e, (void **)&e->tab_j, sizeof(double) * B200FDTD_SPLIT_TABS * e->pitch);
    if (!rc && n_batch > 1) rc = dev_alloc_zero(e, (void **)&e->batch_cw, sizeof(b200fdtd_batch_cw) * (size_t)n_batch);
  } else {
    const int n_eps = (grid->kind == B200FDTD_TM_UPML || grid->kind == B200FDTD_MPI_TM_UPML) ? 1 : 2;
    for (int s = 0; s < n_eps && !rc; s++)
      rc = dev_alloc_zero(e, (void **)&e->eps[s], e->plane * e->rsize);
    if (!rc) rc = dev_alloc_zero(e, (void **)&e->tab_i, sizeof(double) * B200FDTD_UPML_TABS * e->rows);
    if (!rc) rc = dev_alloc_zero(e, (void **)&e->tab_j, sizeof(double) * B200FDTD_UPML_TABS * e->pitch);
    if (!rc && (grid->kind == B200FDTD_TM_UPML || grid->kind == B200FDTD_TE_UPML)) {
      // per-simulation pulse records: a batched engine needs them, a single one uses them for
      // multi-step replay (b200fdtd_run_steps)
      rc = dev_alloc_zero(e, (void **)&e->batch_src, sizeof(b200fdtd_batch_source) * (size_t)n_batch);
      if (!rc) rc = dev_alloc_zero(e, (void **)&e->clock_dev, sizeof(double));
    }
  }
  if (rc) { b200fdtd_destroy(e); return rc; }
  if (cudaStreamSynchronize(e->stream) != cudaSuccess) {
    b200fdtd_destroy(e);
    return b200_fail(B200FDTD_ERR_CUDA, "initial memset failed");
  }
  // The kernels replace x / mu0 by a reciprocal multiply with one FMA correction (div_const,
  // upml_common.cuh), bit-identical to IEEE division for the reference's MU_0_S (2^31 operands in
  // tests/test_gpu_fused.py).  The identity is a property of the divisor: a caller-supplied mu0 is
  // checked once per distinct value here, and an engine whose divisor breaks it is refused.
  if (kind_is_upml(grid->kind) && !e->fp32) {
    static double checked_mu0 = 0.0;
    if (grid->mu0 != checked_mu0) {
      unsigned long long bad = 0;
      rc = (grid->mu0 > 0.0) ? b200_selftest_division(grid->mu0, 1ull << 22, &bad) : B200FDTD_ERR_ARG;
      if (rc || bad) {
        b200fdtd_destroy(e);
        return b200_fail(rc ? rc : B200FDTD_ERR_STATE, "mu0 = %.17g: the exact-division shortcut disagrees with IEEE "
                         "division for %llu of 2^22 operands", grid->mu0, bad);
      }
      checked_mu0 = grid->mu0;
    }
  }
  *out = e;
  return B200FDTD_OK;
}

int b200fdtd_destroy(b200fdtd_engine *e)
{
  if (!e) return B200FDTD_OK;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  e->e_stale = false;               // nobody will read the E arrays again: nothing to bring up to date
  for (int s = 0; s < B200FDTD_MAX_FIELDS; s++) cudaFree(e->field[s]);
  cudaFree(e->eps[0]); cudaFree(e->eps[1]);
  cudaFree(e->tab_i); cudaFree(e->tab_j); cudaFree(e->batch_src); cudaFree(e->batch_cw); cudaFree(e->cw_tab);
  for (int b = 0; b < 2; b++) {
    if (e->cw_stage[b]) cudaFreeHost(e->cw_stage[b]);
    if (e->cw_stage_done[b]) cudaEventDestroy((cudaEvent_t)e->cw_stage_done[b]);
  }
  cudaFree(e->clock_dev);
  if (e->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)e->graph_exec);
  for (int s = 0; s < B200FDTD_MAX_DENSE; s++) cudaFree(e->dense[s]);
  free_ntff(e);
  b200_fused_release(e);
  peer_release(e);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return B200FDTD_OK;
}

// ---- peer halos ----------------------------------------------------------------------
struct PeerBlob {                       // what b200fdtd_peer_export writes (<= 256 bytes)
  cudaIpcMemHandle_t e_arr, h_arr, flags;
  int32_t nj, pitch, rows, kind;
};
static_assert(sizeof(PeerBlob) <= B200FDTD_PEER_BLOB_BYTES, "peer blob too large");

static int peer_slots(const b200fdtd_engine *e, int *e_slot, int *h_slot)
{
  const bool tm = (e->g.kind == B200FDTD_TM_UPML || e->g.kind == B200FDTD_MPI_TM_UPML);
  *e_slot = tm ? (int)B200FDTD_TM_EZ : (int)B200FDTD_TE_EX;
  *h_slot = tm ? (int)B200FDTD_TM_HX : (int)B200FDTD_TE_HZ;
  return 0;
}

int b200fdtd_peer_export(b200fdtd_engine *e, void *blob)
{
  if (!e || !blob) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  if (!kind_is_upml(e->g.kind) || e->n_batch > 1)
    return b200_fail(B200FDTD_ERR_ARG, "peer halos serve unbatched engines of the UPML kinds");
  int rc = select_device(e); if (rc) return rc;
  if (!e->peer.flags) {
    rc = dev_alloc_zero(e, (void **)&e->peer.flags, 2 * sizeof(unsigned long long));
    if (rc) return rc;
    B200_CUDA(cudaStreamSynchronize(e->stream));
  }
  int es, hs;
  peer_slots(e, &es, &hs);
  PeerBlob b;
  memset(&b, 0, sizeof b);
  B200_CUDA(cudaIpcGetMemHandle(&b.e_arr, e->field[es]));
  B200_CUDA(cudaIpcGetMemHandle(&b.h_arr, e->field[hs]));
  B200_CUDA(cudaIpcGetMemHandle(&b.flags, e->peer.flags));
  b.nj = e->g.nj; b.pitch = e->pitch; b.rows = e->rows; b.kind = e->g.kind | (e->fp32 ? 0x100 : 0);
  memset(blob, 0, B200FDTD_PEER_BLOB_BYTES);
  memcpy(blob, &b, sizeof b);
  return B200FDTD_OK;
}

int b200fdtd_peer_attach(b200fdtd_engine *e, int32_t which, const void *blob)
{
  if (!e || !blob || which < 0 || which > 1) return b200_fail(B200FDTD_ERR_ARG, "bad peer argument");
  int rc = select_device(e); if (rc) return rc;
  if (!e->peer.flags) return b200_fail(B200FDTD_ERR_STATE, "peer_attach before peer_export");
  PeerBlob b;
  memcpy(&b, blob, sizeof b);
  if (b.rows != e->rows || b.kind != (e->g.kind | (e->fp32 ? 0x100 : 0)))
    return b200_fail(B200FDTD_ERR_ARG, "neighbour slab has another shape, solver kind or precision");
  void *p_e = nullptr, *p_h = nullptr, *p_f = nullptr;
  B200_CUDA(cudaIpcOpenMemHandle(&p_f, b.flags, cudaIpcMemLazyEnablePeerAccess));
  if (which == 1) {                     // upper neighbour: I store H into its low ghost column
    B200_CUDA(cudaIpcOpenMemHandle(&p_h, b.h_arr, cudaIpcMemLazyEnablePeerAccess));
    e->peer.up_h = (double2 *)p_h;
    e->peer.up_pitch = b.pitch;
    e->peer.up_flag = (unsigned long long *)p_f + 0;
    e->peer.opened[3] = p_h; e->peer.opened[4] = p_f;
  } else {                              // lower neighbour: I store E into its high ghost column
    B200_CUDA(cudaIpcOpenMemHandle(&p_e, b.e_arr, cudaIpcMemLazyEnablePeerAccess));
    e->peer.down_e = (double2 *)p_e;
    e->peer.down_pitch = b.pitch;
    e->peer.down_nj = b.nj;
    e->peer.down_flag = (unsigned long long *)p_f + 1;
    e->peer.opened[0] = p_e; e->peer.opened[1] = p_f;
  }
  e->peer.attached[which] = true;
  return B200FDTD_OK;
}

// The same attachment between two engines of ONE process (one host thread driving several
// devices, or several slabs on one device): plain pointers, peer access enabled on demand.
int b200fdtd_peer_attach_engine(b200fdtd_engine *e, int32_t which, b200fdtd_engine *nb)
{
  if (!e || !nb || e == nb || which < 0 || which > 1) return b200_fail(B200FDTD_ERR_ARG, "bad peer argument");
  if (!kind_is_upml(e->g.kind) || e->n_batch > 1 || nb->n_batch > 1)
    return b200_fail(B200FDTD_ERR_ARG, "peer halos serve unbatched engines of the UPML kinds");
  if (nb->rows != e->rows || nb->g.kind != e->g.kind || nb->fp32 != e->fp32)
    return b200_fail(B200FDTD_ERR_ARG, "neighbour slab has another shape, solver kind or precision");
  b200fdtd_engine *both[2] = { e, nb };
  for (int n = 0; n < 2; n++) {
    int rc = select_device(both[n]); if (rc) return rc;
    if (!both[n]->peer.flags) {
      rc = dev_alloc_zero(both[n], (void **)&both[n]->peer.flags, 2 * sizeof(unsigned long long));
      if (rc) return rc;
      B200_CUDA(cudaStreamSynchronize(both[n]->stream));
    }
  }
  int rc = select_device(e); if (rc) return rc;
  if (nb->device != e->device) {
    int can = 0;
    B200_CUDA(cudaDeviceCanAccessPeer(&can, e->device, nb->device));
    if (!can) return b200_fail(B200FDTD_ERR_STATE, "device %d cannot access device %d", e->device, nb->device);
    cudaError_t err = cudaDeviceEnablePeerAccess(nb->device, 0);
    if (err == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
    else if (err != cudaSuccess) return b200_fail(B200FDTD_ERR_CUDA, "enable peer access: %s", cudaGetErrorString(err));
  }
  int es, hs;
  peer_slots(e, &es, &hs);
  if (which == 1) {                     // upper neighbour: I store H into its low ghost column
    e->peer.up_h = nb->field[hs];
    e->peer.up_pitch = nb->pitch;
    e->peer.up_flag = nb->peer.flags + 0;
  } else {                              // lower neighbour: I store E into its high ghost column
    e->peer.down_e = nb->field[es];
    e->peer.down_pitch = nb->pitch;
    e->peer.down_nj = nb->g.nj;
    e->peer.down_flag = nb->peer.flags + 1;
  }
  e->peer.attached[which] = true;
  e->graph_epoch++;
  return B200FDTD_OK;
}

static void peer_release(b200fdtd_engine *e)
{
  for (int n = 0; n < 6; n++)
    if (e->peer.opened[n]) cudaIpcCloseMemHandle(e->peer.opened[n]);
  cudaFree(e->peer.flags);
  memset(&e->peer, 0, sizeof e->peer);
}

int b200fdtd_set_stream(b200fdtd_engine *e, void *cuda_stream)
{
  if (!e) return b200_fail(B200FDTD_ERR_ARG, "NULL engine");
  int rc = select_device(e); if (rc) return rc;
  B200_CUDA(cudaStreamSynchronize(e->stream));
  if (e->own_stream) { cudaStreamDestroy(e->stream); e->own_stream = false; }
  e->stream = (cudaStream_t)cuda_stream;
  e->graph_epoch++;
  return B200FDTD_OK;
}

// Longest run around the middle of [0, n) on which every listed table holds `want`.
static void run_around_centre(const double *tab, int n, const int *slots, const double *want, int n_slots,
                              int32_t *lo, int32_t *hi)
{
  auto ok = [&](int x) {
    for (int s = 0; s < n_slots; s++)
      if (tab[(size_t)slots[s] * n + x] != want[s]) return false;
    return true;
  };
  const int mid = n / 2;
  *lo = 1; *hi = 0;
  if (n < 1 || !ok(mid)) return;
  int a = mid, b = mid;
  while (a > 0 && ok(a - 1)) a--;
  while (b < n - 1 && ok(b + 1)) b++;
  *lo = a; *hi = b;
}

int b200fdtd_upml_interior(int32_t kind, const double *tab_i, int32_t n_px, const double *tab_j, int32_t n_py,
                           int32_t out[4])
{
  if (!tab_i || !tab_j || !out || n_px < 1 || n_py < 1) return b200_fail(B200FDTD_ERR_ARG, "bad argument");
  if (!kind_is_upml(kind)) return b200_fail(B200FDTD_ERR_ARG, "kind %d has no UPML tables", kind);
  const bool tm = kind == B200FDTD_TM_UPML || kind == B200FDTD_MPI_TM_UPML;
  // The two quotient coefficients (C_BYMY0/1, C_DYJY0/1) are num(j) / den(i): 1 where both hold
  // the sigma == 0 value 2*eps0, which the centre of the grid defines.
  const int den_slot = tm ? (int)B200FDTD_TMI_DEN_BYMY : (int)B200FDTD_TEI_DEN_DYJY;
  const double two_eps = tab_i[(size_t)den_slot * n_px + n_px / 2];
  if (tm) {
    const int si[6] = { B200FDTD_TMI_C_JZ, B200FDTD_TMI_C_JZHXHY, B200FDTD_TMI_C_BXMX1, B200FDTD_TMI_C_BXMX0,
                        B200FDTD_TMI_C_BY, B200FDTD_TMI_DEN_BYMY };
    const double wi[6] = { 1, 1, 1, 1, 1, two_eps };
    const int sj[6] = { B200FDTD_TMJ_C_DZ, B200FDTD_TMJ_C_DZJZ, B200FDTD_TMJ_C_MX, B200FDTD_TMJ_C_MXEZ,
                        B200FDTD_TMJ_NUM_BYMY1, B200FDTD_TMJ_NUM_BYMY0 };
    const double wj[6] = { 1, 1, 1, 1, two_eps, two_eps };
    run_around_centre(tab_i, n_px, si, wi, 6, &out[0], &out[1]);
    run_around_centre(tab_j, n_py, sj, wj, 6, &out[2], &out[3]);
  } else {
    const int si[6] = { B200FDTD_TEI_C_DXJX1, B200FDTD_TEI_C_DXJX0, B200FDTD_TEI_C_DY, B200FDTD_TEI_DEN_DYJY,
                        B200FDTD_TEI_C_MZ, B200FDTD_TEI_C_MZEXEY };
    const double wi[6] = { 1, 1, 1, two_eps, 1, 1 };
    const int sj[6] = { B200FDTD_TEJ_C_JX, B200FDTD_TEJ_C_JXHZ, B200FDTD_TEJ_NUM_DYJY1, B200FDTD_TEJ_NUM_DYJY0,
                        B200FDTD_TEJ_C_BZ, B200FDTD_TEJ_C_BZMZ };
    const double wj[6] = { 1, 1, two_eps, two_eps, 1, 1 };
    run_around_centre(tab_i, n_px, si, wi, 6, &out[0], &out[1]);
    run_around_centre(tab_j, n_py, sj, wj, 6, &out[2], &out[3]);
  }
  if (out[0] > out[1] || out[2] > out[3]) { out[0] = out[2] = 1; out[1] = out[3] = 0; }
  return B200FDTD_OK;
}

int b200fdtd_split_geometry(const int32_t updated[4], const int32_t interior[4], int32_t *rects, int32_t *n_rects)
{
  if (!updated || !interior || !rects || !n_rects) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  if (updated[1] < updated[0] || updated[3] < updated[2] || interior[1] < interior[0] || interior[3] < interior[2] ||
      interior[0] < updated[0] || interior[1] > updated[1] || interior[2] < updated[2] || interior[3] > updated[3])
    return b200_fail(B200FDTD_ERR_ARG, "the interior rectangle must be a non-empty part of the updated one");
  int out[5][7], n = 0;
  const int u[4] = { updated[0], updated[1], updated[2], updated[3] };
  const int in[4] = { interior[0], interior[1], interior[2], interior[3] };
  int rc = b200_split_geometry(u, in, out, &n); if (rc) return rc;
  for (int q = 0; q < n; q++)
    for (int m = 0; m < 7; m++) rects[q * 7 + m] = out[q][m];
  *n_rects = n;
  return B200FDTD_OK;
}

int b200fdtd_get_step_form(b200fdtd_engine *e, int32_t *form)
{
  if (!e || !form) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  *form = kind_is_upml(e->g.kind) && e->have_tabs ? b200_step_form(e) : 0;
  if (kind_is_upml(e->g.kind) && e->have_tabs && b200_want_fused(e, nullptr)) *form = (*form == 2) ? 4 : 3;
  return B200FDTD_OK;
}

int b200fdtd_get_lean_extent(b200fdtd_engine *e, int32_t out[4])
{
  if (!e || !out) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  out[0] = out[2] = 1; out[1] = out[3] = 0;
  if (kind_is_upml(e->g.kind) && e->have_tabs && b200_step_form(e) != 0) {
    out[0] = e->lean_r_lo - 1;  out[1] = e->lean_r_hi - 1;
    out[2] = e->lean_c_lo - B200_JOFF + e->g.j0;  out[3] = e->lean_c_hi - B200_JOFF + e->g.j0;
  }
  return B200FDTD_OK;
}

int b200fdtd_set_upml_tables(b200fdtd_engine *e, const double *tab_i, const double *tab_j)
{
  if (!e || !tab_i || !tab_j) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  int rc = select_device(e); if (rc) return rc;
  const b200fdtd_grid &g = e->g;
  {
    // frame-free region of this slab, clipped to the cells the engine updates
    int32_t x[4];
    rc = b200fdtd_upml_interior(g.kind, tab_i, g.n_px, tab_j, g.n_py, x); if (rc) return rc;
    e->lean_r_lo = e->lean_c_lo = 1;
    e->lean_r_hi = e->lean_c_hi = 0;
    if (x[0] <= x[1]) {
      const int jl = x[2] > g.j0 ? x[2] : g.j0, jh = x[3] < g.j0 + g.nj - 1 ? x[3] : g.j0 + g.nj - 1;
      const int r_lo = x[0] + 1 > e->r_lo + 1 ? x[0] + 1 : e->r_lo + 1, r_hi = x[1] + 1 < e->r_hi ? x[1] + 1 : e->r_hi;
      int c_lo = jl - g.j0 + B200_JOFF, c_hi = jh - g.j0 + B200_JOFF;
      // Row r_lo and column c_lo stay with the frame kernels: their low-side neighbour is the ring
      // or a neighbour slab's halo column, which only the H arrays hold (the lean E kernels read B).
      if (c_lo < e->c_lo + 1) c_lo = e->c_lo + 1;
      if (c_hi > e->c_hi) c_hi = e->c_hi;
      if (r_lo <= r_hi && c_lo <= c_hi) {
        e->lean_r_lo = r_lo; e->lean_r_hi = r_hi; e->lean_c_lo = c_lo; e->lean_c_hi = c_hi;
      }
    }
    e->graph_epoch++;
  }
  // ghost rows/columns get the neutral coefficient 1 (never used by an update)
  std::vector<double> hi((size_t)B200FDTD_UPML_TABS * e->rows, 1.0);
  std::vector<double> hj((size_t)B200FDTD_UPML_TABS * e->pitch, 1.0);
  for (int s = 0; s < B200FDTD_UPML_TABS; s++) {
    for (int i = 0; i < g.n_px; i++) hi[(size_t)s * e->rows + i + 1] = tab_i[(size_t)s * g.n_px + i];
    for (int c = 0; c < g.nj; c++)
      hj[(size_t)s * e->pitch + B200_JOFF + c] = tab_j[(size_t)s * g.n_py + g.j0 + c];
  }
  if (e->fp32) {                        // same tables rounded once to float, behind the same pointers
    std::vector<float> fi(hi.begin(), hi.end()), fj(hj.begin(), hj.end());
    B200_CUDA(cudaMemcpyAsync(e->tab_i, fi.data(), fi.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    B200_CUDA(cudaMemcpyAsync(e->tab_j, fj.data(), fj.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    B200_CUDA(cudaStreamSynchronize(e->stream));
    e->have_tabs = true;
    return B200FDTD_OK;
  }
  B200_CUDA(cudaMemcpyAsync(e->tab_i, hi.data(), hi.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  B200_CUDA(cudaMemcpyAsync(e->tab_j, hj.data(), hj.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  B200_CUDA(cudaStreamSynchronize(e->stream));
  e->have_tabs = true;
  return B200FDTD_OK;
}

int b200fdtd_set_split_tables(b200fdtd_engine *e, const double *tab_i, const double *tab_j)
{
  if (!e || !tab_i || !tab_j) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  if (e->g.kind != B200FDTD_TM && e->g.kind != B200FDTD_TE && e->g.kind != B200FDTD_NS_TM)
    return b200_fail(B200FDTD_ERR_ARG, "the lean split-field form serves kinds 0, 1 and 6");
  int rc = select_device(e); if (rc) return rc;
  const b200fdtd_grid &g = e->g;
  std::vector<double> hi((size_t)B200FDTD_SPLIT_TABS * e->rows, 1.0);
  std::vector<double> hj((size_t)B200FDTD_SPLIT_TABS * e->pitch, 1.0);
  for (int s = 0; s < B200FDTD_SPLIT_TABS; s++) {
    for (int i = 0; i < g.n_px; i++) hi[(size_t)s * e->rows + i + 1] = tab_i[(size_t)s * g.n_px + i];
    for (int c = 0; c < g.nj; c++)
      hj[(size_t)s * e->pitch + B200_JOFF + c] = tab_j[(size_t)s * g.n_py + g.j0 + c];
  }
  B200_CUDA(cudaMemcpyAsync(e->tab_i, hi.data(), hi.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  B200_CUDA(cudaMemcpyAsync(e->tab_j, hj.data(), hj.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  B200_CUDA(cudaStreamSynchronize(e->stream));
  e->have_tabs = true;
  e->split_lean = true;
  e->graph_epoch++;
  return B200FDTD_OK;
}

static int upload_eps(b200fdtd_engine *e, int32_t slot, const double *host_eps, size_t ld);

int b200fdtd_set_eps(b200fdtd_engine *e, int32_t slot, const double *host_eps)
{
  if (!e || !host_eps) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  return upload_eps(e, slot, host_eps + e->g.j0, (size_t)e->g.n_py);
}

int b200fdtd_set_eps_slab(b200fdtd_engine *e, int32_t slot, const double *slab_eps)
{
  if (!e || !slab_eps) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  return upload_eps(e, slot, slab_eps, (size_t)e->g.nj);
}

// src points at (i = 0, first owned column); ld = host row stride in doubles
static int upload_eps(b200fdtd_engine *e, int32_t slot, const double *src, size_t ld)
{
  if (slot < 0 || slot > 1 || (!e->eps[slot] && !kind_is_split(e->g.kind)))
    return b200_fail(B200FDTD_ERR_ARG, "bad eps slot %d", slot);
  int rc = select_device(e); if (rc) return rc;
  if (!e->eps[slot]) { rc = dev_alloc_zero(e, (void **)&e->eps[slot], e->plane * e->rsize); if (rc) return rc; }
  const b200fdtd_grid &g = e->g;
  if (e->fp32) {                        // stage the double map on the device, round once to float
    double *stage = nullptr;
    const size_t count = (size_t)g.n_px * g.nj;
    cudaError_t err = cudaMalloc(&stage, count * sizeof(double));
    if (err != cudaSuccess) return b200_fail(B200FDTD_ERR_NOMEM, "eps staging buffer: %s", cudaGetErrorString(err));
    rc = b200_fill_float(e, (float *)e->eps[slot], e->plane, 1.0f);
    if (!rc) {
      cudaError_t c2 = cudaMemcpy2DAsync(stage, sizeof(double) * g.nj, src, sizeof(double) * ld,
                                         sizeof(double) * g.nj, g.n_px, cudaMemcpyHostToDevice, e->stream);
      if (c2 != cudaSuccess) rc = b200_fail(B200FDTD_ERR_CUDA, "eps upload: %s", cudaGetErrorString(c2));
    }
    if (!rc) rc = b200_narrow_real_region(e, stage, (size_t)g.nj, (float *)e->eps[slot]);
    cudaStreamSynchronize(e->stream);
    cudaFree(stage);
    if (!rc) { e->have_eps[slot] = true; e->eps_epoch++; }
    return rc;
  }
  // ghosts and row padding hold vacuum (1.0) so no kernel can ever divide by zero there
  fill_double_kernel<<<1184, 256, 0, e->stream>>>(e->eps[slot], e->plane, 1.0);
  e->launches++;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpy2DAsync(e->eps[slot] + (size_t)e->pitch + B200_JOFF, sizeof(double) * e->pitch,
                              src, sizeof(double) * ld, sizeof(double) * g.nj, g.n_px,
                              cudaMemcpyHostToDevice, e->stream));
  B200_CUDA(cudaStreamSynchronize(e->stream));
  e->have_eps[slot] = true;
  e->eps_epoch++;
  return B200FDTD_OK;
}

int b200fdtd_set_eps_palette(b200fdtd_engine *e, int32_t slot, const uint16_t *index_first, int64_t ld,
                             const double *table, int32_t n_values)
{
  if (!e || !index_first || !table || n_values < 1 || n_values > 65536 || ld < e->g.nj)
    return b200_fail(B200FDTD_ERR_ARG, "bad palette (%d values)", n_values);
  if (slot < 0 || slot > 1 || (!e->eps[slot] && !kind_is_split(e->g.kind)))
    return b200_fail(B200FDTD_ERR_ARG, "bad eps slot %d", slot);
  int rc = select_device(e); if (rc) return rc;
  if (!e->eps[slot]) { rc = dev_alloc_zero(e, (void **)&e->eps[slot], e->plane * e->rsize); if (rc) return rc; }
  const b200fdtd_grid &g = e->g;
  // the indices travel in row blocks through a small staging buffer (a 512 MB cudaMalloc / cudaFree
  // pair costs more than the copy it would serve)
  const int rows_per_chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)g.n_px, ((size_t)16 << 20) / (size_t)g.nj));
  unsigned short *d_index = nullptr;
  double *d_table = nullptr;
  cudaError_t err = cudaMalloc((void **)&d_index, (size_t)rows_per_chunk * g.nj * sizeof(unsigned short));
  if (err == cudaSuccess) err = cudaMalloc((void **)&d_table, sizeof(double) * (size_t)n_values);
  if (err != cudaSuccess) { cudaFree(d_index); return b200_fail(B200FDTD_ERR_NOMEM, "palette staging: %s", cudaGetErrorString(err)); }
  if (e->fp32) rc = b200_fill_float(e, (float *)e->eps[slot], e->plane, 1.0f);
  else { fill_double_kernel<<<1184, 256, 0, e->stream>>>(e->eps[slot], e->plane, 1.0); e->launches++; }
  if (!rc) {
    err = cudaMemcpyAsync(d_table, table, sizeof(double) * (size_t)n_values, cudaMemcpyHostToDevice, e->stream);
    for (int i0 = 0; i0 < g.n_px && err == cudaSuccess; i0 += rows_per_chunk) {
      const int n_rows = std::min(rows_per_chunk, g.n_px - i0);
      err = cudaMemcpy2DAsync(d_index, sizeof(unsigned short) * g.nj, index_first + (size_t)i0 * (size_t)ld,
                              sizeof(unsigned short) * (size_t)ld, sizeof(unsigned short) * g.nj, n_rows,
                              cudaMemcpyHostToDevice, e->stream);
      if (err != cudaSuccess) break;
      if (e->fp32) expand_palette_kernel<float><<<592, 256, 0, e->stream>>>(d_index, d_table, n_values, (float *)e->eps[slot], e->pitch, i0, n_rows, g.nj);
      else         expand_palette_kernel<double><<<592, 256, 0, e->stream>>>(d_index, d_table, n_values, e->eps[slot], e->pitch, i0, n_rows, g.nj);
      e->launches++;
      err = cudaGetLastError();
    }
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) rc = b200_fail(B200FDTD_ERR_CUDA, "palette upload: %s", cudaGetErrorString(err));
  }
  cudaFree(d_index); cudaFree(d_table);
  if (!rc) { e->have_eps[slot] = true; e->eps_epoch++; }
  return rc;
}

int b200fdtd_set_dense(b200fdtd_engine *e, int32_t slot, const double *host_map)
{
  if (!e || !host_map || slot < 0 || slot >= B200FDTD_MAX_DENSE || !kind_is_split(e->g.kind))
    return b200_fail(B200FDTD_ERR_ARG, "bad dense slot %d for kind %d", slot, e ? e->g.kind : -1);
  int rc = select_device(e); if (rc) return rc;
  if (!e->dense[slot]) { rc = dev_alloc_zero(e, (void **)&e->dense[slot], e->plane * sizeof(double)); if (rc) return rc; }
  const b200fdtd_grid &g = e->g;
  B200_CUDA(cudaMemcpy2DAsync(e->dense[slot] + (size_t)e->pitch + B200_JOFF, sizeof(double) * e->pitch,
                              host_map + g.j0, sizeof(double) * g.n_py, sizeof(double) * g.nj, g.n_px,
                              cudaMemcpyHostToDevice, e->stream));
  B200_CUDA(cudaStreamSynchronize(e->stream));
  e->have_dense[slot] = true;
  e->graph_epoch++;                     // the first upload of a slot allocates it: captured pointers are stale
  return B200FDTD_OK;
}

int b200fdtd_set_split_interior(b200fdtd_engine *e, int32_t i_lo, int32_t i_hi, int32_t j_lo, int32_t j_hi)
{
  if (!e) return b200_fail(B200FDTD_ERR_ARG, "NULL engine");
  if (e->g.kind != B200FDTD_NS_TE) return b200_fail(B200FDTD_ERR_ARG, "the interior form serves the NS-FDTD TE kind (7)");
  e->graph_epoch++;
  e->split_in_r_lo = e->split_in_c_lo = 1;
  e->split_in_r_hi = e->split_in_c_hi = 0;
  if (i_lo > i_hi || j_lo > j_hi) return B200FDTD_OK;             // switched off
  const b200fdtd_grid &g = e->g;
  const int jl = j_lo > g.j0 ? j_lo : g.j0, jh = j_hi < g.j0 + g.nj - 1 ? j_hi : g.j0 + g.nj - 1;
  int r_lo = i_lo + 1, r_hi = i_hi + 1, c_lo = jl - g.j0 + B200_JOFF, c_hi = jh - g.j0 + B200_JOFF;
  if (r_lo < e->r_lo) r_lo = e->r_lo;
  if (r_hi > e->r_hi) r_hi = e->r_hi;
  if (c_lo < e->c_lo) c_lo = e->c_lo;
  if (c_hi > e->c_hi) c_hi = e->c_hi;
  if (r_lo <= r_hi && c_lo <= c_hi) {
    e->split_in_r_lo = r_lo; e->split_in_r_hi = r_hi; e->split_in_c_lo = c_lo; e->split_in_c_hi = c_hi;
  }
  return B200FDTD_OK;
}

int b200fdtd_set_batch_sources(b200fdtd_engine *e, const b200fdtd_batch_source *sources)
{
  if (!e || !sources) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  if (!e->batch_src) return b200_fail(B200FDTD_ERR_ARG, "batch sources serve the serial UPML kinds (2, 3)");
  int rc = select_device(e); if (rc) return rc;
  B200_CUDA(cudaMemcpyAsync(e->batch_src, sources, sizeof(b200fdtd_batch_source) * (size_t)e->n_batch,
                            cudaMemcpyHostToDevice, e->stream));
  B200_CUDA(cudaStreamSynchronize(e->stream));
  e->have_batch_src = true;
  return B200FDTD_OK;
}

int b200fdtd_set_batch_cw(b200fdtd_engine *e, const b200fdtd_batch_cw *sources)
{
  if (!e || !sources) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  if (!e->batch_cw) return b200_fail(B200FDTD_ERR_ARG, "batch CW records serve batched split-field engines");
  int rc = select_device(e); if (rc) return rc;
  B200_CUDA(cudaMemcpyAsync(e->batch_cw, sources, sizeof(b200fdtd_batch_cw) * (size_t)e->n_batch,
                            cudaMemcpyHostToDevice, e->stream));
  B200_CUDA(cudaStreamSynchronize(e->stream));
  e->have_batch_cw = true;
  return B200FDTD_OK;
}

int b200fdtd_select_batch(b200fdtd_engine *e, int32_t index)
{
  if (!e || index < 0 || index >= e->n_batch) return b200_fail(B200FDTD_ERR_ARG, "batch index %d out of range", index);
  e->sel = index;
  return B200FDTD_OK;
}

int b200fdtd_set_ntff_plan(b200fdtd_engine *e, const b200fdtd_ntff_plan *p)
{
  if (!e || !p || (!p->time_shift && p->n_local > 0)) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  int rc = select_device(e); if (rc) return rc;
  const b200fdtd_grid &g = e->g;
  const int nx = p->right - p->left, ny = p->top - p->bottom;
  if (nx <= 0 || ny <= 0 || p->n_points != 2 * nx + 2 * ny || p->max_time < 1 ||
      p->n_bins < 1 || p->n_angles < 1 || p->left + p->sample_di < 0 || p->bottom + p->sample_dj < 0 ||
      p->left < 1 || p->bottom < 1 || p->right >= g.n_px || p->top >= g.n_py ||
      p->right + p->sample_di >= g.n_px || p->top + p->sample_dj >= g.n_py)
    return b200_fail(B200FDTD_ERR_ARG, "inconsistent NTFF plan (box %d..%d x %d..%d, %d points)",
                     p->left, p->right, p->bottom, p->top, p->n_points);
  B200_CUDA(cudaStreamSynchronize(e->stream));
  free_ntff(e);
  NtffState &n = e->ntff;
  n.top = p->top; n.bottom = p->bottom; n.left = p->left; n.right = p->right;
  n.n_points_global = p->n_points;
  n.max_time = p->max_time; n.n_bins = p->n_bins; n.n_angles = p->n_angles;
  n.array_size = p->array_size;
  n.tap_scale = p->tap_scale != 0.0 ? p->tap_scale : 1.0;

  // perimeter in the reference's loop order, keeping the points this slab owns; the cell
  // actually read is (i + sample_di, j + sample_dj) (non-zero for the MPI-variant ids only)
  std::vector<NtffPoint> pts;
  const int di = p->sample_di, dj = p->sample_dj;
  auto push = [&](int i, int j, int edge, int pg) {
    i += di; j += dj;
    if (j < g.j0 || j >= g.j0 + g.nj) return;
    NtffPoint q;
    q.k = (long long)(i + 1) * e->pitch + (j - g.j0) + B200_JOFF;
    q.edge = edge;
    q.p_global = pg;
    pts.push_back(q);
  };
  int pg = 0;
  for (int i = p->left; i < p->right; i++) push(i, p->bottom, 0, pg++);
  for (int j = p->bottom; j < p->top; j++) push(p->right, j, 1, pg++);
  for (int i = p->left; i < p->right; i++) push(i, p->top, 2, pg++);
  for (int j = p->bottom; j < p->top; j++) push(p->left, j, 3, pg++);
  n.n_local = (int)pts.size();

  if (n.n_local != p->n_local) {
    const int have = n.n_local;
    memset(&n, 0, sizeof n);
    return b200_fail(B200FDTD_ERR_ARG, "NTFF plan lists %d local points, slab [%d,+%d) owns %d",
                     p->n_local, g.j0, g.nj, have);
  }

  rc = dev_alloc_zero(e, (void **)&n.pts, sizeof(NtffPoint) * (size_t)(n.n_local ? n.n_local : 1));
  const size_t ts_count = (size_t)n.n_angles * (size_t)(n.n_local ? n.n_local : 1);
  if (!rc) rc = dev_alloc_zero(e, (void **)&n.ts, sizeof(double) * ts_count);
  const size_t nb = (size_t)e->n_batch;
  if (!rc) rc = dev_alloc_zero(e, (void **)&n.hist_e, sizeof(double2) * (size_t)n.n_local * n.max_time * nb);
  if (!rc) rc = dev_alloc_zero(e, (void **)&n.hist_h, sizeof(double2) * (size_t)n.n_local * n.max_time * nb);
  if (!rc) rc = dev_alloc_zero(e, (void **)&n.uw, sizeof(double2) * 3 * (size_t)n.n_angles * n.n_bins * nb);
  if (rc) { free_ntff(e); return rc; }
  if (n.n_local) {
    B200_CUDA(cudaMemcpyAsync(n.pts, pts.data(), sizeof(NtffPoint) * pts.size(), cudaMemcpyHostToDevice, e->stream));
    B200_CUDA(cudaMemcpyAsync(n.ts, p->time_shift, sizeof(double) * (size_t)n.n_angles * n.n_local,
                              cudaMemcpyHostToDevice, e->stream));
  }
  B200_CUDA(cudaStreamSynchronize(e->stream));
  n.ready = true;
  n.steps_recorded = 0;
  e->graph_epoch++;
  e->ntff_epoch++;
  return B200FDTD_OK;
}

static int check_ready(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  if (!e || !a) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  if (e->n_fields == 0) return b200_fail(B200FDTD_ERR_STATE, "an NTFF-only engine has no fields to step");
  if (kind_is_split(e->g.kind)) {
    if (e->n_batch > 1 && !e->have_batch_cw)
      return b200_fail(B200FDTD_ERR_STATE, "step of a batched engine before set_batch_cw");
    if (e->split_lean) {            // 1-D tables + eps (kinds 0, 1) or G arrays + source factor (kind 6)
      const int k = e->g.kind;
      if (k == B200FDTD_NS_TM) {
        const int need[4] = { B200FDTD_STM_C_EZXLX, B200FDTD_STM_C_HXLY, B200FDTD_STM_C_HYLX, B200FDTD_DENSE_SRC0 };
        for (int s = 0; s < 4; s++)
          if (!e->have_dense[need[s]]) return b200_fail(B200FDTD_ERR_STATE, "step before set_dense(%d)", need[s]);
      } else if (!e->have_eps[0] || (k == B200FDTD_TE && !e->have_eps[1])) {
        return b200_fail(B200FDTD_ERR_STATE, "step before set_eps");
      }
      return select_device(e);
    }
    for (int s = 0; s < 8; s++)
      if (!e->have_dense[s]) return b200_fail(B200FDTD_ERR_STATE, "step before set_dense(%d)", s);
    int rc = select_device(e);
    for (int s = B200FDTD_DENSE_SRC0; s <= B200FDTD_DENSE_SRC1 && !rc; s++)     // no source factor uploaded: zeros
      if (!e->dense[s]) rc = dev_alloc_zero(e, (void **)&e->dense[s], e->plane * sizeof(double));
    return rc;
  }
  if (!e->have_tabs) return b200_fail(B200FDTD_ERR_STATE, "step before set_upml_tables");
  if (!e->have_eps[0] || (e->eps[1] && !e->have_eps[1]))
    return b200_fail(B200FDTD_ERR_STATE, "step before set_eps");
  if (e->n_batch > 1 && !e->have_batch_src)
    return b200_fail(B200FDTD_ERR_STATE, "step of a batched engine before set_batch_sources");
  return select_device(e);
}

int b200fdtd_phase_h(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  int rc = check_ready(e, a); if (rc) return rc;
  if (kind_is_split(e->g.kind)) return b200_fail(B200FDTD_ERR_ARG, "phase API serves the UPML kinds only");
  const unsigned long long step = (unsigned long long)a->time;
  // my high ghost column (Ez / Ex) was written by the upper neighbour's E phase of step-1
  if (e->peer.attached[1] && step > 0) { rc = b200_peer_wait(e, 1, step); if (rc) return rc; }
  rc = b200_launch_upml_h(e, a);                // sets h_stale = !store_h; stores the halo column upward
  if (!rc && e->peer.attached[1]) rc = b200_peer_signal(e, e->peer.up_flag, step + 1);
  return rc;
}

int b200fdtd_phase_e(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  int rc = check_ready(e, a); if (rc) return rc;
  if (kind_is_split(e->g.kind)) return b200_fail(B200FDTD_ERR_ARG, "phase API serves the UPML kinds only");
  const unsigned long long step = (unsigned long long)a->time;
  // my low ghost column (Hx / Hz) is written by the lower neighbour's H phase of this step
  if (e->peer.attached[0]) { rc = b200_peer_wait(e, 0, step + 1); if (rc) return rc; }
  rc = b200_launch_upml_e(e, a);                // reads B/mu0 when the H arrays are stale; halo column downward
  // Once told, the lower neighbour may run its next H phase and overwrite my low ghost column of H,
  // which the NTFF sample of THIS step still reads for surface points on my first owned column: with
  // an NTFF plan the signal is left to b200fdtd_phase_sample, which completes the step.
  if (!rc && e->peer.attached[0]) {
    if (e->ntff.ready) e->peer.pending_down = step + 1;
    else rc = b200_peer_signal(e, e->peer.down_flag, step + 1);
  }
  return rc;
}

int b200fdtd_set_option(b200fdtd_engine *e, int32_t option, int32_t value)
{
  if (!e) return b200_fail(B200FDTD_ERR_ARG, "NULL engine");
  int rc = select_device(e); if (rc) return rc;
  e->graph_epoch++;
  switch (option) {
  case B200FDTD_OPT_FUSED:
    if (value == 1 && ((e->g.kind != B200FDTD_TM_UPML && e->g.kind != B200FDTD_TE_UPML) || e->fp32 || e->n_batch > 1))
      return b200_fail(B200FDTD_ERR_ARG, "the one-pass step serves the serial UPML kinds (2, 3), unbatched, in double "
                                         "precision");
    e->use_fused = value == 1;
    e->fused_auto = value == 2;
    return B200FDTD_OK;
  case B200FDTD_OPT_STORE_H:
    rc = b200_refresh_h(e); if (rc) return rc;
    e->store_h = value != 0;
    return B200FDTD_OK;
  case B200FDTD_OPT_BAND_ROWS:
    if (value < 1) return b200_fail(B200FDTD_ERR_ARG, "band rows must be >= 1");
    B200_CUDA(cudaStreamSynchronize(e->stream));
    b200_fused_release(e);
    e->fused.band_h = value;
    return B200FDTD_OK;
  case B200FDTD_OPT_FUSED_SHAPE:
    if (value < 20 || value > 24) return b200_fail(B200FDTD_ERR_ARG, "one-pass launch shapes are 20..24");
    B200_CUDA(cudaStreamSynchronize(e->stream));
    e->fused_variant = value;
    return B200FDTD_OK;
  case B200FDTD_OPT_F32_PAIRS:
    e->f32_pairs = value != 0;
    return B200FDTD_OK;
  case B200FDTD_OPT_UNIT_SPLIT:
    if (value < 0 || value > 2) return b200_fail(B200FDTD_ERR_ARG, "unit split: 0 off, 1 on, 2 auto");
    e->unit_split = value;
    return B200FDTD_OK;
  case B200FDTD_OPT_DERIVED_E:
    rc = b200_refresh_e(e); if (rc) return rc;
    B200_CUDA(cudaStreamSynchronize(e->stream));
    e->derived_e = value != 0;
    e->fused.vac_built = false;
    return B200FDTD_OK;
  case B200FDTD_OPT_LEAN_INTERIOR:
    if (value && !kind_is_upml(e->g.kind))
      return b200_fail(B200FDTD_ERR_ARG, "the lean interior form serves the UPML kinds");
    e->lean_interior = value != 0;
    return B200FDTD_OK;
  default:
    return b200_fail(B200FDTD_ERR_ARG, "unknown option %d", option);
  }
}

int b200fdtd_phase_fused(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  int rc = check_ready(e, a); if (rc) return rc;
  if (!b200_want_fused(e, a)) return b200_fail(B200FDTD_ERR_STATE, "the one-pass step does not serve this engine / source");
  return b200_launch_upml_fused(e, a);
}

int b200fdtd_phase_sample(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  int rc = check_ready(e, a); if (rc) return rc;
  rc = b200_launch_ntff_sample(e, a);
  if (!rc && e->peer.pending_down) {            // the E-phase signal b200fdtd_phase_e held back
    rc = b200_peer_signal(e, e->peer.down_flag, e->peer.pending_down);
    e->peer.pending_down = 0;
  }
  return rc;
}

int b200fdtd_step(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  int rc = check_ready(e, a); if (rc) return rc;
  if (kind_is_split(e->g.kind)) return b200_launch_split_step(e, a);
  const bool e_first = (e->g.kind == B200FDTD_MPI_TM_UPML || e->g.kind == B200FDTD_MPI_TE_UPML);
  const bool peers = e->peer.attached[0] || e->peer.attached[1];
  if (peers && e_first) {
    // mpiTM_UPML.c:196-217 on a y-slab: E, source, [Ez/Ex column down], H, [Hx/Hz column up], NTFF --
    // the reference's two MPI_Sendrecv exchanges as peer stores behind the same two flags.
    const unsigned long long step = (unsigned long long)a->time;
    // my low ghost column of H: the lower neighbour's H phase of step-1.  Its signal also says that
    // phase has finished reading the high ghost column my E phase is about to overwrite.
    if (e->peer.attached[0] && step > 0) { rc = b200_peer_wait(e, 0, step); if (rc) return rc; }
    rc = b200_launch_upml_e(e, a);
    if (!rc && e->peer.attached[0]) rc = b200_peer_signal(e, e->peer.down_flag, step + 1);
    // my high ghost column of E: the upper neighbour's E phase of this step (which has also finished
    // reading the low ghost column my H phase overwrites, as has its NTFF sample of step-1)
    if (!rc && e->peer.attached[1]) rc = b200_peer_wait(e, 1, step + 1);
    if (!rc) rc = b200_launch_upml_h(e, a);
    if (!rc && e->peer.attached[1]) rc = b200_peer_signal(e, e->peer.up_flag, step + 1);
    // the surface sample averages H over columns j-1 and j: on my first owned column that is the
    // lower neighbour's H of THIS step
    if (!rc && e->ntff.ready && e->peer.attached[0]) rc = b200_peer_wait(e, 0, step + 1);
    if (!rc) rc = b200_launch_ntff_sample(e, a);
    return rc;
  }
  if (peers && !b200_want_fused(e, a)) {
    // two-kernel forms on a slab with neighbours: the phase entry points carry the flag protocol
    // (wait for the neighbour's halo, launch, signal) -- never a bare launch next to a peer store
    rc = b200fdtd_phase_h(e, a);
    if (!rc) rc = b200fdtd_phase_e(e, a);
    if (!rc) rc = b200fdtd_phase_sample(e, a);
    return rc;
  }
  if (b200_want_fused(e, a)) {
    // H and E in one pass.  On a y-slab with peer halos: my last column's Hx goes up first (edge
    // kernel; it also saves the old ghost Ez the pass needs), the pass waits for the lower
    // neighbour's, and stores my first column's Ez downward itself.
    const unsigned long long step = (unsigned long long)a->time;
    if (e->peer.attached[1]) {
      if (step > 0) rc = b200_peer_wait(e, 1, step);        // upper neighbour's Ez of step-1 is in my ghost column
      if (!rc) rc = b200_launch_fused_edge(e, a);
      if (!rc) rc = b200_peer_signal(e, e->peer.up_flag, step + 1);
    }
    if (!rc && e->peer.attached[0]) rc = b200_peer_wait(e, 0, step + 1);   // lower neighbour's Hx of this step
    if (!rc) rc = b200_launch_upml_fused(e, a);
    // (the surface sample goes before the signal: see b200fdtd_phase_e)
    if (!rc && e->peer.attached[0]) {
      rc = b200_launch_ntff_sample(e, a);
      if (!rc) rc = b200_peer_signal(e, e->peer.down_flag, step + 1);
      return rc;
    }
  } else if (e_first) {              // mpiTM_UPML.c:196-217: E, source, H, NTFF
    rc = b200_launch_upml_e(e, a);
    if (!rc) rc = b200_launch_upml_h(e, a);
  } else {                    // fdtdTM_upml.c:54-66: H, E, source, NTFF
    rc = b200_launch_upml_h(e, a);
    if (!rc) rc = b200_launch_upml_e(e, a);
  }
  if (!rc) rc = b200_launch_ntff_sample(e, a);
  return rc;
}

// One update() with the time taken from the device clock (capturable: nothing in the launch
// sequence depends on the step number).
static int launch_clocked_step(b200fdtd_engine *e, const b200fdtd_step_args *a)
{
  int rc;
  if (b200_want_fused(e, a)) {
    rc = b200_launch_upml_fused(e, a);
  } else {
    rc = b200_launch_upml_h(e, a);
    if (!rc) rc = b200_launch_upml_e(e, a);
  }
  if (!rc) rc = b200_launch_ntff_sample(e, a);
  if (!rc) rc = b200_launch_clock_advance(e);
  return rc;
}

int b200fdtd_run_steps(b200fdtd_engine *e, double time0, int32_t n_steps)
{
  if (!e || n_steps < 0) return b200_fail(B200FDTD_ERR_ARG, "bad argument");
  if (e->g.kind != B200FDTD_TM_UPML && e->g.kind != B200FDTD_TE_UPML)
    return b200_fail(B200FDTD_ERR_ARG, "multi-step replay serves the serial UPML kinds (2, 3)");
  if (!e->have_batch_src) return b200_fail(B200FDTD_ERR_STATE, "run_steps before set_batch_sources");
  if (e->peer.attached[0] || e->peer.attached[1])
    return b200_fail(B200FDTD_ERR_ARG, "multi-step replay: no peer halos");
  b200fdtd_step_args a;
  memset(&a, 0, sizeof a);
  a.time = time0;
  int rc = check_ready(e, &a); if (rc) return rc;
  if (n_steps == 0) return B200FDTD_OK;
  if (b200_want_fused(e, &a)) {       // side buffers are allocated outside any stream capture
    rc = b200_fused_prepare(e); if (rc) return rc;
  }
  NtffState &n = e->ntff;
  if (n.ready && ((int)time0 < 0 || (int)time0 + n_steps > n.max_time))
    return b200_fail(B200FDTD_ERR_ARG, "steps %d..%d outside the NTFF history [0, %d)", (int)time0,
                     (int)time0 + n_steps - 1, n.max_time);
  B200_CUDA(cudaMemcpyAsync(e->clock_dev, &time0, sizeof(double), cudaMemcpyHostToDevice, e->stream));
  B200_CUDA(cudaStreamSynchronize(e->stream));          // time0 is a stack variable

  const int kChunk = 128;
  e->clock_mode = true;
  int done = 0;
  while (done < n_steps && !rc) {
    if (!e->e_consistent) {
      // after b200fdtd_set_field: one step outside any graph restores E == D/eps, which decides what the
      // one-pass kernels of the captured steps read (fused_kernels.cu, vacuum row-strips)
      rc = launch_clocked_step(e, &a);
      done++;
      continue;
    }
    const int chunk = n_steps - done < kChunk ? n_steps - done : kChunk;
    cudaGraphExec_t exec = (cudaGraphExec_t)e->graph_exec;
    if (exec == nullptr || e->graph_steps != chunk || e->graph_built_epoch != e->graph_epoch) {
      if (exec) { cudaGraphExecDestroy(exec); e->graph_exec = nullptr; }
      // a chunk shorter than 8 steps is not worth a graph
      if (chunk < 8) {
        for (int s = 0; s < chunk && !rc; s++) rc = launch_clocked_step(e, &a);
        done += chunk;
        continue;
      }
      const uint64_t launches_before = e->launches;
      cudaGraph_t graph = nullptr;
      cudaError_t err = cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal);
      if (err != cudaSuccess) { rc = b200_fail(B200FDTD_ERR_CUDA, "begin capture: %s", cudaGetErrorString(err)); break; }
      for (int s = 0; s < chunk && !rc; s++) rc = launch_clocked_step(e, &a);
      err = cudaStreamEndCapture(e->stream, &graph);
      e->graph_launches = e->launches - launches_before;   // kernels in the chunk
      e->launches = launches_before;                    // captured, not launched yet
      if (rc) { if (graph) cudaGraphDestroy(graph); break; }
      if (err != cudaSuccess) { rc = b200_fail(B200FDTD_ERR_CUDA, "end capture: %s", cudaGetErrorString(err)); break; }
      err = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (err != cudaSuccess) { rc = b200_fail(B200FDTD_ERR_CUDA, "graph instantiate: %s", cudaGetErrorString(err)); break; }
      e->graph_exec = exec;
      e->graph_steps = chunk;
      e->graph_built_epoch = e->graph_epoch;
    }
    cudaError_t err = cudaGraphLaunch((cudaGraphExec_t)e->graph_exec, e->stream);
    if (err != cudaSuccess) { rc = b200_fail(B200FDTD_ERR_CUDA, "graph launch: %s", cudaGetErrorString(err)); break; }
    e->launches += e->graph_launches;
    done += chunk;
  }
  e->clock_mode = false;
  // a replayed graph runs none of the launch functions that keep these flags: a getter between two replays
  // of the same graph has refreshed the E arrays, and the second replay leaves them behind again
  if (done > 0 && b200_fused_derives_e(e)) e->e_stale = true;
  if (!rc) {
    e->h_stale = !e->store_h;
    if (n.ready && (int)time0 + n_steps > n.steps_recorded) n.steps_recorded = (int)time0 + n_steps;
  }
  return rc;
}

// Split-field kinds: a chunk of steps from one graph, each kernel reading its own step's CW records
// from a device table refreshed by one copy per chunk.
int b200fdtd_run_split_steps(b200fdtd_engine *e, const b200fdtd_step_args *args, int32_t n_steps)
{
  if (!e || !args || n_steps < 0) return b200_fail(B200FDTD_ERR_ARG, "bad argument");
  if (!kind_is_split(e->g.kind)) return b200_fail(B200FDTD_ERR_ARG, "run_split_steps serves the split-field kinds (0, 1, 6, 7)");
  int rc = check_ready(e, args); if (rc) return rc;
  if (n_steps == 0) return B200FDTD_OK;
  const int kChunk = 128;
  if (!e->cw_tab) {
    rc = dev_alloc_zero(e, (void **)&e->cw_tab, sizeof(b200fdtd_cw) * 2 * kChunk);
    if (rc) return rc;
    for (int b = 0; b < 2; b++) {
      B200_CUDA(cudaMallocHost((void **)&e->cw_stage[b], sizeof(b200fdtd_cw) * 2 * kChunk));
      cudaEvent_t ev;
      B200_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      e->cw_stage_done[b] = ev;
    }
  }
  int done = 0, turn = 0;
  while (done < n_steps && !rc) {
    const int chunk = n_steps - done < kChunk ? n_steps - done : kChunk;
    if (chunk < 8) {                    // not worth a graph
      for (int s = 0; s < chunk && !rc; s++) rc = b200_launch_split_step(e, &args[done + s]);
      done += chunk;
      continue;
    }
    cudaGraphExec_t exec = (cudaGraphExec_t)e->graph_exec;
    if (exec == nullptr || e->graph_steps != chunk || e->graph_built_epoch != e->graph_epoch) {
      if (exec) { cudaGraphExecDestroy(exec); e->graph_exec = nullptr; }
      const uint64_t launches_before = e->launches;
      cudaGraph_t graph = nullptr;
      cudaError_t err = cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal);
      if (err != cudaSuccess) { rc = b200_fail(B200FDTD_ERR_CUDA, "begin capture: %s", cudaGetErrorString(err)); break; }
      for (int s = 0; s < chunk && !rc; s++) {
        e->split_cw_step = e->cw_tab + 2 * s;
        rc = b200_launch_split_step(e, &args[0]);          // ns_r2; the CW records come from the table
      }
      e->split_cw_step = nullptr;
      err = cudaStreamEndCapture(e->stream, &graph);
      e->graph_launches = e->launches - launches_before;
      e->launches = launches_before;
      if (rc) { if (graph) cudaGraphDestroy(graph); break; }
      if (err != cudaSuccess) { rc = b200_fail(B200FDTD_ERR_CUDA, "end capture: %s", cudaGetErrorString(err)); break; }
      err = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (err != cudaSuccess) { rc = b200_fail(B200FDTD_ERR_CUDA, "graph instantiate: %s", cudaGetErrorString(err)); break; }
      e->graph_exec = exec;
      e->graph_steps = chunk;
      e->graph_built_epoch = e->graph_epoch;
    }
    // this chunk's records: pinned twin (once its previous copy has run) -> device table, in stream order
    // after the previous chunk's kernels
    B200_CUDA(cudaEventSynchronize((cudaEvent_t)e->cw_stage_done[turn]));
    for (int s = 0; s < chunk; s++) {
      e->cw_stage[turn][2 * s] = args[done + s].cw[0];
      e->cw_stage[turn][2 * s + 1] = args[done + s].cw[1];
    }
    B200_CUDA(cudaMemcpyAsync(e->cw_tab, e->cw_stage[turn], sizeof(b200fdtd_cw) * 2 * (size_t)chunk,
                              cudaMemcpyHostToDevice, e->stream));
    B200_CUDA(cudaEventRecord((cudaEvent_t)e->cw_stage_done[turn], e->stream));
    cudaError_t err = cudaGraphLaunch((cudaGraphExec_t)e->graph_exec, e->stream);
    if (err != cudaSuccess) { rc = b200_fail(B200FDTD_ERR_CUDA, "graph launch: %s", cudaGetErrorString(err)); break; }
    e->launches += e->graph_launches;
    done += chunk;
    turn ^= 1;
  }
  return rc;
}

int b200fdtd_sync(b200fdtd_engine *e)
{
  if (!e) return b200_fail(B200FDTD_ERR_ARG, "NULL engine");
  int rc = select_device(e); if (rc) return rc;
  B200_CUDA(cudaStreamSynchronize(e->stream));
  return B200FDTD_OK;
}

int b200fdtd_halo_pack(b200fdtd_engine *e, int32_t which, void *dev_buf)
{
  if (!e || !dev_buf || which < 0 || which > 1 || e->n_batch > 1) return b200_fail(B200FDTD_ERR_ARG, "bad halo argument");
  int rc = select_device(e); if (rc) return rc;
  return b200_launch_halo(e, which, dev_buf, true);
}

int b200fdtd_halo_unpack(b200fdtd_engine *e, int32_t which, const void *dev_buf)
{
  if (!e || !dev_buf || which < 0 || which > 1) return b200_fail(B200FDTD_ERR_ARG, "bad halo argument");
  int rc = select_device(e); if (rc) return rc;
  return b200_launch_halo(e, which, const_cast<void *>(dev_buf), false);
}

// Device plane of a field as double2: the array itself, or (single-precision engines) a
// widened temporary the caller frees after its copy has completed.
static int field_plane_f64(b200fdtd_engine *e, int slot, const double2 **plane, double2 **temp)
{
  *temp = nullptr;
  *plane = e->field[slot] + (size_t)e->sel * e->plane;
  if (!e->fp32) return B200FDTD_OK;
  cudaError_t err = cudaMalloc(temp, e->plane * sizeof(double2));
  if (err != cudaSuccess) return b200_fail(B200FDTD_ERR_NOMEM, "getter staging plane: %s", cudaGetErrorString(err));
  int rc = b200_widen_plane(e, (const float2 *)e->field[slot] + (size_t)e->sel * e->plane, *temp, e->plane);
  if (rc) { cudaFree(*temp); *temp = nullptr; return rc; }
  *plane = *temp;
  return B200FDTD_OK;
}

static int copy_field_out(b200fdtd_engine *e, int slot, double *host_first, size_t host_ld_complex)
{
  int rc = b200_refresh_h(e); if (rc) return rc;
  rc = b200_refresh_e(e); if (rc) return rc;
  const double2 *plane; double2 *temp;
  rc = field_plane_f64(e, slot, &plane, &temp); if (rc) return rc;
  const b200fdtd_grid &g = e->g;
  cudaError_t err = cudaMemcpy2DAsync(host_first, sizeof(double2) * host_ld_complex,
                                      plane + (size_t)e->pitch + B200_JOFF, sizeof(double2) * e->pitch,
                                      sizeof(double2) * g.nj, g.n_px, cudaMemcpyDeviceToHost, e->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  cudaFree(temp);
  if (err != cudaSuccess) return b200_fail(B200FDTD_ERR_CUDA, "field download: %s", cudaGetErrorString(err));
  return B200FDTD_OK;
}

int b200fdtd_get_field(b200fdtd_engine *e, int32_t slot, double *host)
{
  if (!e || !host || slot < 0 || slot >= e->n_fields) return b200_fail(B200FDTD_ERR_ARG, "bad field slot %d", slot);
  int rc = select_device(e); if (rc) return rc;
  return copy_field_out(e, slot, host + 2 * (size_t)e->g.j0, (size_t)e->g.n_py);
}

int b200fdtd_get_field_ld(b200fdtd_engine *e, int32_t slot, double *host, int64_t ld)
{
  if (!e || !host || slot < 0 || slot >= e->n_fields || ld < e->g.nj)
    return b200_fail(B200FDTD_ERR_ARG, "bad field slot %d / leading dimension", slot);
  int rc = select_device(e); if (rc) return rc;
  return copy_field_out(e, slot, host, (size_t)ld);
}

int b200fdtd_get_field_slab(b200fdtd_engine *e, int32_t slot, double *host)
{
  if (!e || !host || slot < 0 || slot >= e->n_fields) return b200_fail(B200FDTD_ERR_ARG, "bad field slot %d", slot);
  int rc = select_device(e); if (rc) return rc;
  return copy_field_out(e, slot, host, (size_t)e->g.nj);
}

int b200fdtd_set_field(b200fdtd_engine *e, int32_t slot, const double *host)
{
  if (!e || !host || slot < 0 || slot >= e->n_fields) return b200_fail(B200FDTD_ERR_ARG, "bad field slot %d", slot);
  int rc = select_device(e); if (rc) return rc;
  const b200fdtd_grid &g = e->g;
  rc = b200_refresh_e(e); if (rc) return rc;
  e->e_consistent = false;        // an arbitrary state: E == D/eps is not given (until the next full step)
  e->graph_epoch++;
  double2 *dst = e->field[slot] + (size_t)e->sel * e->plane, *temp = nullptr;
  if (e->fp32) {
    cudaError_t err = cudaMalloc(&temp, e->plane * sizeof(double2));
    if (err != cudaSuccess) return b200_fail(B200FDTD_ERR_NOMEM, "setter staging plane: %s", cudaGetErrorString(err));
    dst = temp;
  }
  cudaError_t err = cudaMemcpy2DAsync(dst + (size_t)e->pitch + B200_JOFF, sizeof(double2) * e->pitch,
                                      host + 2 * (size_t)g.j0, sizeof(double2) * g.n_py,
                                      sizeof(double2) * g.nj, g.n_px, cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess && e->fp32)
    rc = b200_narrow_region(e, temp, (float2 *)e->field[slot] + (size_t)e->sel * e->plane);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  cudaFree(temp);
  if (err != cudaSuccess) return b200_fail(B200FDTD_ERR_CUDA, "field upload: %s", cudaGetErrorString(err));
  return rc;
}

int b200fdtd_field_digest(b200fdtd_engine *e, int32_t slot, uint64_t *digest)
{
  if (!e || !digest || slot < 0 || slot >= e->n_fields) return b200_fail(B200FDTD_ERR_ARG, "bad field slot %d", slot);
  int rc = select_device(e); if (rc) return rc;
  rc = b200_refresh_h(e); if (rc) return rc;
  rc = b200_refresh_e(e); if (rc) return rc;
  unsigned long long *acc = nullptr;
  cudaError_t err = cudaMalloc((void **)&acc, sizeof *acc);
  if (err != cudaSuccess) return b200_fail(B200FDTD_ERR_NOMEM, "digest accumulator: %s", cudaGetErrorString(err));
  cudaMemsetAsync(acc, 0, sizeof *acc, e->stream);
  const unsigned long long *plane =
      (const unsigned long long *)((const char *)e->field[slot] + (size_t)e->sel * e->plane * e->csize);
  if (e->fp32) digest_kernel<1><<<1184, 256, 0, e->stream>>>(plane, e->pitch, e->g.n_px, e->g.nj, e->g.n_py, e->g.j0, acc);
  else         digest_kernel<2><<<1184, 256, 0, e->stream>>>(plane, e->pitch, e->g.n_px, e->g.nj, e->g.n_py, e->g.j0, acc);
  e->launches++;
  unsigned long long host = 0;
  err = cudaMemcpyAsync(&host, acc, sizeof host, cudaMemcpyDeviceToHost, e->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  cudaFree(acc);
  if (err != cudaSuccess) return b200_fail(B200FDTD_ERR_CUDA, "field digest: %s", cudaGetErrorString(err));
  *digest = host;
  return B200FDTD_OK;
}

int b200fdtd_zero_state(b200fdtd_engine *e)
{
  if (!e) return b200_fail(B200FDTD_ERR_ARG, "NULL engine");
  int rc = select_device(e); if (rc) return rc;
  for (int s = 0; s < e->n_fields; s++)
    B200_CUDA(cudaMemsetAsync(e->field[s], 0, e->plane * e->csize * (size_t)e->n_batch, e->stream));
  e->h_stale = false;
  e->e_stale = false;
  e->e_consistent = true;
  // peer-halo flags restart with the step counter; a multi-rank reset must be bracketed by
  // the driver's own barrier (no rank may be mid-step while another zeroes)
  if (e->peer.flags) B200_CUDA(cudaMemsetAsync(e->peer.flags, 0, 2 * sizeof(unsigned long long), e->stream));
  e->peer.pending_down = 0;
  NtffState &n = e->ntff;
  if (n.ready) {
    const size_t nb = (size_t)e->n_batch;
    B200_CUDA(cudaMemsetAsync(n.hist_e, 0, sizeof(double2) * (size_t)n.n_local * n.max_time * nb, e->stream));
    B200_CUDA(cudaMemsetAsync(n.hist_h, 0, sizeof(double2) * (size_t)n.n_local * n.max_time * nb, e->stream));
    B200_CUDA(cudaMemsetAsync(n.uw, 0, sizeof(double2) * 3 * (size_t)n.n_angles * n.n_bins * nb, e->stream));
    n.steps_recorded = 0;
  }
  B200_CUDA(cudaStreamSynchronize(e->stream));
  return B200FDTD_OK;
}

int b200fdtd_ntff_project(b200fdtd_engine *e)
{
  if (!e) return b200_fail(B200FDTD_ERR_ARG, "NULL engine");
  int rc = select_device(e); if (rc) return rc;
  return b200_launch_ntff_project(e);
}

int b200fdtd_ntff_get_uw(b200fdtd_engine *e, int32_t slot, double *host)
{
  if (!e || !host || slot < 0 || slot > 2 || !e->ntff.ready) return b200_fail(B200FDTD_ERR_ARG, "bad U/W request");
  int rc = select_device(e); if (rc) return rc;
  const NtffState &n = e->ntff;
  const size_t count = (size_t)n.n_angles * n.n_bins;
  B200_CUDA(cudaMemcpyAsync(host, n.uw + ((size_t)e->sel * 3 + (size_t)slot) * count, sizeof(double2) * count,
                            cudaMemcpyDeviceToHost, e->stream));
  B200_CUDA(cudaStreamSynchronize(e->stream));
  return B200FDTD_OK;
}

// dst.U/W += src.U/W (both projected): the end-of-run sum over the slabs of one process
int b200fdtd_ntff_add_uw(b200fdtd_engine *dst, b200fdtd_engine *src)
{
  if (!dst || !src || !dst->ntff.ready || !src->ntff.ready || dst->ntff.n_angles != src->ntff.n_angles ||
      dst->ntff.n_bins != src->ntff.n_bins || dst->n_batch != src->n_batch)
    return b200_fail(B200FDTD_ERR_ARG, "U/W blocks of different shape");
  const size_t n = 2ull * 3ull * (size_t)dst->ntff.n_angles * (size_t)dst->ntff.n_bins * (size_t)dst->n_batch;
  int rc = select_device(src); if (rc) return rc;
  B200_CUDA(cudaStreamSynchronize(src->stream));
  rc = select_device(dst); if (rc) return rc;
  double *tmp = nullptr;
  cudaError_t err = cudaMalloc((void **)&tmp, n * sizeof(double));
  if (err != cudaSuccess) return b200_fail(B200FDTD_ERR_NOMEM, "U/W staging: %s", cudaGetErrorString(err));
  err = cudaMemcpyPeerAsync(tmp, dst->device, src->ntff.uw, src->device, n * sizeof(double), dst->stream);
  if (err == cudaSuccess) {
    axpy_double_kernel<<<592, 256, 0, dst->stream>>>((double *)dst->ntff.uw, tmp, n);
    dst->launches++;
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) err = cudaStreamSynchronize(dst->stream);
  cudaFree(tmp);
  if (err != cudaSuccess) return b200_fail(B200FDTD_ERR_CUDA, "U/W sum: %s", cudaGetErrorString(err));
  return B200FDTD_OK;
}

int b200fdtd_ntff_push_samples(b200fdtd_engine *e, int32_t t, const double *e_complex, const double *h_complex)
{
  if (!e || !e_complex || !h_complex || !e->ntff.ready) return b200_fail(B200FDTD_ERR_ARG, "bad sample push");
  NtffState &n = e->ntff;
  if (t < 0 || t >= n.max_time || e->n_batch > 1)
    return b200_fail(B200FDTD_ERR_ARG, "step %d outside the NTFF history [0, %d)", t, n.max_time);
  int rc = select_device(e); if (rc) return rc;
  if (n.n_local > 0) {            // hist[p][t]: one element per point, max_time elements apart
    B200_CUDA(cudaMemcpy2DAsync(n.hist_e + t, sizeof(double2) * (size_t)n.max_time, e_complex, sizeof(double2),
                                sizeof(double2), (size_t)n.n_local, cudaMemcpyHostToDevice, e->stream));
    B200_CUDA(cudaMemcpy2DAsync(n.hist_h + t, sizeof(double2) * (size_t)n.max_time, h_complex, sizeof(double2),
                                sizeof(double2), (size_t)n.n_local, cudaMemcpyHostToDevice, e->stream));
    B200_CUDA(cudaStreamSynchronize(e->stream));      // the caller's buffers are free again
  }
  if (t + 1 > n.steps_recorded) n.steps_recorded = t + 1;
  return B200FDTD_OK;
}

int b200fdtd_ntff_set_uw(b200fdtd_engine *e, int32_t slot, const double *host)
{
  if (!e || !host || slot < 0 || slot > 2 || !e->ntff.ready) return b200_fail(B200FDTD_ERR_ARG, "bad U/W request");
  int rc = select_device(e); if (rc) return rc;
  const NtffState &n = e->ntff;
  const size_t count = (size_t)n.n_angles * n.n_bins;
  B200_CUDA(cudaMemcpyAsync(n.uw + ((size_t)e->sel * 3 + (size_t)slot) * count, host, sizeof(double2) * count,
                            cudaMemcpyHostToDevice, e->stream));
  B200_CUDA(cudaStreamSynchronize(e->stream));
  return B200FDTD_OK;
}

int b200fdtd_ntff_uw_device(b200fdtd_engine *e, void **dev_ptr, uint64_t *n_doubles)
{
  if (!e || !dev_ptr || !n_doubles || !e->ntff.ready) return b200_fail(B200FDTD_ERR_ARG, "bad U/W request");
  *dev_ptr = e->ntff.uw;
  *n_doubles = 2ull * 3ull * (uint64_t)e->ntff.n_angles * (uint64_t)e->ntff.n_bins * (uint64_t)e->n_batch;
  return B200FDTD_OK;
}

int b200fdtd_ntff_spectrum(b200fdtd_engine *e, const b200fdtd_spectrum_args *args, double *out)
{
  if (!e || !args || !out || !args->cos_phi || !args->sin_phi || !args->twiddle)
    return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  int rc = select_device(e); if (rc) return rc;
  return b200_run_ntff_spectrum(e, args, out);
}

int b200fdtd_selftest_division(double divisor, uint64_t samples, uint64_t *mismatches)
{
  if (!mismatches || !(divisor >= 0.0)) return b200_fail(B200FDTD_ERR_ARG, "bad argument");
  unsigned long long bad = 0;
  int rc = b200_selftest_division(divisor, samples, &bad);
  *mismatches = bad;
  return rc;
}

int b200fdtd_ntff_frequency(b200fdtd_engine *e, const b200fdtd_freq_args *args, double *out)
{
  if (!e || !args || !out || !args->cos_a || !args->sin_a) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  int rc = select_device(e); if (rc) return rc;
  return b200_run_ntff_frequency(e, args, out);
}

int b200fdtd_launch_count(b200fdtd_engine *e, uint64_t *count)
{
  if (!e || !count) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  *count = e->launches;
  return B200FDTD_OK;
}

int b200fdtd_device_bytes(b200fdtd_engine *e, uint64_t *bytes)
{
  if (!e || !bytes) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  *bytes = e->dev_bytes;
  return B200FDTD_OK;
}

int b200fdtd_mem_info(int32_t device, uint64_t *free_bytes, uint64_t *total_bytes)
{
  if (!free_bytes || !total_bytes) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return b200_fail(B200FDTD_ERR_NODEVICE, "no CUDA device: the FDTD engine has no CPU path");
  if (device >= 0) B200_CUDA(cudaSetDevice(device));
  size_t f = 0, t = 0;
  B200_CUDA(cudaMemGetInfo(&f, &t));
  *free_bytes = f; *total_bytes = t;
  return B200FDTD_OK;
}

int b200fdtd_timer_start(b200fdtd_engine *e)
{
  if (!e) return b200_fail(B200FDTD_ERR_ARG, "NULL engine");
  int rc = select_device(e); if (rc) return rc;
  B200_CUDA(cudaEventRecord(e->ev0, e->stream));
  return B200FDTD_OK;
}

int b200fdtd_timer_stop(b200fdtd_engine *e, float *ms)
{
  if (!e || !ms) return b200_fail(B200FDTD_ERR_ARG, "NULL argument");
  int rc = select_device(e); if (rc) return rc;
  B200_CUDA(cudaEventRecord(e->ev1, e->stream));
  B200_CUDA(cudaEventSynchronize(e->ev1));
  B200_CUDA(cudaEventElapsedTime(ms, e->ev0, e->ev1));
  return B200FDTD_OK;
}

}  // extern "C"
