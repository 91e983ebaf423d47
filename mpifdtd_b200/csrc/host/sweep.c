/* sweep.c -- the incidence-angle sweep of one structure as batched GPU work.
 *
 * What it replaces (rennone/mpiFDTD main.c:114-138, 183-211): every MPI rank runs one angle
 * at a time -- simulator_init, stepNum x simulator_calc, then simulator_reset() +
 * field_setWaveAngle(next) -- and the ranks stride through the angle list.  All simulations
 * of one structure share grid, permittivity and UPML coefficients; only the source differs.
 * Here a chunk of angles becomes ONE batched engine (mpifdtd_setAngleBatch): each
 * simulator_calc() advances every angle of the chunk in the same kernel launches, and
 * simulator_finish() writes every angle's far-field files, named like the reference's.
 * Several GPUs: give each process its own slice of the angle list (the reference's
 * rank/numProc striding), no communication needed.
 */
#include <stdio.h>
#include <stdlib.h>
#include "b200fdtd.h"
#include "host_internal.h"

/* bytes one simulation of the batch keeps on the device (fields + NTFF history + U/W) */
static double bytes_per_simulation(FieldInfo info, int complex_bytes)
{
  const double n_px = info.width_nm / info.h_u_nm + 2.0 * info.pml;
  const double n_py = info.height_nm / info.h_u_nm + 2.0 * info.pml;
  const double plane = (n_px + 2) * (n_py + 16);
  const double perimeter = 2 * (n_px + n_py);
  return 9 * plane * complex_bytes + 2 * perimeter * info.stepNum * 16.0 + 3 * 360.0 * info.stepNum * 16.0;
}

int mpifdtd_runAngleSweep(FieldInfo field_info, int start_deg, int end_deg, int delta_deg, int max_batch)
{
  if (delta_deg <= 0 || end_deg < start_deg) {
    printf("mpifdtd_runAngleSweep: bad angle range %d..%d step %d\n", start_deg, end_deg, delta_deg);
    exit(2);
  }
  const int total = (end_deg - start_deg) / delta_deg + 1;
  int chunk = total;
  if (max_batch > 0 && chunk > max_batch) chunk = max_batch;

  /* keep the batch within 80 % of the free device memory (shared arrays are small next to it) */
  uint64_t free_b = 0, total_b = 0;
  int rc = b200fdtd_mem_info(-1, &free_b, &total_b);
  if (rc != B200FDTD_OK) { printf("b200fdtd: mem_info failed (%d): %s\n", rc, b200fdtd_last_error()); exit(2); }
  const double per_sim = bytes_per_simulation(field_info, 16);
  const double fit = 0.8 * (double)free_b / per_sim;
  if (fit < 1.0) { printf("mpifdtd_runAngleSweep: one simulation (%.1f GB) does not fit the device\n", per_sim / 1e9); exit(2); }
  if ((double)chunk > fit) chunk = (int)fit;
  if (chunk > 4096) chunk = 4096;

  if (chunk == 1) {
    /* one angle at a time, exactly the reference's loop (main.c:183-211): init once, then
     * reset() -- far-field files + zeroed state -- and field_setWaveAngle() between angles */
    mpifdtd_setAngleBatch(NULL, 0);
    field_info.angle_deg = start_deg;
    simulator_init(field_info);
    for (int k = 0; k < total; k++) {
      while (!simulator_isFinish()) simulator_calc();
      if (k + 1 == total) break;
      simulator_reset();
      field_setWaveAngle(start_deg + (k + 1) * delta_deg);
    }
    simulator_finish();
    return total;
  }

  int *angles = (int *)malloc(sizeof(int) * (size_t)chunk);
  int done = 0;
  while (done < total) {
    const int n = (total - done < chunk) ? total - done : chunk;
    for (int k = 0; k < n; k++) angles[k] = start_deg + (done + k) * delta_deg;
    printf("angle sweep: %d simulation(s) in one batch, %d..%d deg\n", n, angles[0], angles[n - 1]);
    /* a batch of one is the plain single-simulation path */
    mpifdtd_setAngleBatch(angles, n > 1 ? n : 0);
    field_info.angle_deg = angles[0];
    simulator_init(field_info);
    while (!simulator_isFinish()) simulator_calc();
    simulator_finish();                       /* every angle's "<ang>[deg]..." files, then free */
    done += n;
  }
  mpifdtd_setAngleBatch(NULL, 0);
  free(angles);
  return total;
}
