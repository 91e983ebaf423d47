/* mpifdtd_sweep -- batch driver over libmpifdtd_b200.so.
 *
 * Same job as the reference's main.c batch mode (main.c:150-213): for every structure of the
 * selected model (models_isFinish iterates the model's parameter), size the region
 * (calcFieldSize, main.c:71-88, unless config.txt gives it), enter the reference's directory
 * chain <model dir>/.../hu_<h>nm/<solver dir> (moveDir, main.c:57-69) and run the incidence
 * angles start..end step delta -- but each structure's angles go to the GPU as ONE batched
 * engine (mpifdtd_runAngleSweep) instead of one angle per MPI rank.
 *
 *   mpifdtd_sweep [config.txt] [--max-batch N]
 *
 * config.txt: the 11 values of configSample.txt (width, height, h_u, pml, lambda, steps,
 * start/end/delta angle, model id, solver id); without it the defaults of main.c's
 * initParameter() with MIE_CYLINDER / TM_UPML_2D.  Several GPUs: start one process per GPU
 * with RANK / WORLD_SIZE (or OMPI_COMM_WORLD_RANK / _SIZE) set and CUDA_VISIBLE_DEVICES
 * selecting the device; rank r takes angles start + r*delta, step delta*world -- the
 * reference's rank striding (main.c:126-138) -- and only rank 0 opens config.txt: the others
 * receive the struct from it (initConfigFromText, main.c:368-394).  Errors: message + exit(2), as everywhere.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "mpifdtd_plugin.h"

static char root[512];

static int env_int(const char *a, const char *b, int fallback)
{
  const char *v = getenv(a);
  if (v == NULL && b != NULL) v = getenv(b);
  return v != NULL ? atoi(v) : fallback;
}

static void move_dir(const FieldInfo *info)            /* main.c:57-69 */
{
  char buf[128];
  moveDirectory(root);
  models_moveDirectory();
  sprintf(buf, "hu_%dnm", info->h_u_nm);
  makeDirectory(buf);
  moveDirectory(buf);
  simulator_moveDirectory();
}

static void calc_field_size(FieldInfo *info)           /* main.c:71-88 */
{
  int x_nm, y_nm;
  models_needSize(&x_nm, &y_nm);
  info->width_nm  = x_nm + info->h_u_nm * (info->pml + 5) * 2 + 200;
  info->height_nm = y_nm + info->h_u_nm * (info->pml + 5) * 2 + 200;
}

int main(int argc, char **argv)
{
  MpifdtdConfig cfg;
  const char *config_path = NULL;
  int max_batch = 0, have_config = 0;
  for (int a = 1; a < argc; a++) {
    if (strcmp(argv[a], "--max-batch") == 0 && a + 1 < argc) max_batch = atoi(argv[++a]);
    else config_path = argv[a];
  }
  const int rank = env_int("RANK", "OMPI_COMM_WORLD_RANK", 0);
  const int world = env_int("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", 1);
  if (config_path != NULL) {
    /* main.c:368-394: only rank 0 reads the file, the others get the struct from it */
    mpifdtd_initConfigFromText(config_path, rank, world, &cfg, NULL, NULL, NULL);
    have_config = 1;
  } else {                                             /* initParameter(), main.c:90-108 */
    memset(&cfg, 0, sizeof cfg);
    cfg.field_info.h_u_nm = 50;
    cfg.field_info.pml = 15;
    cfg.field_info.lambda_nm = 500;
    cfg.field_info.stepNum = 20000 / cfg.field_info.h_u_nm;
    cfg.startAngle = 0;  cfg.endAngle = 0;  cfg.deltaAngle = 5;
    cfg.ModelType = MIE_CYLINDER;
    cfg.SolverType = TM_UPML_2D;
  }
  if (world < 1 || rank < 0 || rank >= world || cfg.deltaAngle <= 0) {
    printf("mpifdtd_sweep: bad rank %d of %d or angle step %d\n", rank, world, cfg.deltaAngle);
    exit(2);
  }
  if (getcwd(root, sizeof root) == NULL) { printf("cannot read the working directory\n"); exit(2); }

  models_setModel((enum MODEL)cfg.ModelType);
  simulator_setSolver((enum SOLVER)cfg.SolverType);
  const int start = cfg.startAngle + rank * cfg.deltaAngle, delta = cfg.deltaAngle * world;
  int structures = 0, simulations = 0;
  do {
    FieldInfo info = cfg.field_info;
    if (!have_config) calc_field_size(&info);
    printf("structure %d: field size (%d nm, %d nm)\n", structures, info.width_nm, info.height_nm);
    move_dir(&info);
    if (start <= cfg.endAngle) {
      if (cfg.SolverType == TM_UPML_2D || cfg.SolverType == TE_UPML_2D) {
        simulations += mpifdtd_runAngleSweep(info, start, cfg.endAngle, delta, max_batch);
      } else {                                         /* the other solvers: one angle at a time */
        info.angle_deg = start;
        simulator_init(info);
        for (int ang = start; ang <= cfg.endAngle; ang += delta) {
          while (!simulator_isFinish()) simulator_calc();
          simulations++;
          if (ang + delta > cfg.endAngle) break;
          simulator_reset();
          field_setWaveAngle(ang + delta);
        }
        simulator_finish();
      }
    }
    structures++;
  } while (!models_isFinish());                        /* next structure parameter, main.c:126-132 */
  moveDirectory(root);
  printf("mpifdtd_sweep: rank %d of %d ran %d simulation(s) over %d structure(s)\n", rank, world, simulations,
         structures);
  return 0;
}
