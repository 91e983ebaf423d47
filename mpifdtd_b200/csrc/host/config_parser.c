/* config_parser.c -- the config.txt reader (parser.c:3-20, configSample.txt:6-22,
 * and the readConfig the reference keeps commented out at main.c:319-366).
 *
 * File format: one value per line, '#' starts a comment, blank lines and lines
 * that begin with '#' are skipped.  The 11 values, in order: width, height,
 * h_u, pml, lambda, step, start angle, end angle, delta angle, model id
 * (enum MODEL), solver id (enum SOLVER).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mpifdtd_plugin.h"

/* Next payload line into buf (>= 1024 bytes).  Upstream's "cut after '#'" is a
 * no-op (strncpy onto itself, parser.c:14); callers rely on atoi() stopping at
 * the first non-digit, so the trailing comment is left in place here too. */
bool parser_nextLine(FILE *fp, char buf[])
{
  while (fgets(buf, 1024, fp) != NULL) {
    if (buf[0] == '#' || buf[0] == '\0' || buf[0] == '\n')
      continue;
    return true;
  }
  return false;
}

int mpifdtd_readConfig(const char *path, MpifdtdConfig *out)
{
  FILE *fp = fopen(path, "r");
  if (fp == NULL) { printf("cannot open file %s \n", path); exit(2); }
  int *slots[11] = {
    &out->field_info.width_nm, &out->field_info.height_nm, &out->field_info.h_u_nm,
    &out->field_info.pml, &out->field_info.lambda_nm, &out->field_info.stepNum,
    &out->startAngle, &out->endAngle, &out->deltaAngle, &out->ModelType, &out->SolverType };
  char line[1024];
  for (int n = 0; n < 11; n++) {
    if (!parser_nextLine(fp, line)) {
      printf("parse error, config.txt needs 11 values, found %d\n", n);
      exit(2);
    }
    /* lambda goes through strtod upstream (main.c:342), the rest through atoi */
    *slots[n] = (n == 4) ? (int)strtod(line, NULL) : atoi(line);
  }
  fclose(fp);
  out->field_info.angle_deg = out->startAngle;
  return 0;
}

/* ---- initConfigFromText (main.c:368-394, commented out upstream) -----------------------------
 * "It would be bad if everybody read the file at once, so only rank 0 reads and then the config
 * is synchronised": rank 0 parses config.txt, prints the FieldSetting banner and sends the struct
 * -- all ints -- to every other rank; the others receive it.  Upstream's transport is
 * MPI_Send / MPI_Recv of sizeof(Config)/sizeof(int) MPI_INTs with tag 1.  Here the transport is a
 * pair of callbacks (an MPI build passes thin wrappers of MPI_Send / MPI_Recv; the Python harness
 * passes torch.distributed send / recv), and with send == recv == NULL a built-in one for
 * processes on ONE node that were started without any message layer (one `mpifdtd_sweep` per GPU
 * with RANK / WORLD_SIZE in the environment): rank 0 publishes the ints in a file under /dev/shm
 * (written under a temporary name, then renamed: readers never see a partial file), the other
 * ranks poll for it.  The file is keyed by MPIFDTD_JOB_ID, or by the parent process id when that
 * is not set (ranks started by one launcher share it). */
#include <errno.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#define CONFIG_INTS ((int)(sizeof(MpifdtdConfig) / sizeof(int)))

static void bcast_path(char *dst, size_t n)
{
  const char *job = getenv("MPIFDTD_JOB_ID");
  if (job != NULL && job[0] != '\0') snprintf(dst, n, "/dev/shm/mpifdtd_config_%s", job);
  else snprintf(dst, n, "/dev/shm/mpifdtd_config_ppid%ld", (long)getppid());
}

static void print_banner(const MpifdtdConfig *c)                 /* main.c:373-381 */
{
  printf("===========FieldSetting=======\n");
  printf("fieldSize(nm) = (%d, %d) \nh_u = %d \npml = %d\n", c->field_info.width_nm, c->field_info.height_nm,
         c->field_info.h_u_nm, c->field_info.pml);
  printf("lambda(nm) = %d  \nstep = %d\n", c->field_info.lambda_nm, c->field_info.stepNum);
  printf("angle = %d .. %d (delta = %d)\n", c->startAngle, c->endAngle, c->deltaAngle);
  printf("==============================\n");
}

int mpifdtd_initConfigFromText(const char *path, int rank, int n_ranks, MpifdtdConfig *cfg,
                               mpifdtd_send_ints send, mpifdtd_recv_ints recv, void *ctx)
{
  if (cfg == NULL || rank < 0 || n_ranks < 1 || rank >= n_ranks || (send == NULL) != (recv == NULL)) {
    printf("mpifdtd_initConfigFromText: bad rank %d of %d or half a transport\n", rank, n_ranks);
    exit(2);
  }
  char shm[256];
  bcast_path(shm, sizeof shm);
  if (rank == 0) {
    mpifdtd_readConfig(path, cfg);                                /* exit(2) on a missing / short file */
    print_banner(cfg);
    if (send != NULL) {
      for (int i = 1; i < n_ranks; i++)
        if (send((const int *)cfg, CONFIG_INTS, i, ctx) != 0) { printf("config broadcast: send to rank %d failed\n", i); exit(2); }
    } else if (n_ranks > 1) {
      char tmp[300];
      snprintf(tmp, sizeof tmp, "%s.tmp%ld", shm, (long)getpid());
      FILE *fp = fopen(tmp, "wb");
      if (fp == NULL || fwrite(cfg, sizeof(int), CONFIG_INTS, fp) != (size_t)CONFIG_INTS || fclose(fp) != 0 ||
          rename(tmp, shm) != 0) {
        printf("config broadcast: cannot publish %s\n", shm);
        exit(2);
      }
    }
    return 0;
  }
  if (recv != NULL) {
    if (recv((int *)cfg, CONFIG_INTS, 0, ctx) != 0) { printf("config broadcast: receive from rank 0 failed\n"); exit(2); }
    return 0;
  }
  double waited = 0, limit = 120;
  if (getenv("MPIFDTD_BCAST_TIMEOUT_S") != NULL) limit = atof(getenv("MPIFDTD_BCAST_TIMEOUT_S"));
  const time_t not_before = time(NULL) - 120;           /* a leftover of an earlier job under the same key is not ours */
  for (;;) {
    struct stat st;
    FILE *fp = (stat(shm, &st) == 0 && st.st_mtime >= not_before) ? fopen(shm, "rb") : NULL;
    if (fp != NULL) {
      size_t got = fread(cfg, sizeof(int), CONFIG_INTS, fp);
      fclose(fp);
      if (got == (size_t)CONFIG_INTS) return 0;
    }
    if (waited >= limit) { printf("config broadcast: rank 0 never published %s\n", shm); exit(2); }
    struct timespec ts = { 0, 20 * 1000 * 1000 };
    nanosleep(&ts, NULL);
    waited += 0.02;
  }
}

/* rank 0 removes the published file once every rank has its copy (after the job's first barrier,
 * or at exit); harmless when nothing was published */
void mpifdtd_configBroadcastDone(void)
{
  char shm[256];
  bcast_path(shm, sizeof shm);
  unlink(shm);
}
