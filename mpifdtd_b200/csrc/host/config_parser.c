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
