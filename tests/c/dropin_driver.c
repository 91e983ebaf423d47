/* A minimal C driver that uses ONLY the reference's plugin surface, in the order the
 * reference's own main.c does (main.c:150-213): select model and solver, init, step
 * until finished, read a field through the borrowed-pointer getter, finish (which
 * writes "<angle>[deg]_380nm_700nm_b.dat" into cwd).  It is compiled against
 * libmpifdtd_b200.so by tests/test_dropin_cpu.py to prove C linkage; with a GPU it
 * runs end to end, without one it must exit(2) like every reference error path. */
#include <stdio.h>
#include <stdlib.h>
#include "mpifdtd_plugin.h"

int main(int argc, char **argv)
{
  int n = argc > 1 ? atoi(argv[1]) : 96;
  int steps = argc > 2 ? atoi(argv[2]) : 120;
  int solver = argc > 3 ? atoi(argv[3]) : TM_UPML_2D;
  models_setModel(MIE_CYLINDER);
  simulator_setSolver((enum SOLVER)solver);
  FieldInfo info = { n * 20, n * 20, 20, 10, 500, 0, steps };
  simulator_init(info);
  while (!simulator_isFinish())
    simulator_calc();
  FieldInfo_S g = field_getFieldInfo_S();
  double complex *draw = simulator_getDrawingData();
  double *eps = simulator_getEps();
  double peak = 0, solid = 0;
  for (int k = 0; k < g.N_CELL; k++) {
    double v = cnorm(draw[k]);
    if (v > peak) peak = v;
    if (eps[k] > 1.0) solid += 1;
  }
  printf("DRIVER cells=%d steps=%d peak=%.17g solid=%.0f\n", g.N_CELL, (int)field_getTime(), peak, solid);
  simulator_finish();
  return 0;
}
