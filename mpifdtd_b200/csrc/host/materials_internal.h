/* materials_internal.h -- shared by the material-model translation units. */
#ifndef MPIFDTD_MATERIALS_INTERNAL_H
#define MPIFDTD_MATERIALS_INTERNAL_H
#include "mpifdtd_plugin.h"

/* eps callback: permittivity of the cell-sized neighbourhood centred on (x, y)
 * in cell units.  (col,row) = (1,0) averages along x only, (0,1) along y only,
 * (1,1) over the cell area (models.c:136-144). */
typedef double (*material_eps_fn)(double x, double y, int col, int row);

typedef struct MaterialModel {
  const char *dir;                      /* top-level output directory name   */
  material_eps_fn (*select)(void);      /* run at models_setModel time       */
  void (*prepare)(void);                /* run after field_init              */
  void (*need_size)(int *x_nm, int *y_nm);
  bool (*advance)(void);                /* step the parameter sweep; true = exhausted */
  void (*enter_dirs)(void);             /* chdir chain that encodes the parameters    */
} MaterialModel;

extern const MaterialModel material_vacuum, material_mie_cylinder, material_concentric,
    material_multilayer, material_morpho, material_zigzag, material_trace_image;

/* the reference's function.h:7-8 macros, kept as macros because their
 * usual-arithmetic-conversion result type (int vs double) is part of the
 * arithmetic being reproduced */
#define MPIFDTD_MAX(a, b) ((a) > (b) ? (a) : (b))
#define MPIFDTD_MIN(a, b) ((a) < (b) ? (a) : (b))

/* Sub-cell sampling lattice shared by most models: offsets -4.5 ... 4.5 in
 * tenths of a cell (e.g. circleModel.c:46-57). */
#define SUBCELL_SPLIT 10.0
#define FOR_SUBCELL(u) for (double u = -SUBCELL_SPLIT / 2 + 0.5; u < SUBCELL_SPLIT / 2; u += 1)

#endif
