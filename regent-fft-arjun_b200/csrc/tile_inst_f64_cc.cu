// generated list of instantiations: precision double, variant V_CC (see tile_inst.inc)
#define TT double
#define TT_IS_DOUBLE 1
#define VAR V_CC
#define COL_VARIANT 1
#define TABLE_NAME tile_table_f64_cc
#include "tile_inst.inc"
