// generated list of instantiations: precision float, variant V_CC (see tile_inst.inc)
#define TT float
#define TT_IS_FLOAT 1
#define VAR V_CC
#define COL_VARIANT 1
#define TABLE_NAME tile_table_f32_cc
#include "tile_inst.inc"
