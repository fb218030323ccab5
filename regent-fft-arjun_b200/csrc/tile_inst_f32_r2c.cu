// generated list of instantiations: precision float, variant V_RR_R2C (see tile_inst.inc)
#define TT float
#define TT_IS_FLOAT 1
#define VAR V_RR_R2C
#define ROWONLY_VARIANT 1
#define TABLE_NAME tile_table_f32_r2c
#include "tile_inst.inc"
