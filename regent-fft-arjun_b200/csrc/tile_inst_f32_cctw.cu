// generated list of instantiations: precision float, variant V_CC_TW (see tile_inst.inc)
#define TT float
#define TT_IS_FLOAT 1
#define VAR V_CC_TW
#define COL_VARIANT 1
#define TABLE_NAME tile_table_f32_cctw
#include "tile_inst.inc"
