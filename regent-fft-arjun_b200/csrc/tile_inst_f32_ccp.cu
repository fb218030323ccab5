// generated list of instantiations: precision float, variant V_CC_PEER (see tile_inst.inc)
#define TT float
#define VAR V_CC_PEER
#define COL_VARIANT 1
#define TABLE_NAME tile_table_f32_ccp
#include "tile_inst.inc"
