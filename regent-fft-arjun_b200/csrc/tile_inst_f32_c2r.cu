// generated list of instantiations: precision float, variant V_RR_C2R (see tile_inst.inc)
#define TT float
#define VAR V_RR_C2R
#define ROWONLY_VARIANT 1
#define TABLE_NAME tile_table_f32_c2r
#include "tile_inst.inc"
