// generated list of instantiations: precision float, variant V_RC (see tile_inst.inc)
#define TT float
#define TT_IS_FLOAT 1
#define VAR V_RC
#define TABLE_NAME tile_table_f32_rc
#include "tile_inst.inc"
