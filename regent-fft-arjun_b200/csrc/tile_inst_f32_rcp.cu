// generated list of instantiations: precision float, variant V_RC_PEER (see tile_inst.inc)
#define TT float
#define VAR V_RC_PEER
#define TABLE_NAME tile_table_f32_rcp
#include "tile_inst.inc"
