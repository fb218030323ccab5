// generated list of instantiations: precision double, variant V_RC_PEER (see tile_inst.inc)
#define TT double
#define TT_IS_DOUBLE 1
#define VAR V_RC_PEER
#define TABLE_NAME tile_table_f64_rcp
#include "tile_inst.inc"
