// generated list of instantiations: precision double, variant V_RR (see tile_inst.inc)
#define TT double
#define TT_IS_DOUBLE 1
#define VAR V_RR
#define ROWONLY_VARIANT 1
#define TABLE_NAME tile_table_f64_rr
#include "tile_inst.inc"
