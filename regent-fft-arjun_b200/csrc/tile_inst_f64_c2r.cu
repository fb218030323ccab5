// generated list of instantiations: precision double, variant V_RR_C2R (see tile_inst.inc)
#define TT double
#define TT_IS_DOUBLE 1
#define VAR V_RR_C2R
#define ROWONLY_VARIANT 1
#define TABLE_NAME tile_table_f64_c2r
#include "tile_inst.inc"
