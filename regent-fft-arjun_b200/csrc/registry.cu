// registry.cu — lookup of the compiled tile kernels (tile_inst_*.cu tables) by precision, variant and length
#include <cstdlib>

#include "plan_internal.h"

namespace fftb200 {

typedef const TileKernelInfo *(*table_fn)(int *);
static table_fn k_tables[2][V_COUNT] = {
    {tile_table_f32_rr, tile_table_f32_cc, tile_table_f32_cctw, tile_table_f32_rc, tile_table_f32_r2c, tile_table_f32_ccp,
     tile_table_f32_c2r, tile_table_f32_rcp},
    {tile_table_f64_rr, tile_table_f64_cc, tile_table_f64_cctw, tile_table_f64_rc, tile_table_f64_r2c, tile_table_f64_ccp,
     tile_table_f64_c2r, tile_table_f64_rcp}};

const TileKernelInfo *find_tile_kernel(int prec, int variant, int L) {
    int n = 0;
    const TileKernelInfo *t = k_tables[prec][variant](&n);
    // FFTB200_TILE_ALT=k (tuning experiments only): take the k-th alternative row compiled for this length;
    // FFTB200_TILE_ALT_ROW / FFTB200_TILE_ALT_COL do the same for contiguous-axis / strided-axis passes only
    const bool row_only = (variant == V_RR || variant == V_RR_R2C || variant == V_RR_C2R);
    const char *alt = getenv(row_only ? "FFTB200_TILE_ALT_ROW" : "FFTB200_TILE_ALT_COL");
    if (!(alt && *alt)) alt = getenv("FFTB200_TILE_ALT");
    int skip = (alt && *alt) ? atoi(alt) : 0;
    const TileKernelInfo *first = nullptr;
    for (int i = 0; i < n; ++i)
        if (t[i].L == L) {
            if (!first) first = &t[i];
            if (skip-- == 0) return &t[i];
        }
    return first;
}

int max_tile_length(int prec) {
    int n = 0, m = 0;
    const TileKernelInfo *t = k_tables[prec][V_RR](&n);
    for (int i = 0; i < n; ++i) m = t[i].L > m ? t[i].L : m;
    return m;
}

}  // namespace fftb200
