// tile_registry.h — lookup of the compiled shared-memory Stockham kernels (tile_kernel.cuh).
#pragma once
#include "tile_kernel.cuh"

namespace fftb200 {

struct TileKernelInfo {
    void (*fn)(const TileParams);
    int L, R, W, threads, smem_bytes;
    int cluster;  // CTAs per thread-block cluster (1 = ordinary launch); L = cluster * (length one CTA holds)
};

// prec: 0 = fp32 (complex32), 1 = fp64 (complex64).  Returns nullptr when L is not compiled.
const TileKernelInfo *find_tile_kernel(int prec, int variant, int L);
// largest single-tile length compiled for this precision
int max_tile_length(int prec);

// per-(precision, variant) tables, one translation unit each (tile_inst_*.cu)
#define FFTB200_DECL_TABLE(name) const TileKernelInfo *name(int *count)
FFTB200_DECL_TABLE(tile_table_f64_rr);
FFTB200_DECL_TABLE(tile_table_f64_cc);
FFTB200_DECL_TABLE(tile_table_f64_cctw);
FFTB200_DECL_TABLE(tile_table_f64_rc);
FFTB200_DECL_TABLE(tile_table_f64_r2c);
FFTB200_DECL_TABLE(tile_table_f64_ccp);
FFTB200_DECL_TABLE(tile_table_f64_c2r);
FFTB200_DECL_TABLE(tile_table_f64_rcp);
FFTB200_DECL_TABLE(tile_table_f32_rr);
FFTB200_DECL_TABLE(tile_table_f32_cc);
FFTB200_DECL_TABLE(tile_table_f32_cctw);
FFTB200_DECL_TABLE(tile_table_f32_rc);
FFTB200_DECL_TABLE(tile_table_f32_r2c);
FFTB200_DECL_TABLE(tile_table_f32_ccp);
FFTB200_DECL_TABLE(tile_table_f32_c2r);
FFTB200_DECL_TABLE(tile_table_f32_rcp);

}  // namespace fftb200
