// Host-side TMA tensor-map encoding shared by every tcgen05 kernel file.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace reppo {

// Row-major matrix [rows, cols] of `elem_bytes`-byte elements (2 = bf16, 4 = fp32 read as tf32) with leading dimension
// `ld` elements; box = {box_cols, box_rows}; 128-byte swizzle (box_cols * elem_bytes must be 128).  The base must be
// 16-byte aligned and ld * elem_bytes a multiple of 16.  Out-of-bounds elements read as zero.
bool tmap_encode_2d(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_cols, int box_rows,
                    int elem_bytes = 2);

}  // namespace reppo
