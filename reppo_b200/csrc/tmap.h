// Host-side TMA tensor-map encoding shared by every tcgen05 kernel file.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace reppo {

// Row-major matrix [rows, cols] of `elem_bytes`-byte elements (2 = bf16, 4 = fp32 read as tf32) with leading dimension
// `ld` elements; box = {box_cols, box_rows}; 128-byte swizzle (box_cols * elem_bytes must be 128).  The base must be
// 16-byte aligned and ld * elem_bytes a multiple of 16.  Out-of-bounds elements read as zero.
// atom_32b: the 128-byte swizzle with 32-byte atoms (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), the only shared-memory layout
// tcgen05 accepts for MN-major 32-bit (tf32) operands (UMMA layout type SWIZZLE_128B_BASE32B).
bool tmap_encode_2d(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_cols, int box_rows,
                    int elem_bytes = 2, bool atom_32b = false);

}  // namespace reppo
