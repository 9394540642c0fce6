// Host side of the TMA tensor copies: CUtensorMap encoding through the driver entry point (no link-time dependency on
// libcuda; the runtime hands out cuTensorMapEncodeTiled).  Maps are encoded per launch (the tensors live in the caller's
// workspace) and passed to the kernels as __grid_constant__ parameters.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace pca {

// 2-D row-major bf16 tensor (rows, inner) with `row_stride_bytes` between rows; box = (box_inner, box_rows); no swizzle
// (the kernels place 16-byte-wide boxes so that the tile lands in the canonical no-swizzle UMMA operand layout).
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, unsigned long long inner, unsigned long long rows,
                      unsigned long long row_stride_bytes, unsigned box_inner, unsigned box_rows);

// (rows, 64) bf16 tensor viewed as (8 elements, rows, 8 chunks): one box of (8, box_rows, 8) lands as [chunk][row][16 B]
int make_tmap_chunked_bf16(CUtensorMap* out, const void* base, unsigned long long rows, unsigned long long row_stride_bytes,
                           unsigned box_rows);

// 3-D fp32 tensor (batch, rows, inner) with byte strides (batch_stride, row_stride, 4); box = (box_inner, box_rows, 1); no
// swizzle: a box lands in shared memory as box_rows rows of box_inner floats.  Rows / columns outside the tensor are zero-filled.
int make_tmap_3d_f32(CUtensorMap* out, const void* base, unsigned long long inner, unsigned long long rows, unsigned long long batch,
                     unsigned long long row_stride_bytes, unsigned long long batch_stride_bytes, unsigned box_inner, unsigned box_rows);

}  // namespace pca
