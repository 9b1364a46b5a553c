// Host-side TMA descriptor (CUtensorMap) construction with a small cache.
//
// cuTensorMapEncodeTiled is a driver-API entry point; it is resolved at run time through
// cudaGetDriverEntryPoint so the library links against nothing but the CUDA runtime.
// Descriptors are passed to kernels by value (__grid_constant__), so a captured CUDA graph
// bakes them into its kernel nodes exactly like it bakes raw pointers.
#pragma once
#include <cuda.h>
#include "common.cuh"

enum TmaElem { TMA_U8 = 0, TMA_F16 = 1, TMA_BF16 = 2, TMA_I32 = 3 };

// 2-D row-major tensor [rows, cols] with `row_stride_bytes` between rows; box = [box_rows, box_cols].
// swizzle128: CU_TENSOR_MAP_SWIZZLE_128B (box_cols * elem_size must be <= 128 bytes), else no swizzle.
int b2l_make_tmap_2d(CUtensorMap* out, const void* base, TmaElem elem, uint64_t rows, uint64_t cols,
                     uint64_t row_stride_bytes, uint32_t box_rows, uint32_t box_cols, bool swizzle128);
