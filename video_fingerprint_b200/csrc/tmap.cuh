// Host-side CUtensorMap construction. The driver entry point is resolved through the runtime
// (cudaGetDriverEntryPoint), so the library carries no link-time dependency on libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace vfp {

typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                             CUtensorMapFloatOOBfill);

inline PFN_tensorMapEncodeTiled tensor_map_encoder() {
  static PFN_tensorMapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || !p) {
      return nullptr;
    }
    fn = reinterpret_cast<PFN_tensorMapEncodeTiled>(p);
  }
  return fn;
}

inline CUtensorMapSwizzle swizzle_for_row_bytes(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
         : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
         : row_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                           : CU_TENSOR_MAP_SWIZZLE_NONE;
}

// bf16 matrix [rows][cols] with row pitch `ld` elements; box = box_rows x box_cols (cols innermost).
inline int make_tmap_rows_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                               uint32_t box_rows, uint32_t box_cols) {
  PFN_tensorMapEncodeTiled enc = tensor_map_encoder();
  if (!enc) return 1;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(box_cols * 2),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

// fp32 matrix [rows][cols] (dense), box = box_rows x all `cols` columns (cols <= 256), no swizzle.
inline int make_tmap_rows_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  PFN_tensorMapEncodeTiled enc = tensor_map_encoder();
  if (!enc) return 1;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 4};
  cuuint32_t box[2] = {(cuuint32_t)cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

// Row-major output matrix [rows][cols] written by the epilogue in 32-column x 32-row boxes (see EpiBiasActTma).
inline int make_tmap_out(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, bool bf16) {
  PFN_tensorMapEncodeTiled enc = tensor_map_encoder();
  if (!enc) return 1;
  const uint64_t esz = bf16 ? 2 : 4;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * esz};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base),
                   gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   bf16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

// bf16 NHWC activation tensor [frames][H][W][C] read as the A operand of an implicit-GEMM convolution:
// one box = (c_box channels) x (out_w positions along W) x (out_h positions along H) x (n_box frames), visiting
// every `stride`-th pixel along W and H (stride 2 = the 3x3/stride-2 convolutions; stride 1 = a space-to-depth
// input). Out-of-range coordinates (the -1 halo, frames past the end) are zero-filled by the TMA unit.
inline int make_tmap_nhwc_bf16(CUtensorMap* out, const void* base, uint64_t frames, uint32_t H, uint32_t W, uint32_t C,
                               uint32_t c_box, uint32_t out_w, uint32_t out_h, uint32_t n_box, uint32_t stride) {
  PFN_tensorMapEncodeTiled enc = tensor_map_encoder();
  if (!enc) return 1;
  cuuint64_t gdim[4] = {C, W, H, frames};
  cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {c_box, out_w * stride, out_h * stride, n_box};
  cuuint32_t estr[4] = {1, stride, stride, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(c_box * 2),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}
inline int make_tmap_conv_s2_bf16(CUtensorMap* out, const void* base, uint64_t frames, uint32_t H, uint32_t W,
                                  uint32_t C, uint32_t c_box, uint32_t out_w, uint32_t out_h, uint32_t n_box) {
  return make_tmap_nhwc_bf16(out, base, frames, H, W, C, c_box, out_w, out_h, n_box, 2);
}

}  // namespace vfp
