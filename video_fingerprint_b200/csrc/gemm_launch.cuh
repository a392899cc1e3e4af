// Host launch helpers for gemm_tcgen05_kernel.
#pragma once
#include "epilogues.cuh"
#include "gemm_sm100.cuh"
#include "tmap.cuh"

namespace vfp {

inline int device_sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

template <int BLOCK_N, int BLOCK_K, int STAGES, class Epi>
inline cudaError_t launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmShape& shape,
                               const typename Epi::Params& ep, cudaStream_t stream) {
  using L = GemmSmemLayout<BLOCK_N, BLOCK_K, STAGES>;
  static_assert(L::kTotal <= 232448, "shared memory budget");
  auto kernel = gemm_tcgen05_kernel<BLOCK_N, BLOCK_K, STAGES, Epi>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int total = shape.m_tiles * shape.n_tiles;
  if (total <= 0) return cudaSuccess;
  const int grid = total < device_sm_count() ? total : device_sm_count();
  kernel<<<grid, kGemmThreads, L::kTotal, stream>>>(ta, tb, shape, ep);
  return cudaGetLastError();
}

inline GemmShape plain_shape(long long M, int N, int K, int block_n, int block_k, int group_m = 16) {
  GemmShape s{};
  s.m_tiles = (int)((M + kBlockM - 1) / kBlockM);
  s.n_tiles = (N + block_n - 1) / block_n;
  s.k_blocks = K / block_k;
  s.group_m = group_m;
  s.a_conv = 0;
  s.tiles_per_frame = 1;
  s.frames_per_tile = 1;
  s.tile_out_rows = 0;
  s.cblocks_per_tap = 1;
  return s;
}

}  // namespace vfp
