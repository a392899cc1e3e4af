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

template <int BLOCK_N, int BLOCK_K, int STAGES, class Epi, int MT = 1, bool SWAP = false>
inline cudaError_t launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmShape& shape,
                               const typename Epi::Params& ep, cudaStream_t stream) {
  using L = GemmSmemLayout<BLOCK_N, BLOCK_K, STAGES, MT>;
  constexpr int kSmem = L::kTotal + Epi::kExtraSmemBytes;
  static_assert(kSmem <= 232448, "shared memory budget");
  auto kernel = gemm_tcgen05_kernel<BLOCK_N, BLOCK_K, STAGES, Epi, MT, SWAP>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int m_super = (shape.m_tiles + MT - 1) / MT;
  const int total = shape.row_resident ? m_super * shape.n_segments : m_super * shape.n_tiles;
  if (total <= 0) return cudaSuccess;
  const int grid = total < device_sm_count() ? total : device_sm_count();
  kernel<<<grid, gemm_threads<BLOCK_N, Epi>(), kSmem, stream>>>(ta, tb, shape, ep);
  return cudaGetLastError();
}

template <int BLOCK_N, int BLOCK_K, int STAGES, int KB, class Epi>
inline cudaError_t launch_gemm_bres(const CUtensorMap& ta, const CUtensorMap& tb, const GemmShape& shape,
                                    const typename Epi::Params& ep, cudaStream_t stream) {
  using L = GemmBresSmemLayout<BLOCK_N, BLOCK_K, STAGES, KB>;
  constexpr int kSmem = L::kTotal + Epi::kExtraSmemBytes;
  static_assert(kSmem <= 232448, "shared memory budget");
  if (shape.k_blocks != KB) return cudaErrorInvalidValue;
  auto kernel = gemm_bres_tcgen05_kernel<BLOCK_N, BLOCK_K, STAGES, KB, Epi>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const long long total = (long long)shape.m_tiles * shape.n_tiles;
  if (total <= 0) return cudaSuccess;
  const int grid = total < device_sm_count() ? (int)total : device_sm_count();
  kernel<<<grid, gemm_threads<BLOCK_N, Epi>(), kSmem, stream>>>(ta, tb, shape, ep);
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
  s.h_mul = 0;
  s.row_resident = 0;
  s.n_segments = 1;
  s.b_prefetch_tiles = 0;
  return s;
}

// 3x3 / stride-2 / pad-1 convolution read through a stride-2 TMA box: K block kb = tap (kh, kw) x channel slice
inline void conv_taps_strided(GemmShape* s, int cblocks_per_tap) {
  s->h_mul = 2;
  for (int kb = 0; kb < 9 * cblocks_per_tap; ++kb) {
    const int tap = kb / cblocks_per_tap, kh = tap / 3, kw = tap % 3;
    s->tap_c_blk[kb] = (signed char)(kb % cblocks_per_tap);
    s->tap_w[kb] = (signed char)(kw - 1);
    s->tap_h[kb] = (signed char)(kh - 1);
  }
}

// The same convolution over a space-to-depth input [frame][H/2][W/2][(sh*2+sw)*C + c]: it becomes a 2x2 /
// stride-1 window over cells, every tap reads ONE dense channel slice of a cell (no strided traversal).
// tap k in {0,1,2} -> input row 2*o + k - 1 = cell (o + d), sub-row s with (d, s) = (-1,1), (0,0), (0,1).
inline void conv_taps_space_to_depth(GemmShape* s, int cblocks_per_subpixel) {
  s->h_mul = 1;
  const int d[3] = {-1, 0, 0}, sub[3] = {1, 0, 1};
  int kb = 0;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw)
      for (int cb = 0; cb < cblocks_per_subpixel; ++cb, ++kb) {
        s->tap_c_blk[kb] = (signed char)((sub[kh] * 2 + sub[kw]) * cblocks_per_subpixel + cb);
        s->tap_w[kb] = (signed char)d[kw];
        s->tap_h[kb] = (signed char)d[kh];
      }
}

}  // namespace vfp
