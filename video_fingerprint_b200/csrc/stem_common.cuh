// Shared pieces of the fused frame-encoder stem (stem_ts_kernel.cuh): conv1 (3->32, k5 s2) feeding conv2 (32->64, k3 s2,
// tcgen05) through SHARED MEMORY. conv1's output is the largest tensor of the whole forward (64 KB per frame: written once
// and read once it would be 84 GB of HBM traffic per 10k clips); in the fused kernel it never leaves the SM: the frame
// goes in (12-24 KB), conv2's output comes out (32 KB). This header holds the layout both convs agree on, the raw-frame ->
// HWC transposition and small helpers. (The round-1 variant that ran conv1 on mma.sync was removed: the TS-mode tcgen05
// kernel superseded it and is covered by the parity, guard-zone and repeatability tests.)
//
// Work unit = half a frame = 128 conv2 output pixels (cell rows 8*hf .. 8*hf+7 of the 16x16 output). conv1's output is
// kept space-to-depth (cell = 2x2 pixels, sub-pixel (sh, sw)), which turns conv2's stride-2 taps into shifts by whole
// cells: tap kh -> (dh, sh) = (-1,1),(0,0),(0,1), same for kw. Two buffers per unit, each 9 cell rows (1 halo + 8) x
// 16 cells x 128 B (64 channels = both sub-rows sh of one sub-column sw):
//   AL0: sw = 0, K order [sh=0 | sh=1]  -> taps kw = 1
//   AL1: sw = 1, K order [sh=1 | sh=0]  -> taps kw = 2, and taps kw = 0 of the cell to the RIGHT
// (the opposite K orders make the two pixels a quarter-warp stores land in different bank halves).
// A dh=-1 tap is the same buffer addressed one cell row (16 rows = 2 swizzle atoms) higher, so every UMMA descriptor
// stays atom-aligned. The kw = 0 taps need the cell to the LEFT, a one-row shift that a swizzled descriptor cannot
// express; instead they accumulate UNSHIFTED into their own accumulators and the epilogue adds row r-1 into row r
// (one warp shuffle; cell column 0 gets the zero padding). K = 288 real, no padded K at all. The kw = 2 and kw = 0
// taps read the SAME A rows (AL1), so each such pair is ONE N = 128 UMMA against the stacked weights [W_kw2 ; W_kw0]
// (columns 0-63 regular, 64-127 shifted): an SS-mode M128 K16 UMMA costs ~49 cycles at N = 64 but only ~65 at
// N = 128 (tests/cuda/microbench_tensor.cu), so a unit takes 12 UMMAs / ~690 cycles instead of 18 / ~880.
#pragma once
#include "conv1_kernel.cuh"
#include "epilogues.cuh"

namespace vfp {

constexpr int kStemThreads = 768;
constexpr int kStemEpiWarp0 = 4, kStemEpiWarps = 8;
constexpr int kStemABuf = 9 * 16 * 128;            // one of AL0 / AL1: 18432 B
constexpr int kStemUnitBytes = 2 * kStemABuf;      // 36864 B
constexpr int kStemWBlocks = 5;                    // conv2 weight K blocks of 64 (see vfp_weights_create)
constexpr int kStemRawSlotBytes = 8192;
// HWC tile: rows -2..64, 72 pixels per row (8 zero pixels, then columns 0..63; column 64 of a row IS the first zero
// pixel of the next row), 3 channels interleaved: element ((h+2)*72 + w+8)*3 + c
constexpr int kStemTilePitch = 72;
constexpr int kStemTileBytes = 29184;
static_assert((67 * kStemTilePitch + 8) * 6 <= kStemTileBytes, "HWC tile");

template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// relu + round to bf16 of two floats in one instruction; `lo` lands in the low half
__device__ __forceinline__ uint32_t relu_pack_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(src)), "r"(bytes) : "memory");
}
// u8 -> value/255 as bf16 pairs
__device__ __forceinline__ uint32_t u8x2_to_bf16x2(uint32_t lo, uint32_t hi) {
  return pack_bf16x2((float)lo * (1.0f / 255.0f), (float)hi * (1.0f / 255.0f));
}

// Transposer work of one frame: raw planes (three ring slots pl[0..2]; for decoder-layout u8 just three 4 KB chunks of the
// byte stream) -> zero-haloed HWC bf16 tile. Called by `n_threads` threads with ltid = 0 .. n_threads-1.
__device__ __forceinline__ void stem_transpose_frame(const uint8_t* const (&pl)[3], uint8_t* tile, int frame_dtype, int ltid,
                                                     int n_threads) {
  if (frame_dtype == kFrameBF16) {
    // item = 8 pixels of one row: 3 x 16 B in (one per plane), 48 contiguous bytes out
    for (int i = ltid; i < 512; i += n_threads) {
      const int h = i >> 3, w0 = (i & 7) * 8;
      const uint4 a = *reinterpret_cast<const uint4*>(pl[0] + (h * 64 + w0) * 2);
      const uint4 b = *reinterpret_cast<const uint4*>(pl[1] + (h * 64 + w0) * 2);
      const uint4 c = *reinterpret_cast<const uint4*>(pl[2] + (h * 64 + w0) * 2);
      uint4* dst = reinterpret_cast<uint4*>(tile + (h + 2) * (kStemTilePitch * 6) + (w0 + 8) * 6);
      // words of the HWC stream: (a0 b0)(c0 a1)(b1 c1) (a2 b2)(c2 a3)(b3 c3) ...
      uint4 o0, o1, o2;
      o0.x = __byte_perm(a.x, b.x, 0x5410); o0.y = __byte_perm(c.x, a.x, 0x7610); o0.z = __byte_perm(b.x, c.x, 0x7632);
      o0.w = __byte_perm(a.y, b.y, 0x5410); o1.x = __byte_perm(c.y, a.y, 0x7610); o1.y = __byte_perm(b.y, c.y, 0x7632);
      o1.z = __byte_perm(a.z, b.z, 0x5410); o1.w = __byte_perm(c.z, a.z, 0x7610); o2.x = __byte_perm(b.z, c.z, 0x7632);
      o2.y = __byte_perm(a.w, b.w, 0x5410); o2.z = __byte_perm(c.w, a.w, 0x7610); o2.w = __byte_perm(b.w, c.w, 0x7632);
      dst[0] = o0; dst[1] = o1; dst[2] = o2;
    }
  } else if (frame_dtype == kFrameU8) {
    for (int i = ltid; i < 512; i += n_threads) {
      const int h = i >> 3, w0 = (i & 7) * 8;
      const uint2 a = *reinterpret_cast<const uint2*>(pl[0] + h * 64 + w0);
      const uint2 b = *reinterpret_cast<const uint2*>(pl[1] + h * 64 + w0);
      const uint2 c = *reinterpret_cast<const uint2*>(pl[2] + h * 64 + w0);
      uint4* dst = reinterpret_cast<uint4*>(tile + (h + 2) * (kStemTilePitch * 6) + (w0 + 8) * 6);
      uint32_t o[12];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t av = half ? a.y : a.x, bv = half ? b.y : b.x, cv = half ? c.y : c.x;
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {  // pixel pair (2pp, 2pp+1) of this half
          const uint32_t a0 = (av >> (16 * pp)) & 0xFF, a1 = (av >> (16 * pp + 8)) & 0xFF;
          const uint32_t b0 = (bv >> (16 * pp)) & 0xFF, b1 = (bv >> (16 * pp + 8)) & 0xFF;
          const uint32_t c0 = (cv >> (16 * pp)) & 0xFF, c1 = (cv >> (16 * pp + 8)) & 0xFF;
          o[half * 6 + pp * 3 + 0] = u8x2_to_bf16x2(a0, b0);
          o[half * 6 + pp * 3 + 1] = u8x2_to_bf16x2(c0, a1);
          o[half * 6 + pp * 3 + 2] = u8x2_to_bf16x2(b1, c1);
        }
      }
      dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
      dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
      dst[2] = make_uint4(o[8], o[9], o[10], o[11]);
    }
  } else {
    // decoder layout (H, W, 3) u8: the byte stream is already HWC; item = 16 bytes -> 32 bytes out
    for (int i = ltid; i < 768; i += n_threads) {
      const int byte0 = i * 16;
      const int chunk = byte0 >> 12;
      const uint4 q = *reinterpret_cast<const uint4*>((chunk == 0 ? pl[0] : chunk == 1 ? pl[1] : pl[2]) + (byte0 & 4095));
      const int h = byte0 / 192, off = byte0 - h * 192;
      uint4* dst = reinterpret_cast<uint4*>(tile + (h + 2) * (kStemTilePitch * 6) + 48 + 2 * off);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[2 * j] = u8x2_to_bf16x2(w[j] & 0xFF, (w[j] >> 8) & 0xFF);
        o[2 * j + 1] = u8x2_to_bf16x2((w[j] >> 16) & 0xFF, w[j] >> 24);
      }
      dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
      dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
  }
}

}  // namespace vfp
