// Fused frame-encoder stem: conv1 (3->32, k5 s2, mma.sync) feeding conv2 (32->64, k3 s2, tcgen05) through SHARED
// MEMORY. conv1's output is the largest tensor of the whole forward (64 KB per frame: written once and read once it
// is 84 GB of the ~200 GB of HBM traffic of 10k clips), and both stand-alone kernels sit on the HBM roofline. Here it
// never leaves the SM: the frame goes in (12-48 KB), conv2's output comes out (32 KB).
//
// One CTA per SM, persistent over frames, 30 warps (960 threads, 64 registers each - every role is written to fit):
//   warp 0       conv2 weights -> smem by TMA (once), TMEM allocation
//   warp 1       UMMA issuer (one lane)
//   warps 2-5    conv2 epilogue: TMEM -> +bias, ReLU -> bf16 -> swizzled staging -> TMA store
//   warps 6-21   conv1 producers: mma.sync on the staged frame (B fragments re-read from shared memory to save
//                registers), results written straight into conv2's A-operand buffers (K-major SWIZZLE_128B)
//   warps 22-29  frame loaders: global -> padded HWC bf16 tile (double buffered)
// The stand-alone conv1 kernel needs ~20 resident warps per SM to hide its latencies; 16 producer warps is what
// fits next to the other roles.
//
// Work unit = half a frame = 128 conv2 output pixels (cell rows 8*hf .. 8*hf+7 of the 16x16 output). conv1's output is
// kept space-to-depth (cell = 2x2 pixels, sub-pixel (sh, sw)), which turns conv2's stride-2 taps into row/column
// shifts of whole cells: tap kh -> (dh, sh) = (-1,1),(0,0),(0,1), same for kw. Three buffers per unit, each 9 cell rows
// (1 halo + 8) x 16 cells x 128 B (64 channels = sub-rows sh=0,1 of one sw):
//   AL0: sub-column sw=0          -> taps kw=1        AL1: sw=1 -> taps kw=2
//   SH1: AL1 shifted right by one cell (column 0 = zero padding) -> taps kw=0 (input column 2*ow-1)
// A dh=-1 tap is the same buffer addressed one cell row (16 rows = 2 swizzle atoms) higher, so every UMMA descriptor
// stays atom-aligned. K = 3 buffers x (4 K-steps for dh=0 + 2 K-steps (sh=1 only) for dh=-1) x 16 = 288 = 9*32: no
// padded K at all. The 18 dependent UMMAs of a unit alternate between two accumulators (split-K) that the epilogue adds.
#pragma once
#include "conv1_kernel.cuh"
#include "epilogues.cuh"

namespace vfp {

constexpr int kStemThreads = 960;
constexpr int kStemProducerWarp0 = 6, kStemProducerWarps = 16;
constexpr int kStemLoaderWarp0 = 22, kStemLoaderWarps = 8;
constexpr int kStemC1WBytes = 5 * 4 * 32 * 8;      // conv1 B fragments: [kh][n-tile][lane] x 8 B
constexpr int kStemABuf = 9 * 16 * 128;            // one of AL0 / AL1 / SH1: 18432 B
constexpr int kStemUnitBytes = 3 * kStemABuf;      // 55296 B
constexpr int kStemTileBytes = (kC1SmemElems * 2 + 127) / 128 * 128;

struct StemSmem {
  static constexpr int kC1 = 0;                                   // 2 units
  static constexpr int kW = kC1 + 2 * kStemUnitBytes;             // 6 weight tiles of 64 rows x 128 B
  static constexpr int kStage = kW + 6 * 8192;                    // 4 epilogue warps x 2 KB
  static constexpr int kTile = kStage + 4 * 2048;                 // 2 frame tiles
  static constexpr int kC1W = kTile + 2 * kStemTileBytes;
  static constexpr int kBars = kC1W + kStemC1WBytes;
  static constexpr int kTotal = kBars + 256 + 1024;
};
static_assert(StemSmem::kTotal <= 232448, "stem kernel shared memory");
static_assert(StemSmem::kW % 1024 == 0 && StemSmem::kStage % 1024 == 0, "swizzled regions must be 1024-byte aligned");

struct StemParams {
  alignas(64) CUtensorMap tmap_w;    // conv2 weights [64][384] bf16 (fused K order), box 64 rows x 64 K, SWIZZLE_128B
  alignas(64) CUtensorMap tmap_out;  // conv2 output [frames*256][64] bf16, box 32 x 32, SWIZZLE_64B
  const void* frames;
  int frame_dtype;
  long long n_frames;
  const uint32_t* c1_wpack;  // conv1 B fragments
  const float* c1_bias;
  const float* c2_bias;
};

__global__ void __launch_bounds__(kStemThreads, 1) stem_fused_kernel(const __grid_constant__ StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* c1buf = smem + StemSmem::kC1;
  uint8_t* wbuf = smem + StemSmem::kW;
  uint8_t* stagebuf = smem + StemSmem::kStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + StemSmem::kBars);
  uint64_t* w_full = bars;            // [1]
  uint64_t* c1_full = bars + 1;       // [2] producers -> UMMA
  uint64_t* c1_empty = bars + 3;      // [2] UMMA -> producers
  uint64_t* acc_full = bars + 5;      // [2] UMMA -> epilogue
  uint64_t* acc_empty = bars + 7;     // [2] epilogue -> UMMA
  uint64_t* tile_full = bars + 9;     // [2] loaders -> producers
  uint64_t* tile_empty = bars + 11;   // [2] producers -> loaders
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  // frames of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const long long n_local = (p.n_frames > blockIdx.x) ? (p.n_frames - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  // zero everything that is read before it is written: c1 halo rows / SH1 column 0, the tile halos
  for (int i = tid; i < (2 * kStemUnitBytes) / 16; i += kStemThreads) reinterpret_cast<uint4*>(c1buf)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < (2 * kStemTileBytes) / 16; i += kStemThreads) reinterpret_cast<uint4*>(smem + StemSmem::kTile)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < kStemC1WBytes / 8; i += kStemThreads)
    reinterpret_cast<uint2*>(smem + StemSmem::kC1W)[i] = __ldg(reinterpret_cast<const uint2*>(p.c1_wpack) + i);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_w);
    tma_prefetch_desc(&p.tmap_out);
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&c1_full[i], kStemProducerWarps);
      mbar_init(&c1_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
      mbar_init(&tile_full[i], kStemLoaderWarps);
      mbar_init(&tile_empty[i], kStemProducerWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);  // 2 units x 2 split-K accumulators x 64 columns
    tmem_relinquish();
  }
  fence_proxy_async_smem();  // the zero fill is read by the UMMA (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ conv2 weights, once ------------------------------
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, 6 * 8192);
      for (int kb = 0; kb < 6; ++kb) tma_load_2d(&p.tmap_w, w_full, wbuf + kb * 8192, kb * 64, 0);
    }
  } else if (warp == 1) {
    // ------------------------------ UMMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
      mbar_wait(w_full, 0);
      for (long long u = 0; u < 2 * n_local; ++u) {
        const int b = (int)(u & 1);
        const uint32_t ph = (uint32_t)((u >> 1) & 1);
        mbar_wait(&acc_empty[b], ph ^ 1);
        mbar_wait(&c1_full[b], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(c1buf + b * kStemUnitBytes);
        const uint32_t w_base = smem_u32(wbuf);
        const uint32_t d_base = tmem_base + b * 128;
        int step = 0;
#pragma unroll
        for (int dh = 0; dh >= -1; --dh) {
#pragma unroll
          for (int x = 0; x < 3; ++x) {
            // dh = 0: rows start one cell row (16 rows) below the halo row, all 4 K-steps; dh = -1: from the halo row,
            // only the sh = 1 half of the 64 channels (K-steps 2, 3)
            const uint64_t adesc = umma_smem_desc_kmajor<128>(a_base + x * kStemABuf + (dh == 0 ? 16 * 128 : 0));
            const uint64_t bdesc = umma_smem_desc_kmajor<128>(w_base + ((dh == 0 ? 0 : 3) + x) * 8192);
#pragma unroll
            for (int k = (dh == 0 ? 0 : 2); k < 4; ++k, ++step)
              umma_bf16(d_base + (step & 1) * 64, adesc + 2 * k, bdesc + 2 * k, idesc, step >= 2 ? 1u : 0u);
          }
        }
        umma_commit(&c1_empty[b]);
        umma_commit(&acc_full[b]);
      }
    }
  } else if (warp < kStemProducerWarp0) {
    // ------------------------------ conv2 epilogue ------------------------------
    const int quarter = warp & 3;
    uint8_t* dst = stagebuf + (warp - 2) * 2048;
    for (long long u = 0; u < 2 * n_local; ++u) {
      const int b = (int)(u & 1);
      const uint32_t ph = (uint32_t)((u >> 1) & 1);
      const long long frame = blockIdx.x + (u >> 1) * gridDim.x;
      mbar_wait_relaxed(&acc_full[b], ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + b * 128;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint8_t* r0 = dst + lane * 64;
        const int sw = (lane >> 1) & 3;
        if (lane == 0) tma_store_wait_read<0>();  // the previous store has finished reading the staging tile
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // 16 columns at a time keeps this role under 64 registers
          uint32_t v0[16], v1[16];
          tmem_ld_32x16(taddr + c0 + 16 * h, v0);
          tmem_ld_32x16(taddr + 64 + c0 + 16 * h, v1);
          tmem_ld_wait();
          float x[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.c2_bias + c0 + 16 * h + i));
            x[i] = fmaxf(__uint_as_float(v0[i]) + __uint_as_float(v1[i]) + bb.x, 0.0f);
            x[i + 1] = fmaxf(__uint_as_float(v0[i + 1]) + __uint_as_float(v1[i + 1]) + bb.y, 0.0f);
            x[i + 2] = fmaxf(__uint_as_float(v0[i + 2]) + __uint_as_float(v1[i + 2]) + bb.z, 0.0f);
            x[i + 3] = fmaxf(__uint_as_float(v0[i + 3]) + __uint_as_float(v1[i + 3]) + bb.w, 0.0f);
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint4 q;
            q.x = pack_bf16x2(x[8 * c + 0], x[8 * c + 1]);
            q.y = pack_bf16x2(x[8 * c + 2], x[8 * c + 3]);
            q.z = pack_bf16x2(x[8 * c + 4], x[8 * c + 5]);
            q.w = pack_bf16x2(x[8 * c + 6], x[8 * c + 7]);
            *reinterpret_cast<uint4*>(r0 + (((2 * h + c) ^ sw) << 4)) = q;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&p.tmap_out, dst, c0, (int)(frame * 256 + b * 128 + quarter * 32));
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
    }
    if (lane == 0) tma_store_wait_read<0>();
  } else if (warp < kStemLoaderWarp0) {
    // ------------------------------ conv1 producers ------------------------------
    const int pw = warp - kStemProducerWarp0;
    const int g = lane >> 2, tig = lane & 3;
    const uint2* wfrag = reinterpret_cast<const uint2*>(smem + StemSmem::kC1W) + lane;  // [(kh*4+nt)*32 + lane]
    float bia[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      bia[nt][0] = __ldg(p.c1_bias + nt * 8 + 2 * tig);
      bia[nt][1] = __ldg(p.c1_bias + nt * 8 + 2 * tig + 1);
    }
    for (long long li = 0; li < n_local; ++li) {
      const int t = (int)(li & 1);
      const uint32_t* tile32 = reinterpret_cast<const uint32_t*>(smem + StemSmem::kTile + t * kStemTileBytes);
      mbar_wait_relaxed(&tile_full[t], (uint32_t)((li >> 1) & 1));
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        uint8_t* unit = c1buf + hf * kStemUnitBytes;
        mbar_wait_relaxed(&c1_empty[hf], (uint32_t)((li & 1) ^ 1));
        // conv1 output rows of this unit: hf = 0 -> 0..15 ; hf = 1 -> 14..31 (rows 14, 15 are the halo cell row,
        // recomputed instead of shared with the other buffer). m-tile = one row x 16 columns.
        const int row_first = hf == 0 ? 0 : 14;
        const int n_mtiles = (hf == 0 ? 16 : 18) * 2;
#pragma unroll 1
        for (int mt = pw; mt < n_mtiles; mt += kStemProducerWarps) {
          const int oh = row_first + (mt >> 1);
          const int ow0 = (mt & 1) * 16;
          float acc[4][4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[nt][i] = 0.0f;
#pragma unroll
          for (int kh = 0; kh < 5; ++kh) {
            const int base_lo = (((2 * oh + kh) * kC1PadW + 2 * (ow0 + g)) * 3) >> 1;
            const int base_hi = base_lo + 24;
            uint32_t a[4];
            a[0] = tile32[base_lo + tig];
            a[1] = tile32[base_hi + tig];
            a[2] = tile32[base_lo + tig + 4];
            a[3] = tile32[base_hi + tig + 4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              const uint2 wv = wfrag[(kh * 4 + nt) * 32];
              const uint32_t bq[2] = {wv.x, wv.y};
              mma_bf16_16816(acc[nt], a, bq);
            }
          }
          // scatter into the A-operand buffers. pixel (oh, ow): cell (oh/2, ow/2), sub (sh, sw) = (oh%2, ow%2);
          // buffer row = (cell_row - 8*hf + 1) * 16 + cell_col; 16-byte chunk j = sh*4 + nt, stored at j ^ (row % 8).
          const int sh = oh & 1;
          const int rowblk = (oh >> 1) - 8 * hf + 1;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int ow = ow0 + g + 8 * half;
            const int cx = ow >> 1;
            const int row = rowblk * 16 + cx;
            uint8_t* al = unit + (ow & 1) * kStemABuf + row * 128 + tig * 4;
            uint8_t* shf = unit + 2 * kStemABuf + (row + 1) * 128 + tig * 4;
            const bool to_shift = (ow & 1) && cx < 15;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              const float v0 = fmaxf(acc[nt][2 * half] + bia[nt][0], 0.0f);
              const float v1 = fmaxf(acc[nt][2 * half + 1] + bia[nt][1], 0.0f);
              const uint32_t pk = pack_bf16x2(v0, v1);
              const int j = sh * 4 + nt;
              *reinterpret_cast<uint32_t*>(al + ((j ^ (row & 7)) << 4)) = pk;
              if (to_shift) *reinterpret_cast<uint32_t*>(shf + ((j ^ ((row + 1) & 7)) << 4)) = pk;
            }
          }
        }
        fence_proxy_async_smem();  // generic-proxy writes above -> visible to the UMMA reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&c1_full[hf]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&tile_empty[t]);
    }
  } else {
    // ------------------------------ frame loaders ------------------------------
    const int ltid = tid - kStemLoaderWarp0 * 32;
    for (long long li = 0; li < n_local; ++li) {
      const int t = (int)(li & 1);
      mbar_wait_relaxed(&tile_empty[t], (uint32_t)(((li >> 1) & 1) ^ 1));
      const long long frame = blockIdx.x + li * gridDim.x;
      stage_frame_hwc(p.frames, p.frame_dtype, frame, reinterpret_cast<__nv_bfloat16*>(smem + StemSmem::kTile + t * kStemTileBytes), ltid,
                      kStemLoaderWarps * 32);
      __syncwarp();
      if (lane == 0) mbar_arrive(&tile_full[t]);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace vfp
