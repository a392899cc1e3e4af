// Fused frame-encoder stem: conv1 (3->32, k5 s2, mma.sync) feeding conv2 (32->64, k3 s2, tcgen05) through SHARED
// MEMORY. conv1's output is the largest tensor of the whole forward (64 KB per frame: written once and read once it
// is 84 GB of the ~200 GB of HBM traffic of 10k clips), and both stand-alone kernels sit on the HBM / L2->SM fill
// roofline. Here it never leaves the SM: the frame goes in (12-24 KB), conv2's output comes out (32 KB).
//
// One CTA per SM, persistent over frames, 24 warps = 6 warpgroups, registers re-balanced with setmaxnreg:
//   WG0   warp 0        bulk-copy issuer: raw frame planes -> smem ring (cp.async.bulk, 5 slots = 1.7 frames of
//                       prefetch), conv2 weights -> smem once (TMA)
//         warp 1        TMEM allocation + UMMA issuer (one lane)
//         warps 2-3     transposers: raw planar (or decoder-layout) frame -> zero-haloed HWC bf16 tile (double buffered)
//   WG1-2 warps 4-11    conv2 epilogue: TMEM -> split-K sum, shifted-tap add, bias, ReLU -> bf16 -> staging -> TMA store
//   WG3-5 warps 12-23   conv1 producers: mma.sync on the HWC tile, result written straight into conv2's A-operand
//                       buffers (K-major SWIZZLE_128B) with conflict-free 16-byte stores
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
//
// A producer task = two vertically adjacent conv1 output rows x 16 columns (= one cell row x 8 cells, both sh): the
// 7 input rows they share are read once. Fragment rows are pixels ow0+g / ow0+g+8 and the conv1 weight columns are
// permuted so that a thread ends up with 8 CONSECUTIVE channels of each of its pixels = one 16-byte store.
#pragma once
#include "conv1_kernel.cuh"
#include "epilogues.cuh"

namespace vfp {

constexpr int kStemThreads = 768;
constexpr int kStemXposeWarp0 = 2, kStemXposeWarps = 2;
constexpr int kStemEpiWarp0 = 4, kStemEpiWarps = 8;
constexpr int kStemProdWarp0 = 12, kStemProdWarps = 12;
constexpr int kStemTasks0 = 16, kStemTasks1 = 18;  // producer tasks of unit hf=0 / hf=1 (hf=1 recomputes the halo row pair)
constexpr int kStemTasks = kStemTasks0 + kStemTasks1;
constexpr int kStemABuf = 9 * 16 * 128;            // one of AL0 / AL1: 18432 B
constexpr int kStemUnitBytes = 2 * kStemABuf;      // 36864 B
constexpr int kStemWBlocks = 5;                    // conv2 weight K blocks of 64 (see vfp_weights_create)
constexpr int kStemRawSlots = 5, kStemRawSlotBytes = 8192;
// HWC tile: rows -2..64, 72 pixels per row (8 zero pixels, then columns 0..63; column 64 of a row IS the first zero
// pixel of the next row), 3 channels interleaved: element ((h+2)*72 + w+8)*3 + c
constexpr int kStemTilePitch = 72;
constexpr int kStemTileBytes = 29184;
static_assert((67 * kStemTilePitch + 8) * 6 <= kStemTileBytes, "HWC tile");

struct StemSmem {
  static constexpr int kC1 = 0;                                            // 2 units
  static constexpr int kW = kC1 + 2 * kStemUnitBytes;                      // 5 weight tiles of 64 rows x 128 B
  static constexpr int kStage = kW + kStemWBlocks * 8192;                  // 8 epilogue warps x 2 KB
  static constexpr int kRaw = kStage + kStemEpiWarps * 2048;               // raw frame planes
  static constexpr int kTile = kRaw + kStemRawSlots * kStemRawSlotBytes;   // 2 HWC tiles
  static constexpr int kBars = kTile + 2 * kStemTileBytes;
  static constexpr int kTotal = kBars + 512 + 1024;
};
static_assert(StemSmem::kTotal <= 232448, "stem kernel shared memory");
static_assert(StemSmem::kW % 1024 == 0 && StemSmem::kStage % 1024 == 0 && StemSmem::kRaw % 1024 == 0, "alignment");

struct StemParams {
  alignas(64) CUtensorMap tmap_w;    // conv2 weights [64][320] bf16 (fused K order), box 64 rows x 64 K, SWIZZLE_128B
  alignas(64) CUtensorMap tmap_out;  // conv2 output [frames*256][64] bf16, box 32 x 32, SWIZZLE_64B
  const void* frames;                // planar u8 / planar bf16 / decoder-layout u8, 16-byte aligned
  int frame_dtype;
  long long n_frames;
  const uint32_t* c1_wpack;  // conv1 B fragments, output channels permuted (column n of tile nt = channel 8*(n>>1) + 2*nt + (n&1))
  const float* c1_bias;      // [32] natural channel order
  const float* c2_bias;
};

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

__global__ void __launch_bounds__(kStemThreads, 1) stem_fused_kernel(const __grid_constant__ StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  // keep the pointer arithmetic on the __shared__ array (no integer round trip) so accesses compile to LDS / STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* c1buf = smem + StemSmem::kC1;
  uint8_t* wbuf = smem + StemSmem::kW;
  uint8_t* stagebuf = smem + StemSmem::kStage;
  uint8_t* rawbuf = smem + StemSmem::kRaw;
  uint8_t* tilebuf = smem + StemSmem::kTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + StemSmem::kBars);
  uint64_t* w_full = bars;            // [1]
  uint64_t* c1_full = bars + 1;       // [2] producers -> UMMA (one arrive per producer task)
  uint64_t* c1_empty = bars + 3;      // [2] UMMA -> producers
  uint64_t* acc_full = bars + 5;      // [2] UMMA -> epilogue
  uint64_t* acc_empty = bars + 7;     // [2] epilogue -> UMMA
  uint64_t* tile_full = bars + 9;     // [2] transposers -> producers
  uint64_t* tile_empty = bars + 11;   // [2] producers -> transposers
  uint64_t* raw_full = bars + 13;     // [5] bulk copy -> transposers
  uint64_t* raw_empty = bars + 18;    // [5] transposers -> bulk copy issuer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  // frames of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const long long n_local = (p.n_frames > blockIdx.x) ? (p.n_frames - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  // zero everything that is read before it is written: the c1 halo row blocks, the tile halos
  for (int i = tid; i < (2 * kStemUnitBytes) / 16; i += kStemThreads) reinterpret_cast<uint4*>(c1buf)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < (2 * kStemTileBytes) / 16; i += kStemThreads) reinterpret_cast<uint4*>(tilebuf)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_w);
    tma_prefetch_desc(&p.tmap_out);
    mbar_init(w_full, 1);
    mbar_init(&c1_full[0], kStemTasks0);
    mbar_init(&c1_full[1], kStemTasks1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&c1_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kStemEpiWarps);
      mbar_init(&tile_full[i], kStemXposeWarps);
      mbar_init(&tile_empty[i], kStemProdWarps);
    }
    for (int i = 0; i < kStemRawSlots; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_empty[i], kStemXposeWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);  // 2 units x 192 accumulator columns
    tmem_relinquish();
  }
  fence_proxy_async_smem();  // the zero fill is read by the UMMA (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool is_bf16 = p.frame_dtype == kFrameBF16;
  const uint32_t plane_bytes = is_bf16 ? 8192u : 4096u;  // a third of a frame (for decoder-layout u8: just a chunk)

  // setmaxnreg re-balances inside the CTA's own allocation (768 threads x 80 registers = 480 per warpgroup-thread):
  // 56 + 2*56 + 3*104 = 480. An inc that the decs do not cover blocks forever.
  if (warp < kStemEpiWarp0) {
    setmaxnreg_dec<56>();
    if (warp == 0) {
      // ------------------------------ weights once, then the raw-plane ring ------------------------------
      if (lane == 0) {
        mbar_arrive_expect_tx(w_full, kStemWBlocks * 8192);
        for (int kb = 0; kb < kStemWBlocks; ++kb) tma_load_2d(&p.tmap_w, w_full, wbuf + kb * 8192, kb * 64, 0);
        const uint8_t* base = static_cast<const uint8_t*>(p.frames);
        int slot = 0;
        uint32_t ph = 0;
        for (long long li = 0; li < n_local; ++li) {
          const uint8_t* src = base + (size_t)(blockIdx.x + li * gridDim.x) * (3 * plane_bytes);
          if (li + 2 < n_local) bulk_prefetch_l2(base + (size_t)(blockIdx.x + (li + 2) * gridDim.x) * (3 * plane_bytes), 3 * plane_bytes);
          for (int c = 0; c < 3; ++c) {
            mbar_wait_relaxed(&raw_empty[slot], ph ^ 1);
            mbar_arrive_expect_tx(&raw_full[slot], plane_bytes);
            bulk_copy_g2s(rawbuf + slot * kStemRawSlotBytes, src + c * plane_bytes, plane_bytes, &raw_full[slot]);
            if (++slot == kStemRawSlots) { slot = 0; ph ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ------------------------------ UMMA issuer ------------------------------
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
        mbar_wait_relaxed(w_full, 0);
        const uint32_t w_base = smem_u32(wbuf);
        for (long long u = 0; u < 2 * n_local; ++u) {
          const int b = (int)(u & 1);
          const uint32_t ph = (uint32_t)((u >> 1) & 1);
          mbar_wait_relaxed(&acc_empty[b], ph ^ 1);
          mbar_wait_relaxed(&c1_full[b], ph);
          tc_fence_after();
          const uint32_t al0 = smem_u32(c1buf + b * kStemUnitBytes), al1 = al0 + kStemABuf;
          const uint32_t d_base = tmem_base + b * 192;
          // accumulator R (columns 0-63): kw = 1 taps; accumulator M (columns 64-191): kw = 2 taps in its first 64
          // columns, kw = 0 taps (to be shifted by the epilogue) in its last 64.
          //   G_A  AL0 dh=0  kw=1 (4 K-steps)            -> R     weights blk0
          //   G_BC AL1 dh=0  kw=2|0 (4 K-steps, N=128)   -> M     weights blk1 ; blk2 (stacked)
          //   G_D  AL0 dh=-1 kw=1 (sh=1 = A K-steps 2,3) -> R     weights blk3 K-steps 2,3
          //   G_EF AL1 dh=-1 kw=2|0 (A K-steps 0,1, N=128) -> M   weights blk3 ; blk4 K-steps 0,1 (stacked)
          constexpr uint32_t idesc128 = umma_idesc_bf16(128, 128);
          auto issue = [&](uint32_t a_addr, int ka, int wblk, int kb, bool wide, uint32_t accumulate) {
            const uint64_t adesc = umma_smem_desc_kmajor<128>(a_addr) + 2 * ka;
            const uint64_t bdesc = umma_smem_desc_kmajor<128>(w_base + wblk * 8192) + 2 * kb;
            umma_bf16(d_base + (wide ? 64 : 0), adesc, bdesc, wide ? idesc128 : idesc, accumulate);
          };
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            issue(al0 + 2048, k, 0, k, false, k > 0);   // G_A
            issue(al1 + 2048, k, 1, k, true, k > 0);    // G_BC
          }
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            issue(al0, 2 + k, 3, 2 + k, false, 1);      // G_D
            issue(al1, k, 3, k, true, 1);               // G_EF
          }
          umma_commit(&c1_empty[b]);
          umma_commit(&acc_full[b]);
        }
      }
    } else {
      // ------------------------------ transposers: raw planes -> HWC tile ------------------------------
      const int ltid = tid - kStemXposeWarp0 * 32;
      constexpr int kXThreads = kStemXposeWarps * 32;
      int slot0 = 0;       // ring slot of plane 0 of the current frame
      uint32_t n0 = 0;     // plane sequence number of plane 0 of the current frame
      for (long long li = 0; li < n_local; ++li) {
        const int t = (int)(li & 1);
        uint8_t* tile = tilebuf + t * kStemTileBytes;
        mbar_wait_relaxed(&tile_empty[t], (uint32_t)(((li >> 1) & 1) ^ 1));
        const uint8_t* pl[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const uint32_t n = n0 + c;
          const int s = (slot0 + c) % kStemRawSlots;
          mbar_wait_relaxed(&raw_full[s], (n / kStemRawSlots) & 1u);
          pl[c] = rawbuf + s * kStemRawSlotBytes;
        }
        stem_transpose_frame(pl, tile, p.frame_dtype, ltid, kXThreads);
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) mbar_arrive(&raw_empty[(slot0 + c) % kStemRawSlots]);
          mbar_arrive(&tile_full[t]);
        }
        slot0 = (slot0 + 3) % kStemRawSlots;
        n0 += 3;
      }
    }
  } else if (warp < kStemProdWarp0) {
    // ------------------------------ conv2 epilogue ------------------------------
    setmaxnreg_dec<56>();
    const int quarter = warp & 3;                      // TMEM lane quarter of this warp
    const int col_half = (warp - kStemEpiWarp0) >> 2;  // 32 of the 64 output channels
    uint8_t* dst = stagebuf + (warp - kStemEpiWarp0) * 2048;
    uint8_t* r0 = dst + lane * 64;
    const int sw = (lane >> 1) & 3;
    const bool first_cell = (lane & 15) == 0;  // cell column 0: the kw = 0 taps read the zero padding
    for (long long u = 0; u < 2 * n_local; ++u) {
      const int b = (int)(u & 1);
      const uint32_t ph = (uint32_t)((u >> 1) & 1);
      const long long frame = blockIdx.x + (u >> 1) * gridDim.x;
      mbar_wait_relaxed(&acc_full[b], ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + b * 192 + col_half * 32;
      if (lane == 0) tma_store_wait_read<0>();  // the previous store has finished reading the staging tile
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c) {  // 8 columns (= one 16-byte staging chunk) at a time keeps this role under 56 registers
        uint32_t v0[8], v1[8], v2[8];
        tmem_ld_32x8(taddr + 8 * c, v0);         // kw = 1 taps
        tmem_ld_32x8(taddr + 64 + 8 * c, v1);    // kw = 2 taps
        tmem_ld_32x8(taddr + 128 + 8 * c, v2);   // kw = 0 taps of the cell to the left
        tmem_ld_wait();
        uint32_t q[4];
#pragma unroll
        for (int i = 0; i < 8; i += 4) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.c2_bias + col_half * 32 + 8 * c + i));
          const float bias4[4] = {bb.x, bb.y, bb.z, bb.w};
          float x[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float s = __shfl_up_sync(0xffffffffu, __uint_as_float(v2[i + j]), 1);
            if (first_cell) s = 0.0f;
            x[j] = __uint_as_float(v0[i + j]) + __uint_as_float(v1[i + j]) + bias4[j] + s;
          }
          q[i / 2] = relu_pack_bf16x2(x[0], x[1]);
          q[i / 2 + 1] = relu_pack_bf16x2(x[2], x[3]);
        }
        *reinterpret_cast<uint4*>(r0 + ((c ^ sw) << 4)) = make_uint4(q[0], q[1], q[2], q[3]);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&acc_empty[b]);
        tma_store_2d(&p.tmap_out, dst, col_half * 32, (int)(frame * 256 + b * 128 + quarter * 32));
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait_read<0>();
  } else {
    // ------------------------------ conv1 producers ------------------------------
    setmaxnreg_inc<104>();
    const int pw = warp - kStemProdWarp0;
    const int g = lane >> 2, tig = lane & 3;
    uint32_t bfrag[5][4][2];
#pragma unroll
    for (int kh = 0; kh < 5; ++kh)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const uint2 w = __ldg(reinterpret_cast<const uint2*>(p.c1_wpack) + (kh * 4 + nt) * 32 + lane);
        bfrag[kh][nt][0] = w.x;
        bfrag[kh][nt][1] = w.y;
      }
    // this thread's accumulator columns (2tig, 2tig+1) of n-tile nt are channels 8*tig + 2*nt, +1
    float bia[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      bia[nt][0] = __ldg(p.c1_bias + 8 * tig + 2 * nt);
      bia[nt][1] = __ldg(p.c1_bias + 8 * tig + 2 * nt + 1);
    }
    // scatter constants: fragment row g = pixel ow0+g: sub-column sw = g & 1 (buffer AL0 / AL1), cell column (g >> 1);
    // row g+8 = pixel ow0+g+8: same sub-column, 4 cells to the right. A buffer row's swizzle phase is (cell column) & 7.
    const int psw = g & 1;
    const int cxl = g >> 1;
    const uint32_t ph_lo = (uint32_t)cxl, ph_hi = (uint32_t)(cxl + 4);
    const int n_tasks = (int)(kStemTasks * n_local);  // a conv pass is at most 16384 frames
    int cur_li = -1;
    const uint32_t* tile32 = nullptr;

    for (int tg = pw; tg < n_tasks; tg += kStemProdWarps) {
      const int li = tg / kStemTasks;
      const int task = tg - li * kStemTasks;
      if (li != cur_li) {
        if (cur_li >= 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&tile_empty[cur_li & 1]);
        }
        mbar_wait_relaxed(&tile_full[li & 1], (uint32_t)((li >> 1) & 1));
        tile32 = reinterpret_cast<const uint32_t*>(tilebuf + (li & 1) * kStemTileBytes);
        cur_li = li;
      }
      const int hf = task >= kStemTasks0 ? 1 : 0;
      const int tt = task - hf * kStemTasks0;
      // cell row of this task: hf = 0 -> 0..7 ; hf = 1 -> 7..15 (cell row 7 is the halo row of the second unit)
      const int cy = hf == 0 ? (tt >> 1) : 7 + (tt >> 1);
      const int ow0 = (tt & 1) * 16;
      uint8_t* unit = c1buf + hf * kStemUnitBytes;
      mbar_wait_relaxed(&c1_empty[hf], (uint32_t)((li & 1) ^ 1));

      float acc[2][4][4];
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          acc[s][nt][0] = bia[nt][0]; acc[s][nt][1] = bia[nt][1];
          acc[s][nt][2] = bia[nt][0]; acc[s][nt][3] = bia[nt][1];
        }
      // output rows oh = 2*cy + s read tile rows 2*oh + kh = 4*cy + 2*s + kh: 7 distinct rows R = 0..6,
      // row R feeds (s = 0, kh = R) and (s = 1, kh = R - 2). Word offset of (tile row r, pixel ow, k) =
      // r*108 + 3*ow + 9 + k/2.
      const uint32_t* trow = tile32 + (4 * cy) * 108 + 3 * (ow0 + g) + 9 + tig;
#pragma unroll
      for (int R = 0; R < 7; ++R) {
        uint32_t a[4];
        a[0] = trow[R * 108];
        a[1] = trow[R * 108 + 24];
        a[2] = trow[R * 108 + 4];
        a[3] = trow[R * 108 + 28];
        if (R <= 4) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[0][nt], a, bfrag[R <= 4 ? R : 0][nt]);
        }
        if (R >= 2) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[1][nt], a, bfrag[R >= 2 ? R - 2 : 0][nt]);
        }
      }
      // buffer row = (cy - 8*hf + 1) * 16 + cell column; 16-byte chunk (s ^ sw)*4 + tig stored at chunk ^ phase
      const int rowblk = cy - 8 * hf + 1;
      uint8_t* row_lo = unit + psw * kStemABuf + (rowblk * 16 + (ow0 >> 1) + cxl) * 128;
      uint8_t* row_hi = row_lo + 4 * 128;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const uint32_t chunk = (uint32_t)((s ^ psw) * 4 + tig);
        uint4 lo, hi;
        lo.x = relu_pack_bf16x2(acc[s][0][0], acc[s][0][1]); hi.x = relu_pack_bf16x2(acc[s][0][2], acc[s][0][3]);
        lo.y = relu_pack_bf16x2(acc[s][1][0], acc[s][1][1]); hi.y = relu_pack_bf16x2(acc[s][1][2], acc[s][1][3]);
        lo.z = relu_pack_bf16x2(acc[s][2][0], acc[s][2][1]); hi.z = relu_pack_bf16x2(acc[s][2][2], acc[s][2][3]);
        lo.w = relu_pack_bf16x2(acc[s][3][0], acc[s][3][1]); hi.w = relu_pack_bf16x2(acc[s][3][2], acc[s][3][3]);
        *reinterpret_cast<uint4*>(row_lo + ((chunk ^ ph_lo) << 4)) = lo;
        *reinterpret_cast<uint4*>(row_hi + ((chunk ^ (ph_hi & 7)) << 4)) = hi;
      }
      fence_proxy_async_smem();  // generic-proxy writes above -> visible to the UMMA reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&c1_full[hf]);
    }
    if (cur_li >= 0) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&tile_empty[cur_li & 1]);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace vfp
