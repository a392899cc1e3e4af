// On-device restatement of the scanner's frame preprocessing (/root/reference/fingerprint.py:186-214 _preprocess_frames):
// cv2.resize(frame, (new_w, new_h), INTER_AREA) so that the short side becomes 64, centre crop to 64 x 64. The uint8 result
// (decoder layout H, W, 3) feeds the fused stem directly (frame_dtype VFP_FRAME_U8_HWC, which also does the /255).
// Bit-exact with OpenCV 4.x. Down-scaling (both scale factors >= 1) has three code paths (up-scaling: see the end of the file):
//   * general (non-integer scale): float tables of (source index, weight) per destination index (computeResizeAreaTab);
//     per source row   buf = buf + S * alpha   over the x entries in order, then   sum = beta * buf   (first row) or
//     sum = sum + beta * buf; saturate_cast<uchar> = round half to even. Separate fmul / fadd, exactly OpenCV's order.
//   * integer scale factors: integer box sum, then round-half-even(float(sum) * float(1 / area));
//   * both factors == 2: (sum + 2) >> 2.
// One CTA per (output row, frame): every needed source row segment is staged through shared memory with coalesced loads
// (a 1080p frame is read exactly once: 3.5 MB for the 64 x 64 crop), thread (dx, c) then walks its own table entries.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vfp {

struct PreprocessParams {
  const uint8_t* src;    // [frames][H][W][3]
  uint8_t* dst;          // [frames][64][64][3]
  int H, W;
  int mode;              // 0 general, 1 integer scale (float multiply), 2 exact 2 x 2
  float inv_area;        // modes 1, 2
  // tables over the CROPPED destination window (64 entries each): entry ranges and the source span they touch
  const int* x_begin;    // [65] prefix into x_si / x_alpha
  const int* x_si;
  const float* x_alpha;
  const int* y_begin;    // [65]
  const int* y_si;
  const float* y_beta;
  int sx_min, sx_count;  // source columns [sx_min, sx_min + sx_count) cover every x entry
};

constexpr int kPreMaxSpan = 8192;   // bytes of one staged source row segment (sx_count * 3 <= 8192: scale factors up to ~42)
constexpr int kPreRowsPerStage = 8; // source rows staged per barrier pair: their loads are all in flight together

__global__ void __launch_bounds__(192) preprocess_area_kernel(const PreprocessParams p) {
  extern __shared__ __align__(16) uint8_t pre_rows[];   // kPreRowsPerStage slots of `pitch` bytes
  const int dy = blockIdx.x, f = blockIdx.y;
  const int dx = threadIdx.x / 3, c = threadIdx.x - 3 * dx;
  const int xb = p.x_begin[dx], xe = p.x_begin[dx + 1];
  const int yb = p.y_begin[dy], ye = p.y_begin[dy + 1];
  const uint8_t* frame = p.src + (size_t)f * p.H * p.W * 3;
  const int span = p.sx_count * 3;
  const int pitch = (span + 32 + 15) & ~15;
  float sum = 0.0f;
  int isum = 0;
  for (int j0 = yb; j0 < ye; j0 += kPreRowsPerStage) {
    const int nr = min(kPreRowsPerStage, ye - j0);
    __syncthreads();
    // coalesced copy of the row segments: 16-byte vectors between the first and last aligned address, single bytes at the two
    // ends; logical byte i of a segment is staged at slot[off + i] with off = 16 - head, which keeps the vector stores aligned
    for (int r = 0; r < nr; ++r) {
      const uint8_t* srow = frame + ((size_t)p.y_si[j0 + r] * p.W + p.sx_min) * 3;
      uint8_t* slot = pre_rows + r * pitch;
      const int head = (int)((16 - (reinterpret_cast<uintptr_t>(srow) & 15)) & 15);
      const int off = 16 - head;
      const int nvec = span > head ? (span - head) >> 4 : 0;
      const int tail0 = head + 16 * nvec;
      for (int i = threadIdx.x; i < head && i < span; i += 192) slot[off + i] = srow[i];
      const uint4* v = reinterpret_cast<const uint4*>(srow + head);
      for (int i = threadIdx.x; i < nvec; i += 192) *reinterpret_cast<uint4*>(slot + 16 + 16 * i) = __ldg(v + i);
      for (int i = tail0 + threadIdx.x; i < span; i += 192) slot[off + i] = srow[i];
    }
    __syncthreads();
    for (int r = 0; r < nr; ++r) {
      const int j = j0 + r;
      const uint8_t* srow = frame + ((size_t)p.y_si[j] * p.W + p.sx_min) * 3;
      const uint8_t* rowv = pre_rows + r * pitch + 16 - (int)((16 - (reinterpret_cast<uintptr_t>(srow) & 15)) & 15);
      if (p.mode == 0) {
        float buf = 0.0f;
        for (int k = xb; k < xe; ++k) buf = __fadd_rn(buf, __fmul_rn((float)rowv[(p.x_si[k] - p.sx_min) * 3 + c], p.x_alpha[k]));
        const float t = __fmul_rn(p.y_beta[j], buf);
        sum = (j == yb) ? t : __fadd_rn(sum, t);
      } else {
        for (int k = xb; k < xe; ++k) isum += rowv[(p.x_si[k] - p.sx_min) * 3 + c];
      }
    }
  }
  float v;
  if (p.mode == 0) v = sum;
  else if (p.mode == 1) v = __fmul_rn((float)isum, p.inv_area);
  else v = (float)((isum + 2) >> 2);
  const int r = __float2int_rn(v);    // round half to even, like cvRound
  p.dst[(((size_t)f * 64 + dy) * 64 + dx) * 3 + c] = (uint8_t)min(max(r, 0), 255);
}


// Up-scaling (a frame side below 64 px): cv::resize leaves the area code and runs its 8-bit bilinear kernels with the
// INTER_AREA coefficient rule (imgproc/src/resize.cpp: HResizeLinear<uchar, int, short, 2048> and VResizeLinear<uchar, int,
// short, FixedPtCast<int, uchar, 22>>): horizontal pass in int32 with two 11-bit coefficients, vertical pass
// (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2. The tables of the 64 x 64 crop window travel in the kernel
// parameters; such frames are tiny, so every thread reads its four source bytes straight from global memory.
struct PreprocessUpParams {
  const uint8_t* src;    // [frames][H][W][3]
  uint8_t* dst;          // [frames][64][64][3]
  int H, W;
  int xo[64], yo[64];            // source column / row of the first sample
  short xa[64][2], yb[64][2];    // 11-bit fixed-point coefficient pairs
};

__global__ void __launch_bounds__(192) preprocess_linear_kernel(const __grid_constant__ PreprocessUpParams p) {
  const int dy = blockIdx.x, f = blockIdx.y;
  const int dx = threadIdx.x / 3, c = threadIdx.x - 3 * dx;
  const uint8_t* frame = p.src + (size_t)f * p.H * p.W * 3;
  const int s0 = p.xo[dx], s1 = min(s0 + 1, p.W - 1);
  const int r0 = p.yo[dy], r1 = min(r0 + 1, p.H - 1);
  const int a0 = p.xa[dx][0], a1 = p.xa[dx][1], b0 = p.yb[dy][0], b1 = p.yb[dy][1];
  const uint8_t* row0 = frame + (size_t)r0 * p.W * 3;
  const uint8_t* row1 = frame + (size_t)r1 * p.W * 3;
  const int h0 = row0[s0 * 3 + c] * a0 + row0[s1 * 3 + c] * a1;
  const int h1 = row1[s0 * 3 + c] * a0 + row1[s1 * 3 + c] * a1;
  const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
  p.dst[(((size_t)f * 64 + dy) * 64 + dx) * 3 + c] = (uint8_t)v;
}

}  // namespace vfp
