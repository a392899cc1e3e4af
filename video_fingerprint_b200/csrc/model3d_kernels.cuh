// VideoFingerprint3D forward (/root/reference/model.py:406-512), the reference's second model behind create_model("3d").
// Four Conv3d + BatchNorm3d + ReLU blocks, spatial average pool, a temporal Conv1d, attention + average pooling over time,
// a two-layer projector and L2 normalisation. Every Conv3d runs as   im2col (this file)  ->  tcgen05 GEMM (gemm_sm100.cuh,
// bias + ReLU + bf16 TMA-store epilogue)   on bf16 operands with fp32 accumulation; the GEMM output [positions][C_out] IS the
// channels-last activation tensor the next layer's im2col reads. The small tail (pooling, Conv1d, softmax, projector,
// normalise: < 0.2 MFLOP per clip) is one fp32 CTA per clip.
// First version of this model: correct and on the tensor cores, but the explicit im2col costs ~6x the input in HBM traffic;
// an implicit-GEMM loader like the attention model's stem is the obvious next step.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv1_kernel.cuh"   // FrameDtype

namespace vfp {

// Layer 1: Conv3d(3 -> 16, kernel (fs, 5, 5), stride (fs, 2, 2), padding (0, 2, 2)) over planar frames (B*T, 3, 64, 64) of any
// accepted dtype; T is zero-padded to a multiple of fs like model.py:468-471. Row = (b, g, oh, ow) with 32 x 32 output
// positions per group of fs frames; column k = ((kt*5 + kh)*5 + kw)*3 + c, zero-filled up to `kp`.
__global__ void im2col3d_frames_kernel(const void* __restrict__ frames, int frame_dtype, int B, int T, int fs, int groups,
                                       int kp, __nv_bfloat16* __restrict__ out) {
  const long long rows = (long long)B * groups * 1024;
  const int k8n = kp / 8;
  const long long total = rows * k8n;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
    const long long row = it / k8n;
    const int k0 = (int)(it - row * k8n) * 8;
    const int ow = (int)(row & 31), oh = (int)((row >> 5) & 31);
    const long long bg = row >> 10;
    const int g = (int)(bg % groups);
    const long long b = bg / groups;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = k0 + e;
      float x = 0.0f;
      if (k < fs * 75) {
        const int c = k % 3, kw = (k / 3) % 5, kh = (k / 15) % 5, kt = k / 75;
        const int t = g * fs + kt, ih = 2 * oh + kh - 2, iw = 2 * ow + kw - 2;
        if (t < T && ih >= 0 && ih < 64 && iw >= 0 && iw < 64) {
          const size_t idx = (((size_t)(b * T + t) * 3 + c) * 64 + ih) * 64 + iw;
          if (frame_dtype == kFrameU8) x = (float)static_cast<const uint8_t*>(frames)[idx] * (1.0f / 255.0f);
          else if (frame_dtype == kFrameBF16) x = __bfloat162float(static_cast<const __nv_bfloat16*>(frames)[idx]);
          else x = static_cast<const float*>(frames)[idx];
        }
      }
      v[e] = __float2bfloat16(x);
    }
    *reinterpret_cast<uint4*>(out + row * kp + k0) = *reinterpret_cast<const uint4*>(v);
  }
}

// Layers 2-4: Conv3d(C -> *, kernel 3x3x3, stride (st, 2, 2), padding 1) over a channels-last activation [B][Ti][Hi][Wi][C]
// (C a multiple of 8). Row = (b, to, oh, ow), column k = ((kt*3 + kh)*3 + kw)*C + c; one thread moves 8 channels (16 bytes).
__global__ void im2col3d_ndhwc_kernel(const __nv_bfloat16* __restrict__ in, int B, int Ti, int Hi, int Wi, int C, int st, int To,
                                      int Ho, int Wo, int kp, __nv_bfloat16* __restrict__ out) {
  const int k8n = kp / 8;
  const long long rows = (long long)B * To * Ho * Wo;
  const long long total = rows * k8n;
  const int kreal = 27 * C;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
    const long long row = it / k8n;
    const int k0 = (int)(it - row * k8n) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (k0 < kreal) {
      const int tap = k0 / C, c0 = k0 - tap * C;
      const int kw = tap % 3, kh = (tap / 3) % 3, kt = tap / 9;
      long long r = row;
      const int ow = (int)(r % Wo); r /= Wo;
      const int oh = (int)(r % Ho); r /= Ho;
      const int to = (int)(r % To);
      const long long b = r / To;
      const int t = to * st + kt - 1, ih = 2 * oh + kh - 1, iw = 2 * ow + kw - 1;
      if (t >= 0 && t < Ti && ih >= 0 && ih < Hi && iw >= 0 && iw < Wi)
        v = *reinterpret_cast<const uint4*>(in + ((((size_t)b * Ti + t) * Hi + ih) * Wi + iw) * C + c0);
    }
    *reinterpret_cast<uint4*>(out + row * kp + k0) = v;
  }
}

struct Head3dParams {
  const __nv_bfloat16* act;   // layer-4 output [B][T3][16][128]
  int T3;
  const float *tc_w, *tc_b;   // temporal_conv (128, 128, 3), (128)
  const float *ta_w, *ta_b;   // temporal_attention (1, 128, 1), (1)
  const float *p0_w, *p0_b;   // projector.0 (128, 128)
  const float *p3_w, *p3_b;   // projector.3 (D, 128)
  int D;
  float* out;                 // [B][D]
};
constexpr int kHead3dMaxT = 32;   // temporal positions after the two temporal strides (clips of up to 32 * 2 * fs frames)

// model.py:476-509 for one clip per CTA (128 threads = one per channel)
__global__ void __launch_bounds__(128) head3d_kernel(const Head3dParams p) {
  __shared__ float feat[kHead3dMaxT + 2][128];   // spatial means, rows 0 and T3+1 are the Conv1d zero padding
  __shared__ float tc[kHead3dMaxT][128];
  __shared__ float prob[kHead3dMaxT];
  __shared__ float vec[128], hid[128];
  __shared__ float red[4];
  const int b = blockIdx.x, c = threadIdx.x, T3 = p.T3;
  const __nv_bfloat16* a = p.act + (size_t)b * T3 * 16 * 128;
  feat[0][c] = 0.0f;
  feat[T3 + 1][c] = 0.0f;
  for (int t = 0; t < T3; ++t) {
    float s = 0.0f;
    for (int px = 0; px < 16; ++px) s += __bfloat162float(a[((size_t)t * 16 + px) * 128 + c]);
    feat[t + 1][c] = s * (1.0f / 16.0f);   // AdaptiveAvgPool3d((None, 1, 1))
  }
  __syncthreads();
  for (int t = 0; t < T3; ++t) {           // temporal_conv: Conv1d(128, 128, 3, padding 1)
    float s = p.tc_b[c];
    const float* w = p.tc_w + (size_t)c * 128 * 3;
    for (int ci = 0; ci < 128; ++ci)
      s += w[ci * 3] * feat[t][ci] + w[ci * 3 + 1] * feat[t + 1][ci] + w[ci * 3 + 2] * feat[t + 2][ci];
    tc[t][c] = s;
  }
  __syncthreads();
  for (int t = c; t < T3; t += 128) {      // temporal_attention logits
    float s = p.ta_b[0];
    for (int ci = 0; ci < 128; ++ci) s += p.ta_w[ci] * tc[t][ci];
    prob[t] = s;
  }
  __syncthreads();
  if (c == 0) {                             // softmax over time (T3 <= 64)
    float mx = -INFINITY, den = 0.0f;
    for (int t = 0; t < T3; ++t) mx = fmaxf(mx, prob[t]);
    for (int t = 0; t < T3; ++t) { prob[t] = expf(prob[t] - mx); den += prob[t]; }
    for (int t = 0; t < T3; ++t) prob[t] /= den;
  }
  __syncthreads();
  {
    float wsum = 0.0f, avg = 0.0f;
    for (int t = 0; t < T3; ++t) { wsum += tc[t][c] * prob[t]; avg += tc[t][c]; }
    vec[c] = wsum + avg / (float)T3;        // weighted + average pooling
  }
  __syncthreads();
  {
    float s = p.p0_b[c];
    const float* w = p.p0_w + (size_t)c * 128;
    for (int ci = 0; ci < 128; ++ci) s += w[ci] * vec[ci];
    hid[c] = fmaxf(s, 0.0f);
  }
  __syncthreads();
  float e[4];   // D <= 512
  float sq = 0.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int d = c + 128 * j;
    e[j] = 0.0f;
    if (d < p.D) {
      float s = p.p3_b[d];
      const float* w = p.p3_w + (size_t)d * 128;
      for (int ci = 0; ci < 128; ++ci) s += w[ci] * hid[ci];
      e[j] = s;
      sq += s * s;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if ((c & 31) == 0) red[c >> 5] = sq;
  __syncthreads();
  const float nrm = fmaxf(sqrtf(red[0] + red[1] + red[2] + red[3]), 1e-12f);   // F.normalize eps
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int d = c + 128 * j;
    if (d < p.D) p.out[(size_t)b * p.D + d] = e[j] / nrm;
  }
}

}  // namespace vfp
