// Token-stream kernels of the temporal encoder that are NOT dense contractions: clip bookkeeping,
// LayerNorm, the grouped multi-scale temporal convolution, self-attention over one clip, the three-way
// temporal pooling and the final projection + L2 normalisation. Tokens of all clips are PACKED
// ([sum T][256], clip v owns rows cu[v] .. cu[v+1]) - there is no padding, so the reference's B=1
// semantics for variable-length clips (fingerprint.py:244-249) hold by construction: the temporal conv
// zero-pads at clip ends, attention/pooling only ever see a clip's own rows.
#pragma once
#include "sm100_primitives.cuh"
#include "epilogues.cuh"

namespace vfp {

constexpr int kDim = 256;      // temporal_dim
constexpr int kHeads = 8;
constexpr int kHeadDim = 32;

// ---------------------------------------------------------------------------------------------
// token -> (clip, position) maps from the prefix sums
// ---------------------------------------------------------------------------------------------
__global__ void token_map_kernel(const int* __restrict__ cu, int n_clips, int n_tokens, int* __restrict__ tok_pos,
                                 int* __restrict__ tok_len) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tokens) return;
  int lo = 0, hi = n_clips;  // find v with cu[v] <= t < cu[v+1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cu[mid] <= t) lo = mid; else hi = mid;
  }
  tok_pos[t] = t - cu[lo];
  tok_len[t] = cu[lo + 1] - cu[lo];
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over 256 channels, fp32 in -> bf16 out. One warp per token, 8 channels per lane.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layernorm_bf16_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                      __nv_bfloat16* __restrict__ y, int n_tokens) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_tokens) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)warp * kDim) + lane * 2;
  const float4 a = xr[0], b = xr[1];
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / kDim);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] -= mean; q += v[i] * v[i]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.0f / kDim) + 1e-5f);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + lane * 2);
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + lane * 2 + 1);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + lane * 2);
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + lane * 2 + 1);
  uint4 o;
  o.x = pack_bf16x2(v[0] * rstd * g0.x + b0.x, v[1] * rstd * g0.y + b0.y);
  o.y = pack_bf16x2(v[2] * rstd * g0.z + b0.z, v[3] * rstd * g0.w + b0.w);
  o.z = pack_bf16x2(v[4] * rstd * g1.x + b1.x, v[5] * rstd * g1.y + b1.y);
  o.w = pack_bf16x2(v[6] * rstd * g1.z + b1.z, v[7] * rstd * g1.w + b1.w);
  reinterpret_cast<uint4*>(y + (size_t)warp * kDim)[lane] = o;
}

// ---------------------------------------------------------------------------------------------
// TemporalConvBlock + residual:  y = x + relu(bn(groupedconv_k(x)))  for k in {3,5,7,11} concatenated.
// Output channel o (branch o/64, group o%64) reads input channels 4*(o%64) .. +3 (model.py:163-169).
// Weights are BN-folded and zero-padded to 11 centred taps: w[ci][tap][o], so all 256 channels run the
// same loop and weight reads are coalesced.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
temporal_conv_kernel(const float* __restrict__ x, const int* __restrict__ tok_pos, const int* __restrict__ tok_len,
                     const float* __restrict__ w /*[4][11][256]*/, const float* __restrict__ bias /*[256]*/,
                     float* __restrict__ y, int n_tokens) {
  constexpr int TOK = 8;
  const int o = threadIdx.x;
  const int t0 = blockIdx.x * TOK;
  float wr[4][11];
#pragma unroll
  for (int ci = 0; ci < 4; ++ci)
#pragma unroll
    for (int tap = 0; tap < 11; ++tap) wr[ci][tap] = __ldg(w + (ci * 11 + tap) * kDim + o);
  const float b = __ldg(bias + o);
  const int cin = 4 * (o & 63);
  for (int tt = 0; tt < TOK; ++tt) {
    const int t = t0 + tt;
    if (t >= n_tokens) break;
    const int pos = tok_pos[t], len = tok_len[t];
    float acc = b;
#pragma unroll
    for (int tap = 0; tap < 11; ++tap) {
      const int p = pos + tap - 5;
      if (p >= 0 && p < len) {
        const float4 xi = *reinterpret_cast<const float4*>(x + (size_t)(t + tap - 5) * kDim + cin);
        acc += wr[0][tap] * xi.x + wr[1][tap] * xi.y + wr[2][tap] * xi.z + wr[3][tap] * xi.w;
      }
    }
    y[(size_t)t * kDim + o] = x[(size_t)t * kDim + o] + fmaxf(acc, 0.0f);
  }
}

// ---------------------------------------------------------------------------------------------
// Multi-head self-attention over one clip. grid = (clip, head); K and V of the head live in shared
// memory as bf16; each thread owns one query row and runs an online softmax over the clip's keys.
// qkv: [tokens][768] bf16 = [Q | K | V], head h = columns 32h..32h+31 of each (model.py:130-132, 143).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attention_clip_kernel(const __nv_bfloat16* __restrict__ qkv, const int* __restrict__ cu, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  const int clip = blockIdx.x, head = blockIdx.y;
  const int t0 = cu[clip];
  const int T = cu[clip + 1] - t0;
  uint4* ks = reinterpret_cast<uint4*>(att_smem);   // [T][4] uint4 = 32 bf16 per key
  uint4* vs = ks + (size_t)T * 4;
  for (int i = threadIdx.x; i < T * 4; i += blockDim.x) {
    const int t = i >> 2, c = i & 3;
    const uint4* row = reinterpret_cast<const uint4*>(qkv + (size_t)(t0 + t) * (3 * kDim) + head * kHeadDim);
    ks[i] = row[kDim / 8 + c];       // +256 bf16
    vs[i] = row[2 * kDim / 8 + c];   // +512 bf16
  }
  __syncthreads();
  const float scale = 0.17677669529663687f;  // 1/sqrt(32)
  for (int qi = threadIdx.x; qi < T; qi += blockDim.x) {
    float q[kHeadDim];
    {
      const uint4* row = reinterpret_cast<const uint4*>(qkv + (size_t)(t0 + qi) * (3 * kDim) + head * kHeadDim);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 u = row[c];
        const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(p[e]);
          q[c * 8 + 2 * e] = f.x * scale;
          q[c * 8 + 2 * e + 1] = f.y * scale;
        }
      }
    }
    float m = -INFINITY, l = 0.f;
    float o[kHeadDim];
#pragma unroll
    for (int d = 0; d < kHeadDim; ++d) o[d] = 0.f;
    for (int j = 0; j < T; ++j) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 u = ks[j * 4 + c];
        const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(p[e]);
          s += q[c * 8 + 2 * e] * f.x + q[c * 8 + 2 * e + 1] * f.y;
        }
      }
      const float m_new = fmaxf(m, s);
      const float corr = __expf(m - m_new);
      const float pj = __expf(s - m_new);
      l = l * corr + pj;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 u = vs[j * 4 + c];
        const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(p[e]);
          o[c * 8 + 2 * e] = o[c * 8 + 2 * e] * corr + pj * f.x;
          o[c * 8 + 2 * e + 1] = o[c * 8 + 2 * e + 1] * corr + pj * f.y;
        }
      }
      m = m_new;
    }
    const float inv = 1.0f / l;
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(t0 + qi) * kDim + head * kHeadDim);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 u;
      u.x = pack_bf16x2(o[c * 8 + 0] * inv, o[c * 8 + 1] * inv);
      u.y = pack_bf16x2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv);
      u.z = pack_bf16x2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv);
      u.w = pack_bf16x2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv);
      dst[c] = u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// adaptive_pooling (model.py:256-270): per clip and channel, mean_T(x), max_T(x) and
// sum_T x * softmax_T(logit) where logit = relu(W_p x + b_p) was produced by the GEMM before.
// One CTA per clip, one thread per channel, single pass with an online softmax.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
temporal_pool_kernel(const float* __restrict__ x, const float* __restrict__ logits, const int* __restrict__ cu,
                     float* __restrict__ pooled /*[clips][768]*/) {
  const int clip = blockIdx.x, c = threadIdx.x;
  const int t0 = cu[clip], T = cu[clip + 1] - t0;
  float sum = 0.f, mx = -INFINITY, m = -INFINITY, l = 0.f, ws = 0.f;
  for (int t = 0; t < T; ++t) {
    const float xv = x[(size_t)(t0 + t) * kDim + c];
    const float lg = logits[(size_t)(t0 + t) * kDim + c];
    sum += xv;
    mx = fmaxf(mx, xv);
    const float m_new = fmaxf(m, lg);
    const float corr = __expf(m - m_new);
    const float p = __expf(lg - m_new);
    l = l * corr + p;
    ws = ws * corr + p * xv;
    m = m_new;
  }
  float* o = pooled + (size_t)clip * (3 * kDim);
  o[c] = sum / (float)T;
  o[kDim + c] = mx;
  o[2 * kDim + c] = ws / l;
}

// ---------------------------------------------------------------------------------------------
// final_projection (Linear 768->256, ReLU, Linear 256->D) + L2 normalise (model.py:219-224, 292-294),
// fp32 on CUDA cores: 0.5 MFLOP per clip, and the last rounding step before the embedding leaves.
// kClips clips per CTA so the transposed weights are streamed once per CTA, thread = output channel.
// ---------------------------------------------------------------------------------------------
template <int kClips>
__global__ void __launch_bounds__(256)
final_projection_kernel(const float* __restrict__ pooled, const float* __restrict__ w0t /*[768][256]*/,
                        const float* __restrict__ b0, const float* __restrict__ w3t /*[256][D]*/,
                        const float* __restrict__ b3, int D, int n_clips, float* __restrict__ emb /*[clips][D]*/) {
  __shared__ float sp[kClips][3 * kDim];
  __shared__ float sh[kClips][kDim];
  __shared__ float red[kClips][8];
  const int c0 = blockIdx.x * kClips;
  const int tid = threadIdx.x;
  for (int i = tid; i < kClips * 3 * kDim; i += 256) {
    const int v = i / (3 * kDim), k = i - v * 3 * kDim;
    sp[v][k] = (c0 + v < n_clips) ? pooled[(size_t)(c0 + v) * 3 * kDim + k] : 0.f;
  }
  __syncthreads();
  float acc[kClips];
#pragma unroll
  for (int v = 0; v < kClips; ++v) acc[v] = 0.f;
  for (int k = 0; k < 3 * kDim; ++k) {
    const float w = __ldg(w0t + (size_t)k * kDim + tid);
#pragma unroll
    for (int v = 0; v < kClips; ++v) acc[v] += w * sp[v][k];
  }
  const float bb = __ldg(b0 + tid);
#pragma unroll
  for (int v = 0; v < kClips; ++v) sh[v][tid] = fmaxf(acc[v] + bb, 0.f);
  __syncthreads();
  // second layer: D may be smaller/larger than 256 -> loop output channels in strides of 256
  for (int oc = tid; oc < ((D + 255) / 256) * 256; oc += 256) {
    float e[kClips];
#pragma unroll
    for (int v = 0; v < kClips; ++v) e[v] = 0.f;
    if (oc < D) {
      for (int k = 0; k < kDim; ++k) {
        const float w = __ldg(w3t + (size_t)k * D + oc);
#pragma unroll
        for (int v = 0; v < kClips; ++v) e[v] += w * sh[v][k];
      }
      const float b = __ldg(b3 + oc);
#pragma unroll
      for (int v = 0; v < kClips; ++v) e[v] += b;
    }
    // stash un-normalised values in global, accumulate squared norms below
#pragma unroll
    for (int v = 0; v < kClips; ++v) {
      if (oc < D && c0 + v < n_clips) emb[(size_t)(c0 + v) * D + oc] = e[v];
      float q = (oc < D) ? e[v] * e[v] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      if ((tid & 31) == 0) {
        if (oc < 256) red[v][tid >> 5] = q; else red[v][tid >> 5] += q;
      }
    }
  }
  __syncthreads();
  for (int v = 0; v < kClips; ++v) {
    if (c0 + v >= n_clips) break;
    float n2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) n2 += red[v][i];
    const float inv = 1.0f / fmaxf(sqrtf(n2), 1e-12f);
    for (int oc = tid; oc < D; oc += 256) emb[(size_t)(c0 + v) * D + oc] *= inv;
  }
}

// fp32 -> bf16 row copy (pooling GEMM input, join operands)
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float4 a = reinterpret_cast<const float4*>(in)[2 * i];
  const float4 b = reinterpret_cast<const float4*>(in)[2 * i + 1];
  uint4 o;
  o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
  o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
  reinterpret_cast<uint4*>(out)[i] = o;
}

}  // namespace vfp
