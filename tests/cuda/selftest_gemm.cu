// Standalone device self-test for the tcgen05 GEMM skeleton (no torch, no Python):
//   * plain K-major GEMMs at several tile shapes, checked against a host fp32/double reference
//   * 3x3 / stride-2 / pad-1 implicit-GEMM convolutions whose A operand is a strided TMA box
// Build: see video_fingerprint_b200/build.py (target "selftest"). Exit code 0 = all checks passed.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../video_fingerprint_b200/csrc/gemm_launch.cuh"

using namespace vfp;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

static uint32_t rng_state = 12345u;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}
static __nv_bfloat16 f2b(float f) { return __float2bfloat16(f); }
static float b2f(__nv_bfloat16 b) { return __bfloat162float(b); }

template <int BN, int BK, int ST>
static int run_plain(long long M, int N, int K, int act, bool with_res) {
  std::vector<__nv_bfloat16> hA(M * K), hB((size_t)N * K);
  std::vector<float> hbias(N), hres(M * N);
  for (auto& v : hA) v = f2b(frand());
  for (auto& v : hB) v = f2b(frand());
  for (auto& v : hbias) v = frand();
  for (auto& v : hres) v = frand();
  __nv_bfloat16 *dA, *dB, *dObf;
  float *dbias, *dres, *dO;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dbias, N * 4));
  CK(cudaMalloc(&dres, hres.size() * 4));
  CK(cudaMalloc(&dO, M * N * 4));
  CK(cudaMalloc(&dObf, M * N * 2));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, hbias.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dres, hres.data(), hres.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0xFF, M * N * 4));
  CUtensorMap ta, tb;
  if (make_tmap_rows_bf16(&ta, dA, M, K, K, 128, BK) || make_tmap_rows_bf16(&tb, dB, N, K, K, BN, BK)) {
    printf("tensor map encode failed\n");
    return 1;
  }
  GemmShape s = plain_shape(M, N, K, BN, BK, 4);
  EpiBiasAct::Params ep{};
  ep.bias = dbias;
  ep.residual = with_res ? dres : nullptr;
  ep.ld_res = N;
  ep.out_f32 = dO;
  ep.out_bf16 = dObf;
  ep.ld_out = N;
  ep.M = (int)M;
  ep.N = N;
  ep.act = act;
  CK((launch_gemm<BN, BK, ST, EpiBiasAct>(ta, tb, s, ep, 0)));
  CK(cudaDeviceSynchronize());
  std::vector<float> hO(M * N);
  std::vector<__nv_bfloat16> hObf(M * N);
  CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hObf.data(), dObf, hObf.size() * 2, cudaMemcpyDeviceToHost));
  double max_err = 0, max_err_bf = 0;
  for (long long m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)b2f(hA[m * K + k]) * (double)b2f(hB[(size_t)n * K + k]);
      acc += hbias[n];
      if (act == 1) acc = acc > 0 ? acc : 0;
      if (act == 2) acc = 0.5 * acc * (1.0 + erf(acc * 0.7071067811865476));
      if (with_res) acc += hres[m * N + n];
      double e = fabs(acc - hO[m * N + n]);
      if (!(e <= max_err)) max_err = e;
      double eb = fabs(acc - b2f(hObf[m * N + n])) / (1.0 + fabs(acc));
      if (!(eb <= max_err_bf)) max_err_bf = eb;
    }
  const bool ok = max_err < 2e-3 && max_err_bf < 1e-2;
  printf("[plain BN=%d BK=%d] M=%lld N=%d K=%d act=%d res=%d  max_abs_err=%.3e  bf16_rel=%.3e  %s\n", BN, BK, M, N,
         K, act, (int)with_res, max_err, max_err_bf, ok ? "OK" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dbias); cudaFree(dres); cudaFree(dO); cudaFree(dObf);
  return ok ? 0 : 1;
}

// conv 3x3 s2 p1, NHWC, weights [COUT][9*CIN] with k = (kh*3+kw)*CIN + c
template <int CIN, int COUT, int HIN, int BN, int BK, int ST>
static int run_conv(int frames) {
  constexpr int HOUT = HIN / 2;
  constexpr int PIX = HOUT * HOUT;
  const int K = 9 * CIN;
  std::vector<__nv_bfloat16> hX((size_t)frames * HIN * HIN * CIN), hW((size_t)COUT * K);
  std::vector<float> hbias(COUT);
  for (auto& v : hX) v = f2b(frand());
  for (auto& v : hW) v = f2b(frand() * 0.2f);
  for (auto& v : hbias) v = frand();
  __nv_bfloat16 *dX, *dW, *dO;
  float* dbias;
  CK(cudaMalloc(&dX, hX.size() * 2));
  CK(cudaMalloc(&dW, hW.size() * 2));
  CK(cudaMalloc(&dbias, COUT * 4));
  CK(cudaMalloc(&dO, (size_t)frames * PIX * COUT * 2));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, hbias.data(), COUT * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0xFF, (size_t)frames * PIX * COUT * 2));

  GemmShape s{};
  s.a_conv = 1;
  s.group_m = 8;
  s.n_tiles = COUT / BN;
  s.k_blocks = K / BK;
  conv_taps_strided(&s, CIN / BK);
  int out_w = HOUT, out_h, n_box;
  if (PIX >= 128) {
    s.tiles_per_frame = PIX / 128;
    s.frames_per_tile = 1;
    out_h = 128 / HOUT;
    n_box = 1;
    s.m_tiles = frames * s.tiles_per_frame;
  } else {
    s.tiles_per_frame = 1;
    s.frames_per_tile = 128 / PIX;
    out_h = HOUT;
    n_box = s.frames_per_tile;
    s.m_tiles = (frames + s.frames_per_tile - 1) / s.frames_per_tile;
  }
  s.tile_out_rows = out_h;
  CUtensorMap ta, tb;
  if (make_tmap_conv_s2_bf16(&ta, dX, frames, HIN, HIN, CIN, BK, out_w, out_h, n_box) ||
      make_tmap_rows_bf16(&tb, dW, COUT, K, K, BN, BK)) {
    printf("tensor map encode failed (conv)\n");
    return 1;
  }
  EpiBiasAct::Params ep{};
  ep.bias = dbias;
  ep.out_bf16 = dO;
  ep.ld_out = COUT;
  ep.M = frames * PIX;
  ep.N = COUT;
  ep.act = 1;
  CK((launch_gemm<BN, BK, ST, EpiBiasAct>(ta, tb, s, ep, 0)));
  CK(cudaDeviceSynchronize());
  std::vector<__nv_bfloat16> hO((size_t)frames * PIX * COUT);
  CK(cudaMemcpy(hO.data(), dO, hO.size() * 2, cudaMemcpyDeviceToHost));
  double max_err = 0;
  for (int f = 0; f < frames; ++f)
    for (int oh = 0; oh < HOUT; ++oh)
      for (int ow = 0; ow < HOUT; ++ow)
        for (int co = 0; co < COUT; ++co) {
          double acc = hbias[co];
          for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
              const int ih = 2 * oh + kh - 1, iw = 2 * ow + kw - 1;
              if (ih < 0 || iw < 0 || ih >= HIN || iw >= HIN) continue;
              const __nv_bfloat16* px = &hX[(((size_t)f * HIN + ih) * HIN + iw) * CIN];
              const __nv_bfloat16* wr = &hW[(size_t)co * K + (kh * 3 + kw) * CIN];
              for (int c = 0; c < CIN; ++c) acc += (double)b2f(px[c]) * (double)b2f(wr[c]);
            }
          if (acc < 0) acc = 0;
          const double got = b2f(hO[(((size_t)f * HOUT + oh) * HOUT + ow) * COUT + co]);
          const double e = fabs(acc - got) / (1.0 + fabs(acc));
          if (!(e <= max_err)) max_err = e;
        }
  const bool ok = max_err < 1e-2;
  printf("[conv %d->%d %dx%d BN=%d BK=%d] frames=%d  rel_err=%.3e  %s\n", CIN, COUT, HIN, HIN, BN, BK, frames,
         max_err, ok ? "OK" : "FAIL");
  cudaFree(dX); cudaFree(dW); cudaFree(dbias); cudaFree(dO);
  return ok ? 0 : 1;
}

template <int BN, int BK, int ST, int KBLOCKS>
static int run_bres(long long M, int N) {
  const int K = KBLOCKS * BK;
  std::vector<__nv_bfloat16> hA(M * K), hB((size_t)N * K);
  std::vector<float> hbias(N);
  for (auto& v : hA) v = f2b(frand());
  for (auto& v : hB) v = f2b(frand());
  for (auto& v : hbias) v = frand();
  __nv_bfloat16 *dA, *dB;
  float *dbias, *dO;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dbias, N * 4));
  CK(cudaMalloc(&dO, M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, hbias.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0xFF, M * N * 4));
  CUtensorMap ta, tb;
  if (make_tmap_rows_bf16(&ta, dA, M, K, K, 128, BK) || make_tmap_rows_bf16(&tb, dB, N, K, K, BN, BK)) return 1;
  GemmShape s = plain_shape(M, N, K, BN, BK, 4);
  EpiBiasAct::Params ep{};
  ep.bias = dbias; ep.out_f32 = dO; ep.ld_out = N; ep.M = (int)M; ep.N = N;
  CK((launch_gemm_bres<BN, BK, ST, KBLOCKS, EpiBiasAct>(ta, tb, s, ep, 0)));
  CK(cudaDeviceSynchronize());
  std::vector<float> hO(M * N);
  CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0;
  for (long long m = 0; m < M; m += 7)
    for (int n = 0; n < N; ++n) {
      double acc = hbias[n];
      for (int k = 0; k < K; ++k) acc += (double)b2f(hA[m * K + k]) * (double)b2f(hB[(size_t)n * K + k]);
      const double e = fabs(acc - hO[m * N + n]);
      if (!(e <= max_err)) max_err = e;
    }
  const bool ok = max_err < 2e-3;
  printf("[b-resident BN=%d BK=%d KB=%d] M=%lld N=%d  max_abs_err=%.3e  %s\n", BN, BK, KBLOCKS, M, N, max_err, ok ? "OK" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dbias); cudaFree(dO);
  return ok ? 0 : 1;
}

int main() {
  int fails = 0;
  fails += run_bres<256, 64, 4, 4>(300, 256);
  fails += run_bres<256, 64, 4, 4>(40000, 768);     // several column tiles per CTA: B is reloaded mid-range
  fails += run_bres<256, 64, 4, 4>(128 * 200, 1024);
  fails += run_bres<128, 64, 4, 9>(5000, 128);
  fails += run_bres<64, 64, 6, 6>(3000, 64);
  fails += run_plain<256, 64, 4>(128, 256, 64, 0, false);
  fails += run_plain<256, 64, 4>(128, 256, 256, 0, false);
  fails += run_plain<256, 64, 4>(300, 256, 256, 1, true);
  fails += run_plain<256, 64, 4>(5000, 768, 256, 0, false);
  fails += run_plain<256, 64, 4>(1000, 256, 1024, 2, true);
  fails += run_plain<128, 64, 6>(777, 128, 576, 1, false);
  fails += run_plain<64, 32, 8>(515, 64, 288, 1, false);
  fails += run_plain<32, 64, 6>(400, 32, 128, 0, false);
  fails += run_conv<32, 64, 32, 64, 32, 8>(5);
  fails += run_conv<64, 128, 16, 128, 64, 6>(5);
  fails += run_conv<128, 256, 8, 256, 64, 4>(19);
  unsigned int dev_err = 0;
  cudaMemcpyFromSymbol(&dev_err, g_vfp_device_error, sizeof(dev_err));
  printf("device error word: 0x%x\n", dev_err);
  printf(fails ? "SELFTEST FAILED (%d)\n" : "SELFTEST PASSED (%d failures)\n", fails);
  return fails ? 1 : 0;
}
