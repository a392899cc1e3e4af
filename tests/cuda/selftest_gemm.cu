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

template <int BN, int BK, int ST, int MTILES>
static int run_plain_mt(long long M, int N, int K) {
  std::vector<__nv_bfloat16> hA(M * K), hB((size_t)N * K);
  std::vector<float> hbias(N);
  for (auto& v : hA) v = f2b(frand());
  for (auto& v : hB) v = f2b(frand());
  for (auto& v : hbias) v = frand();
  __nv_bfloat16 *dA, *dB, *dO;
  float* dbias;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dbias, N * 4));
  CK(cudaMalloc(&dO, M * N * 2));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, hbias.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0xFF, M * N * 2));
  CUtensorMap ta, tb;
  if (make_tmap_rows_bf16(&ta, dA, M, K, K, 128, BK) || make_tmap_rows_bf16(&tb, dB, N, K, K, BN, BK)) return 1;
  GemmShape s = plain_shape(M, N, K, BN, BK, 4);
  EpiBiasActTma<true>::Params ep{};
  if (make_tmap_out(&ep.tmap_out, dO, M, N, true)) return 1;
  ep.bias = dbias; ep.N = N; ep.act = 1;
  CK((launch_gemm<BN, BK, ST, EpiBiasActTma<true>, MTILES>(ta, tb, s, ep, 0)));
  CK(cudaDeviceSynchronize());
  std::vector<__nv_bfloat16> hO(M * N);
  CK(cudaMemcpy(hO.data(), dO, hO.size() * 2, cudaMemcpyDeviceToHost));
  double max_err = 0;
  for (long long m = 0; m < M; m += 3)
    for (int n = 0; n < N; ++n) {
      double acc = hbias[n];
      for (int k = 0; k < K; ++k) acc += (double)b2f(hA[m * K + k]) * (double)b2f(hB[(size_t)n * K + k]);
      if (acc < 0) acc = 0;
      const double e = fabs(acc - b2f(hO[m * N + n])) / (1.0 + fabs(acc));
      if (!(e <= max_err)) max_err = e;
    }
  const bool ok = max_err < 1e-2;
  printf("[multi-acc MT=%d BN=%d, TMA-store bf16] M=%lld N=%d K=%d  rel_err=%.3e  %s\n", MTILES, BN, M, N, K, max_err, ok ? "OK" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dbias); cudaFree(dO);
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

// ---- micro-benchmarks (run with any argument): operand-fill behaviour of the conv2 shape ----
struct EpiNull : EpiDefaults {
  struct Params { int dummy; };
  __device__ __forceinline__ void begin(const Params&, int, int, int) {}
  __device__ __forceinline__ void end(const Params&, int, int, int) {}
  __device__ __forceinline__ void chunk(const Params&, int, int, int, uint32_t (&)[32], int) {}
};

template <class F>
static float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

static void perf() {
  const int frames = 16384;
  __nv_bfloat16 *c1, *w;
  CK(cudaMalloc(&c1, (size_t)frames * 65536));
  CK(cudaMalloc(&w, 256 * 1152 * 2));
  CK(cudaMemset(c1, 0, (size_t)frames * 65536));
  CK(cudaMemset(w, 0, 256 * 1152 * 2));
  EpiNull::Params ep{0};
  CUtensorMap ta, tb;
  // (1) conv2 as shipped: s2d 4-D box (64 ch x 16 W x 8 H), 6 K blocks, weights resident
  {
    make_tmap_nhwc_bf16(&ta, c1, frames, 16, 16, 128, 64, 16, 8, 1, 1);
    make_tmap_rows_bf16(&tb, w, 64, 384, 384, 64, 64);
    GemmShape s{};
    s.m_tiles = 2 * frames; s.n_tiles = 1; s.k_blocks = 6; s.group_m = 16; s.a_conv = 1; s.tiles_per_frame = 2; s.frames_per_tile = 1;
    s.tile_out_rows = 8; s.h_mul = 1;
    const int c[6] = {1, 0, 1, 1, 0, 1}, dw[6] = {0, 0, 0, -1, -1, -1}, dh[6] = {-1, 0, 0, -1, 0, 0};
    for (int k = 0; k < 6; ++k) { s.tap_c_blk[k] = c[k]; s.tap_w[k] = dw[k]; s.tap_h[k] = dh[k]; }
    printf("conv2 s2d 4-D box, B resident      : %.3f ms\n", time_ms([&] { launch_gemm_bres<64, 64, 6, 6, EpiNull>(ta, tb, s, ep, 0); }));
    for (int k = 0; k < 6; ++k) { s.tap_w[k] = 0; s.tap_h[k] = 0; }
    printf("  same, all taps at offset 0 (L2 re-use): %.3f ms\n", time_ms([&] { launch_gemm_bres<64, 64, 6, 6, EpiNull>(ta, tb, s, ep, 0); }));
  }
  // (2) same bytes as a plain 2-D matrix [frames*256][384] (rows of 768 B, box 128 x 128 B)
  {
    const long long M2 = (long long)frames * 65536 / 768;  // same bytes as c1
    make_tmap_rows_bf16(&ta, c1, (uint64_t)M2, 384, 384, 128, 64);
    make_tmap_rows_bf16(&tb, w, 64, 384, 384, 64, 64);
    GemmShape s = plain_shape(M2, 64, 384, 64, 64, 16);
    printf("(plain 2-D has %.2fx the row tiles of conv2)\n", (double)s.m_tiles / (2.0 * frames));
    printf("plain 2-D [M][384] N=64, B resident : %.3f ms\n", time_ms([&] { launch_gemm_bres<64, 64, 6, 6, EpiNull>(ta, tb, s, ep, 0); }));
    printf("plain 2-D [M][384] N=64, generic    : %.3f ms\n", time_ms([&] { launch_gemm<64, 64, 8, EpiNull>(ta, tb, s, ep, 0); }));
    printf("plain 2-D [M][384] N=64, generic MT=2: %.3f ms\n", time_ms([&] { launch_gemm<64, 64, 4, EpiNull, 2>(ta, tb, s, ep, 0); }));
    printf("plain 2-D [M][384] N=64, generic MT=3: %.3f ms\n", time_ms([&] { launch_gemm<64, 64, 3, EpiNull, 3>(ta, tb, s, ep, 0); }));
    printf("plain 2-D [M][384] N=64, generic MT=4: %.3f ms\n", time_ms([&] { launch_gemm<64, 64, 2, EpiNull, 4>(ta, tb, s, ep, 0); }));
  }
  // (2b) L2-resident operand (50 MB): isolates the MMA/TMEM side from HBM
  {
    const long long M3 = 65536;
    make_tmap_rows_bf16(&ta, c1, (uint64_t)M3, 384, 384, 128, 64);
    make_tmap_rows_bf16(&tb, w, 64, 384, 384, 64, 64);
    GemmShape s = plain_shape(M3, 64, 384, 64, 64, 16);
    const double fl = 2.0 * M3 * 64 * 384 / 1e9;
    float t;
    t = time_ms([&] { launch_gemm<64, 64, 8, EpiNull>(ta, tb, s, ep, 0); }, 20); printf("L2-resident N=64 K=384 MT=1: %.4f ms %.0f TFLOP/s\n", t, fl / t);
    t = time_ms([&] { launch_gemm<64, 64, 4, EpiNull, 2>(ta, tb, s, ep, 0); }, 20); printf("L2-resident N=64 K=384 MT=2: %.4f ms %.0f TFLOP/s\n", t, fl / t);
    t = time_ms([&] { launch_gemm<64, 64, 3, EpiNull, 3>(ta, tb, s, ep, 0); }, 20); printf("L2-resident N=64 K=384 MT=3: %.4f ms %.0f TFLOP/s\n", t, fl / t);
    t = time_ms([&] { launch_gemm<64, 64, 2, EpiNull, 4>(ta, tb, s, ep, 0); }, 20); printf("L2-resident N=64 K=384 MT=4: %.4f ms %.0f TFLOP/s\n", t, fl / t);
    make_tmap_rows_bf16(&ta, c1, (uint64_t)M3, 576, 576, 128, 64);
    make_tmap_rows_bf16(&tb, w, 128, 576, 576, 128, 64);
    GemmShape s3 = plain_shape(M3, 128, 576, 128, 64, 16);
    const double fl3 = 2.0 * M3 * 128 * 576 / 1e9;
    t = time_ms([&] { launch_gemm<128, 64, 6, EpiNull>(ta, tb, s3, ep, 0); }, 20); printf("L2-resident N=128 K=576 MT=1: %.4f ms %.0f TFLOP/s\n", t, fl3 / t);
    t = time_ms([&] { launch_gemm<128, 64, 4, EpiNull, 2>(ta, tb, s3, ep, 0); }, 20); printf("L2-resident N=128 K=576 MT=2: %.4f ms %.0f TFLOP/s\n", t, fl3 / t);
    make_tmap_rows_bf16(&ta, c1, (uint64_t)M3, 1152, 1152, 128, 64);
    make_tmap_rows_bf16(&tb, w, 256, 1152, 1152, 256, 64);
    GemmShape s4 = plain_shape(M3 / 2, 256, 1152, 256, 64, 16);
    const double fl4 = 2.0 * (M3 / 2) * 256 * 1152 / 1e9;
    t = time_ms([&] { launch_gemm<256, 64, 4, EpiNull>(ta, tb, s4, ep, 0); }, 20); printf("L2-resident N=256 K=1152 MT=1: %.4f ms %.0f TFLOP/s\n", t, fl4 / t);
  }
  // (2c) conv2 4-D s2d box on an L2-resident set of frames (512 frames = 32 MB), MT sweep
  {
    const int f2 = 512;
    make_tmap_nhwc_bf16(&ta, c1, f2, 16, 16, 128, 64, 16, 8, 1, 1);
    make_tmap_rows_bf16(&tb, w, 64, 384, 384, 64, 64);
    GemmShape s{};
    s.m_tiles = 2 * f2; s.n_tiles = 1; s.k_blocks = 6; s.group_m = 16; s.a_conv = 1; s.tiles_per_frame = 2; s.frames_per_tile = 1;
    s.tile_out_rows = 8; s.h_mul = 1; s.n_segments = 1;
    const int c[6] = {1, 0, 1, 1, 0, 1}, dw[6] = {0, 0, 0, -1, -1, -1}, dh[6] = {-1, 0, 0, -1, 0, 0};
    for (int k = 0; k < 6; ++k) { s.tap_c_blk[k] = c[k]; s.tap_w[k] = dw[k]; s.tap_h[k] = dh[k]; }
    const double fl = 2.0 * f2 * 256 * 64 * 384 / 1e9;
    float t;
    t = time_ms([&] { launch_gemm<64, 64, 8, EpiNull>(ta, tb, s, ep, 0); }, 20); printf("conv2 4-D box, 512 frames MT=1: %.4f ms %.0f TFLOP/s\n", t, fl / t);
    t = time_ms([&] { launch_gemm<64, 64, 3, EpiNull, 3>(ta, tb, s, ep, 0); }, 20); printf("conv2 4-D box, 512 frames MT=3: %.4f ms %.0f TFLOP/s\n", t, fl / t);
  }
  // (3) N = 256, K = 256 token GEMM shape on the same buffer: [M][256]
  {
    const long long M = 1 << 20;  // 512 MB of the 1 GB buffer
    make_tmap_rows_bf16(&ta, c1, M, 256, 256, 128, 64);
    make_tmap_rows_bf16(&tb, w, 256, 256, 256, 256, 64);
    GemmShape s = plain_shape(M, 256, 256, 256, 64, 16);
    const float t1 = time_ms([&] { launch_gemm_bres<256, 64, 4, 4, EpiNull>(ta, tb, s, ep, 0); });
    const float t2 = time_ms([&] { launch_gemm<256, 64, 4, EpiNull>(ta, tb, s, ep, 0); });
    printf("token GEMM M=1M N=256 K=256: B resident %.3f ms (%.0f TFLOP/s), generic %.3f ms (%.0f TFLOP/s)\n", t1,
           2.0 * M * 256 * 256 / t1 / 1e9, t2, 2.0 * M * 256 * 256 / t2 / 1e9);
  }
  CK(cudaDeviceSynchronize());
  cudaFree(c1); cudaFree(w);
}

int main(int argc, char**) {
  if (argc > 1) { perf(); return 0; }
  int fails = 0;
  fails += run_plain_mt<64, 64, 3, 3>(5000, 64, 384);
  fails += run_plain_mt<128, 64, 4, 2>(3333, 128, 576);
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
