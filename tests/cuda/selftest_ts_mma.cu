// Device self-test + microbenchmark for the "TS" form of tcgen05.mma: the A operand comes from TENSOR MEMORY (written
// there by tcgen05.st from registers), B from a SWIZZLE_128B shared-memory tile. This is the form the fused stem uses
// for conv1 (3 -> 32 channels, K = 75 padded to 80): the im2col rows are built in registers, so they never have to
// exist in shared memory and the tensor core does not spend its smem read port on them.
//   (1) correctness: D[128][32] = A[128][80] * B[32][80]^T against a CPU fp32 reference (bf16 inputs)
//   (2) issue interval of M128 N32 K16 TS-mode UMMAs (1 and 2 accumulators)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/selftest_ts_mma tests/cuda/selftest_ts_mma.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../video_fingerprint_b200/csrc/sm100_primitives.cuh"

using namespace vfp;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kM = 128, kN = 32, kK = 80;

// a: [128][80] bf16, b: [32][80] bf16, d: [128][32] fp32, clk: [4]
__global__ void __launch_bounds__(160, 1) ts_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, float* d, long long* clk, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* wbuf = smem;  // 2 blocks of [32 rows][128 B] SWIZZLE_128B = 8 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 8192 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(wbuf)[i] = 0;
  __syncthreads();
  for (int i = tid; i < kN * kK; i += blockDim.x) {
    const int n = i / kK, k = i % kK;
    const int kb = k >> 6, kk = k & 63;
    const int chunk = (kk >> 3) ^ (n & 7);
    *reinterpret_cast<__nv_bfloat16*>(wbuf + kb * 4096 + n * 128 + chunk * 16 + (kk & 7) * 2) = b[i];
  }
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 4) { tmem_alloc(slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  constexpr uint32_t kColD = 0, kColA = 64;
  if (warp < 4) {
    uint32_t v[40];
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(a + (size_t)tid * kK);
#pragma unroll
    for (int j = 0; j < 40; ++j) v[j] = arow[j];
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16) + kColA;
#pragma unroll
    for (int s = 0; s < 5; ++s)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr + 8 * s),
                   "r"(v[8 * s]), "r"(v[8 * s + 1]), "r"(v[8 * s + 2]), "r"(v[8 * s + 3]), "r"(v[8 * s + 4]), "r"(v[8 * s + 5]),
                   "r"(v[8 * s + 6]), "r"(v[8 * s + 7])
                   : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4 && lane == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(kM, kN);
    const uint32_t w_base = smem_u32(wbuf);
    auto issue = [&](uint32_t dcol, int s, uint32_t acc) {
      const uint64_t bdesc = umma_smem_desc_kmajor<128>(w_base + (s >> 2) * 4096) + 2 * (s & 3);
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
          "}\n" ::"r"(tmem + dcol),
          "r"(tmem + kColA + 8 * s), "l"(bdesc), "r"(idesc), "r"(acc)
          : "memory");
    };
    for (int s = 0; s < 5; ++s) issue(kColD, s, s > 0);
    umma_commit(&bar[0]);
    mbar_wait(&bar[0], 0);
    // timing: chains of 5 K-steps into a scratch accumulator (columns 128..191), 1 or 2 accumulators
    for (int nacc = 1; nacc <= 2; ++nacc) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r)
        for (int s = 0; s < 5; ++s) issue(128 + (nacc == 2 ? (r & 1) * 32 : 0), s, s > 0);
      umma_commit(&bar[1]);
      mbar_wait(&bar[1], (uint32_t)(nacc - 1));
      clk[nacc - 1] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 4) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + kColD, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) d[tid * kN + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int main() {
  std::vector<__nv_bfloat16> ha(kM * kK), hb(kN * kK);
  std::vector<float> fa(kM * kK), fb(kN * kK);
  srand(5);
  for (int i = 0; i < kM * kK; ++i) { ha[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f); fa[i] = __bfloat162float(ha[i]); }
  for (int i = 0; i < kN * kK; ++i) { hb[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f); fb[i] = __bfloat162float(hb[i]); }
  __nv_bfloat16 *da, *db; float* dd; long long* dclk;
  CK(cudaMalloc(&da, ha.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2)); CK(cudaMalloc(&dd, kM * kN * 4)); CK(cudaMalloc(&dclk, 32));
  CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  const int reps = 400;
  ts_kernel<<<1, 160, 16384>>>(da, db, dd, dclk, reps);
  CK(cudaDeviceSynchronize());
  std::vector<float> hd(kM * kN);
  long long clk[2];
  CK(cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(clk, dclk, 16, cudaMemcpyDeviceToHost));
  double max_err = 0;
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < kN; ++n) {
      double s = 0;
      for (int k = 0; k < kK; ++k) s += (double)fa[m * kK + k] * fb[n * kK + k];
      max_err = fmax(max_err, fabs(s - hd[m * kN + n]));
    }
  printf("TS-mode UMMA M128 N32 K80: max abs err %.3e  %s\n", max_err, max_err < 1e-3 ? "PASS" : "FAIL");
  printf("TS-mode issue interval: %.1f clk/UMMA (1 acc), %.1f clk/UMMA (2 acc)\n", clk[0] / (5.0 * reps), clk[1] / (5.0 * reps));
  return max_err < 1e-3 ? 0 : 1;
}
