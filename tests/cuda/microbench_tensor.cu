// Tensor-pipe microbenchmarks behind the design numbers in DESIGN.md (run on a B200: build/microbench_tensor).
//  (1) UMMA issue interval when consecutive tcgen05.mma accumulate into the same / rotating TMEM tiles, N = 32..256
//  (2) legacy mma.sync (HMMA.16816 bf16) throughput per SM for 4..16 resident warps
//  (3) both at once (the fused stem kernel runs conv1 on mma.sync next to conv2 on tcgen05)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../video_fingerprint_b200/csrc/sm100_primitives.cuh"

using namespace vfp;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// warp 0: UMMA issuer (if n_umma > 0); warps 1..: HMMA loops (if n_hmma > 0)
template <int N, int NACC>
__global__ void __launch_bounds__(1024, 1) bench_kernel(int n_umma, int n_hmma, long long* out, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (lane == 0 && n_umma > 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, N);
      const uint64_t adesc = umma_smem_desc_kmajor<128>(smem_u32(smem));
      const uint64_t bdesc = umma_smem_desc_kmajor<128>(smem_u32(smem + 16384));
      t0 = clock64();
      // fully unrolled groups of 16 so that the issuing thread spends no time on index arithmetic
      for (int i = 0; i < n_umma; i += 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          umma_bf16(tmem + (j % NACC) * N, adesc + 2 * (j & 3), bdesc + 2 * (j & 3), idesc, 1u);
      }
      umma_commit(bar);
      mbar_wait(bar, 0);
      t1 = clock64();
      out[blockIdx.x * 4 + 0] = t1 - t0;
    }
  } else if (n_hmma > 0) {
    float acc[8][4];
    uint32_t a[4] = {0x3c003c00u + lane, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x3c003c00u + lane};
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[j][k] = 0.f;
    __syncwarp();
    const long long h0 = clock64();
    for (int i = 0; i < n_hmma; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) hmma(acc[j], a, b);
    }
    const long long h1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[j][0] + acc[j][1] + acc[j][2] + acc[j][3];
    if (s == 123.456f) sink[threadIdx.x] = s;
    if (warp == 1 && lane == 0) out[blockIdx.x * 4 + 1] = h1 - h0;
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int NACC>
static void run(const char* label, int threads, int n_umma, int n_hmma, long long* d_out, float* sink) {
  const int n_acc = NACC;
  static bool cfg = false;
  if (!cfg) { CK(cudaFuncSetAttribute(bench_kernel<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536)); cfg = true; }
  CK(cudaMemset(d_out, 0, 148 * 4 * 8));
  bench_kernel<N, NACC><<<148, threads, 65536>>>(n_umma, n_hmma, d_out, sink);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(148 * 4);
  CK(cudaMemcpy(h.data(), d_out, h.size() * 8, cudaMemcpyDeviceToHost));
  double cu = 0, ch = 0;
  for (int i = 0; i < 148; ++i) { cu += h[i * 4]; ch += h[i * 4 + 1]; }
  cu /= 148; ch /= 148;
  printf("%-34s N=%3d acc=%d warps=%2d :", label, N, n_acc, threads / 32 - 1);
  if (n_umma) printf("  UMMA %.1f clk/inst (ideal %d)", cu / n_umma, 128 * N / 256);
  if (n_hmma) printf("  HMMA %.2f clk per HMMA per SM-quadrant (warp-level %.1f clk/HMMA)", ch / (n_hmma * 8.0) / ((threads / 32 - 1) / 4.0 > 1 ? (threads / 32 - 1) / 4.0 : 1.0) , ch / (n_hmma * 8.0));
  printf("\n");
}

int main() {
  long long* d_out; float* sink;
  CK(cudaMalloc(&d_out, 148 * 4 * 8));
  CK(cudaMalloc(&sink, 4096));
  run<64, 1>("warm-up", 64, 64, 0, d_out, sink);
  run<64, 1>("UMMA chain", 64, 240, 0, d_out, sink);
  run<64, 2>("UMMA chain", 64, 240, 0, d_out, sink);
  run<64, 4>("UMMA chain", 64, 240, 0, d_out, sink);
  run<64, 8>("UMMA chain", 64, 240, 0, d_out, sink);
  run<32, 1>("UMMA chain", 64, 240, 0, d_out, sink);
  run<32, 4>("UMMA chain", 64, 240, 0, d_out, sink);
  run<32, 16>("UMMA chain", 64, 240, 0, d_out, sink);
  run<128, 1>("UMMA chain", 64, 240, 0, d_out, sink);
  run<128, 2>("UMMA chain", 64, 240, 0, d_out, sink);
  run<128, 4>("UMMA chain", 64, 240, 0, d_out, sink);
  run<256, 1>("UMMA chain", 64, 240, 0, d_out, sink);
  run<256, 2>("UMMA chain", 64, 240, 0, d_out, sink);
  for (int w : {4, 12}) run<64, 1>("HMMA only", 32 * (w + 1), 0, 2000, d_out, sink);
  for (int w : {4, 12}) run<64, 4>("HMMA + UMMA (N=64, 4 acc)", 32 * (w + 1), 40000, 2000, d_out, sink);
  for (int w : {12}) run<256, 2>("HMMA + UMMA (N=256, 2 acc)", 32 * (w + 1), 10000, 2000, d_out, sink);
  for (int w : {12}) run<32, 4>("HMMA + UMMA (N=32, 4 acc)", 32 * (w + 1), 40000, 2000, d_out, sink);
  printf("(HMMA column: cycles between HMMA issues on ONE SM sub-partition with warps/4 warps resident there)\n");
  return 0;
}
