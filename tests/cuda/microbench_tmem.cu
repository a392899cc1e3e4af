// TMEM <-> register bandwidth microbenchmark (run on a B200): N warps of one CTA stream tcgen05.ld / tcgen05.st over their
// TMEM lane quarter. The stem kernel and every small-K GEMM epilogue are paced by these numbers.
#include <cstdio>
#include <cstdlib>
#include "../../video_fingerprint_b200/csrc/sm100_primitives.cuh"
using namespace vfp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}

// mode 0: loads, mode 1: stores; every warp does `iters` x (4 x 32 columns) on its lane quarter
__global__ void __launch_bounds__(1024, 1) tmem_bw(int mode, int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128 % 512;
  uint32_t v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = lane + j;
  __syncthreads();
  const long long t0 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_32x32(base + 32 * c, v);
        tmem_ld_wait();
        acc += v[0] + v[31];
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) st32(base + 32 * c, v);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* d; uint32_t* sink;
  CK(cudaMalloc(&d, 64)); CK(cudaMalloc(&sink, 4096));
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 16, 32}) {
      tmem_bw<<<1, warps * 32>>>(mode, iters, d, sink);
      CK(cudaDeviceSynchronize());
      long long h;
      CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
      const double bytes = (double)warps * iters * 4 * 32 * 32 * 4;
      printf("%s %2d warps: %.1f B/clk per SM\n", mode ? "tcgen05.st" : "tcgen05.ld", warps, bytes / h);
    }
  return 0;
}
