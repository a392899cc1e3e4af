// Hand-off latency microbenchmark (run on a B200): round trip between two warps of one CTA through mbarriers.
//   A: arrive(bar1) ; wait(bar2)          B: wait(bar1) ; signal(bar2)
// signal = mbarrier.arrive or tcgen05.commit (with no UMMA outstanding); wait = spin on try_wait / try_wait with a
// suspend-time hint / the library's mbar_wait_relaxed. Numbers feed the buffering depth of the stem kernels.
#include <cstdio>
#include <cstdlib>
#include "../../video_fingerprint_b200/csrc/sm100_primitives.cuh"
using namespace vfp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int WAIT>
__device__ __forceinline__ void do_wait(uint64_t* bar, uint32_t parity) {
  if (WAIT == 0) { while (!mbar_try_wait(bar, parity)) {} }
  else if (WAIT == 1) { while (!mbar_try_wait_hint(bar, parity, 20000u)) {} }
  else mbar_wait_relaxed(bar, parity);
}

template <int WAIT, int COMMIT, int WHOLE_WARP>
__global__ void __launch_bounds__(256, 1) pingpong(int iters, long long* out) {
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 32); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // warp 1 = A, warp 5 = B (different SM sub-partitions: 1 % 4 != 5 % 4 is false -> use 1 and 6)
  if (warp == 1) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (WHOLE_WARP) { __syncwarp(); if (lane == 0) mbar_arrive(&bar[0]); do_wait<WAIT>(&bar[1], i & 1); }
      else if (lane == 0) { mbar_arrive(&bar[0]); do_wait<WAIT>(&bar[1], i & 1); }
    }
    if (lane == 0) out[0] = clock64() - t0;
  } else if (warp == 6) {
    for (int i = 0; i < iters; ++i) {
      if (WHOLE_WARP) {
        do_wait<WAIT>(&bar[0], i & 1);
        __syncwarp();
        if (lane == 0) { if (COMMIT) umma_commit(&bar[1]); else mbar_arrive(&bar[1]); }
      } else if (lane == 0) {
        do_wait<WAIT>(&bar[0], i & 1);
        if (COMMIT) umma_commit(&bar[1]); else mbar_arrive(&bar[1]);
      }
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 32);
}

template <int WAIT, int COMMIT, int WW>
void run(const char* label, long long* d) {
  const int iters = 2000;
  pingpong<WAIT, COMMIT, WW><<<1, 256>>>(iters, d);
  CK(cudaDeviceSynchronize());
  long long h;
  CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
  printf("%-60s %7.1f cycles per round trip\n", label, (double)h / iters);
}

int main() {
  long long* d;
  CK(cudaMalloc(&d, 64));
  run<0, 0, 0>("warm-up", d);
  run<0, 0, 0>("single lanes, arrive/arrive, spin try_wait", d);
  run<1, 0, 0>("single lanes, arrive/arrive, try_wait + 20us hint", d);
  run<2, 0, 0>("single lanes, arrive/arrive, mbar_wait_relaxed", d);
  run<0, 1, 0>("single lanes, arrive/tcgen05.commit, spin try_wait", d);
  run<1, 1, 0>("single lanes, arrive/tcgen05.commit, try_wait + hint", d);
  run<0, 0, 1>("whole warps, arrive/arrive, spin try_wait", d);
  run<1, 0, 1>("whole warps, arrive/arrive, try_wait + hint", d);
  run<2, 0, 1>("whole warps, arrive/arrive, mbar_wait_relaxed", d);
  run<2, 1, 1>("whole warps, arrive/tcgen05.commit, mbar_wait_relaxed", d);
  return 0;
}

// ---- cost of the individual synchronisation instructions, issued back to back by one lane ----
__global__ void __launch_bounds__(128, 1) opcost(int iters, long long* out) {
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 32); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1 && lane == 0) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) umma_commit(&bar[0]);
    long long t1 = clock64();
    out[0] = t1 - t0;
    // wait for the last phase flip so that the barrier is quiescent
    t0 = clock64();
    for (int i = 0; i < iters; ++i) fence_proxy_async_smem();
    t1 = clock64();
    out[1] = t1 - t0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { tc_fence_before(); tc_fence_after(); }
    t1 = clock64();
    out[2] = t1 - t0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) mbar_arrive(&bar[1]);
    t1 = clock64();
    out[3] = t1 - t0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { umma_commit(&bar[0]); umma_commit(&bar[1]); mbar_try_wait(&bar[0], 0); }
    t1 = clock64();
    out[4] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 32);
}

struct OpCostRunner {
  OpCostRunner() {
    long long* d;
    CK(cudaMalloc(&d, 64));
    const int iters = 1000;
    opcost<<<1, 128>>>(iters, d);
    CK(cudaDeviceSynchronize());
    long long h[5];
    CK(cudaMemcpy(h, d, 40, cudaMemcpyDeviceToHost));
    printf("op cost (cycles, back to back, one lane): tcgen05.commit %.1f | fence.proxy.async %.1f | tcgen05.fence pair %.1f | mbarrier.arrive %.1f | 2 commits + try_wait %.1f\n",
           h[0] / (double)iters, h[1] / (double)iters, h[2] / (double)iters, h[3] / (double)iters, h[4] / (double)iters);
  }
} g_opcost_runner;
