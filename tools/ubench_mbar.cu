// mbarrier wake-up latency (sm_100a): warp 1 arrives on a barrier at a recorded clock, warp 0 (another scheduler) waits on
// it with (a) try_wait + 10 ms suspend hint (NANOSLEEP.SYNCS), (b) try_wait + 64 ns hint, (c) try_wait without a hint,
// (d) test_wait spin; prints cycles from the arrive to the waiter's first instruction after the wait.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I <csrc> tools/ubench_mbar.cu -o tools/ab/ubench_mbar
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#include "hh_ptx.cuh"
using namespace hh;

constexpr int ROUNDS = 16;

template <int MODE>
__device__ __forceinline__ void wait_mode(uint64_t* bar, uint32_t parity) {
  if (MODE == 0) {
    while (!mbar_try_wait(bar, parity)) {}
  } else if (MODE == 1) {
    while (!mbar_try_wait_ns(bar, parity, 64)) {}
  } else if (MODE == 2) {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{ .reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.b32 %0, 1, 0, P; }"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
  } else if (MODE >= 4) {
    // a thread that serves two barriers in turn: `bar + 1` never completes, `bar` is the one that is signalled
    const uint32_t ns = MODE == 4 ? 64u : (MODE == 5 ? 0u : 16u);
    for (;;) {
      if (MODE == 7) {
        uint32_t ok = 0;
        asm volatile("{ .reg .pred P; mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2; selp.b32 %0, 1, 0, P; }"
                     : "=r"(ok) : "r"(smem_u32(bar + 1)), "r"(0u) : "memory");
        asm volatile("{ .reg .pred P; mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2; selp.b32 %0, 1, 0, P; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) break;
      } else {
        if (mbar_try_wait_ns(bar + 1, 0, ns)) break;
        if (mbar_try_wait_ns(bar, parity, ns)) break;
      }
    }
  } else {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{ .reg .pred P; mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2; selp.b32 %0, 1, 0, P; }"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
  }
}

// NW waiting warps (1 = only warp 0; 5 = warps 0, 4..7 also wait: several sleepers on the barrier's CTA)
template <int MODE>
__global__ void k(long long* out, int delay) {
  __shared__ uint64_t bars[2], back;
  uint64_t& bar = bars[0];
  __shared__ long long t_arrive[ROUNDS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&back, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == 0) {
    for (int r = 0; r < ROUNDS; ++r) {
      wait_mode<MODE>(&bar, r & 1);
      const long long t = clock64();
      __syncwarp();
      if (lane == 0) {
        out[blockIdx.x * ROUNDS + r] = t - *reinterpret_cast<volatile long long*>(&t_arrive[r]);
        mbar_arrive(&back);
      }
    }
  } else if (warp == 1 && lane == 0) {
    for (int r = 0; r < ROUNDS; ++r) {
      const long long t0 = clock64();
      while (clock64() - t0 < delay) {}
      *reinterpret_cast<volatile long long*>(&t_arrive[r]) = clock64();
      __threadfence_block();
      mbar_arrive(&bar);
      while (!mbar_try_wait_ns(&back, r & 1, 32)) {}
    }
  }
}

template <int MODE>
void run(const char* name, int delay) {
  long long* d;
  cudaMalloc(&d, sizeof(long long) * 148 * ROUNDS);
  k<MODE><<<148, 64>>>(d, delay);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); exit(1); }
  long long h[ROUNDS];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-40s delay %6d:", name, delay);
  for (int r = 0; r < ROUNDS; ++r) printf(" %lld", h[r]);
  printf("\n");
  cudaFree(d);
}

int main() {
  for (int delay : {500, 5000, 50000}) {
    run<0>("try_wait, 10 ms hint (mbar_wait)", delay);
    run<1>("try_wait, 64 ns hint", delay);
    run<2>("try_wait, no hint", delay);
    run<3>("test_wait spin", delay);
    run<4>("two barriers in turn, try_wait 64 ns", delay);
    run<6>("two barriers in turn, try_wait 16 ns", delay);
    run<5>("two barriers in turn, try_wait 0 ns", delay);
    run<7>("two barriers in turn, test_wait", delay);
  }
  return 0;
}
