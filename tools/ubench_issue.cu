// Issue-rate micro-benchmark (sm_100a): cycles per warp instruction of the opcodes the attention softmax is made of, with
// 8 independent register chains per thread, 1 / 2 / 4 warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/ubench_issue.cu -o tools/ab/ubench_issue
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 256;

template <int OP>
__device__ __forceinline__ void body(float (&a)[16], uint32_t (&u)[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(a[8 + k]), "f"(a[15 - k]));
    if (OP == 1) {   // FFMA2
      asm volatile(
          "{ .reg .b64 x, y, z;\n\t"
          "mov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %2};\n\tmov.b64 z, {%3, %3};\n\t"
          "fma.rn.f32x2 x, x, y, z;\n\t"
          "mov.b64 {%0, %1}, x; }"
          : "+f"(a[2 * (k & 3)]), "+f"(a[2 * (k & 3) + 1])
          : "f"(a[8 + k]), "f"(a[15 - (k & 3)]));
    }
    if (OP == 2) {   // FADD2
      asm volatile(
          "{ .reg .b64 x, y;\n\t"
          "mov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %2};\n\t"
          "add.rn.f32x2 x, x, y;\n\t"
          "mov.b64 {%0, %1}, x; }"
          : "+f"(a[2 * (k & 3)]), "+f"(a[2 * (k & 3) + 1])
          : "f"(a[8 + k]));
    }
    if (OP == 3) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "+r"(u[k]) : "f"(a[k]), "f"(__uint_as_float(u[(k + 3) & 7])));
    if (OP == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(a[8 + k]), "f"(a[15 - k]));
    if (OP == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(a[8 + k]));
    if (OP == 6) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
    if (OP == 7) asm volatile("{ .reg .b32 t; shl.b32 t, %1, 23; add.u32 %0, %0, t; }" : "+r"(u[k]) : "r"(u[(k + 1) & 7]));
    if (OP == 8) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(a[8 + k]));
    if (OP == 9) {   // MUFU with one FFMA2 behind each (co-issue)
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
      asm volatile(
          "{ .reg .b64 x, y, z;\n\t"
          "mov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %2};\n\tmov.b64 z, {%3, %3};\n\t"
          "fma.rn.f32x2 x, x, y, z;\n\t"
          "mov.b64 {%0, %1}, x; }"
          : "+f"(a[8 + 2 * (k & 1)]), "+f"(a[9 + 2 * (k & 1)])
          : "f"(a[12]), "f"(a[13]));
    }
    if (OP == 10) {   // MUFU + 3 plain FFMA behind each
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[8]) : "f"(a[12]), "f"(a[13]));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[9]) : "f"(a[12]), "f"(a[13]));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[10]) : "f"(a[12]), "f"(a[13]));
    }
    if (OP == 11) asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u[k]) : "r"(u[(k + 1) & 7]), "r"(u[(k + 2) & 7]));
    if (OP == 12) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(a[8 + k]));
  }
}

template <int OP>
__global__ void __launch_bounds__(512, 1) k(long long* out, float* sink) {
  float a[16];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 0.001f * static_cast<float>(threadIdx.x + i) - 0.3f;
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = threadIdx.x * 977 + i;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    body<OP>(a, u);
    body<OP>(a, u);
    body<OP>(a, u);
    body<OP>(a, u);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float(u[i] & 0x3fffffffu);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) {
    out[(blockIdx.x * 16 + (threadIdx.x >> 5)) * 2] = t0;
    out[(blockIdx.x * 16 + (threadIdx.x >> 5)) * 2 + 1] = t1;
  }
}

template <int OP>
void run(const char* name, int per_body) {
  long long* d_out;
  float* d_sink;
  cudaMalloc(&d_out, sizeof(long long) * 148 * 32);
  cudaMalloc(&d_sink, sizeof(float) * 148 * 512);
  printf("%-34s", name);
  for (int threads : {128, 256, 512}) {
    for (int i = 0; i < 2; ++i) k<OP><<<148, threads>>>(d_out, d_sink);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); exit(1); }
    long long h[32];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    long long lo = h[0], hi = h[1];
    for (int w = 0; w < threads / 32; ++w) {
      if (h[2 * w] < lo) lo = h[2 * w];
      if (h[2 * w + 1] > hi) hi = h[2 * w + 1];
    }
    const double n_inst = static_cast<double>(ITERS) * 4 * per_body * (threads / 128);   // warp instructions per scheduler
    printf("  %dw/sched: %5.2f cyc/inst", threads / 128, static_cast<double>(hi - lo) / n_inst);
  }
  printf("\n");
}

int main() {
  run<0>("FFMA", 8);
  run<1>("FFMA2", 8);
  run<2>("FADD2", 8);
  run<8>("FADD", 8);
  run<12>("FMUL", 8);
  run<3>("F2FP.BF16.PACK_AB", 8);
  run<4>("FMNMX3", 8);
  run<5>("FMNMX", 8);
  run<6>("MUFU.EX2", 8);
  run<7>("SHL + IADD (LEA?)", 8);
  run<11>("PRMT", 8);
  run<9>("MUFU + FFMA2 (per pair of inst)", 8);
  run<10>("MUFU + 3 FFMA (per group of 4)", 8);
  return 0;
}
