// Micro-benchmark of the softmax inner loops of attn_space_tc_kernel in isolation: one thread per TMEM lane walks 256
// fp32 columns (8 chunks of 32) from TMEM and writes 128 packed bf16 columns back -- no MMA, no TMA, no barriers.
// Answers: what does one pass cost per 32-column chunk with 1 / 2 warps per scheduler, and which instruction order /
// MUFU share gets it down.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../helping_hand..../csrc
//   tools/ubench_softmax.cu -o tools/ab/ubench_softmax ; run: tools/ab/ubench_softmax
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "hh_ptx.cuh"

using namespace hh;

constexpr float LOG2E = 1.4426950408889634f;
constexpr int REPS = 16;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }

// 2^x for x <= 0 on the FMA / ALU pipes: round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-3 minimax of 2^f,
// exponent add.  Clamped at -125 (result 2^-125: nothing next to a row sum >= 1).
__device__ __forceinline__ float2 poly_exp2x2(float2 x) {
  const float MAGIC = 12582912.f;   // 1.5 * 2^23
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 xf = fadd2(x, make_float2(MAGIC, MAGIC));
  const float2 fl = fadd2(xf, make_float2(-MAGIC, -MAGIC));
  const float2 f = fadd2(x, make_float2(-fl.x, -fl.y));
  // 2^f on [-0.5, 0.5]: c0 + c1 f + c2 f^2 + c3 f^3
  const float c0 = 0.999928074f, c1 = 0.693260986f, c2 = 0.242611122f, c3 = 0.055171667f;   // rel. error 7.5e-5
  float2 p = ffma2(f, make_float2(c3, c3), make_float2(c2, c2));
  p = ffma2(p, f, make_float2(c1, c1));
  p = ffma2(p, f, make_float2(c0, c0));
  float2 r;
  r.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(xf.x) << 23));
  r.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(xf.y) << 23));
  return r;
}

// ---------------------------------------------------------------- exp-pass chunk bodies: v[32] -> w[16], l2 +=
template <int MODE>
__device__ __forceinline__ void emit(const uint32_t (&v)[32], uint32_t (&w)[16], float2& l2, float ml) {
  const float2 sc = make_float2(LOG2E, LOG2E), sh = make_float2(-ml, -ml);
  if (MODE == 0) {   // the shipped form: ptxas interleaves each pair's FADD2 / pack right behind its two ex2
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float2 t = ffma2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc, sh);
      const float2 e = make_float2(fast_exp2(t.x), fast_exp2(t.y));
      l2 = fadd2(l2, e);
      w[j >> 1] = pack_bf16x2(e.x, e.y);
    }
  } else if (MODE == 1) {   // 4 independent sum chains
    float2 a[4] = {l2, make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float2 t = ffma2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc, sh);
      const float2 e = make_float2(fast_exp2(t.x), fast_exp2(t.y));
      a[(j >> 1) & 3] = fadd2(a[(j >> 1) & 3], e);
      w[j >> 1] = pack_bf16x2(e.x, e.y);
    }
    l2 = fadd2(fadd2(a[0], a[1]), fadd2(a[2], a[3]));
  } else if (MODE == 2) {   // MUFU only: no sum, truncating pack by PRMT (floor of what the rest costs)
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float2 t = ffma2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc, sh);
      const float2 e = make_float2(fast_exp2(t.x), fast_exp2(t.y));
      w[j >> 1] = __byte_perm(__float_as_uint(e.x), __float_as_uint(e.y), 0x7632);
    }
  } else if (MODE == 3) {   // every second pair on the FMA pipe (50 % of the exponentials emulated)
    float2 a[2] = {l2, make_float2(0.f, 0.f)};
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float2 t0 = ffma2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc, sh);
      const float2 t1 = ffma2(make_float2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), sc, sh);
      const float2 e0 = make_float2(fast_exp2(t0.x), fast_exp2(t0.y));
      const float2 e1 = poly_exp2x2(t1);
      a[0] = fadd2(a[0], e0);
      a[1] = fadd2(a[1], e1);
      w[j >> 1] = pack_bf16x2(e0.x, e0.y);
      w[(j >> 1) + 1] = pack_bf16x2(e1.x, e1.y);
    }
    l2 = fadd2(a[0], a[1]);
  } else if (MODE == 4) {   // one pair in four emulated (25 %)
    float2 a[2] = {l2, make_float2(0.f, 0.f)};
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      float2 t[4], e[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        t[q] = ffma2(make_float2(__uint_as_float(v[j + 2 * q]), __uint_as_float(v[j + 2 * q + 1])), sc, sh);
#pragma unroll
      for (int q = 0; q < 3; ++q) e[q] = make_float2(fast_exp2(t[q].x), fast_exp2(t[q].y));
      e[3] = poly_exp2x2(t[3]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a[q & 1] = fadd2(a[q & 1], e[q]);
        w[(j >> 1) + q] = pack_bf16x2(e[q].x, e[q].y);
      }
    }
    l2 = fadd2(a[0], a[1]);
  } else if (MODE == 5) {   // all on the FMA pipe (issue-slot cost of the emulation)
    float2 a[2] = {l2, make_float2(0.f, 0.f)};
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float2 t = ffma2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc, sh);
      const float2 e = poly_exp2x2(t);
      a[(j >> 1) & 1] = fadd2(a[(j >> 1) & 1], e);
      w[j >> 1] = pack_bf16x2(e.x, e.y);
    }
    l2 = fadd2(a[0], a[1]);
  } else if (MODE == 6) {   // ex2 batched 8 ahead of their consumers through exact fake dependencies
    float e[32];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float2 t = ffma2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc, sh);
      e[j] = fast_exp2(t.x);
      e[j + 1] = fast_exp2(t.y);
    }
#pragma unroll
    for (int j = 0; j < 24; ++j) e[j] = fmaf(e[j + 8], 0.f, e[j]);
    float2 a[2] = {l2, make_float2(0.f, 0.f)};
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      a[(j >> 1) & 1] = fadd2(a[(j >> 1) & 1], make_float2(e[j], e[j + 1]));
      w[j >> 1] = pack_bf16x2(e[j], e[j + 1]);
    }
    l2 = fadd2(a[0], a[1]);
  } else if (MODE == 7) {   // no exponentials at all: scale, sum, pack (everything but the MUFU)
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float2 e = ffma2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc, sh);
      l2 = fadd2(l2, e);
      w[j >> 1] = pack_bf16x2(e.x, e.y);
    }
  }
}

template <int MODE>
__device__ __forceinline__ float exp_pass(uint32_t s_col, float ml) {
  uint32_t va[32], vb[32], w[16];
  float2 l2 = make_float2(0.f, 0.f);
  tmem_ld_32x32b_x32(s_col, va);
#pragma unroll 1
  for (int c = 0; c < 8; c += 2) {
    tmem_ld_wait();
    tmem_ld_32x32b_x32(s_col + (c + 1) * 32, vb);
    emit<MODE>(va, w, l2, ml);
    tmem_st_32x32b_x16(s_col + c * 16, w);
    tmem_ld_wait();
    if (c + 2 < 8) tmem_ld_32x32b_x32(s_col + (c + 2) * 32, va);
    emit<MODE>(vb, w, l2, ml);
    tmem_st_32x32b_x16(s_col + (c + 1) * 16, w);
  }
  tmem_st_wait();
  return l2.x + l2.y;
}

// three-stage register pipeline (x <- ld(c+1), t = scaled c, e = 2^t of c-1), G = columns per iteration (32 or 16)
template <int G>
__device__ __forceinline__ float exp_pass_pipe(uint32_t s_col, float ml) {
  constexpr int NP = G / 2, NI = 256 / G;
  const float2 sc = make_float2(LOG2E, LOG2E), sh = make_float2(-ml, -ml);
  uint32_t x[G];
  float2 t[NP], e[NP];
  float2 l2 = make_float2(0.f, 0.f), l2b = make_float2(0.f, 0.f);
  auto ld = [&](int c) {
    if constexpr (G == 32) tmem_ld_32x32b_x32(s_col + c * G, reinterpret_cast<uint32_t(&)[32]>(x));
    else tmem_ld_32x32b_x16(s_col + c * G, reinterpret_cast<uint32_t(&)[16]>(x));
  };
  auto scale = [&]() {
#pragma unroll
    for (int j = 0; j < NP; ++j) t[j] = ffma2(make_float2(__uint_as_float(x[2 * j]), __uint_as_float(x[2 * j + 1])), sc, sh);
  };
  auto exps = [&]() {
#pragma unroll
    for (int j = 0; j < NP; ++j) e[j] = make_float2(fast_exp2(t[j].x), fast_exp2(t[j].y));
  };
  auto emitp = [&](int c) {
    uint32_t w[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      if (j & 1) l2b = fadd2(l2b, e[j]); else l2 = fadd2(l2, e[j]);
      w[j] = pack_bf16x2(e[j].x, e[j].y);
    }
    if constexpr (G == 32) tmem_st_32x32b_x16(s_col + c * NP, reinterpret_cast<uint32_t(&)[16]>(w));
    else tmem_st_32x32b_x8(s_col + c * NP, reinterpret_cast<uint32_t(&)[8]>(w));
  };
  ld(0);
  tmem_ld_wait();
  scale();
  ld(1);
  exps();
  tmem_ld_wait();
  scale();
#pragma unroll 1
  for (int c = 1; c < NI; ++c) {
    ld((c + 1) & (NI - 1));
    emitp(c - 1);
    exps();
    tmem_ld_wait();
    scale();
  }
  emitp(NI - 1);
  tmem_st_wait();
  return l2.x + l2.y + l2b.x + l2b.y;
}

// ---------------------------------------------------------------- max-pass variants
template <int MODE>
__device__ __forceinline__ float max_pass(uint32_t s_col) {
  uint32_t va[32], vb[32];
  float m[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
  auto fold = [&](const uint32_t(&v)[32]) {
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 2)
        asm("max.f32 %0, %0, %1, %2;" : "+f"(m[0]) : "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])));
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 2)
        asm("max.f32 %0, %0, %1, %2;" : "+f"(m[(j >> 1) & 3]) : "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])));
    }
  };
  if (MODE == 2) {   // load-only: TMEM read throughput of this access pattern
    tmem_ld_32x32b_x32(s_col, va);
#pragma unroll 1
    for (int c = 0; c < 8; c += 2) {
      tmem_ld_32x32b_x32(s_col + (c + 1) * 32, vb);
      tmem_ld_wait();
      m[0] = fmaxf(m[0], __uint_as_float(va[0]));
      m[1] = fmaxf(m[1], __uint_as_float(vb[0]));
      if (c + 2 < 8) tmem_ld_32x32b_x32(s_col + (c + 2) * 32, va);
    }
    tmem_ld_wait();
    return fmaxf(m[0], m[1]);
  }
  tmem_ld_32x32b_x32(s_col, va);
#pragma unroll 1
  for (int c = 0; c < 8; c += 2) {
    tmem_ld_wait();
    tmem_ld_32x32b_x32(s_col + (c + 1) * 32, vb);
    fold(va);
    tmem_ld_wait();
    if (c + 2 < 8) tmem_ld_32x32b_x32(s_col + (c + 2) * 32, va);
    fold(vb);
  }
  return fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
}

// test ids: 0..7 exp pass MODE, 10..12 max pass MODE, 20: max pass (4 chains) + exp pass MODE 1 back to back
// BG: tensor-core work running next to the softmax warps (warp 8, one thread): 0 none, 1 S-like (128x256x64 from shared
// memory into TMEM columns [256, 512)), 2 P V-like (A = 128 TMEM columns, B from shared memory, N = 64), 3 both in turn
template <int TEST, int BG, int COMP = 0>
__global__ void __launch_bounds__(512, 1) bench_kernel(long long* out, float* sink) {
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bg_bar;
  __shared__ volatile int stop_flag;
  extern __shared__ uint8_t dyn_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  if (threadIdx.x == 0) {
    mbar_init(&bg_bar, 1);
    fence_mbar_init();
    stop_flag = 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int wq = warp & 3, hf = (warp >> 2) & 1;   // warps 0-3: half 0, 4-7: half 1; warps >= 8 idle here
  const uint32_t s_col = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + static_cast<uint32_t>(hf * 256);
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  const int nsoft = (BG || COMP) ? 4 : min(8, static_cast<int>(blockDim.x) / 32);
  if (warp < nsoft) {
    // fill the 256 columns with logits around 0 (|s| < 8)
    uint32_t w[16];
    for (int c = 0; c < 16; ++c) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        w[j] = __float_as_uint(static_cast<float>(((threadIdx.x * 37 + (c * 16 + j) * 101) % 1024) - 512) * (1.f / 64.f));
      tmem_st_32x32b_x16(s_col + c * 16, w);
    }
    tmem_st_wait();
  }
  __syncthreads();
  if (warp < nsoft) {
    t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REPS; ++r) {
      if (TEST == 8) acc += exp_pass_pipe<32>(s_col, 8.f * LOG2E + acc * 1e-30f);
      else if (TEST == 9) acc += exp_pass_pipe<16>(s_col, 8.f * LOG2E + acc * 1e-30f);
      else if (TEST < 10) acc += exp_pass<TEST>(s_col, 8.f * LOG2E + acc * 1e-30f);
      else if (TEST < 20) acc += max_pass<TEST - 10>(s_col);
      else {
        const float mx = max_pass<1>(s_col);
        acc += exp_pass<1>(s_col, mx * LOG2E);
      }
    }
    t1 = clock64();
  }
  if (COMP != 0 && warp >= 4 && warp < 8) {   // a different instruction stream on the same schedulers
    const uint32_t c_col = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + 256u;
    float a = static_cast<float>(threadIdx.x), b = 1.0001f, cacc = 0.f;
    while (!stop_flag) {
      if (COMP == 1) cacc += max_pass<1>(c_col);
      else if (COMP == 2) {
#pragma unroll
        for (int i = 0; i < 64; ++i) { a = fmaf(a, b, 0.5f); b = fmaf(b, 1.0001f, a * 1e-9f); }
        cacc += a + b;
      } else {
        __nanosleep(200);
      }
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = cacc;
  }
  if (BG != 0 && warp == 8 && lane == 0) {
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dyn_smem) + 1023) & ~uintptr_t(1023));
    const uint32_t a_s = smem_u32(sm), b_s = a_s + 16384;
    const uint32_t idesc_s = umma_idesc_bf16(128, 256), idesc_o = umma_idesc_bf16_bmn(128, 64);
    uint32_t ph = 0;
    int groups = 0;
    while (!stop_flag) {
      if (BG == 1 || (BG == 3 && (groups & 1) == 0)) {
        const uint64_t da = umma_desc_sw128(a_s), db = umma_desc_sw128(b_s);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + 256, da + 2 * k, db + 2 * k, idesc_s, k > 0 ? 1u : 0u);
      } else {
        const uint64_t dv = umma_desc_sw128_mn(b_s);
#pragma unroll
        for (int k = 0; k < 16; ++k)
          umma_bf16_ts(tmem_base + 256 + 192, tmem_base + 256 + 8 * k, dv + static_cast<uint64_t>(k * 128), idesc_o, k > 0 ? 1u : 0u);
      }
      umma_commit(&bg_bar);
      mbar_wait(&bg_bar, ph);
      ph ^= 1u;
      ++groups;
    }
    out[148 * 16 + blockIdx.x] = groups;
  }
  if (warp < nsoft) {
    asm volatile("bar.sync 1, %0;" ::"r"(nsoft * 32) : "memory");
    if (threadIdx.x == 0) stop_flag = 1;
  }
  if (lane == 0 && warp < nsoft) {
    out[(blockIdx.x * 8 + warp) * 2] = t0;
    out[(blockIdx.x * 8 + warp) * 2 + 1] = t1;
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int TEST, int BG = 0, int COMP = 0>
void run(const char* name, int threads) {
  long long* d_out;
  float* d_sink;
  const int grid = 148;
  cudaMalloc(&d_out, sizeof(long long) * grid * 17);
  if (BG) cudaFuncSetAttribute(bench_kernel<TEST, BG, COMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  cudaMalloc(&d_sink, sizeof(float) * grid * 512);
  cudaMemset(d_out, 0, sizeof(long long) * grid * 16);
  for (int i = 0; i < 3; ++i) bench_kernel<TEST, BG, COMP><<<grid, (BG || COMP) ? 288 : threads, BG ? 66 * 1024 : 0>>>(d_out, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: CUDA error %s\n", name, cudaGetErrorString(e));
    exit(1);
  }
  long long h[16];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const int nw = threads / 32 < 8 ? threads / 32 : 8;
  long long lo = h[0], hi = h[1];
  for (int w = 0; w < nw; ++w) {
    if (h[2 * w] < lo) lo = h[2 * w];
    if (h[2 * w + 1] > hi) hi = h[2 * w + 1];
  }
  const double per_pass = static_cast<double>(hi - lo) / REPS;
  printf("%-44s warps/SMSP %d: %8.0f cycles per 256-column pass, %6.1f per 32-column chunk", name, nw / 4, per_pass,
         per_pass / 8);
  if (BG) {
    long long g = 0;
    cudaMemcpy(&g, d_out + 148 * 16, sizeof(g), cudaMemcpyDeviceToHost);
    printf("   [%lld background MMA groups = one per %.0f cycles]", g, static_cast<double>(hi - lo) / (g > 0 ? g : 1));
  }
  printf("\n");
  cudaFree(d_out);
  cudaFree(d_sink);
}

int main() {
  // softmax warps of half 0 only (128 threads) with the tensor core busy on half 1
  run<8, 0, 1>("exp 3-stage x32 | max-pass loop on the other warp", 128);
  run<8, 0, 2>("exp 3-stage x32 | FFMA loop on the other warp", 128);
  run<8, 0, 3>("exp 3-stage x32 | sleeping other warp", 128);
  run<0, 0, 1>("exp shipped | max-pass loop on the other warp", 128);
  run<0, 0, 2>("exp shipped | FFMA loop on the other warp", 128);
  run<8, 3, 1>("exp 3-stage x32 | max-pass loop + MMAs", 128);
  run<8, 0>("exp 3-stage x32, no MMA", 128);
  run<8, 3>("exp 3-stage x32 + both MMA kinds", 128);
  run<9, 0>("exp 3-stage x16, no MMA", 128);
  run<9, 3>("exp 3-stage x16 + both MMA kinds", 128);
  run<0, 0>("exp shipped, no MMA (288-thread CTA)", 128);
  run<0, 1>("exp shipped + S-like MMAs", 128);
  run<0, 2>("exp shipped + P V-like MMAs", 128);
  run<0, 3>("exp shipped + both", 128);
  run<11, 1>("max 4 chains + S-like MMAs", 128);
  run<11, 2>("max 4 chains + P V-like MMAs", 128);
  run<12, 1>("loads only + S-like MMAs", 128);
  run<12, 2>("loads only + P V-like MMAs", 128);
  run<2, 3>("exp MUFU + PRMT + both", 128);
  run<7, 3>("exp without ex2 + both", 128);
  for (int threads : {128, 256}) {
    run<0>("exp: shipped order", threads);
    run<1>("exp: 4 sum chains", threads);
    run<2>("exp: MUFU + PRMT pack only", threads);
    run<3>("exp: 50% on the FMA pipe", threads);
    run<4>("exp: 25% on the FMA pipe", threads);
    run<5>("exp: 100% on the FMA pipe", threads);
    run<6>("exp: ex2 batched 8 ahead", threads);
    run<7>("exp: no ex2 (scale + sum + pack)", threads);
    run<8>("exp: 3-stage register pipeline, x32", threads);
    run<9>("exp: 3-stage register pipeline, x16", threads);
    run<10>("max: 1 chain", threads);
    run<11>("max: 4 chains", threads);
    run<12>("max: loads only", threads);
    run<20>("max(4 chains) + exp(4 chains)", threads);
  }
  return 0;
}
