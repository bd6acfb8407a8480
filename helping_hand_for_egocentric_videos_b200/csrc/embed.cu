// Encoder front end (model/LaviLa.py:218-223, 540-559):
//   im2col_patches      video fp32 -> bf16 patch matrix so the 14x14/s14 (or 16x16/s16) bias-free conv becomes one
//                       tcgen05 GEMM  [B*T*n, 3*p*p (padded)] x [D, 3*p*p]^T
//   assemble_tokens_ln  prepend CLS, add the tiled spatial + repeated temporal position embeddings, apply ln_pre
#include "hh_internal.h"
#include "hh_ptx.cuh"

namespace hh {

namespace {

__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ video, bf16* __restrict__ out, int BT,
                                                     int H, int W, int p, int Kp) {
  const int gw = W / p, gh = H / p;
  const int n = gw * gh;
  const int K = 3 * p * p;
  const int half = Kp >> 1;
  const long long total = static_cast<long long>(BT) * n * half;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % half) * 2;
    const long long row = i / half;
    const int patch = static_cast<int>(row % n);
    const int img = static_cast<int>(row / n);
    float v0 = 0.f, v1 = 0.f;
    if (k < K) {  // p is even, so k and k+1 share (c, i)
      const int c = k / (p * p);
      const int rem = k - c * p * p;
      const int ii = rem / p, jj = rem - ii * p;
      const int py = patch / gw, px = patch - py * gw;
      const float* s = video + ((static_cast<size_t>(img) * 3 + c) * H + (py * p + ii)) * W + px * p + jj;
      const float2 t = *reinterpret_cast<const float2*>(s);
      v0 = t.x;
      v1 = t.y;
    }
    *reinterpret_cast<uint32_t*>(out + row * Kp + k) = pack_bf16x2(v0, v1);
  }
}

// Same patch matrix from raw decoder output: frames uint8 [BT, H, W, 3] (decord / cv2 layout, base/base_dataset.py:
// 322-323).  The loader tail  frames.float() / 255 -> permute -> NormalizeVideo(mean, std)  (data_loader/transforms.py:
// 48-51, constants run/test_EgoMCQ.py:230-233) is applied on the fly with the reference's fp32 operation order, so the
// bf16 patch values are bit-identical to normalising on the host first.  One thread = two horizontally adjacent pixels
// (6 contiguous bytes in, one bf16 pair per channel out); the K..Kp padding columns are zeroed by the tail loop.
struct NormConst { float mean[3], stdv[3]; };

__global__ void __launch_bounds__(256) im2col_u8_kernel(const uint8_t* __restrict__ frames, bf16* __restrict__ out, int BT,
                                                        int H, int W, int p, int Kp, NormConst nc) {
  const int gw = W / p, gh = H / p;
  const int n = gw * gh;
  const int hp = p >> 1;
  const long long total = static_cast<long long>(BT) * n * p * hp;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long tid0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  for (long long i = tid0; i < total; i += stride) {
    const int jj = static_cast<int>(i % hp) * 2;
    long long r = i / hp;
    const int ii = static_cast<int>(r % p);
    r /= p;  // row = img * n + patch
    const int patch = static_cast<int>(r % n);
    const int img = static_cast<int>(r / n);
    const int py = patch / gw, px = patch - py * gw;
    const uint8_t* s = frames + ((static_cast<size_t>(img) * H + (py * p + ii)) * W + px * p + jj) * 3;
    const ushort3 raw = *reinterpret_cast<const ushort3*>(s);  // 6 bytes, 2-byte aligned (even pixel column)
    const float u[6] = {static_cast<float>(raw.x & 0xff), static_cast<float>(raw.x >> 8), static_cast<float>(raw.y & 0xff),
                        static_cast<float>(raw.y >> 8), static_cast<float>(raw.z & 0xff), static_cast<float>(raw.z >> 8)};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v0 = __fdiv_rn(__fsub_rn(__fdiv_rn(u[c], 255.f), nc.mean[c]), nc.stdv[c]);
      const float v1 = __fdiv_rn(__fsub_rn(__fdiv_rn(u[3 + c], 255.f), nc.mean[c]), nc.stdv[c]);
      *reinterpret_cast<uint32_t*>(out + r * Kp + c * p * p + ii * p + jj) = pack_bf16x2(v0, v1);
    }
  }
  const int K = 3 * p * p;
  const int padh = (Kp - K) >> 1;
  if (padh > 0) {
    const long long ptotal = static_cast<long long>(BT) * n * padh;
    for (long long i = tid0; i < ptotal; i += stride) {
      const long long row = i / padh;
      const int k = K + static_cast<int>(i % padh) * 2;
      *reinterpret_cast<uint32_t*>(out + row * Kp + k) = 0u;
    }
  }
}

constexpr int MAX_VEC = 8;

__global__ void __launch_bounds__(256)
assemble_ln_kernel(const float* __restrict__ tok, const float* __restrict__ cls, const float* __restrict__ pos,
                   const float* __restrict__ temporal, const float* __restrict__ w, const float* __restrict__ b,
                   float eps, float* __restrict__ x, int B, int T, int n, int D, bf16* __restrict__ z16,
                   float* __restrict__ stats) {
  const int N = 1 + T * n;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B * N) return;
  const int bi = row / N, t = row - bi * N;
  const int nv = D >> 7;
  const float *src, *pe, *te = nullptr;
  if (t == 0) {
    src = cls;
    pe = pos;
  } else {
    const int f = (t - 1) / n, q = (t - 1) - f * n;
    src = tok + (static_cast<size_t>(bi) * T * n + (t - 1)) * D;
    pe = pos + static_cast<size_t>(1 + q) * D;
    te = temporal + static_cast<size_t>(f) * D;
  }
  float4 v[MAX_VEC];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < MAX_VEC; ++j) {
    if (j < nv) {
      const int c = (j * 32 + lane) * 4;
      float4 a = *reinterpret_cast<const float4*>(src + c);
      float4 p4 = *reinterpret_cast<const float4*>(pe + c);
      if (te) {  // reference sums (pos + temporal) first, then adds it to the token (LaviLa.py:553,557)
        const float4 t4 = *reinterpret_cast<const float4*>(te + c);
        p4.x += t4.x; p4.y += t4.y; p4.z += t4.z; p4.w += t4.w;
      }
      a.x += p4.x; a.y += p4.y; a.z += p4.z; a.w += p4.w;
      v[j] = a;
      s += (a.x + a.y) + (a.z + a.w);
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(D);
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < MAX_VEC; ++j) {
    if (j < nv) {
      v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
      ss += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(D) + eps);
  float* xr = x + static_cast<size_t>(row) * D;
#pragma unroll
  for (int j = 0; j < MAX_VEC; ++j) {
    if (j < nv) {
      const int c = (j * 32 + lane) * 4;
      const float4 ww = *reinterpret_cast<const float4*>(w + c);
      const float4 bb = *reinterpret_cast<const float4*>(b + c);
      float4 y;
      y.x = v[j].x * rstd * ww.x + bb.x;
      y.y = v[j].y * rstd * ww.y + bb.y;
      y.z = v[j].z * rstd * ww.z + bb.z;
      y.w = v[j].w * rstd * ww.w + bb.w;
      *reinterpret_cast<float4*>(xr + c) = y;
      v[j] = y;
    }
  }
  if (z16) {  // operands of the first folded-LayerNorm GEMM: bf16 copy of the row and its (sum, sum of squares)
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < MAX_VEC; ++j) {
      if (j < nv) {
        const int c = (j * 32 + lane) * 4;
        s1 += (v[j].x + v[j].y) + (v[j].z + v[j].w);
        s2 += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
        *reinterpret_cast<uint2*>(z16 + static_cast<size_t>(row) * D + c) =
            make_uint2(pack_bf16x2(v[j].x, v[j].y), pack_bf16x2(v[j].z, v[j].w));
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) *reinterpret_cast<float2*>(stats + static_cast<size_t>(row) * 2) = make_float2(s1, s2);
  }
}

}  // namespace

int im2col_patches(const float* video, bf16* out, int BT, int H, int W, int p, int Kp, cudaStream_t stream) {
  HH_REQUIRE(p % 2 == 0 && H % p == 0 && W % p == 0, "im2col_patches: patch size must be even and divide H, W");
  HH_REQUIRE(Kp % 8 == 0 && Kp >= 3 * p * p, "im2col_patches: padded K");
  HH_REQUIRE((reinterpret_cast<uintptr_t>(video) & 7) == 0 && W % 2 == 0, "im2col_patches: video alignment");
  const long long total = static_cast<long long>(BT) * (H / p) * (W / p) * (Kp / 2);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  im2col_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(video, out, BT, H, W, p, Kp);
  HH_CHECK_LAUNCH("im2col_kernel");
  return 0;
}

int im2col_patches_u8(const uint8_t* frames, const float* mean, const float* stdv, bf16* out, int BT, int H, int W, int p,
                      int Kp, cudaStream_t stream) {
  HH_REQUIRE(p % 2 == 0 && H % p == 0 && W % p == 0, "im2col_patches_u8: patch size must be even and divide H, W");
  HH_REQUIRE(Kp % 8 == 0 && Kp >= 3 * p * p, "im2col_patches_u8: padded K");
  HH_REQUIRE((reinterpret_cast<uintptr_t>(frames) & 1) == 0, "im2col_patches_u8: frames must be 2-byte aligned");
  HH_REQUIRE(stdv[0] != 0.f && stdv[1] != 0.f && stdv[2] != 0.f, "im2col_patches_u8: zero std");
  NormConst nc;
  for (int c = 0; c < 3; ++c) {
    nc.mean[c] = mean[c];
    nc.stdv[c] = stdv[c];
  }
  const long long total = static_cast<long long>(BT) * (H / p) * (W / p) * p * (p / 2);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  im2col_u8_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(frames, out, BT, H, W, p, Kp, nc);
  HH_CHECK_LAUNCH("im2col_u8_kernel");
  return 0;
}

int assemble_tokens_ln(const float* tok, const float* cls, const float* pos, const float* temporal, const float* w,
                       const float* b, float eps, float* x, int B, int T, int n, int D, cudaStream_t stream, bf16* z16,
                       float* stats) {
  HH_REQUIRE(D % 128 == 0 && D <= 128 * MAX_VEC, "assemble_tokens_ln: D must be a multiple of 128, at most 1024");
  HH_REQUIRE((z16 == nullptr) == (stats == nullptr), "assemble_tokens_ln: z16 and stats come together");
  const int rows = B * (1 + T * n);
  const int grid = (rows + 7) / 8;
  assemble_ln_kernel<<<grid, 256, 0, stream>>>(tok, cls, pos, temporal, w, b, eps, x, B, T, n, D, z16, stats);
  HH_CHECK_LAUNCH("assemble_ln_kernel");
  return 0;
}

}  // namespace hh
