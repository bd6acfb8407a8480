// Temporal half of the divided space-time attention (VarAttention with einops '(b n) f d', model/LaviLa.py:246-283)
// and the CLS query row shared by both halves.
//
// attn_time: patch query (f,p) attends {CLS key} U {(f',p) : f'}.  T+1 keys of 64 dims per problem: HBM-bound, no
// tensor cores.  One warp owns 32/TP problems (TP = frames rounded up to 4/8/16/32): the same (clip, patch) for
// consecutive heads, so every row segment it touches is a multiple of 128 contiguous bytes.  K and V are staged
// once in shared memory as fp32; lane (sub-problem, frame) keeps its q row and its 64-wide output in registers.
//
// attn_cls: the CLS query attends all 1+T*n keys (LaviLa.py:258).  One CTA per (clip, head); 8 lanes share a key
// (16 bytes each), 4 keys per warp step, online softmax per 8-lane group, groups merged through shared memory.
#include "hh_internal.h"
#include "hh_ptx.cuh"

namespace hh {

namespace {

constexpr int HD = 64;
constexpr float LOG2E = 1.4426950408889634f;
constexpr int TIME_WARPS = 4;

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// TP: padded frame count (lanes per sub-problem). Keys per problem: T + 1 <= TP + 1.
template <int TP>
__global__ void __launch_bounds__(TIME_WARPS * 32)
attn_time_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int B, int T, int n, int H) {
  constexpr int PPW = 32 / TP;       // problems (heads) per warp
  constexpr int NK = TP + 1;         // key slots per problem
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Ks = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warp) * (2 * PPW * NK * HD);
  float* Vs = Ks + PPW * NK * HD;

  const int D = H * HD;
  const int N = 1 + T * n;
  const int hgroups = (H + PPW - 1) / PPW;
  const long long task = static_cast<long long>(blockIdx.x) * TIME_WARPS + warp;
  const long long ntasks = static_cast<long long>(B) * n * hgroups;
  if (task >= ntasks) return;
  const int hg = static_cast<int>(task % hgroups);
  const int p = static_cast<int>((task / hgroups) % n);
  const int b = static_cast<int>(task / (static_cast<long long>(hgroups) * n));
  const int h0 = hg * PPW;
  const size_t ld = static_cast<size_t>(3) * D;
  const bf16* clip = qkv + static_cast<size_t>(b) * N * ld;

  // ---- stage K, V of all sub-problems: slot 0 = CLS token, slot 1+f = token (f, p); 16-byte chunks
  const int nkeys = T + 1;
  for (int c = lane; c < PPW * nkeys * 8; c += 32) {
    const int ch = c & 7;
    const int slot = (c >> 3) % nkeys;
    const int sp = (c >> 3) / nkeys;
    if (h0 + sp < H) {
      const size_t tokrow = (slot == 0) ? 0 : (1 + static_cast<size_t>(slot - 1) * n + p);
      const bf16* src = clip + tokrow * ld + (h0 + sp) * HD + ch * 8;
      const uint4 ku = *reinterpret_cast<const uint4*>(src + D);
      const uint4 vu = *reinterpret_cast<const uint4*>(src + 2 * D);
      float kf[8], vf[8];
      unpack8(ku, kf);
      unpack8(vu, vf);
      float* kd = Ks + (sp * NK + slot) * HD + ch * 8;
      float* vd = Vs + (sp * NK + slot) * HD + ch * 8;
      *reinterpret_cast<float4*>(kd) = make_float4(kf[0], kf[1], kf[2], kf[3]);
      *reinterpret_cast<float4*>(kd + 4) = make_float4(kf[4], kf[5], kf[6], kf[7]);
      *reinterpret_cast<float4*>(vd) = make_float4(vf[0], vf[1], vf[2], vf[3]);
      *reinterpret_cast<float4*>(vd + 4) = make_float4(vf[4], vf[5], vf[6], vf[7]);
    }
  }
  __syncwarp();

  const int sp = lane / TP, f = lane % TP;
  const bool active = (f < T) && (h0 + sp < H);
  if (!active) return;
  const size_t qrow = 1 + static_cast<size_t>(f) * n + p;
  const bf16* qsrc = clip + qrow * ld + (h0 + sp) * HD;
  float q[HD];
#pragma unroll
  for (int c = 0; c < 8; ++c) unpack8(*reinterpret_cast<const uint4*>(qsrc + c * 8), q + c * 8);

  const float* Kp = Ks + sp * NK * HD;
  const float* Vp = Vs + sp * NK * HD;
  float s[NK];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < NK; ++j) {
    if (j < nkeys) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 k4 = *reinterpret_cast<const float4*>(Kp + j * HD + d);
        acc += q[d] * k4.x + q[d + 1] * k4.y + q[d + 2] * k4.z + q[d + 3] * k4.w;
      }
      s[j] = acc;
      mx = fmaxf(mx, acc);
    } else {
      s[j] = -INFINITY;
    }
  }
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < NK; ++j) {
    s[j] = exp2f((s[j] - mx) * LOG2E);  // exp2f(-inf) = 0 for unused slots
    l += s[j];
  }
  const float inv = 1.f / l;
  // reuse q[] as the output accumulator
#pragma unroll
  for (int d = 0; d < HD; ++d) q[d] = 0.f;
#pragma unroll
  for (int j = 0; j < NK; ++j) {
    if (j < nkeys) {
      const float pj = s[j] * inv;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 v4 = *reinterpret_cast<const float4*>(Vp + j * HD + d);
        q[d] += pj * v4.x; q[d + 1] += pj * v4.y; q[d + 2] += pj * v4.z; q[d + 3] += pj * v4.w;
      }
    }
  }
  bf16* dst = out + (static_cast<size_t>(b) * N + qrow) * D + (h0 + sp) * HD;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 u;
    u.x = pack_bf16x2(q[c * 8 + 0], q[c * 8 + 1]);
    u.y = pack_bf16x2(q[c * 8 + 2], q[c * 8 + 3]);
    u.z = pack_bf16x2(q[c * 8 + 4], q[c * 8 + 5]);
    u.w = pack_bf16x2(q[c * 8 + 6], q[c * 8 + 7]);
    *reinterpret_cast<uint4*>(dst + c * 8) = u;
  }
}

template <int TP>
int launch_time(const bf16* qkv, bf16* out, int B, int T, int n, int H, cudaStream_t stream) {
  constexpr int PPW = 32 / TP;
  const size_t smem = static_cast<size_t>(TIME_WARPS) * 2 * PPW * (TP + 1) * HD * sizeof(float);
  static bool configured = false;
  if (!configured) {
    HH_CHECK_CUDA(cudaFuncSetAttribute(attn_time_kernel<TP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    configured = true;
  }
  const long long ntasks = static_cast<long long>(B) * n * ((H + PPW - 1) / PPW);
  const long long blocks = (ntasks + TIME_WARPS - 1) / TIME_WARPS;
  HH_REQUIRE(blocks < (1ll << 31), "attn_time: grid too large");
  attn_time_kernel<TP><<<static_cast<unsigned>(blocks), TIME_WARPS * 32, smem, stream>>>(qkv, out, B, T, n, H);
  HH_CHECK_LAUNCH("attn_time_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------ CLS query row
constexpr int CLS_WARPS = 8;

__global__ void __launch_bounds__(CLS_WARPS * 32)
attn_cls_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int N, int H) {
  __shared__ float sm_m[CLS_WARPS * 4];
  __shared__ float sm_l[CLS_WARPS * 4];
  __shared__ float sm_o[CLS_WARPS * 4][HD];
  const int D = H * HD;
  const int h = blockIdx.x % H, b = blockIdx.x / H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, sub = lane & 7;  // 4 key groups per warp, 8 lanes x 8 dims per key
  const size_t ld = static_cast<size_t>(3) * D;
  const bf16* clip = qkv + static_cast<size_t>(b) * N * ld + h * HD + sub * 8;

  float q[8];
  unpack8(*reinterpret_cast<const uint4*>(clip), q);  // CLS token is row 0; q section is columns [0, D)
  float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) o[d] = 0.f;

  const int step = CLS_WARPS * 4;
  for (int j0 = 0; j0 < N; j0 += step) {  // uniform trip count: the shuffles below need the whole warp
    const int j = j0 + warp * 4 + grp;
    const bool valid = j < N;
    float s = 0.f, v[8];
    if (valid) {
      const bf16* row = clip + static_cast<size_t>(j) * ld;
      float k[8];
      unpack8(*reinterpret_cast<const uint4*>(row + D), k);
      unpack8(*reinterpret_cast<const uint4*>(row + 2 * D), v);
#pragma unroll
      for (int d = 0; d < 8; ++d) s += q[d] * k[d];
    } else {
#pragma unroll
      for (int d = 0; d < 8; ++d) v[d] = 0.f;
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (valid) {
      const float mn = fmaxf(m, s);
      const float c = exp2f((m - mn) * LOG2E);  // m = -inf on first key -> 0
      const float pj = exp2f((s - mn) * LOG2E);
      l = l * c + pj;
#pragma unroll
      for (int d = 0; d < 8; ++d) o[d] = o[d] * c + pj * v[d];
      m = mn;
    }
  }
  const int slot = warp * 4 + grp;
  if (sub == 0) {
    sm_m[slot] = m;
    sm_l[slot] = l;
  }
#pragma unroll
  for (int d = 0; d < 8; ++d) sm_o[slot][sub * 8 + d] = o[d];
  __syncthreads();
  if (threadIdx.x < HD) {
    float mm = -INFINITY;
    for (int s2 = 0; s2 < CLS_WARPS * 4; ++s2) mm = fmaxf(mm, sm_m[s2]);
    float ll = 0.f, acc = 0.f;
    for (int s2 = 0; s2 < CLS_WARPS * 4; ++s2) {
      const float c = (sm_m[s2] == -INFINITY) ? 0.f : exp2f((sm_m[s2] - mm) * LOG2E);
      ll += sm_l[s2] * c;
      acc += sm_o[s2][threadIdx.x] * c;
    }
    out[static_cast<size_t>(b) * N * D + h * HD + threadIdx.x] = __float2bfloat16(acc / ll);
  }
}

}  // namespace

int attn_time(const bf16* qkv, bf16* out, int B, int T, int n, int H, cudaStream_t stream) {
  HH_REQUIRE(B > 0 && T > 0 && n > 0 && H > 0, "attn_time: empty problem");
  HH_REQUIRE(T <= 32, "attn_time: at most 32 frames");
  if (T <= 4) return launch_time<4>(qkv, out, B, T, n, H, stream);
  if (T <= 8) return launch_time<8>(qkv, out, B, T, n, H, stream);
  if (T <= 16) return launch_time<16>(qkv, out, B, T, n, H, stream);
  return launch_time<32>(qkv, out, B, T, n, H, stream);
}

int attn_cls(const bf16* qkv, bf16* out, int B, int N, int H, cudaStream_t stream) {
  HH_REQUIRE(B > 0 && N > 0 && H > 0, "attn_cls: empty problem");
  attn_cls_kernel<<<B * H, CLS_WARPS * 32, 0, stream>>>(qkv, out, N, H);
  HH_CHECK_LAUNCH("attn_cls_kernel");
  return 0;
}

}  // namespace hh
