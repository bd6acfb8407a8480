// Temporal half of the divided space-time attention (VarAttention with einops '(b n) f d', model/LaviLa.py:246-283)
// and the CLS query row shared by both halves.
//
// attn_time (T <= 16): patch query (f,p) attends {CLS key} U {(f',p) : f'} -- T+1 <= 17 keys of 64 dims, 16 queries.
//   HBM-bound (q,k,v read once, o written once).  One warp owns one head of one clip and walks a chunk of patch
//   positions; per position it pulls the 3*T row segments (128 B each) into its private shared-memory tile with
//   16-byte cp.async, then runs the 16x24x64 and 16x64x32 products on the legacy tensor-core path (mma.sync
//   m16n8k16, fp32 accumulate) with an fp32 softmax in registers.  The 8 warps of a CTA cover 8 adjacent heads, so
//   every global row access of the CTA is a 1 KB contiguous span.
//   The CLS *query* (which attends all 1+T*n keys, LaviLa.py:258) rides along as a second 16-row MMA block whose
//   only live row is q_cls: each warp keeps a running (max, sum, o[64]) over the keys it sees and writes ONE partial
//   per (clip, head, patch-chunk); attn_cls_merge folds the partials and the CLS key itself into output row 0.
// attn_time (16 < T <= 32) keeps the simple SIMT kernel + the stand-alone CLS kernel.
#include <cstdlib>

#include "hh_internal.h"
#include "hh_ptx.cuh"

namespace hh {

namespace {

constexpr int HD = 64;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// ============================================================================================ tensor-core kernel
constexpr int TW = 8;            // warps (= adjacent heads) per CTA
constexpr int TLD = 72;          // smem row stride in bf16 (144 B): conflict-free ldmatrix
constexpr int TQ_ROWS = 32;      // rows 0..15 queries of the patch problems, row 16 = q_cls, 17..31 zero
constexpr int TK_ROWS = 24;      // key slots: 0 = CLS, 1..T frames, rest zero
constexpr int TV_ROWS = 32;      // value slots (second k16 step reads rows 16..31)
constexpr int TWARP_ELEMS = (TQ_ROWS + TK_ROWS + TV_ROWS) * TLD;

__global__ void __launch_bounds__(TW * 32, 2)
attn_time_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ cls_part, int T, int n,
                     int H, int pchunk, int nchunks) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw) + static_cast<size_t>(warp) * TWARP_ELEMS;
  bf16* Ks = Qs + TQ_ROWS * TLD;
  bf16* Vs = Ks + TK_ROWS * TLD;

  const int D = H * HD;
  const int N = 1 + T * n;
  const int hgroups = (H + TW - 1) / TW;
  const int chunk = blockIdx.x % nchunks;
  const int hg = (blockIdx.x / nchunks) % hgroups;
  const int b = blockIdx.x / (nchunks * hgroups);
  const int h = hg * TW + warp;
  if (h >= H) return;  // whole warp; no block-level barriers below
  const size_t ld = static_cast<size_t>(3) * D;
  const bf16* clip = qkv + static_cast<size_t>(b) * N * ld + h * HD;

  // ---- zero the private tile once (padding rows must stay finite), then the per-warp constants: CLS k, v, q
  for (int i = lane; i < TWARP_ELEMS / 8; i += 32) reinterpret_cast<uint4*>(Qs)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  if (lane < 24) {
    const int which = lane >> 3, ch = lane & 7;  // 0: q_cls -> Qs row 16, 1: k_cls -> Ks row 0, 2: v_cls -> Vs row 0
    bf16* dst = (which == 0 ? Qs + 16 * TLD : which == 1 ? Ks : Vs) + ch * 8;
    cp_async_16(dst, clip + which * D + ch * 8, true);
  }
  cp_async_commit();

  const int g = lane >> 2, t = lane & 3;
  const int mi = lane >> 3, lr = lane & 7;
  const int nkeys = T + 1;

  // running softmax state of the CLS query over the keys this warp visits (rows g of the second MMA block;
  // only g == 0 is meaningful)
  float cm = -INFINITY, cl = 0.f;
  float co[8][4];
#pragma unroll
  for (int ni = 0; ni < 8; ++ni) co[ni][0] = co[ni][1] = co[ni][2] = co[ni][3] = 0.f;

  const int p_begin = chunk * pchunk;
  const int p_end = min(n, p_begin + pchunk);
  for (int p = p_begin; p < p_end; ++p) {
    // ---- stage q (rows 0..T-1), k, v (slots 1..T) of patch position p: 3*T rows x 8 chunks of 16 bytes
    for (int c = lane; c < 3 * T * 8; c += 32) {
      const int ch = c & 7;
      const int r = c >> 3;
      const int which = r / T, f = r - which * T;
      const bf16* src = clip + (1 + static_cast<size_t>(f) * n + p) * ld + which * D + ch * 8;
      bf16* dst = (which == 0 ? Qs + f * TLD : which == 1 ? Ks + (1 + f) * TLD : Vs + (1 + f) * TLD) + ch * 8;
      cp_async_16(dst, src, true);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();

    // ---- S = Q K^T for both 16-row blocks (block 1: only row 16 = q_cls is live)
    float s[2][3][4];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
      for (int ni = 0; ni < 3; ++ni) s[mb][ni][0] = s[mb][ni][1] = s[mb][ni][2] = s[mb][ni][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t k01[4], k2[4];
      ldmatrix_x4(k01, smem_u32(Ks + ((mi >> 1) * 8 + lr) * TLD + ks * 16 + (mi & 1) * 8));
      ldmatrix_x4(k2, smem_u32(Ks + (16 + lr) * TLD + ks * 16 + (mi & 1) * 8));  // matrices 2,3 repeat 0,1 (unused)
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        uint32_t qf[4];
        ldmatrix_x4(qf, smem_u32(Qs + (mb * 16 + (mi & 1) * 8 + lr) * TLD + ks * 16 + (mi >> 1) * 8));
        mma_bf16_16816(s[mb][0], qf, k01[0], k01[1]);
        mma_bf16_16816(s[mb][1], qf, k01[2], k01[3]);
        mma_bf16_16816(s[mb][2], qf, k2[0], k2[1]);
      }
    }

    // ---- patch queries: softmax over the nkeys valid slots, O = P V
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int ni = 0; ni < 3; ++ni) {
      const int key = ni * 8 + 2 * t;
      if (key >= nkeys) s[0][ni][0] = s[0][ni][2] = -INFINITY;
      if (key + 1 >= nkeys) s[0][ni][1] = s[0][ni][3] = -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(s[0][ni][0], s[0][ni][1]));
      mx1 = fmaxf(mx1, fmaxf(s[0][ni][2], s[0][ni][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float l0 = 0.f, l1 = 0.f;
    uint32_t pa[2][4];
#pragma unroll
    for (int ni = 0; ni < 3; ++ni) {
      const float p0 = fast_exp2(fmaf(s[0][ni][0], LOG2E, -mx0 * LOG2E)), p1 = fast_exp2(fmaf(s[0][ni][1], LOG2E, -mx0 * LOG2E));
      const float p2 = fast_exp2(fmaf(s[0][ni][2], LOG2E, -mx1 * LOG2E)), p3 = fast_exp2(fmaf(s[0][ni][3], LOG2E, -mx1 * LOG2E));
      l0 += p0 + p1;
      l1 += p2 + p3;
      pa[ni >> 1][(ni & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pa[ni >> 1][(ni & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    pa[1][2] = pa[1][3] = 0u;
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

    // ---- CLS query (row g of block 1): keys 1..T of this position join its running softmax; key 0 (the CLS key
    //      itself) is added once, by the merge kernel
    float cmx = -INFINITY;
#pragma unroll
    for (int ni = 0; ni < 3; ++ni) {
      const int key = ni * 8 + 2 * t;
      if (key >= nkeys || key == 0) s[1][ni][0] = -INFINITY;
      if (key + 1 >= nkeys) s[1][ni][1] = -INFINITY;
      cmx = fmaxf(cmx, fmaxf(s[1][ni][0], s[1][ni][1]));
    }
    cmx = fmaxf(cmx, __shfl_xor_sync(0xffffffffu, cmx, 1));
    cmx = fmaxf(cmx, __shfl_xor_sync(0xffffffffu, cmx, 2));
    const float cmn = fmaxf(cm, cmx);  // finite: T >= 1 gives at least one valid key
    const float ccorr = fast_exp2((cm - cmn) * LOG2E);
    cm = cmn;
    cl *= ccorr;
    uint32_t pc[2][4];
#pragma unroll
    for (int ni = 0; ni < 3; ++ni) {
      const float p0 = fast_exp2(fmaf(s[1][ni][0], LOG2E, -cmn * LOG2E)), p1 = fast_exp2(fmaf(s[1][ni][1], LOG2E, -cmn * LOG2E));
      cl += p0 + p1;
      pc[ni >> 1][(ni & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pc[ni >> 1][(ni & 1) * 2 + 1] = 0u;  // rows g+8 of block 1 are dead
    }
    pc[1][2] = pc[1][3] = 0u;
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
      co[ni][0] *= ccorr;
      co[ni][1] *= ccorr;
    }

    float o[8][4];
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) o[ni][0] = o[ni][1] = o[ni][2] = o[ni][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, smem_u32(Vs + (kk * 16 + (mi & 1) * 8 + lr) * TLD + dp * 16 + (mi >> 1) * 8));
        mma_bf16_16816(o[2 * dp], pa[kk], vf[0], vf[1]);
        mma_bf16_16816(o[2 * dp + 1], pa[kk], vf[2], vf[3]);
        mma_bf16_16816(co[2 * dp], pc[kk], vf[0], vf[1]);
        mma_bf16_16816(co[2 * dp + 1], pc[kk], vf[2], vf[3]);
      }
    }

    // ---- stage the 16x64 result over the (consumed) query rows, then 16-byte coalesced row stores
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    __syncwarp();
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
      *reinterpret_cast<uint32_t*>(Qs + g * TLD + ni * 8 + 2 * t) = pack_bf16x2(o[ni][0] * i0, o[ni][1] * i0);
      *reinterpret_cast<uint32_t*>(Qs + (g + 8) * TLD + ni * 8 + 2 * t) = pack_bf16x2(o[ni][2] * i1, o[ni][3] * i1);
    }
    __syncwarp();
    for (int c = lane; c < T * 8; c += 32) {
      const int f = c >> 3, ch = c & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(Qs + f * TLD + ch * 8);
      bf16* dst = out + (static_cast<size_t>(b) * N + 1 + static_cast<size_t>(f) * n + p) * D + h * HD + ch * 8;
      *reinterpret_cast<uint4*>(dst) = v;
    }
    __syncwarp();  // Qs rows are rewritten by the next position's loads
  }

  // ---- this warp's CLS partial: (m, l, o[64]) for (b, h, chunk); row 0 of block 1 lives in lanes 0..3
  cl += __shfl_xor_sync(0xffffffffu, cl, 1);
  cl += __shfl_xor_sync(0xffffffffu, cl, 2);
  if (g == 0) {
    float* dst = cls_part + ((static_cast<size_t>(b) * H + h) * nchunks + chunk) * (HD + 2);
    if (t == 0) {
      dst[0] = cm;
      dst[1] = cl;
    }
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
      dst[2 + ni * 8 + 2 * t] = co[ni][0];
      dst[2 + ni * 8 + 2 * t + 1] = co[ni][1];
    }
  }
}

// Folds the per-chunk partials of the CLS query and the CLS key/value itself into output row 0 of every (clip, head).
__global__ void __launch_bounds__(HD)
attn_cls_merge_kernel(const bf16* __restrict__ qkv, const float* __restrict__ part, bf16* __restrict__ out, int N, int H,
                      int nparts) {
  __shared__ float red[2];
  const int h = blockIdx.x % H, b = blockIdx.x / H;
  const int d = threadIdx.x;
  const int D = H * HD;
  const bf16* cls = qkv + static_cast<size_t>(b) * N * 3 * D + h * HD;
  // s_cls = q_cls . k_cls (q pre-scaled)
  float prod = __bfloat162float(cls[d]) * __bfloat162float(cls[D + d]);
  prod = warp_sum(prod);
  if ((d & 31) == 0) red[d >> 5] = prod;
  __syncthreads();
  const float s_cls = red[0] + red[1];
  const float* base = part + static_cast<size_t>(blockIdx.x) * nparts * (HD + 2);
  float mm = s_cls;
  for (int i = 0; i < nparts; ++i) mm = fmaxf(mm, base[i * (HD + 2)]);
  float ll = exp2f((s_cls - mm) * LOG2E);
  float oo = ll * __bfloat162float(cls[2 * D + d]);
  for (int i = 0; i < nparts; ++i) {
    const float* pr = base + i * (HD + 2);
    const float c = (pr[0] == -INFINITY) ? 0.f : exp2f((pr[0] - mm) * LOG2E);
    ll += pr[1] * c;
    oo += pr[2 + d] * c;
  }
  out[static_cast<size_t>(b) * N * D + h * HD + d] = __float2bfloat16(oo / ll);
}

// ============================================================================================ SIMT kernel (T > 16)
constexpr int TIME_WARPS = 4;

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
    s[j] = exp2f((s[j] - mx) * LOG2E);
    l += s[j];
  }
  const float inv = 1.f / l;
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

// ------------------------------------------------------------------------------------------ stand-alone CLS query row
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

size_t attn_cls_workspace_bytes(int B, int T, int n, int H) {
  const int parts = (T > (n + 15) / 16) ? T : (n + 15) / 16;
  return static_cast<size_t>(B) * H * parts * (HD + 2) * sizeof(float);
}

int attn_cls_merge(const bf16* qkv, const float* parts, bf16* out, int B, int N, int H, int nparts, cudaStream_t stream) {
  attn_cls_merge_kernel<<<B * H, HD, 0, stream>>>(qkv, parts, out, N, H, nparts);
  HH_CHECK_LAUNCH("attn_cls_merge_kernel");
  return 0;
}

int attn_time(const bf16* qkv, bf16* out, int B, int T, int n, int H, float* cls_ws, cudaStream_t stream) {
  HH_REQUIRE(B > 0 && T > 0 && n > 0 && H > 0, "attn_time: empty problem");
  HH_REQUIRE(T <= 32, "attn_time: at most 32 frames");
  if (T <= 16) {
    HH_REQUIRE(cls_ws != nullptr, "attn_time: CLS workspace");
    const int pchunk = 16;
    const int nchunks = (n + pchunk - 1) / pchunk;
    static const bool use_v1 = std::getenv("HH_ATTN_TIME_V1") != nullptr;
    if (!use_v1) {  // double-buffered, swizzled-tile generation of the same kernel (attn_time_v2.cu)
      int rc = attn_time_v2(qkv, out, B, T, n, H, cls_ws, pchunk, nchunks, stream);
      if (rc) return rc;
      return attn_cls_merge(qkv, cls_ws, out, B, 1 + T * n, H, nchunks, stream);
    }
    const size_t smem = static_cast<size_t>(TW) * TWARP_ELEMS * sizeof(bf16);
    static bool configured = false;
    if (!configured) {
      HH_CHECK_CUDA(cudaFuncSetAttribute(attn_time_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
      configured = true;
    }
    const long long blocks = static_cast<long long>(B) * ((H + TW - 1) / TW) * nchunks;
    HH_REQUIRE(blocks < (1ll << 31), "attn_time: grid too large");
    attn_time_mma_kernel<<<static_cast<unsigned>(blocks), TW * 32, smem, stream>>>(qkv, out, cls_ws, T, n, H, pchunk,
                                                                                   nchunks);
    HH_CHECK_LAUNCH("attn_time_mma_kernel");
    return attn_cls_merge(qkv, cls_ws, out, B, 1 + T * n, H, nchunks, stream);
  }
  int rc = launch_time<32>(qkv, out, B, T, n, H, stream);
  if (rc) return rc;
  return attn_cls(qkv, out, B, 1 + T * n, H, stream);
}

int attn_cls(const bf16* qkv, bf16* out, int B, int N, int H, cudaStream_t stream) {
  HH_REQUIRE(B > 0 && N > 0 && H > 0, "attn_cls: empty problem");
  attn_cls_kernel<<<B * H, CLS_WARPS * 32, 0, stream>>>(qkv, out, N, H);
  HH_CHECK_LAUNCH("attn_cls_kernel");
  return 0;
}

}  // namespace hh
