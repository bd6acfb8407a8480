// Spatial half of the divided space-time attention (VarAttention with einops '(b f) n d', model/LaviLa.py:246-283):
// patch query (f,p) attends {CLS key} U {keys of frame f}.  One CTA per (clip, frame, head): the frame's K and V
// (n x 64 bf16 each) and Q stay resident in shared memory, 8 warps run a flash-style loop with legacy warp MMA
// (mma.sync m16n8k16, fp32 accumulate, fp32 online softmax in the exp2 domain).  The CLS key/value initialise the
// running softmax state (m = q.k_cls, l = 1, O = v_cls), so the key loop only sees the n patch keys.
// The CLS *query* (which attends all 1+T*n keys, LaviLa.py:258) rides along as one extra 16-row block per CTA whose
// only live row is q_cls: its (max, sum, o[64]) over this frame's keys is written as a partial per (clip, head, frame)
// and folded, with the CLS key itself, into output row 0 by attn_cls_merge (attn_time.cu).
#include <cstdlib>
#include "hh_internal.h"
#include "hh_ptx.cuh"

namespace hh {

namespace {

constexpr int HD = 64;
constexpr int LDS = 72;  // smem row stride in bf16 (144 B): 16-byte aligned rows, conflict-free ldmatrix
constexpr float LOG2E = 1.4426950408889634f;

__global__ void __launch_bounds__(256, 2)
attn_space_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ cls_part, int T, int n, int H) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int D = H * HD;
  const int N = 1 + T * n;
  const int h = blockIdx.x % H;
  const int f = (blockIdx.x / H) % T;
  const int b = blockIdx.x / (H * T);
  const int mblocks = (n + 15) >> 4;
  const int kblocks = (n + 63) >> 6;
  const int qrows = mblocks * 16;
  const int krows = kblocks * 64;

  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);          // qrows patch queries + 16 rows for the CLS-query block
  bf16* Ks = Qs + (qrows + 16) * LDS;
  bf16* Vs = Ks + krows * LDS;
  float* kcls = reinterpret_cast<float*>(Vs + krows * LDS);
  float* vcls = kcls + HD;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const size_t ld = static_cast<size_t>(3) * D;
  const bf16* base = qkv + (static_cast<size_t>(b) * N + 1 + static_cast<size_t>(f) * n) * ld + h * HD;

  // ---- stage Q, K, V (zero-filled past n) with 16-byte async copies
  const int maxrows = krows > qrows ? krows : qrows;
  for (int c = tid; c < maxrows * 8; c += blockDim.x) {
    const int r = c >> 3, ch = c & 7;
    const bool valid = r < n;
    const bf16* src = base + static_cast<size_t>(valid ? r : 0) * ld + ch * 8;
    if (r < qrows) cp_async_16(Qs + r * LDS + ch * 8, src, valid);
    if (r < krows) {
      cp_async_16(Ks + r * LDS + ch * 8, src + D, valid);
      cp_async_16(Vs + r * LDS + ch * 8, src + 2 * D, valid);
    }
  }
  {  // CLS-query block: row qrows = q_cls, rows qrows+1 .. qrows+15 = 0
    const bf16* cls = qkv + static_cast<size_t>(b) * N * ld + h * HD;
    if (tid < 16 * 8) {
      const int r = tid >> 3, ch = tid & 7;
      cp_async_16(Qs + (qrows + r) * LDS + ch * 8, cls + ch * 8, r == 0);
    }
    cp_async_commit();
    if (tid < 2 * HD) {
      if (tid < HD) kcls[tid] = __bfloat162float(cls[D + tid]);
      else vcls[tid - HD] = __bfloat162float(cls[2 * D + tid - HD]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  const int mi = lane >> 3, lr = lane & 7;

  for (int mb = warp; mb <= mblocks; mb += 8) {  // block `mblocks` is the CLS query
    const bool is_cls = (mb == mblocks);
    const int r0 = mb * 16;
    // Q fragments for the 4 k-steps over d
    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ldmatrix_x4(qf[ks], smem_u32(Qs + (r0 + (mi & 1) * 8 + lr) * LDS + ks * 16 + (mi >> 1) * 8));

    // ---- CLS key initialises the online softmax
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int c = ks * 16 + 2 * t;
      float2 a0 = unpack_bf16x2(qf[ks][0]), a1 = unpack_bf16x2(qf[ks][1]);
      float2 a2 = unpack_bf16x2(qf[ks][2]), a3 = unpack_bf16x2(qf[ks][3]);
      s0 += a0.x * kcls[c] + a0.y * kcls[c + 1] + a2.x * kcls[c + 8] + a2.y * kcls[c + 9];
      s1 += a1.x * kcls[c] + a1.y * kcls[c + 1] + a3.x * kcls[c + 8] + a3.y * kcls[c + 9];
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    float m0 = s0, m1 = s1;
    float l0 = (t == 0) ? 1.f : 0.f, l1 = l0;  // thread-partial row sums, reduced over the quad at the end
    float o[8][4];
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
      o[ni][0] = o[ni][2] = vcls[ni * 8 + 2 * t];
      o[ni][1] = o[ni][3] = vcls[ni * 8 + 2 * t + 1];
    }
    if (is_cls) {  // the CLS query's partial covers this frame's patch keys only: empty initial state
      m0 = m1 = -INFINITY;
      l0 = l1 = 0.f;
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) o[ni][0] = o[ni][1] = o[ni][2] = o[ni][3] = 0.f;
    }

    for (int kb = 0; kb < kblocks; ++kb) {
      float s[8][4];
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) s[ni][0] = s[ni][1] = s[ni][2] = s[ni][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t kf[4];
          ldmatrix_x4(kf, smem_u32(Ks + (kb * 64 + np * 16 + (mi >> 1) * 8 + lr) * LDS + ks * 16 + (mi & 1) * 8));
          mma_bf16_16816(s[2 * np], qf[ks], kf[0], kf[1]);
          mma_bf16_16816(s[2 * np + 1], qf[ks], kf[2], kf[3]);
        }
      }
      // mask keys past n (only possible in the last block), row max
      float mx0 = -INFINITY, mx1 = -INFINITY;
      if (kb * 64 + 64 > n) {  // ragged last block only (n = 196); n = 256 never takes this path
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
          const int key = kb * 64 + ni * 8 + 2 * t;
          if (key >= n) s[ni][0] = s[ni][2] = -INFINITY;
          if (key + 1 >= n) s[ni][1] = s[ni][3] = -INFINITY;
        }
      }
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) {
        mx0 = fmaxf(mx0, fmaxf(s[ni][0], s[ni][1]));
        mx1 = fmaxf(mx1, fmaxf(s[ni][2], s[ni][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);  // finite: m0/m1 start finite
      const float ml0 = mn0 * LOG2E, ml1 = mn1 * LOG2E;
      const float c0 = fast_exp2(m0 * LOG2E - ml0), c1 = fast_exp2(m1 * LOG2E - ml1);
      m0 = mn0;
      m1 = mn1;
      l0 *= c0;
      l1 *= c1;
      uint32_t pa[4][4];
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) {
        const float p0 = fast_exp2(fmaf(s[ni][0], LOG2E, -ml0)), p1 = fast_exp2(fmaf(s[ni][1], LOG2E, -ml0));
        const float p2 = fast_exp2(fmaf(s[ni][2], LOG2E, -ml1)), p3 = fast_exp2(fmaf(s[ni][3], LOG2E, -ml1));
        l0 += p0 + p1;
        l1 += p2 + p3;
        pa[ni >> 1][(ni & 1) * 2 + 0] = pack_bf16x2(p0, p1);
        pa[ni >> 1][(ni & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        o[ni][0] *= c0; o[ni][1] *= c0; o[ni][2] *= c1; o[ni][3] *= c1;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t vf[4];
          ldmatrix_x4_trans(vf, smem_u32(Vs + (kb * 64 + kk * 16 + (mi & 1) * 8 + lr) * LDS + dp * 16 + (mi >> 1) * 8));
          mma_bf16_16816(o[2 * dp], pa[kk], vf[0], vf[1]);
          mma_bf16_16816(o[2 * dp + 1], pa[kk], vf[2], vf[3]);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    if (is_cls) {  // unnormalised partial (m, l, o[64]) of row 0 -> workspace [b][h][f]
      if (g == 0) {
        float* dst = cls_part + ((static_cast<size_t>(b) * H + h) * T + f) * (HD + 2);
        if (t == 0) {
          dst[0] = m0;
          dst[1] = l0;
        }
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
          dst[2 + ni * 8 + 2 * t] = o[ni][0];
          dst[2 + ni * 8 + 2 * t + 1] = o[ni][1];
        }
      }
      continue;
    }
    const float i0 = 1.f / l0, i1 = 1.f / l1;

    // ---- stage the 16x64 result in this warp's (now free) Q rows, then 16-byte coalesced row stores
    __syncwarp();
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
      *reinterpret_cast<uint32_t*>(Qs + (r0 + g) * LDS + ni * 8 + 2 * t) = pack_bf16x2(o[ni][0] * i0, o[ni][1] * i0);
      *reinterpret_cast<uint32_t*>(Qs + (r0 + g + 8) * LDS + ni * 8 + 2 * t) = pack_bf16x2(o[ni][2] * i1, o[ni][3] * i1);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int c = it * 32 + lane;
      const int r = c >> 3, ch = c & 7;
      if (r0 + r < n) {
        const uint4 v = *reinterpret_cast<const uint4*>(Qs + (r0 + r) * LDS + ch * 8);
        bf16* dst = out + (static_cast<size_t>(b) * N + 1 + static_cast<size_t>(f) * n + r0 + r) * D + h * HD + ch * 8;
        *reinterpret_cast<uint4*>(dst) = v;
      }
    }
  }
}

}  // namespace

int attn_space(const bf16* qkv, bf16* out, int B, int T, int n, int H, float* cls_ws, cudaStream_t stream) {
  HH_REQUIRE(B > 0 && T > 0 && n > 0 && H > 0, "attn_space: empty problem");
  HH_REQUIRE(cls_ws != nullptr, "attn_space: CLS workspace");
  if (attn_space_tc_supported(n)) return attn_space_tc(qkv, out, B, T, n, H, cls_ws, stream);  // tcgen05 path
  const int mblocks = (n + 15) / 16, kblocks = (n + 63) / 64;
  const size_t smem = static_cast<size_t>(mblocks * 16 + 16 + 2 * kblocks * 64) * LDS * sizeof(bf16) + 2 * HD * sizeof(float);
  HH_REQUIRE(smem <= 227 * 1024, "attn_space: patches per frame too large for the resident-K/V kernel");
  static size_t configured = 0;
  if (smem > configured) {
    HH_CHECK_CUDA(cudaFuncSetAttribute(attn_space_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    configured = smem;
  }
  attn_space_kernel<<<B * T * H, 256, smem, stream>>>(qkv, out, cls_ws, T, n, H);
  HH_CHECK_LAUNCH("attn_space_kernel");
  return attn_cls_merge(qkv, cls_ws, out, B, 1 + T * n, H, T, stream);
}

}  // namespace hh
