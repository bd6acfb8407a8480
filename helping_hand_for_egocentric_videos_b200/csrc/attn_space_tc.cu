// Spatial attention on the 5th-generation tensor cores (tcgen05 + TMEM), n <= 256 patches per frame, head dim 64.
// Same operator as attn_space.cu (VarAttention '(b f) n d', model/LaviLa.py:246-283); that file stays as the
// mma.sync statement of it (fallback for n > 256 and differential reference).
//
// One task = one (clip, frame, head): Q, K, V tiles of [n x 64] bf16.  All n keys fit one accumulator tile, so the
// softmax is a single exact pass (no online rescaling):
//     S[256 x 256] = Q K^T     2 halves of 128 rows: tcgen05.mma 128x256x16, A = Q (smem), B = K (smem), D -> TMEM
//     P = exp2(S - max)        one thread per row reads its TMEM lane, writes P back as bf16 over the S columns
//     O[256 x 64]  = P V       tcgen05.mma 128x64x16, A = P (TMEM), B = V (smem, N-major), D -> TMEM
// plus the CLS key/value (one extra logit per row, folded in registers).
//
// Persistent CTA, 12 warps:  0 TMA producer (4-D tensor maps: rows past n are zero-filled / clipped by hardware, two
// tasks in flight)  |  1 MMA issuer  |  2-3 CLS *query* partial over this frame's keys (SIMT; merged across frames by
// attn_cls_merge)  |  4-7 softmax + epilogue of rows 0..127  |  8-11 the same for rows 128..255.
// TMEM (512 columns): half h owns columns [256h, 256h+256): S, then P in the first 128 and O in the next 64.
#include <cstdlib>

#include "hh_internal.h"
#include "hh_ptx.cuh"

namespace hh {

namespace {

constexpr int HD = 64;
constexpr int ROWS = 256;                 // query / key slots per task
constexpr int TILE_BYTES = ROWS * 128;    // 32 KB: [256 rows x 64 bf16], SWIZZLE_128B
constexpr int STAGE_BYTES = 3 * TILE_BYTES;
constexpr int NSTAGE = 2;
constexpr int NTHREADS = 384;
constexpr float LOG2E = 1.4426950408889634f;
constexpr int PART = HD + 2;

struct SmemExtras {
  __nv_bfloat16 cls_q[NSTAGE][HD];   // q, k, v of the CLS token for (clip, head) of the staged task
  __nv_bfloat16 cls_k[NSTAGE][HD];
  __nv_bfloat16 cls_v[NSTAGE][HD];
  float merge[PART];                 // CLS-query partial of warp 3, folded by warp 2
  uint64_t full[NSTAGE], empty[NSTAGE];
  uint64_t s_full[2], p_full[2], o_full[2], t_free[2];
  uint32_t tmem_slot;
};
constexpr int SMEM_BYTES = 1024 + NSTAGE * STAGE_BYTES + static_cast<int>(sizeof(SmemExtras)) + 64;

struct TcArgs {
  const bf16* qkv;
  float* cls_part;
  int B, T, n, H;
};

__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
attn_space_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                     const TcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  SmemExtras* ex = reinterpret_cast<SmemExtras*>(smem + NSTAGE * STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = p.H * HD;
  const int N = 1 + p.T * p.n;
  const int ntasks = p.B * p.T * p.H;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_in);
    tma_prefetch_desc(&tm_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&ex->full[s], 1);
      mbar_init(&ex->empty[s], 10);  // MMA commit + 8 epilogue warps (store has read its smem) + CLS-query warps
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(&ex->s_full[h], 1);
      mbar_init(&ex->p_full[h], 4);
      mbar_init(&ex->o_full[h], 1);
      mbar_init(&ex->t_free[h], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&ex->tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ex->tmem_slot;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (lane == 0) {
      int it = 0;
      for (int task = blockIdx.x; task < ntasks; task += gridDim.x, ++it) {
        const int st = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const int h = task % p.H, f = (task / p.H) % p.T, b = task / (p.H * p.T);
        mbar_wait(&ex->empty[st], ph ^ 1u);
        uint8_t* base = smem + st * STAGE_BYTES;
        mbar_arrive_expect_tx(&ex->full[st], STAGE_BYTES + 3 * HD * 2);
        tma_load_4d(&tm_in, &ex->full[st], base, h * HD, 0, f, b);                       // Q
        tma_load_4d(&tm_in, &ex->full[st], base + TILE_BYTES, D + h * HD, 0, f, b);      // K
        tma_load_4d(&tm_in, &ex->full[st], base + 2 * TILE_BYTES, 2 * D + h * HD, 0, f, b);  // V
        const bf16* cls = p.qkv + static_cast<size_t>(b) * N * 3 * D + h * HD;
        bulk_load_1d(ex->cls_q[st], cls, HD * 2, &ex->full[st]);
        bulk_load_1d(ex->cls_k[st], cls + D, HD * 2, &ex->full[st]);
        bulk_load_1d(ex->cls_v[st], cls + 2 * D, HD * 2, &ex->full[st]);
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 256);
      constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(128, HD);
      int it = 0;
      for (int task = blockIdx.x; task < ntasks; task += gridDim.x, ++it) {
        const int st = it & 1;
        const uint32_t ph = (it >> 1) & 1;   // parity of the stage barriers
        const uint32_t tp = it & 1;          // parity of the per-task barriers
        const uint32_t qs = smem_u32(smem + st * STAGE_BYTES);
        const uint32_t ks = qs + TILE_BYTES, vs = qs + 2 * TILE_BYTES;
        mbar_wait(&ex->full[st], ph);
        tc_fence_after();
        for (int hf = 0; hf < 2; ++hf) {
          mbar_wait(&ex->t_free[hf], tp ^ 1u);   // previous task's O has left these columns
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(qs + hf * (128 * 128));
          const uint64_t db = umma_desc_sw128(ks);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_bf16(tmem_base + hf * 256, da + 2 * k, db + 2 * k, idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&ex->s_full[hf]);
        }
        for (int hf = 0; hf < 2; ++hf) {
          mbar_wait(&ex->p_full[hf], tp);
          tc_fence_after();
          const uint64_t dv = umma_desc_sw128_mn(vs);
#pragma unroll
          for (int k = 0; k < ROWS / 16; ++k)   // 16 keys per instruction = 2 swizzle atoms of V, 8 TMEM columns of P
            umma_bf16_ts(tmem_base + hf * 256 + 128, tmem_base + hf * 256 + 8 * k, dv + static_cast<uint64_t>(k * 128),
                         idesc_o, k > 0 ? 1u : 0u);
          umma_commit(&ex->o_full[hf]);
        }
        umma_commit(&ex->empty[st]);  // every MMA that read this stage's Q, K, V has retired
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ================================================================== CLS-query partial (SIMT, 64 threads)
    const int ww = warp - 2;
    int it = 0;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int h = task % p.H, f = (task / p.H) % p.T, b = task / (p.H * p.T);
      mbar_wait(&ex->full[st], ph);
      const uint32_t ks = smem_u32(smem + st * STAGE_BYTES + TILE_BYTES);
      const uint32_t vs = ks + TILE_BYTES;
      float q[HD];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 u = *reinterpret_cast<const uint4*>(&ex->cls_q[st][c * 8]);
        float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), dd = unpack_bf16x2(u.w);
        q[c * 8 + 0] = a.x; q[c * 8 + 1] = a.y; q[c * 8 + 2] = bb.x; q[c * 8 + 3] = bb.y;
        q[c * 8 + 4] = cc.x; q[c * 8 + 5] = cc.y; q[c * 8 + 6] = dd.x; q[c * 8 + 7] = dd.y;
      }
      // phase A: lane <-> key (4 keys per lane of this warp's 128-key half)
      float s[4];
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = ww * 128 + i * 32 + lane;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = ld_shared_v4(ks + j * 128 + ((c ^ (j & 7)) << 4));
          float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), dd = unpack_bf16x2(u.w);
          acc += q[c * 8 + 0] * a.x + q[c * 8 + 1] * a.y + q[c * 8 + 2] * bb.x + q[c * 8 + 3] * bb.y +
                 q[c * 8 + 4] * cc.x + q[c * 8 + 5] * cc.y + q[c * 8 + 6] * dd.x + q[c * 8 + 7] * dd.y;
        }
        s[i] = (j < p.n) ? acc : -INFINITY;
        m = fmaxf(m, s[i]);
      }
      m = warp_max(m);
      float l = 0.f;
      const float ml = (m == -INFINITY) ? 0.f : m * LOG2E;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s[i] = fast_exp2(fmaf(s[i], LOG2E, -ml));   // -inf -> 0
        l += s[i];
      }
      l = warp_sum(l);
      // phase B: lane <-> output dims (2*lane, 2*lane+1)
      float o0 = 0.f, o1 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll 8
        for (int sl = 0; sl < 32; ++sl) {
          const float pj = __shfl_sync(0xffffffffu, s[i], sl);
          const int j = ww * 128 + i * 32 + sl;
          const float2 v = unpack_bf16x2(ld_shared_u32(vs + j * 128 + (((lane >> 2) ^ (j & 7)) << 4) + (lane & 3) * 4));
          o0 = fmaf(pj, v.x, o0);
          o1 = fmaf(pj, v.y, o1);
        }
      }
      if (ww == 1) {
        if (lane == 0) {
          ex->merge[0] = m;
          ex->merge[1] = l;
        }
        ex->merge[2 + 2 * lane] = o0;
        ex->merge[2 + 2 * lane + 1] = o1;
      }
      asm volatile("bar.sync 2, 64;" ::: "memory");
      if (ww == 0) {
        const float m1 = ex->merge[0], l1 = ex->merge[1];
        const float mm = fmaxf(m, m1);
        const float c0 = (m == -INFINITY) ? 0.f : fast_exp2((m - mm) * LOG2E);
        const float c1 = (m1 == -INFINITY) ? 0.f : fast_exp2((m1 - mm) * LOG2E);
        float* dst = p.cls_part + ((static_cast<size_t>(b) * p.H + h) * p.T + f) * PART;
        if (lane == 0) {
          dst[0] = mm;
          dst[1] = l * c0 + l1 * c1;
        }
        dst[2 + 2 * lane] = o0 * c0 + ex->merge[2 + 2 * lane] * c1;
        dst[2 + 2 * lane + 1] = o1 * c0 + ex->merge[2 + 2 * lane + 1] * c1;
      }
      asm volatile("bar.sync 2, 64;" ::: "memory");  // merge[] reusable; both warps are done with K and V
      if (ww == 0 && lane == 0) mbar_arrive(&ex->empty[st]);
    }
  } else {
    // ================================================================== softmax + epilogue (one thread per row)
    const int hf = (warp - 4) >> 2;           // 0: rows 0..127, 1: rows 128..255
    const int wq = warp & 3;                  // TMEM lane quarter of this warp
    const int r = hf * 128 + wq * 32 + lane;  // row within the task
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + static_cast<uint32_t>(hf * 256);
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    int it = 0;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const uint32_t tp = it & 1;
      const int h = task % p.H, f = (task / p.H) % p.T, b = task / (p.H * p.T);
      const uint32_t qs = smem_u32(smem + st * STAGE_BYTES);
      mbar_wait(&ex->full[st], ph);

      // ---- logit of the CLS key for this row: q_r . k_cls
      float s_cls = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 u = ld_shared_v4(qs + r * 128 + ((static_cast<uint32_t>(c) ^ sw) << 4));
        const uint4 kk = *reinterpret_cast<const uint4*>(&ex->cls_k[st][c * 8]);
        const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
        const float2 k0 = unpack_bf16x2(kk.x), k1 = unpack_bf16x2(kk.y), k2 = unpack_bf16x2(kk.z), k3 = unpack_bf16x2(kk.w);
        s_cls += a0.x * k0.x + a0.y * k0.y + a1.x * k1.x + a1.y * k1.y + a2.x * k2.x + a2.y * k2.y + a3.x * k3.x +
                 a3.y * k3.y;
      }

      mbar_wait(&ex->s_full[hf], tp);
      tc_fence_after();
      const int nch = (p.n + 31) >> 5;   // 32-key chunks that hold valid keys
      // ---- pass 1: row maximum over the n patch keys and the CLS key.  TMEM loads are double-buffered (the load of
      //      chunk c+1 is in flight while chunk c is reduced); 3-input FMNMX3 halves the instruction count.
      float mx = s_cls;
      {
        uint32_t va[32], vb[32];
        auto reduce = [&](const uint32_t (&v)[32], int c) {
          if (c * 32 + 32 <= p.n) {
#pragma unroll
            for (int j = 0; j < 32; j += 2)
              asm("max.f32 %0, %0, %1, %2;" : "+f"(mx) : "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < p.n) mx = fmaxf(mx, __uint_as_float(v[j]));
          }
        };
        tmem_ld_32x32b_x32(t_lane, va);
#pragma unroll
        for (int c = 0; c < ROWS / 32; c += 2) {
          if (c < nch) {
            tmem_ld_wait();
            if (c + 1 < nch) tmem_ld_32x32b_x32(t_lane + (c + 1) * 32, vb);
            reduce(va, c);
          }
          if (c + 1 < nch) {
            tmem_ld_wait();
            if (c + 2 < nch) tmem_ld_32x32b_x32(t_lane + (c + 2) * 32, va);
            reduce(vb, c + 1);
          }
        }
      }
      const float ml = mx * LOG2E;
      const float p_cls = fast_exp2(fmaf(s_cls, LOG2E, -ml));
      // ---- pass 2: P = exp2(S - max) as bf16 pairs, written back over the S columns (16 columns per 32 keys).
      //      Packed FFMA2 / FADD2 for the scale-and-shift and the row sum; the exponentials are the MUFU floor.
      float2 l2 = make_float2(p_cls, 0.f);
      {
        uint32_t va[32], vb[32];
        const float2 sc = make_float2(LOG2E, LOG2E), sh = make_float2(-ml, -ml);
        auto emit = [&](const uint32_t (&v)[32], int c) {
          uint32_t w[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 t = __ffma2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc, sh);
            float2 e = make_float2(fast_exp2(t.x), fast_exp2(t.y));
            if (c * 32 + 32 > p.n) {  // ragged last chunk (n = 196)
              if (c * 32 + j >= p.n) e.x = 0.f;
              if (c * 32 + j + 1 >= p.n) e.y = 0.f;
            }
            l2 = __fadd2_rn(l2, e);
            w[j >> 1] = pack_bf16x2(e.x, e.y);
          }
          tmem_st_32x32b_x16(t_lane + c * 16, w);
        };
        tmem_ld_32x32b_x32(t_lane, va);
#pragma unroll
        for (int c = 0; c < ROWS / 32; c += 2) {
          if (c < nch) {
            tmem_ld_wait();
            if (c + 1 < nch) tmem_ld_32x32b_x32(t_lane + (c + 1) * 32, vb);
            emit(va, c);
          }
          if (c + 1 < nch) {
            tmem_ld_wait();
            if (c + 2 < nch) tmem_ld_32x32b_x32(t_lane + (c + 2) * 32, va);
            emit(vb, c + 1);
          }
        }
        if (nch < ROWS / 32) {  // key chunks past n: P = 0 (V rows there are zero-filled by TMA as well)
          uint32_t z[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) z[j] = 0u;
          for (int c = nch; c < ROWS / 32; ++c) tmem_st_32x32b_x16(t_lane + c * 16, z);
        }
      }
      const float l = l2.x + l2.y;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ex->p_full[hf]);

      // ---- O = P V is in TMEM columns [128, 192) of this half
      mbar_wait(&ex->o_full[hf], tp);
      tc_fence_after();
      uint32_t o[64];
      {
        uint32_t (&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&o[0]);
        uint32_t (&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&o[32]);
        tmem_ld_32x32b_x32(t_lane + 128, lo);
        tmem_ld_32x32b_x32(t_lane + 160, hi);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ex->t_free[hf]);   // the next task's S may overwrite this half

      // ---- normalise (+ CLS value), stage bf16 row into the (consumed) Q tile, bulk-store 32-row boxes
      const float inv = 1.f / l;
      const float pc = p_cls * inv;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 vv = *reinterpret_cast<const uint4*>(&ex->cls_v[st][c * 8]);
        const float2 v0 = unpack_bf16x2(vv.x), v1 = unpack_bf16x2(vv.y), v2 = unpack_bf16x2(vv.z), v3 = unpack_bf16x2(vv.w);
        const uint32_t w0 = pack_bf16x2(fmaf(__uint_as_float(o[c * 8 + 0]), inv, pc * v0.x),
                                        fmaf(__uint_as_float(o[c * 8 + 1]), inv, pc * v0.y));
        const uint32_t w1 = pack_bf16x2(fmaf(__uint_as_float(o[c * 8 + 2]), inv, pc * v1.x),
                                        fmaf(__uint_as_float(o[c * 8 + 3]), inv, pc * v1.y));
        const uint32_t w2 = pack_bf16x2(fmaf(__uint_as_float(o[c * 8 + 4]), inv, pc * v2.x),
                                        fmaf(__uint_as_float(o[c * 8 + 5]), inv, pc * v2.y));
        const uint32_t w3 = pack_bf16x2(fmaf(__uint_as_float(o[c * 8 + 6]), inv, pc * v3.x),
                                        fmaf(__uint_as_float(o[c * 8 + 7]), inv, pc * v3.y));
        st_shared_v4(qs + r * 128 + ((static_cast<uint32_t>(c) ^ sw) << 4), w0, w1, w2, w3);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const int row0 = hf * 128 + wq * 32;
        if (row0 < p.n) {
          tma_store_4d(&tm_out, qs + row0 * 128, h * HD, row0, f, b);
          tma_store_commit();
          tma_store_wait_read<0>();
        }
        mbar_arrive(&ex->empty[st]);
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// [cols, n, T, B] view of a token matrix whose rows are (clip, 1 + frame*n + patch): skips the CLS row of each clip.
int make_map4d(CUtensorMap* map, const bf16* base_row1, int cols, int n, int T, int B, int N, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(-3, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(n), static_cast<cuuint64_t>(T),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[3] = {static_cast<cuuint64_t>(cols) * 2, static_cast<cuuint64_t>(n) * cols * 2,
                        static_cast<cuuint64_t>(N) * cols * 2};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(box_rows), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base_row1), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled (4-D) failed with CUresult " + std::to_string((int)r));
  return 0;
}

}  // namespace

bool attn_space_tc_supported(int n) {
  static const bool off = std::getenv("HH_ATTN_SPACE_MMA_SYNC") != nullptr;
  return !off && n <= ROWS;
}

int attn_space_tc(const bf16* qkv, bf16* out, int B, int T, int n, int H, float* cls_ws, cudaStream_t stream) {
  HH_REQUIRE(B > 0 && T > 0 && n > 0 && n <= ROWS && H > 0, "attn_space_tc: needs 1 <= n <= 256");
  HH_REQUIRE(cls_ws != nullptr, "attn_space_tc: CLS workspace");
  HH_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "attn_space_tc: 16-byte alignment");
  const int D = H * HD, N = 1 + T * n;
  CUtensorMap tm_in, tm_out;
  int rc = make_map4d(&tm_in, qkv + static_cast<size_t>(3) * D, 3 * D, n, T, B, N, ROWS);
  if (rc) return rc;
  rc = make_map4d(&tm_out, out + D, D, n, T, B, N, 32);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    HH_CHECK_CUDA(cudaFuncSetAttribute(attn_space_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  TcArgs a;
  a.qkv = qkv;
  a.cls_part = cls_ws;
  a.B = B;
  a.T = T;
  a.n = n;
  a.H = H;
  const int ntasks = B * T * H;
  int grid = num_sms();
  if (grid > ntasks) grid = ntasks;
  attn_space_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, stream>>>(tm_in, tm_out, a);
  HH_CHECK_LAUNCH("attn_space_tc_kernel");
  return attn_cls_merge(qkv, cls_ws, out, B, N, H, T, stream);
}

}  // namespace hh
