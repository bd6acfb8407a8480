// CLIP text tower: CLIP.encode_text (model/LaviLa.py:660-670) over the
// ResidualAttentionBlock stack (model/openai_model.py:182-216), on the same GEMM / LayerNorm kernels as the video
// tower plus three small kernels of its own: token-embedding gather, causal attention over the 77-token context, and
// the end-of-text row gather.  Residual adds are folded into the following LayerNorm exactly as in the video tower.
#include <cmath>

#include "engine.h"
#include "hh_ptx.cuh"

namespace hh {

namespace {

constexpr int HD = 64;
constexpr int LDS = 72;  // smem row stride in bf16 (144 B): conflict-free ldmatrix
constexpr float LOG2E = 1.4426950408889634f;

// x[g*L + l, :] = token_embedding[tokens[g, l], :] + positional_embedding[l, :]      (LaviLa.py:661-662)
// Out-of-range ids are clamped and reported through *bad (nn.Embedding would device-assert).
__global__ void text_embed_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ table,
                                  const float* __restrict__ pos, float* __restrict__ x, int rows, int L, int W, int vocab,
                                  int* __restrict__ bad) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  long long id = tokens[row];
  if (id < 0 || id >= vocab) {
    if (lane == 0) atomicExch(bad, 1);
    id = id < 0 ? 0 : vocab - 1;
  }
  const float4* src = reinterpret_cast<const float4*>(table + static_cast<size_t>(id) * W);
  const float4* pp = reinterpret_cast<const float4*>(pos + static_cast<size_t>(row % L) * W);
  float4* dst = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * W);
  for (int c = lane; c < W / 4; c += 32) {
    const float4 a = __ldg(src + c), p = __ldg(pp + c);
    dst[c] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
  }
}

// Causal multi-head attention over one sequence (nn.MultiheadAttention with the triu(-inf) mask, openai_model.py:
// 199-201).  One CTA per (sequence, head); warp w owns query rows [16w, 16w+16) and only visits keys <= its last row.
// qkv bf16 [G*L, 3W] (q pre-scaled), out bf16 [G*L, W].
__global__ void __launch_bounds__(256) attn_causal_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int L,
                                                          int H) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int W = H * HD;
  const int h = blockIdx.x % H;
  const int g_ = blockIdx.x / H;
  const int mblocks = (L + 15) >> 4;
  const int rows = mblocks * 16;
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
  bf16* Ks = Qs + rows * LDS;
  bf16* Vs = Ks + rows * LDS;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const size_t ld = static_cast<size_t>(3) * W;
  const bf16* base = qkv + static_cast<size_t>(g_) * L * ld + h * HD;
  for (int c = tid; c < rows * 8; c += blockDim.x) {
    const int r = c >> 3, ch = c & 7;
    const bool valid = r < L;
    const bf16* src = base + static_cast<size_t>(valid ? r : 0) * ld + ch * 8;
    cp_async_16(Qs + r * LDS + ch * 8, src, valid);
    cp_async_16(Ks + r * LDS + ch * 8, src + W, valid);
    cp_async_16(Vs + r * LDS + ch * 8, src + 2 * W, valid);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  const int mi = lane >> 3, lr = lane & 7;
  for (int mb = warp; mb < mblocks; mb += (blockDim.x >> 5)) {
    const int r0 = mb * 16;
    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ldmatrix_x4(qf[ks], smem_u32(Qs + (r0 + (mi & 1) * 8 + lr) * LDS + ks * 16 + (mi >> 1) * 8));
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[8][4];
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) o[ni][0] = o[ni][1] = o[ni][2] = o[ni][3] = 0.f;
    const int row0 = r0 + g, row1 = r0 + g + 8;
    for (int kb = 0; kb <= mb; ++kb) {  // 16 keys per step; block kb == mb carries the diagonal
      float s[2][4];
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) s[ni][0] = s[ni][1] = s[ni][2] = s[ni][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t kf[4];
        ldmatrix_x4(kf, smem_u32(Ks + (kb * 16 + (mi >> 1) * 8 + lr) * LDS + ks * 16 + (mi & 1) * 8));
        mma_bf16_16816(s[0], qf[ks], kf[0], kf[1]);
        mma_bf16_16816(s[1], qf[ks], kf[2], kf[3]);
      }
      if (kb == mb) {  // causal mask: key > query row  (keys past L are > every live row as well)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
          const int key = kb * 16 + ni * 8 + 2 * t;
          if (key > row0) s[ni][0] = -INFINITY;
          if (key + 1 > row0) s[ni][1] = -INFINITY;
          if (key > row1) s[ni][2] = -INFINITY;
          if (key + 1 > row1) s[ni][3] = -INFINITY;
        }
      }
      float mx0 = fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1]));
      float mx1 = fmaxf(fmaxf(s[0][2], s[0][3]), fmaxf(s[1][2], s[1][3]));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      // key 0 <= every row, so after block 0 the running max is finite; inside block kb == mb the diagonal key is live
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
      const float ml0 = mn0 * LOG2E, ml1 = mn1 * LOG2E;
      const float c0 = fast_exp2(m0 * LOG2E - ml0), c1 = fast_exp2(m1 * LOG2E - ml1);
      m0 = mn0;
      m1 = mn1;
      l0 *= c0;
      l1 *= c1;
      uint32_t pa[4];
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) {
        const float p0 = fast_exp2(fmaf(s[ni][0], LOG2E, -ml0)), p1 = fast_exp2(fmaf(s[ni][1], LOG2E, -ml0));
        const float p2 = fast_exp2(fmaf(s[ni][2], LOG2E, -ml1)), p3 = fast_exp2(fmaf(s[ni][3], LOG2E, -ml1));
        l0 += p0 + p1;
        l1 += p2 + p3;
        pa[ni * 2 + 0] = pack_bf16x2(p0, p1);
        pa[ni * 2 + 1] = pack_bf16x2(p2, p3);
      }
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) {
        o[ni][0] *= c0; o[ni][1] *= c0; o[ni][2] *= c1; o[ni][3] *= c1;
      }
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, smem_u32(Vs + (kb * 16 + (mi & 1) * 8 + lr) * LDS + dp * 16 + (mi >> 1) * 8));
        mma_bf16_16816(o[2 * dp], pa, vf[0], vf[1]);
        mma_bf16_16816(o[2 * dp + 1], pa, vf[2], vf[3]);
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
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
      if (r0 + r < L) {
        const uint4 v = *reinterpret_cast<const uint4*>(Qs + (r0 + r) * LDS + ch * 8);
        *reinterpret_cast<uint4*>(out + (static_cast<size_t>(g_) * L + r0 + r) * W + h * HD + ch * 8) = v;
      }
    }
  }
}

// cls[g, :] = x[g, argmax_l tokens[g, l], :]   (first maximal index, as torch.argmax; LaviLa.py:669)
__global__ void text_gather_eot_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ x,
                                       float* __restrict__ cls, int L, int W) {
  const int g = blockIdx.x;
  __shared__ int s_idx;
  if (threadIdx.x < 32) {
    long long best = LLONG_MIN;
    int bi = 0;
    for (int l = threadIdx.x; l < L; l += 32) {
      const long long v = tokens[static_cast<size_t>(g) * L + l];
      if (v > best) {
        best = v;
        bi = l;
      }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      const long long ov = __shfl_xor_sync(0xffffffffu, best, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov > best || (ov == best && oi < bi)) {
        best = ov;
        bi = oi;
      }
    }
    if (threadIdx.x == 0) s_idx = bi;
  }
  __syncthreads();
  const float* src = x + (static_cast<size_t>(g) * L + s_idx) * W;
  for (int c = threadIdx.x; c < W; c += blockDim.x) cls[static_cast<size_t>(g) * W + c] = src[c];
}

// dst[c, r] = src[r, c]   (text_projection [W, E] -> nn.Linear layout [E, W])
__global__ void transpose_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[static_cast<size_t>(r) * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[static_cast<size_t>(c) * rows + r] = tile[threadIdx.x][i];
  }
}

}  // namespace

int attn_causal(const bf16* qkv, bf16* out, int G, int L, int H, cudaStream_t stream) {
  HH_REQUIRE(G > 0 && L > 0 && H > 0, "attn_causal: empty problem");
  HH_REQUIRE(L <= 512, "attn_causal: context length above 512 is not supported by the resident-K/V kernel");
  const int rows = (L + 15) / 16 * 16;
  const size_t smem = static_cast<size_t>(3) * rows * LDS * sizeof(bf16);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    HH_CHECK_CUDA(cudaFuncSetAttribute(attn_causal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    configured = smem;
  }
  const int warps = rows / 16 < 8 ? rows / 16 : 8;
  attn_causal_kernel<<<G * H, warps * 32, smem, stream>>>(qkv, out, L, H);
  HH_CHECK_LAUNCH("attn_causal_kernel");
  return 0;
}

#define RC(expr)         \
  do {                   \
    int _rc = (expr);    \
    if (_rc) return _rc; \
  } while (0)

// ========================================================================================== engine
TextEncoder::TextEncoder(const hh_text_cfg& c) : cfg(c) {
  const int64_t W = cfg.width, E = cfg.embed_dim;
  auto& e = weights.expected;
  e["token_embedding.weight"] = static_cast<int64_t>(cfg.vocab_size) * W;
  e["positional_embedding"] = static_cast<int64_t>(cfg.context_length) * W;
  e["ln_final.weight"] = W;
  e["ln_final.bias"] = W;
  e["text_projection"] = W * E;
  for (int i = 0; i < cfg.layers; ++i) {
    const std::string p = "transformer.resblocks." + std::to_string(i) + ".";
    for (const char* nm : {"ln_1", "ln_2"}) {
      e[p + nm + ".weight"] = W;
      e[p + nm + ".bias"] = W;
    }
    e[p + "attn.in_proj_weight"] = 3 * W * W;
    e[p + "attn.in_proj_bias"] = 3 * W;
    e[p + "attn.out_proj.weight"] = W * W;
    e[p + "attn.out_proj.bias"] = W;
    e[p + "mlp.c_fc.weight"] = 4 * W * W;
    e[p + "mlp.c_fc.bias"] = 4 * W;
    e[p + "mlp.c_proj.weight"] = 4 * W * W;
    e[p + "mlp.c_proj.bias"] = W;
  }
}

int TextEncoder::validate(const hh_text_cfg& c) {
  HH_REQUIRE(c.width % 128 == 0 && c.width <= 1024, "text: width must be a multiple of 128, <= 1024");
  HH_REQUIRE(c.heads > 0 && c.width == c.heads * 64, "text: head dim must be 64");
  HH_REQUIRE(c.layers >= 1 && c.vocab_size >= 1 && c.embed_dim >= 1, "text: layers / vocab / embed_dim");
  HH_REQUIRE(c.context_length >= 1 && c.context_length <= 512, "text: context length 1..512");
  return 0;
}

int TextEncoder::pack(cudaStream_t s) {
  RC(weights.check_complete());
  const int W = cfg.width, E = cfg.embed_dim;
  const float qscale = 1.0f / std::sqrt(64.0f);  // nn.MultiheadAttention scales q by head_dim^-0.5; folded into W_q, b_q
  layers.resize(cfg.layers);
  for (int i = 0; i < cfg.layers; ++i) {
    const std::string p = "transformer.resblocks." + std::to_string(i) + ".";
    Layer& Ly = layers[i];
    RC(Ly.w_qkv.reserve(static_cast<size_t>(3) * W * W * 2));
    RC(pack_weight_bf16(weights.get(p + "attn.in_proj_weight"), static_cast<bf16*>(Ly.w_qkv.ptr), 3 * W, W, W, W, qscale, s));
    RC(Ly.b_qkv.reserve(static_cast<size_t>(3) * W * 4));
    RC(scale_copy_f32(weights.get(p + "attn.in_proj_bias"), static_cast<float*>(Ly.b_qkv.ptr), 3 * W, W, qscale, s));
    RC(Ly.w_proj.reserve(static_cast<size_t>(W) * W * 2));
    RC(pack_weight_bf16(weights.get(p + "attn.out_proj.weight"), static_cast<bf16*>(Ly.w_proj.ptr), W, W, W, 0, 1.f, s));
    RC(Ly.w_fc1.reserve(static_cast<size_t>(4) * W * W * 2));
    RC(pack_weight_bf16(weights.get(p + "mlp.c_fc.weight"), static_cast<bf16*>(Ly.w_fc1.ptr), 4 * W, W, W, 0, 1.f, s));
    RC(Ly.w_fc2.reserve(static_cast<size_t>(4) * W * W * 2));
    RC(pack_weight_bf16(weights.get(p + "mlp.c_proj.weight"), static_cast<bf16*>(Ly.w_fc2.ptr), W, 4 * W, 4 * W, 0, 1.f, s));
  }
  RC(w_projT.reserve(static_cast<size_t>(W) * E * 4));
  {
    dim3 grid((E + 31) / 32, (W + 31) / 32), block(32, 8);
    transpose_f32_kernel<<<grid, block, 0, s>>>(weights.get("text_projection"), static_cast<float*>(w_projT.ptr), W, E);
    HH_CHECK_LAUNCH("transpose_f32_kernel");
  }
  RC(flag.reserve(sizeof(int)));
  weights.dirty = false;
  return 0;
}

int TextEncoder::forward(const int64_t* tokens, int G, float* embed, float* fmap, cudaStream_t s) {
  HH_REQUIRE(G > 0, "text forward: empty batch");
  HH_REQUIRE(tokens != nullptr && (embed != nullptr || fmap != nullptr), "text forward: null buffer");
  if (weights.dirty) RC(pack(s));
  launches = 0;
  const int W = cfg.width, L = cfg.context_length, H = cfg.heads, E = cfg.embed_dim;
  const int chunk = G < max_chunk ? G : max_chunk;
  const size_t Mc = static_cast<size_t>(chunk) * L;
  RC(ws_x.reserve(Mc * W * 4));
  RC(ws_dl.reserve(Mc * W * 2));
  RC(ws_a.reserve(Mc * W * 2));
  RC(ws_qkv.reserve(Mc * 3 * W * 2));
  RC(ws_h.reserve(Mc * 4 * W * 2));
  RC(ws_out.reserve(Mc * W * 4));
  RC(ws_cls.reserve(static_cast<size_t>(chunk) * W * 4));
  float* x = static_cast<float*>(ws_x.ptr);
  bf16* dl = static_cast<bf16*>(ws_dl.ptr);
  bf16* a = static_cast<bf16*>(ws_a.ptr);
  bf16* qkv = static_cast<bf16*>(ws_qkv.ptr);
  bf16* h = static_cast<bf16*>(ws_h.ptr);
  float* cls = static_cast<float*>(ws_cls.ptr);
  int* bad = static_cast<int*>(flag.ptr);
  HH_CHECK_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));

  for (int g0 = 0; g0 < G; g0 += chunk) {
    const int Gc = (G - g0) < chunk ? (G - g0) : chunk;
    const int M = Gc * L;
    const int64_t* tok = tokens + static_cast<size_t>(g0) * L;
    text_embed_kernel<<<(M + 7) / 8, 256, 0, s>>>(tok, weights.get("token_embedding.weight"),
                                                 weights.get("positional_embedding"), x, M, L, W, cfg.vocab_size, bad);
    HH_CHECK_LAUNCH("text_embed_kernel");
    launches += 1;
    const bf16* pending = nullptr;
    auto ln_fused = [&](const bf16* delta, const std::string& nm, bf16* out16, float* out32) {
      LnArgs ln{};
      ln.x = x;
      ln.ldx = W;
      ln.delta = delta;
      ln.xsum_out = delta ? x : nullptr;
      ln.w = weights.get(nm + ".weight");
      ln.b = weights.get(nm + ".bias");
      ln.eps = 1e-5f;
      ln.out_bf16 = out16;
      ln.out_f32 = out32;
      ln.M = M;
      ln.D = W;
      return layernorm_rows(ln, s);
    };
    for (int i = 0; i < cfg.layers; ++i) {
      const std::string p = "transformer.resblocks." + std::to_string(i) + ".";
      const Layer& Ly = layers[i];
      RC(ln_fused(pending, p + "ln_1", a, nullptr));  // x <- x + mlp_out(prev) ; a = ln_1(x)
      RC(gemm_bf16(a, W, static_cast<const bf16*>(Ly.w_qkv.ptr), W, qkv, 3 * W, static_cast<const float*>(Ly.b_qkv.ptr),
                   nullptr, 0, M, 3 * W, W, EPI_BIAS_BF16, s));
      RC(attn_causal(qkv, a, Gc, L, H, s));
      RC(gemm_bf16(a, W, static_cast<const bf16*>(Ly.w_proj.ptr), W, dl, W, weights.get(p + "attn.out_proj.bias"), nullptr, 0,
                   M, W, W, EPI_BIAS_BF16, s));
      RC(ln_fused(dl, p + "ln_2", a, nullptr));  // x <- x + attn_out ; a = ln_2(x)
      RC(gemm_bf16(a, W, static_cast<const bf16*>(Ly.w_fc1.ptr), W, h, 4 * W, weights.get(p + "mlp.c_fc.bias"), nullptr, 0, M,
                   4 * W, W, EPI_BIAS_QGELU_BF16, s));
      RC(gemm_bf16(h, 4 * W, static_cast<const bf16*>(Ly.w_fc2.ptr), 4 * W, dl, W, weights.get(p + "mlp.c_proj.bias"), nullptr,
                   0, M, W, 4 * W, EPI_BIAS_BF16, s));
      pending = dl;
      launches += 7;
    }
    float* fm = fmap ? fmap + static_cast<size_t>(g0) * L * W : static_cast<float*>(ws_out.ptr);
    RC(ln_fused(pending, "ln_final", nullptr, fm));
    launches += 1;
    if (embed) {
      text_gather_eot_kernel<<<Gc, 128, 0, s>>>(tok, fm, cls, L, W);
      HH_CHECK_LAUNCH("text_gather_eot_kernel");
      LinArgs la{};
      la.in = cls; la.ldi = W; la.W = static_cast<const float*>(w_projT.ptr);
      la.out = embed + static_cast<size_t>(g0) * E; la.ldo = E; la.R = Gc; la.N = E; la.K = W;
      RC(linear_f32(la, s));
      launches += 2;
    }
  }
  int hbad = 0;
  HH_CHECK_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
  HH_CHECK_CUDA(cudaStreamSynchronize(s));
  HH_REQUIRE(hbad == 0, "text forward: token id outside [0, vocab_size)");
  return 0;
}

double TextEncoder::flops_per_sequence() const {
  const double W = cfg.width, L = cfg.context_length;
  return cfg.layers * (24.0 * L * W * W + 4.0 * W * L * (L + 1) / 2.0) + 2.0 * W * cfg.embed_dim;
}

}  // namespace hh
