// Box utilities of utils/box_ops.py:9-61 and the matcher cost of model/box_utils.py:75-88, fp32.
//   * one 16-byte load per box (float4 = one box), no intermediate tensors
//   * pairwise kernels: a warp owns one box of the first set (broadcast by shuffle from lane 0's load) and sweeps
//     the second set 32 boxes at a time, so every store is a coalesced 128-byte row segment
//   * arithmetic order follows the reference exactly (iou = inter / (union + 1e-4); giou = iou - (hull-union)/hull)
//     so the Hungarian indices computed from the cost are bit-stable.
#include "hh_internal.h"
#include "hh_ptx.cuh"

namespace hh {

namespace {

__global__ void __launch_bounds__(256)
cxcywh_to_xyxy_kernel(const float4* __restrict__ in, float4* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 b = in[i];
    out[i] = make_float4(b.x - 0.5f * b.z, b.y - 0.5f * b.w, b.x + 0.5f * b.z, b.y + 0.5f * b.w);
  }
}

__global__ void __launch_bounds__(256)
xyxy_to_cxcywh_kernel(const float4* __restrict__ in, float4* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 b = in[i];
    out[i] = make_float4((b.x + b.z) / 2.f, (b.y + b.w) / 2.f, b.z - b.x, b.w - b.y);
  }
}

__device__ __forceinline__ float4 to_xyxy(const float4 b) {  // 0.5*w is exact, so an FMA here changes nothing
  return make_float4(b.x - 0.5f * b.z, b.y - 0.5f * b.w, b.x + 0.5f * b.z, b.y + 0.5f * b.w);
}

struct PairOut {
  float iou, uni, giou;
};

// __fmul_rn/__fadd_rn/__fsub_rn are never contracted into FMAs, so every rounding matches the eager reference.
__device__ __forceinline__ PairOut pair_metrics(const float4 p, const float4 t) {
  const float area1 = __fmul_rn(__fsub_rn(p.z, p.x), __fsub_rn(p.w, p.y));
  const float area2 = __fmul_rn(__fsub_rn(t.z, t.x), __fsub_rn(t.w, t.y));
  const float iw = fmaxf(__fsub_rn(fminf(p.z, t.z), fmaxf(p.x, t.x)), 0.f);
  const float ih = fmaxf(__fsub_rn(fminf(p.w, t.w), fmaxf(p.y, t.y)), 0.f);
  const float inter = __fmul_rn(iw, ih);
  const float uni = __fsub_rn(__fadd_rn(area1, area2), inter);
  const float iou = __fdiv_rn(inter, __fadd_rn(uni, 0.0001f));
  const float hw = fmaxf(__fsub_rn(fmaxf(p.z, t.z), fminf(p.x, t.x)), 0.f);
  const float hh_ = fmaxf(__fsub_rn(fmaxf(p.w, t.w), fminf(p.y, t.y)), 0.f);
  const float hull = __fmul_rn(hw, hh_);
  PairOut o;
  o.iou = iou;
  o.uni = uni;
  o.giou = __fsub_rn(iou, __fdiv_rn(__fsub_rn(hull, uni), hull));
  return o;
}

// MODE 0: xyxy inputs -> iou/union/giou.  MODE 1: cxcywh inputs -> matcher cost.
template <int MODE>
__global__ void __launch_bounds__(256)
pairwise_kernel(const float4* __restrict__ b1, const float4* __restrict__ b2, int N, int M, float* __restrict__ iou,
                float* __restrict__ uni, float* __restrict__ giou, float w_bbox, float w_giou) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  float4 p;
  if (lane == 0) p = b1[row];
  p.x = __shfl_sync(0xffffffffu, p.x, 0);
  p.y = __shfl_sync(0xffffffffu, p.y, 0);
  p.z = __shfl_sync(0xffffffffu, p.z, 0);
  p.w = __shfl_sync(0xffffffffu, p.w, 0);
  const float4 pxy = (MODE == 1) ? to_xyxy(p) : p;
  for (int c = lane; c < M; c += 32) {
    const float4 t = b2[c];
    const size_t o = static_cast<size_t>(row) * M + c;
    if (MODE == 0) {
      const PairOut r = pair_metrics(pxy, t);
      if (iou) iou[o] = r.iou;
      if (uni) uni[o] = r.uni;
      if (giou) giou[o] = r.giou;
    } else {
      const PairOut r = pair_metrics(pxy, to_xyxy(t));
      const float l1 = __fadd_rn(__fadd_rn(__fadd_rn(fabsf(__fsub_rn(p.x, t.x)), fabsf(__fsub_rn(p.y, t.y))),
                                           fabsf(__fsub_rn(p.z, t.z))), fabsf(__fsub_rn(p.w, t.w)));
      giou[o] = __fadd_rn(__fmul_rn(w_bbox, l1), __fmul_rn(w_giou, -r.giou));
    }
  }
}

// ---- matched-pair box loss (SetCriterion.loss_boxes, model/box_utils.py:157-173), forward and backward
// pair k: prediction row src_row[k] of pred (cxcywh) against tgt[k] (cxcywh).  losses[0] = sum |p - t| / num_boxes,
// losses[1] = sum (1 - giou(p, t)) / num_boxes.  One CTA; fixed-order tree reduction (deterministic).
__global__ void __launch_bounds__(256)
box_loss_fwd_kernel(const float4* __restrict__ pred, const long long* __restrict__ src_row, const float4* __restrict__ tgt,
                    int K, float num_boxes, float* __restrict__ losses) {
  __shared__ float s1[256], s2[256];
  float l1 = 0.f, lg = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float4 p = pred[src_row[k]], t = tgt[k];
    l1 += __fadd_rn(__fadd_rn(__fadd_rn(fabsf(__fsub_rn(p.x, t.x)), fabsf(__fsub_rn(p.y, t.y))), fabsf(__fsub_rn(p.z, t.z))),
                    fabsf(__fsub_rn(p.w, t.w)));
    lg += __fsub_rn(1.f, pair_metrics(to_xyxy(p), to_xyxy(t)).giou);
  }
  s1[threadIdx.x] = l1;
  s2[threadIdx.x] = lg;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) {
      s1[threadIdx.x] += s1[threadIdx.x + o];
      s2[threadIdx.x] += s2[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    losses[0] = s1[0] / num_boxes;
    losses[1] = s2[0] / num_boxes;
  }
}

// autograd conventions of the eager reference: maximum/minimum split the gradient evenly on ties, clamp(min=0) passes
// it where the argument is >= 0, sign(0) = 0 for the L1 term.
__device__ __forceinline__ float sel_gt(float a, float b) { return a > b ? 1.f : (a == b ? 0.5f : 0.f); }

__global__ void __launch_bounds__(256)
box_loss_bwd_kernel(const float4* __restrict__ pred, const long long* __restrict__ src_row, const float4* __restrict__ tgt,
                    int K, float num_boxes, const float* __restrict__ g, float4* __restrict__ grad_pred) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const float gb = g[0] / num_boxes, gg = g[1] / num_boxes;
  const float4 pc = pred[src_row[k]], tc = tgt[k];
  const float4 p = to_xyxy(pc), t = to_xyxy(tc);
  auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
  // d(1 - giou) / d(x1, y1, x2, y2)
  const float pw = p.z - p.x, ph = p.w - p.y;
  const float area1 = pw * ph, area2 = (t.z - t.x) * (t.w - t.y);
  const float ltx = fmaxf(p.x, t.x), lty = fmaxf(p.y, t.y), rbx = fminf(p.z, t.z), rby = fminf(p.w, t.w);
  const float iw_raw = rbx - ltx, ih_raw = rby - lty;
  const float iw = fmaxf(iw_raw, 0.f), ih = fmaxf(ih_raw, 0.f);
  const float inter = iw * ih;
  const float uni = area1 + area2 - inter;
  const float den = uni + 0.0001f;
  const float cltx = fminf(p.x, t.x), clty = fminf(p.y, t.y), crbx = fmaxf(p.z, t.z), crby = fmaxf(p.w, t.w);
  const float hw_raw = crbx - cltx, hh_raw = crby - clty;
  const float hw = fmaxf(hw_raw, 0.f), hh_ = fmaxf(hh_raw, 0.f);
  const float hull = hw * hh_;
  // giou = inter/den - 1 + uni/hull
  const float d_inter = 1.f / den, d_uni_iou = -inter / (den * den);
  const float d_uni = d_uni_iou + 1.f / hull;
  const float d_hull = -uni / (hull * hull);
  // uni = area1 + area2 - inter  ->  total derivative w.r.t. inter and area1
  const float t_inter = d_inter - d_uni, t_area1 = d_uni;
  const float g_iw = t_inter * ih * (iw_raw >= 0.f ? 1.f : 0.f), g_ih = t_inter * iw * (ih_raw >= 0.f ? 1.f : 0.f);
  const float g_hw = d_hull * hh_ * (hw_raw >= 0.f ? 1.f : 0.f), g_hh = d_hull * hw * (hh_raw >= 0.f ? 1.f : 0.f);
  float dx1 = -t_area1 * ph - g_iw * sel_gt(p.x, t.x) - g_hw * sel_gt(t.x, p.x);
  float dy1 = -t_area1 * pw - g_ih * sel_gt(p.y, t.y) - g_hh * sel_gt(t.y, p.y);
  float dx2 = t_area1 * ph + g_iw * sel_gt(t.z, p.z) + g_hw * sel_gt(p.z, t.z);
  float dy2 = t_area1 * pw + g_ih * sel_gt(t.w, p.w) + g_hh * sel_gt(p.w, t.w);
  // loss = 1 - giou
  dx1 = -dx1 * gg; dy1 = -dy1 * gg; dx2 = -dx2 * gg; dy2 = -dy2 * gg;
  float4 o;
  o.x = dx1 + dx2 + gb * sgn(pc.x - tc.x);
  o.y = dy1 + dy2 + gb * sgn(pc.y - tc.y);
  o.z = 0.5f * (dx2 - dx1) + gb * sgn(pc.z - tc.z);
  o.w = 0.5f * (dy2 - dy1) + gb * sgn(pc.w - tc.w);
  grad_pred[src_row[k]] = o;
}

int grid_for(long long n) {
  long long blocks = (n + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace

int box_cxcywh_to_xyxy(const float* in, float* out, long long nboxes, cudaStream_t stream) {
  if (nboxes == 0) return 0;
  HH_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "box_cxcywh_to_xyxy: 16-byte alignment");
  cxcywh_to_xyxy_kernel<<<grid_for(nboxes), 256, 0, stream>>>(reinterpret_cast<const float4*>(in),
                                                             reinterpret_cast<float4*>(out), nboxes);
  HH_CHECK_LAUNCH("cxcywh_to_xyxy_kernel");
  return 0;
}

int box_xyxy_to_cxcywh(const float* in, float* out, long long nboxes, cudaStream_t stream) {
  if (nboxes == 0) return 0;
  HH_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "box_xyxy_to_cxcywh: 16-byte alignment");
  xyxy_to_cxcywh_kernel<<<grid_for(nboxes), 256, 0, stream>>>(reinterpret_cast<const float4*>(in),
                                                             reinterpret_cast<float4*>(out), nboxes);
  HH_CHECK_LAUNCH("xyxy_to_cxcywh_kernel");
  return 0;
}

int box_pairwise(const float* b1, const float* b2, int N, int M, float* iou, float* uni, float* giou,
                 cudaStream_t stream) {
  if (N == 0 || M == 0) return 0;
  HH_REQUIRE((reinterpret_cast<uintptr_t>(b1) & 15) == 0 && (reinterpret_cast<uintptr_t>(b2) & 15) == 0,
             "box_pairwise: 16-byte alignment");
  const int grid = (N + 7) / 8;
  pairwise_kernel<0><<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(b1), reinterpret_cast<const float4*>(b2),
                                               N, M, iou, uni, giou, 0.f, 0.f);
  HH_CHECK_LAUNCH("pairwise_kernel<0>");
  return 0;
}

int box_match_cost(const float* pred, const float* tgt, int N, int M, float w_bbox, float w_giou, float* cost,
                   cudaStream_t stream) {
  if (N == 0 || M == 0) return 0;
  HH_REQUIRE((reinterpret_cast<uintptr_t>(pred) & 15) == 0 && (reinterpret_cast<uintptr_t>(tgt) & 15) == 0,
             "box_match_cost: 16-byte alignment");
  const int grid = (N + 7) / 8;
  pairwise_kernel<1><<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(pred),
                                               reinterpret_cast<const float4*>(tgt), N, M, nullptr, nullptr, cost,
                                               w_bbox, w_giou);
  HH_CHECK_LAUNCH("pairwise_kernel<1>");
  return 0;
}

int box_loss_forward(const float* pred, const long long* src_row, const float* tgt, int K, float num_boxes, float* losses,
                     cudaStream_t stream) {
  HH_REQUIRE(K >= 0 && num_boxes > 0.f && losses, "box_loss_forward: bad argument");
  HH_REQUIRE(K == 0 || (pred && src_row && tgt), "box_loss_forward: null buffer");
  box_loss_fwd_kernel<<<1, 256, 0, stream>>>(reinterpret_cast<const float4*>(pred), src_row,
                                            reinterpret_cast<const float4*>(tgt), K, num_boxes, losses);
  HH_CHECK_LAUNCH("box_loss_fwd_kernel");
  return 0;
}

int box_loss_backward(const float* pred, const long long* src_row, const float* tgt, int K, float num_boxes,
                      const float* g_losses, float* grad_pred, long long pred_rows, cudaStream_t stream) {
  HH_REQUIRE(K >= 0 && num_boxes > 0.f && g_losses && grad_pred && pred_rows >= 0, "box_loss_backward: bad argument");
  HH_CHECK_CUDA(cudaMemsetAsync(grad_pred, 0, static_cast<size_t>(pred_rows) * 16, stream));
  if (K == 0) return 0;
  box_loss_bwd_kernel<<<(K + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const float4*>(pred), src_row,
                                                          reinterpret_cast<const float4*>(tgt), K, num_boxes, g_losses,
                                                          reinterpret_cast<float4*>(grad_pred));
  HH_CHECK_LAUNCH("box_loss_bwd_kernel");
  return 0;
}

}  // namespace hh
