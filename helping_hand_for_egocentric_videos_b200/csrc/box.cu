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

}  // namespace hh
