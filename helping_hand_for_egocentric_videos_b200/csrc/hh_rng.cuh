// Counter-based dropout masks for the decoder's training path (nn.Dropout at model/tfm_decoder.py:372-386,
// nn.MultiheadAttention(dropout=0.1) at :365-366): Philox4x32-10 keyed by the step's (seed, offset), one 16-bit lane
// per element, so the backward pass regenerates the forward's masks instead of storing them:
//     r = lane (idx & 7) of philox(key = seed, counter = (idx >> 3, site, offset));   keep  <=>  r >= thr = round(p * 65536)
// `site` = layer * 8 + {0 self-attention probabilities, 1 dropout1, 2 cross-attention probabilities, 3 dropout2,
// 4 FFN inner dropout, 5 dropout3}; idx = flat element index of the dropped tensor (stated where it is used).
// The same function is restated in oracle/hh_oracle.py (philox_keep) so that tests can run the reference graph with
// identical masks.  The distribution is torch's (Bernoulli(1-p) keep, 1/(1-p) scaling); the random stream is ours.
#pragma once
#include <cstdint>

namespace hh {

struct DropCfg {
  uint32_t thr;       // 0: dropout off
  float scale;        // 1 / (1 - p)
  uint32_t seed_lo, seed_hi, offset;
};

__host__ __device__ inline DropCfg drop_off() { return DropCfg{0u, 1.f, 0u, 0u, 0u}; }

__device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 16-bit lane of element idx at dropout site `site`
__device__ __forceinline__ uint32_t drop_bits(const DropCfg& d, uint32_t site, uint64_t idx) {
  uint32_t o[4];
  const uint64_t blk = idx >> 3;
  philox4x32_10(d.seed_lo, d.seed_hi, static_cast<uint32_t>(blk), static_cast<uint32_t>(blk >> 32), site, d.offset, o);
  const uint32_t lane = static_cast<uint32_t>(idx) & 7u, wi = lane >> 1;
  const uint32_t w = wi == 0 ? o[0] : wi == 1 ? o[1] : wi == 2 ? o[2] : o[3];   // selects, not a local-memory array
  return (w >> ((lane & 1u) * 16u)) & 0xFFFFu;
}
// multiplier of element idx: 0 (dropped) or 1/(1-p) (kept)
__device__ __forceinline__ float drop_mult(const DropCfg& d, uint32_t site, uint64_t idx) {
  return drop_bits(d, site, idx) >= d.thr ? d.scale : 0.f;
}
// the 8 lanes of one Philox block (elements 8*blk .. 8*blk+7) at once
__device__ __forceinline__ void drop_block8(const DropCfg& d, uint32_t site, uint64_t blk, uint32_t (&o)[4]) {
  philox4x32_10(d.seed_lo, d.seed_hi, static_cast<uint32_t>(blk), static_cast<uint32_t>(blk >> 32), site, d.offset, o);
}

}  // namespace hh
