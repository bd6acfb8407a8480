// Retrieval metrics of the multi-instance-retrieval evaluation: per-query average precision
// (utils/mAP.py:4-44) and discounted cumulative gain (utils/nDCG.py:3-44) of an N x M similarity matrix against a
// relevancy matrix, float64 like the numpy reference.  One CTA per query row: the row's (similarity, index) pairs are
// sorted in shared memory with a bitonic network (M <= 16384: 192 KB), the relevancies are gathered in ranked order and
// reduced in place -- the ranked matrices the reference materialises (N x M float64, three of them) never exist.
//
// Ordering.  mAP ranks with argsort(-sim), nDCG with argsort(sim)[:, ::-1].  numpy's default sort is not stable, so
// for tied similarities the reference's order is unspecified; this kernel uses the order a STABLE sort gives in each
// formula (ties by ascending index for mAP, by descending index for nDCG), which is also what numpy does for rows of
// fewer than 17 elements (insertion sort) and whenever there are no ties.
#include "hh_internal.h"

namespace hh {

namespace {

constexpr int RT = 1024;  // threads per CTA

struct RetrArgs {
  const double* sim;   // [N, M]
  const double* rel;   // [N, M]
  const double* logs;  // [M] log2(k + 2) (nDCG only; computed by the caller so that it is numpy's own table)
  const int* kcounts;  // optional [N, M] (nDCG): mask of ranks that count; null -> rank < #positive relevancies of the row
  int N, M, P;         // P = M rounded up to a power of two
  int mode;            // 0 = average precision, 1 = DCG
  double* out;         // [N]
};

// numpy's pairwise summation (the reduction np.sum applies along a contiguous axis): blocks of <= 128 elements are
// summed with 8 interleaved accumulators, larger ranges are split at n/2 rounded down to a multiple of 8 and the halves
// added.  The recursion is unrolled onto an explicit stack (depth <= log2(16384 / 128) + 1).
__device__ __forceinline__ double pairwise_leaf(const double* a, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res += a[i];
    return res;
  }
  double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
    r0 += a[i]; r1 += a[i + 1]; r2 += a[i + 2]; r3 += a[i + 3];
    r4 += a[i + 4]; r5 += a[i + 5]; r6 += a[i + 6]; r7 += a[i + 7];
  }
  double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
  for (; i < n; ++i) res += a[i];
  return res;
}

__device__ double pairwise_sum(const double* a, int n) {
  int off[12], len[12], stage[12];
  double left[12];
  int sp = 1;
  off[0] = 0; len[0] = n; stage[0] = 0;
  double result = 0.0;
  while (sp > 0) {
    const int t = sp - 1;
    if (len[t] <= 128) {
      result = pairwise_leaf(a + off[t], len[t]);
      --sp;
      continue;
    }
    int n2 = len[t] / 2;
    n2 -= n2 % 8;
    if (stage[t] == 0) {
      stage[t] = 1;
      off[sp] = off[t]; len[sp] = n2; stage[sp] = 0;
      ++sp;
    } else if (stage[t] == 1) {
      left[t] = result;
      stage[t] = 2;
      off[sp] = off[t] + n2; len[sp] = len[t] - n2; stage[sp] = 0;
      ++sp;
    } else {
      result = left[t] + result;
      --sp;
    }
  }
  return result;
}

// ascending by (key, tie): mode 0 sorts (-sim, idx), mode 1 sorts (-sim, -idx)
__device__ __forceinline__ bool before(double ka, int ia, double kb, int ib) { return ka < kb || (ka == kb && ia < ib); }

__global__ void __launch_bounds__(RT) retrieval_rows_kernel(RetrArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* key = reinterpret_cast<double*>(smraw);          // [P]
  int* idx = reinterpret_cast<int*>(key + a.P);            // [P]
  __shared__ double red[32];
  const int row = blockIdx.x;
  const int tid = threadIdx.x;
  const double* s = a.sim + static_cast<size_t>(row) * a.M;
  const double* r = a.rel + static_cast<size_t>(row) * a.M;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  for (int i = tid; i < a.P; i += RT) {
    if (i < a.M) {
      key[i] = -s[i];
      idx[i] = a.mode == 0 ? i : -i;
    } else {
      key[i] = inf;  // padding sorts last
      idx[i] = 0x7fffffff;
    }
  }
  __syncthreads();
  for (int k = 2; k <= a.P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < a.P; i += RT) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = (i & k) == 0;
          const double ki = key[i], kl = key[l];
          const int ii = idx[i], il = idx[l];
          const bool swap = up ? before(kl, il, ki, ii) : before(ki, ii, kl, il);
          if (swap) {
            key[i] = kl; key[l] = ki;
            idx[i] = il; idx[l] = ii;
          }
        }
      }
      __syncthreads();
    }
  }
  // ranked relevancies replace the keys
  for (int i = tid; i < a.P; i += RT) {
    if (i < a.M) {
      const int src = a.mode == 0 ? idx[i] : -idx[i];
      key[i] = r[src];
    } else {
      key[i] = 0.0;
    }
  }
  __syncthreads();

  // numpy's own arithmetic from here on, so the float64 results reproduce to the last bit: cumsum is sequential
  // (utils/mAP.py:32), np.sum over the contiguous axis is 0 + pairwise_sum (8 accumulators per <= 128-element leaf).
  int nrel = 0;
  if (a.mode == 0) {
    if (tid == 0) {
      double cum = 0.0;
      for (int k = 0; k < a.M; ++k) {
        const double v = key[k];
        cum += v;
        const bool hit = v == 1.0;  // rel(k): only exact ones count (utils/mAP.py:34,39)
        key[k] = hit ? cum / static_cast<double>(k + 1) : 0.0;
        nrel += hit;
      }
    }
  } else {
    double np = 0.0;
    if (!a.kcounts) {  // calculate_k_counts (utils/nDCG.py:46-75): the first #(rel > 0) ranks count
      for (int k = tid; k < a.M; k += RT) np += r[k] > 0.0 ? 1.0 : 0.0;
      for (int o = 16; o; o >>= 1) np += __shfl_xor_sync(0xffffffffu, np, o);
      if ((tid & 31) == 0) red[tid >> 5] = np;
      __syncthreads();
      np = 0.0;
      for (int w = 0; w < RT / 32; ++w) np += red[w];
    }
    const int npos = static_cast<int>(np);
    for (int k = tid; k < a.M; k += RT) {
      const double kc = a.kcounts ? static_cast<double>(a.kcounts[static_cast<size_t>(row) * a.M + k]) : (k < npos ? 1.0 : 0.0);
      key[k] = key[k] * kc / a.logs[k];
    }
  }
  __syncthreads();
  if (tid == 0) {
    const double res = 0.0 + pairwise_sum(key, a.M);
    a.out[row] = a.mode == 0 ? res / static_cast<double>(nrel) : res;
  }
}

}  // namespace

int retrieval_rows(const double* sim, const double* rel, const double* logs, const int* kcounts, int N, int M, int mode,
                   double* out, cudaStream_t stream) {
  HH_REQUIRE(N > 0 && M > 0 && sim && rel && out, "retrieval_rows: bad argument");
  HH_REQUIRE(mode == 0 || (mode == 1 && logs != nullptr), "retrieval_rows: mode 1 (DCG) needs the log2 table");
  HH_REQUIRE(M <= 16384, "retrieval_rows: more than 16384 candidates per query");
  int P = 32;
  while (P < M) P <<= 1;
  const size_t smem = static_cast<size_t>(P) * 12;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    HH_CHECK_CUDA(cudaFuncSetAttribute(retrieval_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    configured = smem;
  }
  RetrArgs a{sim, rel, logs, kcounts, N, M, P, mode, out};
  retrieval_rows_kernel<<<N, RT, smem, stream>>>(a);
  HH_CHECK_LAUNCH("retrieval_rows_kernel");
  return 0;
}

}  // namespace hh
