// Exact linear-sum assignment for batches of small rectangular problems, on the device.
//
// The reference solves one tiny problem per image / per clip on the host with scipy.optimize.linear_sum_assignment after
// a device->host copy of the cost block (HungarianMatcher.forward, model/box_utils.py:89-92: num_queries x n_targets,
// at most 13 x 8; WordContrastiveLoss.forward, model/loss.py:88-93: <= 4 nouns x 12 queries).  scipy is not part of the
// reference tree; its solver is the shortest-augmenting-path algorithm of D. F. Crouse, "On implementing 2D rectangular
// assignment algorithms", IEEE T-AES 52(4), 2016, run in float64 on the (transposed if taller than wide) cost matrix.
// This kernel restates that published algorithm, one thread per problem, float64 duals, including its scan order and
// its tie-breaks (prefer an unassigned column among equal reduced costs), so that the INDICES are the ones the
// reference obtains even when optimal assignments are not unique.  Also here: the class-probability term of the matcher
// cost (model/box_utils.py:66,83-85).
#include <cfloat>

#include "hh_internal.h"

namespace hh {

namespace {

constexpr int MAXD = 32;  // max rows / columns of one problem

struct AssignArgs {
  const float* cost;
  const long long* offset;  // [P] element offset of problem p's (0,0) entry
  const int* ld;            // [P] row stride
  const int* nr;            // [P] rows before compaction
  const int* nc;            // [P] columns
  const unsigned char* row_valid;  // optional [P, row_valid_ld]: rows with 0 are dropped (compacted away) first
  int row_valid_ld;
  int P;
  long long* row_ind;  // [P, out_ld], -1 padded
  long long* col_ind;  // [P, out_ld]
  int* count;          // [P] = min(rows, cols) of the compacted problem
  int out_ld;
};

__global__ void assign_kernel(AssignArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.P) return;
  const float* C = a.cost + a.offset[p];
  const int ld = a.ld[p];
  int rows[MAXD];
  int R = 0;
  if (a.nr[p] > MAXD || a.nc[p] > MAXD || a.nr[p] < 0 || a.nc[p] < 0) {  // caller broke its max_dim promise
    a.count[p] = -1;
    return;
  }
  for (int r = 0; r < a.nr[p]; ++r)
    if (!a.row_valid || a.row_valid[static_cast<size_t>(p) * a.row_valid_ld + r]) rows[R++] = r;
  const int Cn = a.nc[p];
  long long* ri = a.row_ind + static_cast<size_t>(p) * a.out_ld;
  long long* ci = a.col_ind + static_cast<size_t>(p) * a.out_ld;
  for (int k = 0; k < a.out_ld; ++k) ri[k] = ci[k] = -1;
  const bool tr = Cn < R;           // the solver works on the wide orientation
  const int n = tr ? Cn : R;        // solver rows
  const int m = tr ? R : Cn;        // solver columns
  a.count[p] = n;
  if (n == 0) return;
  if (n > a.out_ld) {
    a.count[p] = -1;
    return;
  }
  auto cost_at = [&](int i, int j) -> double {  // solver (i, j)
    return tr ? static_cast<double>(C[static_cast<size_t>(rows[j]) * ld + i])
              : static_cast<double>(C[static_cast<size_t>(rows[i]) * ld + j]);
  };
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  double u[MAXD], v[MAXD], sp[MAXD];
  int col4row[MAXD], row4col[MAXD], path[MAXD], remaining[MAXD];
  unsigned sr, sc;
  for (int i = 0; i < n; ++i) { u[i] = 0.0; col4row[i] = -1; }
  for (int j = 0; j < m; ++j) { v[j] = 0.0; row4col[j] = -1; }
  for (int cur = 0; cur < n; ++cur) {
    double min_val = 0.0;
    int i = cur;
    int nrem = m;
    for (int it = 0; it < m; ++it) {
      remaining[it] = m - it - 1;
      sp[it] = INF;
      path[it] = -1;
    }
    sr = sc = 0u;
    int sink = -1;
    while (sink == -1) {
      int index = -1;
      double lowest = INF;
      sr |= 1u << i;
      for (int it = 0; it < nrem; ++it) {
        const int j = remaining[it];
        const double r = min_val + cost_at(i, j) - u[i] - v[j];
        if (r < sp[j]) {
          path[j] = i;
          sp[j] = r;
        }
        if (sp[j] < lowest || (sp[j] == lowest && row4col[j] == -1)) {
          lowest = sp[j];
          index = it;
        }
      }
      min_val = lowest;
      if (index < 0 || !(min_val < INF)) {  // infeasible (inf / nan costs): report nothing for this problem
        a.count[p] = -1;
        return;
      }
      const int j = remaining[index];
      if (row4col[j] == -1) sink = j;
      else i = row4col[j];
      sc |= 1u << j;
      remaining[index] = remaining[--nrem];
    }
    u[cur] += min_val;
    for (int r = 0; r < n; ++r)
      if (((sr >> r) & 1u) && r != cur) u[r] += min_val - sp[col4row[r]];
    for (int j = 0; j < m; ++j)
      if ((sc >> j) & 1u) v[j] -= min_val - sp[j];
    int j = sink;
    while (true) {
      const int r = path[j];
      row4col[j] = r;
      const int t = col4row[r];
      col4row[r] = j;
      j = t;
      if (r == cur) break;
    }
  }
  // results ordered by the row index of the (compacted) input matrix
  if (!tr) {
    for (int r = 0; r < n; ++r) {
      ri[r] = r;
      ci[r] = col4row[r];
    }
  } else {
    int k = 0;
    for (int r = 0; r < m; ++r)  // input rows = solver columns
      if (row4col[r] != -1) {
        ri[k] = r;
        ci[k] = row4col[r];
        ++k;
      }
  }
}

// cost[r, t] += w * -softmax(logits[r, :])[ids[t]]     (one CTA per prediction row)
__global__ void class_cost_kernel(const float* __restrict__ logits, int ncls, const long long* __restrict__ ids, int M,
                                  float w, float* __restrict__ cost) {
  const int r = blockIdx.x;
  const float* x = logits + static_cast<size_t>(r) * ncls;
  __shared__ float red[32];
  __shared__ float s_max, s_sum;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < ncls; c += blockDim.x) mx = fmaxf(mx, x[c]);
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
    for (int o = 16; o; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
    if (threadIdx.x == 0) s_max = t;
  }
  __syncthreads();
  mx = s_max;
  float sum = 0.f;
  for (int c = threadIdx.x; c < ncls; c += blockDim.x) sum += expf(x[c] - mx);
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) s_sum = t;
  }
  __syncthreads();
  const float inv = 1.f / s_sum;
  for (int t = threadIdx.x; t < M; t += blockDim.x) {
    const long long id = ids[t];
    const float pr = (id >= 0 && id < ncls) ? expf(x[id] - mx) * inv : 0.f;
    cost[static_cast<size_t>(r) * M + t] += w * -pr;
  }
}

}  // namespace

int assign_lsa(const float* cost, const long long* offset, const int* ld, const int* nr, const int* nc,
               const unsigned char* row_valid, int row_valid_ld, int P, int max_dim, long long* row_ind,
               long long* col_ind, int* count, int out_ld, cudaStream_t stream) {
  if (P == 0) return 0;
  HH_REQUIRE(P > 0 && cost && offset && ld && nr && nc && row_ind && col_ind && count, "assign: null argument");
  HH_REQUIRE(max_dim >= 1 && max_dim <= MAXD, "assign: problems larger than 32 x 32 are not supported");
  HH_REQUIRE(out_ld >= 1 && out_ld <= MAXD, "assign: out_ld must be in 1..32");
  AssignArgs a{cost, offset, ld, nr, nc, row_valid, row_valid_ld, P, row_ind, col_ind, count, out_ld};
  assign_kernel<<<(P + 31) / 32, 32, 0, stream>>>(a);
  HH_CHECK_LAUNCH("assign_kernel");
  return 0;
}

int match_cost_class(const float* logits, int N, int ncls, const long long* ids, int M, float w, float* cost,
                     cudaStream_t stream) {
  if (N == 0 || M == 0) return 0;
  HH_REQUIRE(N > 0 && M > 0 && ncls > 0 && logits && ids && cost, "match_cost_class: bad argument");
  class_cost_kernel<<<N, 256, 0, stream>>>(logits, ncls, ids, M, w, cost);
  HH_CHECK_LAUNCH("class_cost_kernel");
  return 0;
}

}  // namespace hh
