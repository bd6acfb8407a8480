// Training-side scoring losses of the path, forward and backward:
//   * EgoNCE.forward (model/loss.py:15-70): log-softmax of sim / temperature over rows and over columns, averaged over
//     the positives of a (verb x noun + diagonal) x pad mask; rows of padded captions are dropped.
//   * WordContrastiveLoss.forward (model/loss.py:78-106): match <= 4 ground-truth nouns to the 12 object-query embeddings
//     by -cosine (hh_assign), then cross-entropy of the matched queries over the noun vocabulary at 1/temperature, with
//     near-synonyms of the target (cosine > noun_threshold) pushed to logit -1.
//   * the backward of sim_matrix (model/metric.py:363-375) that both feed.
// All are small (<= 2560 x 512 similarities, <= 256 x ~2000 logits): one CTA per row / column, fp32, deterministic
// (no atomics), no host synchronisation.
#include "hh_internal.h"
#include "hh_ptx.cuh"

namespace hh {

namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) t = fmaxf(t, red[i]);
  return t;
}

// ---------------------------------------------------------------------------------------------- sim_matrix backward
// inv[r] = 1 / max(|x_r|, eps); live[r] = |x_r| > eps (the clamp passes no gradient to the norm otherwise)
__global__ void row_inv_norm_kernel(const float* __restrict__ x, int rows, int d, float eps, float* __restrict__ inv,
                                    unsigned char* __restrict__ live) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int k = lane; k < d; k += 32) {
    const float v = x[static_cast<size_t>(row) * d + k];
    s += v * v;
  }
  s = warp_sum(s);
  if (lane == 0) {
    const float nrm = sqrtf(s);
    inv[row] = 1.f / fmaxf(nrm, eps);
    live[row] = nrm > eps;
  }
}

// da[i, :] for out = a^ b^T:  t = sum_j G[i,j] * gscale * b^_j ;  da_i = (t - a^_i (a^_i . t)) * inv_a[i]   (live rows)
// G element (i, j) at G[i * gs_i + j * gs_j] (so the same kernel gives db with the strides swapped).
__global__ void __launch_bounds__(256)
sim_bwd_kernel(const float* __restrict__ a, const float* __restrict__ inva, const unsigned char* __restrict__ livea,
               const float* __restrict__ b, const float* __restrict__ invb, const float* __restrict__ G, long long gs_i,
               long long gs_j, const float* __restrict__ gscale, float* __restrict__ da, int Nb, int d) {
  extern __shared__ float sm[];
  float* w = sm;            // [Nb] G[i, j] * invb[j]
  float* red = sm + Nb;     // [32]
  const int i = blockIdx.x;
  const float gsc = gscale ? *gscale : 1.f;
  for (int j = threadIdx.x; j < Nb; j += blockDim.x) w[j] = G[i * gs_i + j * gs_j] * invb[j] * gsc;
  __syncthreads();
  const float ia = inva[i];
  const bool live = livea[i];
  float dotp = 0.f;
  for (int k0 = 0; k0 < d; k0 += blockDim.x) {  // t_k kept in da (scratch) then corrected
    const int k = k0 + threadIdx.x;
    float t = 0.f;
    if (k < d) {
      for (int j = 0; j < Nb; ++j) t += w[j] * b[static_cast<size_t>(j) * d + k];
      da[static_cast<size_t>(i) * d + k] = t;
      dotp += t * a[static_cast<size_t>(i) * d + k] * ia;
    }
  }
  dotp = block_sum(dotp, red);
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    const float t = da[static_cast<size_t>(i) * d + k];
    const float ah = a[static_cast<size_t>(i) * d + k] * ia;
    da[static_cast<size_t>(i) * d + k] = live ? (t - ah * dotp) * ia : t * ia;
  }
}

// ---------------------------------------------------------------------------------------------- EgoNCE
struct NceArgs {
  const float* x;       // [N, M]
  const float* mask_v;  // [N / R, M] or null
  const float* mask_n;  // [N / R, M] or null
  const float* pad;     // [N, M] or null (single-positive branch)
  int N, M, R;
  float inv_t, thr;
  unsigned char* mask_bool;  // [N, M]
  unsigned char* keep;       // [N]
  float* row_lse;            // [N]
  float* row_cnt;            // [N]
  float* row_term;           // [N]
  float* col_lse;            // [M]
  float* col_cnt;            // [M]
  float* col_term;           // [M]
  float* stats;              // [2]: loss, kept rows
};

// one CTA per row: keep flag, mask_bool row, log-sum-exp of the row, positives' mean log-probability
__global__ void __launch_bounds__(128) nce_rows_kernel(NceArgs a) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  const float* xr = a.x + static_cast<size_t>(r) * a.M;
  const int v = r / a.R;  // video (column) this caption belongs to
  float mx = -INFINITY, padded = 0.f;
  for (int c = threadIdx.x; c < a.M; c += blockDim.x) {
    const bool p = a.pad ? a.pad[static_cast<size_t>(r) * a.M + c] != 0.f : true;
    if (!p) padded += 1.f;
    mx = fmaxf(mx, xr[c] * a.inv_t);
  }
  padded = block_sum(padded, red);
  const bool keep = padded == 0.f;  // masked_x.sum(-1) != -inf  <=>  no entry of the row was filled with -inf
  mx = block_max(mx, red);
  float l = 0.f, cnt = 0.f, s = 0.f;
  for (int c = threadIdx.x; c < a.M; c += blockDim.x) {
    const size_t mi = static_cast<size_t>(v) * a.M + c;
    float m = (c == v) ? 1.f : 0.f;
    if (a.mask_v && a.mask_n) m += a.mask_v[mi] * a.mask_n[mi];
    else if (a.mask_n) m += a.mask_n[mi];
    else if (a.mask_v) m += a.mask_v[mi];
    if (a.pad) m *= a.pad[static_cast<size_t>(r) * a.M + c];
    const bool mb = m > a.thr;
    a.mask_bool[static_cast<size_t>(r) * a.M + c] = mb;
    const float z = xr[c] * a.inv_t;
    l += expf(z - mx);
    if (mb) {
      cnt += 1.f;
      s += z;
    }
  }
  l = block_sum(l, red);
  cnt = block_sum(cnt, red);
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float lse = mx + logf(l);
    a.keep[r] = keep;
    a.row_lse[r] = lse;
    a.row_cnt[r] = cnt;
    a.row_term[r] = (s - cnt * lse) / cnt;  // 0/0 = nan when a kept row has no positive, as in the reference
  }
}

// one CTA per column over the kept rows
__global__ void __launch_bounds__(128) nce_cols_kernel(NceArgs a) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  float mx = -INFINITY;
  for (int r = threadIdx.x; r < a.N; r += blockDim.x)
    if (a.keep[r]) mx = fmaxf(mx, a.x[static_cast<size_t>(r) * a.M + c] * a.inv_t);
  mx = block_max(mx, red);
  float l = 0.f, cnt = 0.f, s = 0.f;
  for (int r = threadIdx.x; r < a.N; r += blockDim.x) {
    if (!a.keep[r]) continue;
    const float z = a.x[static_cast<size_t>(r) * a.M + c] * a.inv_t;
    l += expf(z - mx);
    if (a.mask_bool[static_cast<size_t>(r) * a.M + c]) {
      cnt += 1.f;
      s += z;
    }
  }
  l = block_sum(l, red);
  cnt = block_sum(cnt, red);
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float lse = mx + logf(l);
    a.col_lse[c] = lse;
    a.col_cnt[c] = cnt;
    a.col_term[c] = (s - cnt * lse) / cnt;
  }
}

__global__ void __launch_bounds__(256) nce_final_kernel(NceArgs a) {
  __shared__ float red[32];
  float si = 0.f, nk = 0.f, sj = 0.f;
  for (int r = threadIdx.x; r < a.N; r += blockDim.x)
    if (a.keep[r]) {
      si += a.row_term[r];
      nk += 1.f;
    }
  for (int c = threadIdx.x; c < a.M; c += blockDim.x) sj += a.col_term[c];
  si = block_sum(si, red);
  nk = block_sum(nk, red);
  sj = block_sum(sj, red);
  if (threadIdx.x == 0) {
    a.stats[0] = -(si / nk) - (sj / static_cast<float>(a.M));
    a.stats[1] = nk;
  }
}

// dx[r,c] = g * inv_t * [ -(mb/cnt_r - p_row)/n_kept - (mb/cnt_c - p_col)/M ]   (kept rows; 0 elsewhere)
__global__ void nce_bwd_kernel(const float* __restrict__ x, int N, int M, float inv_t,
                               const unsigned char* __restrict__ mask_bool, const unsigned char* __restrict__ keep,
                               const float* __restrict__ row_lse, const float* __restrict__ row_cnt,
                               const float* __restrict__ col_lse, const float* __restrict__ col_cnt,
                               const float* __restrict__ stats, const float* __restrict__ gloss, float* __restrict__ dx) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(N) * M) return;
  const int r = static_cast<int>(idx / M), c = static_cast<int>(idx - static_cast<size_t>(r) * M);
  if (!keep[r]) {
    dx[idx] = 0.f;
    return;
  }
  const float z = x[idx] * inv_t;
  const float mb = mask_bool[idx] ? 1.f : 0.f;
  const float gi = mb / row_cnt[r] - expf(z - row_lse[r]);
  const float gj = mb / col_cnt[c] - expf(z - col_lse[c]);
  dx[idx] = (*gloss) * inv_t * (-(gi / stats[1]) - gj / static_cast<float>(M));
}

// ---------------------------------------------------------------------------------------------- word loss
// cost[b, w, q] = -cos(noun[ind[b,w]], pred[b,q]);  valid[b, w] = ind[b,w] != 0      (model/loss.py:85-90)
__global__ void __launch_bounds__(128)
word_cost_kernel(const float* __restrict__ nouns, int V, int d, const float* __restrict__ pred, int Q,
                 const long long* __restrict__ inds, int Wm, float eps, float* __restrict__ cost,
                 unsigned char* __restrict__ valid, int* __restrict__ bad) {
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int pair = warp; pair < Wm * Q; pair += nw) {
    const int w = pair / Q, q = pair - w * Q;
    long long id = inds[static_cast<size_t>(b) * Wm + w];
    if (id < 0 || id >= V) {
      if (lane == 0) atomicExch(bad, 1);
      id = 0;
    }
    if (q == 0 && lane == 0) valid[static_cast<size_t>(b) * Wm + w] = id != 0;
    const float* g = nouns + static_cast<size_t>(id) * d;
    const float* p = pred + (static_cast<size_t>(b) * Q + q) * d;
    float dot = 0.f, ng = 0.f, np = 0.f;
    for (int k = lane; k < d; k += 32) {
      const float x = g[k], y = p[k];
      dot += x * y;
      ng += x * x;
      np += y * y;
    }
    dot = warp_sum(dot);
    ng = warp_sum(ng);
    np = warp_sum(np);
    if (lane == 0) cost[(static_cast<size_t>(b) * Wm + w) * Q + q] = -(dot / (fmaxf(sqrtf(ng), eps) * fmaxf(sqrtf(np), eps)));
  }
}

// slot (b, w): k = rank of w among the valid words of b -> q = col_ind[b, k];  sel[slot] = pred[b, q], gt[slot] =
// noun[ind], sel_row[slot] = b*Q + q (or -1), col_out[b, w] = q (or -1)
__global__ void __launch_bounds__(128)
word_select_kernel(const float* __restrict__ nouns, int d, const float* __restrict__ pred, int Q,
                   const long long* __restrict__ inds, const unsigned char* __restrict__ valid, int Wm,
                   const long long* __restrict__ col_ind, int K, float* __restrict__ sel, float* __restrict__ gtn,
                   long long* __restrict__ sel_row, long long* __restrict__ col_out) {
  const int slot = blockIdx.x;
  const int b = slot / Wm, w = slot - b * Wm;
  const bool ok = valid[slot];
  int k = 0;
  for (int i = 0; i < w; ++i) k += valid[static_cast<size_t>(b) * Wm + i];
  const long long q = (ok && k < K) ? col_ind[static_cast<size_t>(b) * K + k] : -1;
  if (threadIdx.x == 0) {
    sel_row[slot] = q >= 0 ? static_cast<long long>(b) * Q + q : -1;
    col_out[slot] = q;
  }
  const float* src = q >= 0 ? pred + (static_cast<size_t>(b) * Q + q) * d : nullptr;
  const float* g = q >= 0 ? nouns + static_cast<size_t>(inds[slot]) * d : nullptr;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    sel[static_cast<size_t>(slot) * d + c] = src ? src[c] : 0.f;
    gtn[static_cast<size_t>(slot) * d + c] = g ? g[c] : 0.f;
  }
}

// per slot: logits = where(noun_sim[gt] > thr (diag := 0), -1, sim_all) / t ; loss_slot = lse - logit[gt];
// dlogits (in place over sim_all) = (softmax - onehot) / t on unmasked entries (masked_fill blocks the gradient)
__global__ void __launch_bounds__(256)
word_ce_kernel(float* __restrict__ sim_all, const float* __restrict__ noun_sim, const long long* __restrict__ inds,
               const long long* __restrict__ sel_row, int V, float inv_t, float thr, float* __restrict__ slot_loss) {
  __shared__ float red[32];
  const int slot = blockIdx.x;
  float* z = sim_all + static_cast<size_t>(slot) * V;
  if (sel_row[slot] < 0) {
    for (int v = threadIdx.x; v < V; v += blockDim.x) z[v] = 0.f;
    if (threadIdx.x == 0) slot_loss[slot] = 0.f;
    return;
  }
  const long long gt = inds[slot];
  const float* ns = noun_sim + static_cast<size_t>(slot) * V;
  // noun_sim's diagonal is zeroed before the threshold test (loss.py:99-101)
  auto is_masked = [&](int v) { return (v == gt ? 0.f : ns[v]) > thr; };
  float mx = -INFINITY;
  for (int v = threadIdx.x; v < V; v += blockDim.x) mx = fmaxf(mx, (is_masked(v) ? -1.f : z[v]) * inv_t);
  mx = block_max(mx, red);
  float l = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) l += expf((is_masked(v) ? -1.f : z[v]) * inv_t - mx);
  l = block_sum(l, red);
  const float lse = mx + logf(l);
  const float zgt = (is_masked(static_cast<int>(gt)) ? -1.f : z[gt]) * inv_t;
  __syncthreads();
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const bool masked = is_masked(v);
    const float p = expf((masked ? -1.f : z[v]) * inv_t - lse);
    z[v] = masked ? 0.f : (p - (v == gt ? 1.f : 0.f)) * inv_t;
  }
  if (threadIdx.x == 0) slot_loss[slot] = lse - zgt;
}

// stats[0] = mean slot loss over valid slots, stats[1] = number of valid slots; dlogits *= 1 / n
__global__ void __launch_bounds__(256)
word_final_kernel(const float* __restrict__ slot_loss, const long long* __restrict__ sel_row, int S,
                  const int* __restrict__ bad, float* __restrict__ stats) {
  __shared__ float red[32];
  float s = 0.f, n = 0.f;
  for (int i = threadIdx.x; i < S; i += blockDim.x)
    if (sel_row[i] >= 0) {
      s += slot_loss[i];
      n += 1.f;
    }
  s = block_sum(s, red);
  n = block_sum(n, red);
  if (threadIdx.x == 0) {
    stats[0] = *bad ? NAN : s / n;  // a noun id outside the vocabulary poisons the loss (index_select would assert)
    stats[1] = n;
    stats[2] = 1.f / n;
  }
}

__global__ void scale_by_kernel(float* __restrict__ x, size_t n, const float* __restrict__ a, const float* __restrict__ b) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= (*a) * (b ? *b : 1.f);
}

// dpred[sel_row[slot], :] = dsel[slot, :]  (rows of pred that no noun selected get 0; the target is zeroed first)
__global__ void scatter_rows_kernel(const float* __restrict__ dsel, const long long* __restrict__ sel_row, int d,
                                    float* __restrict__ dpred) {
  const int slot = blockIdx.x;
  const long long row = sel_row[slot];
  if (row < 0) return;
  for (int c = threadIdx.x; c < d; c += blockDim.x) dpred[static_cast<size_t>(row) * d + c] = dsel[static_cast<size_t>(slot) * d + c];
}

// assignment problem table for the word loss: problem b = cost[b] (Wm x Q, rows filtered by valid[b])
__global__ void word_meta_kernel(long long* offset, int* ld, int* nr, int* nc, int B2, int Wm, int Q) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B2) return;
  offset[b] = static_cast<long long>(b) * Wm * Q;
  ld[b] = Q;
  nr[b] = Wm;
  nc[b] = Q;
}

inline size_t align_up(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

struct WordWs {
  float* cost; unsigned char* valid; long long* offset; int* ld; int* nr; int* nc; long long* ri; long long* ci;
  int* cnt; float* gtn; float* noun_sim; float* slot_loss; int* bad; size_t bytes;
};
WordWs word_ws(void* base, int V, int d, int B2, int Q, int Wm) {
  const size_t S = static_cast<size_t>(B2) * Wm;
  const int K = Wm < Q ? Wm : Q;
  WordWs w{};
  size_t o = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + o : nullptr;
    o += align_up(bytes);
    return p;
  };
  w.cost = static_cast<float*>(take(S * Q * 4));
  w.valid = static_cast<unsigned char*>(take(S));
  w.offset = static_cast<long long*>(take(static_cast<size_t>(B2) * 8));
  w.ld = static_cast<int*>(take(static_cast<size_t>(B2) * 4));
  w.nr = static_cast<int*>(take(static_cast<size_t>(B2) * 4));
  w.nc = static_cast<int*>(take(static_cast<size_t>(B2) * 4));
  w.ri = static_cast<long long*>(take(static_cast<size_t>(B2) * K * 8));
  w.ci = static_cast<long long*>(take(static_cast<size_t>(B2) * K * 8));
  w.cnt = static_cast<int*>(take(static_cast<size_t>(B2) * 4));
  w.gtn = static_cast<float*>(take(S * d * 4));
  w.noun_sim = static_cast<float*>(take(S * V * 4));
  w.slot_loss = static_cast<float*>(take(S * 4));
  w.bad = static_cast<int*>(take(4));
  w.bytes = o;
  return w;
}

}  // namespace

// ================================================================================================ host entry points
int sim_matrix_backward(const float* a, const float* b, const float* G, const float* gscale, float* da, float* db, int Na,
                        int Nb, int d, float eps, void* workspace, cudaStream_t s) {
  HH_REQUIRE(Na > 0 && Nb > 0 && d > 0, "sim_matrix_backward: empty problem");
  HH_REQUIRE(workspace != nullptr, "sim_matrix_backward: workspace");
  float* inva = static_cast<float*>(workspace);
  float* invb = inva + Na;
  unsigned char* livea = reinterpret_cast<unsigned char*>(invb + Nb);
  unsigned char* liveb = livea + Na;
  row_inv_norm_kernel<<<(Na + 7) / 8, 256, 0, s>>>(a, Na, d, eps, inva, livea);
  row_inv_norm_kernel<<<(Nb + 7) / 8, 256, 0, s>>>(b, Nb, d, eps, invb, liveb);
  HH_CHECK_LAUNCH("row_inv_norm_kernel");
  if (da) {
    const size_t smem = (static_cast<size_t>(Nb) + 32) * sizeof(float);
    HH_REQUIRE(smem <= 48 * 1024, "sim_matrix_backward: more than 12 k columns");
    sim_bwd_kernel<<<Na, 256, smem, s>>>(a, inva, livea, b, invb, G, Nb, 1, gscale, da, Nb, d);
  }
  if (db) {
    const size_t smem = (static_cast<size_t>(Na) + 32) * sizeof(float);
    HH_REQUIRE(smem <= 48 * 1024, "sim_matrix_backward: more than 12 k rows");
    sim_bwd_kernel<<<Nb, 256, smem, s>>>(b, invb, liveb, a, inva, G, 1, Nb, gscale, db, Na, d);
  }
  HH_CHECK_LAUNCH("sim_bwd_kernel");
  return 0;
}
size_t sim_matrix_backward_workspace_bytes(int Na, int Nb) { return static_cast<size_t>(Na + Nb) * 5 + 64; }

int egonce_forward(const float* x, int N, int M, const float* mask_v, const float* mask_n, int R, const float* pad,
                   float temperature, float vn_threshold, unsigned char* mask_bool, unsigned char* keep, float* saved,
                   cudaStream_t s) {
  HH_REQUIRE(N > 0 && M > 0 && R > 0 && N % R == 0, "egonce: bad shape");
  HH_REQUIRE(x && mask_bool && keep && saved, "egonce: null buffer");
  HH_REQUIRE(N / R == M, "egonce: the positive mask is diagonal over (rows / R) x columns; needs rows == R * columns");
  HH_REQUIRE(temperature > 0.f, "egonce: temperature");
  NceArgs a{};
  a.x = x; a.mask_v = mask_v; a.mask_n = mask_n; a.pad = pad; a.N = N; a.M = M; a.R = R;
  a.inv_t = 1.f / temperature; a.thr = vn_threshold; a.mask_bool = mask_bool; a.keep = keep;
  a.row_lse = saved; a.row_cnt = saved + N; a.row_term = saved + 2 * static_cast<size_t>(N);
  a.col_lse = saved + 3 * static_cast<size_t>(N); a.col_cnt = a.col_lse + M; a.col_term = a.col_cnt + M;
  a.stats = a.col_term + M;
  nce_rows_kernel<<<N, 128, 0, s>>>(a);
  nce_cols_kernel<<<M, 128, 0, s>>>(a);
  nce_final_kernel<<<1, 256, 0, s>>>(a);
  HH_CHECK_LAUNCH("egonce kernels");
  return 0;
}

int egonce_backward(const float* x, int N, int M, float temperature, const unsigned char* mask_bool,
                    const unsigned char* keep, const float* saved, const float* grad_loss, float* grad_x,
                    cudaStream_t s) {
  HH_REQUIRE(N > 0 && M > 0 && x && mask_bool && keep && saved && grad_loss && grad_x, "egonce backward: bad argument");
  const float* col_lse = saved + 3 * static_cast<size_t>(N);
  const size_t total = static_cast<size_t>(N) * M;
  nce_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(
      x, N, M, 1.f / temperature, mask_bool, keep, saved, saved + N, col_lse, col_lse + M, col_lse + 3 * static_cast<size_t>(M),
      grad_loss, grad_x);
  HH_CHECK_LAUNCH("nce_bwd_kernel");
  return 0;
}

size_t word_loss_workspace_bytes(int V, int d, int B2, int Q, int Wm) {
  const size_t fwd = word_ws(nullptr, V, d, B2, Q, Wm).bytes;
  const size_t S = static_cast<size_t>(B2) * Wm;
  const size_t bwd = align_up(sim_matrix_backward_workspace_bytes(static_cast<int>(S), V)) + align_up(S * d * 4);
  return fwd > bwd ? fwd : bwd;
}

int word_loss_forward(const float* nouns, int V, int d, const float* pred, int B2, int Q, const long long* gt_inds, int Wm,
                      float temperature, float noun_threshold, long long* col_ind, float* sel, long long* sel_row,
                      float* dlogits, float* stats, void* workspace, cudaStream_t s) {
  HH_REQUIRE(V > 0 && d > 0 && B2 > 0 && Q > 0 && Wm > 0, "word loss: empty problem");
  HH_REQUIRE(Q <= 32 && Wm <= 32, "word loss: at most 32 queries / nouns per clip");
  HH_REQUIRE(nouns && pred && gt_inds && col_ind && sel && sel_row && dlogits && stats && workspace, "word loss: null buffer");
  const WordWs w = word_ws(workspace, V, d, B2, Q, Wm);
  const int S = B2 * Wm;
  const int K = Wm < Q ? Wm : Q;
  const float eps = 1e-8f;  // sim_matrix default (model/metric.py:363)
  HH_CHECK_CUDA(cudaMemsetAsync(w.bad, 0, 4, s));
  word_cost_kernel<<<B2, 128, 0, s>>>(nouns, V, d, pred, Q, gt_inds, Wm, eps, w.cost, w.valid, w.bad);
  word_meta_kernel<<<(B2 + 127) / 128, 128, 0, s>>>(w.offset, w.ld, w.nr, w.nc, B2, Wm, Q);
  HH_CHECK_LAUNCH("word_cost_kernel");
  int rc = assign_lsa(w.cost, w.offset, w.ld, w.nr, w.nc, w.valid, Wm, B2, Wm > Q ? Wm : Q, w.ri, w.ci, w.cnt, K, s);
  if (rc) return rc;
  word_select_kernel<<<S, 128, 0, s>>>(nouns, d, pred, Q, gt_inds, w.valid, Wm, w.ci, K, sel, w.gtn, sel_row, col_ind);
  HH_CHECK_LAUNCH("word_select_kernel");
  rc = sim_matrix(sel, nouns, dlogits, S, V, d, eps, s);          // sim_all        (loss.py:96)
  if (rc) return rc;
  rc = sim_matrix(w.gtn, nouns, w.noun_sim, S, V, d, eps, s);     // noun_sim rows  (loss.py:98-101)
  if (rc) return rc;
  word_ce_kernel<<<S, 256, 0, s>>>(dlogits, w.noun_sim, gt_inds, sel_row, V, 1.f / temperature, noun_threshold, w.slot_loss);
  word_final_kernel<<<1, 256, 0, s>>>(w.slot_loss, sel_row, S, w.bad, stats);
  HH_CHECK_LAUNCH("word_ce_kernel");
  return 0;
}

int word_loss_backward(const float* nouns, int V, int d, int B2, int Q, int Wm, const float* sel, const long long* sel_row,
                       const float* dlogits, const float* stats, const float* grad_loss, float* d_pred, float* d_nouns,
                       void* workspace, cudaStream_t s) {
  HH_REQUIRE(nouns && sel && sel_row && dlogits && stats && grad_loss && workspace, "word loss backward: null buffer");
  const int S = B2 * Wm;
  char* base = static_cast<char*>(workspace);
  void* simws = base;
  float* dsel = reinterpret_cast<float*>(base + align_up(sim_matrix_backward_workspace_bytes(S, V)));
  // upstream gradient and the 1/n of the mean are applied inside the similarity backward (gscale = g * 1/n)
  float* gs = const_cast<float*>(stats) + 3;  // stats[3] is scratch
  HH_CHECK_CUDA(cudaMemcpyAsync(gs, grad_loss, 4, cudaMemcpyDeviceToDevice, s));
  scale_by_kernel<<<1, 32, 0, s>>>(gs, 1, stats + 2, nullptr);
  int rc = sim_matrix_backward(sel, nouns, dlogits, gs, d_pred ? dsel : nullptr, d_nouns, S, V, d, 1e-8f, simws, s);
  if (rc) return rc;
  if (d_pred) {
    HH_CHECK_CUDA(cudaMemsetAsync(d_pred, 0, static_cast<size_t>(B2) * Q * d * 4, s));
    scatter_rows_kernel<<<S, 128, 0, s>>>(dsel, sel_row, d, d_pred);
    HH_CHECK_LAUNCH("scatter_rows_kernel");
  }
  return 0;
}

}  // namespace hh
