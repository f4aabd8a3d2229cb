// Cell-clustering step on the GPU: the dense k x k affinity of a set of cells and its
// normalised graph Laplacian stay in HBM; the host eigensolver (scipy's LOBPCG, as in the reference) only sees
// "L @ X" products.
//
// Replaces, for one call of `create_affinity` + `normalizedMinCut`
// (linear_split.py:277-338, branching_split.py:175-232, utils.py:131-150, sklearn spectral_embedding):
//   d_mat   = squareform(pdist(features))                 kernel   = exp(-1 * d_mat / (0.2 * (max d - min d)))
//   c_mat   = copy_distance[np.ix_(cells, cells)]         r        = quantile(c_mat, radius)
//   copy_k  = exp(-1 * c_mat / (0.2 * (max c - min c)));  copy_k[c_mat > r] = 0   (branching: >= r)
//   W       = kernel * copy_k
//   L       = csgraph_laplacian(W, normed=True): zero diagonal, w = column sums (rows added in order), dd = sqrt(w)
//             (1 where w == 0), L = -(W / dd) / dd[:, None], unit diagonal
// with the reference's operation order (separately rounded divide / multiply), so that W, dd and L agree with numpy up
// to the last-ulp difference between CUDA's and glibc's exp().
#include <cub/device/device_radix_sort.cuh>
#include <math.h>
#include <stdlib.h>

#include <new>

#include "ph_common.cuh"

struct ph_affinity {
  int device;
  int32_t n, D;
  double* d_copy;    // [n*n] pairwise Euclidean distance of the embedding (phertilizer.py:286); NULL = not stored:
                     // entries are recomputed from d_emb on the fly (low-dimensional embeddings of many cells)
  double* d_emb;     // [n*D]
  double* d_W;       // [cap*cap] affinity, then Laplacian of the current cell set
  double* d_tmp;     // [cap*cap] sort scratch (quantile), allocated on first use
  double* d_feat;    // [cap*8]
  double* d_dd;      // [cap]
  double* d_x;       // [cap*8] matmat operand / result
  double* d_y;
  int32_t* d_cells;  // [cap]
  unsigned long long* d_red;  // [4]: max d bits, max c bits, nan flag, spare
  void* d_sort_tmp;
  size_t sort_tmp_bytes;
  double* h_pin;     // [cap*16] pinned staging
  int32_t cap, k;    // capacity (= n), current cell-set size
  int32_t cap_rows;  // rows of the k x k matrices this context can hold (= cap, or the row block of a sharded context)
  int32_t row_lo, row_hi, use_copy, rows_mode;  // row block held after ph_affinity_build_rows (cells sharded across GPUs)
  double* d_ddf;     // [cap] full sqrt-degree vector (row-block mode)
  double* d_bv;      // [kSlots][cap * kBW] block vectors of the device-resident eigensolver (ph_affinity_bv_*)
  double* d_gram;    // [kGramChunks + 1][kGramMax] partial / final Gram entries
  cudaStream_t st;
};

namespace {

constexpr int kMaxF = 8, kMaxCols = 8;
constexpr int kSlots = 10, kBW = 3;              // block-vector slots of kBW columns each (LOBPCG block size 3)
constexpr int kGramChunks = 148, kGramMax = 9 * 18;  // Gram: up to 3 x 6 slots, reduced in up to 148 row chunks (one wave)

// d(i,j) = sqrt(sum_f (x_if - x_jf)^2), multiply and add rounded separately like scipy's C loop
__device__ __forceinline__ double feat_dist(const double* __restrict__ f, int F, int i, int j) {
  double s = 0.0;
  for (int q = 0; q < F; ++q) {
    const double t = __dsub_rn(f[(size_t)i * F + q], f[(size_t)j * F + q]);
    s = __dadd_rn(s, __dmul_rn(t, t));
  }
  return sqrt(s);
}

// copy_distance[ci][cj]: from the stored matrix, or recomputed with the arithmetic of ph_pairwise_dist (k_pdist)
struct CopySrc {
  const double* copy;
  const double* emb;
  int n, D;
};
__device__ __forceinline__ double copy_at(const CopySrc& s, int ci, int cj) {
  if (s.copy) return s.copy[(size_t)ci * s.n + cj];
  double acc = 0.0;
  for (int q = 0; q < s.D; ++q) {
    const double t = __dsub_rn(s.emb[(size_t)ci * s.D + q], s.emb[(size_t)cj * s.D + q]);
    acc = __dadd_rn(acc, __dmul_rn(t, t));
  }
  return sqrt(acc);
}

// pass 1: max of d and of c over all pairs (their minima are the zero diagonals), NaN flag.  d and c are symmetric bit
// for bit ((a-b)^2 == (b-a)^2), so only tiles on or above the diagonal are evaluated.  32 x 32 tiles, 32 x 8 threads.
__global__ void __launch_bounds__(256) k_aff_stats(const double* __restrict__ feat, int F, const CopySrc src,
                                                   const int32_t* __restrict__ cells, int k, int use_copy,
                                                   unsigned long long* red) {
  if (blockIdx.y > blockIdx.x) return;  // tile row > tile column: the mirror image is evaluated instead
  const int j = blockIdx.x * 32 + threadIdx.x;
  double md = 0.0, mc = 0.0;
  int nan = 0;
  if (j < k) {
    const int cj = use_copy ? cells[j] : 0;
    for (int u = threadIdx.y; u < 32; u += 8) {
      const int i = blockIdx.y * 32 + u;
      if (i >= k) break;
      const double d = feat_dist(feat, F, i, j);
      if (d != d) nan = 1;
      md = fmax(md, d);
      if (use_copy) {
        const double c = copy_at(src, cells[i], cj);
        if (c != c) nan = 1;
        mc = fmax(mc, c);
      }
    }
  }
  // non-negative doubles order like their bit patterns
  for (int o = 16; o > 0; o >>= 1) {
    md = fmax(md, __shfl_xor_sync(0xffffffffu, md, o));
    mc = fmax(mc, __shfl_xor_sync(0xffffffffu, mc, o));
    nan |= __shfl_xor_sync(0xffffffffu, nan, o);
  }
  if (threadIdx.x == 0) {
    if (md > 0.0) atomicMax(red + 0, (unsigned long long)__double_as_longlong(md));
    if (mc > 0.0) atomicMax(red + 1, (unsigned long long)__double_as_longlong(mc));
    if (nan) atomicMax(red + 2, 1ull);
  }
}

__global__ void k_aff_gather_c(const CopySrc src, const int32_t* __restrict__ cells, int k, double* __restrict__ out) {
  const size_t total = (size_t)k * k;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(p / k), j = (int)(p - (size_t)i * k);
    out[p] = copy_at(src, cells[i], cells[j]);
  }
}

// pass 2: W = exp(-1*d/kw) * copy_kernel.  W is symmetric bit for bit (d and c are), so a 32 x 32 tile on or above the
// diagonal is evaluated once and stored twice -- directly, and transposed through shared memory (both coalesced): half the
// exp / sqrt / div work of the k^2 pairs.
__global__ void __launch_bounds__(256) k_aff_build(const double* __restrict__ feat, int F, const CopySrc src,
                                                   const int32_t* __restrict__ cells, int k, int use_copy, double kw,
                                                   double kwc, double r, int strict, double* __restrict__ W,
                                                   unsigned long long* red) {
  if (blockIdx.y > blockIdx.x) return;
  __shared__ double tile[32][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  const int cj = (use_copy && j < k) ? cells[j] : 0;
  int nan = 0;
  for (int u = threadIdx.y; u < 32; u += 8) {
    const int i = blockIdx.y * 32 + u;
    double w = 0.0;
    if (i < k && j < k) {
      const double d = feat_dist(feat, F, i, j);
      w = exp(__ddiv_rn(__dmul_rn(-1.0, d), kw));
      if (use_copy) {
        const double c = copy_at(src, cells[i], cj);
        double ck = exp(__ddiv_rn(__dmul_rn(-1.0, c), kwc));
        if (strict ? (c >= r) : (c > r)) ck = 0.0;
        w = __dmul_rn(w, ck);
      }
      if (w != w) nan = 1;
      W[(size_t)i * k + j] = w;
    }
    tile[u][threadIdx.x] = w;
  }
  if (blockIdx.y != blockIdx.x) {
    __syncthreads();
    const int jj = blockIdx.y * 32 + threadIdx.x;  // column of the mirrored tile = row of this one
    for (int u = threadIdx.y; u < 32; u += 8) {
      const int ii = blockIdx.x * 32 + u;          // row of the mirrored tile = column of this one
      if (ii < k && jj < k) W[(size_t)ii * k + jj] = tile[threadIdx.x][u];
    }
  }
  if (__any_sync(0xffffffffu, nan) && threadIdx.x == 0) atomicMax(red + 2, 1ull);
}

// w_j = sum_i W_ij with a zero diagonal, rows added in order (numpy's axis-0 reduction); dd = sqrt(w), 1 if isolated
__global__ void k_aff_degree(const double* __restrict__ W, int k, double* __restrict__ dd) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  double s = 0.0;
  int i = 0;
  for (; i + 8 <= k; i += 8) {  // eight rows in flight, added in row order
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = W[(size_t)(i + u) * k + j];
#pragma unroll
    for (int u = 0; u < 8; ++u) s = __dadd_rn(s, (i + u) == j ? 0.0 : v[u]);
  }
  for (; i < k; ++i) s = __dadd_rn(s, i == j ? 0.0 : W[(size_t)i * k + j]);
  dd[j] = (s == 0.0) ? 1.0 : sqrt(s);
}

// L = -((W / dd[None, :]) / dd[:, None]), unit diagonal (csgraph_laplacian + _set_diag); one row per block iteration
__global__ void k_aff_laplacian(double* __restrict__ W, int k, const double* __restrict__ dd) {
  for (int i = blockIdx.x; i < k; i += gridDim.x) {
    const double di = dd[i];
    double* row = W + (size_t)i * k;
    for (int j = threadIdx.x; j < k; j += blockDim.x)
      row[j] = (i == j) ? 1.0 : __dmul_rn(__ddiv_rn(__ddiv_rn(row[j], dd[j]), di), -1.0);
  }
}

// Y[i, :] = L[i, :] @ X  (X: k x c row-major, c <= 8); one warp per row, lanes stride the row, fixed reduction order
template <int C>
__global__ void k_aff_matmat(const double* __restrict__ L, int k, const double* __restrict__ X, double* __restrict__ Y, int nrows) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nrows) return;
  const double* row = L + (size_t)warp * k;
  double acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.0;
  for (int j = lane; j < k; j += 32) {
    const double l = row[j];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = fma(l, X[(size_t)j * C + c], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    if (lane == 0) Y[(size_t)warp * C + c] = acc[c];
  }
}

// ---- row-block variants: the rank of a multi-GPU run holds rows [row_lo, row_hi) of the k x k matrices (all columns).
// Every entry is evaluated directly (no mirrored tiles); d, c and hence W are symmetric bit for bit, so the entries -- and
// the row sums below -- are the bits the single-GPU kernels produce.
__global__ void __launch_bounds__(256) k_aff_stats_rect(const double* __restrict__ feat, int F, const CopySrc src,
                                                        const int32_t* __restrict__ cells, int k, int use_copy, int row_lo,
                                                        int row_hi, unsigned long long* red) {
  const int j = blockIdx.x * 32 + threadIdx.x;
  double md = 0.0, mc = 0.0;
  int nan = 0;
  if (j < k) {
    const int cj = use_copy ? cells[j] : 0;
    for (int u = threadIdx.y; u < 32; u += 8) {
      const int i = row_lo + blockIdx.y * 32 + u;
      if (i >= row_hi) break;
      const double d = feat_dist(feat, F, i, j);
      if (d != d) nan = 1;
      md = fmax(md, d);
      if (use_copy) {
        const double c = copy_at(src, cells[i], cj);
        if (c != c) nan = 1;
        mc = fmax(mc, c);
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    md = fmax(md, __shfl_xor_sync(0xffffffffu, md, o));
    mc = fmax(mc, __shfl_xor_sync(0xffffffffu, mc, o));
    nan |= __shfl_xor_sync(0xffffffffu, nan, o);
  }
  if (threadIdx.x == 0) {
    if (md > 0.0) atomicMax(red + 0, (unsigned long long)__double_as_longlong(md));
    if (mc > 0.0) atomicMax(red + 1, (unsigned long long)__double_as_longlong(mc));
    if (nan) atomicMax(red + 2, 1ull);
  }
}

__global__ void __launch_bounds__(256) k_aff_build_rect(const double* __restrict__ feat, int F, const CopySrc src,
                                                        const int32_t* __restrict__ cells, int k, int use_copy, double kw,
                                                        double kwc, double r, int strict, int row_lo, int row_hi,
                                                        double* __restrict__ W, unsigned long long* red) {
  const int j = blockIdx.x * 32 + threadIdx.x;
  const int cj = (use_copy && j < k) ? cells[j] : 0;
  int nan = 0;
  for (int u = threadIdx.y; u < 32; u += 8) {
    const int i = row_lo + blockIdx.y * 32 + u;
    if (i < row_hi && j < k) {
      const double d = feat_dist(feat, F, i, j);
      double w = exp(__ddiv_rn(__dmul_rn(-1.0, d), kw));
      if (use_copy) {
        const double c = copy_at(src, cells[i], cj);
        double ck = exp(__ddiv_rn(__dmul_rn(-1.0, c), kwc));
        if (strict ? (c >= r) : (c > r)) ck = 0.0;
        w = __dmul_rn(w, ck);
      }
      if (w != w) nan = 1;
      W[(size_t)(i - row_lo) * k + j] = w;
    }
  }
  if (__any_sync(0xffffffffu, nan) && threadIdx.x == 0) atomicMax(red + 2, 1ull);
}

// w_i = sum_j W_ij with a zero diagonal, columns added in order: W is symmetric, so this is the sequence of additions
// k_aff_degree performs down column i.  One thread per row, eight loads in flight.
__global__ void k_aff_rowsum(const double* __restrict__ W, int nr, int k, int row_lo, double* __restrict__ dd) {
  const int rr = blockIdx.x * blockDim.x + threadIdx.x;
  if (rr >= nr) return;
  const int i = row_lo + rr;
  const double* row = W + (size_t)rr * k;
  double s = 0.0;
  int j = 0;
  for (; j + 8 <= k; j += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = row[j + u];
#pragma unroll
    for (int u = 0; u < 8; ++u) s = __dadd_rn(s, (j + u) == i ? 0.0 : v[u]);
  }
  for (; j < k; ++j) s = __dadd_rn(s, j == i ? 0.0 : row[j]);
  dd[rr] = (s == 0.0) ? 1.0 : sqrt(s);
}

__global__ void k_aff_laplacian_rect(double* __restrict__ W, int nr, int k, int row_lo, const double* __restrict__ dd) {
  for (int rr = blockIdx.x; rr < nr; rr += gridDim.x) {
    const int i = row_lo + rr;
    const double di = dd[i];
    double* row = W + (size_t)rr * k;
    for (int j = threadIdx.x; j < k; j += blockDim.x)
      row[j] = (i == j) ? 1.0 : __dmul_rn(__ddiv_rn(__ddiv_rn(row[j], dd[j]), di), -1.0);
  }
}

// ---- block-vector kernels of the device-resident eigensolver.  A slot is a k x 3 row-major array.
struct BvList {
  const double* p[6];
  int n;
};
// G[(a, i), (b, j)] = sum_r L_a[r, i] * R_b[r, j]: per chunk of rows a fixed-order partial; the CTA that finishes last
// adds the chunks in order (a function of k only: deterministic) and stores the result straight into page-locked host
// memory, so one launch and one stream synchronisation serve a Gram matrix.
__global__ void __launch_bounds__(256) k_bv_gram(const BvList L, const BvList R, int k, double* __restrict__ part,
                                                 unsigned* __restrict__ ticket, double* __restrict__ out_host) {
  const int nl = L.n * kBW, nr = R.n * kBW, ne = nl * nr;
  const int rows_per = (k + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rows_per, r1 = min(k, r0 + rows_per);
  __shared__ double sh[8][kGramMax];
  __shared__ bool last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // every warp walks rows warp, warp + 8, ... of the chunk; lane = entry (i, j) strided over the nl * nr entries
  for (int e = lane; e < ne; e += 32) {
    const int i = e / nr, j = e - i * nr;
    const double* lp = L.p[i / kBW] + (i % kBW);
    const double* rp = R.p[j / kBW] + (j % kBW);
    double acc = 0.0;
    for (int r = r0 + warp; r < r1; r += 8) acc = fma(lp[(size_t)r * kBW], rp[(size_t)r * kBW], acc);
    sh[warp][e] = acc;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < ne; e += blockDim.x) {
    double t = sh[0][e];
#pragma unroll
    for (int w = 1; w < 8; ++w) t += sh[w][e];
    part[(size_t)blockIdx.x * kGramMax + e] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int e = threadIdx.x; e < ne; e += blockDim.x) {
    double t = 0.0;
    int c = 0;
    for (; c + 8 <= (int)gridDim.x; c += 8) {   // eight independent loads in flight, added in chunk order
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(part + (size_t)(c + u) * kGramMax + e);
#pragma unroll
      for (int u = 0; u < 8; ++u) t += v[u];
    }
    for (; c < (int)gridDim.x; ++c) t += __ldcg(part + (size_t)c * kGramMax + e);
    out_host[e] = t;
  }
  if (threadIdx.x == 0) *ticket = 0;
}
// dst[r, :] = sum_t U_t[r, :] @ C_t   (C_t: 3 x 3 row-major, up to 4 terms; products and sums rounded separately, terms in
// order: R = AX * I + X * (-diag(lambda)) is then the reference's  AX - X * lambda  bit for bit).  dst may alias a source.
struct BvTerms {
  const double* u[4];
  double c[4][kBW * kBW];
  int n;
};
__global__ void k_bv_combine(const BvTerms T, int k, double* __restrict__ dst) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= k) return;
  double o[kBW] = {0.0, 0.0, 0.0};
  bool first[kBW] = {true, true, true};
  for (int t = 0; t < T.n; ++t) {
    double x[kBW];
#pragma unroll
    for (int i = 0; i < kBW; ++i) x[i] = T.u[t][(size_t)r * kBW + i];
#pragma unroll
    for (int i = 0; i < kBW; ++i)
#pragma unroll
      for (int j = 0; j < kBW; ++j) {
        const double c = T.c[t][i * kBW + j];
        if (c != 0.0) {   // (selection / padding coefficients: an exact copy, and nothing added for the zeros)
          const double pr = c == 1.0 ? x[i] : __dmul_rn(x[i], c);
          o[j] = first[j] ? pr : __dadd_rn(o[j], pr);
          first[j] = false;
        }
      }
  }
#pragma unroll
  for (int j = 0; j < kBW; ++j) dst[(size_t)r * kBW + j] = o[j];
}

// rows == NULL: rows row0, row0+1, ...
__global__ void k_gather_rows(const CopySrc src, const int32_t* __restrict__ rows, int row0, int nr,
                              double* __restrict__ out) {
  const int n = src.n;
  const size_t total = (size_t)nr * n;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(p / n), j = (int)(p - (size_t)i * n);
    out[p] = copy_at(src, rows ? rows[i] : row0 + i, j);
  }
}

// nearest donor of every listed cell (impute_mut_features, utils.py:87-129): argmin over donors j != c of
// copy_distance[c][j]; ties[r] = number of donors at that minimum (the caller resolves real ties the reference's way)
__global__ void __launch_bounds__(256) k_nearest(const CopySrc src, const int32_t* __restrict__ rows, int nr,
                                                 const uint8_t* __restrict__ donor, int32_t* __restrict__ out_idx,
                                                 int32_t* __restrict__ out_ties) {
  const int r = blockIdx.x;
  if (r >= nr) return;
  const int c = rows[r];
  double best = INFINITY;
  int bj = -1, cnt = 0;
  for (int j = threadIdx.x; j < src.n; j += blockDim.x) {
    if (!donor[j] || j == c) continue;
    const double d = copy_at(src, c, j);
    if (d < best) {
      best = d;
      bj = j;
      cnt = 1;
    } else if (d == best) {
      ++cnt;
    }
  }
  __shared__ double sd[256];
  __shared__ int sj[256], sc[256];
  sd[threadIdx.x] = best; sj[threadIdx.x] = bj; sc[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      const double d2 = sd[threadIdx.x + o];
      const int j2 = sj[threadIdx.x + o], c2 = sc[threadIdx.x + o];
      if (j2 >= 0) {
        if (sj[threadIdx.x] < 0 || d2 < sd[threadIdx.x]) {
          sd[threadIdx.x] = d2; sj[threadIdx.x] = j2; sc[threadIdx.x] = c2;
        } else if (d2 == sd[threadIdx.x]) {
          sc[threadIdx.x] += c2;
          if (j2 < sj[threadIdx.x]) sj[threadIdx.x] = j2;
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out_idx[r] = sj[0];
    out_ties[r] = sc[0];
  }
}

}  // namespace

extern "C" int ph_affinity_nearest(ph_affinity* a, const int32_t* rows, int32_t n_rows, const uint8_t* donor_mask,
                                   int32_t* out_idx, int32_t* out_ties) {
  if (!a || !rows || !donor_mask || !out_idx || !out_ties) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (n_rows <= 0) return PH_OK;
  if (n_rows > a->n) PH_FAIL(PH_ERR_ARG, "n_rows=%d exceeds the number of cells", n_rows);
  for (int i = 0; i < n_rows; ++i)
    if (rows[i] < 0 || rows[i] >= a->n) PH_FAIL(PH_ERR_ARG, "row %d out of range", rows[i]);
  PH_CUDA(cudaSetDevice(a->device));
  cudaStream_t st = a->st;
  // scratch: rows -> d_cells, donor mask + outputs -> the x / y vectors (n * 8 doubles each)
  uint8_t* d_mask = reinterpret_cast<uint8_t*>(a->d_x);
  int32_t* d_idx = reinterpret_cast<int32_t*>(a->d_y);
  int32_t* d_ties = d_idx + a->n;
  PH_CUDA(cudaMemcpyAsync(a->d_cells, rows, (size_t)n_rows * 4, cudaMemcpyHostToDevice, st));
  PH_CUDA(cudaMemcpyAsync(d_mask, donor_mask, (size_t)a->n, cudaMemcpyHostToDevice, st));
  const CopySrc src{a->d_copy, a->d_emb, a->n, a->D};
  k_nearest<<<n_rows, 256, 0, st>>>(src, a->d_cells, n_rows, d_mask, d_idx, d_ties);
  ph_count_launch();
  PH_CUDA(cudaGetLastError());
  PH_CUDA(cudaMemcpyAsync(out_idx, d_idx, (size_t)n_rows * 4, cudaMemcpyDeviceToHost, st));
  PH_CUDA(cudaMemcpyAsync(out_ties, d_ties, (size_t)n_rows * 4, cudaMemcpyDeviceToHost, st));
  PH_CUDA(cudaStreamSynchronize(st));
  return PH_OK;
}

extern "C" void ph_affinity_destroy(ph_affinity* a) {
  if (!a) return;
  cudaSetDevice(a->device);
  cudaFree(a->d_copy); cudaFree(a->d_emb); cudaFree(a->d_W); cudaFree(a->d_tmp); cudaFree(a->d_feat); cudaFree(a->d_dd); cudaFree(a->d_x);
  cudaFree(a->d_y); cudaFree(a->d_cells); cudaFree(a->d_red); cudaFree(a->d_sort_tmp); cudaFree(a->d_ddf); cudaFree(a->d_bv); cudaFree(a->d_gram);
  if (a->h_pin) cudaFreeHost(a->h_pin);
  if (a->st) cudaStreamDestroy(a->st);
  delete a;
}

namespace {
int affinity_create(const double* embedding, int32_t n, int32_t D, int device, int32_t max_rows, ph_affinity** out) {
  if (!embedding || !out) PH_FAIL(PH_ERR_ARG, "null pointer");
  *out = nullptr;
  if (n < 1 || D < 1 || max_rows < 1 || max_rows > n) PH_FAIL(PH_ERR_ARG, "bad shape n=%d D=%d rows=%d", n, D, max_rows);
  int ndev = 0;
  PH_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) PH_FAIL(PH_ERR_CUDA, "device %d not available (%d CUDA devices)", device, ndev);
  PH_CUDA(cudaSetDevice(device));
  size_t free_b = 0, total_b = 0;
  PH_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const size_t nn = (size_t)n * n * sizeof(double);
  const size_t wn = (size_t)max_rows * n * sizeof(double);  // the affinity / Laplacian rows this context holds
  const size_t margin = (size_t)1 << 30;
  const bool full = max_rows == n;
  // the distance matrix is kept next to the affinity when both fit (and a third n x n block remains for the quantile
  // sort); otherwise a low-dimensional embedding is re-evaluated entry by entry
  bool store = (full ? 3 * nn : nn + 2 * wn) + margin <= free_b;
  if (const char* e = getenv("PH_AFFINITY_STORE")) store = atoi(e) != 0;
  if (!store && D > 16 && nn + wn + margin <= free_b) store = true;
  if ((store ? nn : 0) + wn + margin > free_b)
    PH_FAIL(PH_ERR_UNSUPPORTED, "%d cells need %.1f GB for the dense affinity matrix (%.1f GB free)", n,
            ((store ? nn : 0) + wn) / 1e9, free_b / 1e9);
  if (!store && D > 16)
    PH_FAIL(PH_ERR_UNSUPPORTED, "%d cells x %d embedding dimensions: the distance matrix neither fits nor can be recomputed cheaply", n, D);
  ph_affinity* a = new (std::nothrow) ph_affinity();
  if (!a) PH_FAIL(PH_ERR_ARG, "out of host memory");
  memset(a, 0, sizeof(*a));
  a->device = device;
  a->n = n;
  a->D = D;
  a->cap = n;
  a->cap_rows = max_rows;
  auto body = [&]() -> int {
    PH_CUDA(cudaStreamCreateWithFlags(&a->st, cudaStreamNonBlocking));
    if (store) PH_CUDA(cudaMalloc(&a->d_copy, nn));
    PH_CUDA(cudaMalloc(&a->d_W, wn));
    PH_CUDA(cudaMalloc(&a->d_feat, (size_t)n * kMaxF * 8));
    PH_CUDA(cudaMalloc(&a->d_dd, (size_t)n * 8));
    PH_CUDA(cudaMalloc(&a->d_ddf, (size_t)n * 8));
    PH_CUDA(cudaMalloc(&a->d_bv, (size_t)kSlots * n * kBW * 8));
    PH_CUDA(cudaMalloc(&a->d_gram, (size_t)(kGramChunks + 1) * kGramMax * 8));
    PH_CUDA(cudaMemsetAsync(a->d_gram, 0, (size_t)kGramMax * 8, a->st));   // [0]: the completion ticket of k_bv_gram
    PH_CUDA(cudaMalloc(&a->d_x, (size_t)n * kMaxCols * 8));
    PH_CUDA(cudaMalloc(&a->d_y, (size_t)n * kMaxCols * 8));
    PH_CUDA(cudaMalloc(&a->d_cells, (size_t)n * 4));
    PH_CUDA(cudaMalloc(&a->d_red, 4 * 8));
    PH_CUDA(cudaMallocHost(&a->h_pin, (size_t)n * 16 * 8));
    PH_CUDA(cudaMalloc(&a->d_emb, (size_t)n * D * 8));
    PH_CUDA(cudaMemcpyAsync(a->d_emb, embedding, (size_t)n * D * 8, cudaMemcpyHostToDevice, a->st));
    if (store) {
      int rc = ph_pairwise_dist(a->d_emb, n, D, 0, n, a->d_copy, device, 1, a->st);
      if (rc != PH_OK) return rc;
    }
    PH_CUDA(cudaStreamSynchronize(a->st));
    return PH_OK;
  };
  int rc = body();
  if (rc != PH_OK) {
    ph_affinity_destroy(a);
    return rc;
  }
  *out = a;
  return PH_OK;
}
}  // namespace

extern "C" int ph_affinity_create(const double* embedding, int32_t n, int32_t D, int device, ph_affinity** out) {
  return affinity_create(embedding, n, D, device, n, out);
}

// a context that holds at most `max_rows` rows of the k x k matrices: the row block of one rank when the clustering is
// sharded across GPUs (ph_affinity_*_rows).  Cell sets small enough for k * k entries still go through ph_affinity_build.
extern "C" int ph_affinity_create_rows(const double* embedding, int32_t n, int32_t D, int device, int32_t max_rows,
                                       ph_affinity** out) {
  return affinity_create(embedding, n, D, device, max_rows, out);
}

// rows of the copy-distance matrix (impute_mut_features reads a few rows; `copy_distance` itself is an output only for
// small inputs / tests)
extern "C" int ph_affinity_rows(ph_affinity* a, const int32_t* rows, int32_t n_rows, double* out) {
  if (!a || !out || (!rows && n_rows != a->n)) PH_FAIL(PH_ERR_ARG, "null pointer");
  PH_CUDA(cudaSetDevice(a->device));
  const CopySrc src{a->d_copy, a->d_emb, a->n, a->D};
  if (rows)
    for (int i = 0; i < n_rows; ++i)
      if (rows[i] < 0 || rows[i] >= a->n) PH_FAIL(PH_ERR_ARG, "row %d out of range", rows[i]);
  if (n_rows <= 0) return PH_OK;
  // staged through the W buffer in slabs of at most n rows
  int done = 0;
  while (done < n_rows) {
    const int nr = n_rows - done < a->cap_rows ? n_rows - done : a->cap_rows;
    if (rows) PH_CUDA(cudaMemcpyAsync(a->d_cells, rows + done, (size_t)nr * 4, cudaMemcpyHostToDevice, a->st));
    k_gather_rows<<<1184, 256, 0, a->st>>>(src, rows ? a->d_cells : nullptr, done, nr, a->d_W);
    ph_count_launch();
    PH_CUDA(cudaMemcpyAsync(out + (size_t)done * a->n, a->d_W, (size_t)nr * a->n * 8, cudaMemcpyDeviceToHost, a->st));
    PH_CUDA(cudaStreamSynchronize(a->st));
    done += nr;
  }
  a->k = 0;  // the W buffer no longer holds a Laplacian
  a->rows_mode = 0;
  return PH_OK;
}

/*
 * Affinity + normalised Laplacian of `cells` (k of them) from their features [k*F]; see the header for the formulas.
 *   stats_out[8] = kw, kwc, r, has_nan (0/1), max d, max c, and the two order statistics r was interpolated from
 * When has_nan is set the Laplacian is not built (the reference terminates the clustering, linear_split.py:364-369).
 */
extern "C" int ph_affinity_build(ph_affinity* a, const int32_t* cells, int32_t k, const double* feats, int32_t F,
                                 int use_copy_kernel, double radius, int strict, double* stats_out, double* dd_out) {
  if (!a || !cells || !feats || !stats_out) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (k < 1 || k > a->cap || F < 1 || F > kMaxF) PH_FAIL(PH_ERR_ARG, "bad shape k=%d F=%d", k, F);
  if ((size_t)k * k > (size_t)a->cap_rows * a->n)
    PH_FAIL(PH_ERR_UNSUPPORTED, "%d x %d affinity exceeds this context's %d rows: use the row-block calls", k, k, a->cap_rows);
  for (int i = 0; i < k; ++i)
    if (cells[i] < 0 || cells[i] >= a->n) PH_FAIL(PH_ERR_ARG, "cells[%d]=%d out of range", i, cells[i]);
  PH_CUDA(cudaSetDevice(a->device));
  cudaStream_t st = a->st;
  a->k = 0;
  a->row_lo = a->row_hi = 0;
  a->rows_mode = 0;
  memcpy(a->h_pin, feats, (size_t)k * F * 8);
  memcpy(a->h_pin + (size_t)k * F, cells, (size_t)k * 4);
  PH_CUDA(cudaMemcpyAsync(a->d_feat, a->h_pin, (size_t)k * F * 8, cudaMemcpyHostToDevice, st));
  PH_CUDA(cudaMemcpyAsync(a->d_cells, a->h_pin + (size_t)k * F, (size_t)k * 4, cudaMemcpyHostToDevice, st));
  PH_CUDA(cudaMemsetAsync(a->d_red, 0, 4 * 8, st));
  const int blocks = 1184;
  const CopySrc src{a->d_copy, a->d_emb, a->n, a->D};
  const dim3 tgrid((k + 31) / 32, (k + 31) / 32), tblock(32, 8);  // tiles on or above the diagonal do the work
  k_aff_stats<<<tgrid, tblock, 0, st>>>(a->d_feat, F, src, a->d_cells, k, use_copy_kernel, a->d_red);
  ph_count_launch();
  unsigned long long red[4];
  PH_CUDA(cudaMemcpyAsync(red, a->d_red, sizeof(red), cudaMemcpyDeviceToHost, st));
  PH_CUDA(cudaStreamSynchronize(st));
  double maxd, maxc;
  memcpy(&maxd, &red[0], 8);
  memcpy(&maxc, &red[1], 8);
  // utils.py:75-84  kernel width = 0.2 * (max - min), min = 0 (the diagonal)
  const double kw = 0.2 * (maxd - 0.0), kwc = 0.2 * (maxc - 0.0);
  double r = maxc, q_lo = maxc, q_hi = maxc;
  if (use_copy_kernel && radius < 1.0) {
    // np.quantile(c_mat, radius), method "linear": sort the k*k distances, interpolate between two order statistics
    const size_t N = (size_t)k * k;
    if (N >= ((size_t)1 << 31)) PH_FAIL(PH_ERR_UNSUPPORTED, "quantile over %zu distances", N);
    if (!a->d_tmp) PH_CUDA(cudaMalloc(&a->d_tmp, (size_t)a->cap_rows * a->cap * 8));
    size_t need = 0;
    cub::DoubleBuffer<double> db(a->d_W, a->d_tmp);
    PH_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, need, db, (int)N, 0, 64, st));
    if (need > a->sort_tmp_bytes) {
      cudaFree(a->d_sort_tmp);
      a->d_sort_tmp = nullptr;
      PH_CUDA(cudaMalloc(&a->d_sort_tmp, need));
      a->sort_tmp_bytes = need;
    }
    k_aff_gather_c<<<blocks, 256, 0, st>>>(src, a->d_cells, k, a->d_W);
    PH_CUDA(cub::DeviceRadixSort::SortKeys(a->d_sort_tmp, need, db, (int)N, 0, 64, st));
    ph_count_launch(5);
    // numpy _compute_virtual_index(n, q, alpha=1, beta=1) and _lerp
    const double nq = (double)N, q = radius;
    double vi = nq * q + (1.0 + q * (1.0 - 1.0 - 1.0)) - 1.0;
    double prev = floor(vi);
    long long lo = (long long)prev, hi = lo + 1;
    if (lo < 0) lo = 0;
    if (hi < 0) hi = 0;
    if (lo > (long long)N - 1) lo = (long long)N - 1;
    if (hi > (long long)N - 1) hi = (long long)N - 1;
    const double gamma = vi - prev;
    PH_CUDA(cudaMemcpyAsync(&q_lo, db.Current() + lo, 8, cudaMemcpyDeviceToHost, st));
    PH_CUDA(cudaMemcpyAsync(&q_hi, db.Current() + hi, 8, cudaMemcpyDeviceToHost, st));
    PH_CUDA(cudaStreamSynchronize(st));
    const double diff = q_hi - q_lo;
    r = q_lo + diff * gamma;
    if (gamma >= 0.5) r = q_hi - diff * (1.0 - gamma);
  }
  stats_out[0] = kw;
  stats_out[1] = kwc;
  stats_out[2] = r;
  stats_out[4] = maxd;
  stats_out[5] = maxc;
  stats_out[6] = q_lo;
  stats_out[7] = q_hi;
  if (red[2]) {  // NaN features / distances
    stats_out[3] = 1.0;
    return PH_OK;
  }
  k_aff_build<<<tgrid, tblock, 0, st>>>(a->d_feat, F, src, a->d_cells, k, use_copy_kernel, kw, kwc, r, strict,
                                        a->d_W, a->d_red);
  ph_count_launch();
  PH_CUDA(cudaMemcpyAsync(red, a->d_red, sizeof(red), cudaMemcpyDeviceToHost, st));
  PH_CUDA(cudaStreamSynchronize(st));
  stats_out[3] = red[2] ? 1.0 : 0.0;
  if (red[2]) return PH_OK;
  k_aff_degree<<<(k + 127) / 128, 128, 0, st>>>(a->d_W, k, a->d_dd);
  k_aff_laplacian<<<blocks, 256, 0, st>>>(a->d_W, k, a->d_dd);
  ph_count_launch(2);
  PH_CUDA(cudaGetLastError());
  if (dd_out) {
    PH_CUDA(cudaMemcpyAsync(a->h_pin, a->d_dd, (size_t)k * 8, cudaMemcpyDeviceToHost, st));
    PH_CUDA(cudaStreamSynchronize(st));
    memcpy(dd_out, a->h_pin, (size_t)k * 8);
  }
  a->k = k;
  return PH_OK;
}

// Y = L @ X for the Laplacian built by the last ph_affinity_build; X, Y: k x c row-major host arrays, c <= 8
extern "C" int ph_affinity_matmat(ph_affinity* a, const double* X, int32_t c, double* Y) {
  if (!a || !X || !Y) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (a->k < 1) PH_FAIL(PH_ERR_ARG, "no Laplacian: call ph_affinity_build first");
  if (c < 1 || c > kMaxCols) PH_FAIL(PH_ERR_ARG, "c=%d outside 1..%d", c, kMaxCols);
  PH_CUDA(cudaSetDevice(a->device));
  cudaStream_t st = a->st;
  const int k = a->k;
  if (a->rows_mode) PH_FAIL(PH_ERR_ARG, "the context holds a row block: use ph_affinity_matmat_rows");
  const int nrows = k;
  memcpy(a->h_pin, X, (size_t)k * c * 8);
  PH_CUDA(cudaMemcpyAsync(a->d_x, a->h_pin, (size_t)k * c * 8, cudaMemcpyHostToDevice, st));
  const int threads = 256, blocks = (k * 32 + threads - 1) / threads;
  switch (c) {
    case 1: k_aff_matmat<1><<<blocks, threads, 0, st>>>(a->d_W, k, a->d_x, a->d_y, nrows); break;
    case 2: k_aff_matmat<2><<<blocks, threads, 0, st>>>(a->d_W, k, a->d_x, a->d_y, nrows); break;
    case 3: k_aff_matmat<3><<<blocks, threads, 0, st>>>(a->d_W, k, a->d_x, a->d_y, nrows); break;
    case 4: k_aff_matmat<4><<<blocks, threads, 0, st>>>(a->d_W, k, a->d_x, a->d_y, nrows); break;
    case 5: k_aff_matmat<5><<<blocks, threads, 0, st>>>(a->d_W, k, a->d_x, a->d_y, nrows); break;
    case 6: k_aff_matmat<6><<<blocks, threads, 0, st>>>(a->d_W, k, a->d_x, a->d_y, nrows); break;
    case 7: k_aff_matmat<7><<<blocks, threads, 0, st>>>(a->d_W, k, a->d_x, a->d_y, nrows); break;
    default: k_aff_matmat<8><<<blocks, threads, 0, st>>>(a->d_W, k, a->d_x, a->d_y, nrows); break;
  }
  ph_count_launch();
  PH_CUDA(cudaGetLastError());
  PH_CUDA(cudaMemcpyAsync(a->h_pin + (size_t)k * kMaxCols, a->d_y, (size_t)k * c * 8, cudaMemcpyDeviceToHost, st));
  PH_CUDA(cudaStreamSynchronize(st));
  memcpy(Y, a->h_pin + (size_t)k * kMaxCols, (size_t)k * c * 8);
  return PH_OK;
}

// the k x k matrix currently held (affinity-derived Laplacian) -- tests and small inputs only
extern "C" int ph_affinity_matrix(ph_affinity* a, double* out) {
  if (!a || !out) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (a->k < 1 || a->rows_mode) PH_FAIL(PH_ERR_ARG, "no (whole) Laplacian: call ph_affinity_build first");
  PH_CUDA(cudaSetDevice(a->device));
  PH_CUDA(cudaMemcpy(out, a->d_W, (size_t)a->k * a->k * 8, cudaMemcpyDeviceToHost));
  return PH_OK;
}

// ------------------------------------------------------------------------------------------------
// Row-block calls: the clustering of one cell set sharded across GPUs.  Rank r holds rows [row_lo, row_hi) of the
// affinity / Laplacian; the host language combines the ranks between the steps (max of the statistics, concatenation of
// the degree vector and of the product's row blocks) -- see phertilizer_b200/cluster.py ShardedCopyDistance.  Every
// number is computed by exactly one rank with the arithmetic of the single-GPU kernels, so the assembled results are
// bit-identical to ph_affinity_build / ph_affinity_matmat.  radius < 1 (a quantile of all k^2 distances) is not
// sharded: use the single-GPU call.
// ------------------------------------------------------------------------------------------------
extern "C" int ph_affinity_stats_rows(ph_affinity* a, const int32_t* cells, int32_t k, const double* feats, int32_t F,
                                      int use_copy_kernel, int32_t row_lo, int32_t row_hi, double* out3) {
  if (!a || !cells || !feats || !out3) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (k < 1 || k > a->cap || F < 1 || F > kMaxF) PH_FAIL(PH_ERR_ARG, "bad shape k=%d F=%d", k, F);
  if (row_lo < 0 || row_hi > k || row_lo > row_hi || row_hi - row_lo > a->cap_rows)
    PH_FAIL(PH_ERR_ARG, "row block [%d, %d) outside 0..%d or larger than the context's %d rows", row_lo, row_hi, k, a->cap_rows);
  for (int i = 0; i < k; ++i)
    if (cells[i] < 0 || cells[i] >= a->n) PH_FAIL(PH_ERR_ARG, "cells[%d]=%d out of range", i, cells[i]);
  PH_CUDA(cudaSetDevice(a->device));
  cudaStream_t st = a->st;
  a->k = 0;
  a->row_lo = row_lo;
  a->row_hi = row_hi;
  a->use_copy = use_copy_kernel;
  a->rows_mode = 1;
  memcpy(a->h_pin, feats, (size_t)k * F * 8);
  memcpy(a->h_pin + (size_t)k * F, cells, (size_t)k * 4);
  PH_CUDA(cudaMemcpyAsync(a->d_feat, a->h_pin, (size_t)k * F * 8, cudaMemcpyHostToDevice, st));
  PH_CUDA(cudaMemcpyAsync(a->d_cells, a->h_pin + (size_t)k * F, (size_t)k * 4, cudaMemcpyHostToDevice, st));
  PH_CUDA(cudaMemsetAsync(a->d_red, 0, 4 * 8, st));
  a->k = -k;  // features / cells staged, no matrix yet (F kept in d_red[3] would be overkill: the caller passes it again)
  const int nr = row_hi - row_lo;
  if (nr > 0) {
    const CopySrc src{a->d_copy, a->d_emb, a->n, a->D};
    const dim3 grid((k + 31) / 32, (nr + 31) / 32), block(32, 8);
    k_aff_stats_rect<<<grid, block, 0, st>>>(a->d_feat, F, src, a->d_cells, k, use_copy_kernel, row_lo, row_hi, a->d_red);
    ph_count_launch();
  }
  unsigned long long red[4];
  PH_CUDA(cudaMemcpyAsync(red, a->d_red, sizeof(red), cudaMemcpyDeviceToHost, st));
  PH_CUDA(cudaStreamSynchronize(st));
  memcpy(&out3[0], &red[0], 8);
  memcpy(&out3[1], &red[1], 8);
  out3[2] = red[2] ? 1.0 : 0.0;
  return PH_OK;
}

extern "C" int ph_affinity_build_rows(ph_affinity* a, int32_t F, double kw, double kwc, double r, int strict, double* dd_block,
                                      int* has_nan) {
  if (!a || !dd_block || !has_nan) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (a->k >= 0) PH_FAIL(PH_ERR_ARG, "call ph_affinity_stats_rows first");
  PH_CUDA(cudaSetDevice(a->device));
  cudaStream_t st = a->st;
  const int k = -a->k, nr = a->row_hi - a->row_lo;
  *has_nan = 0;
  if (nr > 0) {
    const CopySrc src{a->d_copy, a->d_emb, a->n, a->D};
    const dim3 grid((k + 31) / 32, (nr + 31) / 32), block(32, 8);
    PH_CUDA(cudaMemsetAsync(a->d_red, 0, 4 * 8, st));
    k_aff_build_rect<<<grid, block, 0, st>>>(a->d_feat, F, src, a->d_cells, k, a->use_copy, kw, kwc, r, strict, a->row_lo, a->row_hi,
                                             a->d_W, a->d_red);
    k_aff_rowsum<<<(nr + 63) / 64, 64, 0, st>>>(a->d_W, nr, k, a->row_lo, a->d_dd);
    ph_count_launch(2);
    PH_CUDA(cudaGetLastError());
    unsigned long long red[4];
    PH_CUDA(cudaMemcpyAsync(red, a->d_red, sizeof(red), cudaMemcpyDeviceToHost, st));
    PH_CUDA(cudaMemcpyAsync(a->h_pin, a->d_dd, (size_t)nr * 8, cudaMemcpyDeviceToHost, st));
    PH_CUDA(cudaStreamSynchronize(st));
    memcpy(dd_block, a->h_pin, (size_t)nr * 8);
    *has_nan = red[2] ? 1 : 0;
  }
  return PH_OK;
}

extern "C" int ph_affinity_laplacian_rows(ph_affinity* a, const double* dd_full) {
  if (!a || !dd_full) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (a->k >= 0) PH_FAIL(PH_ERR_ARG, "call ph_affinity_stats_rows / ph_affinity_build_rows first");
  PH_CUDA(cudaSetDevice(a->device));
  cudaStream_t st = a->st;
  const int k = -a->k, nr = a->row_hi - a->row_lo;
  memcpy(a->h_pin, dd_full, (size_t)k * 8);
  PH_CUDA(cudaMemcpyAsync(a->d_ddf, a->h_pin, (size_t)k * 8, cudaMemcpyHostToDevice, st));
  if (nr > 0) {
    k_aff_laplacian_rect<<<1184, 256, 0, st>>>(a->d_W, nr, k, a->row_lo, a->d_ddf);
    ph_count_launch();
    PH_CUDA(cudaGetLastError());
  }
  PH_CUDA(cudaStreamSynchronize(st));
  a->k = k;  // the block of the Laplacian is in place
  return PH_OK;
}

// Y_block = L[row_lo:row_hi, :] @ X   (X: k x c host array, Y_block: (row_hi - row_lo) x c)
extern "C" int ph_affinity_matmat_rows(ph_affinity* a, const double* X, int32_t c, double* Y_block) {
  if (!a || !X || !Y_block) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (a->k < 1 || !a->rows_mode) PH_FAIL(PH_ERR_ARG, "no Laplacian rows: call ph_affinity_laplacian_rows first");
  if (c < 1 || c > kMaxCols) PH_FAIL(PH_ERR_ARG, "c=%d outside 1..%d", c, kMaxCols);
  PH_CUDA(cudaSetDevice(a->device));
  cudaStream_t st = a->st;
  const int k = a->k, nrows = a->row_hi - a->row_lo;
  memcpy(a->h_pin, X, (size_t)k * c * 8);
  PH_CUDA(cudaMemcpyAsync(a->d_x, a->h_pin, (size_t)k * c * 8, cudaMemcpyHostToDevice, st));
  if (nrows > 0) {
    const int threads = 256, blocks = (nrows * 32 + threads - 1) / threads;
    switch (c) {
      case 1: k_aff_matmat<1><<<blocks, threads, 0, st>>>(a->d_W, k, a->d_x, a->d_y, nrows); break;
      case 2: k_aff_matmat<2><<<blocks, threads, 0, st>>>(a->d_W, k, a->d_x, a->d_y, nrows); break;
      case 3: k_aff_matmat<3><<<blocks, threads, 0, st>>>(a->d_W, k, a->d_x, a->d_y, nrows); break;
      case 4: k_aff_matmat<4><<<blocks, threads, 0, st>>>(a->d_W, k, a->d_x, a->d_y, nrows); break;
      case 5: k_aff_matmat<5><<<blocks, threads, 0, st>>>(a->d_W, k, a->d_x, a->d_y, nrows); break;
      case 6: k_aff_matmat<6><<<blocks, threads, 0, st>>>(a->d_W, k, a->d_x, a->d_y, nrows); break;
      case 7: k_aff_matmat<7><<<blocks, threads, 0, st>>>(a->d_W, k, a->d_x, a->d_y, nrows); break;
      default: k_aff_matmat<8><<<blocks, threads, 0, st>>>(a->d_W, k, a->d_x, a->d_y, nrows); break;
    }
    ph_count_launch();
    PH_CUDA(cudaGetLastError());
    PH_CUDA(cudaMemcpyAsync(a->h_pin + (size_t)k * kMaxCols, a->d_y, (size_t)nrows * c * 8, cudaMemcpyDeviceToHost, st));
  }
  PH_CUDA(cudaStreamSynchronize(st));
  if (nrows > 0) memcpy(Y_block, a->h_pin + (size_t)k * kMaxCols, (size_t)nrows * c * 8);
  return PH_OK;
}

// ------------------------------------------------------------------------------------------------
// Device-resident block vectors for the eigensolver of the cut (phertilizer_b200/cluster.py device_lobpcg): the LOBPCG
// iteration of sklearn's spectral_embedding (utils.py:131-150 -> scipy.sparse.linalg.lobpcg) works on a handful of k x 3
// blocks (X, AX, R, AR, P, AP); keeping them here leaves only 3 x 3 ... 9 x 9 Gram matrices to cross the PCIe bus per
// iteration instead of k x 3 operands and results of every product.  Slots 0 .. 9, each k x 3 row-major.
// ------------------------------------------------------------------------------------------------
namespace {
int bv_check(ph_affinity* a, int slot) {
  if (!a) PH_FAIL(PH_ERR_ARG, "null context");
  if (a->k < 1) PH_FAIL(PH_ERR_ARG, "no Laplacian: build one first");
  if (slot < 0 || slot >= kSlots) PH_FAIL(PH_ERR_ARG, "slot %d outside 0..%d", slot, kSlots - 1);
  return PH_OK;
}
double* bv_ptr(ph_affinity* a, int slot) { return a->d_bv + (size_t)slot * a->cap * kBW; }
}  // namespace

extern "C" int ph_affinity_bv_set(ph_affinity* a, int slot, const double* host_k_by_3) {
  int rc = bv_check(a, slot);
  if (rc != PH_OK) return rc;
  if (!host_k_by_3) PH_FAIL(PH_ERR_ARG, "null pointer");
  PH_CUDA(cudaSetDevice(a->device));
  memcpy(a->h_pin, host_k_by_3, (size_t)a->k * kBW * 8);
  PH_CUDA(cudaMemcpyAsync(bv_ptr(a, slot), a->h_pin, (size_t)a->k * kBW * 8, cudaMemcpyHostToDevice, a->st));
  PH_CUDA(cudaStreamSynchronize(a->st));
  return PH_OK;
}

extern "C" int ph_affinity_bv_get(ph_affinity* a, int slot, double* host_k_by_3) {
  int rc = bv_check(a, slot);
  if (rc != PH_OK) return rc;
  if (!host_k_by_3) PH_FAIL(PH_ERR_ARG, "null pointer");
  PH_CUDA(cudaSetDevice(a->device));
  PH_CUDA(cudaMemcpyAsync(a->h_pin, bv_ptr(a, slot), (size_t)a->k * kBW * 8, cudaMemcpyDeviceToHost, a->st));
  PH_CUDA(cudaStreamSynchronize(a->st));
  memcpy(host_k_by_3, a->h_pin, (size_t)a->k * kBW * 8);
  return PH_OK;
}

// dst = L @ src on the device (whole Laplacian), or -- row-block contexts -- Y_block = L[row_lo:row_hi, :] @ src to the host
// (the caller assembles the ranks' blocks and puts the result back with ph_affinity_bv_set)
extern "C" int ph_affinity_bv_apply(ph_affinity* a, int src, int dst, double* host_block_or_null) {
  int rc = bv_check(a, src);
  if (rc == PH_OK && !a->rows_mode) rc = bv_check(a, dst);
  if (rc != PH_OK) return rc;
  PH_CUDA(cudaSetDevice(a->device));
  cudaStream_t st = a->st;
  const int k = a->k, nrows = a->rows_mode ? a->row_hi - a->row_lo : k;
  if (a->rows_mode && !host_block_or_null) PH_FAIL(PH_ERR_ARG, "row-block context: pass the host buffer of the block");
  double* out = a->rows_mode ? a->d_y : bv_ptr(a, dst);
  if (nrows > 0) {
    const int threads = 256, blocks = (nrows * 32 + threads - 1) / threads;
    k_aff_matmat<kBW><<<blocks, threads, 0, st>>>(a->d_W, k, bv_ptr(a, src), out, nrows);
    ph_count_launch();
    PH_CUDA(cudaGetLastError());
  }
  if (a->rows_mode) {
    if (nrows > 0) PH_CUDA(cudaMemcpyAsync(a->h_pin, out, (size_t)nrows * kBW * 8, cudaMemcpyDeviceToHost, st));
    PH_CUDA(cudaStreamSynchronize(st));
    if (nrows > 0) memcpy(host_block_or_null, a->h_pin, (size_t)nrows * kBW * 8);
  }
  return PH_OK;
}

// out[(a, i), (b, j)] = sum_r left_a[r, i] * right_b[r, j], row-major (3 nl) x (3 nr); nl <= 3, nr <= 6 slots
extern "C" int ph_affinity_bv_gram(ph_affinity* a, const int32_t* left, int32_t nl, const int32_t* right, int32_t nr, double* out) {
  if (!a || !left || !right || !out) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (a->k < 1) PH_FAIL(PH_ERR_ARG, "no Laplacian: build one first");
  if (nl < 1 || nl > 3 || nr < 1 || nr > 6) PH_FAIL(PH_ERR_ARG, "nl=%d / nr=%d out of range", nl, nr);
  BvList L, R;
  L.n = nl;
  R.n = nr;
  for (int i = 0; i < nl; ++i) {
    if (left[i] < 0 || left[i] >= kSlots) PH_FAIL(PH_ERR_ARG, "slot %d out of range", left[i]);
    L.p[i] = bv_ptr(a, left[i]);
  }
  for (int i = 0; i < nr; ++i) {
    if (right[i] < 0 || right[i] >= kSlots) PH_FAIL(PH_ERR_ARG, "slot %d out of range", right[i]);
    R.p[i] = bv_ptr(a, right[i]);
  }
  PH_CUDA(cudaSetDevice(a->device));
  cudaStream_t st = a->st;
  const int ne = nl * kBW * nr * kBW;
  const int chunks = a->k < 128 * kGramChunks ? (a->k + 127) / 128 : kGramChunks;   // a function of k only: fixed order
  k_bv_gram<<<chunks, 256, 0, st>>>(L, R, a->k, a->d_gram + kGramMax, reinterpret_cast<unsigned*>(a->d_gram), a->h_pin);
  ph_count_launch();
  PH_CUDA(cudaGetLastError());
  PH_CUDA(cudaStreamSynchronize(st));   // the last CTA stored the matrix into the page-locked buffer itself
  memcpy(out, a->h_pin, (size_t)ne * 8);
  return PH_OK;
}

// dst = sum_t slot_t @ C_t  (C_t 3 x 3 row-major in coeffs[t * 9 ...]; 1 <= n_terms <= 4); dst may be one of the sources
extern "C" int ph_affinity_bv_combine(ph_affinity* a, int dst, const int32_t* slots, const double* coeffs, int32_t n_terms) {
  int rc = bv_check(a, dst);
  if (rc != PH_OK) return rc;
  if (!slots || !coeffs || n_terms < 1 || n_terms > 4) PH_FAIL(PH_ERR_ARG, "bad terms");
  BvTerms T;
  T.n = n_terms;
  for (int t = 0; t < n_terms; ++t) {
    if (slots[t] < 0 || slots[t] >= kSlots) PH_FAIL(PH_ERR_ARG, "slot %d out of range", slots[t]);
    T.u[t] = bv_ptr(a, slots[t]);
    for (int i = 0; i < kBW * kBW; ++i) T.c[t][i] = coeffs[t * kBW * kBW + i];
  }
  PH_CUDA(cudaSetDevice(a->device));
  k_bv_combine<<<(a->k + 255) / 256, 256, 0, a->st>>>(T, a->k, bv_ptr(a, dst));
  ph_count_launch();
  PH_CUDA(cudaGetLastError());
  return PH_OK;
}
