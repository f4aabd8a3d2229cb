// Dense embedding kernels: K4 (Gaussian read-depth log-likelihood per tree node, clonal_tree.py:395-407) and
// K5 (pairwise Euclidean distance, phertilizer.py:286).  Both are fp64 (the reference's scipy arithmetic is fp64 and
// parity is 1e-9 relative), so they run on the CUDA-core fp64 pipe; tcgen05 has no fp64 MMA.
#include <float.h>
#include <math.h>

#include <algorithm>
#include <vector>

#include "ph_common.cuh"

namespace {

// ---------------------------------------------------------------- K4
// rows[] lists the cells of every node consecutively (stable order), node_off[k]..node_off[k+1].
// stat pass: column sums of f(x) over the rows of node k, f = x (pass 0) or (x-mean)^2 (pass 1).
// block = DX dims x RL row lanes (32 x 8 for wide embeddings, 4 x 64 for the 2-D UMAP case), grid = (ceil(D/DX), K).
// Fixed summation order -> deterministic.
template <int DX, int RL>
__global__ void __launch_bounds__(DX * RL) k_gauss_stat(const double* __restrict__ X, int D, const int32_t* __restrict__ rows,
                                                        const int32_t* __restrict__ node_off,
                                                        const double* __restrict__ mean, int pass,
                                                        double* __restrict__ out) {
  const int k = blockIdx.y;
  const int d = blockIdx.x * DX + threadIdx.x;
  const int beg = node_off[k], end = node_off[k + 1];
  const int cnt = end - beg;
  double acc = 0.0;
  if (d < D && cnt > 0) {
    const double mu = pass ? mean[(size_t)k * D + d] : 0.0;
    auto add = [&](double x) {
      if (pass) {
        const double t = x - mu;
        acc += t * t;
      } else {
        acc += x;
      }
    };
    // four rows in flight per thread (a lone dependent load per iteration left the pass latency-bound); the additions
    // keep their order, so the sums are the same bits as with one row at a time
    int r = beg + threadIdx.y;
    for (; r + 3 * RL < end; r += 4 * RL) {
      const double x0 = X[(size_t)rows[r] * D + d], x1 = X[(size_t)rows[r + RL] * D + d];
      const double x2 = X[(size_t)rows[r + 2 * RL] * D + d], x3 = X[(size_t)rows[r + 3 * RL] * D + d];
      add(x0);
      add(x1);
      add(x2);
      add(x3);
    }
    for (; r < end; r += RL) add(X[(size_t)rows[r] * D + d]);
  }
  __shared__ double sh[RL][DX + 1];
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && d < D) {
    double t = sh[0][threadIdx.x];
#pragma unroll
    for (int y = 1; y < RL; ++y) t += sh[y][threadIdx.x];
    out[(size_t)k * D + d] = cnt > 0 ? t / (double)cnt : 0.0;
  }
}

// per node: eps = 1e6*DBL_EPSILON*max|var|, rank, log_pdet; inv[d] = 1/var (kept) or 0 (dropped)
// cst[k*4 + {0,1,2,3}] = {eps, rank, log_pdet, min var}
__global__ void k_gauss_consts(const double* __restrict__ var, int D, double* __restrict__ inv, double* __restrict__ cst) {
  const int k = blockIdx.x;
  const double* v = var + (size_t)k * D;
  __shared__ double r[1024];
  __shared__ double r2[1024];
  double mx = 0.0, mn = INFINITY;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    mx = fmax(mx, fabs(v[d]));
    mn = fmin(mn, v[d]);
  }
  r[threadIdx.x] = mx;
  r2[threadIdx.x] = mn;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      r[threadIdx.x] = fmax(r[threadIdx.x], r[threadIdx.x + s]);
      r2[threadIdx.x] = fmin(r2[threadIdx.x], r2[threadIdx.x + s]);
    }
    __syncthreads();
  }
  const double eps = 1e6 * DBL_EPSILON * r[0];
  const double vmin = r2[0];
  __syncthreads();
  double lp = 0.0, rk = 0.0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const double s = v[d];
    const bool keep = s > eps;
    inv[(size_t)k * D + d] = keep ? 1.0 / s : 0.0;
    if (keep) {
      lp += log(s);
      rk += 1.0;
    }
  }
  r[threadIdx.x] = lp;
  r2[threadIdx.x] = rk;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      r[threadIdx.x] += r[threadIdx.x + s];
      r2[threadIdx.x] += r2[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    cst[k * 4 + 0] = eps;
    cst[k * 4 + 1] = r2[0];
    cst[k * 4 + 2] = r[0];
    cst[k * 4 + 3] = vmin;
  }
}

// one warp per cell: maha over kept dims, squared residual over dropped dims
__global__ void k_gauss_logpdf(const double* __restrict__ X, int n, int D, const uint8_t* __restrict__ cell_node, int K,
                               const double* __restrict__ mean, const double* __restrict__ inv,
                               const double* __restrict__ cst, double* __restrict__ per_cell) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const int k = cell_node[warp];
  if (k >= K) {
    if (lane == 0) per_cell[warp] = 0.0;
    return;
  }
  const double* x = X + (size_t)warp * D;
  const double* mu = mean + (size_t)k * D;
  const double* iv = inv + (size_t)k * D;
  double maha = 0.0, res = 0.0;
  auto add = [&](double xv, double m, double w) {
    const double t = xv - m;
    const double t2 = t * t;
    maha += t2 * w;
    res += (w == 0.0) ? t2 : 0.0;
  };
  // four strides of the row in flight per lane, accumulated in the original order (same bits)
  int d = lane;
  for (; d + 96 < D; d += 128) {
    const double x0 = x[d], x1 = x[d + 32], x2 = x[d + 64], x3 = x[d + 96];
    const double m0 = mu[d], m1 = mu[d + 32], m2 = mu[d + 64], m3 = mu[d + 96];
    const double w0 = iv[d], w1 = iv[d + 32], w2 = iv[d + 64], w3 = iv[d + 96];
    add(x0, m0, w0);
    add(x1, m1, w1);
    add(x2, m2, w2);
    add(x3, m3, w3);
  }
  for (; d < D; d += 32) add(x[d], mu[d], iv[d]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    maha += __shfl_xor_sync(0xffffffffu, maha, o);
    res += __shfl_xor_sync(0xffffffffu, res, o);
  }
  if (lane == 0) {
    const double eps = cst[k * 4 + 0], rank = cst[k * 4 + 1], lpd = cst[k * 4 + 2];
    const double LOG_2PI = 1.8378770664093453;  // scipy.stats._multivariate._LOG_2PI
    double out = -0.5 * (rank * LOG_2PI + lpd + maha);
    if (rank < (double)D && !(sqrt(res) < 1e3 * eps)) out = -INFINITY;
    per_cell[warp] = out;
  }
}

__global__ void k_node_reduce_d(const double* __restrict__ v, const int32_t* __restrict__ rows,
                                const int32_t* __restrict__ node_off, double* out) {
  const int k = blockIdx.x;
  double acc = 0.0;
  for (int r = node_off[k] + threadIdx.x; r < node_off[k + 1]; r += blockDim.x) acc += v[rows[r]];
  __shared__ double sh[1024];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[k] = sh[0];
}

// ---------------------------------------------------------------- K5
// 64x64 output tile per block, 4x4 outputs per thread, embedding staged through shared memory in slabs of 16 dims.
// Each output accumulates d = 0..D-1 in order with separately rounded multiply and add (scipy's C loop).
constexpr int PT = 64, PK = 16;
__global__ void __launch_bounds__(256) k_pdist(const double* __restrict__ X, int n, int D, int row_begin, int row_end,
                                               double* __restrict__ out) {
  __shared__ double sa[PK][PT + 1];
  __shared__ double sb[PK][PT + 1];
  const int i0 = row_begin + blockIdx.y * PT;
  const int j0 = blockIdx.x * PT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int d0 = 0; d0 < D; d0 += PK) {
    for (int t = threadIdx.x; t < PT * PK; t += 256) {
      const int r = t / PK, c = t % PK;
      const int d = d0 + c;
      const int gi = i0 + r, gj = j0 + r;
      sa[c][r] = (gi < row_end && d < D) ? X[(size_t)gi * D + d] : 0.0;
      sb[c][r] = (gj < n && d < D) ? X[(size_t)gj * D + d] : 0.0;
    }
    __syncthreads();
    const int kmax = min(PK, D - d0);
    for (int c = 0; c < kmax; ++c) {
      double xa[4], xb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) xa[a] = sa[c][ty * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) xb[b] = sb[c][tx * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const double t = __dsub_rn(xa[a], xb[b]);
          acc[a][b] = __dadd_rn(acc[a][b], __dmul_rn(t, t));
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int gi = i0 + ty * 4 + a;
    if (gi >= row_end) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int gj = j0 + tx * 4 + b;
      if (gj < n) out[(size_t)(gi - row_begin) * n + gj] = sqrt(acc[a][b]);
    }
  }
}

// ---------------------------------------------------------------- K5, Gram form on the FP64 tensor cores
// d(i,j)^2 = |x_i|^2 + |x_j|^2 - 2 x_i . x_j with the inner products from DMMA (mma.sync m8n8k4 f64: tcgen05 has no fp64
// type, the fp64 tensor path of sm_100 is the legacy warp-level instruction).  One multiply-add per pair and dimension
// instead of subtract, multiply, add: a third of the direct kernel's instructions, on the tensor pipe.  The cancellation
// of the Gram form costs relative accuracy (|x_i|^2 + |x_j|^2) / d^2 * eps: pairs whose squared distance falls below
// kGramFix * (|x_i|^2 + |x_j|^2) are re-evaluated by direct differences (the epilogue's fix-up), which bounds the
// relative error of every distance by ~ sqrt(D) * eps / kGramFix < 1e-9 -- NOT bit-identical to scipy's loop, hence
// opt-in (ph_pairwise_dist_gram); the direct kernel stays the default because the clustering compares distances exactly.
//
// CTA: 128 rows x 64 columns, 8 warps as 4 x 2, warp tile 32 x 32 = 4 x 4 MMA tiles; slabs of 32 dimensions staged in
// shared memory as [row][32] with the position of dimension c rotated by 4 * (row % 8): the 32 lanes of a fragment load
// (4 consecutive dimensions x 8 consecutive rows) and of a staging store (32 dimensions of one row) both touch every
// 8-byte bank pair exactly twice -- the minimum for 256 bytes.
constexpr double kGramFix = 1e-4;
constexpr int GM = 128, GN = 64, GK = 32;

__global__ void k_row_sqnorm(const double* __restrict__ X, int n, int D, double* __restrict__ sq) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const double* x = X + (size_t)warp * D;
  double acc = 0.0;
  for (int d = lane; d < D; d += 32) acc = fma(x[d], x[d], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) sq[warp] = acc;
}

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) k_pdist_gram(const double* __restrict__ X, const double* __restrict__ sq, int n, int D,
                                                    int row_begin, int row_end, double* __restrict__ out) {
  __shared__ double sa[GM * GK];
  __shared__ double sb[GN * GK];
  const int i0 = row_begin + blockIdx.y * GM;
  const int j0 = blockIdx.x * GN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp >> 1, wn = warp & 1;   // warp tile origin: rows wm * 32, columns wn * 32
  const int g = lane >> 2, t = lane & 3;     // fragment coordinates
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  for (int d0 = 0; d0 < D; d0 += GK) {
    // stage: one row of 32 dimensions per warp instruction (256 contiguous bytes of X), zero beyond the matrix
    for (int r = warp; r < GM; r += 8) {
      const int gi = i0 + r, d = d0 + lane;
      sa[r * GK + ((lane + 4 * (r & 7)) & 31)] = (gi < row_end && d < D) ? X[(size_t)gi * D + d] : 0.0;
    }
    for (int r = warp; r < GN; r += 8) {
      const int gj = j0 + r, d = d0 + lane;
      sb[r * GK + ((lane + 4 * (r & 7)) & 31)] = (gj < n && d < D) ? X[(size_t)gj * D + d] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k0 = 0; k0 < GK; k0 += 4) {
      const int p = (k0 + t + 4 * g) & 31;   // rows of a fragment are r0 + g with r0 a multiple of 8: (r & 7) = g
      double fa[4], fb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) fa[a] = sa[(wm * 32 + a * 8 + g) * GK + p];
#pragma unroll
      for (int b = 0; b < 4; ++b) fb[b] = sb[(wn * 32 + b * 8 + g) * GK + p];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
    }
    __syncthreads();
  }
  // epilogue: thread holds C[row = g][columns 2 t, 2 t + 1] of every 8 x 8 tile
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int gi = i0 + wm * 32 + a * 8 + g;
    if (gi >= row_end) continue;
    const double si = sq[gi];
    const double* xi = X + (size_t)gi * D;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gj = j0 + wn * 32 + b * 8 + 2 * t + e;
        if (gj >= n) continue;
        const double sj = sq[gj];
        double d2 = (si + sj) - 2.0 * acc[a][b][e];
        if (gi == gj) {
          d2 = 0.0;
        } else if (!(d2 > kGramFix * (si + sj))) {
          // near pair (or a rounding-negative value): direct differences, the reference's own arithmetic
          const double* xj = X + (size_t)gj * D;
          double s = 0.0;
          for (int d = 0; d < D; ++d) {
            const double df = __dsub_rn(xi[d], xj[d]);
            s = __dadd_rn(s, __dmul_rn(df, df));
          }
          d2 = s;
        }
        out[(size_t)(gi - row_begin) * n + gj] = sqrt(d2);
      }
    }
  }
}

}  // namespace

extern "C" int ph_gauss_node_ll(ph_handle* h, const double* X, int32_t D, const uint8_t* cell_node, int32_t K,
                                double* per_node, double* per_cell, int on_device, void* stream) {
  if (!h) PH_FAIL(PH_ERR_ARG, "null handle");
  if (!cell_node || !per_node) PH_FAIL(PH_ERR_ARG, "null pointer");
  bool cached = false;
  if (!X) {  // the embedding registered with ph_set_embedding
    if (!h->d_emb || on_device) PH_FAIL(PH_ERR_ARG, "X is NULL and no embedding was registered (ph_set_embedding)");
    X = h->d_emb;
    D = h->emb_D;
    cached = true;
  }
  if (D < 1 || K < 1 || K > PH_MAX_TREE_NODES) PH_FAIL(PH_ERR_ARG, "D=%d K=%d out of range", D, K);
  PH_CUDA(cudaSetDevice(h->device));
  const int n = h->n_cells;
  cudaStream_t st = on_device ? (cudaStream_t)stream : h->stream;
  // host copy of the node labels -> stable per-node row lists
  std::vector<uint8_t> lab(n);
  if (on_device) {
    PH_CUDA(cudaMemcpyAsync(lab.data(), cell_node, n, cudaMemcpyDeviceToHost, st));
    PH_CUDA(cudaStreamSynchronize(st));
  } else {
    memcpy(lab.data(), cell_node, n);
  }
  std::vector<int32_t> off(K + 1, 0), rows(n > 0 ? n : 1);
  for (int i = 0; i < n; ++i)
    if (lab[i] < K) off[lab[i] + 1]++;
  for (int k = 0; k < K; ++k) off[k + 1] += off[k];
  {
    std::vector<int32_t> pos(off.begin(), off.end() - 1);
    for (int i = 0; i < n; ++i)
      if (lab[i] < K) rows[pos[lab[i]]++] = i;
  }
  // scratch lives in the handle (grown on demand): with tens of GB allocated next to it, a cudaMalloc / cudaFree pair per
  // buffer per call cost milliseconds, and find_best_tree scores hundreds of trees
  double *dX = nullptr, *d_mean = nullptr, *d_var = nullptr, *d_inv = nullptr, *d_cst = nullptr, *d_pc = nullptr, *d_pn = nullptr;
  int32_t *d_rows = nullptr, *d_off = nullptr;
  uint8_t* d_lab = nullptr;
  int rc = PH_OK;
  auto body = [&]() -> int {
    const size_t kd = (size_t)K * D;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const bool need_x = !on_device && !cached;
    const size_t need = 3 * up(kd * 8) + up((size_t)K * 4 * 8) + up((size_t)(n > 0 ? n : 1) * 4) + up((size_t)(K + 1) * 4) +
                        up((size_t)n + 16) + up((size_t)K * 8) + up((size_t)n * 8) + (need_x ? up((size_t)n * D * 8) : 0);
    if (need > h->gauss_bytes) {
      cudaFree(h->d_gauss);
      h->d_gauss = nullptr;
      h->gauss_bytes = 0;
      PH_CUDA(cudaMalloc(&h->d_gauss, need));
      h->gauss_bytes = need;
    }
    uint8_t* cur = h->d_gauss;
    auto take = [&](size_t b) { uint8_t* r = cur; cur += up(b); return r; };
    d_mean = (double*)take(kd * 8);
    d_var = (double*)take(kd * 8);
    d_inv = (double*)take(kd * 8);
    d_cst = (double*)take((size_t)K * 4 * 8);
    d_rows = (int32_t*)take((size_t)(n > 0 ? n : 1) * 4);
    d_off = (int32_t*)take((size_t)(K + 1) * 4);
    d_lab = (uint8_t*)take((size_t)n + 16);
    d_pn = (double*)take((size_t)K * 8);
    d_pc = (double*)take((size_t)n * 8);
    if (need_x) dX = (double*)take((size_t)n * D * 8);
    // the small host arrays travel through the handle's page-locked staging buffer (truly asynchronous copies; pageable
    // sources are staged by the driver call by call, which costs milliseconds on some hosts and trees are scored by the hundred)
    uint8_t* hp = h->h_stage;
    const size_t hp_need = up((size_t)n * 4) + up((size_t)(K + 1) * 4) + up((size_t)n) + up((size_t)K * 8) + up((size_t)n * 8);
    const bool staged = !on_device && hp && hp_need <= h->h_stage_bytes;
    int32_t* p_rows = rows.data();
    int32_t* p_off = off.data();
    uint8_t* p_lab = lab.data();
    double *p_pn = nullptr, *p_pc = nullptr;
    if (staged) {
      auto htake = [&](size_t b) { uint8_t* r = hp; hp += up(b); return r; };
      p_rows = (int32_t*)htake((size_t)n * 4);
      p_off = (int32_t*)htake((size_t)(K + 1) * 4);
      p_lab = htake((size_t)n);
      p_pn = (double*)htake((size_t)K * 8);
      p_pc = (double*)htake((size_t)n * 8);
      memcpy(p_rows, rows.data(), (size_t)n * 4);
      memcpy(p_off, off.data(), (size_t)(K + 1) * 4);
      memcpy(p_lab, lab.data(), (size_t)n);
    }
    PH_CUDA(cudaMemcpyAsync(d_rows, p_rows, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    PH_CUDA(cudaMemcpyAsync(d_off, p_off, (size_t)(K + 1) * 4, cudaMemcpyHostToDevice, st));
    const double* Xd = X;
    const uint8_t* labd = cell_node;
    if (!on_device) {
      if (!cached) PH_CUDA(cudaMemcpyAsync(dX, X, (size_t)n * D * 8, cudaMemcpyHostToDevice, st));
      PH_CUDA(cudaMemcpyAsync(d_lab, p_lab, n, cudaMemcpyHostToDevice, st));
      Xd = cached ? X : dX;
      labd = d_lab;
    }
    double* pc = (on_device && per_cell) ? per_cell : d_pc;
    double* pn = on_device ? per_node : d_pn;
    dim3 gs((D + 31) / 32, K), bs(32, 8);
    if (D <= 4) {
      dim3 gs4((D + 3) / 4, K), bs4(4, 64);
      k_gauss_stat<4, 64><<<gs4, bs4, 0, st>>>(Xd, D, d_rows, d_off, nullptr, 0, d_mean);
      k_gauss_stat<4, 64><<<gs4, bs4, 0, st>>>(Xd, D, d_rows, d_off, d_mean, 1, d_var);
    } else {
      k_gauss_stat<32, 8><<<gs, bs, 0, st>>>(Xd, D, d_rows, d_off, nullptr, 0, d_mean);
      k_gauss_stat<32, 8><<<gs, bs, 0, st>>>(Xd, D, d_rows, d_off, d_mean, 1, d_var);
    }
    k_gauss_consts<<<K, 1024, 0, st>>>(d_var, D, d_inv, d_cst);
    k_gauss_logpdf<<<(n * 32 + 255) / 256, 256, 0, st>>>(Xd, n, D, labd, K, d_mean, d_inv, d_cst, pc);
    k_node_reduce_d<<<K, 1024, 0, st>>>(pc, d_rows, d_off, pn);
    ph_count_launch(5);
    PH_CUDA(cudaGetLastError());
    if (!on_device) {
      PH_CUDA(cudaMemcpyAsync(staged ? p_pn : per_node, pn, (size_t)K * 8, cudaMemcpyDeviceToHost, st));
      if (per_cell) PH_CUDA(cudaMemcpyAsync(staged ? p_pc : per_cell, pc, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    }
    PH_CUDA(cudaStreamSynchronize(st));  // host staging vectors go out of scope; the scratch is reused by the next call
    if (staged) {
      memcpy(per_node, p_pn, (size_t)K * 8);
      if (per_cell) memcpy(per_cell, p_pc, (size_t)n * 8);
    }
    return PH_OK;
  };
  rc = body();
  return rc;
}

// keep the embedding (read_depth, data.py:14) on the device: ph_gauss_node_ll(X = NULL) then scores trees without
// re-uploading it (find_best_tree scores hundreds of candidate trees, clonal_tree_list.py:196-234)
extern "C" int ph_set_embedding(ph_handle* h, const double* X, int32_t D) {
  if (!h || !X) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (D < 1) PH_FAIL(PH_ERR_ARG, "D=%d out of range", D);
  PH_CUDA(cudaSetDevice(h->device));
  cudaFree(h->d_emb);
  h->d_emb = nullptr;
  PH_CUDA(cudaMalloc(&h->d_emb, (size_t)h->n_cells * D * 8));
  PH_CUDA(cudaMemcpy(h->d_emb, X, (size_t)h->n_cells * D * 8, cudaMemcpyHostToDevice));
  h->emb_D = D;
  return PH_OK;
}

extern "C" int ph_pairwise_dist(const double* X, int32_t n, int32_t D, int32_t row_begin, int32_t row_end, double* out,
                                int device, int on_device, void* stream) {
  if (!X || !out) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (n < 1 || D < 1 || row_begin < 0 || row_end > n || row_begin > row_end)
    PH_FAIL(PH_ERR_ARG, "bad shape n=%d D=%d rows=[%d,%d)", n, D, row_begin, row_end);
  int ndev = 0;
  PH_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) PH_FAIL(PH_ERR_CUDA, "device %d not available (%d CUDA devices)", device, ndev);
  PH_CUDA(cudaSetDevice(device));
  const int nr = row_end - row_begin;
  if (nr == 0) return PH_OK;
  cudaStream_t st = on_device ? (cudaStream_t)stream : (cudaStream_t)0;
  double *dX = nullptr, *dO = nullptr;
  const double* Xd = X;
  double* Od = out;
  int rc = PH_OK;
  auto body = [&]() -> int {
    if (!on_device) {
      PH_CUDA(cudaMalloc(&dX, (size_t)n * D * 8));
      PH_CUDA(cudaMalloc(&dO, (size_t)nr * n * 8));
      PH_CUDA(cudaMemcpyAsync(dX, X, (size_t)n * D * 8, cudaMemcpyHostToDevice, st));
      Xd = dX;
      Od = dO;
    }
    dim3 grid((n + PT - 1) / PT, (nr + PT - 1) / PT);
    k_pdist<<<grid, 256, 0, st>>>(Xd, n, D, row_begin, row_end, Od);
    ph_count_launch();
    PH_CUDA(cudaGetLastError());
    if (!on_device) {
      PH_CUDA(cudaMemcpyAsync(out, dO, (size_t)nr * n * 8, cudaMemcpyDeviceToHost, st));
      PH_CUDA(cudaStreamSynchronize(st));
    }
    return PH_OK;
  };
  rc = body();
  cudaFree(dX);
  cudaFree(dO);
  return rc;
}

// K5 on the FP64 tensor cores (Gram form + direct fix-up of near pairs, see k_pdist_gram): every distance within 1e-9
// relative of the direct kernel, not bit-identical to it.  Device pointers only (the n x n result of a large embedding
// never fits a host transfer budget anyway); `sq_scratch` [n] doubles of device scratch for the squared row norms.
extern "C" int ph_pairwise_dist_gram(const double* X, int32_t n, int32_t D, int32_t row_begin, int32_t row_end, double* out,
                                     double* sq_scratch, int device, void* stream) {
  if (!X || !out || !sq_scratch) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (n < 1 || D < 1 || row_begin < 0 || row_end > n || row_begin > row_end)
    PH_FAIL(PH_ERR_ARG, "bad shape n=%d D=%d rows=[%d,%d)", n, D, row_begin, row_end);
  int ndev = 0;
  PH_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) PH_FAIL(PH_ERR_CUDA, "device %d not available (%d CUDA devices)", device, ndev);
  PH_CUDA(cudaSetDevice(device));
  const int nr = row_end - row_begin;
  if (nr == 0) return PH_OK;
  cudaStream_t st = (cudaStream_t)stream;
  k_row_sqnorm<<<(n * 32 + 255) / 256, 256, 0, st>>>(X, n, D, sq_scratch);
  dim3 grid((n + GN - 1) / GN, (nr + GM - 1) / GM);
  k_pdist_gram<<<grid, 256, 0, st>>>(X, sq_scratch, n, D, row_begin, row_end, out);
  ph_count_launch(2);
  PH_CUDA(cudaGetLastError());
  return PH_OK;
}
