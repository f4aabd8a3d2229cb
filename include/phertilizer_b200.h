/*
 * phertilizer_b200 -- C ABI of the B200-native likelihood hot path of Phertilizer.
 *
 * The reference (elkebir-group/phertilizer) has no FFI: its seam is the Python `Data` dataclass
 * (phertilizer/data.py:8-20) holding four DENSE n x m numpy arrays (var, total, like0, like1_marg) that the
 * tree operations read through `M[np.ix_(cells, muts)].<reduce>` expressions.  This library owns a
 * device-resident sparse replacement of that state (one handle per GPU) and exports one entry point per
 * family of those expressions.  Every function cites the reference lines it replaces.
 *
 * Conventions
 *   - plain C, no torch / C++ types; all sizes are explicit.
 *   - a "label" array assigns every cell (or SNV) to a cluster: 0 = not in any cluster (excluded from the
 *     np.ix_ subset), 1..C = cluster index.  This is how `cells`, `cellsA`, `cellsB`, `muts`, ... index arrays of
 *     the reference are passed.
 *   - label / node arrays are read 4 bytes at a time: device buffers passed with `on_device != 0` must be 4-byte
 *     aligned and readable up to the next multiple of 4 bytes (any cudaMalloc'd or torch-allocated buffer is).
 *   - `on_device != 0`: every data pointer of the call is a device pointer on the handle's GPU and the call is
 *     enqueued on `stream` (a cudaStream_t passed as void*, may be NULL) WITHOUT synchronising.
 *     `on_device == 0`: pointers are host memory; the call stages them through pinned buffers, runs, copies
 *     the results back and synchronises before returning.
 *   - return value: PH_OK or an error code; `ph_last_error()` gives the message of the last failure on the
 *     calling thread.  There is NO CPU fallback: without a usable GPU every compute call fails with PH_ERR_CUDA.
 *   - one caller thread per handle (the reference is single threaded).
 */
#ifndef PHERTILIZER_B200_H
#define PHERTILIZER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PH_ABI_VERSION 9

#if defined(__GNUC__)
#define PH_API __attribute__((visibility("default")))
#else
#define PH_API
#endif

enum ph_status {
  PH_OK = 0,
  PH_ERR_ARG = 1,         /* bad argument (null pointer, label out of range, ...) */
  PH_ERR_CUDA = 2,        /* CUDA runtime error, message in ph_last_error() */
  PH_ERR_UNSUPPORTED = 3  /* shape outside what the kernels support (see message) */
};

#define PH_MAX_FAST_CLUSTERS 3 /* clusters per pass on the packed-counter fast path */
#define PH_MAX_TREE_NODES 64   /* tree nodes for ph_tree_loglik / general tables */

typedef struct ph_handle ph_handle;

typedef struct ph_info {
  int32_t n_cells, n_snvs, n_codes;
  int64_t nnz;
  int32_t dense_codes_by_snv;  /* D of the SNV-major (CSC-like) copy  */
  int32_t dense_codes_by_cell; /* D of the cell-major (CSR-like) copy */
  int64_t device_bytes;        /* bytes of HBM owned by the handle */
  int32_t device, sm_count;
  int32_t group_by_snv, group_by_cell; /* lines (SNVs / cells) per warp batch */
  int32_t labels_in_smem_by_snv, labels_in_smem_by_cell;
  int32_t chunks_by_snv, chunks_by_cell; /* 65520-index chunks of the minor axis (16-bit entries) */
  int64_t stream_bytes_by_snv, stream_bytes_by_cell; /* bytes one full pass reads from HBM (entries incl. padding + tail) */
  int64_t stream4_bytes_by_snv, stream4_bytes_by_cell; /* same for the batched-4 copy whole-orientation passes use; 0 = not kept */
} ph_info;

PH_API const char* ph_last_error(void);
PH_API int ph_abi_version(void);

/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
PH_API int64_t ph_kernel_launches(void);

/* CUDA devices visible to the process (the `--runs` fan-out puts one worker on each). */
PH_API int ph_device_count(int* out);

/*
 * Build the device-resident observation store.  Replaces `sparse_matrix()` + the four dense matrices of `Data`
 * (phertilizer/phertilizer.py:153-161, 295-306; data.py:8-20).
 *   cell/snv [nnz]  : COO coordinates, 0 <= cell < n_cells, 0 <= snv < n_snvs, no duplicates
 *   code [nnz]      : index of the observation's (var,total) pair in the likelihood table, 0 <= code < n_codes,
 *                     codes should be ordered by descending frequency (code 0 = most frequent pair): the leading
 *                     codes with enough entries per line get the compact 16-bit layout
 *   L0/L1 [n_codes] : log-likelihoods of the pair under y=0 / y=1 (phertilizer.py:121-151)
 *   var_pos[n_codes]: 1 iff var > 0 for the pair (np.count_nonzero(var...) call sites)
 *   coo_on_device   : cell/snv/code are device pointers (they are only read)
 */
PH_API int ph_create(int32_t n_cells, int32_t n_snvs, int64_t nnz, const int32_t* cell, const int32_t* snv,
              const uint16_t* code, int coo_on_device, int32_t n_codes, const double* L0, const double* L1,
              const uint8_t* var_pos, int device, ph_handle** out);
PH_API void ph_destroy(ph_handle* h);
PH_API int ph_get_info(const ph_handle* h, ph_info* out);

/* Tuning knob: SNVs / cells per warp batch (4, 8 or 32; 16 means 8); -1 = automatic (4 for launches with fewer than
 * two batches of 8 per warp of the grid, else 8); 0 = keep. */
PH_API int ph_set_group(ph_handle* h, int group_by_snv, int group_by_cell);

/*
 * K1 -- per-SNV table over cell clusters.
 *   for every listed SNV j and cluster c in 1..C:
 *     S0[j,c] = sum_{i in cluster c} like0[i,j]      S1[j,c] = sum like1[i,j]
 *     Nobs[j,c] = #{i in c : total[i,j] > 0}         Nvar[j,c] = #{i in c : var[i,j] > 0}
 * Replaces: linear_split.py:186-188 (Nvar,Nobs), 224 (Nobs), 229-230 (S0,S1);
 *           branching_split.py:316-317, 332-337; check_reads axis=0 (linear_split.py:393-396,
 *           branching_split.py:428-431); clonal_tree.py:581-591 (find_bad_snvs), 1136 (Seed.strip: sum(var)==0 <=> Nvar==0).
 *   cell_label [n_cells] ; snv_list [n_list] optional subset of SNVs (NULL = all, then n_list = n_snvs)
 *   outputs [n_list * C], row k = k-th listed SNV, column c-1.  Any output pointer may be NULL.
 */
PH_API int ph_snv_table(ph_handle* h, const uint8_t* cell_label, int32_t C, const int32_t* snv_list, int32_t n_list,
                 double* S0, double* S1, int32_t* Nobs, int32_t* Nvar, int on_device, void* stream);

/*
 * K1 + K1a -- fused SNV reassignment of the linear split (linear_split.py:203-237).
 *   cluster 1 of cell_label = cellsA.  For every listed SNV:
 *     nobs_out[k]  = Nobs[j,A]                                   (input of the q75 gate, line 224-227)
 *     label_out[k] = 2 (mutsB) if S0[j,A]/lenA >= S1[j,A]/lenA else 1 (mutsA)     (lines 229-234; ties -> mutsB)
 *   The quantile gate itself stays with the caller (it needs all nobs).
 */
PH_API int ph_assign_linear(ph_handle* h, const uint8_t* cell_label, double lenA, const int32_t* snv_list, int32_t n_list,
                     uint8_t* label_out, int32_t* nobs_out, int on_device, void* stream);

/*
 * K1 + K1a -- fused 3-way SNV reassignment of the branching split (branching_split.py:293-348).
 *   clusters 1, 2 of cell_label = cellsA, cellsB.
 *     candA = S1[j,A] + S0[j,B]; candB = S0[j,A] + S1[j,B]; candC = S1[j,A] + S1[j,B]
 *     label_out[k] = 1 (mutsA) if candA == max, else 2 (mutsB) if candB == max, else 3 (mutsC)   (lines 340-346)
 *     nobs_out[2k], nobs_out[2k+1] = Nobs[j,A], Nobs[j,B]                                         (lines 316-319)
 */
PH_API int ph_assign_branching(ph_handle* h, const uint8_t* cell_label, const int32_t* snv_list, int32_t n_list,
                        uint8_t* label_out, int32_t* nobs_out, int on_device, void* stream);

/*
 * Cell-sharded K1 (multi-GPU): every rank holds a contiguous block of cells and all SNVs.
 * ph_snv_partial writes THIS rank's partial table packed as  [S0 | S1 | Nobs]  (3 * n_list * C doubles, counts are
 * exact in a double) so that one sum all-reduce (NCCL over NVLink) combines the ranks; ph_finish_linear /
 * ph_finish_branching then apply the K1a decision rules of ph_assign_linear / ph_assign_branching to the summed
 * table (C = 1 resp. 2).  Device pointers only; enqueued on `stream` of the current device.
 */
PH_API int ph_snv_partial(ph_handle* h, const uint8_t* cell_label, int32_t C, const int32_t* snv_list, int32_t n_list,
                   double* packed, int on_device, void* stream);
PH_API int ph_finish_linear(const double* packed, int32_t n_list, double lenA, uint8_t* label_out, int32_t* nobs_out,
                     void* stream);
PH_API int ph_finish_branching(const double* packed, int32_t n_list, uint8_t* label_out, int32_t* nobs_out, void* stream);

/*
 * Cell-sharded K1 + K1a over NVLink peer memory (no NCCL on the data path).  Each rank creates a context on its handle,
 * exports an opaque handle blob of ph_peer_handle_bytes() bytes (a CUDA IPC memory handle), the host language gathers
 * the blobs of all ranks (rank order) and passes them to ph_peer_connect.  ph_peer_assign_* then behave like
 * ph_assign_linear / ph_assign_branching on the union of all ranks' cells: `cell_label` are THIS rank's cells, every
 * rank receives the full label_out / nobs_out.  What travels between the GPUs is the integer (cluster, code) histogram
 * of every SNV, so the outputs are bit-identical to the single-GPU call for any number of ranks.  Collective: all
 * ranks must make the same sequence of calls.  Device pointers only, enqueued on `stream`.
 * Limits: 1..8 ranks of one node (1 = the same kernels on a local window), <= 65535 cells per rank, <= 64 (var,total)
 * codes (else PH_ERR_UNSUPPORTED: use ph_snv_partial + an all-reduce).
 */
typedef struct ph_peer ph_peer;
PH_API int ph_peer_handle_bytes(void);
PH_API int ph_peer_create(ph_handle* h, int rank, int world, int32_t total_snvs /* 0: every rank holds all SNVs */, ph_peer** out);
PH_API int ph_peer_export(ph_peer* p, void* handle_out);
PH_API int ph_peer_connect(ph_peer* p, const void* all_handles);
PH_API int ph_peer_assign_linear(ph_peer* p, const uint8_t* cell_label, double lenA, const int32_t* snv_list,
                          int32_t n_list, uint8_t* label_out, int32_t* nobs_out, void* stream);
PH_API int ph_peer_assign_branching(ph_peer* p, const uint8_t* cell_label, const int32_t* snv_list, int32_t n_list,
                             uint8_t* label_out, int32_t* nobs_out, void* stream);
/*
 * SNV-sharded variant: every rank holds ALL cells and the SNVs [snv_offset, snv_offset + n_snvs) of `total_snvs`
 * (ph_peer_create).  Per-SNV tables are independent across SNVs, so the pass needs no exchange of partial tables: the
 * kernel stores the labels and counts of every batch of 8 SNVs as one 64/128-byte record into all ranks' windows
 * while it streams the next batch, and label_out / nobs_out [total_snvs (* 2)] are complete on every rank after the
 * call.  `cell_label` covers all cells.  Blocks must be balanced: at most ceil(total_snvs / world) + 8 SNVs per rank
 * (PH_ERR_UNSUPPORTED otherwise) and must not overlap.
 */
PH_API int ph_peer_assign_linear_snv(ph_peer* p, const uint8_t* cell_label, double lenA, int32_t snv_offset,
                              uint8_t* label_out, int32_t* nobs_out, void* stream);
PH_API int ph_peer_assign_branching_snv(ph_peer* p, const uint8_t* cell_label, int32_t snv_offset, uint8_t* label_out,
                                 int32_t* nobs_out, void* stream);
PH_API int ph_peer_timed_out(ph_peer* p, void* stream); /* 1 if a wait for a peer ever gave up (~3 s; PH_PEER_TIMEOUT_MS) */
/* Diagnostics of the first SNV-sharded wait that gave up: out[24] = {timed out, kind (1 header, 2 id/label word, 3/4 count
 * words), source rank, slot, step wanted, step seen, payload seen, this rank's step counter, then per source rank r:
 * slots its header announced for that step (0xFFFFFFFF = no header), id/label words of that step present in the window}. */
PH_API int ph_peer_debug(ph_peer* p, uint32_t* out);
PH_API void ph_peer_destroy(ph_peer* p);

/*
 * K2 -- per-cell table over SNV clusters (transpose of K1).
 *   A0[i,q] = sum_{j in cluster q} like0[i,j], A1 likewise, Nobs[i,q], Nvar[i,q].
 * Replaces: linear_split.py:301-304 (tmb features), 400-410 (check_metrics), check_reads axis=1;
 *           branching_split.py:166-171 (calc_tmb), 395-411; clonal_tree.py:539-548, 1140 (Seed.strip).
 */
PH_API int ph_cell_table(ph_handle* h, const uint8_t* snv_label, int32_t Q, const int32_t* cell_list, int32_t n_list,
                  double* A0, double* A1, int32_t* Nobs, int32_t* Nvar, int on_device, void* stream);

/*
 * K3 -- block sums over the rectangular blocks (cell cluster c) x (SNV cluster q), c in 1..C, q in 1..Q:
 *   sum0[q-1][c-1] = sum like0, sum1 = sum like1, nobs = #observed, nvar = #(var > 0).
 * Replaces: utils.py:218-220 (check_obs), clonal_tree.py:1144-1146 (Seed.count_obs), phertilizer.py:405,
 *           linear_split.py:549-554 and branching_split.py:460-466 (compute_norm_likelihood).
 * Outputs [Q * C] row-major (SNV cluster major).
 */
PH_API int ph_block_sums(ph_handle* h, const uint8_t* cell_label, int32_t C, const uint8_t* snv_label, int32_t Q,
                  double* sum0, double* sum1, int64_t* nobs, int64_t* nvar, int on_device, void* stream);

/*
 * The two fixed-order reductions behind ph_block_sums (Q <= 3: per-cell table over the SNV clusters, summed by cell
 * cluster) and ph_tree_loglik (per-cell values summed by node), on caller-provided HOST tables.  A cell-sharded store
 * (phertilizer_b200/sharded_store.py) assembles the per-cell rows of all ranks -- every row is computed wholly by the
 * rank owning the cell -- and reduces them with the same kernels, so block sums and tree log-likelihoods are
 * bit-identical for any number of GPUs (the scalars decide which split / tree wins: linear_split.py:549-554,
 * branching_split.py:460-466, clonal_tree_list.py:196-234).
 *   ph_line_reduce: table [n_lines x ncol] (ncol <= 3), line_label in 0..nlab -> out[col * nlab + (label - 1)]
 *   ph_node_reduce: per_node[k] = sum of per_line over the lines with line_node == k
 * Synchronous; n_lines <= max(n_cells, n_snvs) of the handle.
 */
PH_API int ph_line_reduce(ph_handle* h, const double* S0, const double* S1, const int32_t* Nobs, const int32_t* Nvar,
                   const uint8_t* line_label, int32_t n_lines, int32_t ncol, int32_t nlab, double* sum0, double* sum1,
                   int64_t* nobs, int64_t* nvar);
PH_API int ph_node_reduce(ph_handle* h, const double* per_line, const uint8_t* line_node, int32_t n_lines, int32_t K,
                   double* per_node);

/*
 * a-9 -- variant read-count log-likelihood of a clonal tree, per node
 * (clonal_tree.py:932-944 with presence_by_node 343-358 and get_ancestral_muts 979-992):
 *   per_node[k] = sum_{i : cell_node[i]==k} sum_j ( present(k, snv_node[j]) ? like1[i,j] : like0[i,j] )
 *   cell_node [n_cells] : tree node of each cell, 255 = cell not in the tree
 *   snv_node  [n_snvs]  : node whose incoming edge carries the SNV, 255 = SNV not in the tree (always absent)
 *   present [K*K]       : present[k*K + q] != 0 iff SNVs of node q are present in cells of node k
 *                         (q on the root->k path, minus losses)
 *   per_node [K] ; per_cell [n_cells] optional (NULL to skip): the cell's own log-likelihood under its node
 */
PH_API int ph_tree_loglik(ph_handle* h, const uint8_t* cell_node, const uint8_t* snv_node, const uint8_t* present,
                   int32_t K, double* per_node, double* per_cell, int on_device, void* stream);

/*
 * a-11 -- general tables for post_process (clonal_tree.py:612-666): for every listed SNV j and every candidate
 * node k in 0..K-1:  T[j,k] = sum_i ( member[k*K + cell_node[i]] ? like1[i,j] : like0[i,j] )  over ALL cells
 * (cells with cell_node == 255 count as like0), i.e. `np.dot(y,like1)+np.dot(1-y,like0)` of reassign_snv with
 * y = c_dict[k].  ph_cell_node_table is the transpose for reassign_cell (y = y_dict[k] over SNVs).
 */
PH_API int ph_snv_node_table(ph_handle* h, const uint8_t* cell_node, const uint8_t* member, int32_t K,
                      const int32_t* snv_list, int32_t n_list, double* T, int on_device, void* stream);
PH_API int ph_cell_node_table(ph_handle* h, const uint8_t* snv_node, const uint8_t* member, int32_t K,
                       const int32_t* cell_list, int32_t n_list, double* T, int on_device, void* stream);

/*
 * K4 -- Gaussian read-depth log-likelihood of the embedding per tree node (clonal_tree.py:395-407):
 * per node mean and population variance per dimension, then the diagonal `multivariate_normal(mean, diag(var),
 * allow_singular=True).logpdf(x).sum()` with scipy's singular-covariance rules (eps = 1e6*DBL_EPSILON*max|var|,
 * dims with var <= eps dropped, rank-dependent normaliser, -inf for rows off the support).
 *   X [n_cells * D] row-major float64 ; cell_node [n_cells], 255 = skip ; per_node [K] (0 for empty nodes)
 */
/* Register the embedding once; ph_gauss_node_ll(X = NULL, on_device = 0) then uses the device copy. */
PH_API int ph_set_embedding(ph_handle* h, const double* X, int32_t D);
PH_API int ph_gauss_node_ll(ph_handle* h, const double* X, int32_t D, const uint8_t* cell_node, int32_t K,
                     double* per_node, double* per_cell, int on_device, void* stream);

/*
 * Cell-clustering step (linear_split.py:277-390, branching_split.py:175-291,
 * utils.py:131-150).  The context keeps `copy_distance = squareform(pdist(embedding))` (phertilizer.py:286) in HBM.
 * ph_affinity_build forms, for a set of cells and their mutation-burden features, the reference's affinity
 *     W = exp(-1*d_mat/kw) * (exp(-1*c_mat/kwc) with entries c_mat > r (strict: >= r) zeroed)        [use_copy_kernel]
 *     d_mat = squareform(pdist(feats)), c_mat = copy_distance[ix_(cells, cells)], kw/kwc = 0.2*(max-min),
 *     r = np.quantile(c_mat, radius)
 * and overwrites it with the normalised graph Laplacian sklearn's spectral_embedding builds from it
 * (scipy csgraph_laplacian(normed=True) + unit diagonal); dd_out [k] receives its sqrt-degree vector.
 * ph_affinity_matmat then serves Y = L @ X (X, Y: k x c row-major, c <= 8) to the host eigensolver (scipy LOBPCG, as
 * in the reference).  stats_out [8] = kw, kwc, r, has_nan, max d, max c, q_lo, q_hi.  Host pointers; synchronous.
 * ph_affinity_rows copies rows of copy_distance (rows == NULL and n_rows == n: the whole matrix).
 */
typedef struct ph_affinity ph_affinity;
PH_API int ph_affinity_create(const double* embedding, int32_t n, int32_t D, int device, ph_affinity** out);
PH_API int ph_affinity_rows(ph_affinity* a, const int32_t* rows, int32_t n_rows, double* out);
/* impute_mut_features (utils.py:87-129): for every listed cell c the donor j (donor_mask[j] != 0, j != c) with the
 * smallest copy_distance[c][j] (lowest j among equals) and how many donors share that minimum. */
PH_API int ph_affinity_nearest(ph_affinity* a, const int32_t* rows, int32_t n_rows, const uint8_t* donor_mask,
                        int32_t* out_idx, int32_t* out_ties);
PH_API int ph_affinity_build(ph_affinity* a, const int32_t* cells, int32_t k, const double* feats, int32_t F,
                      int use_copy_kernel, double radius, int strict, double* stats_out, double* dd_out);
PH_API int ph_affinity_matmat(ph_affinity* a, const double* X, int32_t c, double* Y);
PH_API int ph_affinity_matrix(ph_affinity* a, double* out); /* the k x k Laplacian (tests) */
/*
 * The same clustering step with the k x k matrices sharded by ROWS across the GPUs of a node (one process per GPU,
 * SURVEY.md section 8e / 8f): a context created with ph_affinity_create_rows holds at most max_rows rows.  Per cell set:
 *   ph_affinity_stats_rows     max of d_mat / c_mat over this rank's rows [row_lo, row_hi) (out3 = max d, max c, has_nan);
 *                              the host language takes the maximum over the ranks -> kw, kwc = 0.2 * max, r = max c
 *   ph_affinity_build_rows     this rank's rows of W and their sqrt-degrees (dd_block); the ranks' blocks concatenated are dd
 *   ph_affinity_laplacian_rows W rows -> Laplacian rows, given the whole dd
 *   ph_affinity_matmat_rows    Y[row_lo:row_hi, :] = L[row_lo:row_hi, :] @ X
 * Every entry, degree and product row is computed by one rank with the arithmetic of the single-GPU kernels (W is symmetric
 * bit for bit, so a row sum in column order IS the column sum in row order): the assembled results are bit-identical to
 * ph_affinity_build / ph_affinity_matmat.  radius must be 1 (the reference CLI's value: r = max c_mat).
 */
PH_API int ph_affinity_create_rows(const double* embedding, int32_t n, int32_t D, int device, int32_t max_rows, ph_affinity** out);
PH_API int ph_affinity_stats_rows(ph_affinity* a, const int32_t* cells, int32_t k, const double* feats, int32_t F,
                           int use_copy_kernel, int32_t row_lo, int32_t row_hi, double* out3);
PH_API int ph_affinity_build_rows(ph_affinity* a, int32_t F, double kw, double kwc, double r, int strict, double* dd_block,
                           int* has_nan);
PH_API int ph_affinity_laplacian_rows(ph_affinity* a, const double* dd_full);
PH_API int ph_affinity_matmat_rows(ph_affinity* a, const double* X, int32_t c, double* Y_block);
PH_API void ph_affinity_destroy(ph_affinity* a);

/*
 * K5 -- pairwise Euclidean distance of the embedding, `squareform(pdist(X))` (phertilizer.py:286):
 *   out[i*n + j] = sqrt(sum_d (X[i,d]-X[j,d])^2) computed by direct differences in float64, rows
 *   [row_begin, row_end) only (row-block sharding);  out has (row_end-row_begin) * n entries.
 * Does not need a handle (no sparse data involved).
 */
PH_API int ph_pairwise_dist(const double* X, int32_t n, int32_t D, int32_t row_begin, int32_t row_end, double* out,
                     int device, int on_device, void* stream);

/*
 * Device-resident block vectors for the eigensolver of the cut (sklearn spectral_embedding -> scipy LOBPCG, utils.py:131-150):
 * ten slots of k x 3 row-major doubles next to the Laplacian of the last ph_affinity_build / ph_affinity_laplacian_rows.
 *   ph_affinity_bv_set / _get   host <-> slot
 *   ph_affinity_bv_apply        slot dst = L @ slot src (whole Laplacian); row-block contexts return this rank's rows of the
 *                               product in host_block instead (the caller assembles the ranks' blocks and sets dst)
 *   ph_affinity_bv_gram         out[(a,i),(b,j)] = <left_a[:, i], right_b[:, j]>, row-major (3 nl) x (3 nr), fixed-order reduction
 *   ph_affinity_bv_combine      slot dst = sum_t slot_t @ C_t  (C_t 3 x 3 row-major, 1..4 terms; products and sums rounded
 *                               separately, in term order; coefficient 1 copies, 0 skips)
 * The LOBPCG recurrence then moves 3 x 3 ... 9 x 18 Gram matrices over PCIe per iteration instead of k x 3 operands.
 */
PH_API int ph_affinity_bv_set(ph_affinity* a, int slot, const double* host_k_by_3);
PH_API int ph_affinity_bv_get(ph_affinity* a, int slot, double* host_k_by_3);
PH_API int ph_affinity_bv_apply(ph_affinity* a, int src, int dst, double* host_block_or_null);
PH_API int ph_affinity_bv_gram(ph_affinity* a, const int32_t* left, int32_t nl, const int32_t* right, int32_t nr, double* out);
PH_API int ph_affinity_bv_combine(ph_affinity* a, int dst, const int32_t* slots, const double* coeffs, int32_t n_terms);

/*
 * K5 on the FP64 tensor cores: d^2 = |x_i|^2 + |x_j|^2 - 2 x_i.x_j with the inner products from DMMA (mma.sync m8n8k4
 * f64 -- tcgen05 has no fp64 type), pairs whose squared distance is below 1e-4 of |x_i|^2 + |x_j|^2 re-evaluated by
 * direct differences.  Every distance within 1e-9 relative of ph_pairwise_dist, NOT bit-identical to it: opt-in, for
 * high-dimensional embeddings (--no-umap with thousands of bins) where the direct kernel is bound by the fp64 pipe.
 * Device pointers only, enqueued on `stream`; sq_scratch: n doubles of device scratch.
 */
PH_API int ph_pairwise_dist_gram(const double* X, int32_t n, int32_t D, int32_t row_begin, int32_t row_end, double* out,
                          double* sq_scratch, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PHERTILIZER_B200_H */
