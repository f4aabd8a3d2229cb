/*
 * TEST INFRASTRUCTURE ONLY -- sparse CPU restatement (plain C + OpenMP) of the reference's likelihood reductions.
 *
 * The reference evaluates `M[np.ix_(cells, muts)].<reduce>` on DENSE n x m numpy arrays (phertilizer/data.py:8-20);
 * at the benchmark shapes those arrays do not fit in memory (100k x 200k float64 = 160 GB each), so the
 * same sums are restated here over CSC / CSR arrays.  Unobserved entries are exact zeros in the reference and
 * contribute nothing to any of these reductions, so skipping them is exact for the integer tables; the float sums
 * run per SNV in ascending cell order -- the order numpy's axis-0 reduce uses.
 *
 * Pinned by tests/test_oracle_sparse.py against oracle/oracle_np.py (itself pinned against the reference's recorded
 * outputs in tests/golden/).  Used by tests/ as a mid-size checker and by bench.py as the CPU baseline
 * ("kind": "port").  Nothing under phertilizer_b200/ links or loads this file.
 *
 *   gcc -O3 -march=native -fopenmp -shared -fPIC -o oracle/_build/liboracle_sparse.so oracle/oracle_sparse.c
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the baseline legs of bench.py ask for all host cores explicitly */
void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Per-SNV table over cell clusters, CSC input (colptr[m+1], row[nnz], code[nnz]); label: 0 = excluded, 1..C.
 * linear_split.py:186-188, 224, 229-230; branching_split.py:316-317, 332-337.  Outputs [m*C]. */
void orc_snv_table(int32_t m, const int64_t* colptr, const int32_t* row, const uint16_t* code, const double* L0,
                   const double* L1, const uint8_t* varpos, const uint8_t* cell_label, int32_t C, double* S0,
                   double* S1, int32_t* Nobs, int32_t* Nvar) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int32_t j = 0; j < m; ++j) {
    double s0[8] = {0}, s1[8] = {0};
    int32_t no[8] = {0}, nv[8] = {0};
    for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
      const int l = cell_label[row[p]];
      if (l) {
        const int c = l - 1, k = code[p];
        s0[c] += L0[k];
        s1[c] += L1[k];
        no[c] += 1;
        nv[c] += varpos[k];
      }
    }
    for (int c = 0; c < C; ++c) {
      S0[(int64_t)j * C + c] = s0[c];
      S1[(int64_t)j * C + c] = s1[c];
      Nobs[(int64_t)j * C + c] = no[c];
      Nvar[(int64_t)j * C + c] = nv[c];
    }
  }
}

/* Per-cell table over SNV clusters, CSR input (rowptr[n+1], col[nnz], code[nnz]); label: 0 = excluded, 1..Q.
 * linear_split.py:301-304 (tmb features), 400-410 (check_metrics); branching_split.py:147-173 (calc_tmb), 395-411;
 * clonal_tree.py:1140 (Seed.strip).  Sums run per cell in ascending SNV order (numpy's axis-1 reduce).  Outputs [n*Q]. */
void orc_cell_table(int32_t n, const int64_t* rowptr, const int32_t* col, const uint16_t* code, const double* L0,
                    const double* L1, const uint8_t* varpos, const uint8_t* snv_label, int32_t Q, double* A0,
                    double* A1, int32_t* Nobs, int32_t* Nvar) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int32_t i = 0; i < n; ++i) {
    double s0[8] = {0}, s1[8] = {0};
    int32_t no[8] = {0}, nv[8] = {0};
    for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
      const int l = snv_label[col[p]];
      if (l) {
        const int q = l - 1, k = code[p];
        s0[q] += L0[k];
        s1[q] += L1[k];
        no[q] += 1;
        nv[q] += varpos[k];
      }
    }
    for (int q = 0; q < Q; ++q) {
      A0[(int64_t)i * Q + q] = s0[q];
      A1[(int64_t)i * Q + q] = s1[q];
      Nobs[(int64_t)i * Q + q] = no[q];
      Nvar[(int64_t)i * Q + q] = nv[q];
    }
  }
}

/* linear_split.py:203-237 without the quantile gate: label 2 (mutsB) if mean like0 >= mean like1 else 1 (mutsA). */
void orc_assign_linear(int32_t m, const int64_t* colptr, const int32_t* row, const uint16_t* code, const double* L0,
                       const double* L1, const uint8_t* cell_label, double lenA, uint8_t* label_out, int32_t* nobs_out) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int32_t j = 0; j < m; ++j) {
    double s0 = 0.0, s1 = 0.0;
    int32_t no = 0;
    for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
      if (cell_label[row[p]] == 1) {
        const int k = code[p];
        s0 += L0[k];
        s1 += L1[k];
        no += 1;
      }
    }
    label_out[j] = (s0 / lenA >= s1 / lenA) ? 2 : 1;
    nobs_out[j] = no;
  }
}

/* branching_split.py:293-348 without the quantile gates. */
void orc_assign_branching(int32_t m, const int64_t* colptr, const int32_t* row, const uint16_t* code, const double* L0,
                          const double* L1, const uint8_t* cell_label, uint8_t* label_out, int32_t* nobs_out) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int32_t j = 0; j < m; ++j) {
    double s0[2] = {0, 0}, s1[2] = {0, 0};
    int32_t no[2] = {0, 0};
    for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
      const int l = cell_label[row[p]];
      if (l == 1 || l == 2) {
        const int k = code[p];
        s0[l - 1] += L0[k];
        s1[l - 1] += L1[k];
        no[l - 1] += 1;
      }
    }
    const double cA = s1[0] + s0[1], cB = s0[0] + s1[1], cC = s1[0] + s1[1];
    double mx = cA > cB ? cA : cB;
    mx = mx > cC ? mx : cC;
    label_out[j] = (cA == mx) ? 1 : ((cB == mx) ? 2 : 3);
    nobs_out[2 * (int64_t)j] = no[0];
    nobs_out[2 * (int64_t)j + 1] = no[1];
  }
}

/* clonal_tree.py:932-944: per node k, sum over its cells of (present(k, node(j)) ? like1 : like0), CSR input. */
void orc_tree_loglik(int32_t n, const int64_t* rowptr, const int32_t* col, const uint16_t* code, const double* L0,
                     const double* L1, const uint8_t* cell_node, const uint8_t* snv_node, const uint8_t* present,
                     int32_t K, double* per_cell) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int32_t i = 0; i < n; ++i) {
    const int k = cell_node[i];
    double acc = 0.0;
    if (k < K) {
      for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
        const int q = snv_node[col[p]];
        const int pr = (q < K) ? present[k * K + q] : 0;
        acc += pr ? L1[code[p]] : L0[code[p]];
      }
    }
    per_cell[i] = acc;
  }
}
