// Shared declarations of the phertilizer_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "phertilizer_b200.h"

#define PH_DMAX 4           // max "dense" (var,total) codes that get their own segment per line
#define PH_BLOCK 1024       // threads per CTA of the pass kernels (one CTA per SM)
#define PH_FIELD_BITS 10    // packed per-lane counter field width (3 labels x 10 bits, excluded label -> bit 30)
#define PH_CHUNK_ITERS 255  // 4 increments/iteration -> a field gains <= 1020 < 1024 per chunk

extern thread_local char ph_err_msg[512];
void ph_count_launch(int n = 1);

#define PH_FAIL(code, ...)                                   \
  do {                                                       \
    snprintf(ph_err_msg, sizeof(ph_err_msg), __VA_ARGS__);   \
    return (code);                                           \
  } while (0)

#define PH_CUDA(expr)                                                                               \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      snprintf(ph_err_msg, sizeof(ph_err_msg), "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e),  \
               __FILE__, __LINE__, cudaGetErrorString(_e));                                         \
      return PH_ERR_CUDA;                                                                           \
    }                                                                                               \
  } while (0)

// One orientation of the observation store: "lines" are SNVs (by_snv, CSC-like) or cells (by_cell, CSR-like).
// Entries of a line are sorted by (segment, word): segments 0..D-1 hold the D most frequent (var,total) codes
// (word = minor index only), segment D ("tail") holds everything else with word = minor | code << minor_bits.
struct PhOrient {
  uint32_t* words;    // [nnz + 8]   (padded so 16-byte loads may run past the end)
  int64_t* seg_off;   // [n_major * (D+1) + 1]
  int32_t n_major, n_minor;
  int32_t D;          // dense codes, 0..PH_DMAX
  int32_t minor_bits; // bits of the minor index inside a tail word
  int32_t group;      // lanes cooperating on one line (4, 8, 16, 32)
  int32_t labels_in_smem;
};

struct ph_handle {
  int device;
  int sm_count;
  int max_smem;  // opt-in dynamic shared memory per block
  int smem_per_sm;
  int32_t n_cells, n_snvs, n_codes;
  int64_t nnz;
  PhOrient by_snv, by_cell;
  double* d_L0;      // [n_codes]
  double* d_L1;      // [n_codes]
  uint8_t* d_varpos; // [n_codes]
  int64_t device_bytes;
  double dense_min_avg;  // a (var,total) code gets its own segment when it has >= this many entries per line on average
  // scratch (device)
  uint8_t* d_lab_minor;  // [max(n,m)] staged labels of the minor axis
  uint8_t* d_lab_major;  // [max(n,m)]
  int32_t* d_list;       // [max(n,m)] staged line list
  double* d_f64;         // [max(n,m) * 2 * 3 ... ] table scratch, see PH_SCRATCH_F64
  int32_t* d_i32;        // table scratch
  uint8_t* d_u8;         // [max(n,m)] label outputs
  double* d_small;       // [4096] small outputs (block sums, per-node)
  uint8_t* d_map;        // [65*65] class / member maps
  unsigned* d_counter;   // batch dispenser of the pass kernels
  // pinned host staging
  uint8_t* h_stage;
  size_t h_stage_bytes;
  cudaStream_t stream;   // private stream for host-pointer calls
};

static inline int ph_ceil_log2(uint64_t x) {
  int b = 0;
  while ((1ull << b) < x) ++b;
  return b;
}
