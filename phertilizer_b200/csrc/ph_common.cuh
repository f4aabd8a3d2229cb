// Shared declarations of the phertilizer_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "phertilizer_b200.h"

#define PH_MAX_PEERS 8     // GPUs of one NVLink/NVSwitch node
#define PH_MAX_SUB 32      // max sub-segments (dense codes x chunks) per line
#define PH_DMAX 4           // max "dense" (var,total) codes that get their own segment per line
#define PH_FIELD_BITS 10    // packed per-lane counter field width (3 labels x 10 bits, excluded label -> bit 30)
#define PH_GE 8             // entries per group: a lane consumes one 16-byte group per step (32-byte groups were not faster)
#define PH_FLUSH_GROUPS 127 // 8 increments per group -> a field gains <= 1016 < 1024 between flushes

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
//
// The minor axis is cut into chunks of CH = 2^chunk_bits - 16 indices, so that an entry is ONE 16-bit half-word
//   idx16 = 16 + (minor - chunk * CH)            (values 0..15 never occur; 0 is the padding entry)
// A line's entries with one of the D most frequent ("dense") (var,total) codes are stored as D * NCH sub-segments
// ordered by (code, chunk); every sub-segment is padded with idx16 = 0 to a whole number of 16-byte GROUPS (PH_GE = 8
// entries), so a line -- and any run of consecutive lines -- is one contiguous, 16-byte aligned stream of groups.
//   half    [n_groups * PH_GE]        the stream
//   sub_off [n_major * NSUB + 1]      first group of every sub-segment (NSUB = D * NCH)
// Entries with rarer codes go to the "tail": tail_words = (label-image address) | code << minor_bits, sorted by
// (code, address) inside a line; tail_off per line.
// The kernels keep a label image of the minor axis in shared memory with the matching layout: chunk c at byte
// c << chunk_bits, 16 "excluded" bytes (what the padding entries hit) followed by the labels of its indices.
struct PhOrient {
  uint16_t* half;
  uint32_t* sub_off;
  uint32_t* tail_words;  // [n_tail + 8]
  int64_t* tail_off;     // [n_major + 1]
  int64_t n_groups, n_tail;
  int32_t n_major, n_minor;
  int32_t D;           // dense codes, 0..PH_DMAX
  int32_t NCH, NSUB;   // chunks of the minor axis, sub-segments per line
  int32_t chunk_bits;  // 16 in production (PH_CHUNK_BITS overrides it for tests)
  int32_t minor_bits;  // bits of the label-image address inside a tail word
  int32_t lab_bytes;   // size of the label image
  int32_t batch;       // lines per warp batch (8 or 32)
  int32_t labels_in_smem;
  // Second copy of the dense stream, "batched-4" (pass kernel LAY = 3): lines sorted by the shape of their sub-segments
  // and taken 8 at a time; inside a batch every sub-segment is padded to the batch's longest in units of 4 groups, and
  // a warp step is 32 consecutive groups = [line a = 0..7][lane c = 0..3] holding group 4 t + c of line a.  Four lanes
  // stream one line, sub-segment boundaries fall on the same step for the whole warp, and with the round order of line a
  // rotated by a banks the 32 gathers of a step hit 32 different banks.
  uint16_t* half4;       // [n_steps4 * 32 * PH_GE] or null
  uint32_t* bstep_off;   // [n_batches4 * NSUB + 1] first warp step of every (batch, sub-segment)
  int32_t* perm4;        // [n_batches4 * 8] line of every batch slot, -1 = none
  int64_t n_steps4;
  int32_t n_batches4;
};

struct ph_handle {
  int device;
  int sm_count;
  int max_smem;  // opt-in dynamic shared memory per block
  int smem_per_sm;
  int32_t n_cells, n_snvs, n_codes;
  int64_t nnz;
  PhOrient by_snv, by_cell;
  uint8_t* d_gauss;  // scratch of ph_gauss_node_ll (grown on demand)
  size_t gauss_bytes;
  double* d_emb;     // [n_cells * emb_D] optional cached embedding (ph_set_embedding)
  int emb_D;
  double* d_L0;      // [n_codes]
  double* d_L1;      // [n_codes]
  uint8_t* d_varpos; // [n_codes]
  int64_t device_bytes;
  int chunk_bits, bank_order, ctas_per_sm, aligned_image, zero_copy_out, batch4;
  double dense_min_avg, dense_min_share;  // a (var,total) code gets its own segment when it has >= this many entries per line on average
  // scratch (device)
  uint8_t* d_lab_minor;  // [max(n,m)] staged labels of the minor axis
  uint8_t* d_lab_major;  // [max(n,m)]
  uint8_t* d_lab_img;    // label image in global memory (only when it does not fit in shared memory)
  size_t lab_img_bytes;
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

// ---- peer-memory exchange helpers shared by ph_pass.cu and ph_peer.cu
constexpr size_t kPeerHdrOff = 2048;   // window: flagA[world] at +0, flagB[world] at +1024, SNV block headers {offset, count, B, record bytes}[world] at +2048
constexpr size_t kPeerRecStride = 128; // bytes reserved per 8 lines of a rank's record region (a record of B lines takes <= 16 B per line)
constexpr long long kSpinLimit = 6000000000ll;  // ~3 s of SM clocks: a lost peer must not hang the GPU

// what the collect phase of an SNV-sharded step needs (ph_peer.cu fills it, the pass kernel runs it)
struct PhCollect {
  const uint8_t* win;     // this rank's window
  size_t off_rec[2];      // the two record regions (parity of the step)
  int cap_recs, max_lines, C;
  uint8_t* label_out;
  int32_t* nobs_out;
  unsigned* done;         // completion counter of the collect phase (self-resetting)
  unsigned* timeout_flag;
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// thread 0 of the block polls `world` flags until each has reached `epoch`; ends with a block barrier
__device__ __forceinline__ void wait_flags(const uint32_t* flags, int world, uint32_t epoch, unsigned* timeout_flag) {
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int q = 0; q < world; ++q) {
      // epochs only grow; signed distance keeps the comparison valid across a 32-bit wrap
      while ((int32_t)(ld_acquire_sys(flags + q) - epoch) < 0) {
        if (clock64() - t0 > kSpinLimit) {
          *timeout_flag = 1;
          break;
        }
        __nanosleep(64);
      }
    }
  }
  __syncthreads();
}

// Unpacks the records of every rank's SNV block (ph_pass.cu, PassArgs::rec_dst: B label bytes, then B * C int32 counts
// from byte 16; B and the record size come with the block's header) into the caller's arrays.  s_hdr: 4 * world ints.
__device__ __forceinline__ void unpack_records(const PhCollect& c, int world, uint32_t epoch, int* s_hdr, int tid, int nt) {
  if (threadIdx.x < 4 * world) s_hdr[threadIdx.x] = (int)__ldcv(reinterpret_cast<const uint32_t*>(c.win + kPeerHdrOff) + threadIdx.x);
  __syncthreads();
  const uint8_t* rec = c.win + c.off_rec[epoch & 1u];
  const int cap_lines = c.cap_recs * 8;
  for (int idx = tid; idx < world * cap_lines; idx += nt) {
    const int r = idx / cap_lines, j = idx - r * cap_lines;
    const int off = s_hdr[4 * r], cnt = s_hdr[4 * r + 1], B = s_hdr[4 * r + 2], rb = s_hdr[4 * r + 3];
    if (B <= 0 || rb <= 0 || j >= cnt || off < 0 || off + j >= c.max_lines) continue;  // (a header that never arrived: timed-out wait)
    const int q = j / B, s = j - q * B;
    const uint8_t* b = rec + ((size_t)r * c.cap_recs) * kPeerRecStride + (size_t)q * rb;
    c.label_out[off + j] = __ldcv(b + s);
    const uint32_t* cn = reinterpret_cast<const uint32_t*>(b + 16) + s * c.C;
    for (int k = 0; k < c.C; ++k) c.nobs_out[(size_t)(off + j) * c.C + k] = (int32_t)__ldcv(cn + k);
  }
}
#endif

static inline int ph_ceil_log2(uint64_t x) {
  int b = 0;
  while ((1ull << b) < x) ++b;
  return b;
}
