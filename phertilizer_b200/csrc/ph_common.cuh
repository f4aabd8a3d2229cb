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
#define PH_PAD 128          // "excluded" bytes at the start of every chunk of the label image: one 4-byte word per shared-memory
                            // bank, so that a padding entry can sit in ANY bank (idx16 = 4 * bank) without colliding with the
                            // real entries of its warp step (ph_build.cu, padded rounds); idx16 = 0 is the plain padding entry

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
// The minor axis is cut into chunks of CH = 2^chunk_bits - PH_PAD indices, so that an entry is ONE 16-bit half-word
//   idx16 = PH_PAD + (minor - chunk * CH)        (values below PH_PAD are padding entries: they gather an "excluded" label)
// A line's entries with one of the D most frequent ("dense") (var,total) codes are stored as D * NCH sub-segments
// ordered by (code, chunk); every sub-segment is padded with idx16 = 0 to a whole number of 16-byte GROUPS (PH_GE = 8
// entries), so a line -- and any run of consecutive lines -- is one contiguous, 16-byte aligned stream of groups.
//   half    [n_groups * PH_GE]        the stream
//   sub_off [n_major * NSUB + 1]      first group of every sub-segment (NSUB = D * NCH)
// Entries with rarer codes go to the "tail": tail_words = (label-image address) | code << minor_bits, sorted by
// (code, address) inside a line; tail_off per line.
// The kernels keep a label image of the minor axis in shared memory with the matching layout: chunk c at byte
// c << chunk_bits, PH_PAD "excluded" bytes (what the padding entries hit) followed by the labels of its indices.
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
  uint32_t* part_acc;    // [n_batches4 * 8 * D * PH_MAX_FAST_CLUSTERS] split-tail counters (ph_pass.cu PassArgs), zero at rest
  unsigned* part_tick;   // [n_batches4] split-tail tickets, zero at rest
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
  int chunk_bits, bank_order, ctas_per_sm, aligned_image, zero_copy_out, batch4, round_fill, phase_chunks, lpt_order, split_sh, split_tail, item_batches;
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
  uint32_t* d_partial;   // phased passes: per-line (code, cluster) counters handed from phase to phase (grown on demand)
  size_t partial_bytes;
  // pinned host staging
  uint8_t* h_stage;
  size_t h_stage_bytes;
  cudaStream_t stream;   // private stream for host-pointer calls
};

// ---- peer-memory exchange helpers shared by ph_pass.cu and ph_peer.cu
constexpr long long kSpinLimit = 6000000000ll;  // ~3 s of SM clocks: a lost peer must not hang the GPU (PH_PEER_TIMEOUT_MS overrides)

// SNV-sharded steps (ph_peer.cu, ph_pass.cu): results travel as SELF-VALIDATING 8-byte words  epoch << 32 | payload
// (the step counter rides in the upper half of every word, NCCL-LL style), so a receiver polls the data itself: no
// system-scope fence, no flag round trip, no header.  Per line of a rank's work order (slot = batch * B + lane):
//   word 0 = global SNV id << 8 | label      (id 0xFFFFFF = empty slot)
//   word 1 = Nobs[A]        word 2 = Nobs[B] (branching only)
// laid out  ll[parity][word][source rank][slot]  in every rank's window, so the B lanes of a batch store B consecutive
// words per (peer, word): one coalesced 64-byte NVLink write.  hdr[parity][source rank] = epoch << 32 | slots used.
// Steps alternate between the two parities: a fast rank may already push step e+1 while a peer still polls step e (it
// cannot reach step e+2 before that peer has pushed step e+1, i.e. finished polling step e).
constexpr int kLLWords = 3;
struct PhLL {
  uint8_t* dst[PH_MAX_PEERS];      // per peer window: &ll[0][0][this rank][0]
  uint64_t* hdr_dst[PH_MAX_PEERS]; // per peer window: &hdr[0][this rank]
  const uint8_t* own;              // own window: &ll[0][0][0][0]
  const uint64_t* hdr_own;         // own window: &hdr[0][0]
  size_t word_stride, par_stride;  // bytes between ll[.][w] and ll[.][w+1] / between the parities
  int cap_slots;                   // slots reserved per source rank
  int max_lines;                   // total SNVs (bounds the unpacked ids)
  uint8_t* label_out;
  int32_t* nobs_out;
  unsigned* done;                  // completion counter of the collect phase (self-resetting)
  unsigned* timeout_flag;          // [0] sticky flag; [1..7] what the first wait that gave up was waiting for (ph_peer_debug)
  long long spin_limit;            // SM clocks a wait may take
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t ld_relaxed_sys_u64(const void* p) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(void* p, uint64_t v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// spins until the word at p carries `epoch` in its upper half; returns the payload (0 and *ok = false on a timeout).
// The first wait to give up records what it was waiting for: {kind, source rank, slot, epoch wanted, epoch seen, payload}.
__device__ __forceinline__ uint32_t ll_wait(const void* p, uint32_t epoch, long long t0, long long limit, unsigned* timeout_flag,
                                            bool* ok, unsigned kind, unsigned src, unsigned slot) {
  for (;;) {
    const uint64_t w = ld_relaxed_sys_u64(p);
    if ((uint32_t)(w >> 32) == epoch) return (uint32_t)w;
    if (clock64() - t0 > limit) {
      if (atomicExch(timeout_flag, 1u) == 0u) {
        timeout_flag[1] = kind; timeout_flag[2] = src; timeout_flag[3] = slot; timeout_flag[4] = epoch;
        timeout_flag[5] = (unsigned)(w >> 32); timeout_flag[6] = (unsigned)w;
      }
      *ok = false;
      return 0u;
    }
    __nanosleep(32);
  }
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
#endif

static inline int ph_ceil_log2(uint64_t x) {
  int b = 0;
  while ((1ull << b) < x) ++b;
  return b;
}
