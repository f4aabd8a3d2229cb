// Device-resident observation store: COO -> two code-segmented sparse copies (by SNV, by cell).
// Replaces the dense `sparse_matrix()` unstack of the reference (phertilizer/phertilizer.py:153-161, 295-306).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <stdlib.h>

#include <new>

#include "ph_common.cuh"

thread_local char ph_err_msg[512] = "";
static int64_t g_launches = 0;
void ph_count_launch(int n) { g_launches += n; }

extern "C" const char* ph_last_error(void) { return ph_err_msg; }
extern "C" int ph_abi_version(void) { return PH_ABI_VERSION; }
extern "C" int64_t ph_kernel_launches(void) { return g_launches; }
extern "C" int ph_device_count(int* out) {
  if (!out) PH_FAIL(PH_ERR_ARG, "null pointer");
  PH_CUDA(cudaGetDeviceCount(out));
  return PH_OK;
}

namespace {

__global__ void k_code_hist(const uint16_t* __restrict__ code, int64_t nnz, unsigned long long* hist, int n_codes,
                            int* bad) {
  __shared__ unsigned int sh[64];  // the hot codes are counted per block first
  if (threadIdx.x < 64) sh[threadIdx.x] = 0;
  __syncthreads();
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) {
    int c = code[i];
    if (c >= n_codes) {
      *bad = 1;
    } else if (c >= 64) {
      atomicAdd(&hist[c], 1ull);
    } else {
      atomicAdd(&sh[c], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < 64 && threadIdx.x < n_codes && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// duplicate check: one key per observation, (SNV, cell); equal neighbours after the sort = the pair is listed twice --
// whatever the two entries' (var,total) codes are (the reference's `unstack` raises on such input, phertilizer.py:153-161)
__global__ void k_pair_keys(const int32_t* __restrict__ snv, const int32_t* __restrict__ cell, int64_t nnz, uint64_t* __restrict__ keys) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) keys[i] = ((uint64_t)(uint32_t)snv[i] << 32) | (uint32_t)cell[i];
}
__global__ void k_adjacent_equal(const uint64_t* __restrict__ keys, int64_t nnz, int* dup) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x + 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride)
    if (keys[i] == keys[i - 1]) *dup = 1;
}

// key = segment id << 32 | payload.  Segment id = line * (NSUB + 1) + s; s = d * NCH + chunk for a dense code d,
// s = NSUB for the tail.  Payload: idx16 (dense; optionally with the shared-memory bank of its label above it so that
// the sort groups a sub-segment's entries by bank) or  minor | code << minor_bits  (tail).
__global__ void k_make_keys(const int32_t* __restrict__ major, const int32_t* __restrict__ minor,
                            const uint16_t* __restrict__ code, int64_t nnz, int D, int NCH, int CH, int chunk_bits,
                            int minor_bits, int n_major, int n_minor, int bank_order, uint64_t* __restrict__ keys, int* bad) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const uint32_t NSUB = (uint32_t)(D * NCH);
  for (; i < nnz; i += stride) {
    uint32_t mj = (uint32_t)major[i], mn = (uint32_t)minor[i], c = code[i];
    if (mj >= (uint32_t)n_major || mn >= (uint32_t)n_minor) {
      *bad = 2;
      mj = 0;
      mn = 0;
    }
    uint32_t s, payload;
    const uint32_t ch = mn / (uint32_t)CH;
    const uint32_t idx = (uint32_t)PH_PAD + (mn - ch * (uint32_t)CH);
    if (c < (uint32_t)D) {
      s = c * (uint32_t)NCH + ch;
      payload = idx | (bank_order ? (((idx >> 2) & 31u) << 16) : 0u);
    } else {
      s = NSUB;
      payload = ((ch << chunk_bits) | idx) | (c << minor_bits);  // label-image address | code
    }
    keys[i] = ((uint64_t)(mj * (NSUB + 1u) + s) << 32) | payload;
  }
}

// seg_off[s] = first position whose segment id >= s  (s in 0..nseg)
__global__ void k_seg_offsets(const uint64_t* __restrict__ keys, int64_t nnz, int64_t nseg, int64_t* __restrict__ off) {
  int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s > nseg) return;
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)(keys[mid] >> 32) < s) lo = mid + 1; else hi = mid;
  }
  off[s] = lo;
}

// Padded rounds.  In round order (k_round_keys) round k of a sub-segment holds the k-th entry of every bank that has one;
// the first rounds are nearly full, the last ones ragged, and once a round is short the entry index loses its phase with
// the bank, so the 32 lanes of a gather collide (2.17 wavefronts per LDS.U8 measured on C3).  The first K rounds -- those
// in which at least `min_fill` of the 32 banks still hold an entry -- are therefore laid out on a fixed grid: entry of
// (rotated) bank b of round k at position 32 k + b, the banks without an entry get a padding entry THAT SITS IN THE
// SAME BANK (idx16 = 4 * bank: the excluded region of the label image has one word per bank).  Those gathers are
// conflict free; the rest of the sub-segment follows compacted as before.
//   plan[2 * sub]     = K, the padded rounds
//   plan[2 * sub + 1] = R, the entries those rounds hold       (padded length = 32 K + cnt - R)
// `sorted`: keys ordered by (segment, bank, idx16).
__global__ void k_round_plan(const uint64_t* __restrict__ sorted, const int64_t* __restrict__ off, int64_t n_major, int NSUB,
                             int min_fill, int32_t* __restrict__ plan) {
  const int64_t sub = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (sub >= n_major * NSUB) return;
  const int64_t line = sub / NSUB;
  const int64_t seg = line * (NSUB + 1) + (sub - line * NSUB);
  const int64_t b0 = off[seg], b1 = off[seg + 1];
  int K = 0, R = 0;
  if (min_fill <= 32 && b1 - b0 >= min_fill) {
    int cnt[32];
    int64_t lo = b0;
    for (int b = 0; b < 32; ++b) {  // first position of bank b + 1 inside the sub-segment
      const uint64_t want = ((uint64_t)seg << 16) | (uint64_t)(b + 1);
      int64_t l = lo, h = b1;
      while (l < h) {
        const int64_t mid = (l + h) >> 1;
        if ((sorted[mid] >> 16) < want) l = mid + 1; else h = mid;
      }
      cnt[b] = (int)(l - lo);
      lo = l;
    }
    for (;;) {  // the occupancy of round K only falls with K
      int nk = 0;
#pragma unroll
      for (int b = 0; b < 32; ++b) nk += cnt[b] > K ? 1 : 0;
      if (nk < min_fill) break;
      R += nk;
      ++K;
    }
  }
  plan[2 * sub] = K;
  plan[2 * sub + 1] = R;
}

// per sub-segment: groups of PH_GE entries (rounded up, padded rounds included); per line: tail entries
__global__ void k_seg_sizes(const int64_t* __restrict__ off, int64_t n_major, int NSUB, const int32_t* __restrict__ plan,
                            int64_t* __restrict__ groups, int64_t* __restrict__ tails) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nseg = n_major * (NSUB + 1);
  if (k >= nseg) return;
  const int64_t line = k / (NSUB + 1);
  const int s = (int)(k - line * (NSUB + 1));
  int64_t cnt = off[k + 1] - off[k];
  if (s < NSUB) {
    const int64_t sub = line * NSUB + s;
    if (plan) cnt += 32 * (int64_t)plan[2 * sub] - plan[2 * sub + 1];
    groups[sub] = (cnt + PH_GE - 1) / PH_GE;
  } else {
    tails[line] = cnt;
  }
}

__global__ void k_narrow(const int64_t* __restrict__ in, int64_t n, uint32_t* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)in[i];
}

// bank_order: a sub-segment's entries are laid out in ROUNDS -- round k holds the k-th entry of every shared-memory
// bank (of the label the entry will gather), in bank order -- and consecutive runs of 32 entries of that sequence become
// the 32 lanes of one LDS of the pass kernel (group g = lane, same slot e in each group).  A full round has one entry per
// bank, so those LDS are conflict free wherever the sub-segment starts; only the ragged last rounds can collide.
//
// k_round_keys: keys sorted by (segment, bank, idx16) -> keys (segment, rank within bank, bank, idx16).
// pos4 != null: the banks of line l are visited starting from bank (pos4[l] & 7), its slot in the batch of the
// batched-4 copy (any rotation keeps a full round conflict free for the flat copy as well).
__global__ void k_round_keys(const uint64_t* __restrict__ sorted, int64_t nnz, const int64_t* __restrict__ off, int NSUB,
                             const int32_t* __restrict__ pos4, uint64_t* __restrict__ out, int* dup) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) {
    const uint64_t k = sorted[i];
    if (i > 0 && sorted[i - 1] == k) *dup = 1;  // the same (line, index) twice
    const int64_t seg = (int64_t)(k >> 32);
    if ((int)(seg % (NSUB + 1)) == NSUB) {  // tail entry: unchanged
      out[i] = k;
      continue;
    }
    const uint64_t kb = k >> 16;  // (segment, bank)
    int64_t lo = off[seg], hi = i;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((sorted[mid] >> 16) < kb) lo = mid + 1; else hi = mid;
    }
    const uint64_t r = (uint64_t)(i - lo);  // < 2048: a bank holds 2048 bytes of one chunk of the label image
    uint64_t low = k & 0x1FFFFFull;         // bank << 16 | idx16
    if (pos4) {
      const uint64_t rot = (uint64_t)(pos4[seg / (NSUB + 1)] & 7);
      low = (low & 0xFFFFull) | ((((low >> 16) - rot) & 31ull) << 16);
    }
    out[i] = (k & 0xFFFFFFFF00000000ull) | (r << 21) | low;
  }
}

// Position of the r-th entry (in round order) inside its padded sub-segment of ng = ceil(cnt / PH_GE) groups: run j
// of 32 entries of the round sequence becomes slot (j % PH_GE) of the 32 groups of block (j / PH_GE).
__device__ __forceinline__ int64_t place(int64_t r, int64_t cnt, int bank_order) {
  if (!bank_order) return r;
  const int64_t ng = (cnt + PH_GE - 1) / PH_GE;
  const int64_t tf = ng >> 5;  // full blocks of 32 groups
  constexpr int64_t BE = 32 * PH_GE;  // entries of a full block
  if (r < BE * tf) {
    const int64_t t = r / BE, j = (r % BE) >> 5, ln = r & 31;
    return ((t << 5) + ln) * PH_GE + j;
  }
  const int64_t r2 = r - BE * tf, gl = ng - (tf << 5);  // ragged last block: gl < 32 groups
  const int64_t e = r2 / gl, ln = r2 - e * gl;
  return ((tf << 5) + ln) * PH_GE + e;
}

// where padded position r (of a sub-segment of cnt padded entries) lives in the two copies of the stream
__device__ __forceinline__ void put_entry(uint16_t v, int64_t r, int64_t cnt, int64_t line, int s, int NSUB, int bank_order,
                                          const uint32_t* __restrict__ sub_off, uint16_t* __restrict__ half,
                                          const int32_t* __restrict__ pos4, const uint32_t* __restrict__ bstep_off,
                                          uint16_t* __restrict__ half4) {
  half[(int64_t)sub_off[line * NSUB + s] * PH_GE + place(r, cnt, bank_order)] = v;
  if (half4) {
    // batched-4 copy: position r of the line's sub-segment; 32 positions are one warp step (4 lanes x 8 slots),
    // lane c = (r % 32) / 8 of line slot a, slot r % 8
    const int p = pos4[line];
    const int64_t b = p >> 3, a = p & 7;
    const int64_t step = (int64_t)bstep_off[b * NSUB + s] + (r >> 5);
    half4[((step << 5) + a * 4 + ((r >> 3) & 3)) * PH_GE + (r & 7)] = v;
  }
}

// the padding entries of the padded rounds (k_round_plan): position 32 k + b holds idx16 = 4 * bank, i.e. the excluded
// word of the very bank the position is reserved for; the real entries overwrite theirs afterwards (k_scatter).
// One warp per sub-segment.
__global__ void k_fill_rounds(const int64_t* __restrict__ off, int64_t n_major, int NSUB, const int32_t* __restrict__ plan,
                              const uint32_t* __restrict__ sub_off, int bank_order, uint16_t* __restrict__ half,
                              const int32_t* __restrict__ pos4, const uint32_t* __restrict__ bstep_off,
                              uint16_t* __restrict__ half4) {
  const int64_t sub = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int b = threadIdx.x & 31;
  if (sub >= n_major * NSUB) return;
  const int K = plan[2 * sub];
  if (K == 0) return;
  const int64_t line = sub / NSUB;
  const int s = (int)(sub - line * NSUB);
  const int64_t seg = line * (NSUB + 1) + s;
  const int64_t cnt = off[seg + 1] - off[seg] + 32 * (int64_t)K - plan[2 * sub + 1];
  const int rot = pos4 ? (pos4[line] & 7) : 0;
  const uint16_t v = (uint16_t)(4 * ((b + rot) & 31));
  for (int k = 0; k < K; ++k) put_entry(v, 32 * (int64_t)k + b, cnt, line, s, NSUB, bank_order, sub_off, half, pos4, bstep_off, half4);
}

__global__ void k_scatter(const uint64_t* __restrict__ keys, int64_t nnz, const int64_t* __restrict__ off, int NSUB,
                          const int32_t* __restrict__ plan, const uint32_t* __restrict__ sub_off,
                          const int64_t* __restrict__ tail_off, int bank_order, uint16_t* __restrict__ half,
                          uint32_t* __restrict__ tail_words, const int32_t* __restrict__ pos4,
                          const uint32_t* __restrict__ bstep_off, uint16_t* __restrict__ half4, int* dup) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) {
    const uint64_t k = keys[i];
    if (i > 0 && keys[i - 1] == k) *dup = 1;
    const int64_t seg = (int64_t)(k >> 32);
    const int64_t line = seg / (NSUB + 1);
    const int s = (int)(seg - line * (NSUB + 1));
    int64_t r = i - off[seg];  // rank in round order
    if (s < NSUB) {
      int64_t cnt = off[seg + 1] - off[seg];
      if (plan) {
        // keys in round order are (segment, round, rotated bank, idx16): padded rounds sit on the 32-wide grid
        const int64_t K = plan[2 * (line * NSUB + s)], R = plan[2 * (line * NSUB + s) + 1];
        const int64_t round = (int64_t)((k >> 21) & 0x7FFull), bank = (int64_t)((k >> 16) & 31ull);
        cnt += 32 * K - R;
        r = round < K ? 32 * round + bank : 32 * K + (r - R);
      }
      put_entry((uint16_t)k, r, cnt, line, s, NSUB, bank_order, sub_off, half, pos4, bstep_off, half4);
    } else {
      tail_words[tail_off[line] + r] = (uint32_t)k;
    }
  }
}

// ---- batched-4 copy: lines sorted by the shape (group counts) of their sub-segments
__global__ void k_iota(int32_t* v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}
__global__ void k_shape_key(const int64_t* __restrict__ groups, const int32_t* __restrict__ order, int n, int NSUB, int q,
                            uint32_t* __restrict__ key) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) key[i] = (uint32_t)groups[(int64_t)order[i] * NSUB + q];
}
// perm4[i] = i-th line in shape order (-1 beyond n), pos4[line] = its rank
__global__ void k_perm_pos(const int32_t* __restrict__ order, int n, int n_pad, int32_t* __restrict__ perm4,
                           int32_t* __restrict__ pos4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  if (i < n) {
    perm4[i] = order[i];
    pos4[order[i]] = i;
  } else {
    perm4[i] = -1;
  }
}
// Batches in the order the pass hands them out: longest first.  The dispenser of k_pass counts up, so the last batches
// drawn -- the ones whose length decides how long the other warps of the grid wait at the end -- are the shortest
// (longest-processing-time-first scheduling).  key[b] = ~(warp steps of batch b): an ascending sort puts the longest first.
__global__ void k_batch_len_key(const uint32_t* __restrict__ steps, int n_batches, int NSUB, uint32_t* __restrict__ key,
                                int32_t* __restrict__ id) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_batches) return;
  uint32_t t = 0;
  for (int q = 0; q < NSUB; ++q) t += steps[(int64_t)b * NSUB + q];
  key[b] = ~t;
  id[b] = b;
}
// perm4 / pos4 with whole batches (blocks of 8 consecutive lines of `order`) permuted by `border`
__global__ void k_perm_blocks(const int32_t* __restrict__ order, int n, const int32_t* __restrict__ border, int n_batches,
                              int32_t* __restrict__ perm4, int32_t* __restrict__ pos4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_batches * 8) return;
  const int src = border[i >> 3] * 8 + (i & 7);
  const int line = src < n ? order[src] : -1;
  perm4[i] = line;
  if (line >= 0) pos4[line] = i;
}
// warp steps of every (batch, sub-segment): the longest of the batch's 8 lines, 4 groups per step
__global__ void k_batch_steps(const int64_t* __restrict__ groups, const int32_t* __restrict__ perm4, int n_batches, int NSUB,
                              uint32_t* __restrict__ steps) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= (int64_t)n_batches * NSUB) return;
  const int64_t b = k / NSUB;
  const int q = (int)(k - b * NSUB);
  int64_t m = 0;
  for (int a = 0; a < 8; ++a) {
    const int line = perm4[b * 8 + a];
    if (line >= 0) {
      const int64_t g = groups[(int64_t)line * NSUB + q];
      m = g > m ? g : m;
    }
  }
  steps[k] = (uint32_t)((m + 3) >> 2);
}

int build_orient(ph_handle* h, PhOrient* o, const int32_t* d_major, const int32_t* d_minor, const uint16_t* d_code,
                 int n_major, int n_minor, const unsigned long long* code_count, uint64_t* kbuf0, uint64_t* kbuf1,
                 void* d_temp, size_t temp_bytes, int* d_flag) {
  const int64_t nnz = h->nnz;
  o->n_major = n_major;
  o->n_minor = n_minor;
  o->chunk_bits = h->chunk_bits;
  const int CH = (1 << o->chunk_bits) - PH_PAD;
  o->NCH = (n_minor + CH - 1) / CH;
  {
    const int last = n_minor - (o->NCH - 1) * CH;
    o->lab_bytes = (((o->NCH - 1) << o->chunk_bits) + PH_PAD + last + 15) & ~15;
  }
  o->minor_bits = ph_ceil_log2((uint64_t)o->lab_bytes);  // a tail word = label-image address | code << minor_bits
  int D = 0;
  // a code gets its own sub-segments when it has >= dense_min_avg entries per line on average AND >= 2 % of all
  // observations (a rarer code costs more in sub-segment bookkeeping than its tail entries do)
  while (D < PH_DMAX && D < h->n_codes && (double)code_count[D] >= h->dense_min_avg * (double)n_major &&
         (double)code_count[D] >= h->dense_min_share * (double)nnz)
    ++D;
  while (D > 0 && D * o->NCH > PH_MAX_SUB) --D;
  o->D = D;
  o->NSUB = D * o->NCH;
  if (h->n_codes > D) {
    if (o->minor_bits >= 32 || (uint64_t)(h->n_codes - 1) >= (1ull << (32 - o->minor_bits)))
      PH_FAIL(PH_ERR_UNSUPPORTED, "%d (var,total) codes do not fit next to a %d-bit label address in a 32-bit word",
              h->n_codes, o->minor_bits);
  }
  const int NSUB = o->NSUB;
  const int64_t nseg = (int64_t)n_major * (NSUB + 1);
  if (nseg >= (1ll << 31)) PH_FAIL(PH_ERR_UNSUPPORTED, "too many segments (%lld)", (long long)nseg);
  o->batch = 0;  // lines per warp batch: automatic (8, or 4 in the quantisation band, see launch_pass); 4 / 8 / 32 via ph_set_group
  o->labels_in_smem = 1;

  const int threads = 256;
  const int blocks = h->sm_count * 8;
  int64_t* d_off = nullptr;      // [nseg + 1] positions in the sorted order
  int64_t* d_sizes = nullptr;    // [n_major * NSUB + 1 | n_major + 1] sizes, then their exclusive sums
  int32_t* d_pos4 = nullptr;     // [n_major] rank of every line in the shape order of the batched-4 copy
  int32_t* d_plan = nullptr;     // [n_major * NSUB][2] padded rounds of every sub-segment (k_round_plan)
  int rc = PH_OK;
  auto body = [&]() -> int {
    const int64_t n_sub = (int64_t)n_major * NSUB;
    PH_CUDA(cudaMalloc(&d_off, (size_t)(nseg + 1) * sizeof(int64_t)));
    PH_CUDA(cudaMalloc(&d_sizes, (size_t)(n_sub + 1 + n_major + 1) * sizeof(int64_t)));
    PH_CUDA(cudaMemset(d_sizes, 0, (size_t)(n_sub + 1 + n_major + 1) * sizeof(int64_t)));
    int64_t* d_groups = d_sizes;
    int64_t* d_tails = d_sizes + n_sub + 1;
    PH_CUDA(cudaMalloc(&o->sub_off, (size_t)(n_sub + 1) * sizeof(uint32_t)));
    PH_CUDA(cudaMalloc(&o->tail_off, (size_t)(n_major + 1) * sizeof(int64_t)));
    h->device_bytes += (n_sub + 1) * (int64_t)sizeof(uint32_t) + (n_major + 1) * (int64_t)sizeof(int64_t);
    const uint64_t* sorted = kbuf0;
    if (nnz > 0) {
      k_make_keys<<<blocks, threads>>>(d_major, d_minor, d_code, nnz, D, o->NCH, CH, o->chunk_bits, o->minor_bits, n_major, n_minor,
                                       h->bank_order, kbuf0, d_flag);
      ph_count_launch();
      cub::DoubleBuffer<uint64_t> db(kbuf0, kbuf1);
      int end_bit = 32 + ph_ceil_log2((uint64_t)nseg + 1);
      if (end_bit > 64) end_bit = 64;
      PH_CUDA(cub::DeviceRadixSort::SortKeys(d_temp, temp_bytes, db, nnz, 0, end_bit));
      ph_count_launch(4);
      sorted = db.Current();
    }
    k_seg_offsets<<<(unsigned)((nseg + 1 + threads - 1) / threads), threads>>>(sorted, nnz, nseg, d_off);
    ph_count_launch();
    if (nnz > 0 && h->bank_order && NSUB > 0 && h->round_fill <= 32) {
      PH_CUDA(cudaMalloc(&d_plan, (size_t)n_sub * 2 * sizeof(int32_t)));
      k_round_plan<<<(unsigned)((n_sub + threads - 1) / threads), threads>>>(sorted, d_off, n_major, NSUB, h->round_fill, d_plan);
      ph_count_launch();
    }
    if (nseg > 0) {
      k_seg_sizes<<<(unsigned)((nseg + threads - 1) / threads), threads>>>(d_off, n_major, NSUB, d_plan, d_groups, d_tails);
      ph_count_launch();
    }
    // ---- batched-4 copy (see PhOrient): shape-sort the lines, fix every line's batch and slot, size the stream
    // (where the per-warp tables of a 1024-thread CTA leave room for the label image -- at a 64 KB-aligned address
    // (LAY = 3) or wherever it fits (LAY = 4) -- or at least for one 64 KB chunk of it: longer minor axes are walked in
    // phases of a few chunks, ph_pass.cu launch_pass)
    const size_t tab4 = 32 * (size_t)(8 * (NSUB + 1) + 8 * D * PH_MAX_FAST_CLUSTERS + 2 * 8 + 3) * 4 + (size_t)h->n_codes * 17 + 256 + 66 * 66;
    const size_t img4 = (size_t)o->lab_bytes < 65536 ? (size_t)o->lab_bytes : 65536;
    const bool want4 = h->batch4 && h->bank_order && nnz > 0 && NSUB > 0 && o->chunk_bits == 16 && n_major >= 8 &&
                       img4 + tab4 <= (size_t)h->max_smem;
    if (want4) {
      const int nb4 = (n_major + 7) / 8;
      uint32_t *d_key = nullptr, *d_key2 = nullptr, *d_steps = nullptr;
      int32_t *d_ord = nullptr, *d_ord2 = nullptr;
      auto sub = [&]() -> int {
        PH_CUDA(cudaMalloc(&d_key, (size_t)n_major * 4));
        PH_CUDA(cudaMalloc(&d_key2, (size_t)n_major * 4));
        PH_CUDA(cudaMalloc(&d_ord, (size_t)n_major * 4));
        PH_CUDA(cudaMalloc(&d_ord2, (size_t)n_major * 4));
        PH_CUDA(cudaMalloc(&o->perm4, (size_t)nb4 * 8 * 4));
        PH_CUDA(cudaMalloc(&d_pos4, (size_t)n_major * 4));
        const unsigned gb = (unsigned)((n_major + threads - 1) / threads);
        k_iota<<<gb, threads>>>(d_ord, n_major);
        // least significant key first; the radix sort is stable, so the last key (the rarest sub-segment) dominates
        for (int q = 0; q < NSUB; ++q) {
          k_shape_key<<<gb, threads>>>(d_groups, d_ord, n_major, NSUB, q, d_key);
          cub::DoubleBuffer<uint32_t> kb(d_key, d_key2);
          cub::DoubleBuffer<int32_t> vb(d_ord, d_ord2);
          size_t need = 0;
          PH_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, kb, vb, n_major, 0, 32));
          if (need > temp_bytes) PH_FAIL(PH_ERR_CUDA, "sort scratch too small");
          PH_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, need, kb, vb, n_major, 0, 32));
          if (vb.Current() != d_ord) {  // keep the current order in d_ord
            int32_t* t = d_ord; d_ord = d_ord2; d_ord2 = t;
          }
          if (kb.Current() != d_key) {
            uint32_t* t = d_key; d_key = d_key2; d_key2 = t;
          }
        }
        k_perm_pos<<<(unsigned)((nb4 * 8 + threads - 1) / threads), threads>>>(d_ord, n_major, nb4 * 8, o->perm4, d_pos4);
        const int64_t nbs = (int64_t)nb4 * NSUB;
        PH_CUDA(cudaMalloc(&d_steps, (size_t)(nbs + 1) * 4));
        PH_CUDA(cudaMemset(d_steps, 0, (size_t)(nbs + 1) * 4));
        k_batch_steps<<<(unsigned)((nbs + threads - 1) / threads), threads>>>(d_groups, o->perm4, nb4, NSUB, d_steps);
        if (h->lpt_order && nb4 > 1) {
          // hand the batches out longest first: sort them by their step count, permute whole batches, recount
          uint32_t *d_bk = nullptr, *d_bk2 = nullptr;
          int32_t *d_bi = nullptr, *d_bi2 = nullptr;
          auto lpt = [&]() -> int {
            PH_CUDA(cudaMalloc(&d_bk, (size_t)nb4 * 4));
            PH_CUDA(cudaMalloc(&d_bk2, (size_t)nb4 * 4));
            PH_CUDA(cudaMalloc(&d_bi, (size_t)nb4 * 4));
            PH_CUDA(cudaMalloc(&d_bi2, (size_t)nb4 * 4));
            k_batch_len_key<<<(unsigned)((nb4 + threads - 1) / threads), threads>>>(d_steps, nb4, NSUB, d_bk, d_bi);
            cub::DoubleBuffer<uint32_t> kb(d_bk, d_bk2);
            cub::DoubleBuffer<int32_t> vb(d_bi, d_bi2);
            size_t need = 0;
            PH_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, kb, vb, nb4, 0, 32));
            if (need > temp_bytes) PH_FAIL(PH_ERR_CUDA, "sort scratch too small");
            PH_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, need, kb, vb, nb4, 0, 32));  // stable: ties keep the shape order
            k_perm_blocks<<<(unsigned)((nb4 * 8 + threads - 1) / threads), threads>>>(d_ord, n_major, vb.Current(), nb4, o->perm4, d_pos4);
            PH_CUDA(cudaMemset(d_steps, 0, (size_t)(nbs + 1) * 4));
            k_batch_steps<<<(unsigned)((nbs + threads - 1) / threads), threads>>>(d_groups, o->perm4, nb4, NSUB, d_steps);
            ph_count_launch(4);
            return PH_OK;
          };
          const int rcl = lpt();
          cudaFree(d_bk);
          cudaFree(d_bk2);
          cudaFree(d_bi);
          cudaFree(d_bi2);
          if (rcl != PH_OK) return rcl;
        }
        PH_CUDA(cudaMalloc(&o->bstep_off, (size_t)(nbs + 1) * 4));
        size_t need = 0;
        PH_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, d_steps, o->bstep_off, (int)(nbs + 1)));
        if (need > temp_bytes) PH_FAIL(PH_ERR_CUDA, "scan scratch too small");
        PH_CUDA(cub::DeviceScan::ExclusiveSum(d_temp, need, d_steps, o->bstep_off, (int)(nbs + 1)));
        ph_count_launch(4 + 2 * NSUB);
        uint32_t total = 0;
        PH_CUDA(cudaMemcpy(&total, o->bstep_off + nbs, 4, cudaMemcpyDeviceToHost));
        o->n_steps4 = total;
        o->n_batches4 = nb4;
        const size_t bytes4 = ((size_t)total + 8) * 32 * PH_GE * sizeof(uint16_t);  // zero = padding entries; slack for the load ring
        PH_CUDA(cudaMalloc(&o->half4, bytes4));
        PH_CUDA(cudaMemset(o->half4, 0, bytes4));
        h->device_bytes += (int64_t)bytes4 + (nbs + 1) * 4 + (int64_t)nb4 * 32;
        const size_t acc_bytes = (size_t)nb4 * 8 * D * PH_MAX_FAST_CLUSTERS * 4;
        PH_CUDA(cudaMalloc(&o->part_acc, acc_bytes));
        PH_CUDA(cudaMemset(o->part_acc, 0, acc_bytes));
        PH_CUDA(cudaMalloc(&o->part_tick, (size_t)nb4 * 4));
        PH_CUDA(cudaMemset(o->part_tick, 0, (size_t)nb4 * 4));
        h->device_bytes += (int64_t)acc_bytes + (int64_t)nb4 * 4;
        return PH_OK;
      };
      const int rc4 = sub();
      cudaFree(d_key);
      cudaFree(d_key2);
      cudaFree(d_ord);
      cudaFree(d_ord2);
      cudaFree(d_steps);
      if (rc4 != PH_OK) return rc4;
    }
    if (nnz > 0 && h->bank_order && NSUB > 0) {
      // second sort: round order inside every sub-segment (segment sizes, hence d_off, do not change)
      uint64_t* other = (sorted == kbuf0) ? kbuf1 : kbuf0;
      k_round_keys<<<blocks, threads>>>(sorted, nnz, d_off, NSUB, o->half4 ? d_pos4 : nullptr, other, d_flag + 1);
      ph_count_launch();
      cub::DoubleBuffer<uint64_t> db(other, const_cast<uint64_t*>(sorted));
      int end_bit = 32 + ph_ceil_log2((uint64_t)nseg + 1);
      if (end_bit > 64) end_bit = 64;
      PH_CUDA(cub::DeviceRadixSort::SortKeys(d_temp, temp_bytes, db, nnz, 0, end_bit));
      ph_count_launch(4);
      sorted = db.Current();
    }
    // exclusive sums in place (the extra trailing zero element receives the total)
    size_t need = 0, need2 = 0;
    PH_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, d_groups, d_groups, (int)(n_sub + 1)));
    PH_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need2, d_tails, d_tails, (int)(n_major + 1)));
    if (need > temp_bytes || need2 > temp_bytes) PH_FAIL(PH_ERR_CUDA, "scan scratch too small");
    PH_CUDA(cub::DeviceScan::ExclusiveSum(d_temp, need, d_groups, d_groups, (int)(n_sub + 1)));
    PH_CUDA(cub::DeviceScan::ExclusiveSum(d_temp, need2, d_tails, d_tails, (int)(n_major + 1)));
    ph_count_launch(2);
    int64_t totals[2];
    PH_CUDA(cudaMemcpy(&totals[0], d_groups + n_sub, sizeof(int64_t), cudaMemcpyDeviceToHost));
    PH_CUDA(cudaMemcpy(&totals[1], d_tails + n_major, sizeof(int64_t), cudaMemcpyDeviceToHost));
    o->n_groups = totals[0];
    o->n_tail = totals[1];
    if (o->n_groups >= (1ll << 32) - 4096) PH_FAIL(PH_ERR_UNSUPPORTED, "%lld groups exceed 32-bit group offsets", (long long)o->n_groups);
    k_narrow<<<(unsigned)((n_sub + 1 + threads - 1) / threads), threads>>>(d_groups, n_sub + 1, o->sub_off);
    ph_count_launch();
    PH_CUDA(cudaMemcpy(o->tail_off, d_tails, (size_t)(n_major + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice));
    const size_t half_bytes = (size_t)(o->n_groups + 160) * PH_GE * sizeof(uint16_t);  // with some zero slack
    PH_CUDA(cudaMalloc(&o->half, half_bytes));
    PH_CUDA(cudaMemset(o->half, 0, half_bytes));
    PH_CUDA(cudaMalloc(&o->tail_words, (size_t)(o->n_tail + 8) * sizeof(uint32_t)));
    PH_CUDA(cudaMemset(o->tail_words, 0, (size_t)(o->n_tail + 8) * sizeof(uint32_t)));
    h->device_bytes += (int64_t)half_bytes + (o->n_tail + 8) * (int64_t)sizeof(uint32_t);
    if (nnz > 0) {
      if (d_plan) {
        k_fill_rounds<<<(unsigned)((n_sub * 32 + threads - 1) / threads), threads>>>(d_off, n_major, NSUB, d_plan, o->sub_off, h->bank_order,
                                                                                    o->half, o->half4 ? d_pos4 : nullptr, o->bstep_off, o->half4);
        ph_count_launch();
      }
      k_scatter<<<blocks, threads>>>(sorted, nnz, d_off, NSUB, d_plan, o->sub_off, o->tail_off, h->bank_order, o->half,
                                     o->tail_words, o->half4 ? d_pos4 : nullptr, o->bstep_off, o->half4, d_flag + 1);
      ph_count_launch();
    }
    PH_CUDA(cudaGetLastError());
    PH_CUDA(cudaDeviceSynchronize());
    return PH_OK;
  };
  rc = body();
  cudaFree(d_off);
  cudaFree(d_sizes);
  cudaFree(d_pos4);
  cudaFree(d_plan);
  return rc;
}

}  // namespace

extern "C" void ph_destroy(ph_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (PhOrient* o : {&h->by_snv, &h->by_cell}) {
    cudaFree(o->half);
    cudaFree(o->sub_off);
    cudaFree(o->tail_words);
    cudaFree(o->tail_off);
    cudaFree(o->half4);
    cudaFree(o->bstep_off);
    cudaFree(o->perm4);
    cudaFree(o->part_acc);
    cudaFree(o->part_tick);
  }
  cudaFree(h->d_lab_img);
  cudaFree(h->d_emb);
  cudaFree(h->d_gauss);
  cudaFree(h->d_L0);
  cudaFree(h->d_L1);
  cudaFree(h->d_varpos);
  cudaFree(h->d_lab_minor);
  cudaFree(h->d_lab_major);
  cudaFree(h->d_list);
  cudaFree(h->d_f64);
  cudaFree(h->d_i32);
  cudaFree(h->d_u8);
  cudaFree(h->d_small);
  cudaFree(h->d_map);
  cudaFree(h->d_counter);
  cudaFree(h->d_partial);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" int ph_create(int32_t n_cells, int32_t n_snvs, int64_t nnz, const int32_t* cell, const int32_t* snv,
                         const uint16_t* code, int coo_on_device, int32_t n_codes, const double* L0, const double* L1,
                         const uint8_t* var_pos, int device, ph_handle** out) {
  if (!out) PH_FAIL(PH_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (n_cells <= 0 || n_snvs <= 0 || nnz < 0 || n_codes <= 0 || n_codes > 65535)
    PH_FAIL(PH_ERR_ARG, "bad sizes n_cells=%d n_snvs=%d nnz=%lld n_codes=%d", n_cells, n_snvs, (long long)nnz, n_codes);
  if (nnz > 0 && (!cell || !snv || !code)) PH_FAIL(PH_ERR_ARG, "COO pointer is NULL");
  if (!L0 || !L1 || !var_pos) PH_FAIL(PH_ERR_ARG, "likelihood table pointer is NULL");
  int ndev = 0;
  PH_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) PH_FAIL(PH_ERR_CUDA, "device %d not available (%d CUDA devices)", device, ndev);
  PH_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PH_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) PH_FAIL(PH_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);

  ph_handle* h = new (std::nothrow) ph_handle();
  if (!h) PH_FAIL(PH_ERR_ARG, "out of host memory");
  memset(h, 0, sizeof(*h));
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->max_smem = (int)prop.sharedMemPerBlockOptin;
  h->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
  h->n_cells = n_cells;
  h->n_snvs = n_snvs;
  h->n_codes = n_codes;
  h->nnz = nnz;
  {
    const char* e = getenv("PH_DENSE_MIN_AVG");  // tuning knob: minimum average entries per line of a dense code
    h->dense_min_avg = e ? atof(e) : 8.0;
    e = getenv("PH_DENSE_MIN_SHARE");
    h->dense_min_share = e ? atof(e) : 0.02;
    e = getenv("PH_CHUNK_BITS");  // test knob: log2 of the label-image stride of one chunk of the minor axis
    h->chunk_bits = e ? atoi(e) : 16;
    if (h->chunk_bits < 8 || h->chunk_bits > 16) h->chunk_bits = 16;
    e = getenv("PH_BANK_ORDER");  // order a sub-segment's entries by shared-memory bank (default on)
    h->bank_order = e ? atoi(e) : 1;
    e = getenv("PH_CTAS");  // pass kernels: CTAs per SM (1 x 1024 threads, or 2 x 512 when two label images fit)
    h->ctas_per_sm = e ? atoi(e) : 1;
    e = getenv("PH_ALIGNED_IMAGE");  // pass kernels: label image at a 2^16-aligned shared address when it fits (default on)
    h->aligned_image = e ? atoi(e) : 1;
    e = getenv("PH_ZERO_COPY_OUT");  // host-pointer calls: kernels store results straight into page-locked host memory (default on)
    h->zero_copy_out = e ? atoi(e) : 1;
    // padded rounds (k_round_plan): rounds in which at least this share of the 32 banks hold an entry sit on a fixed
    // grid with bank-matched padding entries; > 1 disables the padding (round order only)
    e = getenv("PH_ROUND_FILL");
    {
      const double f = e ? atof(e) : 0.75;
      h->round_fill = f > 1.0 ? 33 : (int)(f * 32.0 + 0.999);
      if (h->round_fill < 1) h->round_fill = 1;
    }
    e = getenv("PH_PHASE_CHUNKS");  // test knob: walk label images of more chunks than this in phases (ph_pass.cu launch_pass)
    h->phase_chunks = e ? atoi(e) : 0;
    e = getenv("PH_LPT_ORDER");  // batched-4 copy: batches laid out (= handed out) longest first (default on)
    h->lpt_order = e ? atoi(e) : 1;
    // split tail of the batched-4 pass (ph_pass.cu PassArgs): parts per batch = 2^PH_SPLIT_SH (0 = off), over the last
    // PH_SPLIT_TAIL percent of (warps of the grid) batches
    e = getenv("PH_SPLIT_SH");
    h->split_sh = e ? atoi(e) : 0;  // off: measured slower wherever it was tried -- the fence + ticket
                                     // round trips of a part cost more than the shorter end of the pass returns
    if (h->split_sh < 0 || h->split_sh > 4) h->split_sh = 0;
    e = getenv("PH_SPLIT_TAIL");
    h->split_tail = e ? atoi(e) : 150;
    e = getenv("PH_ITEM_BATCHES");  // batched-4 pass: 8-line batches per item -- 0 = automatic, 1 / 2 / 4 = forced
    h->item_batches = e ? atoi(e) : 0;
    e = getenv("PH_BATCH4");  // keep the batched-4 copy of the dense streams and use the 4-lanes-per-line pass (default on)
    h->batch4 = e ? atoi(e) : 1;
  }

  int rc = PH_OK;
  int32_t *d_cell = nullptr, *d_snv = nullptr;
  uint16_t* d_code = nullptr;
  uint64_t *kbuf0 = nullptr, *kbuf1 = nullptr;
  void* d_temp = nullptr;
  unsigned long long* d_hist = nullptr;
  int* d_flag = nullptr;
  unsigned long long* h_hist = nullptr;
  size_t temp_bytes = 0;
  const int32_t nmax = n_cells > n_snvs ? n_cells : n_snvs;

#define PH_TRY(expr)                 \
  do {                               \
    rc = [&]() -> int {              \
      PH_CUDA(expr);                 \
      return PH_OK;                  \
    }();                             \
    if (rc != PH_OK) goto done;      \
  } while (0)

  PH_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  PH_TRY(cudaMalloc(&h->d_L0, n_codes * sizeof(double)));
  PH_TRY(cudaMalloc(&h->d_L1, n_codes * sizeof(double)));
  PH_TRY(cudaMalloc(&h->d_varpos, n_codes));
  PH_TRY(cudaMemcpy(h->d_L0, L0, n_codes * sizeof(double), cudaMemcpyHostToDevice));
  PH_TRY(cudaMemcpy(h->d_L1, L1, n_codes * sizeof(double), cudaMemcpyHostToDevice));
  PH_TRY(cudaMemcpy(h->d_varpos, var_pos, n_codes, cudaMemcpyHostToDevice));

  if (coo_on_device) {
    d_cell = const_cast<int32_t*>(cell);
    d_snv = const_cast<int32_t*>(snv);
    d_code = const_cast<uint16_t*>(code);
  } else if (nnz > 0) {
    PH_TRY(cudaMalloc(&d_cell, nnz * sizeof(int32_t)));
    PH_TRY(cudaMalloc(&d_snv, nnz * sizeof(int32_t)));
    PH_TRY(cudaMalloc(&d_code, nnz * sizeof(uint16_t)));
    PH_TRY(cudaMemcpy(d_cell, cell, nnz * sizeof(int32_t), cudaMemcpyHostToDevice));
    PH_TRY(cudaMemcpy(d_snv, snv, nnz * sizeof(int32_t), cudaMemcpyHostToDevice));
    PH_TRY(cudaMemcpy(d_code, code, nnz * sizeof(uint16_t), cudaMemcpyHostToDevice));
  }
  PH_TRY(cudaMalloc(&d_hist, n_codes * sizeof(unsigned long long)));
  PH_TRY(cudaMemset(d_hist, 0, n_codes * sizeof(unsigned long long)));
  PH_TRY(cudaMalloc(&d_flag, 4 * sizeof(int)));
  PH_TRY(cudaMemset(d_flag, 0, 4 * sizeof(int)));
  h_hist = new unsigned long long[n_codes];
  if (nnz > 0) {
    k_code_hist<<<h->sm_count * 8, 256>>>(d_code, nnz, d_hist, n_codes, d_flag);
    ph_count_launch();
  }
  PH_TRY(cudaMemcpy(h_hist, d_hist, n_codes * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  {
    int flags[4];
    PH_TRY(cudaMemcpy(flags, d_flag, sizeof(flags), cudaMemcpyDeviceToHost));
    if (flags[0]) {
      snprintf(ph_err_msg, sizeof(ph_err_msg), "code out of range (>= n_codes=%d)", n_codes);
      rc = PH_ERR_ARG;
      goto done;
    }
  }
  if (nnz > 0) {
    PH_TRY(cudaMalloc(&kbuf0, nnz * sizeof(uint64_t)));
    PH_TRY(cudaMalloc(&kbuf1, nnz * sizeof(uint64_t)));
    {
      cub::DoubleBuffer<uint64_t> db(kbuf0, kbuf1);
      PH_TRY(cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, db, nnz, 0, 64));
    }
  }
  if (temp_bytes < (1u << 20)) temp_bytes = 1u << 20;  // also serves the offset scans
  PH_TRY(cudaMalloc(&d_temp, temp_bytes));
  if (nnz > 1) {
    k_pair_keys<<<h->sm_count * 8, 256>>>(d_snv, d_cell, nnz, kbuf0);
    cub::DoubleBuffer<uint64_t> db(kbuf0, kbuf1);
    int end_bit = 32 + ph_ceil_log2((uint64_t)n_snvs + 1);
    if (end_bit > 64) end_bit = 64;
    PH_TRY(cub::DeviceRadixSort::SortKeys(d_temp, temp_bytes, db, nnz, 0, end_bit));
    k_adjacent_equal<<<h->sm_count * 8, 256>>>(db.Current(), nnz, d_flag + 1);
    ph_count_launch(6);
  }
  rc = build_orient(h, &h->by_snv, d_snv, d_cell, d_code, n_snvs, n_cells, h_hist, kbuf0, kbuf1, d_temp, temp_bytes, d_flag);
  if (rc != PH_OK) goto done;
  rc = build_orient(h, &h->by_cell, d_cell, d_snv, d_code, n_cells, n_snvs, h_hist, kbuf0, kbuf1, d_temp, temp_bytes, d_flag);
  if (rc != PH_OK) goto done;
  PH_TRY(cudaDeviceSynchronize());
  {
    int flags[4];
    PH_TRY(cudaMemcpy(flags, d_flag, sizeof(flags), cudaMemcpyDeviceToHost));
    if (flags[0] == 2) {
      snprintf(ph_err_msg, sizeof(ph_err_msg), "COO index out of range");
      rc = PH_ERR_ARG;
      goto done;
    }
    if (flags[1]) {
      snprintf(ph_err_msg, sizeof(ph_err_msg), "duplicate (cell, snv) observation in COO input");
      rc = PH_ERR_ARG;
      goto done;
    }
  }

  // scratch
  PH_TRY(cudaMalloc(&h->d_lab_minor, (size_t)nmax + 16));
  PH_TRY(cudaMalloc(&h->d_lab_major, (size_t)nmax + 16));
  h->lab_img_bytes = (size_t)(h->by_snv.lab_bytes > h->by_cell.lab_bytes ? h->by_snv.lab_bytes : h->by_cell.lab_bytes) + 64;
  PH_TRY(cudaMalloc(&h->d_lab_img, h->lab_img_bytes));
  PH_TRY(cudaMalloc(&h->d_list, (size_t)nmax * sizeof(int32_t)));
  PH_TRY(cudaMalloc(&h->d_f64, (size_t)nmax * 2 * PH_MAX_FAST_CLUSTERS * sizeof(double)));
  PH_TRY(cudaMalloc(&h->d_i32, (size_t)nmax * 2 * PH_MAX_FAST_CLUSTERS * sizeof(int32_t)));
  PH_TRY(cudaMalloc(&h->d_u8, (size_t)nmax + 16));
  PH_TRY(cudaMalloc(&h->d_small, 4096 * sizeof(double)));
  PH_TRY(cudaMalloc(&h->d_map, 65 * 65 + 64));
  PH_TRY(cudaMalloc(&h->d_counter, 64));
  PH_TRY(cudaMemset(h->d_counter, 0, 64));  // [0] batch dispenser, [1] CTA completion counter: both re-arm themselves
  h->device_bytes += (int64_t)nmax * (3 + 4 + 2 * PH_MAX_FAST_CLUSTERS * 12) + 4096 * 8;
  h->h_stage_bytes = (size_t)nmax * (2 * PH_MAX_FAST_CLUSTERS * 12 + 8) + (1 << 16);
  PH_TRY(cudaMallocHost(&h->h_stage, h->h_stage_bytes));

done:
  if (!coo_on_device) {
    cudaFree(d_cell);
    cudaFree(d_snv);
    cudaFree(d_code);
  }
  cudaFree(kbuf0);
  cudaFree(kbuf1);
  cudaFree(d_temp);
  cudaFree(d_hist);
  cudaFree(d_flag);
  delete[] h_hist;
  if (rc != PH_OK) {
    ph_destroy(h);
    return rc;
  }
  cudaDeviceSynchronize();  // scratch initialisation (legacy stream) is complete before any caller stream uses it
  *out = h;
  return PH_OK;
#undef PH_TRY
}

extern "C" int ph_get_info(const ph_handle* h, ph_info* out) {
  if (!h || !out) PH_FAIL(PH_ERR_ARG, "null argument");
  out->n_cells = h->n_cells;
  out->n_snvs = h->n_snvs;
  out->n_codes = h->n_codes;
  out->nnz = h->nnz;
  out->dense_codes_by_snv = h->by_snv.D;
  out->dense_codes_by_cell = h->by_cell.D;
  out->device_bytes = h->device_bytes;
  out->device = h->device;
  out->sm_count = h->sm_count;
  out->group_by_snv = h->by_snv.batch;
  out->group_by_cell = h->by_cell.batch;
  out->chunks_by_snv = h->by_snv.NCH;
  out->chunks_by_cell = h->by_cell.NCH;
  out->stream_bytes_by_snv = h->by_snv.n_groups * PH_GE * 2 + h->by_snv.n_tail * 4;
  out->stream_bytes_by_cell = h->by_cell.n_groups * PH_GE * 2 + h->by_cell.n_tail * 4;
  out->stream4_bytes_by_snv = h->by_snv.half4 ? h->by_snv.n_steps4 * 32 * PH_GE * 2 + h->by_snv.n_tail * 4 : 0;
  out->stream4_bytes_by_cell = h->by_cell.half4 ? h->by_cell.n_steps4 * 32 * PH_GE * 2 + h->by_cell.n_tail * 4 : 0;
  out->labels_in_smem_by_snv = h->by_snv.labels_in_smem;
  out->labels_in_smem_by_cell = h->by_cell.labels_in_smem;
  return PH_OK;
}

extern "C" int ph_set_group(ph_handle* h, int group_by_snv, int group_by_cell) {
  if (!h) PH_FAIL(PH_ERR_ARG, "null handle");
  auto ok = [](int g) { return g == 0 || g == -1 || g == 4 || g == 8 || g == 16 || g == 32; };
  if (!ok(group_by_snv) || !ok(group_by_cell)) PH_FAIL(PH_ERR_ARG, "group must be -1, 0, 4, 8, 16 or 32");
  // lines per warp batch: 4, 8 or 32 (16 is accepted for ABI compatibility and means 8); -1 = automatic (4 when the
  // launch has fewer than two batches of 8 per warp, else 8); 0 = leave unchanged
  auto val = [](int g) { return g == -1 ? 0 : (g == 32 ? 32 : (g == 4 ? 4 : 8)); };
  if (group_by_snv) h->by_snv.batch = val(group_by_snv);
  if (group_by_cell) h->by_cell.batch = val(group_by_cell);
  return PH_OK;
}
