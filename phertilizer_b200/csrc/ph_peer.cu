// Cell-sharded K1 + K1a over NVLink peer memory -- no NCCL on the data path.
//
// Every rank owns a contiguous block of cells (its own ph_handle) and a contiguous block of the listed SNVs.  One step:
//
//   A  k_pass<MODE_COUNTS> (ph_pass.cu): the rank's pass over its observations; the (cluster, code) histogram row of
//      every SNV is stored straight into the window of the rank that OWNS the SNV (16-byte peer stores from the kernel
//      epilogue, overlapping the streaming of the following batches); the last CTA raises this rank's flag everywhere.
//   B  k_peer_finish: the owner waits for the world's flags, sums the `world` integer rows of each of its SNVs,
//      evaluates  sum_code fma(count, L[code])  exactly like the single-GPU holder does, applies the K1a decision rule
//      and stores label + observation counts into EVERY rank's output window; then raises its second flag everywhere.
//   C  k_peer_collect: waits for the owners' flags and copies the window to the caller's output buffers.
// (The SNV-sharded steps further down need none of this: one kernel, self-validating words, no flags.)
//
// The exchanged quantity is an integer histogram, so the result is bit-identical to the single-GPU kernels for any
// number of ranks.  Windows are plain cudaMalloc allocations shared through CUDA IPC handles (exchanged by the host
// language -- torch.distributed in this repository -- as opaque 64-byte blobs).
#include <stdlib.h>

#include "ph_common.cuh"

struct ph_peer {
  ph_handle* h;
  int rank, world;
  int32_t max_lines, blk_max, ncp;
  size_t off_cnt, off_lab, off_nobs, off_ll, bytes;  // off_ll: the LL words of the SNV-sharded steps (ph_common.cuh, PhLL)
  int32_t cap_slots;           // result slots (lines of the work order, padded to whole batches) one rank may push per step
  uint8_t* win[PH_MAX_PEERS];  // win[rank] = own allocation, others = IPC mappings
  bool opened[PH_MAX_PEERS];
  long long spin_limit;        // SM clocks a wait of the SNV-sharded collect may take (PH_PEER_TIMEOUT_MS, default ~3 s)
  unsigned* d_done;            // [0],[1] CTA completion counters (pass, finish), [2] spin-timeout flag,
                               // [3] collect completion counter, [4] step counter ("epoch"), advanced by the last
                               // block of every collect kernel: no launch parameter changes from step to step, so a
                               // step can be captured in a CUDA graph and replayed
};

namespace {

constexpr size_t kFlagBytes = 4096;  // flagA[world] at +0, flagB[world] at +1024, LL headers hdr[2][PH_MAX_PEERS] (8 B each) at +2048
constexpr size_t kHdrOff = 2048;
struct FinishArgs {
  const uint8_t* win;  // own window
  size_t off_cnt;
  int world, rank, blk, blk_max, ncp, n_codes, n_list, C, mode;  // mode 1 = linear, 2 = branching
  double lenA;
  const double* L0;
  const double* L1;
  uint8_t* lab_dst[PH_MAX_PEERS];
  int32_t* nobs_dst[PH_MAX_PEERS];
  uint32_t* flag_dst[PH_MAX_PEERS];
  unsigned* done;
  unsigned* timeout_flag;
  const uint32_t* epoch_ptr;
};

template <int C>
__global__ void __launch_bounds__(256) k_peer_finish(const FinishArgs a) {
  const uint32_t epoch = *a.epoch_ptr + 1u;
  wait_flags(reinterpret_cast<const uint32_t*>(a.win), a.world, epoch, a.timeout_flag);
  const int first = a.rank * a.blk;
  const int n_own = min(a.blk, a.n_list - first);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_own; idx += gridDim.x * blockDim.x) {
    double s0[C], s1[C];
    unsigned nob[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { s0[c] = 0.0; s1[c] = 0.0; nob[c] = 0; }
    const size_t row = (size_t)C * a.ncp;  // uint16 per line
    const size_t slot = (size_t)a.blk_max * 3 * a.ncp;  // uint16 per source rank (sized for C = 3)
    const uint16_t* base = reinterpret_cast<const uint16_t*>(a.win + a.off_cnt) + (size_t)idx * row;
    for (int c = 0; c < C; ++c) {
      for (int k8 = 0; k8 < a.ncp; k8 += 8) {
        unsigned t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int s = 0; s < a.world; ++s) {
          // written by a peer since the last launch: bypass L1
          const uint4 v = __ldcv(reinterpret_cast<const uint4*>(base + s * slot + c * a.ncp + k8));
          t[0] += v.x & 0xFFFFu; t[1] += v.x >> 16; t[2] += v.y & 0xFFFFu; t[3] += v.y >> 16;
          t[4] += v.z & 0xFFFFu; t[5] += v.z >> 16; t[6] += v.w & 0xFFFFu; t[7] += v.w >> 16;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int code = k8 + j;
          if (code < a.n_codes) {  // ascending code order, fma(count, L, sum): the holder's arithmetic (ph_pass.cu)
            s0[c] = fma((double)t[j], a.L0[code], s0[c]);
            s1[c] = fma((double)t[j], a.L1[code], s1[c]);
            nob[c] += t[j];
          }
        }
      }
    }
    const int k = first + idx;
    uint8_t label;
    if (a.mode == 1) {  // linear_split.py:229-234
      const double m0 = s0[0] / a.lenA, m1 = s1[0] / a.lenA;
      label = (m0 >= m1) ? 2 : 1;
    } else {  // branching_split.py:332-346
      const double cA = s1[0] + s0[C > 1 ? 1 : 0];
      const double cB = s0[0] + s1[C > 1 ? 1 : 0];
      const double cC = s1[0] + s1[C > 1 ? 1 : 0];
      const double mx = fmax(cA, fmax(cB, cC));
      label = (cA == mx) ? 1 : ((cB == mx) ? 2 : 3);
    }
    for (int p = 0; p < a.world; ++p) {
      a.lab_dst[p][k] = label;
#pragma unroll
      for (int c = 0; c < C; ++c) a.nobs_dst[p][(size_t)k * C + c] = (int32_t)nob[c];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned prev = atomicAdd(a.done, 1u);
    if (prev == gridDim.x - 1) {
      *a.done = 0;
      __threadfence_system();
      for (int p = 0; p < a.world; ++p)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.flag_dst[p]), "r"(epoch) : "memory");
    }
  }
}

// cell-sharded steps: waits for the owners' flags, copies the output window to the caller, and the last block to finish
// advances the step counter
__global__ void __launch_bounds__(256) k_peer_collect(const uint8_t* win, size_t off_flags, size_t off_lab, size_t off_nobs,
                                                     int world, uint32_t* epoch_ptr, unsigned* done, int n_list, int C,
                                                     uint8_t* label_out, int32_t* nobs_out, unsigned* timeout_flag) {
  const uint32_t epoch = *epoch_ptr + 1u;
  wait_flags(reinterpret_cast<const uint32_t*>(win + off_flags), world, epoch, timeout_flag);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  const uint32_t* src_n = reinterpret_cast<const uint32_t*>(win + off_nobs);
  for (int i = tid; i < n_list * C; i += nt) nobs_out[i] = (int32_t)__ldcv(src_n + i);
  const uint8_t* src_l = win + off_lab;
  for (int i = tid; i < n_list; i += nt) label_out[i] = __ldcv(src_l + i);
  __syncthreads();  // every thread of the block has read the counter
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) {
      *done = 0;
      *epoch_ptr = epoch;
    }
  }
}

}  // namespace

// implemented in ph_pass.cu: the COUNTS pass with its peer destinations
int ph_pass_counts(ph_handle* h, const uint8_t* cell_label, int C, const int32_t* snv_list, int32_t n_list,
                   uint16_t* const* cnt_dst, uint32_t* const* flag_dst, unsigned* done, int blk, int ncp, int world,
                   const uint32_t* epoch_ptr, cudaStream_t st);

int ph_pass_assign_push(ph_handle* h, int mode, const uint8_t* cell_label, double lenA, const PhLL& ll, unsigned* done,
                        int out_off, int world, uint32_t* epoch_ptr, cudaStream_t st);

extern "C" int ph_peer_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" void ph_peer_destroy(ph_peer* p) {
  if (!p) return;
  cudaSetDevice(p->h->device);
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) continue;
    if (p->opened[r]) cudaIpcCloseMemHandle(p->win[r]);
  }
  cudaFree(p->win[p->rank]);
  cudaFree(p->d_done);
  delete p;
}

extern "C" int ph_peer_create(ph_handle* h, int rank, int world, int32_t total_snvs, ph_peer** out) {
  if (!h || !out) PH_FAIL(PH_ERR_ARG, "null argument");
  *out = nullptr;
  // world = 1 is a degenerate but valid context (the window is local): the same kernels, no NVLink traffic
  if (world < 1 || world > PH_MAX_PEERS || rank < 0 || rank >= world)
    PH_FAIL(PH_ERR_ARG, "rank %d / world %d outside 1..%d", rank, world, PH_MAX_PEERS);
  if (total_snvs <= 0) {
    // cell-sharded (every rank holds all SNVs): the limits of the histogram exchange are checked HERE, so that set-up
    // fails on every rank alike (the host language gathers the outcomes) instead of one rank refusing a later step
    // while its peers have already launched and wait for it
    if (h->n_cells > 65535)
      PH_FAIL(PH_ERR_UNSUPPORTED, "%d local cells: the exchanged histogram rows are 16-bit (at most 65535 cells per rank)", h->n_cells);
    if (h->n_codes > 64) PH_FAIL(PH_ERR_UNSUPPORTED, "%d (var,total) codes: histogram exchange supports up to 64", h->n_codes);
    total_snvs = h->n_snvs;
  }
  if (total_snvs < h->n_snvs) PH_FAIL(PH_ERR_ARG, "total_snvs=%d is less than the %d local SNVs", total_snvs, h->n_snvs);
  PH_CUDA(cudaSetDevice(h->device));
  ph_peer* p = new ph_peer();
  memset(p, 0, sizeof(*p));
  p->h = h;
  p->rank = rank;
  p->world = world;
  p->max_lines = total_snvs;
  // (every offset below must be the same on all ranks: a rank computes where ITS words go in a PEER's window from its own
  // copy of this layout.  SNV-sharded ranks hold blocks of different sizes -- only global quantities may enter.)
  p->blk_max = (total_snvs + world - 1) / world;
  p->spin_limit = kSpinLimit;
  if (const char* e = getenv("PH_PEER_TIMEOUT_MS")) {
    const long long ms = atoll(e);
    if (ms > 0) p->spin_limit = ms * 2000000ll;  // ~2 GHz SM clock
  }
  p->ncp = (h->n_codes + 7) & ~7;
  if (p->ncp > 64) p->ncp = 64;  // the histogram exchange (cell-sharded) is refused above 64 codes, see below
  p->off_cnt = kFlagBytes;
  const size_t cnt_bytes = (size_t)world * p->blk_max * 3 * p->ncp * sizeof(uint16_t);
  p->off_lab = (p->off_cnt + cnt_bytes + 255) & ~(size_t)255;
  p->off_nobs = (p->off_lab + (size_t)p->max_lines + 255) & ~(size_t)255;
  // SNV-sharded steps: every rank pushes the LL words of its block (work order, whole batches of 8) into its own region
  // of each window; blocks are expected to be balanced (at most ceil(total / world) + 8 SNVs per rank)
  p->cap_slots = (((total_snvs + world - 1) / world + 8 + 7) / 8) * 8 + 40;  // lines in whole 8-line batches + one 32-line item of slack
  p->off_ll = (p->off_nobs + (size_t)p->max_lines * 3 * sizeof(int32_t) + 255) & ~(size_t)255;
  p->bytes = p->off_ll + (size_t)2 * kLLWords * world * p->cap_slots * 8 + 256;
  cudaError_t e = cudaMalloc(&p->win[rank], p->bytes);
  if (e == cudaSuccess) e = cudaMemset(p->win[rank], 0, p->bytes);
  if (e == cudaSuccess) e = cudaMalloc(&p->d_done, 256);
  if (e == cudaSuccess) e = cudaMemset(p->d_done, 0, 256);
  // the window must BE zero before its handle leaves this call: a peer stores into it as soon as the host-side set-up
  // barrier lets it go, and a memset still queued on this device (another process holding the GPU's time slice, a busy
  // stream) would wipe those words afterwards -- every rank would then wait for data that had already arrived
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    snprintf(ph_err_msg, sizeof(ph_err_msg), "peer window allocation failed: %s", cudaGetErrorString(e));
    ph_peer_destroy(p);
    return PH_ERR_CUDA;
  }
  *out = p;
  return PH_OK;
}

extern "C" int ph_peer_export(ph_peer* p, void* handle_out) {
  if (!p || !handle_out) PH_FAIL(PH_ERR_ARG, "null argument");
  PH_CUDA(cudaSetDevice(p->h->device));
  cudaIpcMemHandle_t hd;
  PH_CUDA(cudaIpcGetMemHandle(&hd, p->win[p->rank]));
  memcpy(handle_out, &hd, sizeof(hd));
  return PH_OK;
}

extern "C" int ph_peer_connect(ph_peer* p, const void* all_handles) {
  if (!p || !all_handles) PH_FAIL(PH_ERR_ARG, "null argument");
  PH_CUDA(cudaSetDevice(p->h->device));
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank || p->opened[r]) continue;
    cudaIpcMemHandle_t hd;
    memcpy(&hd, (const uint8_t*)all_handles + (size_t)r * sizeof(hd), sizeof(hd));
    void* ptr = nullptr;
    PH_CUDA(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    p->win[r] = (uint8_t*)ptr;
    p->opened[r] = true;
  }
  return PH_OK;
}

namespace {
int peer_assign(ph_peer* p, int mode, const uint8_t* cell_label, double lenA, const int32_t* snv_list, int32_t n_list,
                uint8_t* label_out, int32_t* nobs_out, void* stream) {
  if (!p) PH_FAIL(PH_ERR_ARG, "null peer context");
  if (!cell_label || !label_out || !nobs_out) PH_FAIL(PH_ERR_ARG, "null pointer");
  ph_handle* h = p->h;
  if (!snv_list) n_list = h->n_snvs;
  if (n_list <= 0 || n_list > h->n_snvs) PH_FAIL(PH_ERR_ARG, "n_list=%d outside 1..%d", n_list, h->n_snvs);
  if (h->n_cells > 65535)
    PH_FAIL(PH_ERR_UNSUPPORTED, "%d local cells: the exchanged histogram rows are 16-bit (at most 65535 cells per rank)", h->n_cells);
  if (h->n_codes > 64) PH_FAIL(PH_ERR_UNSUPPORTED, "%d (var,total) codes: histogram exchange supports up to 64", h->n_codes);
  for (int r = 0; r < p->world; ++r)
    if (!p->win[r]) PH_FAIL(PH_ERR_ARG, "ph_peer_connect has not been called (rank %d missing)", r);
  PH_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int C = mode == 1 ? 1 : 2;
  const int blk = (n_list + p->world - 1) / p->world;
  const uint32_t* epoch_ptr = reinterpret_cast<const uint32_t*>(p->d_done + 4);
  const size_t slot = (size_t)p->blk_max * 3 * p->ncp;  // uint16 per source rank

  uint16_t* cnt_dst[PH_MAX_PEERS] = {};
  uint32_t* flagA_dst[PH_MAX_PEERS] = {};
  FinishArgs f;
  memset(&f, 0, sizeof(f));
  for (int r = 0; r < p->world; ++r) {
    cnt_dst[r] = reinterpret_cast<uint16_t*>(p->win[r] + p->off_cnt) + (size_t)p->rank * slot;
    flagA_dst[r] = reinterpret_cast<uint32_t*>(p->win[r]) + p->rank;
    f.lab_dst[r] = p->win[r] + p->off_lab;
    f.nobs_dst[r] = reinterpret_cast<int32_t*>(p->win[r] + p->off_nobs);
    f.flag_dst[r] = reinterpret_cast<uint32_t*>(p->win[r] + 1024) + p->rank;
  }
  int rc = ph_pass_counts(h, cell_label, C, snv_list, n_list, cnt_dst, flagA_dst, p->d_done, blk, p->ncp, p->world, epoch_ptr, st);
  if (rc != PH_OK) return rc;

  f.win = p->win[p->rank];
  f.off_cnt = p->off_cnt;
  f.world = p->world;
  f.rank = p->rank;
  f.blk = blk;
  f.blk_max = p->blk_max;
  f.ncp = p->ncp;
  f.n_codes = h->n_codes;
  f.n_list = n_list;
  f.C = C;
  f.mode = mode;
  f.lenA = lenA;
  f.L0 = h->d_L0;
  f.L1 = h->d_L1;
  f.done = p->d_done + 1;
  f.timeout_flag = p->d_done + 2;
  f.epoch_ptr = epoch_ptr;
  const int fb = (blk + 255) / 256 < h->sm_count * 4 ? (blk + 255) / 256 : h->sm_count * 4;
  if (C == 1) k_peer_finish<1><<<fb > 0 ? fb : 1, 256, 0, st>>>(f);
  else k_peer_finish<2><<<fb > 0 ? fb : 1, 256, 0, st>>>(f);
  ph_count_launch();
  const int cb = (n_list * C + 255) / 256 < h->sm_count * 2 ? (n_list * C + 255) / 256 : h->sm_count * 2;
  k_peer_collect<<<cb, 256, 0, st>>>(p->win[p->rank], (size_t)1024, p->off_lab, p->off_nobs, p->world,
                                     reinterpret_cast<uint32_t*>(p->d_done + 4), p->d_done + 3, n_list, C, label_out, nobs_out,
                                     p->d_done + 2);
  ph_count_launch();
  PH_CUDA(cudaGetLastError());
  return PH_OK;
}
}  // namespace

extern "C" int ph_peer_assign_linear(ph_peer* p, const uint8_t* cell_label, double lenA, const int32_t* snv_list,
                                     int32_t n_list, uint8_t* label_out, int32_t* nobs_out, void* stream) {
  return peer_assign(p, 1, cell_label, lenA, snv_list, n_list, label_out, nobs_out, stream);
}

extern "C" int ph_peer_assign_branching(ph_peer* p, const uint8_t* cell_label, const int32_t* snv_list, int32_t n_list,
                                        uint8_t* label_out, int32_t* nobs_out, void* stream) {
  return peer_assign(p, 2, cell_label, 1.0, snv_list, n_list, label_out, nobs_out, stream);
}

// ---- SNV-sharded K1 + K1a: every rank holds ALL cells and the SNVs [snv_offset, snv_offset + n_snvs) of total_snvs.
// Per-SNV tables are independent across SNVs, so the pass itself needs no exchange: the kernel epilogue stores each
// batch's labels and counts as self-validating 8-byte words (step counter in the upper half) into every rank's window
// (NVLink peer stores) while it streams on, and every CTA of the SAME kernel then polls its slice of the words in its own
// window and unpacks them into label_out / nobs_out.  ONE launch per step, no fence, no flag, and no CTA ever waits for
// a CTA that is not resident (the pass phase draws its batches from a dispenser).
namespace {
int peer_assign_snv(ph_peer* p, int mode, const uint8_t* cell_label, double lenA, int32_t snv_offset, uint8_t* label_out,
                    int32_t* nobs_out, void* stream) {
  if (!p) PH_FAIL(PH_ERR_ARG, "null peer context");
  if (!cell_label || !label_out || !nobs_out) PH_FAIL(PH_ERR_ARG, "null pointer");
  ph_handle* h = p->h;
  if (snv_offset < 0 || snv_offset + h->n_snvs > p->max_lines)
    PH_FAIL(PH_ERR_ARG, "SNV block [%d, %d) outside 0..%d", snv_offset, snv_offset + h->n_snvs, p->max_lines);
  for (int r = 0; r < p->world; ++r)
    if (!p->win[r]) PH_FAIL(PH_ERR_ARG, "ph_peer_connect has not been called (rank %d missing)", r);
  PH_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int C = mode == 1 ? 1 : 2;
  if (h->n_snvs <= 0) PH_FAIL(PH_ERR_ARG, "this rank holds no SNVs");
  if (h->n_snvs > p->cap_slots - 40)
    PH_FAIL(PH_ERR_UNSUPPORTED, "%d local SNVs exceed the %d a balanced block may hold (ceil(total / world) + 8)", h->n_snvs,
            p->cap_slots - 40);
  if (p->max_lines >= 0xFFFFFF) PH_FAIL(PH_ERR_UNSUPPORTED, "%d SNVs: the pushed words carry 24-bit SNV ids", p->max_lines);
  // LL words alternate between two parities of the step counter: a fast rank may already push step e+1 into a peer that
  // still polls step e (its own step e+2 cannot start before that peer has pushed step e+1, i.e. finished step e).
  PhLL ll;
  memset(&ll, 0, sizeof(ll));
  const size_t word_stride = (size_t)p->world * p->cap_slots * 8;
  for (int r = 0; r < p->world; ++r) {
    ll.dst[r] = p->win[r] + p->off_ll + (size_t)p->rank * p->cap_slots * 8;
    ll.hdr_dst[r] = reinterpret_cast<uint64_t*>(p->win[r] + kHdrOff) + p->rank;
  }
  ll.own = p->win[p->rank] + p->off_ll;
  ll.hdr_own = reinterpret_cast<const uint64_t*>(p->win[p->rank] + kHdrOff);
  ll.word_stride = word_stride;
  ll.par_stride = kLLWords * word_stride;
  ll.cap_slots = p->cap_slots;
  ll.max_lines = p->max_lines;
  ll.label_out = label_out;
  ll.nobs_out = nobs_out;
  ll.done = p->d_done + 3;
  ll.timeout_flag = p->d_done + 16;   // [16] flag, [17..22] what the first failed wait was waiting for
  ll.spin_limit = p->spin_limit;
  uint32_t* epoch_ptr = reinterpret_cast<uint32_t*>(p->d_done + 4);
  int rc = ph_pass_assign_push(h, mode, cell_label, lenA, ll, p->d_done, snv_offset, p->world, epoch_ptr, st);
  if (rc != PH_OK) return rc;
  PH_CUDA(cudaGetLastError());
  return PH_OK;
}
}  // namespace

extern "C" int ph_peer_assign_linear_snv(ph_peer* p, const uint8_t* cell_label, double lenA, int32_t snv_offset,
                                         uint8_t* label_out, int32_t* nobs_out, void* stream) {
  return peer_assign_snv(p, 1, cell_label, lenA, snv_offset, label_out, nobs_out, stream);
}

extern "C" int ph_peer_assign_branching_snv(ph_peer* p, const uint8_t* cell_label, int32_t snv_offset, uint8_t* label_out,
                                            int32_t* nobs_out, void* stream) {
  return peer_assign_snv(p, 2, cell_label, 1.0, snv_offset, label_out, nobs_out, stream);
}

// 1 if a spin-wait of this context ever gave up (a peer never arrived); synchronises the stream
extern "C" int ph_peer_timed_out(ph_peer* p, void* stream) {
  if (!p) return 0;
  unsigned v = 0;
  cudaSetDevice(p->h->device);
  cudaStreamSynchronize((cudaStream_t)stream);
  cudaMemcpy(&v, p->d_done + 2, sizeof(v), cudaMemcpyDeviceToHost);
  unsigned v2 = 0;
  cudaMemcpy(&v2, p->d_done + 16, sizeof(v2), cudaMemcpyDeviceToHost);
  return (v | v2) ? 1 : 0;
}

// what the first wait of an SNV-sharded step that gave up was waiting for:
// out[0..6] = {timed out, kind (1 header, 2 id/label word, 3 / 4 count words), source rank, slot, epoch wanted, epoch seen, payload seen}
// out[7] = this rank's step counter; out[8 + 2 r], out[9 + 2 r] = for source rank r: slots announced by its header of the
// wanted step (0xFFFFFFFF: header absent) and id/label words of that step present in this rank's window
namespace {
__global__ void k_peer_scan(const uint8_t* own_ll, const uint64_t* hdr_own, size_t par_stride, int cap, int world, uint32_t epoch,
                            uint32_t* out) {
  const int r = blockIdx.x;
  if (r >= world) return;
  const int par = (int)(epoch & 1u);
  __shared__ unsigned cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const uint8_t* base = own_ll + (size_t)par * par_stride + (size_t)r * cap * 8;
  unsigned c = 0;
  for (int j = threadIdx.x; j < cap; j += blockDim.x)
    if ((uint32_t)(ld_relaxed_sys_u64(base + (size_t)j * 8) >> 32) == epoch) ++c;
  atomicAdd(&cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint64_t h = ld_relaxed_sys_u64(hdr_own + par * PH_MAX_PEERS + r);
    out[8 + 2 * r] = (uint32_t)(h >> 32) == epoch ? (uint32_t)h : 0xFFFFFFFFu;
    out[9 + 2 * r] = cnt;
  }
}
}  // namespace

extern "C" int ph_peer_debug(ph_peer* p, uint32_t* out) {
  if (!p || !out) PH_FAIL(PH_ERR_ARG, "null argument");
  PH_CUDA(cudaSetDevice(p->h->device));
  PH_CUDA(cudaDeviceSynchronize());
  memset(out, 0, 24 * sizeof(uint32_t));
  PH_CUDA(cudaMemcpy(out, p->d_done + 16, 7 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  PH_CUDA(cudaMemcpy(out + 7, p->d_done + 4, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (out[0]) {
    uint32_t* d_out = reinterpret_cast<uint32_t*>(p->d_done + 32);   // [32..56) of the 64-word block
    PH_CUDA(cudaMemset(d_out, 0, 24 * sizeof(uint32_t)));
    const size_t word_stride = (size_t)p->world * p->cap_slots * 8;
    k_peer_scan<<<p->world, 256>>>(p->win[p->rank] + p->off_ll, reinterpret_cast<const uint64_t*>(p->win[p->rank] + kHdrOff),
                                   kLLWords * word_stride, p->cap_slots, p->world, out[4], d_out);
    PH_CUDA(cudaDeviceSynchronize());
    uint32_t tmp[24];
    PH_CUDA(cudaMemcpy(tmp, d_out, sizeof(tmp), cudaMemcpyDeviceToHost));
    for (int i = 8; i < 8 + 2 * p->world; ++i) out[i] = tmp[i];
  }
  return PH_OK;
}
