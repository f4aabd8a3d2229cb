// The sparse likelihood passes (K1/K1a/K2/K3/a-9/a-11) -- hand-written for sm_100a.
//
// One kernel template walks either orientation of the store (ph_common.cuh, PhOrient).  A warp takes a BATCH of
// consecutive lines (SNVs of the CSC-like copy, cells of the CSR-like copy) and streams their 16-bit entries as one
// contiguous run of 16-byte groups: lane L reads groups L, L+32, ... through a 4-deep register ring of 16-byte
// `ld.global.nc` loads, independent of where lines and sub-segments begin.  The cluster label of every minor index
// lives in shared memory, pre-scaled to a bit shift, so one observation costs  1/8 LDG.128 + PRMT + LDS.U8 + SHF + IADD:
// the lanes only COUNT observations per (code, cluster) in packed 10-bit fields and flush them with shared-memory
// integer atomics when their group leaves a (line, code) run.  The fp64 log-likelihood sums are
// sum_code count * L[code]  evaluated once per line in a fixed order -> exact integer tables, deterministic fp64
// results, no floating-point atomics anywhere.
//
// Whole-orientation passes run on the batched-4 copy of the stream instead (LAY >= 3): four lanes per line, eight lines per
// warp step, sub-segment boundaries uniform across the warp, so the counters are combined with two shuffles and plain
// shared-memory adds -- no atomics at all -- and the dispenser hands out items of 1 - 4 consecutive 8-line batches, so that
// the code around the streaming loop runs once per 8 - 32 lines (launch_pass picks the size).
//
// HBM traffic: 2 B per observation (+ padding to 16 B per sub-segment, + offsets), the roofline of these kernels.
#include "ph_common.cuh"

namespace {

enum { MODE_TABLE = 0, MODE_LINEAR = 1, MODE_BRANCHING = 2, MODE_MAPPED = 3, MODE_COUNTS = 4 };

constexpr int kBlock = 512;  // threads per CTA of the node-table kernel

struct PassArgs {
  const uint16_t* __restrict__ half;
  const uint32_t* __restrict__ sub_off;
  const uint32_t* __restrict__ tail_words;
  const int64_t* __restrict__ tail_off;
  int D, NCH, NSUB, chunk_bits, minor_bits, n_minor, lab_bytes;
  int tab_bytes;  // LAY >= 2: bytes of everything but the label image (they go below the image when they fit)
  // Phased pass (batched-4 stream only): a minor axis whose label image exceeds shared memory (250 k cells, 500 k SNVs) is
  // walked in phases of `ch_n` chunks from `ch_lo`; each phase stages the image of its chunks only and streams the
  // sub-segments of those chunks; the integer (code, cluster) counters of every line travel from phase to phase through
  // `partial` and the last phase finishes the lines (rare-code tail included, its labels gathered from global memory).
  int ch_lo, ch_n, phased, phase_first, phase_last;
  uint32_t* partial;  // [n_lines * D * NC]
  // Split tail (batched-4 stream, one phase): the dispenser hands out batches [0, split_from) whole and every later batch
  // in 2^split_sh parts (equal ranges of its warp steps), each part to whichever warp draws it.  A part adds its integer
  // (line, code, cluster) counters to the batch's row of `part_acc` (integer atomics: exact, order-free) and takes a
  // ticket; the warp holding the last ticket reads the row back (re-arming it) and finishes the batch's lines.  Nobody
  // waits.  It shortens the end of the pass, where warps run out of work one whole batch apart: the last items are
  // a quarter of a batch long.
  int n_items, split_from, split_sh;
  uint32_t* part_acc;   // [(n_batches - split_from) * 8 * D * NC], zero between launches
  unsigned* part_tick;  // [n_batches - split_from], zero between launches
  const uint16_t* __restrict__ half4;     // LAY = 3: the batched-4 copy of the stream (PhOrient)
  const uint32_t* __restrict__ bstep_off;
  const uint8_t* __restrict__ minor_label;  // raw labels [n_minor] (SMEM_LAB) or the prepared label image (global)
  const int32_t* __restrict__ list;         // optional list of lines
  int n_work;
  const double* __restrict__ L0;
  const double* __restrict__ L1;
  const uint8_t* __restrict__ varpos;
  int n_codes;
  int lut_in_smem;
  int C;
  double* S0; double* S1; int32_t* Nobs; int32_t* Nvar;  // TABLE
  double* NobsF;                                         // TABLE: Nobs as double (packed all-reduce buffer)
  uint8_t* label_out; int32_t* nobs_out; double lenA;    // LINEAR / BRANCHING
  const uint8_t* __restrict__ major_label;               // MAPPED: label of each line (255 = skip)
  const uint8_t* __restrict__ map;                       // MAPPED: [K+1][K+1] bytes, != 0 -> "present"
  int K;
  double* per_line;                                      // MAPPED
  unsigned* counter;                                     // batch dispenser: zero at creation, re-armed by the last CTA of every launch
  // COUNTS (cell-sharded K1, ph_peer.cu): every line's (cluster, code) histogram row -> the window of the rank that
  // owns the line (peer memory over NVLink), then one release flag per peer when the whole grid is done
  uint16_t* cnt_dst[PH_MAX_PEERS];   // per owner: this rank's source slot in the owner's window
  uint32_t* flag_dst[PH_MAX_PEERS];  // per peer: where this rank's "partials landed" flag lives
  unsigned* done;                    // CTA completion counter (local, self-resetting)
  int blk, ncp, world, row_words;
  // LINEAR / BRANCHING with push != 0 (SNV-sharded K1, ph_peer.cu): the holders' results leave as self-validating
  // 8-byte words (epoch << 32 | payload, ph_common.cuh PhLL) stored into this rank's region of EVERY rank's window --
  // the B lanes of a batch write B consecutive words per (peer, word): coalesced NVLink stores that overlap the
  // streaming of the following batches.  No fence, no flag: the SAME launch then polls the words of every rank in its
  // own window and unpacks them into the caller's arrays.  Nothing in the pass phase waits for another CTA (batches come
  // from the dispenser), so the kernel does not depend on all of its CTAs being co-resident.
  PhLL ll;
  int push, out_off;
  uint32_t* epoch_w;          // push: the step counter, advanced by the last CTA out of the collect phase
  const uint32_t* epoch_ptr;  // step counter in device memory (step = *epoch_ptr + 1): launches are CUDA-graph replayable
};

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// raw label byte (0 = excluded, 1..NC = cluster) -> bit position of its 10-bit counter field.  Excluded (and any
// label > NC) lands in the field just above the counted ones; for NC = 3 that is bit 30 and simply spills off the top.
template <int NC>
__device__ __forceinline__ uint32_t scale_labels4(uint32_t w) {
  constexpr uint32_t ex = PH_FIELD_BITS * NC;
  constexpr uint32_t t2 = NC >= 2 ? PH_FIELD_BITS : ex, t3 = NC >= 3 ? 2 * PH_FIELD_BITS : ex;
  constexpr uint32_t table = ex | (0u << 8) | (t2 << 16) | (t3 << 24);
  uint32_t sel = (w & 0x3u) | ((w >> 4) & 0x30u) | ((w >> 8) & 0x300u) | ((w >> 12) & 0x3000u);
  return __byte_perm(table, 0u, sel);
}

// node ids: anything >= K (255 = not in the tree) -> K
__device__ __forceinline__ uint32_t clamp_nodes4(uint32_t w, uint32_t K) {
  uint32_t r = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    uint32_t v = (w >> (8 * b)) & 0xFFu;
    v = v < K ? v : K;
    r |= v << (8 * b);
  }
  return r;
}

// Label image of the minor axis (ph_common.cuh): chunk c at byte c << cb = PH_PAD "excluded" bytes + the chunk's labels.
// NODES = false: labels 0..NC -> counter shifts.  NODES = true: node ids clamped to K, padding slots hold K + 1.
template <bool NODES, int NC>
__device__ __forceinline__ void fill_label_image(uint8_t* img, const uint8_t* __restrict__ raw, int n_minor, int NCH, int cb,
                                                 int K, int tid, int nthreads, int ch_lo = 0) {
  const int CH = (1 << cb) - PH_PAD;
  constexpr int PW = PH_PAD / 4;  // words of the excluded region
  const uint32_t excl = NODES ? (uint32_t)(K + 1) : (uint32_t)(PH_FIELD_BITS * NC);
  for (int c = ch_lo; c < ch_lo + NCH; ++c) {  // NCH chunks from ch_lo (a phase), image chunk c - ch_lo
    const int cnt = min(CH, n_minor - c * CH);
    const int nw = (cnt + 3) >> 2;
    const uint8_t* srcb = raw + (size_t)c * CH;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(srcb);  // buffers are padded to 4 B
    uint32_t* dst = reinterpret_cast<uint32_t*>(img + ((size_t)(c - ch_lo) << cb));
    if (tid < PW) dst[tid] = excl * 0x01010101u;
    // whole 16-byte units of an aligned source go through 128-bit loads and stores (4x fewer round trips on the
    // start-up path of every CTA); the rest, and unaligned sources, word by word
    const int nv = (reinterpret_cast<uintptr_t>(srcb) & 15) == 0 ? (cnt >> 4) : 0;
#ifndef PH_STAGE_UNROLL
#define PH_STAGE_UNROLL 4   // 16-byte loads of the staging loop in flight per thread (one L2 round trip for 4 units)
#endif
    constexpr int SU = PH_STAGE_UNROLL;
    for (int i0 = tid; i0 < nv; i0 += SU * nthreads) {
      uint4 w[SU];
#pragma unroll
      for (int u = 0; u < SU; ++u)
        if (i0 + u * nthreads < nv) w[u] = __ldg(reinterpret_cast<const uint4*>(srcb) + i0 + u * nthreads);
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        if (i0 + u * nthreads >= nv) break;
        uint4 o;
        o.x = NODES ? clamp_nodes4(w[u].x, (uint32_t)K) : scale_labels4<NC>(w[u].x);
        o.y = NODES ? clamp_nodes4(w[u].y, (uint32_t)K) : scale_labels4<NC>(w[u].y);
        o.z = NODES ? clamp_nodes4(w[u].z, (uint32_t)K) : scale_labels4<NC>(w[u].z);
        o.w = NODES ? clamp_nodes4(w[u].w, (uint32_t)K) : scale_labels4<NC>(w[u].w);
        reinterpret_cast<uint4*>(dst + PW)[i0 + u * nthreads] = o;
      }
    }
    for (int i = 4 * nv + tid; i < nw; i += nthreads) {
      const uint32_t w = src[i];
      dst[PW + i] = NODES ? clamp_nodes4(w, (uint32_t)K) : scale_labels4<NC>(w);
    }
  }
}

// per-warp shared-memory words: sub-segment bounds [B][NSUB+1], counters [B][D][NC], line ids [B], line labels [B].
// b4 (batched-4 stream): the bounds are the step offsets of the item's B / 8 batches, (B / 8) NSUB + 1 words (two tables
// of NSUB + 1 for the PH_PIPE build) -- what lets 16-line items fit next to a 200 KB label image.
__host__ __device__ inline int warp_words(int B, int NSUB, int D, int NC, bool b4 = false) {
  int bounds = B * (NSUB + 1);
  if (b4) {
    bounds = (B >= 8 ? B / 8 : 1) * NSUB + 1;
    if (bounds < 2 * (NSUB + 1)) bounds = 2 * (NSUB + 1);
  }
  return (bounds + B * D * NC + 2 * B + 3) & ~3;
}

#ifdef PH_STAMPS
// Timeline probe (variant builds only, scripts/probe_stamps.py): per CTA [entry, image staged, last warp left the
// streaming loop, last warp left the kernel] in globaltimer nanoseconds, plus the batches its warps drew.
__device__ unsigned long long g_stamps[256 * 8];
__device__ __forceinline__ unsigned long long stamp_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define PH_STAMP_SET(i) do { if (threadIdx.x == 0 && blockIdx.x < 256) g_stamps[blockIdx.x * 8 + (i)] = stamp_now(); } while (0)
#define PH_STAMP_MAX(i) do { if ((threadIdx.x & 31) == 0 && blockIdx.x < 256) atomicMax(&g_stamps[blockIdx.x * 8 + (i)], stamp_now()); } while (0)
#define PH_STAMP_ADD(i, v) do { if ((threadIdx.x & 31) == 0 && blockIdx.x < 256) atomicAdd(&g_stamps[blockIdx.x * 8 + (i)], (unsigned long long)(v)); } while (0)
#define PH_STAMP_T(var) const unsigned long long var = stamp_now()
extern "C" __attribute__((visibility("default"))) int ph_debug_stamps(unsigned long long* out, int clear) {
  cudaDeviceSynchronize();
  if (out && cudaMemcpyFromSymbol(out, g_stamps, sizeof(g_stamps)) != cudaSuccess) return 1;
  if (clear) {
    static unsigned long long zeros[256 * 8];
    if (cudaMemcpyToSymbol(g_stamps, zeros, sizeof(zeros)) != cudaSuccess) return 1;
  }
  return 0;
}
#else
#define PH_STAMP_SET(i) do { } while (0)
#define PH_STAMP_MAX(i) do { } while (0)
#define PH_STAMP_ADD(i, v) do { } while (0)
#define PH_STAMP_T(var) do { } while (0)
#endif

// One warp walks a batch of B lines handed out by an atomic counter.  After the streaming phase lane (l % B) finishes
// line l alone ("holder"): rare-code tail, fp64 sums, epilogue, store.
// NC = counted classes.  LAY = 0: any chunk stride; 1: production layout (chunk stride 2^16: the label address is one
// PRMT of the entry and the chunk base); 2: as 1 with the label image placed at a shared-memory address that is a
// multiple of 2^16, so that the PRMT yields the complete shared-window address and the gather is a bare LDS.U8 [reg]
// (no base add per observation).  The bytes below the image hold the per-warp tables.  3: as 2, on the batched-4 copy of
// the stream (PhOrient::half4): four lanes stream one line of the batch, sub-segment boundaries are warp-uniform.
// 4: the batched-4 copy with the image wherever it fits (as 1: one base add per observation) -- minor axes whose image
// leaves no room for the 64 KB alignment (K2 on 200 k SNVs: a 200 KB image).  5 / 6: as 3 / 4, in phases (PassArgs).
#ifndef PH_RING4
#define PH_RING4 4
#endif
#ifndef PH_PIPE
#define PH_PIPE 0   // 1: prime the next batch's ring before the holders of the current one.  Measured on C3 (same box,
                    // scripts/probe_r2.py): 168.7 us against 166.7 us for the plain order -- the ring's registers stay live
                    // across the holders and the compiler gives that back in the streaming loop.  Kept as a switch.
#endif
#ifndef PH_TAIL4
#define PH_TAIL4 0  // 1: rare-code tail four words per round trip.  Measured: 172.7 us against 166.7 us (same reason).
#endif
template <int B, int MODE, int NC, bool SMEM_LAB, int LAY>
__global__ void __launch_bounds__(1024, 1) k_pass(const PassArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr bool S16 = LAY >= 1;
  constexpr bool AL = LAY == 2 || LAY == 3 || LAY == 5;
  constexpr bool B4 = LAY >= 3;
  constexpr bool PHASED = LAY >= 5;   // 5 / 6 = 3 / 4 walking the label image in phases (a compile-time switch: the
                                      // plain pass is at the 64-register cap and pays for every extra live value)
  constexpr bool PUSHABLE = MODE == MODE_LINEAR || MODE == MODE_BRANCHING;
  constexpr bool SPLIT = B4 && !PHASED && MODE != MODE_COUNTS && !(PH_PIPE) && B == 8;  // split tail (PassArgs)
  // batched-4 stream: an item is NB8 consecutive 8-line batches of the stream -- ONE run of warp steps over NB8 * NSUB
  // sub-segments -- and the B = 8 * NB8 holder lanes finish its lines together.  The per-item code outside the streaming
  // loop (about 900 SASS instructions: list / offset loads, rare-code tails, fp64 folds and divisions, stores) is issued
  // once per item whatever B is, and a warp streams nothing meanwhile: 5.8 of 26.8 us per 8-line batch on C3.
  constexpr int NB8 = B4 ? (B >= 8 ? B / 8 : 1) : 1;
  static_assert(!B4 || B == 8 || (B % 8 == 0 && !PHASED && !(PH_PIPE)), "batched-4 items are whole 8-line batches");
  const int nthreads = blockDim.x;
  const int D = a.D, NCH = a.NCH, NSUB = a.NSUB, NS1 = NSUB + 1, cb = a.chunk_bits;
  const int K2 = a.K + 2;
  // ---- shared memory carve-up: [label image][map][L0][L1][varpos][per-warp tables]; LAY = 2: the image starts at the
  // next multiple of 2^16 of the shared window and the tables sit below it when they fit there (else behind it)
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t img_base = AL ? ((sbase + 0xFFFFu) & ~0xFFFFu) : sbase;  // shared-window address of the image
  uint8_t* s_lab = smem + (img_base - sbase);
  size_t off = SMEM_LAB ? (size_t)a.lab_bytes : 0;
  if (AL) off = (img_base - sbase) >= (uint32_t)a.tab_bytes ? 0 : (size_t)(img_base - sbase) + (size_t)a.lab_bytes;
  uint8_t* s_map = smem + off;
  if (MODE == MODE_MAPPED) off += ((size_t)K2 * K2 + 15) & ~(size_t)15;
  double* s_L0 = reinterpret_cast<double*>(smem + off);
  double* s_L1 = s_L0 + (a.lut_in_smem ? a.n_codes : 0);
  uint8_t* s_vp = reinterpret_cast<uint8_t*>(s_L1 + (a.lut_in_smem ? a.n_codes : 0));
  if (a.lut_in_smem) off += ((size_t)a.n_codes * 17 + 15) & ~(size_t)15;
  const int wsz = warp_words(B, NSUB, D, NC, B4) + a.row_words;  // both terms are multiples of 4 words
  uint32_t* s_tab0 = reinterpret_cast<uint32_t*>(smem + off);  // start of the per-warp tables
  uint32_t* s_row32 = s_tab0 + (threadIdx.x >> 5) * wsz;
  uint16_t* s_row = reinterpret_cast<uint16_t*>(s_row32);  // COUNTS: [B][NC][ncp] histogram rows of the batch
  uint32_t* s_b = s_row32 + a.row_words;
  const int n_bounds = B4 ? max((B >= 8 ? B / 8 : 1) * NSUB + 1, 2 * NS1) : B * NS1;  // as warp_words()
  uint32_t* s_cnt = s_b + n_bounds;
  int32_t* s_ln = reinterpret_cast<int32_t*>(s_cnt + B * D * NC);
  uint32_t* s_ml = reinterpret_cast<uint32_t*>(s_ln + B);

  PH_STAMP_SET(0);
  unsigned bidx0 = 0;
  if ((threadIdx.x & 31) == 0) bidx0 = atomicAdd(a.counter, 1u);  // first batch of this warp, consumed after the staging
  if (SMEM_LAB)
    fill_label_image<MODE == MODE_MAPPED, NC>(s_lab, a.minor_label, a.n_minor, PHASED ? a.ch_n : NCH, cb, a.K, threadIdx.x, nthreads,
                                              PHASED ? a.ch_lo : 0);
  if (MODE == MODE_MAPPED) {
    // rows: node of the line (K, K+1: not in the tree -> nothing counts); columns: node of the minor index
    // (K: not in the tree -> absent, K+1: padding -> not counted)
    constexpr uint32_t EX = PH_FIELD_BITS * NC;
    for (int i = threadIdx.x; i < K2 * K2; i += nthreads) {
      const int r = i / K2, q = i - r * K2;
      uint32_t v = EX;
      if (r < a.K && q <= a.K) v = (q < a.K && a.map[r * (a.K + 1) + q]) ? PH_FIELD_BITS : 0;
      s_map[i] = (uint8_t)v;
    }
  }
  if (a.lut_in_smem) {
    for (int i = threadIdx.x; i < a.n_codes; i += nthreads) {
      s_L0[i] = a.L0[i];
      s_L1[i] = a.L1[i];
      s_vp[i] = a.varpos[i];
    }
  }
  __syncthreads();
  PH_STAMP_SET(1);
  const double* L0 = a.lut_in_smem ? s_L0 : a.L0;
  const double* L1 = a.lut_in_smem ? s_L1 : a.L1;
  const uint8_t* VP = a.lut_in_smem ? s_vp : a.varpos;
  const uint8_t* lab = SMEM_LAB ? s_lab : a.minor_label;

  constexpr unsigned FULL = 0xffffffffu;
  constexpr unsigned FM = (1u << PH_FIELD_BITS) - 1u;
  const int lane = threadIdx.x & 31;
  const uint32_t mmask = (a.minor_bits >= 32) ? 0xffffffffu : ((1u << a.minor_bits) - 1u);
  const int n_batches = (a.n_work + B - 1) / B;
  const unsigned n_items = SPLIT ? (unsigned)a.n_items : (unsigned)n_batches;  // dispenser range

  const uint32_t epoch = a.epoch_ptr ? *a.epoch_ptr + 1u : 0u;  // this step (peer modes)
  const int par = (int)(epoch & 1u);
  const uint32_t magic_nch = (65536u + (uint32_t)NCH - 1u) / (uint32_t)NCH;      // exact k / NCH for k < 256 * 32 / NCH
  const uint32_t magic_d = D > 0 ? (65536u + (uint32_t)D - 1u) / (uint32_t)D : 0u;  // exact key / D
  const uint32_t magic_nsub = NSUB > 0 ? (65536u + (uint32_t)NSUB - 1u) / (uint32_t)NSUB : 0u;  // exact q / NSUB (q < 4 NSUB)

  auto label_at = [&](uint32_t addr) -> uint32_t {  // addr = offset into the label image
    return SMEM_LAB ? (uint32_t)lab[addr] : (uint32_t)__ldg(lab + addr);
  };
  auto label_dense = [&](uint32_t addr) -> uint32_t {  // LAY = 2: addr is the complete shared-window address
    if (AL) {
      uint32_t v;
      asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));  // pure: the image is read-only after the staging barrier
      return v;
    }
    return label_at(addr);
  };

  // The dispenser round trip never sits on the critical path: the first batch is drawn before the staging above is
  // published (see `bidx0`), every following one while the holders finish the current batch.
  unsigned bidx = __shfl_sync(FULL, bidx0, 0);
  // batched-4 stream: ring of RD 16-byte loads per lane in flight (RD * 512 B per warp).  The ring of the NEXT batch is
  // primed before the holders finish the current one (`primed`), so the tail / fp64 / store phase of a batch overlaps the
  // first loads of its successor instead of leaving the warp without a byte in flight.
  constexpr int RD = PH_RING4;
  uint4 ring[B4 ? RD : 1];
  bool primed = false;
  int ebuf = 0;  // which of the two sub-segment bound tables (s_b) belongs to the current batch
  if (PUSHABLE && a.push && bidx == 0 && lane < a.world) {
    // the warp that drew the first batch announces how many slots this rank fills in this step
    st_relaxed_sys_u64(a.ll.hdr_dst[lane] + par * PH_MAX_PEERS, ((uint64_t)epoch << 32) | (uint32_t)(n_batches * B));
  }
  while (bidx < n_items) {
    PH_STAMP_T(t_top);
    unsigned bat = bidx;          // batch of this item
    uint32_t part = 0;            // split tail: which of the batch's parts
    bool split = false;
    if (SPLIT && bidx >= (unsigned)a.split_from) {
      const unsigned r = bidx - (unsigned)a.split_from;
      bat = (unsigned)a.split_from + (r >> a.split_sh);
      part = r & ((1u << a.split_sh) - 1u);
      split = true;
    }
    const int first = (int)bat * B;
    const int nl = min(B, a.n_work - first);

    // ---- my line (lane % B): holder duties after the streaming phase
    const int myl = lane % B;
    const int myw = first + myl;
    int line = -1;
    if (myl < nl) line = a.list ? __ldg(a.list + myw) : myw;  // (batched-4: the list is the shape order, -1 = empty slot)
    const bool have = line >= 0;
    uint32_t h_ml = 0;
    bool h_skip = !have;
    if (MODE == MODE_MAPPED) {
      h_ml = (uint32_t)a.K;
      if (have) {
        h_ml = a.major_label[line];
        if (h_ml >= (uint32_t)a.K) {  // cell not in the tree: contributes nothing
          h_skip = true;
          h_ml = (uint32_t)a.K;
        }
      }
    }
    if (lane < B) {
      s_ln[lane] = line;
      s_ml[lane] = h_ml;
    }
    // the holder's tail range: offsets loaded now, words pulled into L2 while the dense part streams
    int64_t h_tb = 0, h_te = 0;
    if (lane < B && have) {
      h_tb = __ldg(a.tail_off + line);
      h_te = __ldg(a.tail_off + line + 1);
      if (part == 0)
        for (int64_t r = h_tb; r < h_te; r += 32)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.tail_words + r));
    }
    for (int k = lane; k < B * D * NC; k += 32) s_cnt[k] = 0;
    if (MODE == MODE_COUNTS)
      for (int k = lane; k < a.row_words; k += 32) s_row32[k] = 0;
    __syncwarp();
    // s_e[k]: end (in groups) of sub-segment k = l * NSUB + s of the batch; s_start[l]: first group of line l
    uint32_t* s_e = s_b + (B4 ? ebuf * NS1 : 0);
    uint32_t* s_start = s_b + B * NSUB;
    if (B4 && !primed) {
      // batched-4 copy: s_e[q] = first warp step of sub-segment q of this item (NB8 batches in a row), s_e[NB8 * NSUB] =
      // its end; past the last batch of the stream every bound is the end of the stream (empty sub-segments)
      const size_t nbs = (size_t)(a.n_work >> 3) * NSUB;
      for (int k = lane; k <= NB8 * NSUB; k += 32) {
        const size_t at = (size_t)bat * (NB8 * NSUB) + k;
        s_e[k] = __ldg(a.bstep_off + (NB8 > 1 && at > nbs ? nbs : at));
      }
    }
    for (int k = lane; k < (B4 ? 0 : B * NSUB); k += 32) {
      const int l = k / NSUB, s = k - l * NSUB;
      const int ln = s_ln[l];
      s_e[k] = ln >= 0 ? __ldg(a.sub_off + (size_t)ln * NSUB + s + 1) : 0u;
    }
    if (!B4 && lane < B) s_start[lane] = (line >= 0 && NSUB > 0) ? __ldg(a.sub_off + (size_t)line * NSUB) : 0u;
    __syncwarp();

    unsigned bnext = 0;
    bool drawn = false;
    PH_STAMP_T(t_stream);
    // ---- dense part, batched-4 copy: the batch is one contiguous run of warp steps; lane (a, c) = (lane / 4, lane % 4)
    // reads group 4 t + c of line a at step t.  Everything about sub-segments is warp-uniform: q, the chunk base, the
    // flush points.  A lane's packed counters are summed over the 4 lanes of its line and added to the line's
    // (code, cluster) counter by one of them -- no MATCH / REDUX / atomics.
    if (B4 && NSUB > 0) {
      const int la = lane >> 2;  // my line of the (current) batch
      int cur_sb = 0;            // NB8 > 1: which batch of the item the run is in
      // a full pass is ONE run of steps over all sub-segments; a phase walks, per dense code, the run of its chunks
      constexpr bool phased = PHASED;
      const int ch_lo = phased ? a.ch_lo : 0;
      uint32_t st = 0, st_end = 0;
      const uint4* gp = nullptr;
      int q = 0, cur_d = 0;
      uint32_t nb_eff = 0, cbase = 0, acc = 0, st_flush = 0;
      const uint8_t* maprow = s_map;
      if (MODE == MODE_MAPPED) maprow = s_map + s_ml[la] * K2;
      constexpr uint32_t kDeadline4 = PH_FLUSH_GROUPS;  // steps: one group per lane and step
      auto flush4 = [&]() {  // warp-uniform call
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          uint32_t v = (acc >> (PH_FIELD_BITS * c)) & FM;
          v += __shfl_xor_sync(FULL, v, 1);
          v += __shfl_xor_sync(FULL, v, 2);
          if ((lane & 3) == 0 && v) s_cnt[((la + 8 * cur_sb) * D + cur_d) * NC + c] += v;
        }
        acc = 0;
      };
      auto slow4 = [&](uint32_t step) {  // warp-uniform: step and all bounds are
        while (step >= s_e[q + 1]) ++q;
        int qs = q, sb = 0;  // sub-segment within its batch, batch within the item
        if (NB8 > 1) {
          sb = (int)(((uint32_t)q * magic_nsub) >> 16);
          qs = q - sb * NSUB;
        }
        const int d = (int)(((uint32_t)qs * magic_nch) >> 16);  // qs / NCH
        if (d != cur_d || (NB8 > 1 && sb != cur_sb) || step - st_flush >= kDeadline4) {
          flush4();
          cur_d = d;
          if (NB8 > 1) {
            cur_sb = sb;
            if (MODE == MODE_MAPPED) maprow = s_map + s_ml[la + 8 * sb] * K2;
          }
          st_flush = step;
        }
        cbase = (((uint32_t)qs - (uint32_t)d * (uint32_t)NCH - (uint32_t)ch_lo) << 16) + (AL ? img_base : 0u);
        nb_eff = min(s_e[q + 1], st_flush + kDeadline4);
      };
      auto one4 = [&](uint32_t addr) -> uint32_t {
        uint32_t sh = label_dense(addr);
        if (MODE == MODE_MAPPED) sh = maprow[sh];
        return 1u << sh;
      };
      auto consume4 = [&](const uint4& w, uint32_t step) {
        if (step >= nb_eff) slow4(step);
        const uint32_t a0 = __byte_perm(w.x, cbase, 0x7610), a1 = __byte_perm(w.x, cbase, 0x7632);
        const uint32_t a2 = __byte_perm(w.y, cbase, 0x7610), a3 = __byte_perm(w.y, cbase, 0x7632);
        const uint32_t a4 = __byte_perm(w.z, cbase, 0x7610), a5 = __byte_perm(w.z, cbase, 0x7632);
        const uint32_t a6 = __byte_perm(w.w, cbase, 0x7610), a7 = __byte_perm(w.w, cbase, 0x7632);
        acc += ((one4(a0) + one4(a1)) + (one4(a2) + one4(a3))) + ((one4(a4) + one4(a5)) + (one4(a6) + one4(a7)));
      };
      const int n_runs = phased ? D : 1;
      for (int run = 0; run < n_runs; ++run) {
        q = phased ? run * NCH + ch_lo : 0;
        st = s_e[q];
        st_end = phased ? s_e[run * NCH + ch_lo + a.ch_n] : s_e[NB8 * NSUB];
        if (SPLIT && split) {  // this item's share of the batch's steps
          const uint32_t len = st_end - st, lo = st + ((len * part) >> a.split_sh);
          st_end = st + ((len * (part + 1u)) >> a.split_sh);
          st = lo;
        }
        gp = reinterpret_cast<const uint4*>(a.half4) + ((size_t)st << 5) + lane;
        nb_eff = st;  // the first step of a run enters the slow path
        st_flush = st;
        if (!primed) {
#pragma unroll
          for (int u = 0; u < RD; ++u) {
            ring[u] = make_uint4(0, 0, 0, 0);
            if (st + u < st_end) ring[u] = ld_stream(gp + 32 * u);
          }
        }
        primed = false;
        while (st < st_end) {
          // the next batch is drawn shortly before this one ends: the dispenser's round trip hides behind the last steps
          if (PH_PIPE && !drawn && run == n_runs - 1 && st + 16 >= st_end) {
            if (lane == 0) bnext = atomicAdd(a.counter, 1u);
            drawn = true;
          }
#pragma unroll
          for (int u = 0; u < RD; ++u) {
            if (u == 0 || st + u < st_end) {
              consume4(ring[u], st + u);
              if (st + u + RD < st_end) ring[u] = ld_stream(gp + 32 * (u + RD));
            }
          }
          st += RD;
          gp += 32 * RD;
        }
        flush4();
      }
    }
    // ---- dense part: maximal runs of lines that are contiguous in memory are streamed as one range of groups
    if (!B4 && NSUB > 0) {
      int l0 = 0;
      while (l0 < nl) {
        int l1 = l0 + 1;
        while (l1 < nl && s_start[l1] == s_e[l1 * NSUB - 1]) ++l1;
        const uint32_t gbeg = s_start[l0], gend = s_e[l1 * NSUB - 1];
        uint32_t g = gbeg + lane;
        const uint4* gp = reinterpret_cast<const uint4*>(a.half) + g;
        // Per-lane cursor: k = current sub-segment.  Everything derived from it -- the counter key = k / NCH =
        // line * D + code, the chunk base of the label address -- only changes when the lane's group index passes
        // nb_eff = min(end of sub-segment k, overflow deadline of the packed counters), so the common path of a
        // group is one compare; the slow path advances k, flushes the counters if the (line, code) key changed.
        int k = l0 * NSUB;
        uint32_t nb_eff = 0, cur_key = 0xffffffffu, cbase = 0, acc = 0, gflush = gbeg;
        const uint8_t* maprow = s_map;
        constexpr uint32_t kDeadline = PH_FLUSH_GROUPS * 32u;

        // lanes leave a (line, code) run within a step or two of each other: the ones flushing the same counter
        // together are summed with REDUX first, so the shared-memory atomic sees one request instead of up to 32
        auto flush = [&]() {
          const unsigned peers = __match_any_sync(__activemask(), cur_key);
          const bool leader = lane == __ffs(peers) - 1;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const uint32_t f = __reduce_add_sync(peers, (acc >> (PH_FIELD_BITS * c)) & FM);
            if (leader && f) atomicAdd(&s_cnt[cur_key * NC + c], f);
          }
        };
        auto slow = [&](uint32_t gg) {
          uint32_t nb = s_e[k];
          while (gg >= nb) nb = s_e[++k];
          const uint32_t key = ((uint32_t)k * magic_nch) >> 16;  // k / NCH
          if (key != cur_key || gg - gflush >= kDeadline) {
            if (acc) flush();
            acc = 0;
            cur_key = key;
            gflush = gg;
            if (MODE == MODE_MAPPED) maprow = s_map + s_ml[(key * magic_d) >> 16] * K2;  // line = key / D
          }
          cbase = (((uint32_t)k - key * (uint32_t)NCH) << cb) + (AL ? img_base : 0u);
          nb_eff = min(nb, gflush + kDeadline);
        };
        auto one = [&](uint32_t addr) -> uint32_t {
          uint32_t sh = label_dense(addr);
          if (MODE == MODE_MAPPED) sh = maprow[sh];
          return 1u << sh;
        };
        auto consume = [&](const uint4& w, uint32_t gg) {
          if (gg >= nb_eff) slow(gg);
          uint32_t a0, a1, a2, a3, a4, a5, a6, a7;
          if (S16) {
            a0 = __byte_perm(w.x, cbase, 0x7610); a1 = __byte_perm(w.x, cbase, 0x7632);
            a2 = __byte_perm(w.y, cbase, 0x7610); a3 = __byte_perm(w.y, cbase, 0x7632);
            a4 = __byte_perm(w.z, cbase, 0x7610); a5 = __byte_perm(w.z, cbase, 0x7632);
            a6 = __byte_perm(w.w, cbase, 0x7610); a7 = __byte_perm(w.w, cbase, 0x7632);
          } else {
            a0 = (w.x & 0xFFFFu) | cbase; a1 = (w.x >> 16) | cbase;
            a2 = (w.y & 0xFFFFu) | cbase; a3 = (w.y >> 16) | cbase;
            a4 = (w.z & 0xFFFFu) | cbase; a5 = (w.z >> 16) | cbase;
            a6 = (w.w & 0xFFFFu) | cbase; a7 = (w.w >> 16) | cbase;
          }
          acc += ((one(a0) + one(a1)) + (one(a2) + one(a3))) + ((one(a4) + one(a5)) + (one(a6) + one(a7)));
        };

        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, r2 = r0, r3 = r0;
        if (g < gend) r0 = ld_stream(gp);
        if (g + 32 < gend) r1 = ld_stream(gp + 32);
        if (g + 64 < gend) r2 = ld_stream(gp + 64);
        if (g + 96 < gend) r3 = ld_stream(gp + 96);
        while (g < gend) {
          consume(r0, g);
          if (g + 128 < gend) r0 = ld_stream(gp + 128);
          if (g + 32 < gend) {
            consume(r1, g + 32);
            if (g + 160 < gend) r1 = ld_stream(gp + 160);
          }
          if (g + 64 < gend) {
            consume(r2, g + 64);
            if (g + 192 < gend) r2 = ld_stream(gp + 192);
          }
          if (g + 96 < gend) {
            consume(r3, g + 96);
            if (g + 224 < gend) r3 = ld_stream(gp + 224);
          }
          g += 128;
          gp += 128;
        }
        if (acc) flush();
        l0 = l1;
      }
    }
    __syncwarp();
    PH_STAMP_T(t_hold);
    if (!drawn && lane == 0) bnext = atomicAdd(a.counter, 1u);  // next batch: the round trip overlaps the holders' work
    bool finisher = true;
    if (SPLIT && split) {
      const unsigned sb = bat - (unsigned)a.split_from;
      uint32_t* gacc = a.part_acc + (size_t)sb * (B * D * NC);
      for (int k = lane; k < B * D * NC; k += 32) {
        const uint32_t v = s_cnt[k];
        if (v) atomicAdd(gacc + k, v);
      }
      __threadfence();
      __syncwarp();
      unsigned t = 0;
      if (lane == 0) t = atomicAdd(a.part_tick + sb, 1u);
      t = __shfl_sync(FULL, t, 0);
      finisher = t == (1u << a.split_sh) - 1u;
      if (finisher) {  // every part's counters are in the row: take them (and leave zeros for the next launch)
        __threadfence();
        for (int k = lane; k < B * D * NC; k += 32) s_cnt[k] = atomicExch(gacc + k, 0u);
        if (lane == 0) a.part_tick[sb] = 0u;
        __syncwarp();
      }
    }
    if (PH_PIPE && B4 && NSUB > 0 && !PHASED) {
      // prime the ring of the next batch now: its first loads fly while the holders below finish this one
      const unsigned bn = __shfl_sync(FULL, bnext, 0);
      if (bn < (unsigned)n_batches) {
        uint32_t* s_n = s_b + (ebuf ^ 1) * NS1;
        for (int k = lane; k <= NSUB; k += 32) s_n[k] = __ldg(a.bstep_off + (size_t)bn * NSUB + k);
        __syncwarp();
        const uint32_t st_n = s_n[0], end_n = s_n[NSUB];
        const uint4* gp_n = reinterpret_cast<const uint4*>(a.half4) + ((size_t)st_n << 5) + lane;
#pragma unroll
        for (int u = 0; u < RD; ++u) {
          ring[u] = make_uint4(0, 0, 0, 0);
          if (st_n + u < end_n) ring[u] = ld_stream(gp_n + 32 * u);
        }
        primed = true;
        ebuf ^= 1;
      }
    }

    // ---- holders: lane l finishes line l.  Every sum is  sum_code fma(count, L[code])  in ascending code order --
    // the dense codes from the batch counters, then the rare codes of the tail (words sorted by (code, address), so a
    // code is one run) -- i.e. a function of the integer (code, cluster) histogram only: the result does not depend on
    // how the observations were split over lanes, batches or GPUs.
    uint32_t pay0 = 0xFFFFFF00u, pay1 = 0, pay2 = 0;  // push: this lane's line as LL payloads (empty slot by default)
    if (lane < B && have && finisher) {
      double s0[NC], s1[NC];
      unsigned nob[NC], nva[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) { s0[c] = 0.0; s1[c] = 0.0; nob[c] = 0; nva[c] = 0; }
      constexpr bool ph = PHASED;
      for (int d = 0; d < D; ++d) {
        const double l0 = L0[d], l1 = L1[d];
        const bool vp = VP[d] != 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          unsigned t = s_cnt[(myl * D + d) * NC + c];
          if (ph) {  // the line's counters of the earlier phases; not the last phase: hand them on
            uint32_t* pp = a.partial + ((size_t)line * D + d) * NC + c;
            if (!a.phase_first) t += *pp;
            if (!a.phase_last) *pp = t;
          }
          if (MODE == MODE_COUNTS) s_row[(myl * NC + c) * a.ncp + d] = (uint16_t)t;
          s0[c] = fma((double)t, l0, s0[c]);
          s1[c] = fma((double)t, l1, s1[c]);
          nob[c] += t;
          nva[c] += vp ? t : 0u;
        }
      }
      {
        const uint8_t* h_maprow = (MODE == MODE_MAPPED) ? s_map + h_ml * K2 : nullptr;
        const int64_t tb = h_tb, te = h_te;
        uint32_t cur = 0xffffffffu;
        unsigned rc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) rc[c] = 0;
        auto fold = [&]() {
          const double l0 = L0[cur], l1 = L1[cur];
          const bool vp = VP[cur] != 0;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            if (MODE == MODE_COUNTS) s_row[(myl * NC + c) * a.ncp + cur] = (uint16_t)rc[c];
            s0[c] = fma((double)rc[c], l0, s0[c]);
            s1[c] = fma((double)rc[c], l1, s1[c]);
            nob[c] += rc[c];
            nva[c] += vp ? rc[c] : 0u;
            rc[c] = 0;
          }
        };
        auto tail_one = [&](uint32_t w) {
          const uint32_t code = w >> a.minor_bits;
          uint32_t sh;
          if (ph) {
            // the shared image holds this phase's chunks only: the rare entries (about 1 %) gather their raw label from
            // global memory -- address in the full image -> minor index -> label -> counter shift / clamped node
            const uint32_t ad = w & mmask, ch = ad >> cb, ix = ad & ((1u << cb) - 1u);
            const uint32_t raw = __ldg(a.minor_label + (size_t)ch * ((1u << cb) - PH_PAD) + (ix - PH_PAD));
            if (MODE == MODE_MAPPED) sh = raw < (uint32_t)a.K ? raw : (uint32_t)a.K;
            else sh = (raw >= 1u && raw <= (uint32_t)NC) ? PH_FIELD_BITS * (raw - 1u) : (uint32_t)(PH_FIELD_BITS * NC);
          } else {
            sh = label_at(w & mmask);
          }
          if (MODE == MODE_MAPPED) sh = h_maprow[sh];
          if (code != cur) {
            if (cur != 0xffffffffu) fold();
            cur = code;
          }
#pragma unroll
          for (int c = 0; c < NC; ++c) rc[c] += (sh == (uint32_t)(PH_FIELD_BITS * c)) ? 1u : 0u;
        };
        // four words per round trip (they sit in L2: prefetched when the batch started); the array has 8 words of slack
        const int64_t te_eff = (ph && !a.phase_last) ? tb : te;
#if PH_TAIL4
        for (int64_t r0 = tb; r0 < te_eff; r0 += 4) {
          uint32_t w4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) w4[u] = __ldg(a.tail_words + r0 + u);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (r0 + u < te_eff) tail_one(w4[u]);
        }
#else
        for (int64_t r = tb; r < te_eff; ++r) tail_one(__ldg(a.tail_words + r));
#endif
        if (cur != 0xffffffffu) fold();
      }
      const int wi = B4 ? line : myw;  // batched-4 walks the lines in shape order: results go to the line's own slot
      if (ph && !a.phase_last) {
        // an earlier phase: the counters went to `partial`, the line is finished by the last phase
      } else if (MODE == MODE_TABLE) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const size_t o = (size_t)wi * NC + c;
          if (a.S0) a.S0[o] = s0[c];
          if (a.S1) a.S1[o] = s1[c];
          if (a.Nobs) a.Nobs[o] = (int32_t)nob[c];
          if (a.Nvar) a.Nvar[o] = (int32_t)nva[c];
          if (a.NobsF) a.NobsF[o] = (double)nob[c];
        }
      } else if (MODE == MODE_LINEAR) {
        // linear_split.py:229-234  like0.mean(axis=0) >= like1.mean(axis=0) -> mutsB (ties -> mutsB)
        const double m0 = s0[0] / a.lenA, m1 = s1[0] / a.lenA;
        const uint8_t lb = (m0 >= m1) ? 2 : 1;
        if (a.push) {
          pay0 = ((uint32_t)(a.out_off + line) << 8) | lb;
          pay1 = nob[0];
        } else {
          a.label_out[wi] = lb;
          a.nobs_out[wi] = (int32_t)nob[0];
        }
      } else if (MODE == MODE_BRANCHING) {
        // branching_split.py:332-346
        const double cA = s1[0] + s0[NC > 1 ? 1 : 0];
        const double cB = s0[0] + s1[NC > 1 ? 1 : 0];
        const double cC = s1[0] + s1[NC > 1 ? 1 : 0];
        const double mx = fmax(cA, fmax(cB, cC));
        const uint8_t lb = (cA == mx) ? 1 : ((cB == mx) ? 2 : 3);
        if (a.push) {
          pay0 = ((uint32_t)(a.out_off + line) << 8) | lb;
          pay1 = nob[0];
          pay2 = nob[NC > 1 ? 1 : 0];
        } else {
          a.label_out[wi] = lb;
          a.nobs_out[2 * (size_t)wi] = (int32_t)nob[0];
          a.nobs_out[2 * (size_t)wi + 1] = (int32_t)nob[NC > 1 ? 1 : 0];
        }
      } else if (MODE == MODE_MAPPED) {  // absent -> field 0 uses like0, present -> field 1 uses like1
        a.per_line[wi] = h_skip ? 0.0 : s0[0] + s1[NC > 1 ? 1 : 0];
      }
    }
    __syncwarp();
    if (PUSHABLE && a.push && finisher) {
      // the batch's LL words -> every rank's window: lane = (peer * words + word, line); the B lanes of a (peer, word)
      // store B consecutive 8-byte words
      constexpr int NW = MODE == MODE_LINEAR ? 2 : 3;
      constexpr int KPI = 32 / B;  // (peer, word) pairs per warp instruction
      const int l = lane % B;
      const uint32_t v0 = __shfl_sync(FULL, pay0, l), v1 = __shfl_sync(FULL, pay1, l);
      const uint32_t v2 = NW == 3 ? __shfl_sync(FULL, pay2, l) : 0u;
      const size_t slot_off = ((size_t)bat * B + l) * 8 + (size_t)par * a.ll.par_stride;
      for (int kk = lane < KPI * B ? lane / B : a.world * NW; kk < a.world * NW; kk += KPI) {
        const int p = kk / NW, wd = kk - p * NW;
        const uint32_t v = wd == 0 ? v0 : (wd == 1 ? v1 : v2);
        st_relaxed_sys_u64(a.ll.dst[p] + slot_off + (size_t)wd * a.ll.word_stride, ((uint64_t)epoch << 32) | v);
      }
    }
    if (MODE == MODE_COUNTS) {
      // the batch's rows are consecutive in the owner's slot: 16-byte stores straight into peer memory
      const int upl = NC * a.ncp / 8;  // 16-byte units per line
      for (int u = lane; u < nl * upl; u += 32) {
        const int l = u / upl;
        const int k = first + l;
        const int owner = k / a.blk;
        uint4* dst = reinterpret_cast<uint4*>(a.cnt_dst[owner] + (size_t)(k - owner * a.blk) * NC * a.ncp) + (u - l * upl);
        *dst = reinterpret_cast<const uint4*>(s_row)[u];
      }
      __syncwarp();
    }
    PH_STAMP_ADD(4, 1);
    bidx = __shfl_sync(FULL, bnext, 0);
#ifdef PH_STAMPS
    {  // per-CTA sums over warps and items: prologue, streaming, holders (+ merge, push, next-item round trip)
      PH_STAMP_T(t_end);
      PH_STAMP_ADD(5, t_stream - t_top);
      PH_STAMP_ADD(6, t_hold - t_stream);
      PH_STAMP_ADD(7, t_end - t_hold);
    }
#endif
  }
  PH_STAMP_MAX(2);
  // The last CTA to finish re-arms the batch dispenser for the next launch (no memset between launches).  Peer modes:
  // all partial rows (or pushed records) of this rank are on their way -- make them visible system-wide first; the last
  // CTA then raises this rank's flag in every peer's window.
  const bool peer_mode = MODE == MODE_COUNTS;
  __syncthreads();
  if (threadIdx.x == 0) {
    // one system-scope fence per CTA: the barrier orders the other threads' peer stores before it (cumulativity)
    if (peer_mode) __threadfence_system();
    const unsigned prev = atomicAdd(a.done, 1u);
    if (prev == gridDim.x - 1) {
      *a.done = 0;
      *a.counter = 0;
      if (peer_mode) {
        __threadfence_system();
        for (int p = 0; p < a.world; ++p)
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.flag_dst[p]), "r"(epoch) : "memory");
      }
    }
  }
  if (PUSHABLE && a.push) {
    // ---- collect phase, same launch: poll the LL words of every rank in this rank's own window (they carry the step
    // counter, so the data validates itself) and unpack them into the caller's arrays.  A CTA only waits for words that
    // resident CTAs of some rank push during their pass phase -- never for a CTA that has not started yet.
    constexpr int NW = MODE == MODE_LINEAR ? 2 : 3;
    int* s_cnt_r = reinterpret_cast<int*>(s_tab0);  // the per-warp tables are free now (all warps passed the barrier above)
    const long long t0 = clock64();
    bool ok = true;
    if (threadIdx.x < a.world) {
      const uint32_t c = ll_wait(a.ll.hdr_own + par * PH_MAX_PEERS + threadIdx.x, epoch, t0, a.ll.spin_limit, a.ll.timeout_flag, &ok, 1u,
                                 threadIdx.x, 0u);
      s_cnt_r[threadIdx.x] = ok ? min((int)c, a.ll.cap_slots) : 0;
    }
    __syncthreads();
    const uint8_t* base = a.ll.own + (size_t)par * a.ll.par_stride;
    const int cap = a.ll.cap_slots;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < a.world * cap; idx += gridDim.x * blockDim.x) {
      const int r = idx / cap, j = idx - r * cap;
      if (j >= s_cnt_r[r]) continue;
      const uint8_t* w = base + ((size_t)r * cap + j) * 8;
      ok = true;
      const uint32_t p0 = ll_wait(w, epoch, t0, a.ll.spin_limit, a.ll.timeout_flag, &ok, 2u, (unsigned)r, (unsigned)j);
      const uint32_t g = p0 >> 8;
      if (!ok || g >= (uint32_t)a.ll.max_lines) continue;  // empty slot (id 0xFFFFFF) or a word that never arrived
      const uint32_t n0 = ll_wait(w + a.ll.word_stride, epoch, t0, a.ll.spin_limit, a.ll.timeout_flag, &ok, 3u, (unsigned)r, (unsigned)j);
      uint32_t n1 = 0;
      if (NW == 3) n1 = ll_wait(w + 2 * a.ll.word_stride, epoch, t0, a.ll.spin_limit, a.ll.timeout_flag, &ok, 4u, (unsigned)r, (unsigned)j);
      if (!ok) continue;
      a.ll.label_out[g] = (uint8_t)(p0 & 0xFFu);
      if (NW == 2) {
        a.ll.nobs_out[g] = (int32_t)n0;
      } else {
        a.ll.nobs_out[2 * (size_t)g] = (int32_t)n0;
        a.ll.nobs_out[2 * (size_t)g + 1] = (int32_t)n1;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned prev = atomicAdd(a.ll.done, 1u);
      if (prev == gridDim.x - 1) {
        *a.ll.done = 0;
        *a.epoch_w = epoch;  // next step
      }
    }
  }
  PH_STAMP_MAX(3);
}

// label image in global memory, for minor axes whose image does not fit in shared memory
template <bool NODES, int NC>
__global__ void k_label_image(uint8_t* img, const uint8_t* __restrict__ raw, int n_minor, int NCH, int cb, int K) {
  fill_label_image<NODES, NC>(img, raw, n_minor, NCH, cb, K, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// deterministic block sums of a per-line table by the line's own label: out[(q-1)*C + c]
__global__ void k_block_reduce(const double* __restrict__ S0, const double* __restrict__ S1,
                               const int32_t* __restrict__ Nobs, const int32_t* __restrict__ Nvar,
                               const uint8_t* __restrict__ line_label, int n_lines, int C, double* out0, double* out1,
                               long long* outn, long long* outv) {
  const int q = blockIdx.x / C + 1, c = blockIdx.x % C;
  double a0 = 0.0, a1 = 0.0;
  long long an = 0, av = 0;
  for (int j = threadIdx.x; j < n_lines; j += blockDim.x) {
    if (line_label[j] == q) {
      const size_t o = (size_t)j * C + c;
      a0 += S0[o];
      a1 += S1[o];
      an += Nobs[o];
      av += Nvar[o];
    }
  }
  __shared__ double r0[1024], r1[1024];
  __shared__ long long rn[1024], rv[1024];
  r0[threadIdx.x] = a0; r1[threadIdx.x] = a1; rn[threadIdx.x] = an; rv[threadIdx.x] = av;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      r0[threadIdx.x] += r0[threadIdx.x + s];
      r1[threadIdx.x] += r1[threadIdx.x + s];
      rn[threadIdx.x] += rn[threadIdx.x + s];
      rv[threadIdx.x] += rv[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out0[blockIdx.x] = r0[0]; out1[blockIdx.x] = r1[0]; outn[blockIdx.x] = rn[0]; outv[blockIdx.x] = rv[0];
  }
}

// deterministic sums of a per-line table [n_lines x ncol] over the lines of every label 1..nlab:
// out[col * nlab + (lab - 1)].  Fixed order (thread-strided partial sums, then a tree): a function of the table alone, so
// a table assembled from several GPUs reduces to the very bits a single GPU gets (sharded_store.py).
__global__ void k_line_reduce(const double* __restrict__ S0, const double* __restrict__ S1, const int32_t* __restrict__ Nobs,
                              const int32_t* __restrict__ Nvar, const uint8_t* __restrict__ line_label, int n_lines, int ncol,
                              int nlab, double* out0, double* out1, long long* outn, long long* outv) {
  const int col = blockIdx.x / nlab, lab = blockIdx.x % nlab + 1;
  double a0 = 0.0, a1 = 0.0;
  long long an = 0, av = 0;
  for (int j = threadIdx.x; j < n_lines; j += blockDim.x) {
    if (line_label[j] == lab) {
      const size_t o = (size_t)j * ncol + col;
      a0 += S0[o];
      a1 += S1[o];
      an += Nobs[o];
      av += Nvar[o];
    }
  }
  __shared__ double r0[1024], r1[1024];
  __shared__ long long rn[1024], rv[1024];
  r0[threadIdx.x] = a0; r1[threadIdx.x] = a1; rn[threadIdx.x] = an; rv[threadIdx.x] = av;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      r0[threadIdx.x] += r0[threadIdx.x + s];
      r1[threadIdx.x] += r1[threadIdx.x + s];
      rn[threadIdx.x] += rn[threadIdx.x + s];
      rv[threadIdx.x] += rv[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out0[blockIdx.x] = r0[0]; out1[blockIdx.x] = r1[0]; outn[blockIdx.x] = rn[0]; outv[blockIdx.x] = rv[0];
  }
}

// deterministic per-node sum of per-line values
__global__ void k_node_reduce(const double* __restrict__ v, const uint8_t* __restrict__ label, int n, double* out) {
  const int k = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (label[i] == k) acc += v[i];
  __shared__ double r[1024];
  r[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) r[threadIdx.x] += r[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[k] = r[0];
}

// ---- general node tables (a-11): per line, sums per minor node r in 0..K (K = "not in tree"), then
//      T[line,k] = sum_r ( member[k][r] ? S1[r] : S0[r] ).  G lanes per line, counters in shared memory.
template <int G, bool SMEM_LAB>
__global__ void __launch_bounds__(kBlock, 1) k_node_table(const PassArgs a, const uint8_t* __restrict__ member, double* T) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int K = a.K, K1 = K + 1;
  constexpr int groups_per_block = kBlock / G;
  size_t off = 0;
  uint8_t* s_lab = smem;
  if (SMEM_LAB) off = (size_t)a.lab_bytes;
  double* s_acc0 = reinterpret_cast<double*>(smem + off);
  double* s_acc1 = s_acc0 + groups_per_block * K1;
  unsigned* s_cnt = reinterpret_cast<unsigned*>(s_acc1 + groups_per_block * K1);
  uint8_t* s_mem = reinterpret_cast<uint8_t*>(s_cnt + groups_per_block * K1);
  if (SMEM_LAB) fill_label_image<true, 1>(s_lab, a.minor_label, a.n_minor, a.NCH, a.chunk_bits, K, threadIdx.x, kBlock);
  for (int i = threadIdx.x; i < K * K1; i += kBlock) s_mem[i] = (i % K1) < K ? member[(i / K1) * K + (i % K1)] : 0;
  __syncthreads();
  const uint8_t* lab = SMEM_LAB ? s_lab : a.minor_label;
  const int lane = threadIdx.x & 31, g = lane & (G - 1);
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  const int grp = threadIdx.x / G;
  double* acc0 = s_acc0 + grp * K1;
  double* acc1 = s_acc1 + grp * K1;
  unsigned* cnt = s_cnt + grp * K1;
  const int total_groups = gridDim.x * groups_per_block;
  const int D = a.D, NCH = a.NCH, NSUB = a.NSUB, cb = a.chunk_bits;
  const uint32_t mmask = (a.minor_bits >= 32) ? 0xffffffffu : ((1u << a.minor_bits) - 1u);
  auto node_at = [&](uint32_t addr) -> uint32_t { return SMEM_LAB ? (uint32_t)lab[addr] : (uint32_t)__ldg(lab + addr); };
  for (int wi = blockIdx.x * groups_per_block + grp; wi < a.n_work; wi += total_groups) {
    const int line = a.list ? a.list[wi] : wi;
    const uint32_t* so = a.sub_off + (size_t)line * NSUB;
    for (int r = g; r < K1; r += G) { acc0[r] = 0.0; acc1[r] = 0.0; cnt[r] = 0; }
    __syncwarp(gmask);
    for (int d = 0; d < D; ++d) {
      for (int ch = 0; ch < NCH; ++ch) {
        const int64_t pb = (int64_t)so[d * NCH + ch] * PH_GE, pe = (int64_t)so[d * NCH + ch + 1] * PH_GE;
        for (int64_t p = pb + g; p < pe; p += G) {
          const uint32_t idx = a.half[p];
          if (idx >= PH_PAD) atomicAdd(&cnt[node_at(((uint32_t)ch << cb) | idx)], 1u);  // idx < PH_PAD = padding
        }
      }
      __syncwarp(gmask);
      const double l0 = a.L0[d], l1 = a.L1[d];
      for (int r = g; r < K1; r += G) {
        const double c = (double)cnt[r];
        acc0[r] = fma(c, l0, acc0[r]);
        acc1[r] = fma(c, l1, acc1[r]);
        cnt[r] = 0;
      }
      __syncwarp(gmask);
    }
    if (g == 0) {  // rare codes: serial, in storage order (deterministic)
      for (int64_t p = a.tail_off[line]; p < a.tail_off[line + 1]; ++p) {
        const uint32_t w = a.tail_words[p];
        const uint32_t code = w >> a.minor_bits;
        const uint32_t r = node_at(w & mmask);
        acc0[r] += a.L0[code];
        acc1[r] += a.L1[code];
      }
    }
    __syncwarp(gmask);
    for (int k = g; k < K; k += G) {
      double t = 0.0;
      const uint8_t* mrow = s_mem + k * K1;
      for (int r = 0; r < K1; ++r) t += mrow[r] ? acc1[r] : acc0[r];
      T[(size_t)wi * K + k] = t;
    }
    __syncwarp(gmask);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------

// shared memory of one k_pass CTA with `threads` threads; 0 = does not fit
size_t pass_smem_bytes(const ph_handle* h, const PhOrient* o, int mode, int NC, int K, bool smem_lab, int threads, int B,
                       int row_words, int* lut_in_smem, long long image_bytes = -1, bool b4 = false) {
  size_t b = 0;
  if (smem_lab) b += image_bytes >= 0 ? (size_t)image_bytes : (size_t)o->lab_bytes;
  if (mode == MODE_MAPPED) b += (((size_t)(K + 2) * (K + 2)) + 15) & ~(size_t)15;
  const size_t warps = (size_t)(threads / 32) * (warp_words(B, o->NSUB, o->D, NC, b4) + row_words) * 4 + 16;
  const size_t lut = ((size_t)h->n_codes * 17 + 15) & ~(size_t)15;
  *lut_in_smem = (b + warps + lut <= (size_t)h->max_smem && h->n_codes <= 4096) ? 1 : 0;
  if (*lut_in_smem) b += lut;
  b += warps;
  return b <= (size_t)h->max_smem ? b : 0;
}

template <int MODE, int NC>
int launch_pass(ph_handle* h, const PhOrient* o, PassArgs& a, cudaStream_t st) {
  if (a.n_work <= 0) return PH_OK;
  // lines per warp batch: 8; 32 and 4 are knobs (ph_set_group).  Pushed records hold at most 8 lines.
  int B = o->batch == 32 ? (a.push ? 8 : 32) : (o->batch == 4 ? 4 : 8);
  if (o->batch == 0) {
    // automatic: between one and two batches of 8 per warp of the grid the work quantises badly (some warps stream two
    // batches, most one: 6250 batches on 4736 warps when 4 GPUs share C3, 71 -> 63 us) -- halve the batch there.  With
    // fewer batches than warps, or many per warp, 8 is faster (measured, scripts/probe_snv_shard.py).
    const int64_t nb8 = ((int64_t)a.n_work + 7) / 8, warps = (int64_t)h->sm_count * 32;
    B = (nb8 > warps && nb8 < 2 * warps) ? 4 : 8;
  }
  if (MODE == MODE_COUNTS) a.row_words = B * NC * a.ncp / 2;  // ncp is a multiple of 8: whole 16-byte units
  int lut = 0, threads = 0;
  bool smem_lab = true, aligned = false;
  size_t smem = 0;
  // first choice: the label image at a 2^16-aligned shared address (LAY = 2).  Costs up to 64 KB of address space below
  // the image (the tables live there), so it needs  64 KB + image + tables  to fit -- true for up to ~150 k minor indices.
  // With it, whole-orientation passes (no line list, no peer push) run on the batched-4 copy of the stream (LAY = 3).
  // Whole-orientation passes (no line list) run on the batched-4 copy of the stream when the store keeps one: LAY = 3
  // with the aligned image, LAY = 4 (image wherever it fits) otherwise.
  bool use4 = false, al4 = false;
  const bool can4 = MODE != MODE_COUNTS && o->half4 && !a.list && h->batch4 && o->chunk_bits == 16;
  int n_phases = 1, cpp = o->NCH;   // phased pass: chunks per phase (see PassArgs)
  size_t tab4 = 0;
  // batched-4 items of 2 or 4 consecutive batches (k_pass NB8) where every warp of the grid still gets several items:
  // the per-item code runs once per 16 / 24 / 32 lines instead of once per 8.  Not in phases.  (Pushed results: the LL
  // slots of an item are B consecutive ones, ph_peer.cu sizes the windows for the largest item.)
  int B4items = 8;
  if (can4 && h->item_batches != 1) {
    // rounds of items per warp x (batches per item + the per-item share, about a quarter of a batch's streaming time):
    // C3 by SNV, 25 000 batches on 4 736 warps: 6 x 1.25 / 3 x 2.25 / 2 x 3.25 / 2 x 4.25 (measured 159 / 150 / - / 169 us).
    // Only sizes whose per-warp tables (and the likelihood table) still fit next to the label image.
    const int64_t warps = (int64_t)h->sm_count * 32;
    double best = 0.0;
    for (int nb8 = 1; nb8 <= 4; ++nb8) {
      if (h->item_batches >= 2 && h->item_batches <= 4 && nb8 != h->item_batches) continue;  // forced (PH_ITEM_BATCHES)
      if (nb8 > 1) {
        int l8 = 0, ln = 0;
        pass_smem_bytes(h, o, MODE, NC, a.K, true, 1024, 8, a.row_words, &l8, -1, true);
        if (!pass_smem_bytes(h, o, MODE, NC, a.K, true, 1024, 8 * nb8, a.row_words, &ln, -1, true) || ln < l8) continue;
      }
      const int64_t items = ((int64_t)o->n_batches4 + nb8 - 1) / nb8;
      const double cost = (double)((items + warps - 1) / warps) * (nb8 + 0.25);
      if (best == 0.0 || cost < best) {
        best = cost;
        B4items = 8 * nb8;
      }
    }
  }
  if (o->chunk_bits == 16 && (h->aligned_image || can4)) {
    int Bt = can4 ? B4items : B;
    size_t full = pass_smem_bytes(h, o, MODE, NC, a.K, true, 1024, Bt, a.row_words, &lut, -1, can4);
    while (can4 && !full && Bt > 8) {  // the tables of the larger item do not fit next to the image: smaller items
      Bt -= 8;
      full = pass_smem_bytes(h, o, MODE, NC, a.K, true, 1024, Bt, a.row_words, &lut, -1, true);
    }
    if (can4) B4items = Bt;
    const bool fits_al = h->aligned_image && full && full + 65536 <= (size_t)h->max_smem;
    if (fits_al) {
      smem = full + 65536;
      threads = 1024;
      aligned = true;
      a.tab_bytes = (int)(full - (size_t)o->lab_bytes);
    }
    const int force_cpp = h->phase_chunks > 0 && h->phase_chunks < o->NCH ? h->phase_chunks : 0;  // test knob PH_PHASE_CHUNKS
    if (can4 && ((!full && !fits_al) || force_cpp)) {
      // the image of the whole minor axis does not fit: walk it in phases of `cpp` chunks.  Aligned image (2 chunks of
      // 64 KB above a 64 KB boundary) when that takes no more phases than the 3 chunks an unaligned image holds.
      int lut0 = 0;
      tab4 = pass_smem_bytes(h, o, MODE, NC, a.K, true, 1024, 8, a.row_words, &lut0, 0, true);
      if (tab4 && tab4 + 65536 <= (size_t)h->max_smem) {
        const int cpp_u = (int)(((size_t)h->max_smem - tab4) >> 16);
        const int cpp_a = h->aligned_image && (size_t)h->max_smem >= tab4 + 65536 + 65536 ? (int)(((size_t)h->max_smem - 65536 - tab4) >> 16) : 0;
        const int ph_u = (o->NCH + cpp_u - 1) / cpp_u;
        const int ph_a = cpp_a > 0 ? (o->NCH + cpp_a - 1) / cpp_a : 1 << 30;
        al4 = ph_a <= ph_u;
        cpp = al4 ? cpp_a : cpp_u;
        if (force_cpp && force_cpp <= cpp) cpp = force_cpp;
        n_phases = (o->NCH + cpp - 1) / cpp;
        lut = lut0;
        if (al4) a.tab_bytes = (int)tab4;
      }
    }
    if (can4 && (fits_al || full || n_phases > 1)) {
      if (n_phases == 1) {
        if (!fits_al) smem = full;
        al4 = fits_al;
      }
      threads = 1024;
      use4 = true;
      aligned = al4;
      B = n_phases > 1 ? 8 : B4items;
      a.list = o->perm4;              // the lines in shape order, 8 per batch, -1 = empty slot
      a.n_work = o->n_batches4 * 8;
      a.half4 = o->half4;
      a.bstep_off = o->bstep_off;
      // split tail (PassArgs): the last `split_tail` x (warps of the grid) batches go out in 2^split_sh parts
      a.n_items = a.split_from = o->n_batches4;
      a.split_sh = 0;
      if (n_phases == 1 && B == 8 && o->part_acc && h->split_sh > 0 && h->split_tail > 0) {
        const int64_t n_split = std::min<int64_t>(o->n_batches4, (int64_t)h->sm_count * 32 * h->split_tail / 100);
        a.split_from = (int)(o->n_batches4 - n_split);
        a.split_sh = h->split_sh;
        a.n_items = (int)(a.split_from + (n_split << h->split_sh));
        a.part_acc = o->part_acc;
        a.part_tick = o->part_tick;
      }
      if (n_phases > 1) smem = 1;     // set per phase below
    }
  }
  for (int pass = 0; pass < 2 && !smem; ++pass) {
    smem_lab = pass == 0;
    for (threads = 1024; threads >= 256 && !smem; threads >>= 1) smem = pass_smem_bytes(h, o, MODE, NC, a.K, smem_lab, threads, B, a.row_words, &lut);
    if (smem) threads <<= 1;
  }
  if (!smem) PH_FAIL(PH_ERR_UNSUPPORTED, "pass kernel tables do not fit in shared memory (K=%d, %d sub-segments)", a.K, o->NSUB);
  a.lut_in_smem = lut;
  if (!smem_lab) {
    // the label image goes to global memory (gathered through L1/L2)
    if ((size_t)o->lab_bytes > h->lab_img_bytes) PH_FAIL(PH_ERR_UNSUPPORTED, "label image scratch too small");
    k_label_image<MODE == MODE_MAPPED, NC><<<h->sm_count, 256, 0, st>>>(h->d_lab_img, a.minor_label, o->n_minor, o->NCH,
                                                                        o->chunk_bits, a.K);
    ph_count_launch();
    a.minor_label = h->d_lab_img;
  }
  // one CTA of up to 32 warps per SM; PH_CTAS=2 halves the CTA instead when two copies of the image fit
  int ctas = 1;
  if (h->ctas_per_sm == 2 && threads == 1024 && !aligned && !use4) {
    int lut2 = 0;
    const size_t smem2 = pass_smem_bytes(h, o, MODE, NC, a.K, smem_lab, 512, B, a.row_words, &lut2);
    if (smem2 && lut2 == lut && 2 * (smem2 + 1024) <= (size_t)h->smem_per_sm) {
      ctas = 2;
      threads = 512;
      smem = smem2;
    }
  }
  const int grid = h->sm_count * ctas;
  a.counter = h->d_counter;  // zero at creation, re-armed by the last CTA of every launch
  if (!a.done) a.done = h->d_counter + 1;
#define PH_LAUNCH(BB, SL, SS)                                                                                    \
  do {                                                                                                           \
    auto kern = k_pass<BB, MODE, NC, SL, SS>;                                                                    \
    static int attr_dev = -1; /* opt-in shared memory: once per instantiation and device */                     \
    if (attr_dev != h->device) {                                                                                 \
      PH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));        \
      attr_dev = h->device;                                                                                      \
    }                                                                                                            \
    kern<<<grid, threads, smem, st>>>(a);                                                                        \
  } while (0)
#define PH_LAUNCH_B(BB)                                                        \
  do {                                                                         \
    if (!smem_lab) PH_LAUNCH(BB, false, 0);                                    \
    else if (aligned) PH_LAUNCH(BB, true, 2);                                  \
    else if (o->chunk_bits == 16) PH_LAUNCH(BB, true, 1);                      \
    else PH_LAUNCH(BB, true, 0);                                               \
  } while (0)
  if (a.push && (int64_t)((a.n_work + B - 1) / B) * B > (int64_t)a.ll.cap_slots)
    PH_FAIL(PH_ERR_UNSUPPORTED, "%d result slots exceed the %d reserved per rank in the peer window", ((a.n_work + B - 1) / B) * B, a.ll.cap_slots);
  if (use4) {
    if constexpr (MODE != MODE_COUNTS) {
      if (n_phases > 1) {
        const size_t need = (size_t)o->n_major * o->D * PH_MAX_FAST_CLUSTERS * sizeof(uint32_t);
        if (need > h->partial_bytes) {
          cudaFree(h->d_partial);
          h->d_partial = nullptr;
          h->partial_bytes = 0;
          PH_CUDA(cudaMalloc(&h->d_partial, need));
          h->partial_bytes = need;
        }
        a.partial = h->d_partial;
        a.phased = 1;
      }
      const int push = a.push;
      for (int p = 0; p < n_phases; ++p) {
        if (n_phases > 1) {
          a.ch_lo = p * cpp;
          a.ch_n = o->NCH - a.ch_lo < cpp ? o->NCH - a.ch_lo : cpp;
          a.phase_first = p == 0;
          a.phase_last = p == n_phases - 1;
          a.push = a.phase_last ? push : 0;   // results leave with the last phase only
          // image of this phase's chunks: whole 64 KB chunks, the last chunk of the axis as long as its labels
          const int CH = (1 << o->chunk_bits) - PH_PAD;
          const int last = o->n_minor - (o->NCH - 1) * CH;
          const size_t tail_b = (a.ch_lo + a.ch_n == o->NCH) ? (size_t)((PH_PAD + last + 15) & ~15) : (size_t)65536;
          a.lab_bytes = (int)(((size_t)(a.ch_n - 1) << 16) + tail_b);
          smem = tab4 + (size_t)a.lab_bytes + (al4 ? 65536 : 0);
        }
        if (n_phases > 1) {
          if (al4) PH_LAUNCH(8, true, 5);
          else PH_LAUNCH(8, true, 6);
        } else if (B == 32) {
          if (al4) PH_LAUNCH(32, true, 3);
          else PH_LAUNCH(32, true, 4);
        } else if (B == 24) {
          if (al4) PH_LAUNCH(24, true, 3);
          else PH_LAUNCH(24, true, 4);
        } else if (B == 16) {
          if (al4) PH_LAUNCH(16, true, 3);
          else PH_LAUNCH(16, true, 4);
        } else if (al4) PH_LAUNCH(8, true, 3);
        else PH_LAUNCH(8, true, 4);
        if (p + 1 < n_phases) ph_count_launch();
      }
      a.push = push;
    }
  } else if (B == 32) PH_LAUNCH_B(32);
  else if (B == 4) PH_LAUNCH_B(4);
  else PH_LAUNCH_B(8);
#undef PH_LAUNCH_B
#undef PH_LAUNCH
  ph_count_launch();
  PH_CUDA(cudaGetLastError());
  const_cast<PhOrient*>(o)->labels_in_smem = smem_lab ? 1 : 0;
  return PH_OK;
}

template <int MODE>
int launch_table(ph_handle* h, const PhOrient* o, PassArgs& a, cudaStream_t st) {
  if (a.C == 1) return launch_pass<MODE, 1>(h, o, a, st);
  if (a.C == 2) return launch_pass<MODE, 2>(h, o, a, st);
  return launch_pass<MODE, 3>(h, o, a, st);
}

void fill_common(const ph_handle* h, const PhOrient* o, PassArgs& a) {
  memset(&a, 0, sizeof(a));
  a.half = o->half;
  a.sub_off = o->sub_off;
  a.tail_words = o->tail_words;
  a.tail_off = o->tail_off;
  a.D = o->D;
  a.NCH = o->NCH;
  a.NSUB = o->NSUB;
  a.chunk_bits = o->chunk_bits;
  a.minor_bits = o->minor_bits;
  a.n_minor = o->n_minor;
  a.lab_bytes = o->lab_bytes;
  a.L0 = h->d_L0;
  a.L1 = h->d_L1;
  a.varpos = h->d_varpos;
  a.n_codes = h->n_codes;
}

// host-pointer helpers: carve the pinned staging buffer
struct Pin {
  uint8_t* p;
  size_t left;
  void* take(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes > left) return nullptr;
    void* r = p;
    p += bytes;
    left -= bytes;
    return r;
  }
};

int check_labels(const uint8_t* lab, int n, int maxv, const char* what) {
  uint8_t mx = 0;  // branch-free maximum first (vectorises); the scan for the offender only runs on failure
  for (int i = 0; i < n; ++i) mx = lab[i] > mx ? lab[i] : mx;
  if (mx <= maxv) return PH_OK;
  for (int i = 0; i < n; ++i)
    if (lab[i] > maxv) PH_FAIL(PH_ERR_ARG, "%s[%d] = %d exceeds %d", what, i, lab[i], maxv);
  return PH_OK;
}

// page-locked host memory can be the source / target of the async copies directly (no staging memcpy)
bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

// Results of a batched-4 pass land in device scratch in LINE order while the kernel walks the lines in shape order
// (scattered 1- to 8-byte stores: fine in L2, slow as individual PCIe writes).  This kernel then streams the scratch
// arrays into the page-locked host buffers with coalesced stores -- one launch instead of one D2H copy per array.
struct CopySegs {
  const uint8_t* src[4];
  uint8_t* dst[4];
  size_t bytes[4];
  int n;
};
__global__ void k_copy_out(const CopySegs c) {
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
  for (int s = 0; s < c.n; ++s) {
    const size_t words = c.bytes[s] >> 2;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(c.src[s]);
    uint32_t* dst = reinterpret_cast<uint32_t*>(c.dst[s]);
    for (size_t i = tid; i < words; i += nt) dst[i] = src[i];
    for (size_t i = (words << 2) + tid; i < c.bytes[s]; i += nt) c.dst[s][i] = c.src[s][i];  // last 0..3 bytes
  }
}
int copy_out(const ph_handle* h, CopySegs& c, cudaStream_t st) {
  size_t total = 0;
  for (int s = 0; s < c.n; ++s) {
    total += c.bytes[s];
    if ((reinterpret_cast<uintptr_t>(c.src[s]) | reinterpret_cast<uintptr_t>(c.dst[s])) & 3) PH_FAIL(PH_ERR_ARG, "unaligned result buffer");
  }
  if (!total) return PH_OK;
  const size_t want = (total / 4 + 255) / 256;
  const int blocks = (int)(want < (size_t)h->sm_count * 4 ? (want ? want : 1) : (size_t)h->sm_count * 4);
  k_copy_out<<<blocks, 256, 0, st>>>(c);
  ph_count_launch();
  PH_CUDA(cudaGetLastError());
  return PH_OK;
}

// device-side alias of a page-locked host buffer (UVA: the same address when the allocation is mapped), else null
template <typename T>
T* device_view(T* host) {
  cudaPointerAttributes at;
  if (!host || cudaPointerGetAttributes(&at, host) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return at.type == cudaMemoryTypeHost ? static_cast<T*>(at.devicePointer) : nullptr;
}

// common driver for the per-line "fast" passes (TABLE / LINEAR / BRANCHING)
int run_fast(ph_handle* h, PhOrient* o, int mode, const uint8_t* minor_label, int C, double lenA, const int32_t* list,
             int32_t n_list, double* S0, double* S1, int32_t* Nobs, int32_t* Nvar, uint8_t* label_out,
             int32_t* nobs_out, int on_device, void* stream) {
  if (!h) PH_FAIL(PH_ERR_ARG, "null handle");
  if (!minor_label) PH_FAIL(PH_ERR_ARG, "label array is NULL");
  if (C < 1 || C > PH_MAX_FAST_CLUSTERS) PH_FAIL(PH_ERR_ARG, "C=%d outside 1..%d", C, PH_MAX_FAST_CLUSTERS);
  if (!list) n_list = o->n_major;
  if (n_list < 0 || n_list > o->n_major) PH_FAIL(PH_ERR_ARG, "n_list=%d outside 0..%d", n_list, o->n_major);
  PH_CUDA(cudaSetDevice(h->device));
  PassArgs a;
  fill_common(h, o, a);
  a.C = C;
  a.lenA = lenA;
  a.n_work = n_list;
  const int nobs_per = mode == MODE_BRANCHING ? 2 : 1;
  if (on_device) {
    cudaStream_t st = (cudaStream_t)stream;
    a.minor_label = minor_label;
    a.list = list;
    a.S0 = S0; a.S1 = S1; a.Nobs = Nobs; a.Nvar = Nvar; a.label_out = label_out; a.nobs_out = nobs_out;
    if (mode == MODE_TABLE) return launch_table<MODE_TABLE>(h, o, a, st);
    if (!label_out || !nobs_out) PH_FAIL(PH_ERR_ARG, "output pointer is NULL");
    if (mode == MODE_LINEAR) return launch_pass<MODE_LINEAR, 1>(h, o, a, st);
    return launch_pass<MODE_BRANCHING, 2>(h, o, a, st);
  }
  // host pointers: validate, stage through pinned memory, run on the private stream, copy back, sync
  int rc = check_labels(minor_label, o->n_minor, C, "label");
  if (rc != PH_OK) return rc;
  if (list)
    for (int i = 0; i < n_list; ++i)
      if (list[i] < 0 || list[i] >= o->n_major) PH_FAIL(PH_ERR_ARG, "list[%d]=%d out of range", i, list[i]);
  cudaStream_t st = h->stream;
  Pin pin{h->h_stage, h->h_stage_bytes};
  const uint8_t* p_lab = minor_label;
  if (!is_pinned(minor_label)) {
    uint8_t* q = (uint8_t*)pin.take(o->n_minor);
    memcpy(q, minor_label, o->n_minor);
    p_lab = q;
  }
  PH_CUDA(cudaMemcpyAsync(h->d_lab_minor, p_lab, o->n_minor, cudaMemcpyHostToDevice, st));
  a.minor_label = h->d_lab_minor;
  if (list) {
    int32_t* p_list = (int32_t*)pin.take((size_t)n_list * 4);
    memcpy(p_list, list, (size_t)n_list * 4);
    PH_CUDA(cudaMemcpyAsync(h->d_list, p_list, (size_t)n_list * 4, cudaMemcpyHostToDevice, st));
    a.list = h->d_list;
  }
  const size_t nt = (size_t)n_list * C;
  double* dS0 = h->d_f64;
  double* dS1 = h->d_f64 + (size_t)o->n_major * PH_MAX_FAST_CLUSTERS;
  int32_t* dNo = h->d_i32;
  int32_t* dNv = h->d_i32 + (size_t)o->n_major * PH_MAX_FAST_CLUSTERS;
  // Results go to page-locked host memory.  Zero-copy (default): the holders' stores of the kernel epilogue target the
  // mapped host buffers themselves -- posted PCIe writes that overlap the pass -- so no device-to-host copy follows the
  // kernel (C3 linear step: 1 MB of results, ~27 us of copies after a 195 us kernel).  Fallback: device scratch + D2H.
  if (mode == MODE_TABLE) {
    double* pS0 = S0 ? (double*)pin.take(nt * 8) : nullptr;
    double* pS1 = S1 ? (double*)pin.take(nt * 8) : nullptr;
    int32_t* pNo = Nobs ? (int32_t*)pin.take(nt * 4) : nullptr;
    int32_t* pNv = Nvar ? (int32_t*)pin.take(nt * 4) : nullptr;
    double *vS0 = device_view(pS0), *vS1 = device_view(pS1);
    int32_t *vNo = device_view(pNo), *vNv = device_view(pNv);
    const bool zc = h->zero_copy_out && (!pS0 || vS0) && (!pS1 || vS1) && (!pNo || vNo) && (!pNv || vNv);
    // a whole-orientation pass may run on the batched-4 copy, whose results come in shape order: scratch + k_copy_out
    const bool b4 = !list && o->half4 && h->batch4;
    const bool via_scratch = zc && b4;   // (the staging buffer carves are 256-byte aligned)
    const bool zc_direct = zc && !b4;
    if (zc_direct) {
      a.S0 = vS0; a.S1 = vS1; a.Nobs = vNo; a.Nvar = vNv;
    } else {
      a.S0 = S0 ? dS0 : nullptr; a.S1 = S1 ? dS1 : nullptr; a.Nobs = Nobs ? dNo : nullptr; a.Nvar = Nvar ? dNv : nullptr;
    }
    rc = launch_table<MODE_TABLE>(h, o, a, st);
    if (rc != PH_OK) return rc;
    if (via_scratch) {
      CopySegs c{};
      auto add = [&](const void* src, void* dst, size_t bytes) {
        if (!dst) return;
        c.src[c.n] = (const uint8_t*)src; c.dst[c.n] = (uint8_t*)dst; c.bytes[c.n] = bytes; ++c.n;
      };
      add(dS0, vS0, nt * 8); add(dS1, vS1, nt * 8); add(dNo, vNo, nt * 4); add(dNv, vNv, nt * 4);
      rc = copy_out(h, c, st);
      if (rc != PH_OK) return rc;
    } else if (!zc_direct) {
      if (pS0) PH_CUDA(cudaMemcpyAsync(pS0, dS0, nt * 8, cudaMemcpyDeviceToHost, st));
      if (pS1) PH_CUDA(cudaMemcpyAsync(pS1, dS1, nt * 8, cudaMemcpyDeviceToHost, st));
      if (pNo) PH_CUDA(cudaMemcpyAsync(pNo, dNo, nt * 4, cudaMemcpyDeviceToHost, st));
      if (pNv) PH_CUDA(cudaMemcpyAsync(pNv, dNv, nt * 4, cudaMemcpyDeviceToHost, st));
    }
    PH_CUDA(cudaStreamSynchronize(st));
    if (pS0) memcpy(S0, pS0, nt * 8);
    if (pS1) memcpy(S1, pS1, nt * 8);
    if (pNo) memcpy(Nobs, pNo, nt * 4);
    if (pNv) memcpy(Nvar, pNv, nt * 4);
    return PH_OK;
  }
  if (!label_out || !nobs_out) PH_FAIL(PH_ERR_ARG, "output pointer is NULL");
  const bool direct = is_pinned(label_out) && is_pinned(nobs_out);  // caller's buffers are page-locked: no staging memcpy
  uint8_t* pL = direct ? label_out : (uint8_t*)pin.take(n_list);
  int32_t* pN = direct ? nobs_out : (int32_t*)pin.take((size_t)n_list * 4 * nobs_per);
  uint8_t* vL = device_view(pL);
  int32_t* vN = device_view(pN);
  const bool zc = h->zero_copy_out && vL && vN;
  // a whole-orientation pass may run on the batched-4 copy, whose results come in shape order (scattered stores):
  // they go to device scratch, then k_copy_out streams them to the host buffers (or plain D2H copies if unaligned)
  const bool b4 = !list && o->half4 && h->batch4;
  const bool via_scratch = zc && b4 && (((reinterpret_cast<uintptr_t>(vL) | reinterpret_cast<uintptr_t>(vN)) & 3) == 0);
  const bool zc_direct = zc && !b4;
  a.label_out = zc_direct ? vL : h->d_u8;
  a.nobs_out = zc_direct ? vN : dNo;
  rc = mode == MODE_LINEAR ? launch_pass<MODE_LINEAR, 1>(h, o, a, st) : launch_pass<MODE_BRANCHING, 2>(h, o, a, st);
  if (rc != PH_OK) return rc;
  if (via_scratch) {
    CopySegs c{};
    c.n = 2;
    c.src[0] = h->d_u8; c.dst[0] = vL; c.bytes[0] = (size_t)n_list;
    c.src[1] = (const uint8_t*)dNo; c.dst[1] = (uint8_t*)vN; c.bytes[1] = (size_t)n_list * 4 * nobs_per;
    rc = copy_out(h, c, st);
    if (rc != PH_OK) return rc;
  } else if (!zc_direct) {
    PH_CUDA(cudaMemcpyAsync(pL, h->d_u8, n_list, cudaMemcpyDeviceToHost, st));
    PH_CUDA(cudaMemcpyAsync(pN, dNo, (size_t)n_list * 4 * nobs_per, cudaMemcpyDeviceToHost, st));
  }
  PH_CUDA(cudaStreamSynchronize(st));
  if (!direct) {
    memcpy(label_out, pL, n_list);
    memcpy(nobs_out, pN, (size_t)n_list * 4 * nobs_per);
  }
  return PH_OK;
}

}  // namespace

extern "C" int ph_snv_table(ph_handle* h, const uint8_t* cell_label, int32_t C, const int32_t* snv_list, int32_t n_list,
                            double* S0, double* S1, int32_t* Nobs, int32_t* Nvar, int on_device, void* stream) {
  if (!h) PH_FAIL(PH_ERR_ARG, "null handle");
  return run_fast(h, &h->by_snv, MODE_TABLE, cell_label, C, 1.0, snv_list, n_list, S0, S1, Nobs, Nvar, nullptr, nullptr,
                  on_device, stream);
}

extern "C" int ph_cell_table(ph_handle* h, const uint8_t* snv_label, int32_t Q, const int32_t* cell_list, int32_t n_list,
                             double* A0, double* A1, int32_t* Nobs, int32_t* Nvar, int on_device, void* stream) {
  if (!h) PH_FAIL(PH_ERR_ARG, "null handle");
  return run_fast(h, &h->by_cell, MODE_TABLE, snv_label, Q, 1.0, cell_list, n_list, A0, A1, Nobs, Nvar, nullptr, nullptr,
                  on_device, stream);
}

extern "C" int ph_assign_linear(ph_handle* h, const uint8_t* cell_label, double lenA, const int32_t* snv_list,
                                int32_t n_list, uint8_t* label_out, int32_t* nobs_out, int on_device, void* stream) {
  if (!h) PH_FAIL(PH_ERR_ARG, "null handle");
  return run_fast(h, &h->by_snv, MODE_LINEAR, cell_label, 1, lenA, snv_list, n_list, nullptr, nullptr, nullptr, nullptr,
                  label_out, nobs_out, on_device, stream);
}

extern "C" int ph_assign_branching(ph_handle* h, const uint8_t* cell_label, const int32_t* snv_list, int32_t n_list,
                                   uint8_t* label_out, int32_t* nobs_out, int on_device, void* stream) {
  if (!h) PH_FAIL(PH_ERR_ARG, "null handle");
  return run_fast(h, &h->by_snv, MODE_BRANCHING, cell_label, 2, 1.0, snv_list, n_list, nullptr, nullptr, nullptr, nullptr,
                  label_out, nobs_out, on_device, stream);
}

extern "C" int ph_block_sums(ph_handle* h, const uint8_t* cell_label, int32_t C, const uint8_t* snv_label, int32_t Q,
                             double* sum0, double* sum1, int64_t* nobs, int64_t* nvar, int on_device, void* stream) {
  if (!h) PH_FAIL(PH_ERR_ARG, "null handle");
  if (!cell_label || !snv_label || !sum0 || !sum1 || !nobs || !nvar) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (C < 1 || C > PH_MAX_FAST_CLUSTERS || Q < 1 || Q > 16) PH_FAIL(PH_ERR_ARG, "C=%d / Q=%d out of range", C, Q);
  PH_CUDA(cudaSetDevice(h->device));
  // Q <= 3: the per-CELL table over the SNV clusters (cell-major copy) reduced by cell cluster -- a cell's row is
  // computed wholly where the cell lives, so a cell-sharded store assembles the very same table and reduces it to the
  // very same bits (ph_line_reduce).  Q > 3: per-SNV table over the cell clusters, reduced by SNV cluster.
  const bool by_cell = Q <= PH_MAX_FAST_CLUSTERS;
  PhOrient* o = by_cell ? &h->by_cell : &h->by_snv;
  cudaStream_t st = on_device ? (cudaStream_t)stream : h->stream;
  const uint8_t* d_cl = cell_label;
  const uint8_t* d_sl = snv_label;
  Pin pin{h->h_stage, h->h_stage_bytes};
  if (!on_device) {
    int rc = check_labels(cell_label, h->n_cells, by_cell ? 255 : C, "cell_label");
    if (rc != PH_OK) return rc;
    if (by_cell) {
      rc = check_labels(snv_label, h->n_snvs, Q, "snv_label");
      if (rc != PH_OK) return rc;
    }
    uint8_t* p1 = (uint8_t*)pin.take(h->n_cells);
    uint8_t* p2 = (uint8_t*)pin.take(h->n_snvs);
    memcpy(p1, cell_label, h->n_cells);
    memcpy(p2, snv_label, h->n_snvs);
    PH_CUDA(cudaMemcpyAsync(h->d_lab_minor, p1, h->n_cells, cudaMemcpyHostToDevice, st));
    PH_CUDA(cudaMemcpyAsync(h->d_u8, p2, h->n_snvs, cudaMemcpyHostToDevice, st));
    d_cl = h->d_lab_minor;
    d_sl = h->d_u8;
  }
  PassArgs a;
  fill_common(h, o, a);
  a.C = by_cell ? Q : C;
  a.n_work = o->n_major;
  a.minor_label = by_cell ? d_sl : d_cl;
  double* dS0 = h->d_f64;
  double* dS1 = h->d_f64 + (size_t)o->n_major * PH_MAX_FAST_CLUSTERS;
  int32_t* dNo = h->d_i32;
  int32_t* dNv = h->d_i32 + (size_t)o->n_major * PH_MAX_FAST_CLUSTERS;
  a.S0 = dS0; a.S1 = dS1; a.Nobs = dNo; a.Nvar = dNv;
  int rc = launch_table<MODE_TABLE>(h, o, a, st);
  if (rc != PH_OK) return rc;
  const int nb = Q * C;
  double* d0 = on_device ? sum0 : h->d_small;
  double* d1 = on_device ? sum1 : h->d_small + 64;
  long long* dn = on_device ? (long long*)nobs : (long long*)(h->d_small + 128);
  long long* dv = on_device ? (long long*)nvar : (long long*)(h->d_small + 192);
  if (by_cell)  // table columns = SNV clusters, lines labelled by cell cluster: out[(q-1) * C + (c-1)]
    k_line_reduce<<<nb, 1024, 0, st>>>(dS0, dS1, dNo, dNv, d_cl, o->n_major, Q, C, d0, d1, dn, dv);
  else
    k_block_reduce<<<nb, 1024, 0, st>>>(dS0, dS1, dNo, dNv, d_sl, o->n_major, C, d0, d1, dn, dv);
  ph_count_launch();
  PH_CUDA(cudaGetLastError());
  if (!on_device) {
    double* p = (double*)pin.take(256 * 8);
    PH_CUDA(cudaMemcpyAsync(p, h->d_small, 256 * 8, cudaMemcpyDeviceToHost, st));
    PH_CUDA(cudaStreamSynchronize(st));
    memcpy(sum0, p, nb * 8);
    memcpy(sum1, p + 64, nb * 8);
    memcpy(nobs, p + 128, nb * 8);
    memcpy(nvar, p + 192, nb * 8);
  }
  return PH_OK;
}

// The reductions of ph_block_sums / ph_tree_loglik on caller-provided per-line tables (host pointers): what a sharded
// store runs on the table it assembled from all ranks -- same kernels, same order, same bits as one GPU.
extern "C" int ph_line_reduce(ph_handle* h, const double* S0, const double* S1, const int32_t* Nobs, const int32_t* Nvar,
                              const uint8_t* line_label, int32_t n_lines, int32_t ncol, int32_t nlab, double* sum0,
                              double* sum1, int64_t* nobs, int64_t* nvar) {
  if (!h) PH_FAIL(PH_ERR_ARG, "null handle");
  if (!S0 || !S1 || !Nobs || !Nvar || !line_label || !sum0 || !sum1 || !nobs || !nvar) PH_FAIL(PH_ERR_ARG, "null pointer");
  const int nmax = h->n_cells > h->n_snvs ? h->n_cells : h->n_snvs;
  if (n_lines < 1 || n_lines > nmax || ncol < 1 || ncol > PH_MAX_FAST_CLUSTERS || nlab < 1 || ncol * nlab > 64)
    PH_FAIL(PH_ERR_ARG, "bad shape n_lines=%d ncol=%d nlab=%d", n_lines, ncol, nlab);
  PH_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const size_t nt = (size_t)n_lines * ncol;
  double* dS0 = h->d_f64;
  double* dS1 = h->d_f64 + (size_t)nmax * PH_MAX_FAST_CLUSTERS;
  int32_t* dNo = h->d_i32;
  int32_t* dNv = h->d_i32 + (size_t)nmax * PH_MAX_FAST_CLUSTERS;
  PH_CUDA(cudaMemcpyAsync(dS0, S0, nt * 8, cudaMemcpyHostToDevice, st));
  PH_CUDA(cudaMemcpyAsync(dS1, S1, nt * 8, cudaMemcpyHostToDevice, st));
  PH_CUDA(cudaMemcpyAsync(dNo, Nobs, nt * 4, cudaMemcpyHostToDevice, st));
  PH_CUDA(cudaMemcpyAsync(dNv, Nvar, nt * 4, cudaMemcpyHostToDevice, st));
  PH_CUDA(cudaMemcpyAsync(h->d_lab_major, line_label, (size_t)n_lines, cudaMemcpyHostToDevice, st));
  const int nb = ncol * nlab;
  k_line_reduce<<<nb, 1024, 0, st>>>(dS0, dS1, dNo, dNv, h->d_lab_major, n_lines, ncol, nlab, h->d_small, h->d_small + 64,
                                     (long long*)(h->d_small + 128), (long long*)(h->d_small + 192));
  ph_count_launch();
  PH_CUDA(cudaGetLastError());
  double host[256];
  PH_CUDA(cudaMemcpyAsync(host, h->d_small, sizeof(host), cudaMemcpyDeviceToHost, st));
  PH_CUDA(cudaStreamSynchronize(st));
  memcpy(sum0, host, nb * 8);
  memcpy(sum1, host + 64, nb * 8);
  memcpy(nobs, host + 128, nb * 8);
  memcpy(nvar, host + 192, nb * 8);
  return PH_OK;
}

extern "C" int ph_node_reduce(ph_handle* h, const double* per_line, const uint8_t* line_node, int32_t n_lines, int32_t K,
                              double* per_node) {
  if (!h) PH_FAIL(PH_ERR_ARG, "null handle");
  if (!per_line || !line_node || !per_node) PH_FAIL(PH_ERR_ARG, "null pointer");
  const int nmax = h->n_cells > h->n_snvs ? h->n_cells : h->n_snvs;
  if (n_lines < 1 || n_lines > nmax || K < 1 || K > PH_MAX_TREE_NODES) PH_FAIL(PH_ERR_ARG, "bad shape n_lines=%d K=%d", n_lines, K);
  PH_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  PH_CUDA(cudaMemcpyAsync(h->d_f64, per_line, (size_t)n_lines * 8, cudaMemcpyHostToDevice, st));
  PH_CUDA(cudaMemcpyAsync(h->d_lab_major, line_node, (size_t)n_lines, cudaMemcpyHostToDevice, st));
  k_node_reduce<<<K, 1024, 0, st>>>(h->d_f64, h->d_lab_major, n_lines, h->d_small);
  ph_count_launch();
  PH_CUDA(cudaGetLastError());
  double host[PH_MAX_TREE_NODES];
  PH_CUDA(cudaMemcpyAsync(host, h->d_small, (size_t)K * 8, cudaMemcpyDeviceToHost, st));
  PH_CUDA(cudaStreamSynchronize(st));
  memcpy(per_node, host, (size_t)K * 8);
  return PH_OK;
}

extern "C" int ph_tree_loglik(ph_handle* h, const uint8_t* cell_node, const uint8_t* snv_node, const uint8_t* present,
                              int32_t K, double* per_node, double* per_cell, int on_device, void* stream) {
  if (!h) PH_FAIL(PH_ERR_ARG, "null handle");
  if (!cell_node || !snv_node || !present || !per_node) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (K < 1 || K > PH_MAX_TREE_NODES) PH_FAIL(PH_ERR_ARG, "K=%d outside 1..%d", K, PH_MAX_TREE_NODES);
  PH_CUDA(cudaSetDevice(h->device));
  PhOrient* o = &h->by_cell;
  cudaStream_t st = on_device ? (cudaStream_t)stream : h->stream;
  Pin pin{h->h_stage, h->h_stage_bytes};
  const int K1 = K + 1;
  const uint8_t *d_cn = cell_node, *d_sn = snv_node;
  // (K+1)x(K+1) class map: row = node of the cell, column = node of the SNV (K = not in tree -> absent)
  uint8_t* p_map = (uint8_t*)pin.take((size_t)K1 * K1);
  if (on_device) {
    uint8_t tmp[65 * 65];
    PH_CUDA(cudaMemcpyAsync(tmp, present, (size_t)K * K, cudaMemcpyDeviceToHost, st));
    PH_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < K1; ++k)
      for (int q = 0; q < K1; ++q) p_map[k * K1 + q] = (k < K && q < K) ? tmp[k * K + q] : 0;
  } else {
    for (int k = 0; k < K1; ++k)
      for (int q = 0; q < K1; ++q) p_map[k * K1 + q] = (k < K && q < K) ? present[k * K + q] : 0;
    uint8_t* p1 = (uint8_t*)pin.take(h->n_cells);
    uint8_t* p2 = (uint8_t*)pin.take(h->n_snvs);
    memcpy(p1, cell_node, h->n_cells);
    memcpy(p2, snv_node, h->n_snvs);
    PH_CUDA(cudaMemcpyAsync(h->d_lab_major, p1, h->n_cells, cudaMemcpyHostToDevice, st));
    PH_CUDA(cudaMemcpyAsync(h->d_lab_minor, p2, h->n_snvs, cudaMemcpyHostToDevice, st));
    d_cn = h->d_lab_major;
    d_sn = h->d_lab_minor;
  }
  PH_CUDA(cudaMemcpyAsync(h->d_map, p_map, (size_t)K1 * K1, cudaMemcpyHostToDevice, st));
  PassArgs a;
  fill_common(h, o, a);
  a.C = 2;
  a.K = K;
  a.n_work = o->n_major;
  a.minor_label = d_sn;
  a.major_label = d_cn;
  a.map = h->d_map;
  double* d_pc = (on_device && per_cell) ? per_cell : h->d_f64;
  a.per_line = d_pc;
  int rc = launch_pass<MODE_MAPPED, 2>(h, o, a, st);
  if (rc != PH_OK) return rc;
  double* d_pn = on_device ? per_node : h->d_small;
  k_node_reduce<<<K, 1024, 0, st>>>(d_pc, d_cn, h->n_cells, d_pn);
  ph_count_launch();
  PH_CUDA(cudaGetLastError());
  if (!on_device) {
    double* p = (double*)pin.take((size_t)K * 8);
    double* pc = per_cell ? (double*)pin.take((size_t)h->n_cells * 8) : nullptr;
    PH_CUDA(cudaMemcpyAsync(p, d_pn, (size_t)K * 8, cudaMemcpyDeviceToHost, st));
    if (pc) PH_CUDA(cudaMemcpyAsync(pc, d_pc, (size_t)h->n_cells * 8, cudaMemcpyDeviceToHost, st));
    PH_CUDA(cudaStreamSynchronize(st));
    memcpy(per_node, p, (size_t)K * 8);
    if (pc) memcpy(per_cell, pc, (size_t)h->n_cells * 8);
  }
  return PH_OK;
}

namespace {
int run_node_table(ph_handle* h, PhOrient* o, const uint8_t* minor_node, const uint8_t* member, int32_t K,
                   const int32_t* list, int32_t n_list, double* T, int on_device, void* stream) {
  if (!minor_node || !member || !T) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (K < 1 || K > PH_MAX_TREE_NODES) PH_FAIL(PH_ERR_ARG, "K=%d outside 1..%d", K, PH_MAX_TREE_NODES);
  if (!list) n_list = o->n_major;
  if (n_list < 0 || n_list > o->n_major) PH_FAIL(PH_ERR_ARG, "n_list=%d out of range", n_list);
  if (n_list == 0) return PH_OK;
  PH_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = on_device ? (cudaStream_t)stream : h->stream;
  Pin pin{h->h_stage, h->h_stage_bytes};
  PassArgs a;
  fill_common(h, o, a);
  a.K = K;
  a.n_work = n_list;
  const uint8_t* d_member = member;
  double* dT = T;
  if (on_device) {
    a.minor_label = minor_node;
    a.list = list;
  } else {
    uint8_t* p1 = (uint8_t*)pin.take(o->n_minor);
    memcpy(p1, minor_node, o->n_minor);
    PH_CUDA(cudaMemcpyAsync(h->d_lab_minor, p1, o->n_minor, cudaMemcpyHostToDevice, st));
    a.minor_label = h->d_lab_minor;
    uint8_t* pm = (uint8_t*)pin.take((size_t)K * K);
    memcpy(pm, member, (size_t)K * K);
    PH_CUDA(cudaMemcpyAsync(h->d_map, pm, (size_t)K * K, cudaMemcpyHostToDevice, st));
    d_member = h->d_map;
    if (list) {
      for (int i = 0; i < n_list; ++i)
        if (list[i] < 0 || list[i] >= o->n_major) PH_FAIL(PH_ERR_ARG, "list[%d]=%d out of range", i, list[i]);
      int32_t* pl = (int32_t*)pin.take((size_t)n_list * 4);
      memcpy(pl, list, (size_t)n_list * 4);
      PH_CUDA(cudaMemcpyAsync(h->d_list, pl, (size_t)n_list * 4, cudaMemcpyHostToDevice, st));
      a.list = h->d_list;
    }
    if ((size_t)n_list * K > (size_t)(o->n_major > o->n_minor ? o->n_major : o->n_minor) * 2 * PH_MAX_FAST_CLUSTERS)
      PH_FAIL(PH_ERR_UNSUPPORTED, "node table of %d x %d exceeds the staging scratch; pass fewer lines per call", n_list, K);
    dT = h->d_f64;
  }
  const int G = 8;
  const int gpb = kBlock / G;
  const int K1 = K + 1;
  size_t extra = (size_t)gpb * K1 * (8 + 8 + 4) + (size_t)K * K1 + 16;
  size_t lab = (size_t)o->lab_bytes;
  bool smem_lab = lab + extra <= (size_t)h->max_smem;
  size_t smem = (smem_lab ? lab : 0) + extra;
  if (!smem_lab) {
    if (lab > h->lab_img_bytes) PH_FAIL(PH_ERR_UNSUPPORTED, "label image scratch too small");
    k_label_image<true, 1><<<h->sm_count, 256, 0, st>>>(h->d_lab_img, a.minor_label, o->n_minor, o->NCH, o->chunk_bits, K);
    ph_count_launch();
    a.minor_label = h->d_lab_img;
  }
  if (smem_lab) {
    auto kern = k_node_table<8, true>;
    PH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    kern<<<h->sm_count, kBlock, smem, st>>>(a, d_member, dT);
  } else {
    auto kern = k_node_table<8, false>;
    PH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    kern<<<h->sm_count, kBlock, smem, st>>>(a, d_member, dT);
  }
  ph_count_launch();
  PH_CUDA(cudaGetLastError());
  if (!on_device) {
    const size_t nb = (size_t)n_list * K * 8;
    double* p = (double*)pin.take(nb);
    if (!p) PH_FAIL(PH_ERR_UNSUPPORTED, "node table too large for the pinned staging buffer");
    PH_CUDA(cudaMemcpyAsync(p, dT, nb, cudaMemcpyDeviceToHost, st));
    PH_CUDA(cudaStreamSynchronize(st));
    memcpy(T, p, nb);
  }
  return PH_OK;
}
}  // namespace

extern "C" int ph_snv_node_table(ph_handle* h, const uint8_t* cell_node, const uint8_t* member, int32_t K,
                                 const int32_t* snv_list, int32_t n_list, double* T, int on_device, void* stream) {
  if (!h) PH_FAIL(PH_ERR_ARG, "null handle");
  return run_node_table(h, &h->by_snv, cell_node, member, K, snv_list, n_list, T, on_device, stream);
}

extern "C" int ph_cell_node_table(ph_handle* h, const uint8_t* snv_node, const uint8_t* member, int32_t K,
                                  const int32_t* cell_list, int32_t n_list, double* T, int on_device, void* stream) {
  if (!h) PH_FAIL(PH_ERR_ARG, "null handle");
  return run_node_table(h, &h->by_cell, snv_node, member, K, cell_list, n_list, T, on_device, stream);
}

// ------------------------------------------------------------------------------------------------
// cell-sharded (multi-GPU) K1: per-rank partial tables packed for ONE all-reduce, then the K1a epilogue on the sum
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void k_finish_linear(const double* __restrict__ t, int k, double lenA, uint8_t* __restrict__ label,
                                int32_t* __restrict__ nobs) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  const double m0 = t[j] / lenA, m1 = t[(size_t)k + j] / lenA;  // linear_split.py:229-232
  label[j] = (m0 >= m1) ? 2 : 1;
  nobs[j] = (int32_t)t[2 * (size_t)k + j];
}
__global__ void k_finish_branching(const double* __restrict__ t, int k, uint8_t* __restrict__ label,
                                   int32_t* __restrict__ nobs) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  const size_t kc = 2 * (size_t)k;  // packed [3][k][2]
  const double s0a = t[2 * (size_t)j], s0b = t[2 * (size_t)j + 1];
  const double s1a = t[kc + 2 * (size_t)j], s1b = t[kc + 2 * (size_t)j + 1];
  const double cA = s1a + s0b, cB = s0a + s1b, cC = s1a + s1b;  // branching_split.py:332-346
  const double mx = fmax(cA, fmax(cB, cC));
  label[j] = (cA == mx) ? 1 : ((cB == mx) ? 2 : 3);
  nobs[2 * (size_t)j] = (int32_t)t[2 * kc + 2 * (size_t)j];
  nobs[2 * (size_t)j + 1] = (int32_t)t[2 * kc + 2 * (size_t)j + 1];
}
}  // namespace

extern "C" int ph_snv_partial(ph_handle* h, const uint8_t* cell_label, int32_t C, const int32_t* snv_list, int32_t n_list,
                              double* packed, int on_device, void* stream) {
  if (!h) PH_FAIL(PH_ERR_ARG, "null handle");
  if (!on_device) PH_FAIL(PH_ERR_UNSUPPORTED, "ph_snv_partial works on device buffers only (it feeds an all-reduce)");
  if (!cell_label || !packed) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (C < 1 || C > PH_MAX_FAST_CLUSTERS) PH_FAIL(PH_ERR_ARG, "C=%d outside 1..%d", C, PH_MAX_FAST_CLUSTERS);
  PhOrient* o = &h->by_snv;
  if (!snv_list) n_list = o->n_major;
  if (n_list < 0 || n_list > o->n_major) PH_FAIL(PH_ERR_ARG, "n_list=%d out of range", n_list);
  PH_CUDA(cudaSetDevice(h->device));
  PassArgs a;
  fill_common(h, o, a);
  a.C = C;
  a.n_work = n_list;
  a.minor_label = cell_label;
  a.list = snv_list;
  const size_t kc = (size_t)n_list * C;
  a.S0 = packed;
  a.S1 = packed + kc;
  a.NobsF = packed + 2 * kc;
  return launch_table<MODE_TABLE>(h, o, a, (cudaStream_t)stream);
}

extern "C" int ph_finish_linear(const double* packed, int32_t n_list, double lenA, uint8_t* label_out, int32_t* nobs_out,
                                void* stream) {
  if (!packed || !label_out || !nobs_out) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (n_list <= 0) return PH_OK;
  k_finish_linear<<<(n_list + 255) / 256, 256, 0, (cudaStream_t)stream>>>(packed, n_list, lenA, label_out, nobs_out);
  ph_count_launch();
  PH_CUDA(cudaGetLastError());
  return PH_OK;
}

extern "C" int ph_finish_branching(const double* packed, int32_t n_list, uint8_t* label_out, int32_t* nobs_out, void* stream) {
  if (!packed || !label_out || !nobs_out) PH_FAIL(PH_ERR_ARG, "null pointer");
  if (n_list <= 0) return PH_OK;
  k_finish_branching<<<(n_list + 255) / 256, 256, 0, (cudaStream_t)stream>>>(packed, n_list, label_out, nobs_out);
  ph_count_launch();
  PH_CUDA(cudaGetLastError());
  return PH_OK;
}

// ------------------------------------------------------------------------------------------------
// cell-sharded K1 over peer memory (ph_peer.cu): the COUNTS pass
// ------------------------------------------------------------------------------------------------
int ph_pass_counts(ph_handle* h, const uint8_t* cell_label, int C, const int32_t* snv_list, int32_t n_list,
                   uint16_t* const* cnt_dst, uint32_t* const* flag_dst, unsigned* done, int blk, int ncp, int world,
                   const uint32_t* epoch_ptr, cudaStream_t st) {
  PhOrient* o = &h->by_snv;
  PassArgs a;
  fill_common(h, o, a);
  a.C = C;
  a.n_work = n_list;
  a.minor_label = cell_label;
  a.list = snv_list;
  for (int r = 0; r < world; ++r) {
    a.cnt_dst[r] = cnt_dst[r];
    a.flag_dst[r] = flag_dst[r];
  }
  a.done = done;
  a.blk = blk;
  a.ncp = ncp;
  a.world = world;
  a.epoch_ptr = epoch_ptr;
  return C == 1 ? launch_pass<MODE_COUNTS, 1>(h, o, a, st) : launch_pass<MODE_COUNTS, 2>(h, o, a, st);
}

// SNV-sharded K1 + K1a (ph_peer.cu): this rank's SNVs, results pushed as LL words to every rank's window and collected
// by the same launch
int ph_pass_assign_push(ph_handle* h, int mode, const uint8_t* cell_label, double lenA, const PhLL& ll, unsigned* done,
                        int out_off, int world, uint32_t* epoch_ptr, cudaStream_t st) {
  PhOrient* o = &h->by_snv;
  PassArgs a;
  fill_common(h, o, a);
  a.C = mode == MODE_LINEAR ? 1 : 2;
  a.lenA = lenA;
  a.n_work = o->n_major;
  a.minor_label = cell_label;
  a.ll = ll;
  a.done = done;
  a.world = world;
  a.epoch_ptr = epoch_ptr;
  a.epoch_w = epoch_ptr;
  a.push = 1;
  a.out_off = out_off;
  return mode == MODE_LINEAR ? launch_pass<MODE_LINEAR, 1>(h, o, a, st) : launch_pass<MODE_BRANCHING, 2>(h, o, a, st);
}
