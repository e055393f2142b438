// engine.cuh -- internal declarations shared by the translation units of libodmsspe_b200.so.
// Nothing here is part of the ABI (that is include/od_msspe_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/od_msspe_b200.h"

#define MSSPE_CUDA_TRY(ctx, expr)                                                                       \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      (ctx)->set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);     \
      return MSSPE_ERR_CUDA;                                                                            \
    }                                                                                                   \
  } while (0)

// One direction's inverted index + greedy state, all device-resident.
struct DirIndex {
  uint64_t n_records = 0;     // R: valid (segment, k-mer) records
  uint64_t n_codes = 0;       // D: distinct words
  uint64_t* codes = nullptr;  // [D] ascending 2-bit codes
  uint32_t* post_off = nullptr;  // [D+1] CSR offsets into postings
  uint32_t* postings = nullptr;  // [R] segment ids, ascending inside each list
  uint32_t* fwd_ids = nullptr;   // [G*s] code id per (segment, slot), 0xFFFFFFFF if none (forward index)
  uint32_t* list_part = nullptr; // [D] partition_no shared by ALL postings of the list, or >= 2^31 when they span several
  // greedy state
  uint32_t* freq = nullptr;      // [D] live count per code
  unsigned long long* acc = nullptr;  // [D] arrival-counter|partial-sum for lists that span count tiles
  uint32_t* ignored = nullptr;   // [ceil(G/32)] covered-segment bitmask
  uint32_t* cov = nullptr;       // [65536] partition_coverage
  uint32_t* pmark = nullptr;     // [2048] scratch partition bitmap
  struct SelectCtl* ctl = nullptr;  // device control block
  msspe_candidate* out = nullptr;   // [max_iterations] winners, device
  uint32_t out_capacity = 0;
  uint32_t* tile_first = nullptr;   // [n_tiles] first code whose postings reach into each count tile
  uint32_t n_tiles = 0;
  // scoring stream of the persistent greedy kernel (built on first use): the lists with >= 2 postings in code order,
  // followed by the postings of the single-posting lists.  A k-mer with one posting can never be selected (freq == 1
  // stops the loop before the push, main.rs:354-360), so its posting only has to be counted, not reduced per list.
  bool stream_built = false;
  uint32_t* s_postings = nullptr;   // [R]   lists region [0, s_list_post), tail [s_list_post, R)
  uint32_t* s_off = nullptr;        // [s_lists + 1]
  uint4* s_id = nullptr;            // [s_lists] per stream list: code id, list_part, its range in the complete CSR
  uint32_t* s_tile_first = nullptr; // [s_tiles]
  uint32_t* s_cost = nullptr;       // [tiles + 2] exclusive prefix of the tile costs (range balancing)
  uint32_t s_lists = 0, s_list_post = 0, s_tiles = 0;
  // partition view of the index for MSSPE_SELECT_PARTITIONED (select_part.cu; built on first use): the codes grouped by
  // the partition all their postings share ("units"), lists that span several partitions kept apart
  bool pv_built = false, pv_dist = false;   // pv_dist: built for the multi-GPU loop (words of other ranks count as multi-partition)
  uint32_t pv_units = 0, pv_single = 0, pv_multi_n = 0;
  uint64_t pv_multi_postings = 0;
  uint32_t* pv_ucode_off = nullptr;  // [U+1] CSR over pv_ucodes
  uint32_t* pv_ucodes = nullptr;     // [D] code ids: single-partition lists grouped by unit (ascending inside), then the multi-partition lists
  uint32_t* pv_fwdl = nullptr;       // [G*s] forward index renumbered: position in pv_ucodes | 0x80000000 for multi lists, 0xFFFFFFFF none
  uint32_t* pv_useg_off = nullptr;   // [U+1] CSR over pv_usegs
  uint32_t* pv_usegs = nullptr;      // [G] segment ids grouped by partition (ascending inside)
};

constexpr int MSSPE_CNT_THREADS = 512;  // threads per block of the K3 kernels
constexpr int MSSPE_CNT_TILE = 512;     // postings per warp tile of the coverage scoring (32 lanes x 4 x uint4)

// Device control block of the greedy loop (one per direction).
struct SelectCtl {
  unsigned int gmax;               // max live frequency of this iteration
  unsigned int n_tied;
  unsigned long long best_key;     // (f32 score bits << 32) | ~code_id
  unsigned int done;               // 1 = loop finished
  unsigned int n_out;              // winners pushed
  unsigned int iterations;         // find_most_freq_kmer calls made
  unsigned int pad;
  unsigned long long evals;        // sum of live records counted (reference inner-loop executions)
  unsigned long long postings_read;
  // persistent kernel: per-iteration slots, double-buffered by iteration parity
  unsigned int pg[2];              // max live frequency
  unsigned int pt[2];              // number of tied k-mers
  unsigned long long pk[2];        // best (score bits << 32 | ~code id)
  unsigned int plive[2];           // live postings counted in this iteration (all blocks)
  unsigned int resume_it;          // persistent kernel left for a stream compaction: iteration to resume at
  unsigned int exit_compact;       // 1 = it left because less than half of the streamed postings were still live
  unsigned long long t_count_ns;   // time inside the coverage-scoring phases (block 0, %globaltimer)
  unsigned long long t_tie_ns;     // time inside the tie-break phases
  unsigned long long t_total_ns;
  unsigned long long t_dbg[8];     // fine-grained phase timers of block 0 (diagnostic)
  unsigned long long t_fine[24];   // per-step clocks, only written by a -DMSSPE_FINE_TIMERS build (diagnostic)
};

// Kernel classes of the per-kernel profile (msspe_get_kernel_profile): launches and algorithmic bytes are always counted,
// device time only while msspe_set_profiling is on (two CUDA events around every launch, on the launching stream).
enum {
  KP_ENCODE = 0, KP_SORT_HIST, KP_SORT_SCATTER, KP_SCAN, KP_CSR, KP_VIEW, KP_GREEDY_UNIT, KP_GREEDY_MERGE, KP_GREEDY_VERIFY,
  KP_GREEDY_WHOLE, KP_THERMO, KP_DIMER, KP_N
};
struct KProfClass { double ms = 0.0; uint32_t launches = 0; uint64_t bytes = 0; };
struct KProfPending { int cls; cudaEvent_t e0, e1; };

struct msspe_ctx {
  KProfClass kprof[KP_N];
  std::vector<KProfPending> kprof_pending;
  std::vector<cudaEvent_t> kprof_pool;
  msspe_config cfg{};
  int device = 0;
  int sm_count = 148;
  size_t smem_optin = 48 * 1024;
  cudaStream_t stream = nullptr;      // main stream (own or caller's)
  cudaStream_t stream2 = nullptr;     // second direction
  std::thread* reserve_thread = nullptr;   // msspe_reserve_pool: background first touch of the pool
  void* h_stage = nullptr; size_t h_stage_bytes = 0;   // pinned staging for small results (grown on demand)
  bool own_stream = true;
  cudaEvent_t ev[8] = {};
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::string err;
  // genomes
  uint32_t n_records = 0;
  std::vector<uint64_t> h_offsets;    // [n+1]
  std::vector<uint64_t> h_seg_base;   // [n+1] first segment index of each record
  uint8_t* d_bases = nullptr;         // device bases (owned unless borrowed)
  bool bases_borrowed = false;
  uint64_t bases_bytes = 0;
  uint64_t* d_offsets = nullptr;      // [n+1]
  uint64_t* d_seg_base = nullptr;     // [n+1]
  uint64_t n_segments = 0;            // G
  uint32_t slots = 0;                 // s = w - k + 1
  uint32_t max_partition = 0;
  uint32_t uniform_parts = 0;         // > 0: every record has exactly this many partitions (equal-length alignment)
  uint16_t* d_seg_part = nullptr;     // [G]
  uint32_t* d_seg_rec = nullptr;      // [G]
  bool loaded = false, built = false;
  DirIndex dir[2];
  // thermodynamic tables
  msspe_thal_raw_params raw{};
  bool raw_set = false;
  struct ThalDeviceTables* d_thal = nullptr;  // device copy of static tables (ntthal stand-ins: msspe_set_thal_params applies)
  struct ThalDeviceTables* d_thal_p3 = nullptr;  // Primer3's compiled-in tables (primer3_core stand-ins), never overridden
  msspe_thal_raw_params* raw_p3 = nullptr;
  struct ThalDimerConsts* d_p3_consts = nullptr;  // per-run constants of the primer3_core stand-ins (fixed conditions): built once
  struct ThalDimerConsts* h_p3_consts = nullptr;
  void* dist = nullptr;                        // select_dist.cu: NCCL communicator of the multi-GPU loop
  msspe_dimer_edge* xd_edges = nullptr;       // msspe_cross_dimer_device: lists of the last call (ctx-owned)
  uint64_t* xd_nostruct = nullptr;
  // pinned staging
  SelectCtl* h_ctl = nullptr;  // [2] pinned
  msspe_timing timing{};
  bool profiling = false;

  void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
  }
};

// ---- device scan / sort utilities (kmer_build.cu) ----
int msspe_exclusive_scan_u32(msspe_ctx* ctx, const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* d_total,
                             cudaStream_t st);

int msspe_radix_sort_pairs(msspe_ctx* ctx, uint64_t** key_a, uint32_t** val_a, uint64_t** key_b, uint32_t** val_b, uint64_t n,
                           uint32_t key_bits, cudaStream_t st);

// ---- stages ----
void msspe_join_reserve(msspe_ctx* c);
bool msspe_build_fast_applicable(const msspe_ctx* ctx);   // kmer_build_fast.cu
int msspe_build_fast(msspe_ctx* ctx);
int msspe_free_index(msspe_ctx* ctx);
int msspe_select_prepare_static(msspe_ctx* ctx, int dir, cudaStream_t st);  // select.cu: tile tables
int msspe_select_prepare_stream(msspe_ctx* ctx, int dir, cudaStream_t st);  // select.cu: scoring stream (lazy)
// select_part.cu: per-partition greedy sequences + merge (MSSPE_SELECT_PARTITIONED)
int msspe_select_partitioned(msspe_ctx* ctx, int ndirs, const int* dirs, uint32_t max_iter, uint32_t mms, msspe_candidate** outs,
                             uint32_t** n_outs);
int msspe_partition_view(msspe_ctx* ctx, int dir, cudaStream_t st);   // builds DirIndex::pv_* (lazy)
bool msspe_partitioned_applicable(msspe_ctx* ctx, int ndirs, const int* dirs, uint32_t max_iter);
void msspe_dist_free(msspe_ctx* ctx);                                  // select_dist.cu
int msspe_thal_upload_tables(msspe_ctx* ctx);
void msspe_thal_free_tables(msspe_ctx* ctx);

static inline uint64_t div_up_u64(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// ---- per-kernel profile (ctx.cu) ----
cudaEvent_t msspe_kprof_begin(msspe_ctx* c, cudaStream_t st);
void msspe_kprof_end(msspe_ctx* c, int cls, cudaEvent_t e0, cudaStream_t st, uint64_t alg_bytes);
// usage: { KPROF(c, KP_X, st, bytes) kernel<<<...>>>(...); }   -- the guard's destructor records the closing event
struct KProfGuard {
  msspe_ctx* c; int cls; cudaStream_t st; uint64_t bytes; cudaEvent_t e0;
  KProfGuard(msspe_ctx* c_, int cls_, cudaStream_t st_, uint64_t bytes_) : c(c_), cls(cls_), st(st_), bytes(bytes_), e0(msspe_kprof_begin(c_, st_)) {}
  ~KProfGuard() { msspe_kprof_end(c, cls, e0, st, bytes); }
};
#define KPROF(c, cls, st, bytes) KProfGuard _kp_guard_##__LINE__((c), (cls), (st), (uint64_t)(bytes));

// Device memory comes from the stream-ordered pool (cudaMallocAsync): no device-wide synchronisation per
// allocation, and freed blocks are cached by the pool (release threshold = unlimited, set in msspe_create).
static inline void msspe_dev_free(msspe_ctx* c, void* p) { if (p) cudaFreeAsync(p, c->stream); }
