// select.cu -- K3: the device-resident greedy set-cover loop.  Replaces find_candidates_kmers
// (od-msspe/src/main.rs:331-406), find_most_freq_kmer (:285-329) and partition_tie_score (:261-283).
//
// Four implementations of the same loop live here; all return identical winners, frequencies, tie counts, f32 scores
// and reference-equivalent evals (every parity test runs several of them):
//   greedy_persistent_kernel    MSSPE_SELECT_RECOUNT: the reference's algorithm -- every live (segment, k-mer) record is
//                               examined in every iteration (main.rs:292-309) -- as ONE cooperative launch for both
//                               directions: phase A recount of a "scoring stream" against a shared-memory bitmask,
//                               phase B arg-max with the partition-diversity tie score, two grid barriers per iteration;
//                               the host compacts the stream between launches when less than half of it is live.
//   greedy_incremental_kernel   MSSPE_SELECT_INCREMENTAL: counts kept exact through the forward index, histogram of
//                               counts for the maximum; no shared-memory bitmask, any number of segments.
//   count_kernel / tie_kernel / update_kernel (+ freq_max_kernel)   MSSPE_SELECT_BATCHED: the same phases as one launch
//                               each per iteration (per-launch profiling; count_kernel also serves msspe_shard_count).
// The building blocks of the recount and of the tie score are in select_device.cuh.  Also here: the scoring-stream
// build and compaction, msspe_coverage, and the per-rank primitives of the genome-sharded loop (msspe_shard_*).
#include <cooperative_groups.h>

#include <algorithm>

#include "engine.cuh"
#include "select_device.cuh"

namespace cg = cooperative_groups;

namespace {

using namespace msspe_sel;
constexpr int TIE_THREADS = 256, UPD_THREADS = 1024;

__global__ void tile_first_kernel(const uint32_t* __restrict__ post_off, uint32_t n_codes, uint32_t n_tiles,
                                  uint32_t* __restrict__ tile_first) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  const uint32_t start = t * (uint32_t)CNT_TILE;
  uint32_t lo = 0, hi = n_codes;  // post_off[lo] <= start < post_off[hi]
  while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (post_off[mid] <= start) lo = mid; else hi = mid; }
  tile_first[t] = lo;
}

// Cost of every tile of a scoring stream for the range balancing of the persistent kernel, in units of ~3 warp
// instructions: a list tile costs a fixed part (gather, scan, first round of list lanes) plus a share per list that
// starts in it (every 31 of them is another round); a tile of the counted-only tail is a gather and a popcount.
__global__ void tile_cost_kernel(const uint32_t* __restrict__ tile_first, uint32_t list_tiles, uint32_t n_lists, uint32_t all_tiles,
                                 uint32_t* __restrict__ cost) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > all_tiles) return;
  uint32_t c = 0u;  // cost[all_tiles] = 0: the exclusive scan then ends with the total
  if (t < list_tiles) c = 72u + ((t + 1u < list_tiles ? tile_first[t + 1u] : n_lists) - tile_first[t]);
  else if (t < all_tiles) c = 60u;
  cost[t] = c;
}

// ---- scoring stream (see DirIndex): lists with >= 2 postings in code order, single-posting lists as a tail ----
__global__ void stream_flag_kernel(const uint32_t* __restrict__ post_off, uint32_t n_codes, uint32_t* __restrict__ multi,
                                   uint32_t* __restrict__ mlen) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_codes) return;
  const uint32_t len = post_off[c + 1] - post_off[c];
  multi[c] = len >= 2u ? 1u : 0u;
  mlen[c] = len >= 2u ? len : 0u;
}

// One warp per 32 consecutive lists; short lists are copied by their lane, long ones by the whole warp.
__global__ void __launch_bounds__(256)
stream_copy_kernel(const uint32_t* __restrict__ post_off, const uint32_t* __restrict__ postings, uint32_t n_codes,
                   const uint32_t* __restrict__ rank, const uint32_t* __restrict__ moff, uint32_t n_lists, uint32_t n_list_post,
                   uint32_t tail_begin, const uint32_t* __restrict__ list_part,
                   uint32_t* __restrict__ s_postings, uint32_t* __restrict__ s_off, uint4* __restrict__ s_id) {
  const int lane = threadIdx.x & 31;
  const uint32_t c = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u + (uint32_t)lane;
  uint32_t a = 0, len = 0, dst = 0;
  if (c < n_codes) {
    a = post_off[c]; len = post_off[c + 1] - a;
    const uint32_t r = rank[c];
    if (len >= 2u) { dst = moff[c]; s_off[r] = dst; s_id[r] = make_uint4(c, list_part[c], a, a + len); }
    else if (len == 1u) dst = tail_begin + (c - r);  // singles before c = c - (lists with >= 2 postings before c)
    if (len <= 8u) for (uint32_t i = 0; i < len; i++) s_postings[dst + i] = postings[a + i];
  }
  if (c == 0) s_off[n_lists] = n_list_post;
  unsigned big = __ballot_sync(0xffffffffu, len > 8u);
  while (big) {
    const int l = __ffs(big) - 1;
    big &= big - 1;
    const uint32_t la = __shfl_sync(0xffffffffu, a, l), ll = __shfl_sync(0xffffffffu, len, l), ld = __shfl_sync(0xffffffffu, dst, l);
    for (uint32_t i = lane; i < ll; i += 32) s_postings[ld + i] = postings[la + i];
  }
}

// ---- dead-posting compaction of the scoring stream (between launches of the persistent kernel) ----
// live postings of every list of the current stream under the covered-segment bitmask; one warp per 32 lists
__global__ void __launch_bounds__(256)
stream_live_kernel(const uint32_t* __restrict__ s_off, const uint32_t* __restrict__ s_postings, uint32_t n_lists,
                   const uint32_t* __restrict__ ignored, uint32_t* __restrict__ cnt, uint32_t* __restrict__ multi,
                   uint32_t* __restrict__ mlen, uint32_t* __restrict__ single) {
  const int lane = threadIdx.x & 31;
  const uint32_t c = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u + (uint32_t)lane;
  uint32_t a = 0, len = 0, live = 0;
  if (c < n_lists) {
    a = s_off[c]; len = s_off[c + 1] - a;
    if (len <= 8u) for (uint32_t i = 0; i < len; i++) { const uint32_t g = s_postings[a + i]; live += (~ignored[g >> 5] >> (g & 31u)) & 1u; }
  }
  unsigned big = __ballot_sync(0xffffffffu, len > 8u);
  while (big) {
    const int l = __ffs(big) - 1;
    big &= big - 1;
    const uint32_t la = __shfl_sync(0xffffffffu, a, l), ll = __shfl_sync(0xffffffffu, len, l);
    uint32_t n = 0;
    for (uint32_t i = lane; i < ll; i += 32) { const uint32_t g = s_postings[la + i]; n += (~ignored[g >> 5] >> (g & 31u)) & 1u; }
    n = __reduce_add_sync(0xffffffffu, n);
    if (lane == l) live = n;
  }
  if (c < n_lists) { cnt[c] = live; multi[c] = live >= 2u ? 1u : 0u; mlen[c] = live >= 2u ? live : 0u; single[c] = live == 1u ? 1u : 0u; }
}

// order-preserving move of the live postings: lists that still have >= 2 of them keep a list (new rank = number of
// such lists before), a list left with one contributes it to the counted-only tail, dead lists vanish
__global__ void __launch_bounds__(256)
stream_compact_kernel(const uint32_t* __restrict__ s_off, const uint32_t* __restrict__ s_postings, const uint4* __restrict__ s_id,
                      uint32_t n_lists, const uint32_t* __restrict__ ignored, const uint32_t* __restrict__ cnt,
                      const uint32_t* __restrict__ rank, const uint32_t* __restrict__ moff, const uint32_t* __restrict__ srank,
                      uint32_t new_lists, uint32_t new_list_post, uint32_t new_tail_begin, uint32_t* __restrict__ n_postings,
                      uint32_t* __restrict__ n_off, uint4* __restrict__ n_id) {
  const int lane = threadIdx.x & 31;
  const uint32_t c = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u + (uint32_t)lane;
  uint32_t a = 0, len = 0, dst = 0, n = 0;
  if (c < n_lists) {
    a = s_off[c]; len = s_off[c + 1] - a; n = cnt[c];
    if (n >= 2u) { dst = moff[c]; n_off[rank[c]] = dst; n_id[rank[c]] = s_id[c]; }
    else if (n == 1u) dst = new_tail_begin + srank[c];
    if (n && len <= 8u)
      for (uint32_t i = 0; i < len; i++) { const uint32_t g = s_postings[a + i]; if ((~ignored[g >> 5] >> (g & 31u)) & 1u) n_postings[dst++] = g; }
  }
  if (c == 0) n_off[new_lists] = new_list_post;
  unsigned big = __ballot_sync(0xffffffffu, n && len > 8u);
  while (big) {
    const int l = __ffs(big) - 1;
    big &= big - 1;
    const uint32_t la = __shfl_sync(0xffffffffu, a, l), ll = __shfl_sync(0xffffffffu, len, l);
    uint32_t ld = __shfl_sync(0xffffffffu, dst, l);
    for (uint32_t base = 0; base < ll; base += 32) {
      const uint32_t i = base + lane;
      const uint32_t g = i < ll ? s_postings[la + i] : 0u;
      const bool live = i < ll && ((~ignored[g >> 5] >> (g & 31u)) & 1u);
      const unsigned m = __ballot_sync(0xffffffffu, live);
      if (live) n_postings[ld + __popc(m & ((1u << lane) - 1u))] = g;
      ld += __popc(m);
    }
  }
}

// live postings of the old tail, appended to the new tail in any order (the tail is only counted)
__global__ void stream_tail_compact_kernel(const uint32_t* __restrict__ s_postings, uint32_t tail_begin, uint32_t tail_end,
                                           const uint32_t* __restrict__ ignored, uint32_t* __restrict__ n_postings, uint32_t dst_begin,
                                           uint32_t* counter) {
  const uint32_t i = tail_begin + blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t g = i < tail_end ? s_postings[i] : 0u;
  const bool live = i < tail_end && ((~ignored[g >> 5] >> (g & 31u)) & 1u);
  const unsigned m = __ballot_sync(0xffffffffu, live);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (live) n_postings[dst_begin + base + __popc(m & ((1u << lane) - 1u))] = g;
}

template <bool SMEM_MASK>
__global__ void __launch_bounds__(CNT_THREADS)
count_kernel(const uint32_t* __restrict__ postings, const uint32_t* __restrict__ post_off,
             const uint32_t* __restrict__ tile_first, uint32_t n_codes, uint32_t n_post, uint32_t n_tiles,
             const uint32_t* __restrict__ ignored, uint32_t mask_words, uint32_t* __restrict__ freq,
             unsigned long long* __restrict__ acc, SelectCtl* ctl) {
  if (ld_volatile(&ctl->done)) return;
  extern __shared__ uint32_t smask[];
  __shared__ uint32_t s_max;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int WARPS = CNT_THREADS / 32;
  if (tid == 0) s_max = 0u;
  if (SMEM_MASK)
    for (uint32_t i = tid; i < mask_words; i += CNT_THREADS) smask[i] = ignored[i];
  __syncthreads();
  const uint32_t* mask = SMEM_MASK ? smask : ignored;
  uint32_t mymax = 0;
  CountJob J;
  J.postings = postings; J.post_off = post_off; J.n_codes = n_codes; J.n_post = n_post; J.mask = mask; J.freq = freq; J.acc = acc;
  count_job_range(J, n_tiles, blockIdx.x * WARPS + warp, gridDim.x * WARPS);
  J.tail_end = n_post;
  J.c_first = J.t_begin < J.t_end ? __ldg(tile_first + J.t_begin) : 0u;
  const unsigned long long live_total = warp_count_range<SMEM_MASK>(J, mymax, lane);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mymax = max(mymax, __shfl_xor_sync(0xffffffffu, mymax, o));
  if (lane == 0) { if (mymax) atomicMax(&s_max, mymax); if (live_total) atomicAdd(&ctl->evals, live_total); }
  __syncthreads();
  if (tid == 0 && s_max) atomicMax(&ctl->gmax, s_max);
}

// max over freq[] (incremental mode: freq is maintained by decrements, so only the max is needed)
__global__ void __launch_bounds__(256)
freq_max_kernel(const uint32_t* __restrict__ freq, uint32_t n_codes, SelectCtl* ctl) {
  if (ld_volatile(&ctl->done)) return;
  uint32_t m = 0;
  unsigned long long s = 0;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n_codes; c += gridDim.x * blockDim.x) {
    const uint32_t f = freq[c]; m = max(m, f); s += f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  if ((threadIdx.x & 31) == 0) { if (m) atomicMax(&ctl->gmax, m); if (s) atomicAdd(&ctl->evals, s); }
}

__global__ void __launch_bounds__(TIE_THREADS)
tie_kernel(const uint32_t* __restrict__ freq, uint32_t n_codes, const uint32_t* __restrict__ post_off,
           const uint32_t* __restrict__ postings, const uint32_t* __restrict__ ignored,
           const uint16_t* __restrict__ seg_part, const uint32_t* __restrict__ cov, SelectCtl* ctl, uint32_t p_words) {
  if (ld_volatile(&ctl->done)) return;
  const uint32_t gmax = ld_volatile(&ctl->gmax);
  if (gmax == 0) return;
  extern __shared__ uint32_t seen_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* seen = seen_all + (size_t)warp * p_words;
  const uint32_t n_chunks = (n_codes + 31u) / 32u;
  const uint32_t warps_total = gridDim.x * (TIE_THREADS / 32);
  for (uint32_t chunk = blockIdx.x * (TIE_THREADS / 32) + warp; chunk < n_chunks; chunk += warps_total) {
    const uint32_t c = chunk * 32u + lane;
    const bool tied = c < n_codes && freq[c] == gmax;
    unsigned m = __ballot_sync(0xffffffffu, tied);
    if (m == 0) continue;
    if (lane == 0) atomicAdd(&ctl->n_tied, (unsigned)__popc(m));
    while (m) {
      const int l = __ffs(m) - 1;
      m &= m - 1;
      const uint32_t cc = chunk * 32u + l;
      const float score = warp_tie_score<false, false>(cc, post_off, postings, ignored, seg_part, cov, seen, p_words, lane);
      if (lane == 0) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xFFFFFFFFu - cc);
        atomicMax(&ctl->best_key, key);
      }
    }
  }
}

template <bool INCREMENTAL>
__global__ void __launch_bounds__(UPD_THREADS)
update_kernel(const uint64_t* __restrict__ codes, const uint32_t* __restrict__ post_off,
              const uint32_t* __restrict__ postings, uint32_t* ignored, const uint16_t* __restrict__ seg_part,
              uint32_t* cov, uint32_t* pmark, SelectCtl* ctl, msspe_candidate* out, uint32_t max_iterations,
              uint32_t mms, const uint32_t* __restrict__ fwd_ids, uint32_t slots, uint32_t* freq) {
  if (ld_volatile(&ctl->done)) return;
  const uint32_t gmax = ld_volatile(&ctl->gmax);
  const unsigned long long key = *reinterpret_cast<volatile unsigned long long*>(&ctl->best_key);
  const uint32_t n_tied = ld_volatile(&ctl->n_tied);
  const uint32_t n_out = ld_volatile(&ctl->n_out);
  __syncthreads();
  const bool stop_before = gmax <= 1u;  // None (no live k-mer) or freq == 1: main.rs:353-366
  if (threadIdx.x == 0) {
    ctl->iterations += 1;
    if (stop_before) ctl->done = 1;
  }
  if (stop_before) return;
  const uint32_t c = 0xFFFFFFFFu - (uint32_t)key;
  const uint32_t a = post_off[c], b = post_off[c + 1];
  // main.rs:371-378: every posting (also already covered ones) is inserted; distinct partitions get +1
  for (uint32_t i = a + threadIdx.x; i < b; i += UPD_THREADS) {
    const uint32_t seg = postings[i];
    const uint32_t bit = 1u << (seg & 31u);
    const uint32_t oldw = atomicOr(&ignored[seg >> 5], bit);
    if (INCREMENTAL && !(oldw & bit)) {  // newly covered: its k-mers lose one live segment
      for (uint32_t q = 0; q < slots; q++) {
        const uint32_t id = fwd_ids[(uint64_t)seg * slots + q];
        if (id != 0xFFFFFFFFu) atomicSub(&freq[id], 1u);
      }
    }
    const uint32_t p = seg_part[seg];
    const uint32_t pbit = 1u << (p & 31u);
    const uint32_t old = atomicOr(&pmark[p >> 5], pbit);
    if (!(old & pbit)) atomicAdd(&cov[p], 1u);
  }
  __syncthreads();
  for (uint32_t i = a + threadIdx.x; i < b; i += UPD_THREADS) pmark[seg_part[postings[i]] >> 5] = 0u;
  if (threadIdx.x == 0) {
    msspe_candidate w;
    w.code = codes[c]; w.freq = gmax; w.n_tied = n_tied; w.tie_score = __uint_as_float((uint32_t)(key >> 32)); w.reserved = 0;
    out[n_out] = w;
    ctl->n_out = n_out + 1;
    if (gmax < mms || n_out + 1 >= max_iterations) ctl->done = 1;  // main.rs:387-390 and the for-loop bound :344
    ctl->gmax = 0; ctl->best_key = 0ull; ctl->n_tied = 0;
  }
}


// ------------------------------------------------------------------------------------------------------------
// The whole greedy loop as ONE persistent cooperative kernel (both directions at once): the covered-segment
// bitmask of each direction stays in shared memory for the lifetime of the kernel and every block applies the
// winner's postings to its own copy, so an iteration costs two grid barriers and no launch:
//   phase A  apply previous winner (smem mask; block 0 also the global mask + partition_coverage), then the
//            coverage scoring of this block's posting tiles -> freq[], atomicMax of the best frequency
//   barrier
//   phase B  stop tests; scan freq[] for ties, warp-per-k-mer partition score, atomicMax of (score, ~id)
//   barrier
//   winner   every block reads the winner; block 0 writes the output record
struct GreedyDir {
  const uint32_t* postings; const uint32_t* post_off; const uint32_t* tile_first; const uint64_t* codes;
  uint32_t n_codes, n_post, n_tiles, pad;
  const uint4* list_id;          // per list of the scoring stream: .x code id, .y the partition ALL its postings lie in or >= 2^31 (several),
                                 // .z/.w its range in the complete CSR (full_postings)
  uint32_t tail_t0, tail_t1, tail_end, stream_total;  // tiles / end of the counted-only tail; postings streamed per recount
  const uint32_t* full_off; const uint32_t* full_postings;  // the complete CSR: main.rs:371-378 walks ALL postings of a winner
  uint32_t done0, pad2;          // resumed launch: this direction had already finished
  const uint32_t* cost_prefix;   // [tail_t1 + 1] exclusive prefix of the tile costs, cost_prefix[tail_t1] = total
  uint32_t* ignored; uint32_t* freq; unsigned long long* acc; uint32_t* cov; SelectCtl* ctl; msspe_candidate* out;
};
struct GreedyArgs {
  GreedyDir d[2];
  int ndirs; uint32_t mask_words, p_words, max_iter, mms;
  uint32_t it0;       // first iteration of this launch (> 0: resumed after a stream compaction)
  uint32_t compact_min;  // leave for a compaction when < half of a stream of at least this many postings is live (0 = never)
  uint32_t n_part;   // max partition_no + 1
  uint32_t n_fp;     // entries of the block-cooperative tie-score scratch (n_part, or 0 = too many partitions)
  const uint16_t* seg_part;
  uint32_t uniform_parts;  // > 0: every record has this many partitions, partition_no = segment % uniform_parts (no table lookup)
  unsigned int* barrier;  // grid barrier arrival counter (zeroed before the launch)
};

// Grid-wide barrier for a cooperatively launched (co-resident) grid: one arrival counter, monotonically
// growing target.  __syncthreads + cumulative fence publish the block's writes; readers use L2-coherent loads.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    while (ld_volatile(counter) < target) { }
    __threadfence();
  }
  __syncthreads();
}
// cooperative_groups' own grid barrier: 1.21 us against 1.45 us for the counter above in isolation (296 x 512,
// tools/microbench/grid_barrier.cu).  Both persistent kernels use it for their two barriers per iteration (recount kernel,
// same-box A/B after the tie-score work: cfg2 7.90 vs 8.20 ms, cfg5/8 shard within 2 %); the counter above remains for
// the extra barrier of the global-bitmask variant.
__device__ __forceinline__ void grid_barrier_cg() { cg::this_grid().sync(); }

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Block 0's share of main.rs:371-378: EVERY posting of the winner (also the already covered ones, which the
// compacted stream no longer holds, hence the complete CSR) is marked in the global bitmask, and every distinct
// partition among them gets partition_coverage += 1.  pm = shared-memory partition bitmap, left zeroed.
template <int THREADS>
__device__ __forceinline__ void apply_winner_global(const GreedyDir& D, const uint16_t* __restrict__ seg_part, uint32_t uniform_parts, uint32_t list, uint32_t* pm) {
  const int tid = threadIdx.x;
  const uint4 e = __ldg(D.list_id + list);  // .z/.w: the list's range in the complete CSR
  const uint32_t a = e.z, b = e.w;
  for (uint32_t i = a + tid; i < b; i += THREADS) {
    const uint32_t sg = __ldg(D.full_postings + i);
    atomicOr(&D.ignored[sg >> 5], 1u << (sg & 31u));
    const uint32_t p = partition_of(seg_part, uniform_parts, sg);
    const uint32_t pbit = 1u << (p & 31u);
    const uint32_t old = atomicOr(&pm[p >> 5], pbit);
    if (!(old & pbit)) atomicAdd(&D.cov[p], 1u);
  }
  __syncthreads();
  for (uint32_t i = a + tid; i < b; i += THREADS) pm[partition_of(seg_part, uniform_parts, __ldg(D.full_postings + i)) >> 5] = 0u;
  __syncthreads();
}

// The same for one direction by HALF a block (threads [lt0, lt0 + THREADS/2)), without the block barriers: block 0
// applies the winners of both directions side by side (two chains of four dependent L2 round trips in parallel
// instead of one after the other).  The caller synchronises and calls clear_partition_marks afterwards.
template <int THREADS>
__device__ __forceinline__ void apply_winner_half(const GreedyDir& D, const uint16_t* __restrict__ seg_part, uint32_t uniform_parts, uint32_t list, uint32_t* pm, int lt) {
  const uint4 e = __ldg(D.list_id + list);
  const uint32_t a = e.z, b = e.w;
  for (uint32_t i = a + lt; i < b; i += THREADS / 2) {
    const uint32_t sg = __ldg(D.full_postings + i);
    atomicOr(&D.ignored[sg >> 5], 1u << (sg & 31u));
    const uint32_t p = partition_of(seg_part, uniform_parts, sg);
    const uint32_t pbit = 1u << (p & 31u);
    const uint32_t old = atomicOr(&pm[p >> 5], pbit);
    if (!(old & pbit)) atomicAdd(&D.cov[p], 1u);
  }
}

// Tied k-mers whose postings all lie in one partition (list_id[].y < 2^31): score = 0.0 + 1/(partition_coverage + 1),
// one thread each; scored entries of tied[] are flagged with bit 31.  (Tie-storm chunks; the usual case is inlined.)
__device__ __forceinline__ void single_partition_ties(const GreedyDir& D, uint32_t* tied, uint32_t nt, const uint32_t* covp, bool cov_shared, int par, int tid) {
  if ((uint32_t)tid >= nt) return;
  const uint32_t cc = tied[tid];
  const uint4 e = __ldg(D.list_id + cc);
  if (e.y >> 31) return;
  const uint32_t cv = cov_shared ? covp[e.y] : __ldcg(D.cov + e.y);
  const float sc1 = __fdiv_rn(1.0f, __fadd_rn(__uint2float_rn(cv), 1.0f));
  atomicMax(&D.ctl->pk[par], ((unsigned long long)__float_as_uint(sc1) << 32) | (unsigned long long)(0xFFFFFFFFu - cc));
  tied[tid] = cc | 0x80000000u;
}

// The recount job of one warp for direction d (a literal at every call site).
template <bool SMEM_MASK>
__device__ __forceinline__ void make_job(CountJob& J, const GreedyArgs& A, const int d, uint32_t* smask, uint32_t t_begin, uint32_t t_end, uint32_t c_first) {
  const GreedyDir& D = A.d[d];
  J.postings = D.postings; J.post_off = D.post_off; J.n_codes = D.n_codes; J.n_post = D.n_post;
  J.mask = SMEM_MASK ? smask + (size_t)d * A.mask_words : D.ignored; J.freq = D.freq; J.acc = D.acc;
  J.t_begin = t_begin; J.t_end = t_end; J.list_tiles = D.n_tiles; J.tail_end = D.tail_end;
  J.c_first = c_first;
}

// First tile of warp gw's range when the tiles [0, n) are split over nw warps by cost: the first t whose prefix
// reaches ceil(total * gw / nw).  Ranges of consecutive warps tile [0, n) without gaps; a range may be empty.
__device__ __forceinline__ uint32_t balanced_begin(const uint32_t* __restrict__ P, uint32_t n, uint32_t gw, uint32_t nw) {
  const unsigned long long total = P[n];
  const uint32_t target = (uint32_t)((total * gw + nw - 1u) / nw);
  uint32_t lo = 0, hi = n;  // answer in [lo, hi]; P[n] = total >= target
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(P + mid) >= target) hi = mid; else lo = mid + 1u; }
  return lo;
}

template <bool SMEM_MASK, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 1024 ? 1 : 2)
greedy_persistent_kernel(const GreedyArgs A) {
  extern __shared__ __align__(16) unsigned char dsm[];
  unsigned int bar_target = 0;
  __shared__ uint32_t s_tied[2 * THREADS];
  __shared__ uint32_t s_cnt2[2], s_sc[2], s_max[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int WARPS = THREADS / 32;
  // dynamic smem: [masks ndirs*mask_words] [pm 2*p_words] [seen p_words] [fp n_fp] [lst n_fp (u64)] [cov 2*n_fp]
  uint32_t* smask = reinterpret_cast<uint32_t*>(dsm);
  uint32_t* pm = smask + (SMEM_MASK ? (size_t)A.ndirs * A.mask_words : 0);
  uint32_t* seen = pm + 2u * A.p_words;
  uint32_t* fp = seen + A.p_words;
  unsigned long long* lst = reinterpret_cast<unsigned long long*>(
      dsm + (((size_t)(reinterpret_cast<unsigned char*>(fp + A.n_fp) - dsm) + 7) & ~(size_t)7));
  uint32_t* s_cov = reinterpret_cast<uint32_t*>(lst + A.n_fp);  // [2][n_fp] this iteration's partition_coverage
  if (SMEM_MASK)  // the global bitmask is all zero at the first launch and current at a resumed one
    for (int d = 0; d < A.ndirs; d++)
      for (uint32_t i = tid; i < A.mask_words; i += THREADS) smask[(size_t)d * A.mask_words + i] = __ldcg(A.d[d].ignored + i);
  for (uint32_t i = tid; i < 2u * A.p_words; i += THREADS) pm[i] = 0u;
  // Every block keeps its own copy of both partition_coverage tables in shared memory: loaded here (zero at the
  // first launch, current at a resumed one), then kept up to date from the winners (a winner whose postings lie in one
  // partition raises exactly that entry; the rare multi-partition winner makes the block reload the table).  296 blocks
  // fetching the same few lines every iteration was a 1 us hot spot in L2.
  __shared__ uint32_t s_reload[2];
  if (A.n_fp)
    for (int d = 0; d < A.ndirs; d++)
      for (uint32_t q = tid; q < A.n_part; q += THREADS) s_cov[(size_t)d * A.n_fp + q] = __ldcg(A.d[d].cov + q);
  if (tid == 0) { s_cnt2[0] = 0u; s_cnt2[1] = 0u; s_max[0] = 0u; s_max[1] = 0u; s_reload[0] = 0u; s_reload[1] = 0u; }
  __syncthreads();
  // block 0 keeps the global state (bitmask, partition_coverage, output); the other blocks score tiles
  const bool solo = gridDim.x == 1;
  const bool worker = solo || blockIdx.x > 0;
  const uint32_t wid = solo ? 0u : blockIdx.x - 1u, nworkers = solo ? 1u : gridDim.x - 1u;
  bool done[2] = {A.d[0].done0 != 0u, A.ndirs < 2 || A.d[1].done0 != 0u};
  // block-uniform loop state lives in shared memory (every thread writes the same value before it reads it), so
  // that the streaming loop of phase A has the registers to itself
  __shared__ uint32_t s_win[2], s_gsave[2], s_gd[2], s_pl[2], s_pt[2], s_npush[2];
  __shared__ unsigned long long s_key[2], s_pread[2];
  __shared__ unsigned long long s_evals[2];
  __shared__ uint32_t s_live[2];  // this block's live postings of the current iteration
  if (tid == 0) { s_evals[0] = 0ull; s_evals[1] = 0ull; s_live[0] = 0u; s_live[1] = 0u; s_pread[0] = 0ull; s_pread[1] = 0ull; s_npush[0] = A.it0; s_npush[1] = A.it0; }
  __syncthreads();
  // first k-mer of this warp's tile range; recomputed only when the set of running directions changes
  __shared__ uint32_t s_cfirst[WARPS], s_tb[WARPS], s_te[WARPS];  // this warp's cost-balanced tile range
  uint32_t cfirst_sig = 0xFFFFFFFFu;
  const bool lead = blockIdx.x == 0 && tid == 0;
  const bool wlead = blockIdx.x == (gridDim.x > 1 ? 1 : 0) && tid == 0;  // a worker block's clock (diagnostic)
  // phase timers live in shared memory (only one thread touches them): no registers for diagnostics
  __shared__ unsigned long long s_tm[12];  // 0 t_begin, 1 ta, 2 tb, 3 t_count, 4 t_tie, 5 wt0, 6.. dbg
  if (tid == 0) for (int q = 0; q < 12; q++) s_tm[q] = 0ull;
  __syncthreads();
  if (lead) s_tm[0] = globaltimer_ns();
#ifdef MSSPE_FINE_TIMERS  // build-time diagnostic: per-step clocks of worker block 1 (slots 0..11) and block 0 (12..23)
  __shared__ unsigned long long s_ft[12], s_flast;
  if (tid == 0) { for (int q = 0; q < 12; q++) s_ft[q] = 0ull; s_flast = globaltimer_ns(); }
#define FSTAMP(k) if ((wlead || lead) && tid == 0) { const unsigned long long t_ = globaltimer_ns(); s_ft[k] += t_ - s_flast; s_flast = t_; }
#else
#define FSTAMP(k)
#endif
  for (uint32_t it = A.it0;; it++) {
    const int par = it & 1;
    if (lead) s_tm[1] = globaltimer_ns();
    if (wlead) s_tm[5] = globaltimer_ns();
    // ---------------- phase A ----------------
    // The worker warps are split between the running directions in proportion to their tiles; every warp owns one
    // contiguous tile range of one direction and issues its first loads before the bitmask update below.
    RangeState S;
    int myd = 0;
    uint32_t my_gw = 0, my_nw = 1;
    if (worker) {
      const uint32_t W = nworkers * WARPS;
      my_gw = wid * WARPS + warp; my_nw = W;
      const uint32_t sig = (done[0] ? 0u : 1u) | (done[1] ? 0u : 2u);
      if (sig == 3u) {
        const uint32_t t0 = A.d[0].tail_t1, t1 = A.d[1].tail_t1;
        uint32_t W0 = (uint32_t)(((unsigned long long)W * t0 + (t0 + t1) / 2u) / (unsigned long long)max(t0 + t1, 1u));
        W0 = min(max(W0, 1u), W - 1u);
        if (my_gw < W0) { my_nw = W0; } else { myd = 1; my_gw -= W0; my_nw = W - W0; }
      } else {
        myd = done[0] ? 1 : 0;
      }
      if (sig != cfirst_sig) {  // the ranges only change when a direction finishes
        cfirst_sig = sig;
        if (lane == 0) {
          const GreedyDir& D = A.d[myd];
          const uint32_t tb = balanced_begin(D.cost_prefix, D.tail_t1, my_gw, my_nw);
          const uint32_t te = my_gw + 1u == my_nw ? D.tail_t1 : balanced_begin(D.cost_prefix, D.tail_t1, my_gw + 1u, my_nw);
          s_tb[warp] = tb; s_te[warp] = te;
          s_cfirst[warp] = tb < min(te, D.n_tiles) ? __ldg(D.tile_first + tb) : 0u;
        }
        __syncwarp();
      }
      // the two branches are the same code with the direction as a literal: the job's pointers then stay in the
      // kernel-parameter constant bank instead of occupying registers across the streaming loop
      if (myd == 0) { CountJob J; make_job<SMEM_MASK>(J, A, 0, smask, s_tb[warp], s_te[warp], s_cfirst[warp]); warp_count_begin(J, S, lane); }
      else          { CountJob J; make_job<SMEM_MASK>(J, A, 1, smask, s_tb[warp], s_te[warp], s_cfirst[warp]); warp_count_begin(J, S, lane); }
    }
    if (it > A.it0) {  // main.rs:371-378 for the previous winners (a direction that is not done has pushed one per iteration)
      const bool both = A.ndirs == 2 && !done[0] && !done[1];
      uint4 we = make_uint4(0u, 0xFFFFFFFFu, 0u, 0u);  // thread d: the winner's list entry, for the block's partition_coverage copy
      if (A.n_fp) {
        if (tid == 0 && !done[0]) we = __ldg(A.d[0].list_id + s_win[0]);
        else if (tid == 1 && A.ndirs > 1 && !done[1]) we = __ldg(A.d[1].list_id + s_win[1]);
      }
      for (int d = 0; d < A.ndirs; d++) {
        if (done[d]) continue;
        const GreedyDir& D = A.d[d];
        uint32_t* mask = SMEM_MASK ? smask + (size_t)d * A.mask_words : D.ignored;
        const uint32_t a = D.post_off[s_win[d]], b = D.post_off[s_win[d] + 1];  // live part is enough for the bitmask
        if (SMEM_MASK)
          for (uint32_t i = a + tid; i < b; i += THREADS) { const uint32_t sg = __ldg(D.postings + i); atomicOr(&mask[sg >> 5], 1u << (sg & 31u)); }
        if (blockIdx.x == 0 && !both) apply_winner_global<THREADS>(D, A.seg_part, A.uniform_parts, s_win[d], pm);
      }
      if (blockIdx.x == 0 && both) {  // half a block per direction (the direction is a literal in each branch)
        if (tid < THREADS / 2) apply_winner_half<THREADS>(A.d[0], A.seg_part, A.uniform_parts, s_win[0], pm, tid);
        else                   apply_winner_half<THREADS>(A.d[1], A.seg_part, A.uniform_parts, s_win[1], pm + A.p_words, tid - THREADS / 2);
        __syncthreads();
        for (uint32_t i = tid; i < 2u * A.p_words; i += THREADS) pm[i] = 0u;
      }
      if (we.y != 0xFFFFFFFFu) {  // threads 0 / 1 of a direction that pushed a winner
        if (!(we.y >> 31)) s_cov[(size_t)tid * A.n_fp + we.y] += 1u;
        else s_reload[tid] = 1u;
      }
      __syncthreads();
      if (!SMEM_MASK) grid_barrier(A.barrier, bar_target);  // the workers read the global bitmask block 0 has just updated
    }
    FSTAMP(0)
    if (worker) {
      uint32_t mymax = 0;
      unsigned long long live;
      if (myd == 0) { CountJob J; make_job<SMEM_MASK>(J, A, 0, smask, s_tb[warp], s_te[warp], s_cfirst[warp]); live = warp_count_run<SMEM_MASK>(J, S, mymax, lane); }
      else          { CountJob J; make_job<SMEM_MASK>(J, A, 1, smask, s_tb[warp], s_te[warp], s_cfirst[warp]); live = warp_count_run<SMEM_MASK>(J, S, mymax, lane); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mymax = max(mymax, __shfl_xor_sync(0xffffffffu, mymax, o));
      if (lane == 0) { if (mymax) atomicMax(&s_max[myd], mymax); if (live) atomicAdd(&s_live[myd], (uint32_t)live); }
    }
    FSTAMP(1)
    __syncthreads();
    if (tid == 0 && worker)
      for (int d = 0; d < A.ndirs; d++)
        if (!done[d]) {
          if (s_max[d]) atomicMax(&A.d[d].ctl->pg[par], s_max[d]);
          if (s_live[d]) atomicAdd(&A.d[d].ctl->plive[par], s_live[d]);
          s_evals[d] += s_live[d]; s_max[d] = 0u; s_live[d] = 0u;
        }
    if (wlead) { const unsigned long long t = globaltimer_ns(); s_tm[8] += t - s_tm[5]; s_tm[5] = t; }
    FSTAMP(2)
    grid_barrier_cg();
    FSTAMP(3)
    if (wlead) { const unsigned long long t = globaltimer_ns(); s_tm[9] += t - s_tm[5]; s_tm[5] = t; }
    if (lead) { s_tm[2] = globaltimer_ns(); s_tm[3] += s_tm[2] - s_tm[1]; }
    // ---------------- phase B ----------------
    // Both directions at once: the maxima, this block's slice of both freq[] arrays and both partition_coverage
    // tables are requested before anything is waited for (one L2 round trip instead of one per direction and step).
    // One thread per direction fetches the maximum and shares it: the same word requested by every warp of the
    // grid (4736 x 2 L2 requests for one line) is a hot spot in its L2 slice.
    // (unconditionally: done[] is indexed by loop variables and lives in local memory, whose L1 lines the barrier's fence
    // has just invalidated -- testing it first would put a second L2 round trip in front of this load)
    if (tid < 2 && tid < A.ndirs) s_gd[tid] = __ldcg(&A.d[tid].ctl->pg[par]);
    if (A.n_fp)  // a multi-partition winner: re-read the table (block 0 finished updating it before the barrier)
      for (int d = 0; d < A.ndirs; d++)
        if (!done[d] && s_reload[d]) for (uint32_t q = tid; q < A.n_part; q += THREADS) s_cov[(size_t)d * A.n_fp + q] = __ldcg(A.d[d].cov + q);
    uint32_t f_first[2][2];
    {
      const uint32_t stride = gridDim.x * THREADS, c = blockIdx.x * THREADS + tid;
#pragma unroll
      for (int d = 0; d < 2; d++)
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const uint32_t cc = c + (uint32_t)q * stride;
          f_first[d][q] = (d < A.ndirs && cc < A.d[d].n_codes) ? __ldcg(A.d[d].freq + cc) : 0u;  // a maximum is >= 2; a finished direction is skipped below
        }
    }
    __syncthreads();
    FSTAMP(4)
    if (tid == 0) { s_reload[0] = 0u; s_reload[1] = 0u; }
    const uint32_t gd[2] = {done[0] ? 0u : s_gd[0], done[1] ? 0u : s_gd[1]};
    for (int d = 0; d < A.ndirs; d++) {
      if (done[d]) continue;
      const GreedyDir& D = A.d[d];
      const uint32_t g = gd[d];
      s_gsave[d] = g;
      // next iteration's slots: plain stores only -- block 0 is on the critical path of every barrier, its lead thread must not wait for L2
      if (lead) { D.ctl->pg[par ^ 1] = 0; D.ctl->pt[par ^ 1] = 0; D.ctl->pk[par ^ 1] = 0ull; D.ctl->plive[par ^ 1] = 0; s_pread[d] += D.stream_total; }
      if (g <= 1u) {  // None or freq == 1: stop before the push (main.rs:353-366)
        done[d] = true;
        if (lead) { D.ctl->iterations = it + 1; D.ctl->n_out = it; D.ctl->done = 1; }
        continue;
      }
      // collect this block's k-mers at the maximum (grid-stride over freq[]; the first two strides are already here)
      uint32_t* tied = s_tied + d * THREADS;
      {
        const uint32_t stride = gridDim.x * THREADS;
        uint32_t c = blockIdx.x * THREADS + tid;
        if (f_first[d][0] == g) { const uint32_t q = atomicAdd(&s_cnt2[d], 1u); if (q < (uint32_t)THREADS) tied[q] = c; }
        if (f_first[d][1] == g) { const uint32_t q = atomicAdd(&s_cnt2[d], 1u); if (q < (uint32_t)THREADS) tied[q] = c + stride; }
        for (c += 2u * stride; c < D.n_codes; c += stride)
          if (__ldcg(D.freq + c) == g) { const uint32_t q = atomicAdd(&s_cnt2[d], 1u); if (q < (uint32_t)THREADS) tied[q] = c; }
      }
    }
    __syncthreads();
    FSTAMP(5)
    // A k-mer whose postings all lie in ONE partition p (list_id[].y, fixed at index build; the rule in a pre-aligned
    // alignment) scores 0.0 + 1/(partition_coverage[p] + 1) whichever of them are live: one thread each, no list scan,
    // both directions in the same round trip.  What is scored here is flagged (bit 31) in the tied list.
    {
      const bool on0 = !done[0] && s_cnt2[0] <= (uint32_t)THREADS, on1 = A.ndirs > 1 && !done[1] && s_cnt2[1] <= (uint32_t)THREADS;
      uint4 e0 = make_uint4(0u, 0x80000000u, 0u, 0u), e1 = e0;
      const uint32_t c0 = on0 && (uint32_t)tid < s_cnt2[0] ? s_tied[tid] : 0xFFFFFFFFu;
      const uint32_t c1 = on1 && (uint32_t)tid < s_cnt2[1] ? s_tied[THREADS + tid] : 0xFFFFFFFFu;
      if (c0 != 0xFFFFFFFFu) e0 = __ldg(A.d[0].list_id + c0);
      if (c1 != 0xFFFFFFFFu) e1 = __ldg(A.d[1].list_id + c1);
      if (!(e0.y >> 31)) {
        const uint32_t cv = A.n_fp ? s_cov[e0.y] : __ldcg(A.d[0].cov + e0.y);
        const float sc1 = __fdiv_rn(1.0f, __fadd_rn(__uint2float_rn(cv), 1.0f));
        atomicMax(&A.d[0].ctl->pk[par], ((unsigned long long)__float_as_uint(sc1) << 32) | (unsigned long long)(0xFFFFFFFFu - c0));
        s_tied[tid] = c0 | 0x80000000u;
      }
      if (!(e1.y >> 31)) {
        const uint32_t cv = A.n_fp ? s_cov[(size_t)A.n_fp + e1.y] : __ldcg(A.d[1].cov + e1.y);
        const float sc1 = __fdiv_rn(1.0f, __fadd_rn(__uint2float_rn(cv), 1.0f));
        atomicMax(&A.d[1].ctl->pk[par], ((unsigned long long)__float_as_uint(sc1) << 32) | (unsigned long long)(0xFFFFFFFFu - c1));
        s_tied[THREADS + tid] = c1 | 0x80000000u;
      }
    }
    // the usual case ends here: everything tied in this block was scored above (one barrier, which also tells)
    const uint32_t nh0 = done[0] ? 0u : s_cnt2[0], nh1 = (A.ndirs < 2 || done[1]) ? 0u : s_cnt2[1];
    const bool mine_left = ((uint32_t)tid < min(nh0, (uint32_t)THREADS) && !(s_tied[tid] >> 31)) ||
                           ((uint32_t)tid < min(nh1, (uint32_t)THREADS) && !(s_tied[THREADS + tid] >> 31)) ||
                           nh0 > (uint32_t)THREADS || nh1 > (uint32_t)THREADS;
    const bool slow = __syncthreads_or(mine_left ? 1 : 0) != 0;
    FSTAMP(6)
    if (!slow && tid == 0) {
      if (nh0) atomicAdd(&A.d[0].ctl->pt[par], nh0);
      if (nh1) atomicAdd(&A.d[1].ctl->pt[par], nh1);
      s_cnt2[0] = 0u; s_cnt2[1] = 0u;
    }
    if (slow)
    for (int d = 0; d < A.ndirs; d++) {
      if (done[d]) continue;
      const GreedyDir& D = A.d[d];
      const uint32_t g = gd[d];
      const uint32_t* mask = SMEM_MASK ? smask + (size_t)d * A.mask_words : D.ignored;
      const uint32_t* covp = A.n_fp ? s_cov + (size_t)d * A.n_fp : D.cov;
      uint32_t* tied = s_tied + d * THREADS;
      const uint32_t n_here = s_cnt2[d];
      __syncthreads();
      if (tid == 0) s_cnt2[d] = 0u;
      // the usual case: the block's ties fit the list; otherwise (tie storms) re-collect them chunk by chunk
      const uint32_t n_chunks = n_here <= (uint32_t)THREADS ? 1u : (D.n_codes + gridDim.x * THREADS - 1u) / (gridDim.x * THREADS);
      for (uint32_t chunk = 0; chunk < n_chunks; chunk++) {
        uint32_t nt = n_here;
        if (n_here > (uint32_t)THREADS) {
          __syncthreads();
          const uint32_t c = (chunk * gridDim.x + blockIdx.x) * THREADS + tid;
          if (c < D.n_codes && __ldcg(D.freq + c) == g) tied[atomicAdd(&s_cnt2[d], 1u)] = c;
          __syncthreads();
          nt = s_cnt2[d];
          __syncthreads();
          if (tid == 0) s_cnt2[d] = 0u;
        }
        if (n_here > (uint32_t)THREADS) {  // (the usual case was scored before this loop, both directions at once)
          single_partition_ties(D, tied, nt, covp, A.n_fp != 0u, par, tid);
          __syncthreads();
        }
        for (uint32_t t = 0; t < nt; t++) {
          const uint32_t cc = tied[t];
          if (cc >> 31) continue;  // block-uniform: scored above
          float score;
          if (A.n_fp) {
            score = block_tie_score<SMEM_MASK, false, THREADS>(cc, D.post_off, D.postings, mask, A.seg_part, A.uniform_parts, covp, A.n_part, fp, lst, s_sc);
          } else {  // more partitions than the shared-memory scratch holds: one warp, bitmap of seen partitions
            if (warp == 0) {
              const float sw = warp_tie_score<SMEM_MASK, true>(cc, D.post_off, D.postings, mask, A.seg_part, D.cov, seen, A.p_words, lane);
              if (lane == 0) s_sc[1] = __float_as_uint(sw);
            }
            __syncthreads();
            score = __uint_as_float(s_sc[1]);
          }
          if (tid == 0) atomicMax(&D.ctl->pk[par], ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xFFFFFFFFu - cc));
          __syncthreads();
        }
      }
      if (tid == 0 && n_here) atomicAdd(&D.ctl->pt[par], n_here);
      __syncthreads();
    }
    if (wlead) { const unsigned long long t = globaltimer_ns(); s_tm[10] += t - s_tm[5]; s_tm[5] = t; }
    FSTAMP(7)
    grid_barrier_cg();
    FSTAMP(8)
    if (wlead) { const unsigned long long t = globaltimer_ns(); s_tm[11] += t - s_tm[5]; s_tm[5] = t; }
    if (lead) s_tm[4] += globaltimer_ns() - s_tm[2];
    // ---------------- winner ----------------
    bool all_done = true;
    if (tid < 2 && tid < A.ndirs && !(tid == 0 ? done[0] : done[1])) {  // again one thread per direction asks L2
      s_key[tid] = __ldcg(&A.d[tid].ctl->pk[par]);
      if (blockIdx.x == 0) s_pt[tid] = __ldcg(&A.d[tid].ctl->pt[par]);
      if (A.compact_min) s_pl[tid] = __ldcg(&A.d[tid].ctl->plive[par]);
    }
    __syncthreads();
    FSTAMP(9)
    for (int d = 0; d < A.ndirs; d++) {
      if (done[d]) continue;
      const GreedyDir& D = A.d[d];
      const unsigned long long key = s_key[d];
      const uint32_t c = 0xFFFFFFFFu - (uint32_t)key;
      s_win[d] = c;
      const uint32_t g = s_gsave[d];
      if (lead) {
        msspe_candidate w;
        // .code holds the stream list for now; the words are filled in when the launch ends (no dependent loads here)
        w.code = c; w.freq = g; w.n_tied = s_pt[d]; w.tie_score = __uint_as_float((uint32_t)(key >> 32)); w.reserved = 0;
        D.out[it] = w;  // a direction that is not done has pushed one winner per iteration
        s_npush[d] = it + 1u;
      }
      if (g < A.mms || it + 1u >= A.max_iter) {  // main.rs:387-390 and the loop bound :344
        done[d] = true;
        if (lead) { D.ctl->iterations = it + 1; D.ctl->n_out = it + 1u; D.ctl->done = 1; }
      }
      all_done = all_done && done[d];
    }
    FSTAMP(10)
    if (all_done) break;
    if (A.compact_min) {  // less than half of what a running direction streams is still live: leave for a compaction
      bool leave = false;
      for (int d = 0; d < A.ndirs; d++)
        if (!done[d] && A.d[d].stream_total >= A.compact_min && 2u * s_pl[d] < A.d[d].stream_total) leave = true;
      if (leave) {
        for (int d = 0; d < A.ndirs; d++) {
          if (done[d]) continue;
          if (blockIdx.x == 0) apply_winner_global<THREADS>(A.d[d], A.seg_part, A.uniform_parts, s_win[d], pm);  // nothing stays pending
          if (lead) { A.d[d].ctl->resume_it = it + 1u; A.d[d].ctl->exit_compact = 1u; }
        }
        break;
      }
    }
  }
  if (blockIdx.x == 0) {  // the winners of this launch: stream list -> word
    __syncthreads();
    for (int d = 0; d < A.ndirs; d++)
      for (uint32_t i = A.it0 + tid; i < s_npush[d]; i += THREADS)
        A.d[d].out[i].code = A.d[d].codes[__ldg(A.d[d].list_id + (uint32_t)A.d[d].out[i].code).x];
  }
  for (int d = 0; d < A.ndirs; d++) {
    if (lead) A.d[d].ctl->postings_read += s_pread[d];
    if (tid == 0 && s_evals[d]) atomicAdd(&A.d[d].ctl->evals, s_evals[d]);
    if (lead) { A.d[d].ctl->t_count_ns += s_tm[3]; A.d[d].ctl->t_tie_ns += s_tm[4]; A.d[d].ctl->t_total_ns += globaltimer_ns() - s_tm[0]; }
    if (wlead) for (int q = 0; q < 4; q++) A.d[d].ctl->t_dbg[4 + q] += s_tm[8 + q];
#ifdef MSSPE_FINE_TIMERS
    if (d == 0 && (wlead || lead)) for (int q = 0; q < 12; q++) A.d[0].ctl->t_fine[(lead ? 12 : 0) + q] += s_ft[q];
#endif
  }
}

struct DirRun {
  DirIndex* D; cudaStream_t st; uint32_t* tile_first; uint32_t n_tiles; bool smem_mask; uint32_t mask_words;
  unsigned count_grid; size_t count_smem; unsigned tie_grid; size_t tie_smem; uint32_t p_words;
};

int prepare_dir(msspe_ctx* c, int dir, uint32_t max_iter, cudaStream_t st, DirRun* r) {
  DirIndex& D = c->dir[dir];
  r->D = &D; r->st = st;
  const uint64_t G = c->n_segments;
  if (D.out_capacity < max_iter) {
    msspe_dev_free(c, D.out); D.out = nullptr;
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.out, (uint64_t)(max_iter ? max_iter : 1) * sizeof(msspe_candidate), c->stream));
    D.out_capacity = max_iter;
  }
  r->mask_words = (uint32_t)div_up_u64(G, 32) + 1u;
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.ignored, 0, (size_t)r->mask_words * 4, st));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.cov, 0, 65536 * 4, st));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.pmark, 0, 2048 * 4, st));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.ctl, 0, sizeof(SelectCtl), st));
  r->n_tiles = D.n_tiles;
  r->tile_first = D.tile_first;
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.acc, 0, (uint64_t)(r->n_tiles + 1) * 8, st));
  struct { int multiProcessorCount; size_t sharedMemPerBlockOptin; } prop = {c->sm_count, c->smem_optin};
  const size_t mask_bytes = (size_t)r->mask_words * 4;
  r->smem_mask = mask_bytes + 8192 <= (size_t)prop.sharedMemPerBlockOptin;
  r->count_smem = r->smem_mask ? mask_bytes : 0;
  int per_sm = 1;
  if (r->smem_mask) {
    MSSPE_CUDA_TRY(c, cudaFuncSetAttribute(count_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mask_bytes));
    MSSPE_CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, count_kernel<true>, CNT_THREADS, r->count_smem));
  } else {
    MSSPE_CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, count_kernel<false>, CNT_THREADS, 0));
  }
  if (per_sm < 1) per_sm = 1;
  const unsigned resident = (unsigned)prop.multiProcessorCount * (unsigned)per_sm;
  r->count_grid = r->n_tiles < resident ? r->n_tiles : resident;
  r->p_words = (c->max_partition + 32u) / 32u;
  r->tie_smem = (size_t)r->p_words * 4 * (TIE_THREADS / 32);
  const uint32_t n_chunks = (uint32_t)div_up_u64(D.n_codes, 32);
  const unsigned tie_blocks = (n_chunks + (TIE_THREADS / 32) - 1) / (TIE_THREADS / 32);
  r->tie_grid = tie_blocks < (unsigned)prop.multiProcessorCount * 8u ? tie_blocks : (unsigned)prop.multiProcessorCount * 8u;
  if (r->tie_smem > 48 * 1024)
    MSSPE_CUDA_TRY(c, cudaFuncSetAttribute(tie_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r->tie_smem));
  return MSSPE_OK;
}

void launch_iteration(msspe_ctx* c, DirRun& r, uint32_t max_iter, uint32_t mms, uint32_t mode, bool first,
                      std::vector<cudaEvent_t>* pev) {
  DirIndex& D = *r.D;
  const uint32_t nc = (uint32_t)D.n_codes, np = (uint32_t)D.n_records;
  const bool recount = mode == MSSPE_SELECT_RECOUNT || first;
  if (recount) {
    if (r.n_tiles) {
      if (pev) {  // profiling: bracket this launch with events on its own stream
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        pev->push_back(e0); pev->push_back(e1);
        cudaEventRecord(e0, r.st);
      }
      if (r.smem_mask)
        count_kernel<true><<<r.count_grid, CNT_THREADS, r.count_smem, r.st>>>(D.postings, D.post_off, r.tile_first, nc, np, r.n_tiles,
                                                                               D.ignored, r.mask_words, D.freq, D.acc, D.ctl);
      else
        count_kernel<false><<<r.count_grid, CNT_THREADS, 0, r.st>>>(D.postings, D.post_off, r.tile_first, nc, np, r.n_tiles,
                                                                     D.ignored, r.mask_words, D.freq, D.acc, D.ctl);
      c->timing.kernel_launches++;
      if (pev) cudaEventRecord(pev->back(), r.st);
    }
  } else if (nc) {
    freq_max_kernel<<<min(592u, (nc + 255u) / 256u), 256, 0, r.st>>>(D.freq, nc, D.ctl);
    c->timing.kernel_launches++;
  }
  if (nc) {
    tie_kernel<<<r.tie_grid, TIE_THREADS, r.tie_smem, r.st>>>(D.freq, nc, D.post_off, D.postings, D.ignored, c->d_seg_part, D.cov,
                                                             D.ctl, r.p_words);
    c->timing.kernel_launches++;
  }
  if (mode == MSSPE_SELECT_INCREMENTAL)
    update_kernel<true><<<1, UPD_THREADS, 0, r.st>>>(D.codes, D.post_off, D.postings, D.ignored, c->d_seg_part, D.cov, D.pmark, D.ctl,
                                                      D.out, max_iter, mms, D.fwd_ids, c->slots, D.freq);
  else
    update_kernel<false><<<1, UPD_THREADS, 0, r.st>>>(D.codes, D.post_off, D.postings, D.ignored, c->d_seg_part, D.cov, D.pmark, D.ctl,
                                                       D.out, max_iter, mms, D.fwd_ids, c->slots, D.freq);
  c->timing.kernel_launches++;
}

// The stream a direction currently scores: the pristine one of the index, or a compacted working copy.
struct ScoreStream {
  uint32_t* postings; uint32_t* off; uint4* id; uint32_t* tile_first; uint32_t* cost;
  uint32_t lists, list_post, tiles, tail_end;
  bool owned;
};

void free_stream(msspe_ctx* c, ScoreStream& S) {
  if (!S.owned) return;
  msspe_dev_free(c, S.postings); msspe_dev_free(c, S.off); msspe_dev_free(c, S.id); msspe_dev_free(c, S.tile_first); msspe_dev_free(c, S.cost);
  S.owned = false;
}

// Exclusive prefix of the tile costs of a stream (list tiles, then tail tiles); entry [all_tiles] = total.
int build_cost_prefix(msspe_ctx* c, const uint32_t* tile_first, uint32_t list_tiles, uint32_t n_lists, uint32_t tail_end, uint32_t** out, cudaStream_t st) {
  const uint32_t all_tiles = (uint32_t)div_up_u64(tail_end, CNT_TILE) > list_tiles ? (uint32_t)div_up_u64(tail_end, CNT_TILE) : list_tiles;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(out, ((uint64_t)all_tiles + 2) * 4, c->stream));
  tile_cost_kernel<<<(all_tiles + 256u) / 256u, 256, 0, st>>>(tile_first, list_tiles, n_lists, all_tiles, *out);
  c->timing.kernel_launches++;
  return msspe_exclusive_scan_u32(c, *out, *out, (uint64_t)all_tiles + 1, nullptr, st);
}

// Drop the dead postings of a stream (order-preserving), given the current global bitmask of the direction.
int compact_stream(msspe_ctx* c, DirIndex& D, ScoreStream& S, cudaStream_t st) {
  const uint32_t nl = S.lists, tail_begin = S.tiles * (uint32_t)CNT_TILE;
  uint32_t *cnt = nullptr, *multi = nullptr, *mlen = nullptr, *single = nullptr, *d_tot = nullptr;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&cnt, ((uint64_t)nl + 1) * 4, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&multi, ((uint64_t)nl + 1) * 4, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&mlen, ((uint64_t)nl + 1) * 4, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&single, ((uint64_t)nl + 1) * 4, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_tot, 16, st));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(d_tot, 0, 16, st));
  uint32_t tot[4] = {0, 0, 0, 0};  // lists kept, their postings, lists left with one posting, live postings of the old tail
  if (nl) {
    stream_live_kernel<<<(unsigned)div_up_u64(nl, 256), 256, 0, st>>>(S.off, S.postings, nl, D.ignored, cnt, multi, mlen, single);
    c->timing.kernel_launches++;
    int rc = msspe_exclusive_scan_u32(c, multi, multi, nl, d_tot, st);
    if (!rc) rc = msspe_exclusive_scan_u32(c, mlen, mlen, nl, d_tot + 1, st);
    if (!rc) rc = msspe_exclusive_scan_u32(c, single, single, nl, d_tot + 2, st);
    if (rc) return rc;
  }
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(tot, d_tot, 12, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  ScoreStream N{};
  N.lists = tot[0]; N.list_post = tot[1]; N.tiles = (uint32_t)div_up_u64(N.list_post, CNT_TILE); N.owned = true;
  const uint32_t n_tail_begin = N.tiles * (uint32_t)CNT_TILE;
  const uint64_t cap = (uint64_t)n_tail_begin + tot[2] + (S.tail_end - tail_begin) + CNT_TILE;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&N.postings, cap * 4, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&N.off, ((uint64_t)N.lists + 1) * 4, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&N.id, ((uint64_t)N.lists + 1) * 16, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&N.tile_first, ((uint64_t)N.tiles + 1) * 4, c->stream));
  if (nl) {
    stream_compact_kernel<<<(unsigned)div_up_u64(nl, 256), 256, 0, st>>>(S.off, S.postings, S.id, nl, D.ignored, cnt, multi, mlen, single, N.lists,
                                                                         N.list_post, n_tail_begin, N.postings, N.off, N.id);
    c->timing.kernel_launches++;
  } else {
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(N.off, 0, 4, st));
  }
  if (S.tail_end > tail_begin) {
    stream_tail_compact_kernel<<<(unsigned)div_up_u64(S.tail_end - tail_begin, 256), 256, 0, st>>>(S.postings, tail_begin, S.tail_end, D.ignored, N.postings,
                                                                                               n_tail_begin + tot[2], d_tot + 3);
    c->timing.kernel_launches++;
  }
  if (N.tiles) {
    tile_first_kernel<<<(N.tiles + 255) / 256, 256, 0, st>>>(N.off, N.lists, N.tiles, N.tile_first);
    c->timing.kernel_launches++;
  }
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&tot[3], d_tot + 3, 4, cudaMemcpyDeviceToHost, st));
  for (uint32_t* p : {cnt, multi, mlen, single, d_tot}) MSSPE_CUDA_TRY(c, cudaFreeAsync(p, st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  N.tail_end = n_tail_begin + tot[2] + tot[3];
  int rc2 = build_cost_prefix(c, N.tile_first, N.tiles, N.lists, N.tail_end, &N.cost, st);
  if (rc2) return rc2;
  free_stream(c, S);
  S = N;
  return MSSPE_OK;
}

int run_select_persistent(msspe_ctx* c, int ndirs, const int* dirs, uint32_t max_iter, uint32_t mms,
                          msspe_candidate** outs, uint32_t** n_outs) {
  cudaStream_t st = c->stream;
  const uint64_t G = c->n_segments;
  GreedyArgs A{};
  A.ndirs = ndirs; A.max_iter = max_iter; A.mms = mms; A.seg_part = c->d_seg_part;
  A.uniform_parts = (c->uniform_parts && c->uniform_parts <= 65536u) ? c->uniform_parts : 0u;
  A.barrier = c->dir[dirs[0]].pmark;  // 8 KB scratch, unused by this kernel
  A.mask_words = (uint32_t)div_up_u64(G, 32) + 1u;
  A.p_words = (c->max_partition + 32u) / 32u;
  A.compact_min = 1u << 20;  // compaction passes cost ~0.3 ms of launches and syncs: not worth it for small streams
  if (const char* e = getenv("MSSPE_COMPACT_MIN")) A.compact_min = (uint32_t)strtoul(e, nullptr, 10);
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[2], st));
  ScoreStream cur[2] = {};
  for (int i = 0; i < ndirs; i++) {
    DirIndex& D = c->dir[dirs[i]];
    if (D.out_capacity < max_iter) {
      msspe_dev_free(c, D.out); D.out = nullptr;
      MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.out, (uint64_t)max_iter * sizeof(msspe_candidate), c->stream));
      D.out_capacity = max_iter;
    }
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.ignored, 0, (size_t)A.mask_words * 4, st));
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.cov, 0, 65536 * 4, st));
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.ctl, 0, sizeof(SelectCtl), st));
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.acc, 0, (uint64_t)(D.n_tiles + 1) * 8, st));
    int rc = msspe_select_prepare_stream(c, dirs[i], st);
    if (rc) return rc;
    cur[i] = ScoreStream{D.s_postings, D.s_off, D.s_id, D.s_tile_first, D.s_cost, D.s_lists, D.s_list_post, D.s_tiles,
                         D.s_tiles * (uint32_t)CNT_TILE + ((uint32_t)D.n_records - D.s_list_post), false};
    memset(&c->h_ctl[i], 0, sizeof(SelectCtl));
  }
  A.n_part = c->max_partition + 1u;
  A.n_fp = A.n_part <= 4096u ? A.n_part : 0u;
  // tied k-mers with short lists are scored by one warp each (a partition bitmap per warp); needs the shared-memory
  // partition_coverage copy (n_fp) and small bitmaps
  const size_t aux = (size_t)3 * A.p_words * 4 + (size_t)A.n_fp * 4 + (size_t)A.n_fp * 8 + (size_t)2 * A.n_fp * 4 + 16;  // pm (x2), seen, fp, lst, cov
  const size_t mask_bytes = (size_t)ndirs * A.mask_words * 4;
  const bool smem_mask = mask_bytes + aux + (size_t)4096 + 1024 <= c->smem_optin;
  const size_t smem = aux + (smem_mask ? mask_bytes : 0);
  if (smem + (size_t)4096 + 1024 > c->smem_optin) {
    c->set_error("msspe_select: %u partitions need %zu B of shared memory (device offers %zu)", c->max_partition + 1, smem, c->smem_optin);
    return MSSPE_ERR_CAPACITY;
  }
  // 512-thread blocks, as many per SM as fit (<= 3); when the bitmask leaves room for one block only, 1024 threads
  int per_sm = 0, threads = 512;
  void* fn = smem_mask ? (void*)greedy_persistent_kernel<true, 512> : (void*)greedy_persistent_kernel<false, 512>;
  MSSPE_CUDA_TRY(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MSSPE_CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&per_sm, fn, 512, smem, cudaOccupancyDefault));
  if (per_sm < 2 || getenv("MSSPE_PERSIST_1024")) {
    fn = smem_mask ? (void*)greedy_persistent_kernel<true, 1024> : (void*)greedy_persistent_kernel<false, 1024>;
    threads = 1024;
    MSSPE_CUDA_TRY(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MSSPE_CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&per_sm, fn, 1024, smem, cudaOccupancyDefault));
  }
  if (per_sm < 1) { c->set_error("msspe_select: persistent kernel does not fit on an SM"); return MSSPE_ERR_CAPACITY; }
  int want = 2;
  if (const char* e = getenv("MSSPE_PERSIST_BLOCKS_PER_SM")) want = atoi(e) > 0 ? atoi(e) : want;
  if (per_sm > want) per_sm = want;
  const unsigned grid = (unsigned)c->sm_count * (unsigned)per_sm;
  uint32_t launches = 0, compactions = 0;
  for (;;) {  // one launch per stretch between stream compactions
    for (int i = 0; i < ndirs; i++) {
      DirIndex& D = c->dir[dirs[i]];
      GreedyDir& g = A.d[i];
      const ScoreStream& S = cur[i];
      g.postings = S.postings; g.post_off = S.off; g.tile_first = S.tile_first; g.codes = D.codes; g.list_id = S.id;
      g.n_codes = S.lists; g.n_post = S.list_post; g.n_tiles = S.tiles;
      g.tail_t0 = S.tiles; g.tail_end = S.tail_end; g.tail_t1 = (uint32_t)div_up_u64(S.tail_end, CNT_TILE);
      g.stream_total = S.list_post + (S.tail_end - S.tiles * (uint32_t)CNT_TILE);
      g.full_off = D.post_off; g.full_postings = D.postings; g.cost_prefix = S.cost;
      g.done0 = c->h_ctl[i].done;
      g.ignored = D.ignored; g.freq = D.freq; g.acc = D.acc; g.cov = D.cov; g.ctl = D.ctl; g.out = D.out;
    }
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(A.barrier, 0, 4, st));
    void* kargs[] = {(void*)&A};
    { KPROF(c, KP_GREEDY_WHOLE, st, 0) MSSPE_CUDA_TRY(c, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), kargs, smem, st)); }
    launches++;
    for (int i = 0; i < ndirs; i++)
      MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&c->h_ctl[i], c->dir[dirs[i]].ctl, sizeof(SelectCtl), cudaMemcpyDeviceToHost, st));
    MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
    bool resume = false;
    uint32_t it0 = 0;
    for (int i = 0; i < ndirs; i++)
      if (!c->h_ctl[i].done && c->h_ctl[i].exit_compact) { resume = true; it0 = c->h_ctl[i].resume_it; }
    if (!resume) break;
    for (int i = 0; i < ndirs; i++) {
      DirIndex& D = c->dir[dirs[i]];
      if (c->h_ctl[i].done) continue;
      int rc = compact_stream(c, D, cur[i], st);
      if (rc) return rc;
      // the per-iteration slots of the control block restart clean; everything cumulative stays
      SelectCtl& h = c->h_ctl[i];
      h.pg[0] = h.pg[1] = h.pt[0] = h.pt[1] = h.plive[0] = h.plive[1] = 0; h.pk[0] = h.pk[1] = 0ull; h.exit_compact = 0;
      MSSPE_CUDA_TRY(c, cudaMemcpyAsync(D.ctl, &h, sizeof(SelectCtl), cudaMemcpyHostToDevice, st));
    }
    A.it0 = it0;
    compactions++;
  }
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[3], st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  float ms = 0.f;
  MSSPE_CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]));
  for (int i = 0; i < ndirs; i++) {
    const int d = dirs[i];
    const uint32_t n = c->h_ctl[i].n_out;
    if (n) MSSPE_CUDA_TRY(c, cudaMemcpyAsync(outs[i], c->dir[d].out, (size_t)n * sizeof(msspe_candidate), cudaMemcpyDeviceToHost, st));
    *n_outs[i] = n;
    c->timing.select_ms[d] = ms;
    c->timing.select_evals[d] = c->h_ctl[i].evals;
    c->timing.select_iterations[d] = c->h_ctl[i].iterations;
    c->timing.select_postings_read[d] = c->h_ctl[i].postings_read;  // lists + counted-only tail of the stream of each iteration
    c->timing.count_kernel_launches[d] = c->h_ctl[i].iterations;
    // phases of the two directions are interleaved inside one kernel: attribute the phase time once (to dir 0)
    c->timing.count_kernel_ms[d] = i == 0 ? (float)(c->h_ctl[i].t_count_ns * 1e-6) : 0.f;
    if (getenv("MSSPE_DEBUG_TIMERS") && i == 0)
      fprintf(stderr, "[msspe] grid %u x %d, smem %zu B, smem_mask %d, ndirs %d, launches %u (stream compactions %u)\n", grid, threads, smem, (int)smem_mask, ndirs, launches, compactions);
    if (getenv("MSSPE_DEBUG_TIMERS") && i == 0)
      fprintf(stderr, "[msspe] persistent greedy: total %.3f ms (wall %.3f), iterations %u | coverage-scoring phases %.3f ms, arg-max phases %.3f ms (block 0 clock, barriers included)\n",
              c->h_ctl[i].t_total_ns * 1e-6, ms, c->h_ctl[i].iterations, c->h_ctl[i].t_count_ns * 1e-6, c->h_ctl[i].t_tie_ns * 1e-6);
    if (getenv("MSSPE_DEBUG_TIMERS") && i == 0 && c->h_ctl[i].t_fine[3]) {
      fprintf(stderr, "[msspe]   fine (us/iteration) 0 update+sync | 1 count | 2 sync+atomics | 3 barrier1 | 4 max | 5 collect | 6 single-partition ties | 7 other ties | 8 barrier2 | 9 key | 10 winner\n");
      for (int b = 0; b < 2; b++) {
        fprintf(stderr, "[msspe]   %s:", b ? "block 0" : "block 1");
        for (int q = 0; q < 11; q++) fprintf(stderr, " %.2f", c->h_ctl[i].t_fine[12 * b + q] * 1e-3 / (c->h_ctl[i].iterations ? c->h_ctl[i].iterations : 1));
        fprintf(stderr, "\n");
      }
    }
    if (getenv("MSSPE_DEBUG_TIMERS") && i == 0)
      fprintf(stderr, "[msspe]   worker block 1: phaseA work %.3f sync %.3f | phaseB work %.3f sync %.3f ms\n", c->h_ctl[i].t_dbg[4] * 1e-6,
              c->h_ctl[i].t_dbg[5] * 1e-6, c->h_ctl[i].t_dbg[6] * 1e-6, c->h_ctl[i].t_dbg[7] * 1e-6);
    free_stream(c, cur[i]);
  }
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  return MSSPE_OK;
}

// ------------------------------------------------------------------------------------------------------------
// The greedy loop with INCREMENTAL counts as one persistent cooperative kernel (MSSPE_SELECT_INCREMENTAL; also what
// a recount request falls back to when a direction's bitmask does not fit shared memory).  freq[] is exact at all
// times: when a winner covers a segment for the first time, each of the segment's k-mers (forward index) loses one
// live segment.  hist[f] = number of k-mers with f live segments, so the maximum only walks down the histogram and
// the number of tied k-mers is hist[max].  Per iteration, for both directions:
//   phase 1  maximum from the histogram; stop tests; scan freq[] for the tied k-mers, score them (partition_tie_score,
//            global bitmask), atomicMax of (score, ~id)
//   barrier
//   phase 2  the winner's postings (ALL of them, main.rs:371-378) are split over the whole grid: bitmask bit,
//            partition_coverage through a global partition bitmap, and for every newly covered segment one warp
//            decrements freq[] / moves hist[] for its k-mers
//   barrier
// Reference-equivalent evals = the live records a recount would have examined = a running total minus the decrements.
struct IncDir {
  const uint32_t* post_off; const uint32_t* postings; const uint64_t* codes; const uint32_t* fwd_ids;
  uint32_t* freq; uint32_t* hist; uint32_t* ignored; uint32_t* cov; uint32_t* pmark;  // pmark: [2][2048] partition bitmaps by parity
  SelectCtl* ctl; msspe_candidate* out;
  uint32_t n_codes, n_post, gmax0, pad;
  const uint32_t* by_freq;   // code ids by descending initial count (ties: ascending id)
  const uint32_t* list_part; // [D] the partition all postings of the list lie in, or >= 2^31 (several)
  const uint32_t* cnt_ge;    // [gmax0 + 2] cnt_ge[f] = k-mers whose INITIAL count is >= f: only they can ever be at f
};
struct IncArgs {
  IncDir d[2];
  int ndirs; uint32_t slots, max_iter, mms, n_part, n_fp, p_words, uniform_parts;
  const uint16_t* seg_part;
  unsigned int* barrier;
};

__global__ void hist_init_kernel(const uint32_t* __restrict__ freq, uint32_t n_codes, uint32_t* hist) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_codes) atomicAdd(&hist[freq[c]], 1u);
}

__global__ void freq_key_kernel(const uint32_t* __restrict__ freq, uint32_t n_codes, uint64_t* __restrict__ key, uint32_t* __restrict__ val) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_codes) { key[c] = (uint64_t)(0xFFFFFFFFu - freq[c]); val[c] = c; }  // ascending key = descending count
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2)
greedy_incremental_kernel(const IncArgs A) {
  extern __shared__ __align__(16) unsigned char dsm[];
  constexpr int WARPS = THREADS / 32;
  uint32_t* seen = reinterpret_cast<uint32_t*>(dsm);
  uint32_t* fp = seen + A.p_words;
  unsigned long long* lst = reinterpret_cast<unsigned long long*>(dsm + (((size_t)(reinterpret_cast<unsigned char*>(fp + A.n_fp) - dsm) + 7) & ~(size_t)7));
  __shared__ uint32_t s_tied[2 * THREADS];
  __shared__ uint32_t s_cnt2[2], s_sc[2], s_g[2], s_nt[2], s_can[2], s_dec, s_red[2 * (THREADS / 32)];
  uint32_t* s_cov = reinterpret_cast<uint32_t*>(lst + A.n_fp);  // [2][n_fp] this iteration's partition_coverage
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // bit d = direction d is finished.  A register: a bool[2] indexed by a loop variable lives in LOCAL memory, and the
  // fence of every grid barrier invalidates L1, so each `if (DONE(d))` after a barrier was a round trip to L2.
  uint32_t donem = A.ndirs < 2 ? 2u : 0u;
#define DONE(d) ((donem >> (d)) & 1u)
  __shared__ unsigned long long live[2], evals[2];  // the lead thread's bookkeeping of reference-equivalent evals
  if (tid == 0) { live[0] = A.d[0].n_post; live[1] = A.ndirs > 1 ? A.d[1].n_post : 0ull; evals[0] = 0ull; evals[1] = 0ull; }
  if (tid == 0) { s_cnt2[0] = s_cnt2[1] = 0u; s_g[0] = A.d[0].gmax0; s_g[1] = A.ndirs > 1 ? A.d[1].gmax0 : 0u; s_dec = 0u; }
  __syncthreads();
  const bool lead = blockIdx.x == 0 && tid == 0;
  const bool tlead = blockIdx.x == (gridDim.x > 1 ? 1 : 0) && tid == 0;  // diagnostic clock of one block
  __shared__ unsigned long long s_key[2];
  __shared__ uint32_t s_npush[2];  // winners pushed (lead thread)
  __shared__ uint32_t s_reload[2]; // the last winner's postings span several partitions: re-read partition_coverage
  __shared__ uint32_t s_wa[2], s_wb[2], s_wp[2];  // the winner's posting range and list_part (fetched by one thread per direction)
  __shared__ unsigned long long s_t[6];  // 0 last stamp, 1 walk, 2 collect+score, 3 barrier 1, 4 apply, 5 barrier 2
  if (tid == 0) { for (int q = 0; q < 6; q++) s_t[q] = 0ull; s_npush[0] = 0u; s_npush[1] = 0u; s_reload[0] = 0u; s_reload[1] = 0u; }
  __syncthreads();
  if (tlead) s_t[0] = globaltimer_ns();
#define INC_STAMP(slot) if (tlead) { const unsigned long long t_ = globaltimer_ns(); s_t[slot] += t_ - s_t[0]; s_t[0] = t_; }
  for (uint32_t it = 0;; it++) {
    const int par = it & 1;
    // ---------------- phase 1 ----------------
    // The maximum never grows: walk down the histogram from the previous one, THREADS bins per step, both directions
    // in the same pass (early on the occupied bins are far apart: a serial walk would be hundreds of dependent L2 loads).
    // The thread that finds the maximum already holds hist[max] (the tie count) and has asked for cnt_ge[max] too.
    // In the same round trip: this iteration's partition_coverage tables into shared memory.
    // (the block's copy of partition_coverage is kept up to date from the winners; a multi-partition winner -- rare --
    // makes it re-read the table here)
    if (A.n_fp)
      for (int d = 0; d < A.ndirs; d++)
        if (!DONE(d) && (it == 0 || s_reload[d])) for (uint32_t q = tid; q < A.n_part; q += THREADS) s_cov[(size_t)d * A.n_fp + q] = __ldcg(A.d[d].cov + q);
    {
      uint32_t g[2] = {s_g[0], s_g[1]};
      bool open[2] = {!DONE(0), A.ndirs > 1 && !DONE(1)};
      __syncthreads();
      if (tid == 0) { s_reload[0] = 0u; s_reload[1] = 0u; }
      while (open[0] || open[1]) {
        uint32_t hv[2] = {0u, 0u}, cg_[2] = {0u, 0u}, mine[2] = {0u, 0u};
#pragma unroll
        for (int d = 0; d < 2; d++)
          if (open[d] && g[d] >= (uint32_t)tid && g[d] - (uint32_t)tid > 0u) {
            mine[d] = g[d] - (uint32_t)tid;                       // bins g, g-1, ..., g-THREADS+1
            hv[d] = __ldcg(A.d[d].hist + mine[d]);
            cg_[d] = __ldg(A.d[d].cnt_ge + mine[d]);
          }
#pragma unroll
        for (int d = 0; d < 2; d++) {
          const uint32_t wmax = __reduce_max_sync(0xffffffffu, hv[d] ? mine[d] : 0u);
          if (lane == 0) s_red[d * WARPS + warp] = wmax;
        }
        __syncthreads();
#pragma unroll
        for (int d = 0; d < 2; d++) {
          if (!open[d]) continue;
          uint32_t best = 0u;
#pragma unroll
          for (int w2 = 0; w2 < WARPS; w2++) best = max(best, s_red[d * WARPS + w2]);
          if (best || g[d] < (uint32_t)THREADS) {
            if (best ? (mine[d] == best && hv[d]) : tid == 0) { s_g[d] = best; s_nt[d] = best ? hv[d] : 0u; s_can[d] = best ? cg_[d] : 0u; }
            open[d] = false;
          } else {
            g[d] -= (uint32_t)THREADS;
          }
        }
        __syncthreads();
      }
    }
    INC_STAMP(1)
    for (int d = 0; d < A.ndirs; d++) {
      if (DONE(d)) continue;
      const IncDir& D = A.d[d];
      const uint32_t g = s_g[d];
      if (lead) evals[d] += live[d];                           // what the recount of this call would have examined
      if (g <= 1u) {  // None or freq == 1: stop before the push (main.rs:353-366)
        donem |= 1u << (d);
        if (lead) { D.ctl->iterations = it + 1; D.ctl->n_out = it; D.ctl->done = 1; D.ctl->evals = evals[d]; }
        continue;
      }
      // counts only fall: a k-mer can be at g only if its initial count was >= g, i.e. it sits in the first cnt_ge[g]
      // entries of the order by descending initial count -- a few thousand entries while g is large
      uint32_t* tied = s_tied + d * THREADS;
      const uint32_t stride = gridDim.x * THREADS, n_can = s_can[d];
      // entry x goes to block x % gridDim.x: k-mers of equal initial count are neighbours in this order, and so are
      // the tied ones -- dealt out round-robin every block scores about the same number of them
      for (uint32_t x = blockIdx.x + gridDim.x * (uint32_t)tid; x < n_can; x += stride) {
        const uint32_t c = __ldg(D.by_freq + x);
        if (__ldcg(D.freq + c) == g) { const uint32_t q = atomicAdd(&s_cnt2[d], 1u); if (q < (uint32_t)THREADS) tied[q] = c; }
      }
    }
    __syncthreads();
    // The usual case: the block's ties fit the lists and all of them lie in one partition each (list_part): score =
    // 0.0 + 1/(partition_coverage[p] + 1), one thread per tied k-mer, both directions in one round trip, one barrier.
    bool slow;
    {
      const uint32_t nh0 = DONE(0) ? 0u : s_cnt2[0], nh1 = (A.ndirs < 2 || DONE(1)) ? 0u : s_cnt2[1];
      const bool fit = nh0 <= (uint32_t)THREADS && nh1 <= (uint32_t)THREADS;
      const uint32_t c0 = fit && (uint32_t)tid < nh0 ? s_tied[tid] : 0xFFFFFFFFu;
      const uint32_t c1 = fit && (uint32_t)tid < nh1 ? s_tied[THREADS + tid] : 0xFFFFFFFFu;
      uint32_t lp0 = 0x80000000u, lp1 = 0x80000000u;
      if (c0 != 0xFFFFFFFFu) lp0 = __ldg(A.d[0].list_part + c0);
      if (c1 != 0xFFFFFFFFu) lp1 = __ldg(A.d[1].list_part + c1);
      if (!(lp0 >> 31)) {
        const uint32_t cv = A.n_fp ? s_cov[lp0] : __ldcg(A.d[0].cov + lp0);
        const float sc1 = __fdiv_rn(1.0f, __fadd_rn(__uint2float_rn(cv), 1.0f));
        atomicMax(&A.d[0].ctl->pk[par], ((unsigned long long)__float_as_uint(sc1) << 32) | (unsigned long long)(0xFFFFFFFFu - c0));
        s_tied[tid] = c0 | 0x80000000u;
      }
      if (!(lp1 >> 31)) {
        const uint32_t cv = A.n_fp ? s_cov[(size_t)A.n_fp + lp1] : __ldcg(A.d[1].cov + lp1);
        const float sc1 = __fdiv_rn(1.0f, __fadd_rn(__uint2float_rn(cv), 1.0f));
        atomicMax(&A.d[1].ctl->pk[par], ((unsigned long long)__float_as_uint(sc1) << 32) | (unsigned long long)(0xFFFFFFFFu - c1));
        s_tied[THREADS + tid] = c1 | 0x80000000u;
      }
      const bool mine_left = !fit || (c0 != 0xFFFFFFFFu && (lp0 >> 31)) || (c1 != 0xFFFFFFFFu && (lp1 >> 31));
      slow = __syncthreads_or(mine_left ? 1 : 0) != 0;
      if (!slow && tid == 0) { s_cnt2[0] = 0u; s_cnt2[1] = 0u; }
    }
    if (slow)
    for (int d = 0; d < A.ndirs; d++) {
      if (DONE(d)) continue;
      const IncDir& D = A.d[d];
      const uint32_t g = s_g[d];
      uint32_t* tied = s_tied + d * THREADS;
      const uint32_t n_here = s_cnt2[d];
      __syncthreads();
      if (tid == 0) s_cnt2[d] = 0u;
      const uint32_t n_chunks = n_here <= (uint32_t)THREADS ? 1u : (D.n_codes + gridDim.x * THREADS - 1u) / (gridDim.x * THREADS);
      for (uint32_t chunk = 0; chunk < n_chunks; chunk++) {
        uint32_t nt = n_here;
        if (n_here > (uint32_t)THREADS) {  // tie storm: re-collect chunk by chunk
          __syncthreads();
          const uint32_t x = chunk * gridDim.x * THREADS + blockIdx.x + gridDim.x * (uint32_t)tid;
          const uint32_t c = x < D.n_codes ? __ldg(D.by_freq + x) : 0u;
          if (x < s_can[d] && __ldcg(D.freq + c) == g) tied[atomicAdd(&s_cnt2[d], 1u)] = c;
          __syncthreads();
          nt = s_cnt2[d];
          __syncthreads();
          if (tid == 0) s_cnt2[d] = 0u;
        }
        // tie-storm chunks: single-partition lists as above (the usual case was scored and flagged before this loop)
        if (n_here > (uint32_t)THREADS && (uint32_t)tid < nt) {
          const uint32_t cc = tied[tid];
          const uint32_t lp = __ldg(D.list_part + cc);
          if (!(lp >> 31)) {
            const uint32_t cv = A.n_fp ? s_cov[(size_t)d * A.n_fp + lp] : __ldcg(D.cov + lp);
            const float sc1 = __fdiv_rn(1.0f, __fadd_rn(__uint2float_rn(cv), 1.0f));
            atomicMax(&D.ctl->pk[par], ((unsigned long long)__float_as_uint(sc1) << 32) | (unsigned long long)(0xFFFFFFFFu - cc));
            tied[tid] = cc | 0x80000000u;  // scored
          }
        }
        __syncthreads();
        for (uint32_t t = 0; t < nt; t++) {
          const uint32_t cc = tied[t];
          if (cc >> 31) continue;  // block-uniform: scored above
          float score;
          if (A.n_fp) {
            score = block_tie_score<false, false, THREADS>(cc, D.post_off, D.postings, D.ignored, A.seg_part, A.uniform_parts, s_cov + (size_t)d * A.n_fp, A.n_part, fp, lst, s_sc);
          } else {
            if (warp == 0) {
              const float sw = warp_tie_score<false, true>(cc, D.post_off, D.postings, D.ignored, A.seg_part, D.cov, seen, A.p_words, lane);
              if (lane == 0) s_sc[1] = __float_as_uint(sw);
            }
            __syncthreads();
            score = __uint_as_float(s_sc[1]);
          }
          if (tid == 0) atomicMax(&D.ctl->pk[par], ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xFFFFFFFFu - cc));
          __syncthreads();
        }
      }
    }
    INC_STAMP(2)
    grid_barrier_cg();
    INC_STAMP(3)
    // ---------------- phase 2 ----------------
    bool all_done = true;
    // one thread per direction fetches the winner's key (every warp of the grid asking for the same word is an L2 hot spot)
    // ... and the winner's posting range and list_part, so that 296 requests instead of 4736 go to those lines
    if (tid < 2 && tid < A.ndirs && !DONE(tid)) {
      const unsigned long long key = __ldcg(&A.d[tid].ctl->pk[par]);
      const uint32_t c = 0xFFFFFFFFu - (uint32_t)key;
      s_key[tid] = key;
      if (c < A.d[tid].n_codes) {
        s_wa[tid] = __ldg(A.d[tid].post_off + c); s_wb[tid] = __ldg(A.d[tid].post_off + c + 1u); s_wp[tid] = __ldg(A.d[tid].list_part + c);
      }
    }
    __syncthreads();
    for (int d = 0; d < A.ndirs; d++) {
      if (DONE(d)) continue;
      const IncDir& D = A.d[d];
      const uint32_t g = s_g[d];
      const unsigned long long key = s_key[d];
      const uint32_t c = 0xFFFFFFFFu - (uint32_t)key;
      if (lead) {
        msspe_candidate w;
        w.code = c; w.freq = g; w.n_tied = s_nt[d];  // the word is filled in after the loop (no dependent load here)
        s_npush[d] = it + 1u; w.tie_score = __uint_as_float((uint32_t)(key >> 32)); w.reserved = 0;
        D.out[it] = w;
        D.ctl->pk[par ^ 1] = 0ull; D.ctl->plive[par ^ 1] = 0u;  // next iteration's slots (nobody reads them in this phase)
      }
      if (g < A.mms || it + 1u >= A.max_iter) {  // main.rs:387-390 and the loop bound :344
        donem |= 1u << (d);
        if (lead) { D.ctl->iterations = it + 1; D.ctl->n_out = it + 1u; D.ctl->done = 1; D.ctl->evals = evals[d]; }
        continue;
      }
      all_done = false;
      if (A.n_fp && tid == 32 * d) {  // this block's partition_coverage copy
        const uint32_t lp = s_wp[d];
        if (!(lp >> 31)) s_cov[(size_t)d * A.n_fp + lp] += 1u; else s_reload[d] = 1u;
      }
      // main.rs:371-378 over the whole grid, then the decrements of the newly covered segments
      uint32_t* pmark = D.pmark + (size_t)par * 2048u;
      if (blockIdx.x == 0) for (uint32_t q = tid; q < 2048u; q += THREADS) D.pmark[(size_t)(par ^ 1) * 2048u + q] = 0u;
      const uint32_t a = s_wa[d], b = s_wb[d];
      uint32_t dec = 0;
      // one warp per posting, dealt out over the whole grid: lane 0 marks the segment (bitmask, partition_coverage); if it
      // was not covered before, the 32 lanes take its k-mers (forward index): each loses one live segment
      for (uint32_t i = a + blockIdx.x * WARPS + warp; i < b; i += gridDim.x * WARPS) {
        const uint32_t sg = __ldg(D.postings + i);
        // the segment's first 64 forward-index entries are requested before the bitmask answer is known
        const uint32_t* fw = D.fwd_ids + (uint64_t)sg * A.slots;
        const uint32_t id0 = (uint32_t)lane < A.slots ? __ldg(fw + lane) : 0xFFFFFFFFu;
        const uint32_t id1 = (uint32_t)lane + 32u < A.slots ? __ldg(fw + lane + 32) : 0xFFFFFFFFu;
        uint32_t newly = 0u;
        if (lane == 0) {
          const uint32_t bit = 1u << (sg & 31u);
          const uint32_t old = atomicOr(&D.ignored[sg >> 5], bit);
          const uint32_t p = partition_of(A.seg_part, A.uniform_parts, sg);
          const uint32_t pbit = 1u << (p & 31u);
          const uint32_t pold = atomicOr(&pmark[p >> 5], pbit);
          if (!(pold & pbit)) atomicAdd(&D.cov[p], 1u);
          newly = (old & bit) ? 0u : 1u;
        }
        if (__shfl_sync(0xffffffffu, newly, 0)) {
          uint32_t f0 = 0u, f1 = 0u;
          if (id0 != 0xFFFFFFFFu) f0 = atomicSub(&D.freq[id0], 1u);
          if (id1 != 0xFFFFFFFFu) f1 = atomicSub(&D.freq[id1], 1u);
          // neighbouring k-mers of a window mostly have the same count: one histogram update per distinct count
          // of the warp instead of one per k-mer (late in a run all counts fall into a handful of bins)
          {
            const bool on0 = id0 != 0xFFFFFFFFu, on1 = id1 != 0xFFFFFFFFu;
            const unsigned g0 = __match_any_sync(0xffffffffu, on0 ? f0 : 0xFFFFFF00u + (uint32_t)lane);
            if (on0 && lane == __ffs(g0) - 1) { const uint32_t n = (uint32_t)__popc(g0); atomicSub(&D.hist[f0], n); atomicAdd(&D.hist[f0 - 1u], n); }
            if (A.slots > 32u) {
              const unsigned g1 = __match_any_sync(0xffffffffu, on1 ? f1 : 0xFFFFFF00u + (uint32_t)lane);
              if (on1 && lane == __ffs(g1) - 1) { const uint32_t n = (uint32_t)__popc(g1); atomicSub(&D.hist[f1], n); atomicAdd(&D.hist[f1 - 1u], n); }
            }
            dec += (on0 ? 1u : 0u) + (on1 ? 1u : 0u);
          }
          for (uint32_t q = lane + 64u; q < A.slots; q += 32) {
            const uint32_t id = __ldg(fw + q);
            if (id != 0xFFFFFFFFu) {
              const uint32_t f = atomicSub(&D.freq[id], 1u);
              atomicSub(&D.hist[f], 1u);
              atomicAdd(&D.hist[f - 1u], 1u);
              dec++;
            }
          }
        }
      }
      dec = __reduce_add_sync(0xffffffffu, dec);
      if (lane == 0 && dec) atomicAdd(&s_dec, dec);
      __syncthreads();
      if (tid == 0) { if (s_dec) atomicAdd(&D.ctl->plive[par], s_dec); s_dec = 0u; }
      __syncthreads();
    }
    INC_STAMP(4)
    if (all_done) break;
    grid_barrier_cg();
    INC_STAMP(5)
    if (lead)  // only the lead thread reports evals
      for (int d = 0; d < A.ndirs; d++)
        if (!DONE(d)) live[d] -= (unsigned long long)__ldcg(&A.d[d].ctl->plive[par]);
  }
#undef INC_STAMP
  if (blockIdx.x == 0) {  // winners: code id -> word
    __syncthreads();
    for (int d = 0; d < A.ndirs; d++) {
      const uint32_t n = s_npush[d];
      for (uint32_t i = tid; i < n; i += THREADS) A.d[d].out[i].code = A.d[d].codes[(uint32_t)A.d[d].out[i].code];
    }
  }
  if (tlead) for (int q = 0; q < 5; q++) A.d[0].ctl->t_dbg[q] = s_t[q + 1];
}
#undef DONE

int run_select_incremental(msspe_ctx* c, int ndirs, const int* dirs, uint32_t max_iter, uint32_t mms,
                           msspe_candidate** outs, uint32_t** n_outs) {
  cudaStream_t st = c->stream;
  const uint64_t G = c->n_segments;
  IncArgs A{};
  A.ndirs = ndirs; A.slots = c->slots; A.max_iter = max_iter; A.mms = mms; A.seg_part = c->d_seg_part;
  A.uniform_parts = (c->uniform_parts && c->uniform_parts <= 65536u) ? c->uniform_parts : 0u;
  A.n_part = c->max_partition + 1u;
  A.n_fp = A.n_part <= 4096u ? A.n_part : 0u;
  A.p_words = (c->max_partition + 32u) / 32u;
  const uint32_t mask_words = (uint32_t)div_up_u64(G, 32) + 1u;
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[2], st));
  uint32_t* d_hist[2] = {nullptr, nullptr}; uint32_t* d_pmark[2] = {nullptr, nullptr}; unsigned int* d_bar = nullptr;
  uint32_t* d_byfreq[2] = {nullptr, nullptr}; uint32_t* d_cntge[2] = {nullptr, nullptr};
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_bar, 4, st));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(d_bar, 0, 4, st));
  A.barrier = d_bar;
  for (int i = 0; i < ndirs; i++) {
    DirIndex& D = c->dir[dirs[i]];
    if (D.out_capacity < max_iter) {
      msspe_dev_free(c, D.out); D.out = nullptr;
      MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.out, (uint64_t)max_iter * sizeof(msspe_candidate), c->stream));
      D.out_capacity = max_iter;
    }
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.ignored, 0, (size_t)mask_words * 4, st));
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.cov, 0, 65536 * 4, st));
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.ctl, 0, sizeof(SelectCtl), st));
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.acc, 0, (uint64_t)(D.n_tiles + 1) * 8, st));
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.freq, 0, (D.n_codes ? D.n_codes : 1) * 4, st));
    // exact initial counts: one recount of the complete CSR against the empty bitmask
    if (D.n_tiles) {
      const unsigned cap = (unsigned)c->sm_count * 2u;
      const unsigned blocks = (unsigned)div_up_u64(D.n_tiles, CNT_THREADS / 32);
      count_kernel<false><<<blocks < cap ? blocks : cap, CNT_THREADS, 0, st>>>(D.postings, D.post_off, D.tile_first, (uint32_t)D.n_codes, (uint32_t)D.n_records,
                                                                              D.n_tiles, D.ignored, mask_words, D.freq, D.acc, D.ctl);
      c->timing.kernel_launches++;
    }
    MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&c->h_ctl[i], D.ctl, sizeof(SelectCtl), cudaMemcpyDeviceToHost, st));
  }
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  for (int i = 0; i < ndirs; i++) {
    DirIndex& D = c->dir[dirs[i]];
    const uint32_t gmax0 = c->h_ctl[i].gmax;
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_hist[i], ((uint64_t)gmax0 + 2) * 4, st));
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(d_hist[i], 0, ((uint64_t)gmax0 + 2) * 4, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_pmark[i], 2 * 2048 * 4, st));
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(d_pmark[i], 0, 2 * 2048 * 4, st));
    if (D.n_codes) { hist_init_kernel<<<(unsigned)div_up_u64(D.n_codes, 256), 256, 0, st>>>(D.freq, (uint32_t)D.n_codes, d_hist[i]); c->timing.kernel_launches++; }
    {  // k-mers by descending initial count, and how many of them start at or above each count
      const uint64_t nc = D.n_codes ? D.n_codes : 1;
      uint64_t *ka = nullptr, *kb = nullptr; uint32_t *va = nullptr, *vb = nullptr;
      MSSPE_CUDA_TRY(c, cudaMallocAsync(&ka, nc * 8, st)); MSSPE_CUDA_TRY(c, cudaMallocAsync(&kb, nc * 8, st));
      MSSPE_CUDA_TRY(c, cudaMallocAsync(&va, nc * 4, st)); MSSPE_CUDA_TRY(c, cudaMallocAsync(&vb, nc * 4, st));
      if (D.n_codes) {
        freq_key_kernel<<<(unsigned)div_up_u64(D.n_codes, 256), 256, 0, st>>>(D.freq, (uint32_t)D.n_codes, ka, va);
        c->timing.kernel_launches++;
        int rc = msspe_radix_sort_pairs(c, &ka, &va, &kb, &vb, D.n_codes, 32, st);
        if (rc) return rc;
      }
      d_byfreq[i] = va;
      MSSPE_CUDA_TRY(c, cudaFreeAsync(ka, st)); MSSPE_CUDA_TRY(c, cudaFreeAsync(kb, st)); MSSPE_CUDA_TRY(c, cudaFreeAsync(vb, st));
      std::vector<uint32_t> hh((size_t)gmax0 + 2);
      MSSPE_CUDA_TRY(c, cudaMemcpyAsync(hh.data(), d_hist[i], hh.size() * 4, cudaMemcpyDeviceToHost, st));
      MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
      uint32_t run = 0;
      for (size_t f = hh.size(); f-- > 0;) { run += hh[f]; hh[f] = run; }
      MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_cntge[i], hh.size() * 4, st));
      MSSPE_CUDA_TRY(c, cudaMemcpyAsync(d_cntge[i], hh.data(), hh.size() * 4, cudaMemcpyHostToDevice, st));
      MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));  // hh goes out of scope
    }
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.ctl, 0, sizeof(SelectCtl), st));
    IncDir& g = A.d[i];
    g.post_off = D.post_off; g.postings = D.postings; g.codes = D.codes; g.fwd_ids = D.fwd_ids; g.freq = D.freq; g.hist = d_hist[i];
    g.ignored = D.ignored; g.cov = D.cov; g.pmark = d_pmark[i]; g.ctl = D.ctl; g.out = D.out;
    g.n_codes = (uint32_t)D.n_codes; g.n_post = (uint32_t)D.n_records; g.gmax0 = gmax0;
    g.by_freq = d_byfreq[i]; g.cnt_ge = d_cntge[i]; g.list_part = D.list_part;
  }
  const size_t smem = (size_t)A.p_words * 4 + (size_t)A.n_fp * 4 + (size_t)A.n_fp * 8 + (size_t)2 * A.n_fp * 4 + 32;
  void* fn = (void*)greedy_incremental_kernel<512>;
  MSSPE_CUDA_TRY(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  MSSPE_CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&per_sm, fn, 512, smem, cudaOccupancyDefault));
  if (per_sm < 1) { c->set_error("msspe_select: incremental kernel does not fit on an SM"); return MSSPE_ERR_CAPACITY; }
  if (per_sm > 2) per_sm = 2;
  void* kargs[] = {(void*)&A};
  { KPROF(c, KP_GREEDY_WHOLE, st, 0) MSSPE_CUDA_TRY(c, cudaLaunchCooperativeKernel(fn, dim3((unsigned)c->sm_count * (unsigned)per_sm), dim3(512), kargs, smem, st)); }
  for (int i = 0; i < ndirs; i++)
    MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&c->h_ctl[i], c->dir[dirs[i]].ctl, sizeof(SelectCtl), cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[3], st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  float ms = 0.f;
  MSSPE_CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]));
  for (int i = 0; i < ndirs; i++) {
    const int d = dirs[i];
    const uint32_t n = c->h_ctl[i].n_out;
    if (n) MSSPE_CUDA_TRY(c, cudaMemcpyAsync(outs[i], c->dir[d].out, (size_t)n * sizeof(msspe_candidate), cudaMemcpyDeviceToHost, st));
    *n_outs[i] = n;
    c->timing.select_ms[d] = ms;
    c->timing.select_evals[d] = c->h_ctl[i].evals;
    c->timing.select_iterations[d] = c->h_ctl[i].iterations;
    c->timing.select_postings_read[d] = c->dir[d].n_records;  // the one initial recount
    c->timing.count_kernel_launches[d] = 1;
    c->timing.count_kernel_ms[d] = 0.f;
    if (getenv("MSSPE_DEBUG_TIMERS") && i == 0)
      fprintf(stderr, "[msspe] incremental greedy (%u iterations, wall %.3f ms incl. set-up), one block's clock: histogram walk %.3f | collect+score %.3f | barrier %.3f | apply+decrement %.3f | barrier %.3f ms\n",
              c->h_ctl[i].iterations, ms, c->h_ctl[i].t_dbg[0] * 1e-6, c->h_ctl[i].t_dbg[1] * 1e-6, c->h_ctl[i].t_dbg[2] * 1e-6, c->h_ctl[i].t_dbg[3] * 1e-6, c->h_ctl[i].t_dbg[4] * 1e-6);
    MSSPE_CUDA_TRY(c, cudaFreeAsync(d_hist[i], st));
    MSSPE_CUDA_TRY(c, cudaFreeAsync(d_pmark[i], st));
    MSSPE_CUDA_TRY(c, cudaFreeAsync(d_byfreq[i], st));
    MSSPE_CUDA_TRY(c, cudaFreeAsync(d_cntge[i], st));
  }
  MSSPE_CUDA_TRY(c, cudaFreeAsync(d_bar, st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  return MSSPE_OK;
}

int run_select(msspe_ctx* c, int ndirs, const int* dirs, uint32_t max_iter, uint32_t mms, uint32_t mode,
               msspe_candidate** outs, uint32_t** n_outs) {
  if (!c->built) { c->set_error("msspe_select: index not built"); return MSSPE_ERR_STATE; }
  const bool batched = (mode & MSSPE_SELECT_BATCHED) != 0;
  mode &= ~(uint32_t)MSSPE_SELECT_BATCHED;
  if (mode > MSSPE_SELECT_PARTITIONED) { c->set_error("msspe_select: unknown mode %u", mode); return MSSPE_ERR_INVALID; }
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  // AUTO: the per-partition decomposition wherever (almost) every list lies in one partition (any pre-aligned input);
  // measured on B200 against the two whole-index loops in DESIGN.md
  if (mode == MSSPE_SELECT_AUTO && !batched && !getenv("MSSPE_AUTO_NO_PARTITIONED") && msspe_partitioned_applicable(c, ndirs, dirs, max_iter))
    mode = MSSPE_SELECT_PARTITIONED;
  if (mode == MSSPE_SELECT_PARTITIONED) return msspe_select_partitioned(c, ndirs, dirs, max_iter, mms, outs, n_outs);
  if (mode == MSSPE_SELECT_AUTO) {  // measured on B200: incremental ahead at cfg2 (4.5 M postings per direction: 8.6 vs 9.4 ms), cfg3 (16.8 vs
                                    // 23.4 ms) and beyond; below that its set-up (order by initial count, histogram) is what counts
    uint64_t most = 0;
    for (int i = 0; i < ndirs; i++) most = std::max<uint64_t>(most, c->dir[dirs[i]].n_records);
    mode = most >= (1ull << 21) ? MSSPE_SELECT_INCREMENTAL : MSSPE_SELECT_RECOUNT;
  }
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  if (!batched && mode == MSSPE_SELECT_INCREMENTAL && max_iter > 0) return run_select_incremental(c, ndirs, dirs, max_iter, mms, outs, n_outs);
  if (!batched && mode == MSSPE_SELECT_RECOUNT && max_iter > 0) {
    // both covered-segment bitmasks in shared memory if they fit; otherwise one direction per launch
    const size_t one_mask = ((size_t)div_up_u64(c->n_segments, 32) + 1) * 4;
    // a bitmask that does not fit shared memory would have to be gathered from L2 for every posting of every recount
    // (measured 0.9 TB/s): the incremental kernel returns the identical result without streaming at all
    if (one_mask + 16384 > c->smem_optin && !getenv("MSSPE_FORCE_RECOUNT")) return run_select_incremental(c, ndirs, dirs, max_iter, mms, outs, n_outs);
    if (ndirs == 2 && 2 * one_mask + 16384 > c->smem_optin && one_mask + 16384 <= c->smem_optin) {
      for (int i = 0; i < 2; i++) {
        int rc = run_select_persistent(c, 1, &dirs[i], max_iter, mms, &outs[i], &n_outs[i]);
        if (rc) return rc;
      }
      return MSSPE_OK;
    }
    return run_select_persistent(c, ndirs, dirs, max_iter, mms, outs, n_outs);
  }
  DirRun runs[2];
  cudaStream_t streams[2] = {c->stream, c->stream2};
  cudaEvent_t ev_b[2] = {c->ev[2], c->ev[4]}, ev_e[2] = {c->ev[3], c->ev[5]};
  if (ndirs == 2) {  // fork the second stream off the main one
    MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev_fork, c->stream));
    MSSPE_CUDA_TRY(c, cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
  }
  for (int i = 0; i < ndirs; i++) {
    MSSPE_CUDA_TRY(c, cudaEventRecord(ev_b[i], streams[i]));
    int rc = prepare_dir(c, dirs[i], max_iter, streams[i], &runs[i]);
    if (rc) return rc;
  }
  std::vector<cudaEvent_t> pev[2];
  bool finished[2] = {max_iter == 0, max_iter == 0};
  uint32_t issued = 0;
  const uint32_t BATCH = 32;
  while (issued < max_iter && !(finished[0] && (ndirs == 1 || finished[1]))) {
    const uint32_t nb = (max_iter - issued) < BATCH ? (max_iter - issued) : BATCH;
    for (uint32_t it = 0; it < nb; it++)
      for (int i = 0; i < ndirs; i++)
        if (!finished[i]) launch_iteration(c, runs[i], max_iter, mms, mode, issued + it == 0, c->profiling ? &pev[i] : nullptr);
    issued += nb;
    for (int i = 0; i < ndirs; i++)
      if (!finished[i]) MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&c->h_ctl[i], runs[i].D->ctl, sizeof(SelectCtl), cudaMemcpyDeviceToHost, streams[i]));
    for (int i = 0; i < ndirs; i++)
      if (!finished[i]) {
        MSSPE_CUDA_TRY(c, cudaStreamSynchronize(streams[i]));
        if (c->h_ctl[i].done) finished[i] = true;
      }
  }
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  for (int i = 0; i < ndirs; i++) {
    MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&c->h_ctl[i], runs[i].D->ctl, sizeof(SelectCtl), cudaMemcpyDeviceToHost, streams[i]));
    MSSPE_CUDA_TRY(c, cudaEventRecord(ev_e[i], streams[i]));
    MSSPE_CUDA_TRY(c, cudaStreamSynchronize(streams[i]));
    const uint32_t n = c->h_ctl[i].n_out;
    if (n) MSSPE_CUDA_TRY(c, cudaMemcpy(outs[i], runs[i].D->out, (size_t)n * sizeof(msspe_candidate), cudaMemcpyDeviceToHost));
    *n_outs[i] = n;
    const int d = dirs[i];
    MSSPE_CUDA_TRY(c, cudaEventElapsedTime(&c->timing.select_ms[d], ev_b[i], ev_e[i]));
    c->timing.select_evals[d] = c->h_ctl[i].evals;
    c->timing.select_iterations[d] = c->h_ctl[i].iterations;
    const uint64_t count_launches = mode == MSSPE_SELECT_RECOUNT ? c->h_ctl[i].iterations : (c->h_ctl[i].iterations ? 1 : 0);
    c->timing.select_postings_read[d] = count_launches * runs[i].D->n_records;
    c->timing.count_kernel_launches[d] = (uint32_t)count_launches;
    float sum_ms = 0.f;
    for (size_t e = 0; e + 1 < pev[i].size(); e += 2) {
      float ms = 0.f;
      if (e / 2 < count_launches && cudaEventElapsedTime(&ms, pev[i][e], pev[i][e + 1]) == cudaSuccess) sum_ms += ms;
      cudaEventDestroy(pev[i][e]); cudaEventDestroy(pev[i][e + 1]);
    }
    c->timing.count_kernel_ms[d] = sum_ms;
  }
  if (ndirs == 2) {  // join
    MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev_join, c->stream2));
    MSSPE_CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  }
  return MSSPE_OK;
}

}  // namespace

int msspe_select_prepare_static(msspe_ctx* c, int dir, cudaStream_t st) {
  DirIndex& D = c->dir[dir];
  D.n_tiles = (uint32_t)div_up_u64(D.n_records, CNT_TILE);
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.acc, (uint64_t)(D.n_tiles + 1) * 8, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.tile_first, (uint64_t)(D.n_tiles + 1) * 4, c->stream));
  if (D.n_tiles) {
    tile_first_kernel<<<(D.n_tiles + 255) / 256, 256, 0, st>>>(D.post_off, (uint32_t)D.n_codes, D.n_tiles, D.tile_first);
    c->timing.kernel_launches++;
    MSSPE_CUDA_TRY(c, cudaGetLastError());
  }
  return MSSPE_OK;
}

int msspe_select_prepare_stream(msspe_ctx* c, int dir, cudaStream_t st) {
  DirIndex& D = c->dir[dir];
  if (D.stream_built) return MSSPE_OK;
  const uint32_t nc = (uint32_t)D.n_codes, R = (uint32_t)D.n_records;
  uint32_t* multi = nullptr; uint32_t* mlen = nullptr; uint32_t* d_tot = nullptr;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&multi, ((uint64_t)nc + 1) * 4, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&mlen, ((uint64_t)nc + 1) * 4, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_tot, 8, st));
  uint32_t tot[2] = {0, 0};
  if (nc) {
    stream_flag_kernel<<<(nc + 255u) / 256u, 256, 0, st>>>(D.post_off, nc, multi, mlen);
    c->timing.kernel_launches++;
    int rc = msspe_exclusive_scan_u32(c, multi, multi, nc, d_tot, st);
    if (rc) return rc;
    rc = msspe_exclusive_scan_u32(c, mlen, mlen, nc, d_tot + 1, st);
    if (rc) return rc;
    MSSPE_CUDA_TRY(c, cudaMemcpyAsync(tot, d_tot, 8, cudaMemcpyDeviceToHost, st));
    MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  }
  D.s_lists = tot[0]; D.s_list_post = tot[1];
  D.s_tiles = (uint32_t)div_up_u64(D.s_list_post, CNT_TILE);
  const uint32_t tail_begin = D.s_tiles * (uint32_t)CNT_TILE;  // the tail starts on a tile boundary
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.s_postings, ((uint64_t)R + CNT_TILE) * 4, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.s_off, ((uint64_t)D.s_lists + 1) * 4, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.s_id, ((uint64_t)D.s_lists + 1) * 16, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.s_tile_first, ((uint64_t)D.s_tiles + 1) * 4, c->stream));
  if (nc) {
    stream_copy_kernel<<<(unsigned)div_up_u64(nc, 256), 256, 0, st>>>(D.post_off, D.postings, nc, multi, mlen, D.s_lists, D.s_list_post,
                                                                      tail_begin, D.list_part, D.s_postings, D.s_off, D.s_id);
    c->timing.kernel_launches++;
  } else {
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.s_off, 0, 4, st));
  }
  if (D.s_tiles) {
    tile_first_kernel<<<(D.s_tiles + 255) / 256, 256, 0, st>>>(D.s_off, D.s_lists, D.s_tiles, D.s_tile_first);
    c->timing.kernel_launches++;
  }
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  {
    const uint32_t tail_end = D.s_tiles * (uint32_t)CNT_TILE + (R - D.s_list_post);
    int rc = build_cost_prefix(c, D.s_tile_first, D.s_tiles, D.s_lists, tail_end, &D.s_cost, st);
    if (rc) return rc;
  }
  MSSPE_CUDA_TRY(c, cudaFreeAsync(multi, st));
  MSSPE_CUDA_TRY(c, cudaFreeAsync(mlen, st));
  MSSPE_CUDA_TRY(c, cudaFreeAsync(d_tot, st));
  D.stream_built = true;
  return MSSPE_OK;
}

extern "C" int msspe_select(msspe_ctx* c, uint8_t dir, uint32_t max_iterations, uint32_t mms, uint32_t mode,
                            msspe_candidate* out, uint32_t* n_out) {
  if (!c) return MSSPE_ERR_INVALID;
  if (dir > 1 || !n_out || (max_iterations && !out)) { c->set_error("msspe_select: bad argument"); return MSSPE_ERR_INVALID; }
  int dirs[1] = {dir};
  msspe_candidate* outs[1] = {out};
  uint32_t* ns[1] = {n_out};
  *n_out = 0;
  return run_select(c, 1, dirs, max_iterations, mms, mode, outs, ns);
}

extern "C" int msspe_select_both(msspe_ctx* c, uint32_t max_iterations, uint32_t mms, uint32_t mode,
                                 msspe_candidate* out_fwd, uint32_t* n_fwd, msspe_candidate* out_rev, uint32_t* n_rev) {
  if (!c) return MSSPE_ERR_INVALID;
  if (!n_fwd || !n_rev || (max_iterations && (!out_fwd || !out_rev))) { c->set_error("msspe_select_both: bad argument"); return MSSPE_ERR_INVALID; }
  int dirs[2] = {0, 1};
  msspe_candidate* outs[2] = {out_fwd, out_rev};
  uint32_t* ns[2] = {n_fwd, n_rev};
  *n_fwd = *n_rev = 0;
  return run_select(c, 2, dirs, max_iterations, mms, mode, outs, ns);
}

// ---- coverage of the final primer set: device side of print_coverage_report (main.rs:518-537) ----
namespace {
__global__ void coverage_mark_kernel(const uint64_t* __restrict__ sel, uint32_t n_sel, const uint64_t* __restrict__ codes,
                                     uint32_t n_codes, const uint32_t* __restrict__ post_off,
                                     const uint32_t* __restrict__ postings, uint8_t* __restrict__ covered) {
  const uint32_t s = blockIdx.x;
  if (s >= n_sel || n_codes == 0) return;
  const uint64_t want = sel[s];
  uint32_t lo = 0, hi = n_codes;  // first index with codes[idx] >= want
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (codes[mid] < want) lo = mid + 1; else hi = mid; }
  if (lo >= n_codes || codes[lo] != want) return;
  for (uint32_t i = post_off[lo] + threadIdx.x; i < post_off[lo + 1]; i += blockDim.x) covered[postings[i]] = 1;
}
}  // namespace

extern "C" int msspe_coverage(msspe_ctx* c, const uint64_t* fwd_codes, uint32_t n_fwd, const uint64_t* rev_codes, uint32_t n_rev,
                              uint8_t* covered, uint16_t* partition_no, uint32_t* record_of_segment, uint64_t capacity) {
  if (!c) return MSSPE_ERR_INVALID;
  if (!c->built) { c->set_error("msspe_coverage: index not built"); return MSSPE_ERR_STATE; }
  const uint64_t G = c->n_segments;
  if (capacity < G || (G && !covered)) { c->set_error("msspe_coverage: need capacity for %llu segments", (unsigned long long)G); return MSSPE_ERR_CAPACITY; }
  if ((n_fwd && !fwd_codes) || (n_rev && !rev_codes)) { c->set_error("msspe_coverage: null argument"); return MSSPE_ERR_INVALID; }
  if (G == 0) return MSSPE_OK;
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  uint8_t* d_cov = nullptr; uint64_t* d_sel = nullptr;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_cov, G, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_sel, (size_t)(n_fwd + n_rev + 1) * 8, st));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(d_cov, 0, G, st));
  if (n_fwd) MSSPE_CUDA_TRY(c, cudaMemcpyAsync(d_sel, fwd_codes, (size_t)n_fwd * 8, cudaMemcpyHostToDevice, st));
  if (n_rev) MSSPE_CUDA_TRY(c, cudaMemcpyAsync(d_sel + n_fwd, rev_codes, (size_t)n_rev * 8, cudaMemcpyHostToDevice, st));
  for (int d = 0; d < 2; d++) {
    const uint32_t ns = d == 0 ? n_fwd : n_rev;
    if (!ns) continue;
    DirIndex& D = c->dir[d];
    coverage_mark_kernel<<<ns, 256, 0, st>>>(d_sel + (d == 0 ? 0 : n_fwd), ns, D.codes, (uint32_t)D.n_codes, D.post_off, D.postings, d_cov);
    c->timing.kernel_launches++;
  }
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(covered, d_cov, G, cudaMemcpyDeviceToHost, st));
  if (partition_no) MSSPE_CUDA_TRY(c, cudaMemcpyAsync(partition_no, c->d_seg_part, G * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
  if (record_of_segment) MSSPE_CUDA_TRY(c, cudaMemcpyAsync(record_of_segment, c->d_seg_rec, G * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaFreeAsync(d_cov, st));
  MSSPE_CUDA_TRY(c, cudaFreeAsync(d_sel, st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  return MSSPE_OK;
}


// ---- genome-sharded selection: per-rank device work of one iteration (collectives live above the ABI) ----
namespace {
__global__ void __launch_bounds__(256)
shard_firstpos_kernel(const uint32_t* __restrict__ ids, uint32_t n, uint32_t n_part, const uint32_t* __restrict__ post_off,
                      const uint32_t* __restrict__ postings, const uint32_t* __restrict__ ignored,
                      const uint16_t* __restrict__ seg_part, uint32_t* __restrict__ first_pos) {
  const uint32_t t = blockIdx.x;
  if (t >= n) return;
  uint32_t* fp = first_pos + (size_t)t * n_part;
  for (uint32_t p = threadIdx.x; p < n_part; p += blockDim.x) fp[p] = 0xFFFFFFFFu;
  __syncthreads();
  const uint32_t c = ids[t];
  if (c == MSSPE_NO_LOCAL_ID) return;
  const uint32_t a = post_off[c], b = post_off[c + 1];
  for (uint32_t i = a + threadIdx.x; i < b; i += blockDim.x) {
    const uint32_t seg = postings[i];
    if (!((ignored[seg >> 5] >> (seg & 31u)) & 1u)) atomicMin(&fp[seg_part[seg]], i - a);
  }
}

__global__ void __launch_bounds__(1024)
shard_apply_kernel(uint32_t c, const uint32_t* __restrict__ post_off, const uint32_t* __restrict__ postings, uint32_t* ignored,
                   const uint16_t* __restrict__ seg_part, uint8_t* part_flags) {
  const uint32_t a = post_off[c], b = post_off[c + 1];
  for (uint32_t i = a + threadIdx.x; i < b; i += blockDim.x) {
    const uint32_t seg = postings[i];
    atomicOr(&ignored[seg >> 5], 1u << (seg & 31u));
    part_flags[seg_part[seg]] = 1;
  }
}
}  // namespace

extern "C" int msspe_shard_begin(msspe_ctx* c, uint8_t dir) {
  if (!c || dir > 1) return MSSPE_ERR_INVALID;
  if (!c->built) { c->set_error("msspe_shard_begin: index not built"); return MSSPE_ERR_STATE; }
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  DirIndex& D = c->dir[dir];
  const size_t mask_words = div_up_u64(c->n_segments, 32) + 1;
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.ignored, 0, mask_words * 4, c->stream));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.acc, 0, (uint64_t)(D.n_tiles + 1) * 8, c->stream));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.freq, 0, (D.n_codes ? D.n_codes : 1) * 4, c->stream));
  return MSSPE_OK;
}

extern "C" int msspe_shard_buffers(msspe_ctx* c, uint8_t dir, const uint64_t** d_codes, const uint32_t** d_freq, uint64_t* n_codes) {
  if (!c || dir > 1) return MSSPE_ERR_INVALID;
  if (!c->built) { c->set_error("msspe_shard_buffers: index not built"); return MSSPE_ERR_STATE; }
  if (d_codes) *d_codes = c->dir[dir].codes;
  if (d_freq) *d_freq = c->dir[dir].freq;
  if (n_codes) *n_codes = c->dir[dir].n_codes;
  return MSSPE_OK;
}

extern "C" int msspe_shard_count(msspe_ctx* c, uint8_t dir, uint64_t* live) {
  if (!c || dir > 1) return MSSPE_ERR_INVALID;
  if (!c->built) { c->set_error("msspe_shard_count: index not built"); return MSSPE_ERR_STATE; }
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  DirIndex& D = c->dir[dir];
  cudaStream_t st = c->stream;
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.ctl, 0, sizeof(SelectCtl), st));
  if (D.n_tiles) {
    const uint32_t mask_words = (uint32_t)div_up_u64(c->n_segments, 32) + 1u;
    const size_t mask_bytes = (size_t)mask_words * 4;
    const bool smem_mask = mask_bytes + 8192 <= c->smem_optin;
    constexpr int WARPS = CNT_THREADS / 32;
    const unsigned blocks = (unsigned)div_up_u64(D.n_tiles, WARPS);
    const unsigned cap = (unsigned)c->sm_count * (smem_mask && mask_bytes > 100 * 1024 ? 1u : 2u);
    const unsigned grid = blocks < cap ? blocks : cap;
    if (smem_mask) {
      MSSPE_CUDA_TRY(c, cudaFuncSetAttribute(count_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mask_bytes));
      count_kernel<true><<<grid, CNT_THREADS, mask_bytes, st>>>(D.postings, D.post_off, D.tile_first, (uint32_t)D.n_codes, (uint32_t)D.n_records,
                                                                D.n_tiles, D.ignored, mask_words, D.freq, D.acc, D.ctl);
    } else {
      count_kernel<false><<<grid, CNT_THREADS, 0, st>>>(D.postings, D.post_off, D.tile_first, (uint32_t)D.n_codes, (uint32_t)D.n_records,
                                                        D.n_tiles, D.ignored, mask_words, D.freq, D.acc, D.ctl);
    }
    c->timing.kernel_launches++;
    MSSPE_CUDA_TRY(c, cudaGetLastError());
  }
  if (live) {
    MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&c->h_ctl[dir], D.ctl, sizeof(SelectCtl), cudaMemcpyDeviceToHost, st));
    MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
    *live = c->h_ctl[dir].evals;
  }
  return MSSPE_OK;
}

extern "C" int msspe_shard_firstpos(msspe_ctx* c, uint8_t dir, const uint32_t* local_ids, uint32_t n, uint32_t n_part, uint32_t* first_pos) {
  if (!c || dir > 1 || (n && (!local_ids || !first_pos))) return MSSPE_ERR_INVALID;
  if (!c->built) { c->set_error("msspe_shard_firstpos: index not built"); return MSSPE_ERR_STATE; }
  if (n_part <= c->max_partition && c->n_segments) { c->set_error("msspe_shard_firstpos: n_part %u <= max partition %u", n_part, c->max_partition); return MSSPE_ERR_INVALID; }
  if (n == 0) return MSSPE_OK;
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  DirIndex& D = c->dir[dir];
  for (uint32_t i = 0; i < n; i++)
    if (local_ids[i] != MSSPE_NO_LOCAL_ID && local_ids[i] >= D.n_codes) { c->set_error("msspe_shard_firstpos: id %u out of range", local_ids[i]); return MSSPE_ERR_INVALID; }
  cudaStream_t st = c->stream;
  uint32_t* d_ids = nullptr; uint32_t* d_fp = nullptr;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_ids, (size_t)n * 4, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_fp, (size_t)n * n_part * 4, st));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(d_ids, local_ids, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  shard_firstpos_kernel<<<n, 256, 0, st>>>(d_ids, n, n_part, D.post_off, D.postings, D.ignored, c->d_seg_part, d_fp);
  c->timing.kernel_launches++;
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(first_pos, d_fp, (size_t)n * n_part * 4, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaFreeAsync(d_ids, st));
  MSSPE_CUDA_TRY(c, cudaFreeAsync(d_fp, st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  return MSSPE_OK;
}

extern "C" int msspe_shard_apply(msspe_ctx* c, uint8_t dir, uint32_t local_id, uint32_t n_part, uint8_t* part_flags) {
  if (!c || dir > 1 || !part_flags) return MSSPE_ERR_INVALID;
  if (!c->built) { c->set_error("msspe_shard_apply: index not built"); return MSSPE_ERR_STATE; }
  memset(part_flags, 0, n_part);
  if (local_id == MSSPE_NO_LOCAL_ID) return MSSPE_OK;
  DirIndex& D = c->dir[dir];
  if (local_id >= D.n_codes) { c->set_error("msspe_shard_apply: id %u out of range", local_id); return MSSPE_ERR_INVALID; }
  if (n_part <= c->max_partition) { c->set_error("msspe_shard_apply: n_part %u <= max partition %u", n_part, c->max_partition); return MSSPE_ERR_INVALID; }
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  uint8_t* d_flags = nullptr;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_flags, n_part, st));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(d_flags, 0, n_part, st));
  shard_apply_kernel<<<1, 1024, 0, st>>>(local_id, D.post_off, D.postings, D.ignored, c->d_seg_part, d_flags);
  c->timing.kernel_launches++;
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(part_flags, d_flags, n_part, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaFreeAsync(d_flags, st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  return MSSPE_OK;
}
