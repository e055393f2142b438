// select_device.cuh -- device building blocks of K3 shared by the launch-per-phase path (select.cu) and the
// persistent cooperative kernel (select_persistent.cu).
#pragma once
#include "engine.cuh"

namespace msspe_sel {

constexpr int CNT_THREADS = MSSPE_CNT_THREADS, CNT_ITEMS = MSSPE_CNT_ITEMS, CNT_TILE = MSSPE_CNT_TILE;  // 8192 postings = 32 KB

__device__ __forceinline__ unsigned int ld_volatile(const unsigned int* p) { return *reinterpret_cast<const volatile unsigned int*>(p); }

struct CountScratch {
  __align__(16) uint8_t nib[CNT_TILE / 4];
  uint16_t bits[CNT_THREADS];
  uint32_t tbase[CNT_THREADS + 1];
  uint32_t wsum[CNT_THREADS / 32];
};

// mask word load: shared-memory copy (plain) or the global bitmask (L2-coherent load, other blocks write it)
template <bool SMEM_MASK>
__device__ __forceinline__ uint32_t mask_word(const uint32_t* mask, uint32_t w) {
  return SMEM_MASK ? mask[w] : __ldcg(mask + w);
}

// One tile of the coverage scoring (main.rs:292-309): live flag per posting, per-k-mer sums through a tile
// prefix, freq[] writes, running maximum.  Returns the number of live postings of the tile (all threads).
// Ends with a __syncthreads() so the scratch can be reused immediately.
template <bool SMEM_MASK>
__device__ __forceinline__ uint32_t count_tile(uint32_t tile, const uint32_t* __restrict__ postings,
                                               const uint32_t* __restrict__ post_off, const uint32_t* __restrict__ tile_first,
                                               uint32_t n_codes, uint32_t n_post, const uint32_t* mask, uint32_t* freq,
                                               unsigned long long* acc, CountScratch& s, uint32_t& mymax) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tile_start = tile * (uint32_t)CNT_TILE;
  const uint32_t tile_end = min(n_post, tile_start + (uint32_t)CNT_TILE);
  const uint32_t first = __ldg(tile_first + tile);  // issued first; consumed after the posting loads are in flight
  // phase 1: coalesced 128-bit loads; one live-bit nibble per uint4
#pragma unroll
  for (int j = 0; j < CNT_ITEMS / 4; j++) {
    const uint32_t n = j * CNT_THREADS + tid;
    const uint32_t pos = tile_start + 4u * n;
    uint32_t nibble = 0;
    if (pos + 3u < tile_end) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(postings + pos));
      nibble = ((~mask_word<SMEM_MASK>(mask, v.x >> 5) >> (v.x & 31u)) & 1u) |
               (((~mask_word<SMEM_MASK>(mask, v.y >> 5) >> (v.y & 31u)) & 1u) << 1) |
               (((~mask_word<SMEM_MASK>(mask, v.z >> 5) >> (v.z & 31u)) & 1u) << 2) |
               (((~mask_word<SMEM_MASK>(mask, v.w >> 5) >> (v.w & 31u)) & 1u) << 3);
    } else {
      for (uint32_t e = 0; e < 4u; e++)
        if (pos + e < tile_end) { const uint32_t sg = __ldg(postings + pos + e); nibble |= ((~mask_word<SMEM_MASK>(mask, sg >> 5) >> (sg & 31u)) & 1u) << e; }
    }
    s.nib[n] = (uint8_t)nibble;
  }
  // first round of list bounds: in flight while the tile prefix is built
  uint32_t c = first + tid;
  uint32_t pa = c < n_codes ? __ldg(post_off + c) : 0xFFFFFFFFu;
  uint32_t pb = c < n_codes ? __ldg(post_off + c + 1) : 0xFFFFFFFFu;
  __syncthreads();
  // phase 2: per-thread 16-posting bit groups and their exclusive prefix over the tile
  const uint32_t wv = *reinterpret_cast<const uint32_t*>(&s.nib[4 * tid]);
  const uint32_t b16 = (wv & 0xFu) | (((wv >> 8) & 0xFu) << 4) | (((wv >> 16) & 0xFu) << 8) | (((wv >> 24) & 0xFu) << 12);
  const uint32_t cnt = __popc(b16);
  uint32_t inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) s.wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t ws = lane < CNT_THREADS / 32 ? s.wsum[lane] : 0u, wi = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
    if (lane < CNT_THREADS / 32) s.wsum[lane] = wi - ws;
    if (lane == CNT_THREADS / 32 - 1) s.tbase[CNT_THREADS] = wi;
  }
  __syncthreads();
  s.bits[tid] = (uint16_t)b16;
  s.tbase[tid] = s.wsum[warp] + inc - cnt;
  __syncthreads();
  const uint32_t live = s.tbase[CNT_THREADS];
  // phase 3: per-k-mer sums from prefix differences
  for (; c < n_codes; c += CNT_THREADS, pa = c < n_codes ? __ldg(post_off + c) : 0xFFFFFFFFu, pb = c < n_codes ? __ldg(post_off + c + 1) : 0xFFFFFFFFu) {
    const uint32_t a = pa;
    if (a >= tile_end) break;
    const uint32_t b = pb;
    const uint32_t lo = max(a, tile_start) - tile_start, hi = min(b, tile_end) - tile_start;
    const uint32_t plo = s.tbase[lo >> 4] + __popc((uint32_t)s.bits[lo >> 4] & ((1u << (lo & 15u)) - 1u));
    const uint32_t phi = hi == (uint32_t)CNT_TILE ? live : s.tbase[hi >> 4] + __popc((uint32_t)s.bits[hi >> 4] & ((1u << (hi & 15u)) - 1u));
    const uint32_t sum = phi - plo;
    if (a >= tile_start && b <= tile_end) {
      freq[c] = sum;
      mymax = max(mymax, sum);
    } else {  // list spans tiles: the last arriving tile owns the total
      const uint32_t first_tile = a / (uint32_t)CNT_TILE;
      const uint32_t parts = (b - 1u) / (uint32_t)CNT_TILE - first_tile + 1u;
      const unsigned long long old = atomicAdd(&acc[first_tile], (1ull << 32) | (unsigned long long)sum);
      if ((uint32_t)(old >> 32) + 1u == parts) {
        const uint32_t total = (uint32_t)old + sum;
        freq[c] = total;
        acc[first_tile] = 0ull;
        mymax = max(mymax, total);
      }
    }
  }
  __syncthreads();
  return live;
}

// Warp-cooperative partition_tie_score (main.rs:261-283) of code c.  `seen` = this warp's partition bitmap.
// COHERENT: ignored/cov are being written by other blocks of the same launch -> L2-coherent loads.
template <bool SMEM_MASK, bool COHERENT>
__device__ __forceinline__ float warp_tie_score(uint32_t c, const uint32_t* __restrict__ post_off, const uint32_t* __restrict__ postings,
                                                const uint32_t* mask, const uint16_t* __restrict__ seg_part, const uint32_t* cov,
                                                uint32_t* seen, uint32_t p_words, int lane) {
  for (uint32_t w = lane; w < p_words; w += 32) seen[w] = 0u;
  __syncwarp();
  const uint32_t a = post_off[c], b = post_off[c + 1];
  float score = 0.0f;
  for (uint32_t base = a; base < b; base += 32) {
    const uint32_t i = base + lane;
    const bool valid = i < b;
    const uint32_t seg = valid ? __ldg(postings + i) : 0u;
    const uint32_t mw = SMEM_MASK ? mask[seg >> 5] : (COHERENT ? __ldcg(mask + (seg >> 5)) : mask[seg >> 5]);
    const bool live = valid && !((mw >> (seg & 31u)) & 1u);
    const uint32_t p = live ? (uint32_t)seg_part[seg] : 0xFFFF0000u + (uint32_t)lane;
    const unsigned peers = __match_any_sync(0xffffffffu, p);
    const bool first = live && (lane == __ffs(peers) - 1);
    const bool isnew = first && !((seen[p >> 5] >> (p & 31u)) & 1u);
    const unsigned newmask = __ballot_sync(0xffffffffu, isnew);
    float term = 0.0f;
    if (isnew) {
      atomicOr(&seen[p >> 5], 1u << (p & 31u));
      const uint32_t cv = COHERENT ? __ldcg(cov + p) : cov[p];
      term = __fdiv_rn(1.0f, __fadd_rn(__uint2float_rn(cv), 1.0f));  // 1.0 / (already_covered as f32 + 1.0)
    }
    __syncwarp();
    unsigned mm = newmask;
    while (mm) {  // score += term, strictly in postings order
      const int l = __ffs(mm) - 1;
      mm &= mm - 1;
      score = __fadd_rn(score, __shfl_sync(0xffffffffu, term, l));
    }
  }
  return score;
}


// Block-cooperative partition_tie_score for long posting lists (throughput- instead of latency-bound): every
// thread looks at its share of the postings at once and records, per partition, the position of the first live
// posting (atomicMin); the distinct partitions are then ordered by that position and the f32 terms are added
// sequentially in that order -- the same order, hence the same rounding, as the scan of main.rs:268-281.
// fp[n_part] u32 and lst[n_part] u64 are shared-memory scratch; sc = {count, score bits}.
template <bool SMEM_MASK, bool COHERENT, int THREADS>
__device__ __forceinline__ float block_tie_score(uint32_t c, const uint32_t* __restrict__ post_off, const uint32_t* __restrict__ postings,
                                                 const uint32_t* mask, const uint16_t* __restrict__ seg_part, const uint32_t* cov,
                                                 uint32_t n_part, uint32_t* fp, unsigned long long* lst, uint32_t* sc) {
  const int tid = threadIdx.x;
  for (uint32_t p = tid; p < n_part; p += THREADS) fp[p] = 0xFFFFFFFFu;
  if (tid == 0) sc[0] = 0u;
  __syncthreads();
  const uint32_t a = post_off[c], b = post_off[c + 1];
  for (uint32_t i = a + tid; i < b; i += THREADS) {
    const uint32_t seg = __ldg(postings + i);
    const uint32_t mw = SMEM_MASK ? mask[seg >> 5] : (COHERENT ? __ldcg(mask + (seg >> 5)) : mask[seg >> 5]);
    if (!((mw >> (seg & 31u)) & 1u)) atomicMin(&fp[seg_part[seg]], i);
  }
  __syncthreads();
  for (uint32_t p = tid; p < n_part; p += THREADS) {
    const uint32_t f = fp[p];
    if (f != 0xFFFFFFFFu) lst[atomicAdd(&sc[0], 1u)] = ((unsigned long long)f << 32) | p;
  }
  __syncthreads();
  if (tid < 32) {
    const uint32_t n = sc[0];
    float score = 0.0f;
    if (n <= 32u) {
      const unsigned long long mine = (uint32_t)tid < n ? lst[tid] : ~0ull;
      float term = 0.0f;
      uint32_t rank = 0;
      if ((uint32_t)tid < n) {
        const uint32_t p = (uint32_t)mine;
        const uint32_t cv = COHERENT ? __ldcg(cov + p) : cov[p];
        term = __fdiv_rn(1.0f, __fadd_rn(__uint2float_rn(cv), 1.0f));
      }
      for (uint32_t j = 0; j < n; j++) { const unsigned long long o = __shfl_sync(0xffffffffu, mine, j); rank += (o < mine) ? 1u : 0u; }
      for (uint32_t r = 0; r < n; r++) {  // rank r adds next (positions are distinct, so ranks are a permutation)
        const unsigned src = __ballot_sync(0xffffffffu, (uint32_t)tid < n && rank == r);
        score = __fadd_rn(score, __shfl_sync(0xffffffffu, term, __ffs(src) - 1));
      }
    } else if (tid == 0) {  // many partitions (low-complexity repeats): plain insertion sort, then the sequential sum
      for (uint32_t i = 1; i < n; i++) {
        const unsigned long long v = lst[i];
        uint32_t j = i;
        while (j > 0 && lst[j - 1] > v) { lst[j] = lst[j - 1]; j--; }
        lst[j] = v;
      }
      for (uint32_t i = 0; i < n; i++) {
        const uint32_t p = (uint32_t)lst[i];
        const uint32_t cv = COHERENT ? __ldcg(cov + p) : cov[p];
        score = __fadd_rn(score, __fdiv_rn(1.0f, __fadd_rn(__uint2float_rn(cv), 1.0f)));
      }
    }
    if (tid == 0) sc[1] = __float_as_uint(score);
  }
  __syncthreads();
  return __uint_as_float(sc[1]);
}

}  // namespace msspe_sel
