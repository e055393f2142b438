// select_device.cuh -- device building blocks of K3 shared by the launch-per-phase path (select.cu) and the
// persistent cooperative kernel (select_persistent.cu).
#pragma once
#include "engine.cuh"

namespace msspe_sel {

constexpr int CNT_THREADS = MSSPE_CNT_THREADS, CNT_TILE = MSSPE_CNT_TILE;  // block size; postings per WARP tile

__device__ __forceinline__ unsigned int ld_volatile(const unsigned int* p) { return *reinterpret_cast<const volatile unsigned int*>(p); }

// mask word load: shared-memory copy (plain) or the global bitmask (L2-coherent load, other blocks write it)
template <bool SMEM_MASK>
__device__ __forceinline__ uint32_t mask_word(const uint32_t* mask, uint32_t w) {
  return SMEM_MASK ? mask[w] : __ldcg(mask + w);
}
template <bool SMEM_MASK>
__device__ __forceinline__ uint32_t live_bit(const uint32_t* mask, uint32_t seg) {
  return (~mask_word<SMEM_MASK>(mask, seg >> 5) >> (seg & 31u)) & 1u;
}

// Live postings before position pos (0..511) of the current warp tile, from the packed per-lane scan state.
__device__ __forceinline__ uint32_t tile_prefix(uint32_t pos, uint32_t excl, uint32_t nibs, uint32_t b1, uint32_t b2, uint32_t b3) {
  const uint32_t j = pos >> 7, l = (pos >> 2) & 31u, e = pos & 3u;
  const uint32_t ex = __shfl_sync(0xffffffffu, excl, l), nb = __shfl_sync(0xffffffffu, nibs, l);
  const uint32_t base = j == 0 ? 0u : (j == 1 ? b1 : (j == 2 ? b2 : b3));
  return base + ((ex >> (8u * j)) & 0xFFu) + __popc((nb >> (4u * j)) & ((1u << e) - 1u));
}

// One WARP tile (512 postings = 2 KB) of the coverage scoring (main.rs:292-309), no block-level barrier:
//   lane l loads four coalesced uint4 (block j = postings [128j, 128j+128) of the tile, lane l owns 4 of them),
//   gathers the covered-segment bit of each posting, keeps 4 live nibbles; one packed shuffle scan gives the
//   exclusive live-count of every (block, lane).  A k-mer's live count is prefix(next list start) - prefix(own
//   list start): every lane evaluates the prefix at its own list start (two shuffles) and takes the upper bound
//   from its neighbour lane.  Lists that cross a tile boundary are assembled by the last arriving tile through an
//   arrival-counter|partial-sum word (acc[first tile of the list]).
// Returns the live postings of the tile (uniform over the warp); mymax is per lane.
// The global loads of one warp tile, issued one tile ahead of their use (register double buffering).
struct TileLoad {
  uint4 v[4];
  uint32_t first;
};
__device__ __forceinline__ void tile_issue(TileLoad& L, uint32_t wt, const uint32_t* __restrict__ postings,
                                           const uint32_t* __restrict__ tile_first, uint32_t n_post, int lane) {
  const uint32_t tile_start = wt * (uint32_t)CNT_TILE;
  L.first = __ldg(tile_first + wt);
  if (tile_start + (uint32_t)CNT_TILE <= n_post) {
#pragma unroll
    for (int j = 0; j < 4; j++) L.v[j] = __ldg(reinterpret_cast<const uint4*>(postings + tile_start + (uint32_t)(j * 32 + lane) * 4u));
  } else {  // the last, partial tile: out-of-range slots read posting 0 and are masked out later
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t pos = tile_start + (uint32_t)(j * 32 + lane) * 4u;
      L.v[j].x = __ldg(postings + (pos + 0u < n_post ? pos + 0u : 0u));
      L.v[j].y = __ldg(postings + (pos + 1u < n_post ? pos + 1u : 0u));
      L.v[j].z = __ldg(postings + (pos + 2u < n_post ? pos + 2u : 0u));
      L.v[j].w = __ldg(postings + (pos + 3u < n_post ? pos + 3u : 0u));
    }
  }
}

template <bool SMEM_MASK>
__device__ __forceinline__ uint32_t warp_count_tile(const TileLoad& L, uint32_t wt,
                                                    const uint32_t* __restrict__ post_off,
                                                    uint32_t n_codes, uint32_t n_post, const uint32_t* mask, uint32_t* freq,
                                                    unsigned long long* acc, uint32_t& mymax, int lane) {
  const uint32_t tile_start = wt * (uint32_t)CNT_TILE;
  const uint32_t tile_end = min(n_post, tile_start + (uint32_t)CNT_TILE);
  const uint32_t first = L.first;
  const bool partial = tile_start + (uint32_t)CNT_TILE > n_post;
  uint32_t nibs = 0, cnts = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    uint32_t nib = live_bit<SMEM_MASK>(mask, L.v[j].x) | (live_bit<SMEM_MASK>(mask, L.v[j].y) << 1) |
                   (live_bit<SMEM_MASK>(mask, L.v[j].z) << 2) | (live_bit<SMEM_MASK>(mask, L.v[j].w) << 3);
    if (partial) {
      const uint32_t pos = tile_start + (uint32_t)(j * 32 + lane) * 4u;
      const uint32_t valid = pos >= tile_end ? 0u : (tile_end - pos >= 4u ? 0xFu : (1u << (tile_end - pos)) - 1u);
      nib &= valid;
    }
    nibs |= nib << (4 * j);
    cnts |= (uint32_t)__popc(nib) << (8 * j);
  }
  // list starts of the first 32*CB k-mers of the tile (CB per lane) + the one after them; in flight during the scan
  constexpr int CB = 4;
  uint32_t c0 = first + lane;
  uint32_t pa[CB];
#pragma unroll
  for (int r = 0; r < CB; r++) { const uint32_t c = c0 + 32u * r; pa[r] = c <= n_codes ? __ldg(post_off + c) : 0xFFFFFFFFu; }
  uint32_t pnext = first + 32u * CB <= n_codes ? __ldg(post_off + first + 32u * CB) : 0xFFFFFFFFu;
  // packed inclusive scan over lanes: field j (8 bits) = live count of block j up to this lane (<= 128)
  uint32_t inc = cnts;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
  const uint32_t excl = inc - cnts;
  const uint32_t b1 = tot & 0xFFu, b2 = b1 + ((tot >> 8) & 0xFFu), b3 = b2 + ((tot >> 16) & 0xFFu), live = b3 + (tot >> 24);
  for (;;) {
    // prefix at every active list start of this batch
    uint32_t plo[CB];
    int nr = 0;
#pragma unroll
    for (int r = 0; r < CB; r++) {
      if (r == nr && __any_sync(0xffffffffu, pa[r] < tile_end)) {
        const uint32_t lo = pa[r] < tile_end ? max(pa[r], tile_start) - tile_start : 0u;
        const uint32_t v = tile_prefix(lo, excl, nibs, b1, b2, b3);
        plo[r] = pa[r] < tile_end ? v : live;
        nr = r + 1;
      } else {
        plo[r] = live;
      }
    }
    const bool more = nr == CB && pnext < tile_end;  // uniform: another batch of lists starts inside this tile
    uint32_t pnext_pre = live;
    if (more) pnext_pre = tile_prefix(pnext - tile_start, excl, nibs, b1, b2, b3);
#pragma unroll
    for (int r = 0; r < CB; r++) {
      if (r >= nr) break;
      // the next list's start and its prefix: neighbour lane, or lane 0 of the next round, or the batch successor
      const uint32_t nxt_r = r + 1 < CB ? pa[r + 1 < CB ? r + 1 : r] : pnext;
      const uint32_t nxt_p = r + 1 < CB ? plo[r + 1 < CB ? r + 1 : r] : pnext_pre;
      uint32_t pb = __shfl_down_sync(0xffffffffu, pa[r], 1), phi = __shfl_down_sync(0xffffffffu, plo[r], 1);
      const uint32_t pb31 = __shfl_sync(0xffffffffu, nxt_r, 0), phi31 = __shfl_sync(0xffffffffu, nxt_p, 0);
      if (lane == 31) { pb = pb31; phi = phi31; }
      if (pa[r] < tile_end) {
        const uint32_t c = c0 + 32u * r;
        const uint32_t sum = (pb < tile_end ? phi : live) - plo[r];
        if (pa[r] >= tile_start && pb <= tile_end) {
          freq[c] = sum;
          mymax = max(mymax, sum);
        } else {  // list spans tiles: the last arriving tile owns the total
          const uint32_t first_tile = pa[r] / (uint32_t)CNT_TILE;
          const uint32_t parts = (pb - 1u) / (uint32_t)CNT_TILE - first_tile + 1u;
          const unsigned long long old = atomicAdd(&acc[first_tile], (1ull << 32) | (unsigned long long)sum);
          if ((uint32_t)(old >> 32) + 1u == parts) {
            const uint32_t total = (uint32_t)old + sum;
            freq[c] = total;
            acc[first_tile] = 0ull;
            mymax = max(mymax, total);
          }
        }
      }
    }
    if (!more) break;
    c0 += 32u * CB;
#pragma unroll
    for (int r = 0; r < CB; r++) { const uint32_t c = c0 + 32u * r; pa[r] = c <= n_codes ? __ldg(post_off + c) : 0xFFFFFFFFu; }
    pnext = c0 - lane + 32u * CB <= n_codes ? __ldg(post_off + (c0 - lane) + 32u * CB) : 0xFFFFFFFFu;
  }
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
