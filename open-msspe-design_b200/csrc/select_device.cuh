// select_device.cuh -- device building blocks of K3 shared by the kernels of select.cu: the warp-tile recount
// (gather, packed scan, list windows, range carry) and the partition_tie_score variants.
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

// The global loads of one warp tile (512 postings = 2 KB), issued one tile ahead of their use.
struct TileLoad {
  uint4 v[4];
};
__device__ __forceinline__ void tile_issue(TileLoad& L, uint32_t wt, const uint32_t* __restrict__ postings, uint32_t n_post, int lane) {
  const uint32_t tile_start = wt * (uint32_t)CNT_TILE;
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

// covered bit of segment `seg` pushed into the top of the history word h (h >> 1 | bit << 31)
template <bool SMEM_MASK>
__device__ __forceinline__ uint32_t push_covered(uint32_t h, const uint32_t* mask, uint32_t seg) {
  return __funnelshift_r(h, mask_word<SMEM_MASK>(mask, seg >> 5) >> (seg & 31u), 1);
}

// Everything one warp needs to know about its share of a recount (one direction).
struct CountJob {
  const uint32_t* postings; const uint32_t* post_off;
  uint32_t n_codes, n_post;
  uint32_t t_begin, t_end;   // this warp's contiguous tile range
  uint32_t list_tiles;       // tiles [0, list_tiles) hold lists; tiles from list_tiles on hold the counted-only tail
  uint32_t tail_end;         // end of the tail (postings); = n_post when there is no tail
  uint32_t c_first;          // k-mer whose list contains the first posting of the range (tile_first[t_begin])
  const uint32_t* mask; uint32_t* freq; unsigned long long* acc;
};

// The sliding window of list starts: lane l holds post_off[c0 + l] (and the next two windows, already in flight).
struct ListWindow {
  uint32_t c0, pa, pa1, pa2;
};
__device__ __forceinline__ uint32_t list_start(const CountJob& J, uint32_t c) { return c <= J.n_codes ? __ldg(J.post_off + c) : 0xFFFFFFFFu; }

// One WARP tile of the coverage scoring (main.rs:292-309), no block-level barrier, uniform control flow.
//   Tile position q = 128 j + 4 l + e is held by lane l (coalesced uint4 number j, element e).  Every lane has gathered
//   the covered bit of its 16 postings from the bitmask (tile_gather: 16 bits, bit 4j+e) and one packed shuffle scan gives
//   the exclusive live count of every (j, l).  Lane l < 31 of the window owns k-mer c0 + l: it evaluates the live
//   prefix at the (tile-clamped) start of its list with two shuffles and takes the prefix at the end of the list
//   from lane l+1 (lane 31 only supplies that bound, which is why the window advances by 31).  A warp owns a
//   contiguous range of tiles, so a list cut by a tile boundary inside the range is carried in a register
//   (`carry` = its live postings so far); only lists that leave the range go through a tiles-covered|partial-sum word
//   (acc[first tile of the list]) and are completed by the range whose arrival covers the last missing tile.
// Returns the live postings of the tile (uniform over the warp); mymax is per lane.
// covered bits of the 16 postings of this lane: bit 16 + 4 j + e  <->  tile position 128 j + 4 lane + e
template <bool SMEM_MASK>
__device__ __forceinline__ uint32_t tile_gather(const TileLoad& L, const uint32_t* mask) {
  uint32_t h = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    h = push_covered<SMEM_MASK>(h, mask, L.v[j].x);
    h = push_covered<SMEM_MASK>(h, mask, L.v[j].y);
    h = push_covered<SMEM_MASK>(h, mask, L.v[j].z);
    h = push_covered<SMEM_MASK>(h, mask, L.v[j].w);
  }
  return h;
}

// bits of this lane's 16 tile positions (bit 4 j + e <-> position 128 j + 4 lane + e) that lie before tile_len
__device__ __forceinline__ uint32_t tile_valid_bits(uint32_t tile_len, int lane) {
  uint32_t valid = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint32_t q = (uint32_t)(j * 32 + lane) * 4u;
    const uint32_t v = q >= tile_len ? 0u : (tile_len - q >= 4u ? 0xFu : (1u << (tile_len - q)) - 1u);
    valid |= v << (4 * j);
  }
  return valid;
}

__device__ __forceinline__ uint32_t warp_count_tile(uint32_t h, uint32_t wt, const CountJob& J, ListWindow& W,
                                                    uint32_t& carry, uint32_t& mymax, int lane) {
  const uint32_t tile_start = wt * (uint32_t)CNT_TILE;
  const uint32_t tile_len = min(J.n_post - tile_start, (uint32_t)CNT_TILE);
  const uint32_t tile_end = tile_start + tile_len;
  uint32_t nibs = ~h >> 16;  // live bit of tile position 128 j + 4 lane + e at bit 4 j + e
  if (tile_len < (uint32_t)CNT_TILE) nibs &= tile_valid_bits(tile_len, lane);  // the last tile: drop the slots past the end
  const uint32_t cnts = (uint32_t)__popc(nibs & 0xFu) | ((uint32_t)__popc(nibs & 0xF0u) << 8) |
                        ((uint32_t)__popc(nibs & 0xF00u) << 16) | ((uint32_t)__popc(nibs & 0xF000u) << 24);
  // packed inclusive scan over lanes: field j (8 bits) = live count of block j up to this lane (<= 128)
  uint32_t inc = cnts;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
  const uint32_t excl = inc - cnts;
  const uint32_t b1 = tot & 0xFFu, b2 = b1 + ((tot >> 8) & 0xFFu), b3 = b2 + ((tot >> 16) & 0xFFu), live = b3 + (tot >> 24);
  const uint32_t b01 = b1 << 16, b23 = b2 | (b3 << 16);  // live postings before block j as 16-bit fields (block 0: none)
  const uint32_t range_lo = J.t_begin * (uint32_t)CNT_TILE;
  const bool last_of_range = wt + 1u == min(J.t_end, J.list_tiles);  // the last LIST tile of the range
  uint32_t outsum = 0;  // live postings so far of the list that continues into the next tile of the range
  for (;;) {
    // live postings of the tile before the start of this lane's list (list starts clamped into the tile), branch-free:
    // position 512 (= end of a full tile) is looked up as element 4 of the last nibble of lane 31
    const uint32_t lo = min(max(W.pa, tile_start), tile_end) - tile_start;  // 0 .. 512
    const uint32_t lq = min(lo, (uint32_t)CNT_TILE - 1u);
    const uint32_t j = lq >> 7, e = lo - (lq & ~3u);
    const uint32_t ex = __shfl_sync(0xffffffffu, excl, lq >> 2), nb = __shfl_sync(0xffffffffu, nibs, lq >> 2);
    const uint32_t base = __byte_perm(b01, b23, 0x10u + j * 0x22u);     // 16-bit field j, zero-extended (bytes 0,1 of b01 are 0)
    const uint32_t exj = __byte_perm(ex, 0u, 0x4440u + j);               // byte j of the packed exclusive counts
    const uint32_t plo = base + exj + (uint32_t)__popc((nb >> (4u * j)) & ((1u << e) - 1u));
    const uint32_t pb = __shfl_down_sync(0xffffffffu, W.pa, 1), phi = __shfl_down_sync(0xffffffffu, plo, 1);
    if (lane < 31 && W.pa < tile_end && pb > tile_start) {  // the list overlaps this tile
      const uint32_t total = phi - plo + (W.pa < tile_start ? carry : 0u);
      const bool ends_here = pb <= tile_end;
      if (ends_here && W.pa >= range_lo) {  // the whole list lies inside this warp's range
        J.freq[W.c0 + lane] = total;
        mymax = max(mymax, total);
      } else if (!ends_here && !last_of_range) {
        outsum = total;
      } else {  // the list leaves the range: the last arriving range owns the total
        // arrivals are counted in TILES of the list covered by the arriving range, so ranges of any shape work
        const uint32_t t_a = W.pa / (uint32_t)CNT_TILE, t_b = (pb - 1u) / (uint32_t)CNT_TILE;
        const uint32_t covered = min(t_b, J.t_end - 1u) - max(t_a, J.t_begin) + 1u;
        const unsigned long long old = atomicAdd(&J.acc[t_a], ((unsigned long long)covered << 32) | (unsigned long long)total);
        if ((uint32_t)(old >> 32) + covered == t_b - t_a + 1u) {
          const uint32_t sum = (uint32_t)old + total;
          J.freq[W.c0 + lane] = sum;
          J.acc[t_a] = 0ull;
          mymax = max(mymax, sum);
        }
      }
    }
    if (__shfl_sync(0xffffffffu, W.pa, 31) >= tile_end) break;  // uniform: no further list starts inside this tile
    W.c0 += 31u;
    W.pa = W.pa1; W.pa1 = W.pa2;
    W.pa2 = list_start(J, W.c0 + 62u + (uint32_t)lane);
  }
  carry = __reduce_add_sync(0xffffffffu, outsum);
  return live;
}

// This warp's share of one recount: its contiguous tile range.  One register buffer: as soon as the covered bits of
// a tile have been gathered its registers take the loads of the next tile, which are in flight while the tile is
// scored.  (An additional prefetch.global.L2 of the tile after that was measured: 4.37 TB/s with it at distance 2,
// 4.22 at distance 4, 4.66 without -- 32 warps x 2 KB in flight per SM already cover the latency.)  Tiles past the lists hold the
// counted-only tail (single-posting lists of the scoring stream): gather and popcount, no per-list reduction.
struct RangeState {
  TileLoad A;
  ListWindow W;
};
__device__ __forceinline__ uint32_t tile_bound(const CountJob& J, uint32_t wt) { return wt < J.list_tiles ? J.n_post : J.tail_end; }
// the first loads of the range (postings of the first tile, three windows of list starts): nothing here depends on
// the bitmask, so the caller can issue them before the mask update of the iteration
__device__ __forceinline__ void warp_count_begin(const CountJob& J, RangeState& S, int lane) {
  if (J.t_begin >= J.t_end) return;
  tile_issue(S.A, J.t_begin, J.postings, tile_bound(J, J.t_begin), lane);
  if (J.t_begin < J.list_tiles) {
    S.W.c0 = J.c_first;
    S.W.pa = list_start(J, S.W.c0 + (uint32_t)lane);
    S.W.pa1 = list_start(J, S.W.c0 + 31u + (uint32_t)lane);
    S.W.pa2 = list_start(J, S.W.c0 + 62u + (uint32_t)lane);
  }
}

template <bool SMEM_MASK>
__device__ __forceinline__ unsigned long long warp_count_run(const CountJob& J, RangeState& S, uint32_t& mymax, int lane) {
  unsigned long long live = 0;
  uint32_t carry = 0, tail_live = 0;
  for (uint32_t wt = J.t_begin; wt < J.t_end; wt++) {
    const uint32_t h = tile_gather<SMEM_MASK>(S.A, J.mask);
    if (wt + 1u < J.t_end) tile_issue(S.A, wt + 1u, J.postings, tile_bound(J, wt + 1u), lane);
    if (wt < J.list_tiles) {
      live += warp_count_tile(h, wt, J, S.W, carry, mymax, lane);
    } else {
      uint32_t nibs = ~h >> 16;
      const uint32_t tile_len = min(J.tail_end - wt * (uint32_t)CNT_TILE, (uint32_t)CNT_TILE);
      if (tile_len < (uint32_t)CNT_TILE) nibs &= tile_valid_bits(tile_len, lane);
      tail_live += (uint32_t)__popc(nibs);
    }
  }
  if (J.t_end > J.list_tiles) live += __reduce_add_sync(0xffffffffu, tail_live);
  return live;
}

template <bool SMEM_MASK>
__device__ __forceinline__ unsigned long long warp_count_range(const CountJob& J, uint32_t& mymax, int lane) {
  RangeState S;
  warp_count_begin(J, S, lane);
  return warp_count_run<SMEM_MASK>(J, S, mymax, lane);
}

// Tile range of warp `gw` out of `n_warps`, equal tile counts: T = ceil(n_tiles / n_warps) consecutive tiles per warp
// (stand-alone count kernel; the persistent kernel balances its ranges by cost, select.cu).
__device__ __forceinline__ void count_job_range(CountJob& J, uint32_t n_tiles, uint32_t gw, uint32_t n_warps) {
  const uint32_t T = max((n_tiles + n_warps - 1u) / n_warps, 1u);
  J.t_begin = min(n_tiles, gw * T);
  J.t_end = min(n_tiles, J.t_begin + T);
  J.list_tiles = n_tiles;
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
// fp[n_part] u32 and lst[n_part] u64 are shared-memory scratch; sc = {count, score bits}.  COHERENT: cov is the global
// table (written by block 0 inside this launch) rather than the block's shared-memory copy of this iteration.
// Segment.partition_no (main.rs:227) of segment `seg`: its index inside its record when all records have equally
// many partitions (pre-aligned, equal-length genomes: no table lookup), else from the table.
__device__ __forceinline__ uint32_t partition_of(const uint16_t* __restrict__ seg_part, uint32_t uniform_parts, uint32_t seg) {
  return uniform_parts ? seg % uniform_parts : (uint32_t)seg_part[seg];
}

template <bool SMEM_MASK, bool COHERENT, int THREADS>
__device__ __forceinline__ float block_tie_score(uint32_t c, const uint32_t* __restrict__ post_off, const uint32_t* __restrict__ postings,
                                                 const uint32_t* mask, const uint16_t* __restrict__ seg_part, uint32_t uniform_parts, const uint32_t* cov,
                                                 uint32_t n_part, uint32_t* fp, unsigned long long* lst, uint32_t* sc) {
  const int tid = threadIdx.x;
  for (uint32_t p = tid; p < n_part; p += THREADS) fp[p] = 0xFFFFFFFFu;
  if (tid == 0) sc[0] = 0u;
  __syncthreads();
  const uint32_t a = post_off[c], b = post_off[c + 1];
  for (uint32_t i = a + tid; i < b; i += THREADS) {
    const uint32_t seg = __ldg(postings + i);
    const uint32_t mw = SMEM_MASK ? mask[seg >> 5] : __ldcg(mask + (seg >> 5));  // the global bitmask is written inside this launch
    if (!((mw >> (seg & 31u)) & 1u)) atomicMin(&fp[partition_of(seg_part, uniform_parts, seg)], i);
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
