// kmer_build_fast.cu -- K1 + K2 for words of up to 16 bases (2-bit code and record index in ONE u64): the index build
// of get_segment_manager + make_kmer_segments_windows_mapping (od-msspe/src/main.rs:196-255) in one pass per stage.
//
//   K1 encode_keys_packed_kernel (search windows <= 64 bases) / encode_keys_kernel (longer windows)
//                           the w bases of the head (or tail) search window -> 2-bit, one key per slot at a FIXED stride:
//                           key = (code << idx_bits) | (segment * slots + slot),  all ones for a slot that is invalid
//                           (non-ACGT base, main.rs:163-171) or a repeat inside its window (itertools unique()).  No count
//                           pass, no scan, no compaction: 99 % of the slots are valid.  The packed form works on bit planes
//                           of the window in registers and finds repeats by comparing the window with itself at every
//                           distance (tests/test_encode_model.py is its executable specification).
//   K2 radix sort           LSD over the 2k code bits only, 3 passes of <= 11 bits (k <= 16; MSSPE_SORT_BITS narrows the
//                           digits): per pass a block histogram (block-major), a tiled two-kernel scan over digits x tiles,
//                           one scatter.  Keys carry their record index in the low bits, so a stable sort leaves every
//                           posting list ascending by segment (main.rs:250) without a second key.  The scatter ranks its
//                           tile warp by warp (uniform round: one __match_all_sync; all-different round: tag and read back;
//                           mixed: __match_any_sync or ballots), stages the tile in shared memory in digit order and writes
//                           it out with consecutive threads on consecutive addresses.  Pass 0 drops the all-ones keys, so R
//                           (valid records) is only known on the device; later passes read it there -- the host
//                           synchronises ONCE per build, for both directions together.
//   CSR                     head flags -> scan -> codes / post_off / postings / fwd_ids / list_part as in kmer_build.cu.
// Algorithmic bytes: encode w B read + 8 B written per slot; histogram 8 B per key; scatter 16 B per
// key ("2 x 8 B per record per pass"); CSR 8 + 4 B read, 4 + 4 B written per record.
#include <algorithm>

#include "engine.cuh"

namespace {

constexpr int ENCF_WARPS = 8;
constexpr int SORT_T = 256, SORT_ITEMS = 16, SORT_TILE = SORT_T * SORT_ITEMS, SORT_WARPS = SORT_T / 32;
constexpr unsigned long long KEY_NONE = ~0ull;

__device__ __forceinline__ uint32_t base2f(uint8_t c) {
  switch (c) {
    case 'A': case 'a': return 0u;
    case 'C': case 'c': return 1u;
    case 'G': case 'g': return 2u;
    case 'T': case 't': case 'U': case 'u': return 3u;
    default: return 4u;
  }
}

// Packed path (search windows of up to 64 bases; the reference's default is 50): no shared memory.  A warp takes
// ENC_SEGS consecutive segments with all their loads issued first; a window is ONE packed 2-bit string in registers
// (position p at bits [126 - 2p, 127 - 2p] of ph:pl, built with warp OR-reductions) so that a slot's word is a funnel
// shift, plus three 64-bit planes over the positions (low bit, high bit, is-ACGT; ballots).  Lane l owns slots l and
// l + 32.  A slot is valid when k consecutive positions are ACGT (a run test on the third plane).  The first occurrence
// inside the window (itertools unique()) needs no comparison of words: the word at slot q + d repeats the word at slot
// q exactly when the window equals itself shifted by d over k consecutive positions, so lane l tests the distances
// l + 1 and l + 33 with shifts and ANDs on the planes and one OR-reduction merges the verdicts.  (__match_any_sync on
// the 64-bit words cost 128 SM cycles per segment here: 2 cycles per distinct value, measured, tools/microbench.)
constexpr int ENC_SEGS = 4;

__device__ __forceinline__ uint32_t base2p(uint32_t c) {   // A0 C1 G2 T3 (U = T), 4 = anything else; either case
  const uint32_t lc = c | 0x20u;
  const bool ok = lc == 'a' || lc == 'c' || lc == 'g' || lc == 't' || lc == 'u';
  const uint32_t x = (c >> 1) & 3u;
  return ok ? (x ^ (x >> 1)) : 4u;
}

template <int DIR>
__global__ void __launch_bounds__(ENCF_WARPS * 32)
encode_keys_packed_kernel(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, const uint64_t* __restrict__ seg_base,
                          uint32_t n_records, uint32_t uniform_parts, uint64_t n_segments, uint32_t W, uint32_t S, uint32_t w, uint32_t k,
                          uint32_t slots, uint32_t idx_bits, unsigned long long* __restrict__ keys, uint16_t* __restrict__ seg_part,
                          uint32_t* __restrict__ seg_rec) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t gw = ((uint64_t)blockIdx.x * ENCF_WARPS + (threadIdx.x >> 5)) * ENC_SEGS;
  if (gw >= n_segments) return;
  uint32_t v0[ENC_SEGS], v1[ENC_SEGS];
#pragma unroll
  for (int s = 0; s < ENC_SEGS; s++) {
    const uint64_t g = gw + s;
    v0[s] = v1[s] = 'N';
    if (g < n_segments) {
      uint32_t lo = 0, hi = n_records;
      uint64_t j;
      if (uniform_parts) {
        if (n_segments <= 0xFFFFFFFFull) { lo = (uint32_t)g / uniform_parts; j = (uint32_t)g - lo * uniform_parts; }   // a 64-bit division is ~80 instructions
        else { lo = (uint32_t)(g / uniform_parts); j = g - (uint64_t)lo * uniform_parts; }
      }
      else { while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (seg_base[mid] <= g) lo = mid; else hi = mid; } j = g - seg_base[lo]; }
      const uint64_t win = offsets[lo] + j * (uint64_t)S;
      const uint64_t src = DIR == 0 ? win : win + (W - w);
      if (DIR == 0 && lane == 0) { seg_part[g] = (uint16_t)j; seg_rec[g] = lo; }
      if (lane < w) v0[s] = bases[src + lane];
      if (lane + 32u < w) v1[s] = bases[src + lane + 32u];
    }
  }
  const unsigned long long kmask = (k == 32u) ? ~0ull : ((1ull << (2u * k)) - 1ull);
  // bit p of run_k(m): m has ones at p .. p + k - 1
  auto run_k = [k](unsigned long long m) {
    uint32_t len = 1;
    while (2u * len <= k) { m &= m >> len; len *= 2u; }
    if (len < k) m &= m >> (k - len);
    return m;
  };
#pragma unroll
  for (int s = 0; s < ENC_SEGS; s++) {
    const uint64_t g = gw + s;
    if (g >= n_segments) break;                                  // uniform over the warp
    const uint32_t b0 = lane < w ? base2p(v0[s]) : 4u, b1 = lane + 32u < w ? base2p(v1[s]) : 4u;
    const unsigned long long c0 = (unsigned long long)(b0 & 3u) << (62 - 2 * lane), c1 = (unsigned long long)(b1 & 3u) << (62 - 2 * lane);
    const unsigned long long ph = ((unsigned long long)__reduce_or_sync(0xffffffffu, (uint32_t)(c0 >> 32)) << 32) | __reduce_or_sync(0xffffffffu, (uint32_t)c0);
    const unsigned long long pl = ((unsigned long long)__reduce_or_sync(0xffffffffu, (uint32_t)(c1 >> 32)) << 32) | __reduce_or_sync(0xffffffffu, (uint32_t)c1);
    // planes, position p at bit p
    const unsigned long long lo = ((unsigned long long)__ballot_sync(0xffffffffu, b1 & 1u) << 32) | __ballot_sync(0xffffffffu, b0 & 1u);
    const unsigned long long hi = ((unsigned long long)__ballot_sync(0xffffffffu, b1 & 2u) << 32) | __ballot_sync(0xffffffffu, b0 & 2u);
    const unsigned long long val = ((unsigned long long)__ballot_sync(0xffffffffu, b1 < 4u) << 32) | __ballot_sync(0xffffffffu, b0 < 4u);
    const unsigned long long okm = run_k(val);                     // slots whose k positions are all ACGT (main.rs:163-171)
    unsigned long long dup = 0ull;                                 // slots that repeat an earlier slot's word
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const uint32_t dist = lane + 1u + 32u * h;
      if (dist < slots) {
        const unsigned long long eq = ~((lo ^ (lo >> dist)) | (hi ^ (hi >> dist))) & val & (val >> dist);   // position p equals position p + dist
        dup |= run_k(eq) << dist;
      }
    }
    dup = ((unsigned long long)__reduce_or_sync(0xffffffffu, (uint32_t)(dup >> 32)) << 32) | __reduce_or_sync(0xffffffffu, (uint32_t)dup);
    const unsigned long long keepm = okm & ~dup;
    const unsigned long long rec = g * slots + lane;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const uint32_t q = lane + 32u * h;
      if (q < slots) {
        const unsigned long long x = h ? (pl << (2u * lane)) : (lane == 0u ? ph : ((ph << (2u * lane)) | (pl >> (64u - 2u * lane))));
        unsigned long long cd = x >> (64u - 2u * k);               // bases q .. q + k - 1, first base in the top bits
        if (DIR == 1) {                                            // reverse complement (main.rs:148-161): complement, reverse the 2-bit groups
          const unsigned long long r = __brevll((~cd) & kmask) >> (64u - 2u * k);
          cd = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
        }
        keys[rec + 32u * h] = ((keepm >> q) & 1ull) ? ((cd << idx_bits) | (rec + 32u * h)) : KEY_NONE;
      }
    }
  }
}

// General path (search windows longer than 64 bases).  dynamic smem per warp: slots_pad u64 codes + w bytes
template <int DIR>
__global__ void __launch_bounds__(ENCF_WARPS * 32)
encode_keys_kernel(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, const uint64_t* __restrict__ seg_base,
                   uint32_t n_records, uint32_t uniform_parts, uint64_t n_segments, uint32_t W, uint32_t S, uint32_t w, uint32_t k,
                   uint32_t slots, uint32_t idx_bits, unsigned long long* __restrict__ keys, uint16_t* __restrict__ seg_part,
                   uint32_t* __restrict__ seg_rec) {
  extern __shared__ __align__(8) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t slots_pad = (slots + 1u) & ~1u;
  const size_t per_warp = (size_t)slots_pad * 8 + ((w + 7u) & ~7u);
  unsigned long long* scode = reinterpret_cast<unsigned long long*>(smem_raw + warp * per_warp);
  uint8_t* sb = reinterpret_cast<uint8_t*>(scode + slots_pad);
  const uint64_t g = (uint64_t)blockIdx.x * ENCF_WARPS + warp;
  if (g >= n_segments) return;
  uint32_t lo = 0, hi = n_records;
  if (uniform_parts) lo = (uint32_t)(g / uniform_parts);
  else while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (seg_base[mid] <= g) lo = mid; else hi = mid; }
  const uint64_t j = g - seg_base[lo];
  const uint64_t win = offsets[lo] + j * (uint64_t)S;
  const uint64_t src = DIR == 0 ? win : win + (W - w);
  if (DIR == 0 && lane == 0) { seg_part[g] = (uint16_t)j; seg_rec[g] = lo; }
  for (uint32_t t = lane; t < w; t += 32) sb[t] = (uint8_t)base2f(bases[src + t]);
  __syncwarp();
  for (uint32_t q0 = 0; q0 < slots; q0 += 32) {
    const uint32_t q = q0 + lane;
    unsigned long long code = KEY_NONE;
    if (q < slots) {
      unsigned long long cd = 0; bool ok = true;
      if (DIR == 0) { for (uint32_t t = 0; t < k; t++) { const uint32_t b = sb[q + t]; ok &= (b < 4u); cd = (cd << 2) | (b & 3u); } }
      else { for (uint32_t t = 0; t < k; t++) { const uint32_t b = sb[q + k - 1 - t]; ok &= (b < 4u); cd = (cd << 2) | ((3u - b) & 3u); } }  // reverse complement
      if (ok) code = cd;
    }
    // first occurrence inside the window: earlier lanes of this round, then the rounds before
    const unsigned peers = __match_any_sync(0xffffffffu, code);
    bool keep = code != KEY_NONE && lane == __ffs(peers) - 1;
    for (uint32_t p = 0; keep && p < q0; p++) keep = scode[p] != code;
    if (q < slots) scode[q] = code;
    __syncwarp();
    if (q < slots) {
      const unsigned long long rec = g * slots + q;
      keys[rec] = keep ? ((code << idx_bits) | rec) : KEY_NONE;
    }
  }
}

struct SortPass { int shift, bits; };

// hist[block * nbuckets + d]; pass 0 (n_dev == nullptr) counts only valid keys of the n0 slots and adds them up in *n_out
__global__ void __launch_bounds__(SORT_T)
fast_hist_kernel(const unsigned long long* __restrict__ keys, uint64_t n0, const uint32_t* __restrict__ n_dev, int shift, uint32_t nbuckets,
                 uint32_t* __restrict__ hist, uint32_t nb, uint32_t* __restrict__ n_out) {
  extern __shared__ uint32_t h[];
  for (uint32_t i = threadIdx.x; i < nbuckets; i += SORT_T) h[i] = 0u;
  __syncthreads();
  const uint64_t n = n_dev ? (uint64_t)*n_dev : n0;
  const uint64_t base = (uint64_t)blockIdx.x * SORT_TILE;
  uint32_t valid = 0;
  unsigned long long key[SORT_ITEMS];
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {       // all loads first: sixteen independent requests in flight per thread
    const uint64_t p = base + (uint64_t)i * SORT_T + threadIdx.x;
    key[i] = p < n ? keys[p] : KEY_NONE;
  }
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++)
    if (key[i] != KEY_NONE) { atomicAdd(&h[(uint32_t)(key[i] >> shift) & (nbuckets - 1u)], 1u); valid++; }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < nbuckets; i += SORT_T) hist[(uint64_t)blockIdx.x * nbuckets + i] = h[i];   // [block][digit]: coalesced
  if (n_out) {
    for (int o = 16; o > 0; o >>= 1) valid += __shfl_down_sync(0xffffffffu, valid, o);
    if ((threadIdx.x & 31) == 0 && valid) atomicAdd(n_out, valid);
  }
}

// Offsets of one pass from the block histograms hist[block][digit] (every block's row is contiguous, so the histogram
// kernel writes it and the scatter kernel reads it with consecutive threads on consecutive addresses; the digit-major layout
// of the first version cost each of them one 32-byte sector per digit and block -- as much traffic as the keys).  A block of
// this kernel owns 32 consecutive digits (lane = digit) and all tiles: its warps split the tiles, sum their share, exchange
// the sums, then rewrite their share as running (exclusive) counts; dtot[d] receives the digit's total, which
// fast_digit_base_kernel turns into the digit's base.  Two launches per pass instead of the seven of the generic scan.
constexpr int DSCAN_WARPS = 16, DSCAN_SPLIT = 8, DSCAN_CHUNKS = DSCAN_WARPS * DSCAN_SPLIT;   // the tiles are cut into 128 chunks of rows (256: 36 us against 30)
__device__ __forceinline__ void dscan_range(uint32_t nb, uint32_t chunk, uint32_t* r0, uint32_t* r1) {
  const uint32_t per = (nb + DSCAN_CHUNKS - 1) / DSCAN_CHUNKS;
  *r0 = chunk * per < nb ? chunk * per : nb;
  *r1 = *r0 + per < nb ? *r0 + per : nb;
}
// grid (nbuckets / 32, DSCAN_SPLIT), one warp per chunk, lane = digit: csum[chunk][d] = sum of the chunk's rows
__global__ void __launch_bounds__(DSCAN_WARPS * 32)
fast_digit_partial_kernel(const uint32_t* __restrict__ hist, uint32_t nb, uint32_t nbuckets, uint32_t* __restrict__ csum) {
  const uint32_t lane = threadIdx.x & 31u, chunk = blockIdx.y * DSCAN_WARPS + (threadIdx.x >> 5);
  const uint32_t d = blockIdx.x * 32u + lane;
  if (d >= nbuckets) return;
  uint32_t r0, r1;
  dscan_range(nb, chunk, &r0, &r1);
  uint32_t sum = 0;
#pragma unroll 8
  for (uint32_t r = r0; r < r1; r++) sum += hist[(uint64_t)r * nbuckets + d];
  csum[(uint64_t)chunk * nbuckets + d] = sum;
}
// same grid: the chunk's rows become running (exclusive) counts of their digit; the last chunk leaves the digit's total
__global__ void __launch_bounds__(DSCAN_WARPS * 32)
fast_digit_scan_kernel(uint32_t* __restrict__ hist, uint32_t nb, uint32_t nbuckets, const uint32_t* __restrict__ csum, uint32_t* __restrict__ dtot) {
  const uint32_t lane = threadIdx.x & 31u, chunk = blockIdx.y * DSCAN_WARPS + (threadIdx.x >> 5);
  const uint32_t d = blockIdx.x * 32u + lane;
  // sum of the chunks before this one: the block loads all chunk sums of its 32 digits once (its warps share the rows)
  __shared__ uint32_t s_c[DSCAN_CHUNKS][32];
  const uint32_t need = (blockIdx.y + 1u) * DSCAN_WARPS;       // chunks up to the end of this block's split
  for (uint32_t c2 = threadIdx.x >> 5; c2 < need; c2 += DSCAN_WARPS) s_c[c2][lane] = d < nbuckets ? csum[(uint64_t)c2 * nbuckets + d] : 0u;
  __syncthreads();
  if (d >= nbuckets) return;
  uint32_t run = 0;
  for (uint32_t c2 = 0; c2 < chunk; c2++) run += s_c[c2][lane];
  uint32_t r0, r1;
  dscan_range(nb, chunk, &r0, &r1);
  for (uint32_t r = r0; r < r1; r += 8) {       // eight rows in flight: the loads of a group are issued before its stores
    uint32_t v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = r + q < r1 ? hist[(uint64_t)(r + q) * nbuckets + d] : 0u;
#pragma unroll
    for (int q = 0; q < 8; q++) { if (r + q < r1) hist[(uint64_t)(r + q) * nbuckets + d] = run; run += v[q]; }
  }
  if (chunk == DSCAN_CHUNKS - 1) dtot[d] = run;
}

__global__ void __launch_bounds__(1024)
fast_digit_base_kernel(uint32_t* __restrict__ dtot, uint32_t nbuckets) {   // nbuckets <= 2048: two per thread
  __shared__ uint32_t s_w[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t a = 2u * tid < nbuckets ? dtot[2 * tid] : 0u, b = 2u * tid + 1u < nbuckets ? dtot[2 * tid + 1] : 0u;
  uint32_t inc = a + b;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  uint32_t run = inc - (a + b);
  for (int w2 = 0; w2 < warp; w2++) run += s_w[w2];
  if (2u * tid < nbuckets) dtot[2 * tid] = run;
  if (2u * tid + 1u < nbuckets) dtot[2 * tid + 1] = run + a;
}

// dynamic smem: gadj[nbuckets] u32 | { stage[SORT_TILE] u64  over  loc_off[nbuckets] u32 | cnt[SORT_WARPS][nbuckets] u16 | tag[SORT_WARPS][nbuckets] u8 }
// Ranking: a warp visits its 512 consecutive keys 32 at a time (stable).  A round's 32 digits are almost always distinct
// (<= 2048 buckets), so every lane marks its digit's tag with its lane id and reads it back: when nobody was overwritten
// the round is a plain load + store of the warp's counters; only a round with a repeated digit pays for __match_any_sync
// (whose cost grows with the number of distinct values: it was 45 % of this kernel's stall samples when used every round).
template <int MINB>
__global__ void __launch_bounds__(SORT_T, MINB)
fast_scatter_kernel(const unsigned long long* __restrict__ keys, uint64_t n0, const uint32_t* __restrict__ n_dev, int shift, uint32_t nbuckets,
                    const uint32_t* __restrict__ offs, const uint32_t* __restrict__ dbase, uint32_t nb, unsigned long long* __restrict__ out) {
  extern __shared__ __align__(8) unsigned char sraw[];
  uint32_t* gadj = reinterpret_cast<uint32_t*>(sraw);
  unsigned char* ubase = sraw + (size_t)nbuckets * 4;
  unsigned long long* stage = reinterpret_cast<unsigned long long*>(ubase);
  uint32_t* loc_off = reinterpret_cast<uint32_t*>(ubase);
  uint16_t* cnt = reinterpret_cast<uint16_t*>(loc_off + nbuckets);
  uint8_t* tag = reinterpret_cast<uint8_t*>(cnt + (size_t)SORT_WARPS * nbuckets);
  __shared__ uint32_t s_wsum[SORT_WARPS];
  __shared__ uint32_t s_total;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint64_t n = n_dev ? (uint64_t)*n_dev : n0;
  const uint64_t tile0 = (uint64_t)blockIdx.x * SORT_TILE;
  if (tile0 >= n) return;
  for (uint32_t i = tid; i < SORT_WARPS * nbuckets / 2; i += SORT_T) reinterpret_cast<uint32_t*>(cnt)[i] = 0u;
  // each warp owns 512 consecutive keys.  Only their digits stay in registers (two per register, 0xFFFF = no key): the
  // keys themselves are read again from L2 when they are staged, so the ranking rounds run with ~48 registers instead
  // of 80 + spills (local-memory traffic was 37 % of this kernel's sectors) and more blocks fit an SM
  const uint64_t wbase = tile0 + (uint64_t)warp * (32 * SORT_ITEMS);
  uint32_t dig2[SORT_ITEMS / 2];
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i += 2) {
    const uint64_t p0 = wbase + (uint64_t)i * 32 + lane, p1 = p0 + 32;
    const unsigned long long k0 = p0 < n ? keys[p0] : KEY_NONE, k1 = p1 < n ? keys[p1] : KEY_NONE;
    const uint32_t e0 = k0 != KEY_NONE ? ((uint32_t)(k0 >> shift) & (nbuckets - 1u)) : 0xFFFFu;
    const uint32_t e1 = k1 != KEY_NONE ? ((uint32_t)(k1 >> shift) & (nbuckets - 1u)) : 0xFFFFu;
    dig2[i >> 1] = e0 | (e1 << 16);
  }
  __syncthreads();
  uint32_t rank2[SORT_ITEMS / 2];   // two u16 ranks per register
  uint16_t* wc = cnt + (size_t)warp * nbuckets;
  uint8_t* wt = tag + (size_t)warp * nbuckets;
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const uint32_t d = (dig2[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
    const bool in = d != 0xFFFFu;
    uint32_t r = 0;
    int same = 0;
    __match_all_sync(0xffffffffu, d, &same);
    if (same) {                                // one digit for the whole round (sorted input: the usual case after pass 0)
      if (in) {
        uint32_t old = 0;
        if (lane == 0) { old = wc[d]; wc[d] = (uint16_t)(old + 32u); }
        r = __shfl_sync(0xffffffffu, old, 0) + lane;
      }
    } else {
      if (in) wt[d] = (uint8_t)lane;
      __syncwarp();
      const bool lost = in && wt[d] != (uint8_t)lane;
      if (__any_sync(0xffffffffu, lost)) {      // a digit occurs twice in this round
        unsigned peers;
        if (nbuckets <= 256u) {                 // narrow digits repeat in most rounds: eight ballots (2.3 SM cycles each) beat
          peers = __ballot_sync(0xffffffffu, in) ^ (in ? 0u : 0xffffffffu);   // __match_any_sync (2 cycles per distinct value)
#pragma unroll
          for (int bq = 0; bq < 8; bq++) { const unsigned v = __ballot_sync(0xffffffffu, (d >> bq) & 1u); peers &= ((d >> bq) & 1u) ? v : ~v; }
        } else peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (in && lane == leader) { old = wc[d]; wc[d] = (uint16_t)(old + __popc(peers)); }
        old = __shfl_sync(0xffffffffu, old, leader);
        r = old + __popc(peers & ((1u << lane) - 1u));
      } else if (in) { r = wc[d]; wc[d] = (uint16_t)(r + 1u); }
    }
    if (i & 1) rank2[i >> 1] |= r << 16; else rank2[i >> 1] = r;
    __syncwarp();
  }
  __syncthreads();
  // per digit: exclusive prefix over the warps; block totals into loc_off
  for (uint32_t d = tid; d < nbuckets; d += SORT_T) {
    uint32_t run = 0;
#pragma unroll
    for (int w2 = 0; w2 < SORT_WARPS; w2++) { const uint32_t c2 = cnt[(size_t)w2 * nbuckets + d]; cnt[(size_t)w2 * nbuckets + d] = (uint16_t)run; run += c2; }
    loc_off[d] = run;
  }
  __syncthreads();
  // exclusive scan of the block totals over the digits (each thread a contiguous run of nbuckets / SORT_T digits)
  {
    const uint32_t per = (nbuckets + SORT_T - 1) / SORT_T;
    const uint32_t d0 = tid * per;
    uint32_t sum = 0;
    for (uint32_t d = d0; d < d0 + per && d < nbuckets; d++) sum += loc_off[d];
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    uint32_t wb = 0;
    for (int w2 = 0; w2 < warp; w2++) wb += s_wsum[w2];
    if (tid == SORT_T - 1) s_total = wb + inc;
    uint32_t run = wb + inc - sum;
    for (uint32_t d = d0; d < d0 + per && d < nbuckets; d++) {
      const uint32_t c2 = loc_off[d];
      loc_off[d] = run;
      gadj[d] = dbase[d] + offs[(uint64_t)blockIdx.x * nbuckets + d] - run;     // global position = gadj[d] + position in the staged tile
      run += c2;
    }
  }
  __syncthreads();
  // position of every key in the staged tile, then the stage takes over the counters' memory
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const uint32_t d = (dig2[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
    if (d != 0xFFFFu) {
      const uint32_t r = loc_off[d] + wc[d] + ((rank2[i >> 1] >> (16 * (i & 1))) & 0xFFFFu);
      rank2[i >> 1] = (i & 1) ? ((rank2[i >> 1] & 0xFFFFu) | (r << 16)) : ((rank2[i >> 1] & 0xFFFF0000u) | r);
    }
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < SORT_ITEMS; h += 4) {      // four keys back from L2 at a time
    unsigned long long kk[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint64_t p = wbase + (uint64_t)(h + q) * 32 + lane;
      kk[q] = p < n ? keys[p] : KEY_NONE;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int i = h + q;
      if (((dig2[i >> 1] >> (16 * (i & 1))) & 0xFFFFu) != 0xFFFFu) stage[(rank2[i >> 1] >> (16 * (i & 1))) & 0xFFFFu] = kk[q];
    }
  }
  __syncthreads();
  const uint32_t total = s_total;
  for (uint32_t p = tid; p < total; p += SORT_T) {
    const unsigned long long kk = stage[p];
    out[gadj[(uint32_t)(kk >> shift) & (nbuckets - 1u)] + p] = kk;
  }
}

__global__ void fast_heads_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ n_dev, uint32_t idx_bits,
                                  uint32_t* __restrict__ flags, uint64_t n0) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n0) return;
  const uint64_t n = *n_dev;
  flags[i] = (i < n && (i == 0 || (keys[i] >> idx_bits) != (keys[i - 1] >> idx_bits))) ? 1u : 0u;
}

// as build_csr_kernel (kmer_build.cu), from the packed keys
__global__ void fast_csr_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ escan, uint64_t n, uint32_t idx_bits,
                                uint32_t slots, uint32_t n_codes, uint64_t* __restrict__ codes, uint32_t* __restrict__ post_off,
                                uint32_t* __restrict__ postings, uint32_t* __restrict__ fwd_ids, const uint16_t* __restrict__ seg_part,
                                uint32_t uniform_parts, uint32_t* __restrict__ list_part) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) post_off[n_codes] = (uint32_t)n;
  if (i >= n) return;
  const unsigned long long key = __ldcs(&keys[i]);
  const unsigned long long code = key >> idx_bits;
  const uint32_t rec = (uint32_t)(key & ((1ull << idx_bits) - 1ull));
  const bool head = (i == 0 || (keys[i - 1] >> idx_bits) != code);
  const uint32_t cid = escan[i] + (head ? 1u : 0u) - 1u;
  const uint32_t seg = rec / slots;
  __stcs(&postings[i], seg);        // streamed once: leave the L2 to the scattered forward-index lines
  fwd_ids[rec] = cid;
  const uint32_t part = uniform_parts ? seg % uniform_parts : (uint32_t)seg_part[seg];
  if (head) {
    codes[cid] = code; post_off[cid] = (uint32_t)i;
    atomicOr(&list_part[cid], part);
  } else {
    const uint32_t pseg = (uint32_t)(keys[i - 1] & ((1ull << idx_bits) - 1ull)) / slots;
    const uint32_t ppart = uniform_parts ? pseg % uniform_parts : (uint32_t)seg_part[pseg];
    if (ppart != part) atomicOr(&list_part[cid], 0x80000000u);
  }
}

size_t scatter_smem(uint32_t nbk) {   // gadj + max(stage, loc_off + cnt + tag)
  return (size_t)nbk * 4 + std::max<size_t>((size_t)SORT_TILE * 8, (size_t)nbk * 4 + (size_t)SORT_WARPS * nbk * 3);
}

uint32_t bits_needed(uint64_t v) { uint32_t b = 1; while (b < 64 && (v >> b) != 0ull) b++; return b; }

struct FastDir {
  unsigned long long* ka = nullptr; unsigned long long* kb = nullptr; uint32_t* hist = nullptr; uint32_t* flags = nullptr;
  uint32_t* d_cnt = nullptr;   // [0] R  [1] n_codes
  uint32_t* dtot = nullptr;    // [nbuckets] digit totals -> digit bases of the current pass
  uint32_t* csum = nullptr;    // [DSCAN_CHUNKS][nbuckets] row-chunk sums of the block histograms
};

}  // namespace

bool msspe_build_fast_applicable(const msspe_ctx* c) {
  if (getenv("MSSPE_BUILD_LEGACY")) return false;
  const uint64_t GS = c->n_segments * (uint64_t)c->slots;
  if (GS == 0 || c->slots == 0) return false;
  const uint32_t idx_bits = bits_needed(GS);
  return c->cfg.kmer_size <= 16 && 2 * c->cfg.kmer_size + idx_bits <= 63 && GS < 0xFFFFFFFFull;
}

// Both directions; one host synchronisation in total (R and the number of distinct words of both directions).
int msspe_build_fast(msspe_ctx* c) {
  cudaStream_t st = c->stream;
  const uint64_t G = c->n_segments;
  const uint32_t slots = c->slots, w = c->cfg.search_windows_size, k = c->cfg.kmer_size;
  const uint64_t GS = G * slots;
  const uint32_t idx_bits = bits_needed(GS);
  const uint32_t code_bits = 2 * k;
  // digit width: <= 11 bits (three passes for k <= 16) or <= 8 bits (four passes, but a 4096-key tile then leaves in runs of
  // ~16 keys = 128 B per digit instead of ~4 keys = 32 B); MSSPE_SORT_BITS picks, see DESIGN.md section 4
  const int max_digit = getenv("MSSPE_SORT_BITS") ? std::min(11, std::max(4, atoi(getenv("MSSPE_SORT_BITS")))) : 11;
  const int passes = (int)((code_bits + max_digit - 1) / max_digit);
  SortPass pass[8];
  {
    int left = (int)code_bits, shift = (int)idx_bits;
    for (int p = 0; p < passes; p++) { const int b = (left + (passes - p) - 1) / (passes - p); pass[p].shift = shift; pass[p].bits = b; shift += b; left -= b; }
  }
  const uint32_t nb = (uint32_t)div_up_u64(GS, SORT_TILE);
  const uint32_t uni = (c->uniform_parts && c->uniform_parts <= 65536u) ? c->uniform_parts : 0u;
  const uint32_t slots_pad = (slots + 1u) & ~1u;
  const size_t enc_smem = ((size_t)slots_pad * 8 + ((w + 7u) & ~7u)) * ENCF_WARPS;
  if (enc_smem > 48 * 1024) {
    MSSPE_CUDA_TRY(c, cudaFuncSetAttribute(encode_keys_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc_smem));
    MSSPE_CUDA_TRY(c, cudaFuncSetAttribute(encode_keys_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc_smem));
  }
  int maxbits = 1;
  for (int p = 0; p < passes; p++) maxbits = std::max(maxbits, pass[p].bits);
  const uint32_t max_buckets = 1u << maxbits;
  const size_t sc_smem_max = scatter_smem(max_buckets);
  static const int scatter_minb = getenv("MSSPE_SCATTER_MINB") ? atoi(getenv("MSSPE_SCATTER_MINB")) : 4;
  MSSPE_CUDA_TRY(c, cudaFuncSetAttribute(fast_scatter_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem_max));
  MSSPE_CUDA_TRY(c, cudaFuncSetAttribute(fast_scatter_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem_max));
  FastDir F[2];
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[2], st));
  for (int dir = 0; dir < 2; dir++) {
    FastDir& f = F[dir];
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&f.ka, GS * 8, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&f.kb, GS * 8, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&f.hist, (uint64_t)max_buckets * nb * 4, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&f.flags, GS * 4, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&f.d_cnt, 16, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&f.dtot, (uint64_t)max_buckets * 4, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&f.csum, (uint64_t)DSCAN_CHUNKS * max_buckets * 4, st));
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(f.d_cnt, 0, 16, st));
    {
      KPROF(c, KP_ENCODE, st, G * w + GS * 8)
      if (w <= 64u) {
        const unsigned grid = (unsigned)div_up_u64(G, ENCF_WARPS * ENC_SEGS);
        if (dir == 0) encode_keys_packed_kernel<0><<<grid, ENCF_WARPS * 32, 0, st>>>(c->d_bases, c->d_offsets, c->d_seg_base, c->n_records, c->uniform_parts, G, c->cfg.window_size,
                                                                                     c->cfg.overlap_size, w, k, slots, idx_bits, f.ka, c->d_seg_part, c->d_seg_rec);
        else encode_keys_packed_kernel<1><<<grid, ENCF_WARPS * 32, 0, st>>>(c->d_bases, c->d_offsets, c->d_seg_base, c->n_records, c->uniform_parts, G, c->cfg.window_size,
                                                                            c->cfg.overlap_size, w, k, slots, idx_bits, f.ka, c->d_seg_part, c->d_seg_rec);
      } else {
        const unsigned grid = (unsigned)div_up_u64(G, ENCF_WARPS);
        if (dir == 0) encode_keys_kernel<0><<<grid, ENCF_WARPS * 32, enc_smem, st>>>(c->d_bases, c->d_offsets, c->d_seg_base, c->n_records, c->uniform_parts, G, c->cfg.window_size,
                                                                                    c->cfg.overlap_size, w, k, slots, idx_bits, f.ka, c->d_seg_part, c->d_seg_rec);
        else encode_keys_kernel<1><<<grid, ENCF_WARPS * 32, enc_smem, st>>>(c->d_bases, c->d_offsets, c->d_seg_base, c->n_records, c->uniform_parts, G, c->cfg.window_size,
                                                                            c->cfg.overlap_size, w, k, slots, idx_bits, f.ka, c->d_seg_part, c->d_seg_rec);
      }
    }
    for (int p = 0; p < passes; p++) {
      const uint32_t nbk = 1u << pass[p].bits;
      const uint32_t* n_dev = p == 0 ? nullptr : f.d_cnt;
      { KPROF(c, KP_SORT_HIST, st, GS * 8)
        fast_hist_kernel<<<nb, SORT_T, nbk * 4, st>>>(f.ka, GS, n_dev, pass[p].shift, nbk, f.hist, nb, p == 0 ? f.d_cnt : nullptr); }
      { KPROF(c, KP_SCAN, st, (uint64_t)nbk * nb * 8)
        fast_digit_partial_kernel<<<dim3((nbk + 31u) / 32u, DSCAN_SPLIT), DSCAN_WARPS * 32, 0, st>>>(f.hist, nb, nbk, f.csum);
        fast_digit_scan_kernel<<<dim3((nbk + 31u) / 32u, DSCAN_SPLIT), DSCAN_WARPS * 32, 0, st>>>(f.hist, nb, nbk, f.csum, f.dtot);
        fast_digit_base_kernel<<<1, 1024, 0, st>>>(f.dtot, nbk); }
      const size_t sm = scatter_smem(nbk);
      { KPROF(c, KP_SORT_SCATTER, st, GS * 16)
        if (scatter_minb == 4) fast_scatter_kernel<4><<<nb, SORT_T, sm, st>>>(f.ka, GS, n_dev, pass[p].shift, nbk, f.hist, f.dtot, nb, f.kb);
        else fast_scatter_kernel<3><<<nb, SORT_T, sm, st>>>(f.ka, GS, n_dev, pass[p].shift, nbk, f.hist, f.dtot, nb, f.kb); }
      std::swap(f.ka, f.kb);
    }
    { KPROF(c, KP_CSR, st, GS * 12)
      fast_heads_kernel<<<(unsigned)div_up_u64(GS, 256), 256, 0, st>>>(f.ka, f.d_cnt, idx_bits, f.flags, GS); }
    int rc = msspe_exclusive_scan_u32(c, f.flags, f.flags, GS, f.d_cnt + 1, st);
    if (rc) return rc;
  }
  uint32_t h_cnt[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  for (int dir = 0; dir < 2; dir++) MSSPE_CUDA_TRY(c, cudaMemcpyAsync(h_cnt[dir], F[dir].d_cnt, 8, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[3], st));
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  for (int dir = 0; dir < 2; dir++) {
    FastDir& f = F[dir];
    DirIndex& D = c->dir[dir];
    const uint32_t R = h_cnt[dir][0], n_codes = h_cnt[dir][1];
    D.n_records = R; D.n_codes = n_codes;
    const uint64_t Dn = n_codes ? n_codes : 1, Rn = R ? R : 1;
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.fwd_ids, (GS ? GS : 1) * 4, st));
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.fwd_ids, 0xFF, (GS ? GS : 1) * 4, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.codes, Dn * 8, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.post_off, (Dn + 1) * 4, st));
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.post_off, 0, (Dn + 1) * 4, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.postings, Rn * 4, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.list_part, Dn * 4, st));
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.list_part, 0, Dn * 4, st));
    if (R) {
      KPROF(c, KP_CSR, st, (uint64_t)R * 20)
      fast_csr_kernel<<<(unsigned)div_up_u64(R, 256), 256, 0, st>>>(f.ka, f.flags, R, idx_bits, slots, n_codes, D.codes, D.post_off, D.postings, D.fwd_ids,
                                                                   c->d_seg_part, uni, D.list_part);
    }
    MSSPE_CUDA_TRY(c, cudaFreeAsync(f.ka, st)); MSSPE_CUDA_TRY(c, cudaFreeAsync(f.kb, st)); MSSPE_CUDA_TRY(c, cudaFreeAsync(f.hist, st));
    MSSPE_CUDA_TRY(c, cudaFreeAsync(f.flags, st)); MSSPE_CUDA_TRY(c, cudaFreeAsync(f.d_cnt, st)); MSSPE_CUDA_TRY(c, cudaFreeAsync(f.dtot, st)); MSSPE_CUDA_TRY(c, cudaFreeAsync(f.csum, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.freq, Dn * 4, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.ignored, (div_up_u64(G, 32) + 1) * 4, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.cov, 65536 * 4, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.pmark, 2048 * 4, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.ctl, sizeof(SelectCtl), st));
    int rc = msspe_select_prepare_static(c, dir, st);
    if (rc) return rc;
  }
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[4], st));
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  float a = 0.f, b = 0.f;
  MSSPE_CUDA_TRY(c, cudaEventElapsedTime(&a, c->ev[2], c->ev[3]));
  MSSPE_CUDA_TRY(c, cudaEventElapsedTime(&b, c->ev[3], c->ev[4]));
  c->timing.encode_ms = 0.f;         // K1 and the sort are interleaved per direction: one figure for the build
  c->timing.index_ms = a + b;
  return MSSPE_OK;
}
