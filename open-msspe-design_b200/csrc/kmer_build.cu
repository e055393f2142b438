// kmer_build.cu -- K1 (window slicing + 2-bit encoding + per-window de-duplication) and K2 (stable LSD radix
// sort -> CSR inverted index + forward index).  Replaces get_segment_manager (od-msspe/src/main.rs:196-235:
// partitioning_sequence :173-181, get_sequence_on_search_windows :183-187, find_kmers :163-171,
// reverse_complement :148-161) and make_kmer_segments_windows_mapping (:237-255).
//
// Layout in HBM (per direction d):
//   rec_code[R] u64, rec_idx[R] u32   valid records in (segment, slot) order (transient, sort input)
//   codes[D] u64 ascending            distinct words
//   post_off[D+1] u32, postings[R] u32  CSR: segments containing each word, ascending
//   fwd_ids[G*s] u32                  forward index: code id of (segment, slot) or 0xFFFFFFFF
// Only the first/last w bases of each W-window are read: 2w/S of the genome bytes (40 % at 500/250/50).
#include "engine.cuh"

namespace {

constexpr int ENC_WARPS = 8;

__device__ __forceinline__ uint32_t base2(uint8_t c) {
  // to_records upper-cases and maps U->T (main.rs:117-118); find_kmers keeps only "ATCGU" (main.rs:167)
  switch (c) {
    case 'A': case 'a': return 0u;
    case 'C': case 'c': return 1u;
    case 'G': case 'g': return 2u;
    case 'T': case 't': case 'U': case 'u': return 3u;
    default: return 4u;
  }
}

enum EncMode { ENC_COUNT = 0, ENC_EMIT = 1, ENC_DENSE = 2 };

// One warp per segment, one direction per launch (template DIR).  Dynamic smem per warp: w bytes + s u64.
template <int MODE, int DIR>
__global__ void __launch_bounds__(ENC_WARPS * 32)
encode_windows_kernel(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets,
                      const uint64_t* __restrict__ seg_base, uint32_t n_records, uint32_t uniform_parts, uint64_t n_segments, uint32_t W,
                      uint32_t S, uint32_t w, uint32_t k, uint32_t slots, uint32_t* __restrict__ counts,
                      const uint32_t* __restrict__ rec_off, uint64_t* __restrict__ rec_code,
                      uint32_t* __restrict__ rec_idx, uint64_t* __restrict__ dense, uint16_t* __restrict__ seg_part,
                      uint32_t* __restrict__ seg_rec) {
  extern __shared__ __align__(8) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t slots_pad = (slots + 1u) & ~1u;
  const size_t per_warp = (size_t)slots_pad * 8 + ((w + 7u) & ~7u);
  uint64_t* scode = reinterpret_cast<uint64_t*>(smem_raw + warp * per_warp);
  uint8_t* sb = reinterpret_cast<uint8_t*>(scode + slots_pad);
  const uint64_t g = (uint64_t)blockIdx.x * ENC_WARPS + warp;
  if (g >= n_segments) return;
  // record of this segment: last r with seg_base[r] <= g
  uint32_t lo = 0, hi = n_records;  // invariant: seg_base[lo] <= g < seg_base[hi]
  if (uniform_parts) {              // equal-length alignment: no search (14 dependent loads at 12,500 records)
    lo = (uint32_t)(g / uniform_parts);
  } else {
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (seg_base[mid] <= g) lo = mid; else hi = mid;
    }
  }
  const uint64_t j = g - seg_base[lo];
  const uint64_t win = offsets[lo] + j * (uint64_t)S;
  const uint64_t src = DIR == 0 ? win : win + (W - w);
  for (uint32_t t = lane; t < w; t += 32) sb[t] = (uint8_t)base2(bases[src + t]);
  if (MODE == ENC_COUNT && DIR == 0 && lane == 0) { seg_part[g] = (uint16_t)j; seg_rec[g] = lo; }
  __syncwarp();
  // slot codes
  for (uint32_t q = lane; q < slots; q += 32) {
    uint64_t code = 0; bool ok = true;
    if (DIR == 0) {
      for (uint32_t t = 0; t < k; t++) { uint32_t b = sb[q + t]; ok &= (b < 4u); code = (code << 2) | (b & 3u); }
    } else {  // reverse complement of tail[q..q+k): first output base = complement of the last input base
      for (uint32_t t = 0; t < k; t++) { uint32_t b = sb[q + k - 1 - t]; ok &= (b < 4u); code = (code << 2) | ((3u - b) & 3u); }
    }
    scode[q] = ok ? code : MSSPE_NO_KMER;
  }
  __syncwarp();
  // itertools unique(): keep the first occurrence inside the window
  uint32_t base_rank = 0;
  for (uint32_t q0 = 0; q0 < slots; q0 += 32) {
    const uint32_t q = q0 + lane;
    bool keep = false; uint64_t code = MSSPE_NO_KMER;
    if (q < slots) {
      code = scode[q];
      keep = code != MSSPE_NO_KMER;
      for (uint32_t p = 0; keep && p < q; p++) keep = scode[p] != code;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (MODE == ENC_EMIT) {
      if (keep) {
        const uint64_t pos = (uint64_t)rec_off[g] + base_rank + __popc(m & ((1u << lane) - 1u));
        rec_code[pos] = code;
        rec_idx[pos] = (uint32_t)(g * slots + q);
      }
    } else if (MODE == ENC_DENSE) {
      if (q < slots) dense[g * slots + q] = keep ? code : MSSPE_NO_KMER;
    }
    base_rank += __popc(m);
  }
  if (MODE == ENC_COUNT && lane == 0) counts[g] = base_rank;
}

// ---- exclusive scan (u32) : block-local scan + recursive scan of block sums ----
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS)
scan_block_kernel(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* bsum) {  // in may alias out
  __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = base + i < n ? in[base + i] : 0u; sum += v[i]; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t ws = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0u, wi = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - ws;
    if (lane == SCAN_THREADS / 32 - 1 && bsum) bsum[blockIdx.x] = wi;
  }
  __syncthreads();
  uint32_t run = warp_sums[warp] + inc - sum;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i < n) out[base + i] = run; run += v[i]; }
}

__global__ void scan_add_kernel(uint32_t* __restrict__ out, uint64_t n, const uint32_t* __restrict__ boff) {
  const uint64_t i = (uint64_t)blockIdx.x * SCAN_TILE + threadIdx.x;
  const uint32_t add = boff[blockIdx.x];
#pragma unroll
  for (int t = 0; t < SCAN_ITEMS; t++) { uint64_t p = i + (uint64_t)t * SCAN_THREADS; if (p < n) out[p] += add; }
}

__global__ void scan_total_kernel(const uint32_t* in_last, const uint32_t* out_last, uint32_t* total) {
  *total = *in_last + *out_last;
}

}  // namespace

// out[i] = sum_{t<i} in[i]; *d_total (optional, device) = sum of all.  in may alias out.
int msspe_exclusive_scan_u32(msspe_ctx* ctx, const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* d_total, cudaStream_t st) {
  if (n == 0) { if (d_total) MSSPE_CUDA_TRY(ctx, cudaMemsetAsync(d_total, 0, 4, st)); return MSSPE_OK; }
  const uint64_t nb = div_up_u64(n, SCAN_TILE);
  uint32_t* bsum = nullptr;
  uint32_t* last_in = nullptr;
  if (d_total) {  // in may alias out: save in[n-1] first
    MSSPE_CUDA_TRY(ctx, cudaMallocAsync(&last_in, 4, st));
    MSSPE_CUDA_TRY(ctx, cudaMemcpyAsync(last_in, in + (n - 1), 4, cudaMemcpyDeviceToDevice, st));
  }
  if (nb > 1) MSSPE_CUDA_TRY(ctx, cudaMallocAsync(&bsum, nb * sizeof(uint32_t), st));
  { KPROF(ctx, KP_SCAN, st, n * 8) scan_block_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, n, bsum); }
  if (nb > 1) {
    int rc = msspe_exclusive_scan_u32(ctx, bsum, bsum, nb, nullptr, st);
    if (rc) return rc;
    { KPROF(ctx, KP_SCAN, st, n * 8) scan_add_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(out, n, bsum); }
    MSSPE_CUDA_TRY(ctx, cudaFreeAsync(bsum, st));
  }
  if (d_total) {
    { KPROF(ctx, KP_SCAN, st, 8) scan_total_kernel<<<1, 1, 0, st>>>(last_in, out + (n - 1), d_total); }
    MSSPE_CUDA_TRY(ctx, cudaFreeAsync(last_in, st));
  }
  MSSPE_CUDA_TRY(ctx, cudaGetLastError());
  return MSSPE_OK;
}

namespace {

// ---- stable LSD radix sort, 8-bit digits, u64 keys + u32 payload ----
constexpr int RS_THREADS = 256, RS_WARPS = RS_THREADS / 32, RS_ITEMS = 16;
constexpr int RS_WARP_TILE = 32 * RS_ITEMS, RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void __launch_bounds__(RS_THREADS)
radix_hist_kernel(const uint64_t* __restrict__ keys, uint64_t n, int shift, uint32_t* __restrict__ hist, uint32_t nb) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
  for (int i = 0; i < RS_ITEMS; i++) {
    const uint64_t p = base + (uint64_t)i * RS_THREADS + threadIdx.x;
    if (p < n) atomicAdd(&h[(keys[p] >> shift) & 0xFFu], 1u);
  }
  __syncthreads();
  hist[(uint64_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS)
radix_scatter_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t n, int shift,
                     const uint32_t* __restrict__ offs, uint32_t nb, uint64_t* __restrict__ out_keys,
                     uint32_t* __restrict__ out_vals) {
  __shared__ uint32_t cnt[RS_WARPS][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0;
  __syncthreads();
  // each warp owns a contiguous run of 512 keys, visited in order, 32 at a time => stable
  const uint64_t wbase = (uint64_t)blockIdx.x * RS_TILE + (uint64_t)warp * RS_WARP_TILE;
  uint64_t key[RS_ITEMS];
  uint32_t val[RS_ITEMS], rank[RS_ITEMS];
  unsigned peers[RS_ITEMS];
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {  // all loads and all digit matches first: independent, their latencies overlap
    const uint64_t p = wbase + (uint64_t)i * 32 + lane;
    const bool in = p < n;
    key[i] = in ? keys[p] : 0ull;
    val[i] = in ? vals[p] : 0u;
  }
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    const bool in = wbase + (uint64_t)i * 32 + lane < n;
    const uint32_t d = in ? (uint32_t)((key[i] >> shift) & 0xFFu) : 0x100u;  // 0x100: out-of-range lanes group apart
    peers[i] = __match_any_sync(0xffffffffu, d);
  }
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {  // the counters advance in record order (stability): this part stays sequential
    const bool in = wbase + (uint64_t)i * 32 + lane < n;
    const uint32_t d = (uint32_t)((key[i] >> shift) & 0xFFu);
    const int leader = __ffs(peers[i]) - 1;
    uint32_t old = 0;
    if (in && lane == leader) { old = cnt[warp][d]; cnt[warp][d] = old + __popc(peers[i]); }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[i] = old + __popc(peers[i] & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  {  // thread d: exclusive prefix of digit d over the warps of this block, plus the global base
    const int d = threadIdx.x;
    uint32_t run = offs[(uint64_t)d * nb + blockIdx.x];
#pragma unroll
    for (int w2 = 0; w2 < RS_WARPS; w2++) { uint32_t c = cnt[w2][d]; cnt[w2][d] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    const uint64_t p = wbase + (uint64_t)i * 32 + lane;
    if (p < n) {
      const uint32_t d = (uint32_t)((key[i] >> shift) & 0xFFu);
      const uint32_t pos = cnt[warp][d] + rank[i];
      out_keys[pos] = key[i];
      out_vals[pos] = val[i];
    }
  }
}

__global__ void mark_heads_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t* __restrict__ flags) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// flags_scan[i] = exclusive scan of head flags; head(i) <=> (i == 0 || keys differ)
// Also list_part[cid] (zeroed by the caller): the partition_no (main.rs:227) of the list's postings if they all share
// one -- the rule in a pre-aligned alignment, where a word sits in the same column of every genome -- else bit 31 set.
// partition_tie_score (main.rs:261-283) of such a list is 1/(partition_coverage[p] + 1) whichever postings are live.
__global__ void build_csr_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx,
                                 const uint32_t* __restrict__ escan, uint64_t n, uint32_t slots, uint32_t n_codes,
                                 uint64_t* __restrict__ codes, uint32_t* __restrict__ post_off,
                                 uint32_t* __restrict__ postings, uint32_t* __restrict__ fwd_ids,
                                 const uint16_t* __restrict__ seg_part, uint32_t uniform_parts, uint32_t* __restrict__ list_part) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) post_off[n_codes] = (uint32_t)n;
  if (i >= n) return;
  const bool head = (i == 0 || keys[i] != keys[i - 1]);
  const uint32_t cid = escan[i] + (head ? 1u : 0u) - 1u;
  const uint32_t rec = idx[i];
  const uint32_t seg = rec / slots;
  postings[i] = seg;
  fwd_ids[rec] = cid;
  const uint32_t part = uniform_parts ? seg % uniform_parts : (uint32_t)seg_part[seg];
  if (head) {
    codes[cid] = keys[i]; post_off[cid] = (uint32_t)i;
    atomicOr(&list_part[cid], part);
  } else {
    const uint32_t pseg = idx[i - 1] / slots;
    const uint32_t ppart = uniform_parts ? pseg % uniform_parts : (uint32_t)seg_part[pseg];
    if (ppart != part) atomicOr(&list_part[cid], 0x80000000u);
  }
}

}  // namespace

// Stable LSD radix sort of (u64 key, u32 payload) pairs by the low `key_bits` bits; ping-pongs between the a and b
// buffers and leaves the result in *key_a / *val_a (the pointers are swapped as needed).
int msspe_radix_sort_pairs(msspe_ctx* c, uint64_t** key_a, uint32_t** val_a, uint64_t** key_b, uint32_t** val_b, uint64_t n, uint32_t key_bits,
                           cudaStream_t st) {
  if (n == 0) return MSSPE_OK;
  const uint32_t nb = (uint32_t)div_up_u64(n, RS_TILE);
  uint32_t* hist = nullptr;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&hist, (uint64_t)256 * nb * sizeof(uint32_t), st));
  const int passes = (int)((key_bits + 7) / 8);
  for (int p = 0; p < passes; p++) {
    { KPROF(c, KP_SORT_HIST, st, n * 8) radix_hist_kernel<<<nb, RS_THREADS, 0, st>>>(*key_a, n, 8 * p, hist, nb); }
    int rc = msspe_exclusive_scan_u32(c, hist, hist, (uint64_t)256 * nb, nullptr, st);
    if (rc) return rc;
    { KPROF(c, KP_SORT_SCATTER, st, n * 24) radix_scatter_kernel<<<nb, RS_THREADS, 0, st>>>(*key_a, *val_a, n, 8 * p, hist, nb, *key_b, *val_b); }
    std::swap(*key_a, *key_b); std::swap(*val_a, *val_b);
  }
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  MSSPE_CUDA_TRY(c, cudaFreeAsync(hist, st));
  return MSSPE_OK;
}

namespace {

void free_dir(msspe_ctx* c, DirIndex& d) {
  msspe_dev_free(c, d.codes); msspe_dev_free(c, d.post_off); msspe_dev_free(c, d.postings); msspe_dev_free(c, d.fwd_ids); msspe_dev_free(c, d.list_part); msspe_dev_free(c, d.freq);
  msspe_dev_free(c, d.acc); msspe_dev_free(c, d.ignored); msspe_dev_free(c, d.cov); msspe_dev_free(c, d.pmark); msspe_dev_free(c, d.ctl); msspe_dev_free(c, d.out);
  msspe_dev_free(c, d.tile_first);
  msspe_dev_free(c, d.pv_ucode_off); msspe_dev_free(c, d.pv_ucodes); msspe_dev_free(c, d.pv_fwdl); msspe_dev_free(c, d.pv_useg_off); msspe_dev_free(c, d.pv_usegs);
  msspe_dev_free(c, d.s_postings); msspe_dev_free(c, d.s_off); msspe_dev_free(c, d.s_id); msspe_dev_free(c, d.s_tile_first); msspe_dev_free(c, d.s_cost);
  d = DirIndex();
}

template <int DIR>
int launch_encode(msspe_ctx* c, int mode, uint32_t* counts, const uint32_t* rec_off, uint64_t* rec_code,
                  uint32_t* rec_idx, uint64_t* dense, cudaStream_t st) {
  const uint32_t slots = c->slots, w = c->cfg.search_windows_size;
  const uint32_t slots_pad = (slots + 1u) & ~1u;
  const size_t per_warp = (size_t)slots_pad * 8 + ((w + 7u) & ~7u);
  const size_t smem = per_warp * ENC_WARPS;
  const unsigned grid = (unsigned)div_up_u64(c->n_segments, ENC_WARPS);
  if (grid == 0) return MSSPE_OK;
#define ENC_ARGS c->d_bases, c->d_offsets, c->d_seg_base, c->n_records, c->uniform_parts, c->n_segments, c->cfg.window_size,        \
    c->cfg.overlap_size, w, c->cfg.kmer_size, slots, counts, rec_off, rec_code, rec_idx, dense, c->d_seg_part, c->d_seg_rec
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(encode_windows_kernel<ENC_COUNT, DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(encode_windows_kernel<ENC_EMIT, DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(encode_windows_kernel<ENC_DENSE, DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  {
    KPROF(c, KP_ENCODE, st, c->n_segments * (w + (mode == ENC_COUNT ? 4u : slots * (mode == ENC_EMIT ? 12u : 8u))))
    if (mode == ENC_COUNT) encode_windows_kernel<ENC_COUNT, DIR><<<grid, ENC_WARPS * 32, smem, st>>>(ENC_ARGS);
    else if (mode == ENC_EMIT) encode_windows_kernel<ENC_EMIT, DIR><<<grid, ENC_WARPS * 32, smem, st>>>(ENC_ARGS);
    else encode_windows_kernel<ENC_DENSE, DIR><<<grid, ENC_WARPS * 32, smem, st>>>(ENC_ARGS);
  }
#undef ENC_ARGS
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  return MSSPE_OK;
}

int encode_dispatch(msspe_ctx* c, int dir, int mode, uint32_t* counts, const uint32_t* rec_off, uint64_t* rec_code,
                    uint32_t* rec_idx, uint64_t* dense, cudaStream_t st) {
  return dir == 0 ? launch_encode<0>(c, mode, counts, rec_off, rec_code, rec_idx, dense, st)
                  : launch_encode<1>(c, mode, counts, rec_off, rec_code, rec_idx, dense, st);
}

int build_direction(msspe_ctx* c, int dir, float* enc_ms, float* idx_ms) {
  cudaStream_t st = c->stream;
  DirIndex& D = c->dir[dir];
  const uint64_t G = c->n_segments;
  const uint32_t slots = c->slots;
  uint32_t* counts = nullptr; uint32_t* d_total = nullptr;
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[2], st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&counts, (G + 1) * sizeof(uint32_t), st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_total, sizeof(uint32_t), st));
  int rc = encode_dispatch(c, dir, ENC_COUNT, counts, nullptr, nullptr, nullptr, nullptr, st);
  if (rc) return rc;
  rc = msspe_exclusive_scan_u32(c, counts, counts, G, d_total, st);
  if (rc) return rc;
  uint32_t R = 0;
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&R, d_total, 4, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  D.n_records = R;
  uint64_t* key_a = nullptr; uint64_t* key_b = nullptr; uint32_t* val_a = nullptr; uint32_t* val_b = nullptr;
  const uint64_t Ra = R ? R : 1;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&key_a, Ra * 8, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&val_a, Ra * 4, st));
  rc = encode_dispatch(c, dir, ENC_EMIT, nullptr, counts, key_a, val_a, nullptr, st);
  if (rc) return rc;
  MSSPE_CUDA_TRY(c, cudaFreeAsync(counts, st));
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[3], st));
  // ---- K2 ----
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.fwd_ids, (G * slots != 0 ? G * slots : 1) * sizeof(uint32_t), c->stream));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.fwd_ids, 0xFF, (G * slots != 0 ? G * slots : 1) * sizeof(uint32_t), st));
  uint32_t n_codes = 0;
  if (R > 0) {
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&key_b, Ra * 8, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&val_b, Ra * 4, st));
    rc = msspe_radix_sort_pairs(c, &key_a, &val_a, &key_b, &val_b, R, 2 * c->cfg.kmer_size, st);
    if (rc) return rc;
    // sorted records are in key_a / val_a
    uint32_t* flags = nullptr;
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&flags, (uint64_t)R * 4, st));
    { KPROF(c, KP_CSR, st, (uint64_t)R * 12) mark_heads_kernel<<<(unsigned)div_up_u64(R, 256), 256, 0, st>>>(key_a, R, flags); }
    rc = msspe_exclusive_scan_u32(c, flags, flags, R, d_total, st);
    if (rc) return rc;
    MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&n_codes, d_total, 4, cudaMemcpyDeviceToHost, st));
    MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
    D.n_codes = n_codes;
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.codes, (uint64_t)n_codes * 8, c->stream));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.post_off, ((uint64_t)n_codes + 1) * 4, c->stream));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.postings, (uint64_t)R * 4, c->stream));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.list_part, (uint64_t)(n_codes ? n_codes : 1) * 4, c->stream));
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.list_part, 0, (uint64_t)(n_codes ? n_codes : 1) * 4, st));
    { KPROF(c, KP_CSR, st, (uint64_t)R * 24)
      build_csr_kernel<<<(unsigned)div_up_u64(R, 256), 256, 0, st>>>(key_a, val_a, flags, R, slots, n_codes, D.codes,
                                                                  D.post_off, D.postings, D.fwd_ids, c->d_seg_part,
                                                                  (c->uniform_parts && c->uniform_parts <= 65536u) ? c->uniform_parts : 0u, D.list_part); }
    MSSPE_CUDA_TRY(c, cudaGetLastError());
    MSSPE_CUDA_TRY(c, cudaFreeAsync(flags, st));
    MSSPE_CUDA_TRY(c, cudaFreeAsync(key_b, st));
    MSSPE_CUDA_TRY(c, cudaFreeAsync(val_b, st));
  } else {
    D.n_codes = 0;
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.codes, 8, c->stream));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.post_off, 4, c->stream));
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(D.post_off, 0, 4, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.postings, 4, c->stream));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.list_part, 4, c->stream));
  }
  MSSPE_CUDA_TRY(c, cudaFreeAsync(key_a, st));
  MSSPE_CUDA_TRY(c, cudaFreeAsync(val_a, st));
  MSSPE_CUDA_TRY(c, cudaFreeAsync(d_total, st));
  // greedy state
  const uint64_t Dn = n_codes ? n_codes : 1;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.freq, Dn * 4, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.ignored, (div_up_u64(G, 32) + 1) * 4, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.cov, 65536 * 4, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.pmark, 2048 * 4, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.ctl, sizeof(SelectCtl), c->stream));
  rc = msspe_select_prepare_static(c, dir, st);
  if (rc) return rc;
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[4], st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  float a = 0, b = 0;
  MSSPE_CUDA_TRY(c, cudaEventElapsedTime(&a, c->ev[2], c->ev[3]));
  MSSPE_CUDA_TRY(c, cudaEventElapsedTime(&b, c->ev[3], c->ev[4]));
  *enc_ms += a; *idx_ms += b;
  return MSSPE_OK;
}

}  // namespace

int msspe_free_index(msspe_ctx* c) {
  free_dir(c, c->dir[0]); free_dir(c, c->dir[1]);
  if (c->d_seg_part) msspe_dev_free(c, c->d_seg_part);
  if (c->d_seg_rec) msspe_dev_free(c, c->d_seg_rec);
  c->d_seg_part = nullptr; c->d_seg_rec = nullptr;
  c->built = false;
  return MSSPE_OK;
}

extern "C" int msspe_build_index(msspe_ctx* c) {
  if (!c) return MSSPE_ERR_INVALID;
  if (!c->loaded) { c->set_error("msspe_build_index: no genomes loaded"); return MSSPE_ERR_STATE; }
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  msspe_join_reserve(c);             // msspe_reserve_pool: the pool's first touch ran beside the ingest
  msspe_free_index(c);
  const uint64_t G = c->n_segments;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&c->d_seg_part, (G ? G : 1) * sizeof(uint16_t), c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&c->d_seg_rec, (G ? G : 1) * sizeof(uint32_t), c->stream));
  c->timing.encode_ms = c->timing.index_ms = 0.f;
  if (msspe_build_fast_applicable(c)) {   // words of <= 16 bases: code + record index in one u64 (kmer_build_fast.cu)
    int rc = msspe_build_fast(c);
    if (rc) { msspe_free_index(c); return rc; }
    c->built = true;
    return MSSPE_OK;
  }
  for (int d = 0; d < 2; d++) {
    int rc = build_direction(c, d, &c->timing.encode_ms, &c->timing.index_ms);
    if (rc) { msspe_free_index(c); return rc; }
  }
  c->built = true;
  return MSSPE_OK;
}

extern "C" int msspe_get_segment_kmers(msspe_ctx* c, uint8_t dir, uint64_t* codes, uint64_t capacity) {
  if (!c || dir > 1) return MSSPE_ERR_INVALID;
  if (!c->loaded) { c->set_error("msspe_get_segment_kmers: no genomes loaded"); return MSSPE_ERR_STATE; }
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  const uint64_t need = c->n_segments * c->slots;
  if (capacity < need || (need && !codes)) { c->set_error("segment k-mer table needs %llu entries", (unsigned long long)need); return MSSPE_ERR_CAPACITY; }
  if (need == 0) return MSSPE_OK;
  uint64_t* dense = nullptr;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&dense, need * 8, c->stream));
  int rc = encode_dispatch(c, dir, ENC_DENSE, nullptr, nullptr, nullptr, nullptr, dense, c->stream);
  if (rc == MSSPE_OK) {
    cudaError_t e = cudaMemcpyAsync(codes, dense, need * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { c->set_error("copy of segment k-mers failed: %s", cudaGetErrorString(e)); rc = MSSPE_ERR_CUDA; }
  }
  msspe_dev_free(c, dense);
  return rc;
}

extern "C" int msspe_get_index(msspe_ctx* c, uint8_t dir, uint64_t* n_codes, uint64_t* n_postings, uint64_t* codes,
                               uint64_t* offsets, uint32_t* postings) {
  if (!c || dir > 1) return MSSPE_ERR_INVALID;
  if (!c->built) { c->set_error("msspe_get_index: index not built"); return MSSPE_ERR_STATE; }
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  DirIndex& D = c->dir[dir];
  if (n_codes) *n_codes = D.n_codes;
  if (n_postings) *n_postings = D.n_records;
  if (codes && D.n_codes) MSSPE_CUDA_TRY(c, cudaMemcpy(codes, D.codes, D.n_codes * 8, cudaMemcpyDeviceToHost));
  if (offsets) {
    std::vector<uint32_t> tmp(D.n_codes + 1);
    MSSPE_CUDA_TRY(c, cudaMemcpy(tmp.data(), D.post_off, (D.n_codes + 1) * 4, cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i <= D.n_codes; i++) offsets[i] = tmp[i];
  }
  if (postings && D.n_records) MSSPE_CUDA_TRY(c, cudaMemcpy(postings, D.postings, D.n_records * 4, cudaMemcpyDeviceToHost));
  return MSSPE_OK;
}
