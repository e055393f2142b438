// graph.cu -- SURVEY section 8 rows f-1 and f-2, device side.
//   msspe_vertex_cover      the conflict graph and the greedy vertex cover of od-msspe/src/main.rs:754-815
//                           (graphdb.rs is only its container) as an n x n adjacency bit matrix with popcount
//                           degrees and ONE persistent single-block kernel for the whole removal loop.
//   msspe_coverage_summary  the aggregation of print_coverage_report (main.rs:518-574) as device reductions over
//                           the resident segment arrays: per-record and per-partition (covered, total) counts.
#include <algorithm>
#include <numeric>

#include "engine.cuh"

namespace {

constexpr int VC_THREADS = 1024;

__global__ void vc_edges_kernel(const uint32_t* __restrict__ ea, const uint32_t* __restrict__ eb, uint64_t n_edges, uint32_t words,
                                uint32_t* adj) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const uint32_t a = ea[e], b = eb[e];
  // conflicts[a].insert(b); conflicts[b].insert(a)  (HashSet: duplicates collapse; a == b is a self conflict)
  atomicOr(&adj[(uint64_t)a * words + (b >> 5)], 1u << (b & 31u));
  atomicOr(&adj[(uint64_t)b * words + (a >> 5)], 1u << (a & 31u));
}

// deg[v] = |conflicts[v]| (one warp per row)
__global__ void vc_degree_kernel(const uint32_t* __restrict__ adj, uint32_t n, uint32_t words, uint32_t* __restrict__ deg) {
  const uint32_t v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= n) return;
  uint32_t s = 0;
  for (uint32_t w = threadIdx.x & 31u; w < words; w += 32) s += (uint32_t)__popc(adj[(uint64_t)v * words + w]);
  s = __reduce_add_sync(0xffffffffu, s);
  if ((threadIdx.x & 31u) == 0) deg[v] = s;
}

// The loop of main.rs:776-798 in one launch: among the primers that are not deleted and still have an active
// (not deleted) neighbour pick max (active count, word) -- rank[] is the position of the word in ascending order, so
// the greatest rank is the lexicographically greatest word -- delete it and take one off the active count of each
// of its neighbours.  deg[] counts a live self conflict too, exactly as `neighbors.iter().filter(..).count()` does.
__global__ void __launch_bounds__(VC_THREADS)
vc_greedy_kernel(const uint32_t* __restrict__ adj, uint32_t n, uint32_t words, const uint32_t* __restrict__ rank,
                 const uint32_t* __restrict__ order, uint32_t* deg,
                 uint32_t* alive, uint8_t* __restrict__ deleted, uint32_t* n_deleted) {
  __shared__ unsigned long long s_best[VC_THREADS / 32];
  __shared__ unsigned long long s_pick;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t removed = 0;
  for (;;) {
    unsigned long long best = 0ull;  // (active << 32 | rank + 1); 0 = nobody
    for (uint32_t v = tid; v < n; v += VC_THREADS) {
      const uint32_t d = deg[v];
      if (d && ((alive[v >> 5] >> (v & 31u)) & 1u)) best = max(best, ((unsigned long long)d << 32) | (unsigned long long)(rank[v] + 1u));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
    if (warp == 0) {
      unsigned long long b = lane < VC_THREADS / 32 ? s_best[lane] : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
      if (lane == 0) s_pick = b;
    }
    __syncthreads();
    const unsigned long long pick = s_pick;
    if (pick == 0ull) break;
    const uint32_t v = order[(uint32_t)pick - 1u];  // the node with that rank
    if (tid == 0) { deleted[v] = 1; alive[v >> 5] &= ~(1u << (v & 31u)); }
    __syncthreads();
    for (uint32_t w = tid; w < words; w += VC_THREADS) {
      uint32_t bits = adj[(uint64_t)v * words + w] & alive[w];
      while (bits) {
        const uint32_t u = w * 32u + (uint32_t)__ffs(bits) - 1u;
        bits &= bits - 1u;
        deg[u] -= 1u;  // each u belongs to exactly one thread of this step
      }
    }
    removed++;
    __syncthreads();
  }
  if (tid == 0) *n_deleted = removed;
}

// ---- coverage summary ----
__global__ void cov_mark_kernel(const uint64_t* __restrict__ sel, uint32_t n_sel, const uint64_t* __restrict__ codes, uint32_t n_codes,
                                const uint32_t* __restrict__ post_off, const uint32_t* __restrict__ postings, uint32_t* covered_bits) {
  const uint32_t s = blockIdx.x;
  if (s >= n_sel || n_codes == 0) return;
  const uint64_t want = sel[s];
  uint32_t lo = 0, hi = n_codes;  // first index with codes[idx] >= want
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (codes[mid] < want) lo = mid + 1; else hi = mid; }
  if (lo >= n_codes || codes[lo] != want) return;
  for (uint32_t i = post_off[lo] + threadIdx.x; i < post_off[lo + 1]; i += blockDim.x) {
    const uint32_t g = postings[i];
    atomicOr(&covered_bits[g >> 5], 1u << (g & 31u));
  }
}

// per-record / per-partition (covered, total): shared-memory histogram of the partitions, one atomic per segment
// for the records (segments of a record are consecutive, so a warp usually hits one or two records)
template <bool SMEM_HIST>
__global__ void __launch_bounds__(256)
cov_reduce_kernel(const uint32_t* __restrict__ covered_bits, const uint16_t* __restrict__ seg_part, const uint32_t* __restrict__ seg_rec,
                  uint64_t n_segments, uint32_t n_part, uint32_t* __restrict__ rec_cov, uint32_t* __restrict__ rec_tot,
                  uint32_t* __restrict__ part_cov, uint32_t* __restrict__ part_tot, unsigned long long* n_covered) {
  extern __shared__ uint32_t sh[];  // [n_part] covered, [n_part] total
  if (SMEM_HIST) {
    for (uint32_t p = threadIdx.x; p < 2 * n_part; p += blockDim.x) sh[p] = 0u;
    __syncthreads();
  }
  uint32_t mine = 0;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_segments; g += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t c = (covered_bits[g >> 5] >> (g & 31u)) & 1u;
    const uint32_t p = seg_part[g], r = seg_rec[g];
    atomicAdd(SMEM_HIST ? &sh[n_part + p] : &part_tot[p], 1u);
    atomicAdd(&rec_tot[r], 1u);
    if (c) { atomicAdd(SMEM_HIST ? &sh[p] : &part_cov[p], 1u); atomicAdd(&rec_cov[r], 1u); mine++; }
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_covered, (unsigned long long)mine);
  if (!SMEM_HIST) return;
  __syncthreads();
  for (uint32_t p = threadIdx.x; p < n_part; p += blockDim.x) {
    if (sh[p]) atomicAdd(&part_cov[p], sh[p]);
    if (sh[n_part + p]) atomicAdd(&part_tot[p], sh[n_part + p]);
  }
}

}  // namespace

extern "C" int msspe_vertex_cover(msspe_ctx* c, const uint64_t* codes, uint32_t n, const uint32_t* edge_a, const uint32_t* edge_b,
                                  uint64_t n_edges, uint8_t* deleted, uint32_t* n_deleted) {
  if (!c) return MSSPE_ERR_INVALID;
  if ((n && (!codes || !deleted)) || (n_edges && (!edge_a || !edge_b))) { c->set_error("msspe_vertex_cover: null argument"); return MSSPE_ERR_INVALID; }
  if (n_deleted) *n_deleted = 0;
  if (n == 0) return MSSPE_OK;
  if (n > 65536u) { c->set_error("msspe_vertex_cover: %u primers exceed the 65536-node adjacency matrix", n); return MSSPE_ERR_CAPACITY; }
  for (uint64_t e = 0; e < n_edges; e++)
    if (edge_a[e] >= n || edge_b[e] >= n) { c->set_error("msspe_vertex_cover: edge %llu names node %u/%u of %u", (unsigned long long)e, edge_a[e], edge_b[e], n); return MSSPE_ERR_INVALID; }
  // rank of every word in ascending order (equal-length words: code order = String order, main.rs:789)
  std::vector<uint32_t> order(n), rank(n);
  std::iota(order.begin(), order.end(), 0u);
  std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return codes[x] < codes[y]; });
  for (uint32_t i = 0; i + 1 < n; i++)
    if (codes[order[i]] == codes[order[i + 1]]) { c->set_error("msspe_vertex_cover: the primers must be distinct words (the reference keys its graph by word)"); return MSSPE_ERR_INVALID; }
  for (uint32_t i = 0; i < n; i++) rank[order[i]] = i;
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const uint32_t words = (n + 31u) / 32u;
  uint32_t *d_order = nullptr, *d_adj = nullptr, *d_rank = nullptr, *d_deg = nullptr, *d_alive = nullptr, *d_ea = nullptr, *d_eb = nullptr, *d_nd = nullptr;
  uint8_t* d_del = nullptr;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_adj, (uint64_t)n * words * 4, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_rank, (uint64_t)n * 4, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_deg, (uint64_t)n * 4, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_alive, (uint64_t)words * 4, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_del, n, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_nd, 4, st));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(d_adj, 0, (uint64_t)n * words * 4, st));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(d_del, 0, n, st));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(d_rank, rank.data(), (uint64_t)n * 4, cudaMemcpyHostToDevice, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_order, (uint64_t)n * 4, st));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(d_order, order.data(), (uint64_t)n * 4, cudaMemcpyHostToDevice, st));
  std::vector<uint32_t> alive(words, 0xFFFFFFFFu);
  if (n & 31u) alive[words - 1] = (1u << (n & 31u)) - 1u;
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(d_alive, alive.data(), (uint64_t)words * 4, cudaMemcpyHostToDevice, st));
  if (n_edges) {
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_ea, n_edges * 4, st));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_eb, n_edges * 4, st));
    MSSPE_CUDA_TRY(c, cudaMemcpyAsync(d_ea, edge_a, n_edges * 4, cudaMemcpyHostToDevice, st));
    MSSPE_CUDA_TRY(c, cudaMemcpyAsync(d_eb, edge_b, n_edges * 4, cudaMemcpyHostToDevice, st));
    vc_edges_kernel<<<(unsigned)div_up_u64(n_edges, 256), 256, 0, st>>>(d_ea, d_eb, n_edges, words, d_adj);
    c->timing.kernel_launches++;
  }
  vc_degree_kernel<<<(n + 7u) / 8u, 256, 0, st>>>(d_adj, n, words, d_deg);
  vc_greedy_kernel<<<1, VC_THREADS, 0, st>>>(d_adj, n, words, d_rank, d_order, d_deg, d_alive, d_del, d_nd);
  c->timing.kernel_launches += 2;
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  uint32_t nd = 0;
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(deleted, d_del, n, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&nd, d_nd, 4, cudaMemcpyDeviceToHost, st));
  for (void* p : {(void*)d_order, (void*)d_adj, (void*)d_rank, (void*)d_deg, (void*)d_alive, (void*)d_del, (void*)d_nd, (void*)d_ea, (void*)d_eb})
    if (p) MSSPE_CUDA_TRY(c, cudaFreeAsync(p, st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  if (n_deleted) *n_deleted = nd;
  return MSSPE_OK;
}

extern "C" int msspe_coverage_summary(msspe_ctx* c, const uint64_t* fwd_codes, uint32_t n_fwd, const uint64_t* rev_codes, uint32_t n_rev,
                                      uint32_t* rec_covered, uint32_t* rec_total, uint32_t n_records, uint32_t* part_covered,
                                      uint32_t* part_total, uint32_t n_part, uint64_t* n_covered) {
  if (!c) return MSSPE_ERR_INVALID;
  if (!c->built) { c->set_error("msspe_coverage_summary: index not built"); return MSSPE_ERR_STATE; }
  if ((n_fwd && !fwd_codes) || (n_rev && !rev_codes) || !rec_covered || !rec_total || !part_covered || !part_total) {
    c->set_error("msspe_coverage_summary: null argument"); return MSSPE_ERR_INVALID;
  }
  const uint64_t G = c->n_segments;
  if (n_records < c->n_records || (G && n_part <= c->max_partition)) {
    c->set_error("msspe_coverage_summary: need %u record and %u partition slots", c->n_records, c->max_partition + 1); return MSSPE_ERR_CAPACITY;
  }
  memset(rec_covered, 0, (size_t)n_records * 4); memset(rec_total, 0, (size_t)n_records * 4);
  memset(part_covered, 0, (size_t)n_part * 4); memset(part_total, 0, (size_t)n_part * 4);
  if (n_covered) *n_covered = 0;
  if (G == 0) return MSSPE_OK;
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const uint32_t np = c->max_partition + 1u, nr = c->n_records;
  const uint64_t bit_words = div_up_u64(G, 32);
  uint32_t* d_bits = nullptr; uint64_t* d_sel = nullptr; uint32_t* d_out = nullptr; unsigned long long* d_n = nullptr;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_bits, bit_words * 4, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_sel, (size_t)(n_fwd + n_rev + 1) * 8, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_out, ((uint64_t)2 * nr + 2 * np) * 4, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_n, 8, st));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(d_bits, 0, bit_words * 4, st));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(d_out, 0, ((uint64_t)2 * nr + 2 * np) * 4, st));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(d_n, 0, 8, st));
  if (n_fwd) MSSPE_CUDA_TRY(c, cudaMemcpyAsync(d_sel, fwd_codes, (size_t)n_fwd * 8, cudaMemcpyHostToDevice, st));
  if (n_rev) MSSPE_CUDA_TRY(c, cudaMemcpyAsync(d_sel + n_fwd, rev_codes, (size_t)n_rev * 8, cudaMemcpyHostToDevice, st));
  for (int d = 0; d < 2; d++) {
    const uint32_t ns = d == 0 ? n_fwd : n_rev;
    if (!ns) continue;
    DirIndex& D = c->dir[d];
    cov_mark_kernel<<<ns, 256, 0, st>>>(d_sel + (d == 0 ? 0 : n_fwd), ns, D.codes, (uint32_t)D.n_codes, D.post_off, D.postings, d_bits);
    c->timing.kernel_launches++;
  }
  uint32_t *d_rc = d_out, *d_rt = d_out + nr, *d_pc = d_out + 2 * (uint64_t)nr, *d_pt = d_pc + np;
  const size_t smem = (size_t)2 * np * 4;
  const unsigned grid = (unsigned)std::min<uint64_t>(div_up_u64(G, 256), (uint64_t)c->sm_count * 8);
  if (smem <= 32 * 1024) cov_reduce_kernel<true><<<grid, 256, smem, st>>>(d_bits, c->d_seg_part, c->d_seg_rec, G, np, d_rc, d_rt, d_pc, d_pt, d_n);
  else cov_reduce_kernel<false><<<grid, 256, 0, st>>>(d_bits, c->d_seg_part, c->d_seg_rec, G, np, d_rc, d_rt, d_pc, d_pt, d_n);
  c->timing.kernel_launches++;
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  unsigned long long ncov = 0;
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(rec_covered, d_rc, (size_t)nr * 4, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(rec_total, d_rt, (size_t)nr * 4, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(part_covered, d_pc, (size_t)np * 4, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(part_total, d_pt, (size_t)np * 4, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&ncov, d_n, 8, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaFreeAsync(d_bits, st));
  MSSPE_CUDA_TRY(c, cudaFreeAsync(d_sel, st));
  MSSPE_CUDA_TRY(c, cudaFreeAsync(d_out, st));
  MSSPE_CUDA_TRY(c, cudaFreeAsync(d_n, st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  if (n_covered) *n_covered = ncov;
  return MSSPE_OK;
}
