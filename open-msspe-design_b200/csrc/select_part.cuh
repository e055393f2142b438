// select_part.cuh -- device code of the per-partition greedy loop (see select_part.cu for the algorithm), shared by the
// single-GPU driver (select_part.cu) and the multi-GPU driver (select_dist.cu).  Internal: anonymous namespace per TU.
#pragma once
#include <cooperative_groups.h>

#include <algorithm>

#include "engine.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr uint32_t TK_LIVE = 0xFFFFFFFFu;                  // segment not covered
constexpr uint32_t LID_NONE = 0xFFFFFFFFu, LID_MULTI = 0x80000000u;
constexpr uint32_t ST_FINISHED = 1u, ST_ROLLBACK = 2u, ST_EXTEND = 4u;
constexpr uint32_t T_INF = 0xFFFFFFFFu;
constexpr int EXT_T = 512;        // threads of the unit kernel
constexpr int NEWCAP = 2048;      // newly covered segments handled per batch of the apply phase
constexpr int VER_T = 256;        // threads of the verify kernel
constexpr int TP_SLOTS = 64;      // distinct partitions of a multi-partition list handled by the parallel score path

struct PEntry { uint32_t freq, cid, tied, pad; unsigned long long live_before; unsigned long long code; };   // code = the word: rank-independent tie-break

struct PartCtl {
  uint32_t t_final, done, n_out, iterations;
  unsigned long long evals;
  unsigned long long live_all;     // live records of all units (evals of the iteration that finds nothing)
  uint32_t E, H, cutbound, V, terminal, do_terminal, fmin, t_hi;
  uint32_t n_stage, n_viol, vmin, n_ext;
  uint32_t rounds, rollbacks, wmax, last_viol;
  uint32_t clipped, tq, tq_next, pad2;
  unsigned long long work_bytes;   // algorithmic bytes of the unit kernel: 4 B per k-mer count scanned, 8 B per posting of a winner (id + cover
                                   // token), 4 B per forward-index record taken out of (or put back into) the live counts   // multi-GPU loop: iteration whose tie scores need the first-seen order of the partitions (pending / next)
};

struct PartDir {
  const uint64_t* codes; const uint32_t* post_off; const uint32_t* postings;
  const uint32_t* ucode_off; const uint32_t* ucodes; const uint32_t* fwdl; const uint32_t* useg_off; const uint32_t* usegs;
  uint32_t n_single, n_multi;
  uint32_t* pfreq; uint32_t* token; unsigned long long* ulive; PEntry* entries; uint32_t* pos; uint32_t* rfin; uint32_t* ulen;
  uint32_t* status; uint32_t* ext_cov; uint32_t* elist; uint32_t* order; uint32_t* tied; unsigned long long* tot_live; uint32_t* mt;
  uint32_t* ub; uint32_t* stage; uint4* viol; uint32_t* touch;
  uint32_t* win_freq; uint32_t* win_cov; unsigned long long* win_code;   // the merged winner of every window position (plan kernel)
  PartCtl* ctl; msspe_candidate* out;
};

struct PartArgs {
  PartDir d[2];
  int ndirs;
  uint32_t U, CAP, slots, max_iter, mms, uniform_parts, nsteps;
  uint32_t max_ahead;   // a unit keeps at most this many not-yet-final entries (the multi-GPU exchange has fixed-size records)
  const uint16_t* seg_part;
};

__device__ __forceinline__ uint32_t part_of(const PartArgs& A, uint32_t seg) {
  return A.uniform_parts ? seg % A.uniform_parts : (uint32_t)A.seg_part[seg];
}

template <int T>
__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < T / 32 ? sh[lane] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sh[32] = v;
  }
  __syncthreads();
  return sh[32];
}

template <int T>
__device__ __forceinline__ uint32_t block_min_u32(uint32_t v, unsigned long long* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_down_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < T / 32 ? (uint32_t)sh[lane] : 0xFFFFFFFFu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_down_sync(0xffffffffu, v, o));
    if (lane == 0) sh[32] = v;
  }
  __syncthreads();
  return (uint32_t)sh[32];
}

// ---- unit kernel ---------------------------------------------------------------------------------------------------
// One thread-block CLUSTER of C CTAs per unit (C = 1, 2, 4 or 8 by unit size): the CTAs split the unit's k-mers for the
// arg-max, the winner's postings for the cover step and the unit's segments for a roll-back, and exchange their partial
// results through distributed shared memory (every CTA pushes its part into the others' buffers, one hardware cluster
// barrier, every CTA combines the same C parts) -- two cluster barriers per greedy step, no global-memory round trip.
struct ExPart { unsigned long long key; long long delta; uint32_t cnt; uint32_t pad; };

template <int C>
__device__ __forceinline__ void cluster_sync_all() {
  if (C == 1) __syncthreads();
  else cg::this_cluster().sync();   // barrier.cluster arrive.release / wait.acquire: the CTAs' global atomics are ordered by it
}

// s_ex[par][r] of every CTA <- this CTA's part; after the barrier every CTA holds all C parts
template <int C>
__device__ __forceinline__ void exchange(ExPart (*s_ex)[8], int par, uint32_t rank, unsigned long long key, uint32_t cnt, long long delta) {
  if (threadIdx.x == 0) {
    ExPart e; e.key = key; e.delta = delta; e.cnt = cnt; e.pad = 0u;
    if (C == 1) s_ex[par][0] = e;
    else {
      cg::cluster_group cl = cg::this_cluster();
#pragma unroll
      for (int rr = 0; rr < C; rr++) { ExPart* dst = cl.map_shared_rank(&s_ex[par][rank], rr); *dst = e; }
    }
  }
  cluster_sync_all<C>();
}

// +1 / -1 on the live counts of the k-mers of `n` segments held in s_list (bit 31 of an entry = "revive": +1).
// Flattened over (segment, slot): consecutive lanes read consecutive forward-index entries (coalesced); four independent
// loads are in flight per thread before the first atomic -- the loop is bound by L2 latency, not by bandwidth.
__device__ __forceinline__ long long apply_list(const PartDir& D, const uint32_t* s_list, uint32_t n, uint32_t slots) {
  long long d = 0;
  const uint32_t items = n * slots;
  for (uint32_t x0 = threadIdx.x; x0 < items; x0 += 4 * EXT_T) {
    uint32_t l[4], sg[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t x = x0 + k * EXT_T;
      l[k] = LID_NONE; sg[k] = 0u;
      if (x < items) {
        const uint32_t si = x / slots, q = x - si * slots;
        const uint32_t e = s_list[si];
        sg[k] = e & 0x80000000u;
        l[k] = __ldg(D.fwdl + (unsigned long long)(e & 0x7FFFFFFFu) * slots + q);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (l[k] == LID_NONE) continue;
      if (sg[k]) { d--; if (!(l[k] & LID_MULTI)) atomicAdd(D.pfreq + l[k], 1u); }
      else { d++; if (!(l[k] & LID_MULTI)) atomicSub(D.pfreq + l[k], 1u); }
    }
  }
  return d;
}

template <int C>
__global__ void __launch_bounds__(EXT_T) part_extend_kernel(PartArgs A) {
  const PartDir& D = A.d[blockIdx.y];
  if (D.ctl->done) return;
  const uint32_t u = blockIdx.x / C, rank = blockIdx.x % C;   // cluster = C consecutive blocks of the x dimension
  const uint32_t st = D.status[u];
  if (!(st & (ST_ROLLBACK | ST_EXTEND))) return;              // the whole cluster leaves together
  __shared__ unsigned long long sh[34];
  __shared__ unsigned long long s_key[EXT_T / 32];
  __shared__ uint32_t s_cnt[EXT_T / 32];
  __shared__ uint32_t s_new[NEWCAP];
  __shared__ uint32_t s_n;
  __shared__ ExPart s_ex[2][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t o0 = D.ucode_off[u], o1 = D.ucode_off[u + 1];
  const uint32_t g0 = D.useg_off[u], g1 = D.useg_off[u + 1];
  const uint32_t slot0 = u * A.CAP, slots = A.slots;
  uint32_t len = D.ulen[u];
  long long live = (long long)D.ulive[u];
  bool finished = (st & ST_FINISHED) != 0;
  int par = 0;
  long long delta = 0;                                        // records this CTA took out of the live count, not yet exchanged
  unsigned long long work = 0;                                // algorithmic bytes (counted by CTA 0 of the cluster)
  if (C > 1) cg::this_cluster().sync();                       // nobody pushes into a CTA that has not started yet
  if (st & ST_ROLLBACK) {
    // back to the unit's state at the external winner's iteration: segments covered by the truncated entries live again,
    // segments the external winner newly covered leave the counts.  Cost ~ changed segments, not the unit.
    const uint32_t keep = D.rfin[u];
    const uint32_t ext_live = A.U * A.CAP + 2u * (D.ctl->n_ext - 1u);   // token of "was live, now covered by the newest external winner"
    len = keep; finished = false;
    const uint32_t gshare = (g1 - g0 + C - 1) / C;
    const uint32_t gb = min(g1, g0 + rank * gshare), ge = min(g1, gb + gshare);
    for (uint32_t b = gb; b < ge; b += NEWCAP) {
      if (tid == 0) s_n = 0u;
      __syncthreads();
      const uint32_t be = min(ge, b + (uint32_t)NEWCAP);
      for (uint32_t i = b + tid; i < be; i += EXT_T) {
        const uint32_t g = D.usegs[i];
        const uint32_t tk = __ldcg(D.token + g);
        if (tk >= slot0 + keep && tk < slot0 + A.CAP) { D.token[g] = TK_LIVE; s_new[atomicAdd(&s_n, 1u)] = g | 0x80000000u; }
        else if (tk == ext_live) s_new[atomicAdd(&s_n, 1u)] = g;
      }
      __syncthreads();
      delta += apply_list(D, s_new, s_n, slots);
      __syncthreads();
    }
    delta = (long long)block_sum_u64<EXT_T>((unsigned long long)delta, sh);
    exchange<C>(s_ex, par, rank, 0ull, 0u, delta);
    for (int rr = 0; rr < C; rr++) { const long long dd = s_ex[par][rr].delta; live -= dd; work += 4ull * (unsigned long long)(dd < 0 ? -dd : dd); }
    work += 4ull * (g1 - g0);
    par ^= 1; delta = 0;
  }
  const uint32_t rf0 = D.rfin[u];
  for (uint32_t step = 0; step < A.nsteps && !finished && len - rf0 < A.max_ahead; step++) {
    if (len >= A.CAP) { finished = true; break; }  // entry number CAP = max_iterations can never be among the first max_iterations
    // arg-max over the unit's k-mers: (live count, then smaller word = smaller index), and how many share the count
    unsigned long long bk = 0ull; uint32_t bc = 0u;
    for (uint32_t j0 = o0 + rank * EXT_T + tid; j0 < o1; j0 += 4 * C * EXT_T) {   // four loads in flight per thread
      uint32_t fv[4];
#pragma unroll
      for (int k = 0; k < 4; k++) { const uint32_t j = j0 + k * C * EXT_T; fv[k] = j < o1 ? __ldcg(D.pfreq + j) : 0xFFFFFFFFu; }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t f = fv[k], j = j0 + k * C * EXT_T;
        if (j >= o1) continue;
        const unsigned long long key = ((unsigned long long)f << 32) | (unsigned long long)(0xFFFFFFFFu - (j - o0));
        const uint32_t bf = (uint32_t)(bk >> 32);
        if (f > bf || bc == 0u) { bk = key; bc = 1u; }
        else if (f == bf) { bc++; if (key > bk) bk = key; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ok = __shfl_down_sync(0xffffffffu, bk, o);
      const uint32_t oc = __shfl_down_sync(0xffffffffu, bc, o);
      if (oc) {
        const uint32_t f = (uint32_t)(ok >> 32), bf = (uint32_t)(bk >> 32);
        if (bc == 0u || f > bf) { bk = ok; bc = oc; }
        else if (f == bf) { bc += oc; if (ok > bk) bk = ok; }
      }
    }
    if (lane == 0) { s_key[warp] = bk; s_cnt[warp] = bc; }
    __syncthreads();
    if (tid == 0) {
      bk = 0ull; bc = 0u;
      for (int w2 = 0; w2 < EXT_T / 32; w2++) {
        const unsigned long long ok = s_key[w2]; const uint32_t oc = s_cnt[w2];
        if (!oc) continue;
        const uint32_t f = (uint32_t)(ok >> 32), bf = (uint32_t)(bk >> 32);
        if (bc == 0u || f > bf) { bk = ok; bc = oc; }
        else if (f == bf) { bc += oc; if (ok > bk) bk = ok; }
      }
    }
    exchange<C>(s_ex, par, rank, bk, bc, delta);      // + what the previous step took out of the live records
    unsigned long long wk = 0ull; uint32_t wc = 0u;
    work += 4ull * (o1 - o0);
    for (int rr = 0; rr < C; rr++) {
      const ExPart e = s_ex[par][rr];
      live -= e.delta; work += 4ull * (unsigned long long)e.delta;
      if (!e.cnt) continue;
      const uint32_t f = (uint32_t)(e.key >> 32), bf = (uint32_t)(wk >> 32);
      if (wc == 0u || f > bf) { wk = e.key; wc = e.cnt; }
      else if (f == bf) { wc += e.cnt; if (e.key > wk) wk = e.key; }
    }
    par ^= 1; delta = 0;
    const uint32_t fmax = wc ? (uint32_t)(wk >> 32) : 0u;
    if (fmax < 2u) { finished = true; break; }      // freq == 1 stops before the push (main.rs:354-360); 0 = None
    const uint32_t jwin = o0 + (0xFFFFFFFFu - (uint32_t)wk);
    const uint32_t cid = D.ucodes[jwin];
    const uint32_t slot = slot0 + len;
    if (rank == 0 && tid == 0) { PEntry e; e.freq = fmax; e.cid = cid; e.tied = wc; e.pad = 0u; e.live_before = (unsigned long long)live; e.code = D.codes[cid]; D.entries[slot] = e; }
    const uint32_t pb = D.post_off[cid], pe = D.post_off[cid + 1];
    work += 8ull * (pe - pb);
    const uint32_t share = (pe - pb + C - 1) / C;      // an equal share of the winner's postings for every CTA of the cluster
    const uint32_t mb = min(pe, pb + rank * share), me = min(pe, mb + share);
    for (uint32_t b = mb; b < me; b += NEWCAP) {       // main.rs:371-378 for this unit: cover the winner's live segments ...
      if (tid == 0) s_n = 0u;
      __syncthreads();
      const uint32_t be = min(me, b + (uint32_t)NEWCAP);
      for (uint32_t i = b + tid; i < be; i += EXT_T) {
        const uint32_t g = __ldg(D.postings + i);
        if (__ldcg(D.token + g) == TK_LIVE) { D.token[g] = slot; s_new[atomicAdd(&s_n, 1u)] = g; }
      }
      __syncthreads();
      delta += apply_list(D, s_new, s_n, slots);      // ... and take their k-mers out of the live counts
      __syncthreads();
    }
    delta = (long long)block_sum_u64<EXT_T>((unsigned long long)delta, sh);
    cluster_sync_all<C>();                           // every CTA's decrements are in before anybody scans again
    len++;
    if (fmax < A.mms) { finished = true; break; }   // pushed, then break (main.rs:387-390)
  }
  exchange<C>(s_ex, par, rank, 0ull, 0u, delta);      // the last step's share of the live records
  for (int rr = 0; rr < C; rr++) { live -= s_ex[par][rr].delta; work += 4ull * (unsigned long long)s_ex[par][rr].delta; }
  if (rank == 0 && tid == 0) { D.ulen[u] = len; D.ulive[u] = (unsigned long long)live; D.status[u] = finished ? ST_FINISHED : 0u; atomicAdd(&D.ctl->work_bytes, work); }
  if (C > 1) cg::this_cluster().sync();               // no CTA exits while a peer may still push into its shared memory
}

// initial state: live count of a single-partition list = its length; live records per unit
__global__ void part_init_freq_kernel(PartArgs A) {
  const PartDir& D = A.d[blockIdx.y];
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= D.n_single) return;
  const uint32_t c = D.ucodes[j];
  D.pfreq[j] = D.post_off[c + 1] - D.post_off[c];
}
__global__ void part_init_live_kernel(PartArgs A, unsigned long long G) {
  const PartDir& D = A.d[blockIdx.y];
  const unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  uint32_t n = 0;
  const uint32_t* f = D.fwdl + g * A.slots;
  for (uint32_t q = 0; q < A.slots; q++) n += __ldg(f + q) != LID_NONE;
  if (n) atomicAdd(D.ulive + part_of(A, (uint32_t)g), (unsigned long long)n);
}

// ---- list of non-final entries ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) part_gather_kernel(PartArgs A) {
  const PartDir& D = A.d[blockIdx.x];
  PartCtl* C = D.ctl;
  if (C->done) return;
  __shared__ unsigned long long sh[34];
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0u;
  unsigned long long live = 0;
  __syncthreads();
  for (uint32_t u0 = 0; u0 < A.U; u0 += 1024) {
    const uint32_t u = u0 + tid;
    uint32_t n = 0, rf = 0;
    if (u < A.U) { rf = D.rfin[u]; n = D.ulen[u] - rf; live += D.ulive[u]; }
    uint32_t inc = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_w[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
      s_w[lane] = wi - w;
      if (lane == 31) sh[33] = wi;
    }
    __syncthreads();
    const uint32_t off = s_base + s_w[warp] + inc - n;
    for (uint32_t r = 0; r < n; r++) D.elist[off + r] = u * A.CAP + rf + r;
    __syncthreads();
    if (tid == 0) s_base += (uint32_t)sh[33];
    __syncthreads();
  }
  const unsigned long long la = block_sum_u64<1024>(live, sh);
  for (uint32_t i = tid; i <= A.max_iter; i += 1024) D.mt[i] = 0u;
  if (tid == 0) { C->E = s_base; C->live_all = la; C->n_stage = 0u; C->n_viol = 0u; C->vmin = T_INF; C->tq_next = T_INF; C->rounds++; }
}

// key order of the merge: higher frequency, then lower partition_coverage (= higher 1/(cov+1), main.rs:320-324), then smaller word
__device__ __forceinline__ bool better(uint32_t fa, uint32_t ca, unsigned long long ia, uint32_t fb, uint32_t cb, unsigned long long ib) {
  return fa > fb || (fa == fb && (ca < cb || (ca == cb && ia < ib)));
}

__global__ void __launch_bounds__(256) part_merge_kernel(PartArgs A) {
  const PartDir& D = A.d[blockIdx.y];
  const PartCtl* C = D.ctl;
  if (C->done) return;
  const uint32_t E = C->E, t_final = C->t_final;
  const int lane = threadIdx.x & 31;
  const uint32_t wpb = blockDim.x >> 5;
  for (uint32_t xi = blockIdx.x * wpb + (threadIdx.x >> 5); xi < E; xi += gridDim.x * wpb) {
    const uint32_t slot = D.elist[xi];
    const uint32_t ux = slot / A.CAP, rx = slot - ux * A.CAP;
    const PEntry ex = D.entries[slot];
    const uint32_t covx = rx + D.ext_cov[ux];
    uint32_t cnt = 0, tied = 0; unsigned long long live = 0;
    for (uint32_t q = lane; q < A.U; q += 32) {
      const uint32_t rf = D.rfin[q], ln = D.ulen[q];
      uint32_t idx;                                  // the entry of unit q that is its head at the time x is chosen
      if (q == ux) idx = rx;
      else {
        const uint32_t ec = D.ext_cov[q];
        uint32_t lo = rf, n = ln - rf;
        while (n > 0) {
          const uint32_t half = n >> 1, mid = lo + half;
          const PEntry* e = D.entries + (unsigned long long)q * A.CAP + mid;
          if (better(e->freq, mid + ec, e->code, ex.freq, covx, ex.code)) { lo = mid + 1; n -= half + 1; } else n = half;
        }
        idx = lo;
      }
      cnt += idx - rf;
      if (idx < ln) {
        const PEntry* e = D.entries + (unsigned long long)q * A.CAP + idx;
        live += e->live_before;
        if (e->freq == ex.freq) tied += e->tied;
      } else live += D.ulive[q];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cnt += __shfl_down_sync(0xffffffffu, cnt, o); tied += __shfl_down_sync(0xffffffffu, tied, o); live += __shfl_down_sync(0xffffffffu, live, o);
    }
    if (lane == 0) {
      D.pos[slot] = t_final + cnt;
      if (t_final + cnt <= A.max_iter) { D.order[cnt] = slot; D.tied[cnt] = tied; D.tot_live[cnt] = live; }
    }
  }
}

__global__ void __launch_bounds__(1024) part_plan_kernel(PartArgs A) {
  const PartDir& D = A.d[blockIdx.x];
  PartCtl* C = D.ctl;
  if (C->done) return;
  __shared__ unsigned long long sh[34];
  const int tid = threadIdx.x;
  const uint32_t E = C->E, t_final = C->t_final;
  uint32_t h = T_INF;
  for (uint32_t u = tid; u < A.U; u += 1024) {
    if (!(D.status[u] & ST_FINISHED)) {
      const uint32_t rf = D.rfin[u], ln = D.ulen[u];
      h = min(h, ln > rf ? D.pos[u * A.CAP + ln - 1] + 1u : t_final);
    }
  }
  const uint32_t H = block_min_u32<1024>(h, sh);
  const uint32_t room = A.max_iter - t_final;           // positions still open
  uint32_t first = T_INF;
  for (uint32_t i = tid; i < E && i < room; i += 1024)
    if (D.entries[D.order[i]].freq < A.mms) first = min(first, i);
  first = block_min_u32<1024>(first, sh);
  __shared__ uint32_t s_V;
  if (tid == 0) {
    uint32_t cutbound = A.max_iter, terminal = 0u;
    if (first != T_INF) cutbound = min(cutbound, t_final + first + 1u);
    else if (E < room) { cutbound = t_final + E; terminal = 1u; }
    // verify at most wmax iterations ahead: the earliest external winner is all that counts, and it is usually near
    uint32_t V = min(H, cutbound);
    const uint32_t wmax = C->wmax ? C->wmax : A.max_iter;
    bool clipped = false;
    if (V - t_final > wmax) { V = t_final + wmax; clipped = true; }
    const uint32_t do_term = (terminal && H == T_INF && !clipped) ? 1u : 0u;
    uint32_t fmin = T_INF;
    if (do_term) fmin = 2u;
    else if (V > t_final) fmin = D.entries[D.order[V - 1u - t_final]].freq;
    C->clipped = clipped ? 1u : 0u;
    s_V = V;
    C->H = H; C->cutbound = cutbound; C->V = V; C->terminal = terminal; C->do_terminal = do_term; C->fmin = fmin; C->t_hi = V + do_term;
  }
  __syncthreads();
  for (uint32_t i = tid; i < s_V - t_final; i += 1024) {   // what the verify kernels compare the multi-partition lists with
    const uint32_t slot = D.order[i];
    const uint32_t ux = slot / A.CAP, rx = slot - ux * A.CAP;
    const PEntry* e = D.entries + slot;
    D.win_freq[i] = e->freq; D.win_cov[i] = rx + D.ext_cov[ux]; D.win_code[i] = e->code;
  }
}

__global__ void __launch_bounds__(256) part_stage_kernel(PartArgs A) {
  const PartDir& D = A.d[blockIdx.y];
  PartCtl* C = D.ctl;
  if (C->done) return;
  const uint32_t fmin = C->fmin;
  const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < D.n_multi && D.ub[m] >= fmin) D.stage[atomicAdd(&C->n_stage, 1u)] = m;
}

// cover time of a segment: the global position of the entry (or external winner) that covered it
__device__ __forceinline__ uint32_t time_of(const PartDir& D, uint32_t g) {
  const uint32_t tk = __ldcg(D.token + g);
  return tk == TK_LIVE ? T_INF : __ldcg(D.pos + tk);
}

// partition_coverage[p] at iteration t: winners unit p supplied before t + external winners that touched it (all final)
__device__ uint32_t cov_at(const PartArgs& A, const PartDir& D, uint32_t p, uint32_t t) {
  uint32_t lo = 0, n = D.ulen[p];
  const uint32_t* ps = D.pos + (unsigned long long)p * A.CAP;
  while (n > 0) { const uint32_t half = n >> 1; if (__ldcg(ps + lo + half) < t) { lo += half + 1; n -= half + 1; } else n = half; }
  return D.ext_cov[p] + lo;
}

// partition_tie_score (main.rs:261-283) of a multi-partition list at iteration t: f32 terms added sequentially in the
// order in which the partitions are first seen among the LIVE postings (ascending segment index).
__device__ float multi_score(const PartArgs& A, const PartDir& D, uint32_t a, uint32_t b, uint32_t t, uint32_t* tp, uint32_t* tfirst,
                             uint32_t* s_flag, uint32_t* seen, float* s_score) {
  const int tid = threadIdx.x;
  for (int k = tid; k < TP_SLOTS; k += VER_T) { tp[k] = 0xFFFFFFFFu; tfirst[k] = 0xFFFFFFFFu; }
  if (tid == 0) *s_flag = 0u;
  __syncthreads();
  for (uint32_t i = a + tid; i < b; i += VER_T) {
    const uint32_t g = __ldg(D.postings + i);
    if (time_of(D, g) < t) continue;
    const uint32_t p = part_of(A, g);
    int k = 0;
    for (; k < TP_SLOTS; k++) {
      const uint32_t old = atomicCAS(&tp[k], 0xFFFFFFFFu, p);
      if (old == 0xFFFFFFFFu || old == p) { atomicMin(&tfirst[k], i); break; }
    }
    if (k == TP_SLOTS) *s_flag = 1u;
  }
  __syncthreads();
  if (tid == 0) {
    float score = 0.0f;
    if (*s_flag) {   // more distinct partitions than slots (degenerate inputs): the reference's loop as it stands
      const uint32_t words = (A.U + 31u) / 32u;
      for (uint32_t w2 = 0; w2 < words; w2++) seen[w2] = 0u;
      for (uint32_t i = a; i < b; i++) {
        const uint32_t g = __ldg(D.postings + i);
        if (time_of(D, g) < t) continue;
        const uint32_t p = part_of(A, g);
        if (!((seen[p >> 5] >> (p & 31u)) & 1u)) {
          seen[p >> 5] |= 1u << (p & 31u);
          score = __fadd_rn(score, __fdiv_rn(1.0f, __fadd_rn((float)cov_at(A, D, p, t), 1.0f)));
        }
      }
    } else {
      int n = 0;
      for (int k = 0; k < TP_SLOTS; k++) if (tp[k] != 0xFFFFFFFFu) { tp[n] = tp[k]; tfirst[n] = tfirst[k]; n++; }
      for (int i = 1; i < n; i++) {
        const uint32_t kp = tp[i], kf = tfirst[i]; int j = i - 1;
        while (j >= 0 && tfirst[j] > kf) { tp[j + 1] = tp[j]; tfirst[j + 1] = tfirst[j]; j--; }
        tp[j + 1] = kp; tfirst[j + 1] = kf;
      }
      for (int i = 0; i < n; i++) score = __fadd_rn(score, __fdiv_rn(1.0f, __fadd_rn((float)cov_at(A, D, tp[i], t), 1.0f)));
    }
    *s_score = score;
  }
  __syncthreads();
  return *s_score;
}

// dynamic shared memory: h[max_iter + 2] (cover-time histogram -> exclusive prefix), wf[max_iter + 2], seen[(U + 31) / 32]
__global__ void __launch_bounds__(VER_T) part_verify_kernel(PartArgs A) {
  const PartDir& D = A.d[blockIdx.y];
  PartCtl* C = D.ctl;
  if (C->done) return;
  extern __shared__ uint32_t dsm[];
  uint32_t* h = dsm;
  uint32_t* wf = dsm + (A.max_iter + 2u);             // winning frequency of every iteration of the window
  uint32_t* seen = wf + (A.max_iter + 2u);
  __shared__ unsigned long long sh[34];
  __shared__ uint32_t tp[TP_SLOTS], tfirst[TP_SLOTS];
  __shared__ uint32_t s_flag, s_carry;
  __shared__ float s_score;
  const int tid = threadIdx.x;
  const uint32_t n_stage = C->n_stage, t_final = C->t_final, V = C->V, t_hi = C->t_hi, fmin = C->fmin;
  const uint32_t W = t_hi - t_final;
  if (blockIdx.x >= n_stage) return;
  for (uint32_t i = tid; i < V - t_final; i += VER_T) wf[i] = D.win_freq[i];
  __syncthreads();
  for (uint32_t si = blockIdx.x; si < n_stage; si += gridDim.x) {
    const uint32_t m = D.stage[si];
    const uint32_t c = D.ucodes[D.n_single + m];
    const uint32_t a = D.post_off[c], b = D.post_off[c + 1];
    for (uint32_t i = tid; i < W; i += VER_T) h[i] = 0u;
    if (tid == 0) s_carry = 0u;
    __syncthreads();
    unsigned long long l0 = 0;
    for (uint32_t i = a + tid; i < b; i += VER_T) {
      const uint32_t tm = time_of(D, __ldg(D.postings + i));
      if (tm >= t_final) { l0++; if (tm < t_hi) atomicAdd(&h[tm - t_final], 1u); }
    }
    const uint32_t L0 = (uint32_t)block_sum_u64<VER_T>(l0, sh);
    if (tid == 0) D.ub[m] = L0;                      // live count at t_final: an upper bound for every later iteration
    if (L0 < fmin) continue;
    // h[i] -> postings covered before iteration t_final + i (exclusive prefix); live count there = L0 - h[i]
    for (uint32_t i0 = 0; i0 < W; i0 += VER_T) {
      const uint32_t i = i0 + tid;
      const uint32_t v = i < W ? h[i] : 0u;
      uint32_t inc = v;
      const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
      __syncthreads();
      if (lane == 31) sh[warp] = inc;
      __syncthreads();
      uint32_t wbase = 0;
      for (int w2 = 0; w2 < warp; w2++) wbase += (uint32_t)sh[w2];
      const uint32_t carry = s_carry;
      if (i < W) h[i] = carry + wbase + inc - v;
      __syncthreads();
      if (tid == VER_T - 1) s_carry = carry + wbase + inc;
      __syncthreads();
    }
    // first iteration with a strictly larger live count than the merged winner (terminal iteration: any count >= 2)
    uint32_t fs = T_INF;
    for (uint32_t i = tid; i < W; i += VER_T) {
      const uint32_t t = t_final + i, cnt = L0 - h[i];
      const bool strict = t < V ? cnt > wf[i] : cnt >= 2u;
      if (strict) { fs = i; break; }
    }
    fs = block_min_u32<VER_T>(fs, sh);
    // equal counts before that: the tie-break decides, in ascending iteration order, until one wins
    const uint32_t lim = min(fs, V - t_final);
    uint32_t tv = T_INF, tv_cnt = 0; float tv_score = 0.0f;
    for (uint32_t cur = 0; cur < lim;) {
      uint32_t nx = T_INF;
      for (uint32_t i = cur + tid; i < lim; i += VER_T) if (L0 - h[i] == wf[i]) { nx = i; break; }
      nx = block_min_u32<VER_T>(nx, sh);
      if (nx == T_INF) break;
      const uint32_t i = nx, t = t_final + nx;
      if (tid == 0) atomicAdd(D.mt + i, 1u);
      const float sc = multi_score(A, D, a, b, t, tp, tfirst, &s_flag, seen, &s_score);
      const float wsc = __fdiv_rn(1.0f, __fadd_rn((float)D.win_cov[i], 1.0f));
      if (sc > wsc || (sc == wsc && D.codes[c] < D.win_code[i])) { tv = t; tv_cnt = wf[i]; tv_score = sc; break; }
      cur = nx + 1u;
    }
    if (tv == T_INF && fs != T_INF) {
      tv = t_final + fs; tv_cnt = L0 - h[fs];
      tv_score = multi_score(A, D, a, b, tv, tp, tfirst, &s_flag, seen, &s_score);
    }
    if (tv != T_INF && tid == 0) {
      const uint32_t k = atomicAdd(&C->n_viol, 1u);
      D.viol[k] = make_uint4(tv, tv_cnt, __float_as_uint(tv_score), c);
      atomicMin(&C->vmin, tv);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) part_finalize_kernel(PartArgs A) {
  const PartDir& D = A.d[blockIdx.x];
  PartCtl* C = D.ctl;
  if (C->done) return;
  __shared__ unsigned long long sh[34];
  __shared__ unsigned long long s_best;
  __shared__ uint32_t s_bestc, s_same;
  const int tid = threadIdx.x;
  const uint32_t t_final = C->t_final, V = C->V, vmin = C->vmin, n_viol = C->n_viol;
  const bool viol = vmin != T_INF;
  const uint32_t t_new = viol ? vmin : V;
  // winners before t_new are final
  unsigned long long ev = 0;
  for (uint32_t i = tid; i < t_new - t_final; i += 1024) {
    const uint32_t slot = D.order[i];
    const PEntry e = D.entries[slot];
    const uint32_t ux = slot / A.CAP, rx = slot - ux * A.CAP;
    msspe_candidate o;
    o.code = e.code; o.freq = e.freq; o.n_tied = D.tied[i] + D.mt[i];
    o.tie_score = __fdiv_rn(1.0f, __fadd_rn((float)(rx + D.ext_cov[ux]), 1.0f)); o.reserved = 0u;
    D.out[t_final + i] = o;
    ev += D.tot_live[i];
  }
  ev = block_sum_u64<1024>(ev, sh);
  for (uint32_t u = tid; u < A.U; u += 1024) {
    uint32_t lo = D.rfin[u], n = D.ulen[u] - lo;
    const uint32_t* ps = D.pos + (unsigned long long)u * A.CAP;
    while (n > 0) { const uint32_t half = n >> 1; if (ps[lo + half] < t_new) { lo += half + 1; n -= half + 1; } else n = half; }
    D.rfin[u] = lo;
  }
  if (tid == 0) { s_best = 0ull; s_bestc = 0u; s_same = 0u; }
  __syncthreads();
  if (viol) {
    // the external winner: best (count, score, smaller word) among the lists that win at t_new
    for (uint32_t k = tid; k < n_viol; k += 1024) {
      const uint4 v = D.viol[k];
      if (v.x == t_new) atomicMax(&s_best, ((unsigned long long)v.y << 32) | (unsigned long long)v.z);  // scores are >= 0: bit order = value order
    }
    __syncthreads();
    for (uint32_t k = tid; k < n_viol; k += 1024) {
      const uint4 v = D.viol[k];
      if (v.x == t_new && ((((unsigned long long)v.y << 32) | (unsigned long long)v.z) == s_best)) atomicMax(&s_bestc, 0xFFFFFFFFu - v.w);
      if (v.x == t_new && v.y == (uint32_t)(s_best >> 32)) atomicAdd(&s_same, 1u);
    }
    __syncthreads();
    const uint32_t c = 0xFFFFFFFFu - s_bestc, cnt = (uint32_t)(s_best >> 32);
    const uint32_t iw = t_new - t_final;
    const bool has_entry = t_new < V;                // a merged winner stood at t_new (not the terminal iteration)
    const uint32_t j = C->n_ext;
    const uint32_t a = D.post_off[c], b = D.post_off[c + 1];
    if (tid == 0) {
      msspe_candidate o;
      o.code = D.codes[c]; o.freq = cnt; o.tie_score = __uint_as_float((uint32_t)s_best); o.reserved = 0u;
      const bool tie_case = has_entry && D.entries[D.order[iw]].freq == cnt;
      o.n_tied = tie_case ? D.tied[iw] + D.mt[iw] : s_same;
      D.out[t_new] = o;
      D.pos[A.U * A.CAP + 2u * j] = t_new; D.pos[A.U * A.CAP + 2u * j + 1u] = t_new;
      C->evals += ev + (has_entry ? D.tot_live[iw] : C->live_all);
      C->iterations += iw + 1u;
    }
    __syncthreads();
    for (uint32_t i = a + tid; i < b; i += 1024) {   // main.rs:371-378: cover its live postings, partition_coverage for ALL its partitions
      const uint32_t g = D.postings[i];
      const uint32_t p = part_of(A, g);
      if (atomicExch(D.touch + p, 1u) == 0u) atomicAdd(D.ext_cov + p, 1u);
      const uint32_t tk = __ldcg(D.token + g);    // two codes: "was live" (leaves the counts at the roll-back) / "was covered by a truncated entry"
      if (tk == TK_LIVE || __ldcg(D.pos + tk) >= t_new) { D.token[g] = A.U * A.CAP + 2u * j + (tk != TK_LIVE ? 1u : 0u); D.status[p] = ST_ROLLBACK | ST_EXTEND; }
    }
    __syncthreads();
    for (uint32_t i = a + tid; i < b; i += 1024) D.touch[part_of(A, D.postings[i])] = 0u;
    if (tid == 0) {
      C->n_ext = j + 1u; C->rollbacks++;
      const uint32_t gap = t_new - C->last_viol;     // window of the next rounds: twice the distance to the previous external winner
      C->last_viol = t_new; C->wmax = min(A.max_iter, max(32u, 2u * gap));
      C->t_final = t_new + 1u;
      if (cnt < A.mms || t_new + 1u >= A.max_iter) { C->done = 1u; C->n_out = t_new + 1u; }
    }
    return;
  }
  const uint32_t cutbound = C->cutbound, H = C->H, terminal = C->terminal;
  const bool done = V == cutbound && (!terminal || H == T_INF);
  if (!done && !C->clipped) {                            // a clipped window was not limited by the horizon: nobody has to extend
    const uint32_t bound = terminal ? A.max_iter : cutbound;
    for (uint32_t u = tid; u < A.U; u += 1024) {
      const uint32_t st = D.status[u];
      if (st & ST_FINISHED) continue;
      const uint32_t rf = D.rfin[u], ln = D.ulen[u];
      const uint32_t last = ln > rf ? D.pos[u * A.CAP + ln - 1] + 1u : t_new;
      if (last < bound) D.status[u] = st | ST_EXTEND;
    }
  }
  if (tid == 0) {
    C->evals += ev; C->iterations += t_new - t_final;
    C->t_final = t_new;
    if (C->wmax) C->wmax = min(A.max_iter, 2u * C->wmax);   // a clean window: look twice as far next time
    if (done) {
      if (C->do_terminal) { C->evals += C->live_all; C->iterations += 1u; }  // the call that found freq == 1 / nothing still counted
      C->done = 1u; C->n_out = t_new;
    }
  }
}

// ---- partition view of the index -------------------------------------------------------------------------------------
__global__ void pv_key_kernel(const uint32_t* __restrict__ list_part, uint32_t n, uint32_t U, uint64_t* __restrict__ key, uint32_t* __restrict__ val) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const uint32_t lp = list_part[c];
  key[c] = (lp & 0x80000000u) ? (uint64_t)U : (uint64_t)lp;
  val[c] = c;
}
__global__ void pv_bounds_kernel(const uint64_t* __restrict__ key, uint32_t n, uint32_t U, uint32_t* __restrict__ off) {
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u > U) return;
  uint32_t lo = 0, len = n;
  while (len > 0) { const uint32_t half = len >> 1; if (key[lo + half] < (uint64_t)u) { lo += half + 1; len -= half + 1; } else len = half; }
  off[u] = lo;
}
__global__ void pv_lid_kernel(const uint32_t* __restrict__ ucodes, uint32_t n, uint32_t n_single, uint32_t* __restrict__ lid) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  lid[ucodes[i]] = i < n_single ? i : (LID_MULTI | (i - n_single));
}
__global__ void pv_fwdl_kernel(const uint32_t* __restrict__ fwd_ids, uint64_t n, const uint32_t* __restrict__ lid, uint32_t* __restrict__ fwdl) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t c = fwd_ids[i];
  fwdl[i] = c == 0xFFFFFFFFu ? LID_NONE : __ldg(lid + c);
}
__global__ void pv_segkey_kernel(const uint16_t* __restrict__ seg_part, uint32_t uniform_parts, uint64_t G, uint64_t* __restrict__ key, uint32_t* __restrict__ val) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  key[g] = uniform_parts ? g % uniform_parts : (uint64_t)seg_part[g];
  val[g] = (uint32_t)g;
}
__global__ void pv_mlen_kernel(const uint32_t* __restrict__ ucodes, uint32_t n_single, uint32_t n_multi, const uint32_t* __restrict__ post_off,
                               uint32_t* __restrict__ ub, unsigned long long* total) {
  const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_multi) return;
  const uint32_t c = ucodes[n_single + m];
  const uint32_t len = post_off[c + 1] - post_off[c];
  if (ub) ub[m] = len;
  if (total) atomicAdd(total, (unsigned long long)len);
}
__global__ void pv_status_kernel(uint32_t* status, uint32_t U) {
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < U) status[u] = ST_EXTEND;
}

template <int C>
int launch_extend_c(msspe_ctx* c, const PartArgs& A, cudaStream_t st) {
  if (C == 1) { part_extend_kernel<1><<<dim3(A.U, A.ndirs), EXT_T, 0, st>>>(A); return MSSPE_OK; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(A.U * C, A.ndirs); cfg.blockDim = dim3(EXT_T); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  MSSPE_CUDA_TRY(c, cudaLaunchKernelEx(&cfg, part_extend_kernel<C>, A));
  return MSSPE_OK;
}
int launch_extend(msspe_ctx* c, const PartArgs& A, int csize, cudaStream_t st) {
  switch (csize) {
    case 8: return launch_extend_c<8>(c, A, st);
    case 4: return launch_extend_c<4>(c, A, st);
    case 2: return launch_extend_c<2>(c, A, st);
    default: return launch_extend_c<1>(c, A, st);
  }
}

uint32_t bits_for(uint32_t v) { uint32_t b = 1; while ((v >> b) != 0u && b < 32) b++; return b; }

}  // namespace
