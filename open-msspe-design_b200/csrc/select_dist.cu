// select_dist.cu -- ONE design job over several GPUs: the per-partition greedy loop (select_part.cu) with the partitions
// (alignment columns) sharded over the ranks.  Replaces find_candidates_kmers (od-msspe/src/main.rs:331-406) for both
// directions; every rank returns the complete candidate lists, bit-identical to the single-GPU loop on the whole input.
//
// Each rank's ctx holds, for EVERY genome, the columns of a contiguous range of partitions, builds the index of those
// columns (K1 + K2, no communication) and runs the unit kernel for its own partitions only.  What is global is small
// and is REPLICATED: every round the ranks all-gather their units' not-yet-final entries (fixed-size records, a few KB)
// and every rank merges the same table with the same kernels as one GPU does (part_gather / part_merge / part_plan).
// Lists that lie inside one rank are checked there (part_stage / part_verify); a word that occurs on several ranks is
// a "cross list": its live count per iteration is the SUM of per-rank cover-time histograms, so the ranks all-reduce one
// buffer per round (histograms of the staged cross lists, each rank's best local external winner, per-partition
// liveness of the cross lists for the tie score) and every rank takes the same decision.  Two NCCL collectives per
// round, enqueued on the ctx stream; rounds are batched without host round trips exactly as on one GPU.  NCCL is bound
// with dlopen (the library the host process already uses, e.g. torch's), no link-time dependency.
//
// The f32 tie score of a cross list (main.rs:268-281) is a sum in first-seen partition order.  With one or two live
// partitions the sum is commutative; from three on the order matters, so the iteration is "asked": the loop finalises
// up to it, and the next round's buffer carries the first live genome of every partition at that iteration.
// Limits of this path (MSSPE_ERR_CAPACITY, never a wrong answer): more than DIST_XCAP cross lists staged in one round;
// a cross list with postings in more than DIST_SL partitions of one rank (or 24 live partitions in total).
#include <dlfcn.h>
#include <nccl.h>
#include <time.h>

#include "select_part.cuh"

namespace {

// per call (DistArgs): kmax = not-yet-final entries a unit may hold (fixed-size exchange record), wmax = iterations verified
// ahead per round, xw = wmax + 4 words per cross-list row ([0] live count at t_final, [1 + i] cover-time histogram)
constexpr uint32_t DIST_XCAP = 32768;   // most cross lists staged per round (error beyond); 4096 was short at the complete configs[4] on 8 ranks
constexpr uint32_t DIST_XSTAGE = 256;   // rows of the exchange buffer when the job has more cross lists than that: the window adapts
constexpr uint32_t DIST_SL = 3;         // partitions of one cross list a rank can report per round
constexpr uint32_t DIST_PW = 3;         // words per reported partition: unit + 1, last cover time, first live genome at the queried iteration
constexpr uint32_t VR_WORDS = 8;        // local best record: tv, cnt, score bits, code lo, code hi, n_same, has, pad (+ touched-unit mask)

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api(std::string* err) {
  static NcclApi api;
  if (api.lib) return &api;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { if (err) *err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return nullptr; }
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
  api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
  api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.AllReduce) { if (err) *err = "libnccl.so.2 lacks the expected symbols"; return nullptr; }
  api.lib = h;
  return &api;
}

struct DistState { NcclApi* api; ncclComm_t comm; int rank, world; uint32_t rows_hint; };   // rows_hint: exchange-buffer rows that sufficed last time

#define MSSPE_NCCL_TRY(ctx, ds, expr)                                                                        \
  do {                                                                                                       \
    ncclResult_t _r = (expr);                                                                                \
    if (_r != ncclSuccess) {                                                                                 \
      (ctx)->set_error("%s failed: %s", #expr, (ds)->api->GetErrorString ? (ds)->api->GetErrorString(_r) : "?"); \
      return MSSPE_ERR_CUDA;                                                                                 \
    }                                                                                                        \
  } while (0)

// per-direction distributed state (device pointers)
struct DistDir {
  // exchange of the units' entries
  PEntry* send_e; uint4* send_h;        // [UM * KMAX], [UM] {nnf, status, rfin, 0} ; ulive in send_l
  unsigned long long* send_l;           // [UM]
  unsigned char* sendbuf; unsigned char* recvbuf;
  uint32_t* g_extcov; uint32_t* g_rfin; // [U_pad]
  // cross lists
  uint32_t n_x; const unsigned long long* xcodes; const uint32_t* xlen_desc; uint32_t* ub_x; uint32_t* x_local; const uint32_t* m_xid; uint32_t* xparts;
  uint32_t* xflags; uint32_t* xstage; uint32_t* n_xstage;   // staged cross lists of this round (xid order)
  uint32_t* xbuf;                       // the all-reduced buffer
  uint4* cviol;                         // [XCAP] cross-list violations {tv, cnt, score bits, xid}
  uint32_t* err;
};

struct DistArgs {
  DistDir x[2];
  int rank, world;
  uint32_t UM, U_pad, U_loc, UW;        // units per rank (max), padded total, local units, mask words
  uint32_t xcap;                        // cross lists that can be staged per round: min(DIST_XCAP, cross lists of the job)
  uint32_t kmax, xw, wmax;              // not-yet-final entries per unit in the exchange record; words per cross-list row; window cap
  uint32_t xb_off_mt, xb_off_pb, xb_off_vr, xb_off_err, xb_words, vr_words;
};

__device__ __forceinline__ uint32_t gu_of(const DistArgs& X, uint32_t u_local) { return (uint32_t)X.rank * X.UM + u_local; }

// ---- exchange of the unit tables ------------------------------------------------------------------------------------
__global__ void dist_pack_kernel(PartArgs A, DistArgs X) {
  const PartDir& D = A.d[blockIdx.y];
  const DistDir& Q = X.x[blockIdx.y];
  const uint32_t u = blockIdx.x;        // < UM
  const int tid = threadIdx.x;
  uint32_t nnf = 0, st = ST_FINISHED, rf = 0; unsigned long long ul = 0;
  if (u < X.U_loc) { rf = D.rfin[u]; nnf = D.ulen[u] - rf; st = D.status[u]; ul = D.ulive[u]; }
  for (uint32_t i = tid; i < X.kmax; i += blockDim.x) {
    PEntry e; e.freq = 0; e.cid = 0; e.tied = 0; e.pad = 0; e.live_before = 0; e.code = 0;
    if (i < nnf) e = D.entries[(unsigned long long)u * A.CAP + rf + i];
    Q.send_e[(unsigned long long)u * X.kmax + i] = e;
  }
  if (tid == 0) { Q.send_h[u] = make_uint4(nnf, st, rf, 0u); Q.send_l[u] = ul; }
  if (u == 0) for (uint32_t i = tid; i < X.xw && i < A.max_iter + 2u; i += blockDim.x) D.mt[i] = 0u;   // ties of this rank's local lists (part_verify adds to it)
}

// recvbuf = world blocks of {entries[UM*KMAX], hdr[UM], live[UM]} -> the view V (contiguous over the padded units)
__global__ void dist_unpack_kernel(PartArgs AV, DistArgs X, size_t rank_stride) {
  const PartDir& V = AV.d[blockIdx.y];
  const DistDir& Q = X.x[blockIdx.y];
  const uint32_t gu = blockIdx.x;       // < U_pad
  const uint32_t r = gu / X.UM, u = gu - r * X.UM;
  const unsigned char* blk = Q.recvbuf + (size_t)r * rank_stride;   // both directions of a rank travel in one block
  const PEntry* e = reinterpret_cast<const PEntry*>(blk) + (size_t)u * X.kmax;
  const uint4* h = reinterpret_cast<const uint4*>(blk + (size_t)X.UM * X.kmax * sizeof(PEntry));
  const unsigned long long* l = reinterpret_cast<const unsigned long long*>(blk + (size_t)X.UM * X.kmax * sizeof(PEntry) + (size_t)X.UM * sizeof(uint4));
  for (uint32_t i = threadIdx.x; i < X.kmax; i += blockDim.x) V.entries[(unsigned long long)gu * X.kmax + i] = e[i];
  if (threadIdx.x == 0) {
    const uint4 hh = h[u];
    V.ulen[gu] = hh.x; V.rfin[gu] = 0u; V.status[gu] = hh.y; V.ulive[gu] = l[u];
    Q.g_rfin[gu] = hh.z;
    V.ext_cov[gu] = Q.g_extcov[gu] + hh.z;      // partition_coverage of the unit's first not-yet-final entry
  }
}

// The verify window must not stage more cross lists than the exchange buffer has rows: a list can only matter where its
// global length reaches the winning frequency, so the window ends where more than xcap lists are that long (xlen_desc =
// the global lengths, descending; replicated, so every rank clips alike).
__global__ void dist_clip_kernel(PartArgs AV, DistArgs X) {
  const PartDir& V = AV.d[blockIdx.x];
  const DistDir& Q = X.x[blockIdx.x];
  PartCtl* C = V.ctl;
  if (C->done || threadIdx.x != 0 || Q.n_x <= X.xcap) return;
  const uint32_t t_final = C->t_final;
  uint32_t Vend = C->V;
  if (Vend <= t_final) return;
  const uint32_t need = Q.xlen_desc[X.xcap];             // the (xcap + 1)-th longest list: frequencies at or below it stage too many
  if (V.win_freq[Vend - 1u - t_final] > need && !(C->do_terminal && need >= 2u)) return;
  uint32_t lo = 0, n = Vend - t_final;                  // first window position whose winning frequency is <= need
  while (n > 0) { const uint32_t half = n >> 1; if (V.win_freq[lo + half] > need) { lo += half + 1; n -= half + 1; } else n = half; }
  if (lo == 0) { atomicOr(Q.err, 1u); return; }
  Vend = t_final + lo;
  C->V = Vend; C->t_hi = Vend; C->do_terminal = 0u; C->clipped = 1u; C->fmin = V.win_freq[lo - 1u];
}

// positions of the merged view back into this rank's slots (cover tokens refer to them)
__global__ void dist_posback_kernel(PartArgs A, PartArgs AV, DistArgs X) {
  const PartDir& D = A.d[blockIdx.y];
  const PartDir& V = AV.d[blockIdx.y];
  if (D.ctl->done) return;
  const uint32_t u = blockIdx.x;
  if (u >= X.U_loc) return;
  const uint32_t gu = gu_of(X, u);
  const uint32_t rf = D.rfin[u], nnf = D.ulen[u] - rf;
  for (uint32_t i = threadIdx.x; i < nnf; i += blockDim.x) D.pos[(unsigned long long)u * A.CAP + rf + i] = V.pos[(unsigned long long)gu * X.kmax + i];
}

// ---- local multi-partition lists: best external-winner candidate of this rank -----------------------------------------
__global__ void __launch_bounds__(256) dist_localbest_kernel(PartArgs A, DistArgs X) {
  const PartDir& D = A.d[blockIdx.x];
  const DistDir& Q = X.x[blockIdx.x];
  PartCtl* C = D.ctl;
  if (C->done) return;
  __shared__ unsigned long long s_best;
  __shared__ uint32_t s_bestc, s_same;
  const int tid = threadIdx.x;
  const uint32_t nv = C->n_viol, vmin = C->vmin;
  uint32_t* vr = Q.xbuf + X.xb_off_vr + (size_t)X.rank * X.vr_words;
  for (uint32_t i = tid; i < X.vr_words; i += 256) vr[i] = 0u;
  // ties of the local lists (counted by part_verify into D.mt) travel in the summed buffer
  const uint32_t W = C->t_hi - C->t_final;
  for (uint32_t i = tid; i < W && i < X.xw && i < A.max_iter + 2u; i += 256) Q.xbuf[X.xb_off_mt + i] = D.mt[i];
  if (tid == 0) { s_best = 0ull; s_bestc = 0u; s_same = 0u; Q.xbuf[X.xb_off_err] = *Q.err ? 1u : 0u; }   // a limit hit on any rank stops all of them
  __syncthreads();
  if (vmin == T_INF) return;
  for (uint32_t k = tid; k < nv; k += 256) { const uint4 v = D.viol[k]; if (v.x == vmin) atomicMax(&s_best, ((unsigned long long)v.y << 32) | v.z); }
  __syncthreads();
  for (uint32_t k = tid; k < nv; k += 256) {
    const uint4 v = D.viol[k];
    if (v.x == vmin && ((((unsigned long long)v.y << 32) | v.z) == s_best)) atomicMax(&s_bestc, 0xFFFFFFFFu - v.w);
    if (v.x == vmin && v.y == (uint32_t)(s_best >> 32)) atomicAdd(&s_same, 1u);
  }
  __syncthreads();
  const uint32_t c = 0xFFFFFFFFu - s_bestc;
  const unsigned long long code = D.codes[c];
  if (tid == 0) {
    vr[0] = vmin; vr[1] = (uint32_t)(s_best >> 32); vr[2] = (uint32_t)s_best; vr[3] = (uint32_t)code; vr[4] = (uint32_t)(code >> 32);
    vr[5] = s_same; vr[6] = 1u; vr[7] = c;
  }
  // units its postings touch (all of them: main.rs:371-378 raises partition_coverage for covered ones too)
  for (uint32_t i = D.post_off[c] + tid; i < D.post_off[c + 1]; i += 256) {
    const uint32_t gu = gu_of(X, part_of(A, D.postings[i]));
    atomicOr(&vr[VR_WORDS + (gu >> 5)], 1u << (gu & 31u));
  }
}

// ---- cross lists ---------------------------------------------------------------------------------------------------
__global__ void dist_xflag_kernel(PartArgs A, DistArgs X) {
  const DistDir& Q = X.x[blockIdx.y];
  const PartCtl* C = A.d[blockIdx.y].ctl;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Q.n_x) return;
  Q.xflags[i] = (!C->done && Q.ub_x[i] >= C->fmin) ? 1u : 0u;
}
__global__ void dist_xscatter_kernel(PartArgs A, DistArgs X, int d, const uint32_t* scan) {
  const DistDir& Q = X.x[d];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Q.n_x) return;
  const bool f = (!A.d[d].ctl->done && Q.ub_x[i] >= A.d[d].ctl->fmin);
  if (f) { const uint32_t s = scan[i]; if (s < X.xcap) Q.xstage[s] = i; }
  if (i == Q.n_x - 1) { const uint32_t n = scan[i] + (f ? 1u : 0u); *Q.n_xstage = n; if (n > X.xcap) atomicOr(Q.err, 1u); }
}

// one block per staged cross list: this rank's cover-time histogram of the window, its live count at t_final, and the
// last cover time of its postings per local partition (liveness of the partition at any iteration)
__global__ void __launch_bounds__(128) dist_xhist_kernel(PartArgs A, DistArgs X) {
  const PartDir& D = A.d[blockIdx.y];
  const DistDir& Q = X.x[blockIdx.y];
  const PartCtl* C = D.ctl;
  if (C->done) return;
  extern __shared__ uint32_t h[];
  __shared__ uint32_t tp[8], tmx[8], tnf[8];
  __shared__ unsigned long long sh[34];
  const int tid = threadIdx.x;
  const uint32_t ns = min(*Q.n_xstage, X.xcap), t_final = C->t_final, t_hi = C->t_hi, tq = C->tq;
  for (uint32_t s = blockIdx.x; s < ns; s += gridDim.x) {
    const uint32_t xid = Q.xstage[s];
    const uint32_t m = Q.x_local[xid];
    if (m == 0xFFFFFFFFu) continue;                  // not on this rank: its row stays zero
    const uint32_t c = D.ucodes[D.n_single + m];
    const uint32_t a = D.post_off[c], b = D.post_off[c + 1];
    for (uint32_t i = tid; i < X.xw; i += 128) h[i] = 0u;
    if (tid < 8) { tp[tid] = 0xFFFFFFFFu; tmx[tid] = 0u; tnf[tid] = 0xFFFFFFFFu; }
    __syncthreads();
    unsigned long long l0 = 0;
    for (uint32_t i = a + tid; i < b; i += 128) {
      const uint32_t g = __ldg(D.postings + i);
      const uint32_t tm = time_of(D, g);
      if (tm >= t_final) { l0++; if (tm < t_hi) atomicAdd(&h[1u + tm - t_final], 1u); }
      const uint32_t p = part_of(A, g);
      int k = 0;
      for (; k < 8; k++) {
        const uint32_t old = atomicCAS(&tp[k], 0xFFFFFFFFu, p);
        if (old == 0xFFFFFFFFu || old == p) {
          atomicMax(&tmx[k], tm);
          if (tq != T_INF && tm >= tq) atomicMin(&tnf[k], g / A.uniform_parts);   // genome of the first posting still live at the queried iteration
          break;
        }
      }
      if (k == 8) atomicOr(Q.err, 2u);
    }
    const uint32_t L0 = (uint32_t)block_sum_u64<128>(l0, sh);
    uint32_t* row = Q.xbuf + (size_t)s * X.xw;
    for (uint32_t i = 1 + tid; i < X.xw; i += 128) row[i] = h[i];
    if (tid == 0) {
      row[0] = L0;
      uint32_t* pb = Q.xbuf + X.xb_off_pb + ((size_t)s * X.world + X.rank) * (DIST_SL * DIST_PW);
      uint32_t n = 0;
      for (int k = 0; k < 8; k++) if (tp[k] != 0xFFFFFFFFu) {
        if (n < DIST_SL) { pb[DIST_PW * n] = gu_of(X, tp[k]) + 1u; pb[DIST_PW * n + 1] = tmx[k]; pb[DIST_PW * n + 2] = tnf[k]; }
        n++;
      }
      if (n > DIST_SL) atomicOr(Q.err, 2u);
    }
    __syncthreads();
  }
}

// partition_coverage of padded unit gu at iteration t from the replicated view
__device__ uint32_t cov_view(const PartArgs& AV, const PartDir& V, uint32_t gu, uint32_t t) {
  uint32_t lo = 0, n = V.ulen[gu];
  const uint32_t* ps = V.pos + (unsigned long long)gu * AV.CAP;
  while (n > 0) { const uint32_t half = n >> 1; if (ps[lo + half] < t) { lo += half + 1; n -= half + 1; } else n = half; }
  return V.ext_cov[gu] + lo;            // ext_cov of the view already holds the unit's final entries
}

// every rank, on the summed buffer: would a staged cross list have beaten the merged winner of some iteration?
__global__ void dist_decide_kernel(PartArgs AV, DistArgs X) {
  const PartDir& V = AV.d[blockIdx.y];
  const DistDir& Q = X.x[blockIdx.y];
  PartCtl* C = V.ctl;
  if (C->done) return;
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t ns = min(*Q.n_xstage, X.xcap);
  if (s >= ns) return;
  const uint32_t xid = Q.xstage[s];
  const uint32_t* row = Q.xbuf + (size_t)s * X.xw;
  const uint32_t L0 = row[0], t_final = C->t_final, Vend = C->V, t_hi = C->t_hi, fmin = C->fmin;
  Q.ub_x[xid] = L0;
  Q.cviol[s] = make_uint4(T_INF, 0u, 0u, xid);
  if (L0 < fmin) return;
  const unsigned long long code = Q.xcodes[xid];
  uint32_t covered = 0;
  for (uint32_t i = 0; i < t_hi - t_final; i++) {
    const uint32_t t = t_final + i, cnt = L0 - covered;
    covered += row[1 + i];
    const bool term = t >= Vend;
    const uint32_t F = term ? 1u : V.win_freq[i];
    if (cnt < F || (term && cnt < 2u)) continue;
    // partitions with a live posting at t (a partition is live as long as its last cover time is >= t)
    uint32_t lg[24], ln[24]; int live = 0; bool over = false;
    for (int r = 0; r < X.world; r++) {
      const uint32_t* pb = Q.xbuf + X.xb_off_pb + ((size_t)s * X.world + r) * (DIST_SL * DIST_PW);
      for (uint32_t k = 0; k < DIST_SL; k++)
        if (pb[DIST_PW * k] && pb[DIST_PW * k + 1] >= t) { if (live < 24) { lg[live] = pb[DIST_PW * k] - 1u; ln[live] = pb[DIST_PW * k + 2]; live++; } else over = true; }
    }
    if (over) { atomicOr(Q.err, 2u); return; }
    const bool strict = term || cnt > F;
    if (!strict) atomicAdd(V.mt + i, 1u);
    // the reference's f32 tie score (main.rs:268-281): terms in the order the partitions are first seen among the live
    // postings.  One or two terms: the sum is commutative.  Three or more: the order matters, and only the queried
    // iteration carries it (first live genome per partition) -- otherwise ask for it and stop here.
    if (live > 2) {
      if (t != C->tq) { atomicMin(&C->tq_next, t); return; }
      for (int a2 = 1; a2 < live; a2++) {   // order by (first live genome, partition): global segment order
        const uint32_t kg = lg[a2], kn = ln[a2]; int b2 = a2 - 1;
        while (b2 >= 0 && (ln[b2] > kn || (ln[b2] == kn && lg[b2] > kg))) { lg[b2 + 1] = lg[b2]; ln[b2 + 1] = ln[b2]; b2--; }
        lg[b2 + 1] = kg; ln[b2 + 1] = kn;
      }
    }
    float score = 0.0f;
    for (int a2 = 0; a2 < live; a2++) score = __fadd_rn(score, __fdiv_rn(1.0f, __fadd_rn((float)cov_view(AV, V, lg[a2], t), 1.0f)));
    bool wins = strict;
    if (!wins) {
      const float wsc = __fdiv_rn(1.0f, __fadd_rn((float)V.win_cov[i], 1.0f));
      wins = score > wsc || (score == wsc && code < V.win_code[i]);
    }
    if (wins) {
      Q.cviol[s] = make_uint4(t, cnt, __float_as_uint(score), xid);
      atomicMin(&C->vmin, t);
      return;
    }
  }
}

// every rank, identically: finalise the winners before the horizon / before the earliest external winner; apply it
__global__ void __launch_bounds__(1024) dist_finalize_kernel(PartArgs A, PartArgs AV, DistArgs X) {
  const PartDir& D = A.d[blockIdx.x];
  const PartDir& V = AV.d[blockIdx.x];
  const DistDir& Q = X.x[blockIdx.x];
  PartCtl* C = D.ctl;
  if (C->done) return;
  __shared__ unsigned long long sh[34];
  __shared__ unsigned long long s_best, s_code;
  __shared__ uint32_t s_tv, s_same, s_src;
  const int tid = threadIdx.x;
  const uint32_t t_final = C->t_final, Vend = C->V;
  const uint32_t ns = min(*Q.n_xstage, X.xcap);
  if (tid == 0 && Q.xbuf[X.xb_off_err]) atomicOr(Q.err, 8u);
  // earliest external winner over the ranks' local candidates and the cross lists
  if (tid == 0) { s_tv = T_INF; s_best = 0ull; s_code = ~0ull; s_same = 0u; s_src = 0xFFFFFFFFu; }
  __syncthreads();
  for (uint32_t r = tid; r < (uint32_t)X.world; r += 1024) { const uint32_t* vr = Q.xbuf + X.xb_off_vr + (size_t)r * X.vr_words; if (vr[6]) atomicMin(&s_tv, vr[0]); }
  for (uint32_t s = tid; s < ns; s += 1024) atomicMin(&s_tv, Q.cviol[s].x);
  __syncthreads();
  // an iteration whose tie needs the first-seen order of >= 3 partitions comes first: finalise up to it and ask (tq)
  const uint32_t tqn = C->tq_next;
  const bool ask = tqn != T_INF && tqn <= s_tv;
  const uint32_t tv = ask ? T_INF : s_tv;
  const bool viol = tv != T_INF;
  const uint32_t t_new = ask ? tqn : (viol ? tv : Vend);
  if (viol) {   // best (count, score, smaller word) among everything that wins at tv; words are unique, so the key decides
    for (uint32_t r = tid; r < (uint32_t)X.world; r += 1024) { const uint32_t* vr = Q.xbuf + X.xb_off_vr + (size_t)r * X.vr_words; if (vr[6] && vr[0] == tv) atomicMax(&s_best, ((unsigned long long)vr[1] << 32) | vr[2]); }
    for (uint32_t s = tid; s < ns; s += 1024) { const uint4 v = Q.cviol[s]; if (v.x == tv) atomicMax(&s_best, ((unsigned long long)v.y << 32) | v.z); }
    __syncthreads();
    for (uint32_t r = tid; r < (uint32_t)X.world; r += 1024) {
      const uint32_t* vr = Q.xbuf + X.xb_off_vr + (size_t)r * X.vr_words;
      if (vr[6] && vr[0] == tv) {
        if ((((unsigned long long)vr[1] << 32) | vr[2]) == s_best) atomicMin(&s_code, ((unsigned long long)vr[4] << 32) | vr[3]);
        if (vr[1] == (uint32_t)(s_best >> 32)) atomicAdd(&s_same, vr[5]);
      }
    }
    for (uint32_t s = tid; s < ns; s += 1024) {
      const uint4 v = Q.cviol[s];
      if (v.x == tv) {
        if ((((unsigned long long)v.y << 32) | v.z) == s_best) atomicMin(&s_code, Q.xcodes[v.w]);
        if (v.y == (uint32_t)(s_best >> 32)) atomicAdd(&s_same, 1u);
      }
    }
    __syncthreads();
    // who holds it: a rank's local list (src = rank) or a cross list (src = world + staged slot)
    for (uint32_t r = tid; r < (uint32_t)X.world; r += 1024) { const uint32_t* vr = Q.xbuf + X.xb_off_vr + (size_t)r * X.vr_words; if (vr[6] && vr[0] == tv && ((((unsigned long long)vr[4] << 32) | vr[3]) == s_code)) s_src = r; }
    for (uint32_t s = tid; s < ns; s += 1024) { const uint4 v = Q.cviol[s]; if (v.x == tv && Q.xcodes[v.w] == s_code) s_src = (uint32_t)X.world + s; }
    __syncthreads();
  }
  // winners before t_new are final (from the replicated view: identical on every rank)
  unsigned long long ev = 0;
  for (uint32_t i = tid; i < t_new - t_final; i += 1024) {
    msspe_candidate o;
    o.code = V.win_code[i]; o.freq = V.win_freq[i]; o.n_tied = V.tied[i] + V.mt[i] + Q.xbuf[X.xb_off_mt + i];
    o.tie_score = __fdiv_rn(1.0f, __fadd_rn((float)V.win_cov[i], 1.0f)); o.reserved = 0u;
    D.out[t_final + i] = o;
    ev += V.tot_live[i];
  }
  ev = block_sum_u64<1024>(ev, sh);
  for (uint32_t u = tid; u < X.U_loc; u += 1024) {   // this rank's units: entries before t_new are final
    uint32_t lo = D.rfin[u], n = D.ulen[u] - lo;
    const uint32_t* ps = D.pos + (unsigned long long)u * A.CAP;
    while (n > 0) { const uint32_t half = n >> 1; if (ps[lo + half] < t_new) { lo += half + 1; n -= half + 1; } else n = half; }
    D.rfin[u] = lo;
  }
  __syncthreads();
  if (viol) {
    const uint32_t cnt = (uint32_t)(s_best >> 32);
    const uint32_t iw = t_new - t_final;
    const bool has_entry = t_new < Vend;
    const uint32_t j = C->n_ext;
    const uint32_t src = s_src;
    const bool cross = src >= (uint32_t)X.world;
    if (tid == 0) {
      msspe_candidate o;
      o.code = s_code; o.freq = cnt; o.tie_score = __uint_as_float((uint32_t)s_best); o.reserved = 0u;
      const bool tie_case = has_entry && V.win_freq[iw] == cnt;
      o.n_tied = tie_case ? V.tied[iw] + V.mt[iw] + Q.xbuf[X.xb_off_mt + iw] : s_same;
      D.out[t_new] = o;
      D.pos[(unsigned long long)A.U * A.CAP + 2u * j] = t_new; D.pos[(unsigned long long)A.U * A.CAP + 2u * j + 1u] = t_new;
      C->evals += ev + (has_entry ? V.tot_live[iw] : C->live_all);
      C->iterations += iw + 1u;
    }
    // partition_coverage of every unit the winner touches (replicated table)
    const uint32_t* mask = cross ? Q.xparts + (size_t)Q.cviol[src - X.world].w * X.UW : Q.xbuf + X.xb_off_vr + (size_t)src * X.vr_words + VR_WORDS;
    for (uint32_t gu = tid; gu < X.U_pad; gu += 1024)
      if ((mask[gu >> 5] >> (gu & 31u)) & 1u) {
        Q.g_extcov[gu] += 1u;
        if (gu / X.UM == (uint32_t)X.rank) D.ext_cov[gu - (uint32_t)X.rank * X.UM] += 1u;
      }
    // this rank's postings of the winner (main.rs:371-378)
    uint32_t c = 0xFFFFFFFFu;
    if (cross) { const uint32_t m = Q.x_local[Q.cviol[src - X.world].w]; if (m != 0xFFFFFFFFu) c = D.ucodes[D.n_single + m]; }
    else if (src == (uint32_t)X.rank) c = Q.xbuf[X.xb_off_vr + (size_t)src * X.vr_words + 7];
    if (c != 0xFFFFFFFFu) {
      for (uint32_t i = D.post_off[c] + tid; i < D.post_off[c + 1]; i += 1024) {
        const uint32_t g = D.postings[i];
        const uint32_t tk = __ldcg(D.token + g);
        if (tk == TK_LIVE || __ldcg(D.pos + tk) >= t_new) { D.token[g] = A.U * A.CAP + 2u * j + (tk != TK_LIVE ? 1u : 0u); D.status[part_of(A, g)] = ST_ROLLBACK | ST_EXTEND; }
      }
    }
    if (tid == 0) {
      C->n_ext = j + 1u; C->rollbacks++;
      const uint32_t gap = t_new - C->last_viol;
      C->last_viol = t_new; C->wmax = min(X.wmax, max(32u, 2u * gap));
      C->t_final = t_new + 1u; C->tq = T_INF;
      if (cnt < A.mms || t_new + 1u >= A.max_iter) { C->done = 1u; C->n_out = t_new + 1u; }
    }
    return;
  }
  const uint32_t cutbound = C->cutbound, H = C->H, terminal = C->terminal;
  const bool done = !ask && Vend == cutbound && (!terminal || H == T_INF);
  if (!done && !C->clipped) {
    const uint32_t bound = terminal ? A.max_iter : cutbound;
    for (uint32_t u = tid; u < X.U_loc; u += 1024) {
      const uint32_t st = D.status[u];
      if (st & ST_FINISHED) continue;
      const uint32_t rf = D.rfin[u], ln = D.ulen[u];
      const uint32_t last = ln > rf ? D.pos[(unsigned long long)u * A.CAP + ln - 1] + 1u : t_new;
      if (last < bound) D.status[u] = st | ST_EXTEND;
    }
  }
  if (tid == 0) {
    C->evals += ev; C->iterations += t_new - t_final;
    C->t_final = t_new;
    C->tq = ask ? tqn : T_INF;
    if (C->wmax && !ask) C->wmax = min(X.wmax, 2u * C->wmax);
    if (done) {
      if (C->do_terminal) { C->evals += C->live_all; C->iterations += 1u; }
      C->done = 1u; C->n_out = t_new;
    }
  }
}

// ---- set-up: which words occur on several ranks ----------------------------------------------------------------------
__global__ void dist_crossflag_kernel(const unsigned long long* __restrict__ codes, uint32_t n, const unsigned long long* __restrict__ all, const uint32_t* __restrict__ counts,
                                      uint32_t dmax, int rank, int world, const uint32_t* __restrict__ list_part, uint32_t* __restrict__ lp_out,
                                      uint32_t* __restrict__ own) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const unsigned long long code = codes[c];
  bool cross = false, lower = false;
  for (int r = 0; r < world && !cross; r++) {        // ranks in ascending order: the first hit tells whether a lower rank holds the word
    if (r == rank) continue;
    const unsigned long long* a = all + (size_t)r * dmax;
    uint32_t lo = 0, len = counts[2 * r];
    while (len > 0) { const uint32_t half = len >> 1; if (a[lo + half] < code) { lo += half + 1; len -= half + 1; } else len = half; }
    cross = lo < counts[2 * r] && a[lo] == code;
    lower = cross && r < rank;
  }
  lp_out[c] = list_part[c] | (cross ? 0x80000000u : 0u);
  own[c] = (cross && !lower) ? 1u : 0u;               // the lowest rank holding a cross word reports it
}
__global__ void dist_ownscatter_kernel(const unsigned long long* __restrict__ codes, uint32_t n, const uint32_t* __restrict__ own, const uint32_t* __restrict__ scan,
                                       unsigned long long* __restrict__ out) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n && own[c]) out[scan[c]] = codes[c];
}
__global__ void dist_fill_kernel(unsigned long long* p, uint64_t n, unsigned long long v) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void dist_padcodes_kernel(const unsigned long long* __restrict__ codes, uint32_t n, uint32_t dmax, unsigned long long pad, unsigned long long* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < dmax) out[i] = i < n ? codes[i] : pad;
}
// in the sorted concatenation of all ranks' codes: first element of a run of >= 2 equal words
// one warp per local multi-partition list: is it a cross list, and which units do its postings touch?  (One THREAD per list
// walked up to one posting per genome sequentially: 1.8 ms of the set-up at 40,000 genomes.)
__global__ void __launch_bounds__(256) dist_xmap_kernel(PartArgs A, DistArgs X, int d, uint32_t* m_xid, uint32_t* xlen_local) {
  const PartDir& D = A.d[d];
  const DistDir& Q = X.x[d];
  const uint32_t m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
  if (m >= D.n_multi) return;
  const uint32_t c = D.ucodes[D.n_single + m];
  const unsigned long long code = D.codes[c];
  uint32_t lo = 0, len = Q.n_x;
  while (len > 0) { const uint32_t half = len >> 1; if (Q.xcodes[lo + half] < code) { lo += half + 1; len -= half + 1; } else len = half; }
  const bool is_x = lo < Q.n_x && Q.xcodes[lo] == code;
  if (lane == 0) m_xid[m] = is_x ? lo : 0xFFFFFFFFu;
  if (is_x) {
    const uint32_t p0 = D.post_off[c], p1 = D.post_off[c + 1];
    if (lane == 0) {
      Q.x_local[lo] = m;
      xlen_local[lo] = p1 - p0;
      D.ub[m] = 0u;                                   // never staged by the local check: the cross check owns it
    }
    uint32_t last = 0xFFFFFFFFu;                      // consecutive postings mostly share a unit: one atomic per change
    for (uint32_t i = p0 + lane; i < p1; i += 32) {
      const uint32_t gu = gu_of(X, part_of(A, D.postings[i]));
      if (gu != last) { atomicOr(&Q.xparts[(size_t)lo * X.UW + (gu >> 5)], 1u << (gu & 31u)); last = gu; }
    }
  }
}

__global__ void dist_lenkey_kernel(const uint32_t* __restrict__ len, uint32_t n, uint64_t* __restrict__ key) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) key[i] = (uint64_t)(0xFFFFFFFFu - len[i]);     // ascending sort of the complement = descending lengths
}
__global__ void dist_lenunkey_kernel(const uint64_t* __restrict__ key, uint32_t n, uint32_t* __restrict__ len) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) len[i] = 0xFFFFFFFFu - (uint32_t)key[i];
}
__global__ void dist_status_kernel(PartArgs A) {
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < A.U) A.d[blockIdx.y].status[u] = ST_EXTEND;
}

}  // namespace

extern "C" int msspe_dist_unique_id(uint8_t* out128) {
  if (!out128) return MSSPE_ERR_INVALID;
  std::string err;
  NcclApi* api = nccl_api(&err);
  if (!api) return MSSPE_ERR_STATE;
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return MSSPE_ERR_CUDA;
  memcpy(out128, id.internal, 128);
  return MSSPE_OK;
}

extern "C" int msspe_dist_init(msspe_ctx* c, const uint8_t* id128, int rank, int world) {
  if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return MSSPE_ERR_INVALID;
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  std::string err;
  NcclApi* api = nccl_api(&err);
  if (!api) { c->set_error("%s", err.c_str()); return MSSPE_ERR_STATE; }
  if (c->dist) { DistState* o = (DistState*)c->dist; if (api->CommDestroy) api->CommDestroy(o->comm); delete o; c->dist = nullptr; }
  DistState* ds = new DistState{api, nullptr, rank, world, 0u};
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  ncclResult_t r = api->CommInitRank(&ds->comm, world, id, rank);
  if (r != ncclSuccess) { c->set_error("ncclCommInitRank failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?"); delete ds; return MSSPE_ERR_CUDA; }
  c->dist = ds;
  return MSSPE_OK;
}

void msspe_dist_free(msspe_ctx* c) {
  if (!c->dist) return;
  DistState* ds = (DistState*)c->dist;
  if (ds->api->CommDestroy) ds->api->CommDestroy(ds->comm);
  delete ds;
  c->dist = nullptr;
}

namespace {
struct ScratchList {   // stream-ordered scratch of one attempt, returned to the pool however the attempt ends
  std::vector<void*> p; cudaStream_t st;
  ~ScratchList() { for (void* q : p) cudaFreeAsync(q, st); }
};
int dist_select_impl(msspe_ctx* c, uint32_t max_iter, uint32_t mms, msspe_candidate* out_fwd, uint32_t* n_fwd, msspe_candidate* out_rev,
                     uint32_t* n_rev, uint32_t xstage_rows, bool* retry);
}  // namespace

extern "C" int msspe_select_both_dist(msspe_ctx* c, uint32_t max_iter, uint32_t mms, msspe_candidate* out_fwd, uint32_t* n_fwd,
                                      msspe_candidate* out_rev, uint32_t* n_rev) {
  if (!c) return MSSPE_ERR_INVALID;
  // rows of the per-round exchange buffer: few suffice when few cross-rank lists are as long as the winning frequencies
  // (the verify window adapts to them); an input with many long cross-rank lists reports it on every rank in the same
  // round, and every rank retries with more rows
  int rc = MSSPE_OK;
  if (!c->dist) { c->set_error("msspe_select_both_dist: msspe_dist_init first"); return MSSPE_ERR_STATE; }
  DistState* ds0 = (DistState*)c->dist;
  for (uint32_t rows = std::min(ds0->rows_hint ? ds0->rows_hint : DIST_XSTAGE, DIST_XCAP); ; rows = std::min(rows * 4u, DIST_XCAP)) {
    bool retry = false;
    rc = dist_select_impl(c, max_iter, mms, out_fwd, n_fwd, out_rev, n_rev, rows, &retry);
    if (!(retry && rows < DIST_XCAP)) { if (rc == MSSPE_OK) ds0->rows_hint = rows; break; }   // replicated outcome: every rank keeps the same hint
  }
  return rc;
}

namespace {
int dist_select_impl(msspe_ctx* c, uint32_t max_iter, uint32_t mms, msspe_candidate* out_fwd, uint32_t* n_fwd, msspe_candidate* out_rev,
                     uint32_t* n_rev, uint32_t xstage_rows, bool* retry) {
  if (!c) return MSSPE_ERR_INVALID;
  if (!n_fwd || !n_rev || (max_iter && (!out_fwd || !out_rev))) { c->set_error("msspe_select_both_dist: bad argument"); return MSSPE_ERR_INVALID; }
  if (!c->dist) { c->set_error("msspe_select_both_dist: msspe_dist_init first"); return MSSPE_ERR_STATE; }
  if (!c->built) { c->set_error("msspe_select_both_dist: index not built"); return MSSPE_ERR_STATE; }
  DistState* ds = (DistState*)c->dist;
  NcclApi* N = ds->api;
  const int rank = ds->rank, world = ds->world;
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  *n_fwd = *n_rev = 0;
  for (int d = 0; d < 2; d++) { c->timing.select_evals[d] = 0; c->timing.select_iterations[d] = 0; c->timing.select_ms[d] = 0.f; }
  if (max_iter == 0) return MSSPE_OK;
  const uint64_t G = c->n_segments;
  const uint32_t U_loc = G ? c->max_partition + 1u : 0u;
  if (G && !(c->uniform_parts && c->uniform_parts <= 65536u)) { c->set_error("msspe_select_both_dist: a column shard has records of equal length (every genome, the same columns)"); return MSSPE_ERR_INVALID; }
  if (G >= 0x80000000ull) { c->set_error("msspe_select_both_dist: at most 2^31 segments per rank"); return MSSPE_ERR_CAPACITY; }
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[2], st));
  const bool dbg = getenv("MSSPE_DEBUG_TIMERS") != nullptr;
  auto now_ms = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return 1e3 * ts.tv_sec + 1e-6 * ts.tv_nsec; };
  const double t_begin = now_ms();
  double t_setup = 0.0;
  double t_ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // MSSPE_DEBUG_TIMERS: set-up phases (stream synchronised at each mark)
  double t_mark = t_begin;
  auto mark = [&](int ph) { if (dbg) { cudaStreamSynchronize(st); const double t = now_ms(); t_ph[ph] += t - t_mark; t_mark = t; } };
  ScratchList scratch; scratch.st = st;
  auto alloc = [&](void** p, uint64_t bytes, int fill) -> int {
    MSSPE_CUDA_TRY(c, cudaMallocAsync(p, bytes ? bytes : 4, st));
    scratch.p.push_back(*p);
    if (fill >= 0) MSSPE_CUDA_TRY(c, cudaMemsetAsync(*p, fill, bytes ? bytes : 4, st));
    return MSSPE_OK;
  };
#define DA(ptr, bytes, fill) { int rc2 = alloc((void**)&(ptr), (bytes), (fill)); if (rc2) return rc2; }
  // ---- 1. sizes of all ranks: local units, distinct words per direction ----
  uint32_t* d_cnt = nullptr; uint32_t* d_cnts = nullptr;
  DA(d_cnt, 16, 0); DA(d_cnts, 16ull * world, 0);
  uint32_t h_cnt[4] = {(uint32_t)c->dir[0].n_codes, (uint32_t)c->dir[1].n_codes, U_loc, 0};
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(d_cnt, h_cnt, 16, cudaMemcpyHostToDevice, st));
  MSSPE_NCCL_TRY(c, ds, N->AllGather(d_cnt, d_cnts, 4, ncclUint32, ds->comm, st));
  std::vector<uint32_t> h_cnts(4 * world);
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(h_cnts.data(), d_cnts, 16ull * world, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  mark(0);
  uint32_t UM = 1, dmax[2] = {1, 1};
  for (int r = 0; r < world; r++) { UM = std::max(UM, h_cnts[4 * r + 2]); dmax[0] = std::max(dmax[0], h_cnts[4 * r]); dmax[1] = std::max(dmax[1], h_cnts[4 * r + 1]); }
  const uint32_t U_pad = UM * (uint32_t)world, UW = (U_pad + 31u) / 32u;
  const uint64_t CAP = max_iter;
  if ((uint64_t)std::max(U_loc, 1u) * CAP + 2ull * max_iter >= 0x7FFFFFF0ull) { c->set_error("msspe_select_both_dist: entry table too large"); return MSSPE_ERR_CAPACITY; }
  const uint32_t kbits = 2 * c->cfg.kmer_size;
  if (kbits >= 63) { c->set_error("msspe_select_both_dist: k-mer size %u not supported across ranks", c->cfg.kmer_size); return MSSPE_ERR_INVALID; }
  const unsigned long long pad = 1ull << kbits;

  PartArgs A{}, AV{};
  A.ndirs = 2; A.U = U_loc; A.CAP = (uint32_t)CAP; A.slots = c->slots; A.max_iter = max_iter; A.mms = mms;
  A.uniform_parts = (c->uniform_parts && c->uniform_parts <= 65536u) ? c->uniform_parts : 0u;
  // not-yet-final entries a unit may hold = size of its exchange record: about three times the average share of the winners
  uint32_t kmax = 32;
  while (kmax < 256 && (uint64_t)kmax * U_pad < 3ull * max_iter) kmax <<= 1;
  A.seg_part = c->d_seg_part; A.max_ahead = kmax;
  AV = A; AV.U = U_pad; AV.CAP = kmax; AV.max_ahead = kmax;
  DistArgs X{};
  X.rank = rank; X.world = world; X.UM = UM; X.U_pad = U_pad; X.U_loc = U_loc; X.UW = UW;
  X.vr_words = VR_WORDS + UW; X.kmax = kmax;
  const size_t block_bytes = ((size_t)UM * kmax * sizeof(PEntry) + (size_t)UM * sizeof(uint4) + (size_t)UM * 8 + 15) & ~(size_t)15;
  uint32_t max_nx = 0, max_multi = 0;
  uint32_t* xscan[2] = {nullptr, nullptr};
  // both directions share one all-gather and one all-reduce per round
  unsigned char* send_all = nullptr; unsigned char* recv_all = nullptr; uint32_t* xbuf_all = nullptr;
  DA(send_all, 2 * block_bytes, 0); DA(recv_all, 2 * block_bytes * world, 0);

  for (int d = 0; d < 2; d++) {
    DirIndex& I = c->dir[d];
    DistDir& Q = X.x[d];
    const uint32_t nc = (uint32_t)I.n_codes;
    // ---- 2. words that occur on several ranks ("cross lists") ----
    unsigned long long* sendc = nullptr; unsigned long long* allc = nullptr; uint32_t* lp_dist = nullptr;
    DA(sendc, (uint64_t)dmax[d] * 8, -1); DA(allc, (uint64_t)dmax[d] * 8 * world, -1); DA(lp_dist, ((uint64_t)nc + 1) * 4, -1);
    dist_padcodes_kernel<<<(dmax[d] + 255u) / 256u, 256, 0, st>>>((const unsigned long long*)I.codes, nc, dmax[d], pad, sendc);
    MSSPE_NCCL_TRY(c, ds, N->AllGather(sendc, allc, dmax[d], ncclUint64, ds->comm, st));
    // counts[2 * r + d'] layout expected by the kernel: use d_cnts with stride 4 -> repack on the host (tiny)
    uint32_t* d_cn = nullptr;
    DA(d_cn, 8ull * world, -1);
    std::vector<uint32_t> cn(2 * world);
    for (int r = 0; r < world; r++) { cn[2 * r] = h_cnts[4 * r + d]; cn[2 * r + 1] = 0; }
    MSSPE_CUDA_TRY(c, cudaMemcpyAsync(d_cn, cn.data(), 8ull * world, cudaMemcpyHostToDevice, st));
    MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));   // cn goes out of scope
    uint32_t* own = nullptr; uint32_t* own_scan = nullptr; uint32_t* d_nown = nullptr; uint32_t* d_nowns = nullptr;
    DA(own, ((uint64_t)nc + 1) * 4, 0); DA(own_scan, ((uint64_t)nc + 1) * 4, 0); DA(d_nown, 4, 0); DA(d_nowns, 4ull * world, 0);
    if (nc) {
      dist_crossflag_kernel<<<(nc + 255u) / 256u, 256, 0, st>>>((const unsigned long long*)I.codes, nc, allc, d_cn, dmax[d], rank, world, I.list_part, lp_dist, own);
      int rc0 = msspe_exclusive_scan_u32(c, own, own_scan, nc, d_nown, st);
      if (rc0) return rc0;
    }
    mark(1);
    // the cross words themselves, identically on every rank: every word is reported by the lowest rank that holds it
    MSSPE_NCCL_TRY(c, ds, N->AllGather(d_nown, d_nowns, 1, ncclUint32, ds->comm, st));
    std::vector<uint32_t> h_nown(world);
    MSSPE_CUDA_TRY(c, cudaMemcpyAsync(h_nown.data(), d_nowns, 4ull * world, cudaMemcpyDeviceToHost, st));
    // partition view with the cross lists treated as multi-partition lists (its own synchronisations cover the copy above)
    I.pv_built = false;
    msspe_dev_free(c, I.pv_ucode_off); msspe_dev_free(c, I.pv_ucodes); msspe_dev_free(c, I.pv_fwdl); msspe_dev_free(c, I.pv_useg_off); msspe_dev_free(c, I.pv_usegs);
    I.pv_ucode_off = I.pv_ucodes = I.pv_fwdl = I.pv_useg_off = I.pv_usegs = nullptr;
    uint32_t* saved = I.list_part;
    I.list_part = lp_dist;
    int rc = msspe_partition_view(c, d, st);
    I.list_part = saved;
    I.pv_dist = true;
    if (rc) return rc;
    MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
    mark(2);
    uint32_t n_x = 0, own_max = 1;
    for (int r = 0; r < world; r++) { n_x += h_nown[r]; own_max = std::max(own_max, h_nown[r]); }
    Q.n_x = n_x;
    max_nx = std::max(max_nx, n_x);
    const uint32_t na = own_max * (uint32_t)world;
    unsigned long long* own_send = nullptr; uint64_t *ka = nullptr, *kb = nullptr; uint32_t *va = nullptr, *vb = nullptr;
    DA(own_send, (uint64_t)own_max * 8, -1); DA(ka, (uint64_t)na * 8, -1); DA(kb, (uint64_t)na * 8, -1); DA(va, (uint64_t)na * 4, 0); DA(vb, (uint64_t)na * 4, 0);
    dist_fill_kernel<<<(own_max + 255u) / 256u, 256, 0, st>>>(own_send, own_max, pad);
    if (nc) dist_ownscatter_kernel<<<(nc + 255u) / 256u, 256, 0, st>>>((const unsigned long long*)I.codes, nc, own, own_scan, own_send);
    MSSPE_NCCL_TRY(c, ds, N->AllGather(own_send, ka, own_max, ncclUint64, ds->comm, st));
    rc = msspe_radix_sort_pairs(c, &ka, &va, &kb, &vb, na, kbits + 1, st);   // the paddings sort behind every word
    if (rc) return rc;
    Q.xcodes = reinterpret_cast<const unsigned long long*>(ka);             // first n_x entries
    mark(3);
    // ---- 3. state of the loop (as on one GPU) + the replicated view + exchange buffers ----
    if (I.out_capacity < max_iter) {
      msspe_dev_free(c, I.out); I.out = nullptr;
      MSSPE_CUDA_TRY(c, cudaMallocAsync(&I.out, (uint64_t)max_iter * sizeof(msspe_candidate), c->stream));
      I.out_capacity = max_iter;
    }
    PartDir& P = A.d[d];
    P.codes = I.codes; P.post_off = I.post_off; P.postings = I.postings; P.ucode_off = I.pv_ucode_off; P.ucodes = I.pv_ucodes;
    P.fwdl = I.pv_fwdl; P.useg_off = I.pv_useg_off; P.usegs = I.pv_usegs; P.n_single = I.pv_single; P.n_multi = I.pv_multi_n;
    P.out = I.out;
    max_multi = std::max(max_multi, P.n_multi);
    const uint64_t ne = (uint64_t)std::max(U_loc, 1u) * CAP;
    DA(P.pfreq, ((uint64_t)I.n_codes + 1) * 4, -1);
    DA(P.token, (G + 1) * 4, 0xFF);
    DA(P.ulive, (uint64_t)std::max(U_loc, 1u) * 8, 0);
    DA(P.entries, ne * sizeof(PEntry), -1);
    DA(P.pos, (ne + 2ull * max_iter + 2) * 4, 0xFF);
    DA(P.rfin, (uint64_t)std::max(U_loc, 1u) * 4, 0);
    DA(P.ulen, (uint64_t)std::max(U_loc, 1u) * 4, 0);
    DA(P.status, (uint64_t)std::max(U_loc, 1u) * 4, 0);
    DA(P.ext_cov, (uint64_t)std::max(U_loc, 1u) * 4, 0);
    DA(P.mt, ((uint64_t)max_iter + 2) * 4, 0);
    DA(P.ub, ((uint64_t)P.n_multi + 1) * 4, 0);
    DA(P.stage, ((uint64_t)P.n_multi + 1) * 4, -1);
    DA(P.viol, ((uint64_t)P.n_multi + 1) * 16, -1);
    DA(P.touch, (uint64_t)std::max(U_loc, 1u) * 4, 0);
    DA(P.ctl, sizeof(PartCtl), 0);
    DA(P.win_freq, ((uint64_t)max_iter + 2) * 4, 0);
    DA(P.win_cov, ((uint64_t)max_iter + 2) * 4, 0);
    DA(P.win_code, ((uint64_t)max_iter + 2) * 8, 0);
    P.elist = nullptr; P.order = nullptr; P.tied = nullptr; P.tot_live = nullptr;
    PartDir& Vw = AV.d[d];
    Vw = P;
    const uint64_t nv = (uint64_t)U_pad * kmax;
    DA(Vw.entries, nv * sizeof(PEntry), 0);
    DA(Vw.pos, (nv + 2) * 4, 0xFF);
    DA(Vw.rfin, (uint64_t)U_pad * 4, 0);
    DA(Vw.ulen, (uint64_t)U_pad * 4, 0);
    DA(Vw.status, (uint64_t)U_pad * 4, 0);
    DA(Vw.ext_cov, (uint64_t)U_pad * 4, 0);
    DA(Vw.ulive, (uint64_t)U_pad * 8, 0);
    DA(Vw.elist, nv * 4, -1);
    DA(Vw.order, ((uint64_t)max_iter + 2) * 4, 0);
    DA(Vw.tied, ((uint64_t)max_iter + 2) * 4, 0);
    DA(Vw.tot_live, ((uint64_t)max_iter + 2) * 8, 0);
    DA(Vw.mt, ((uint64_t)max_iter + 2) * 4, 0);      // ties of cross lists (the same on every rank)
    // the window arrays and the control block are shared by the local and the view side
    Q.sendbuf = send_all + (size_t)d * block_bytes; Q.recvbuf = recv_all + (size_t)d * block_bytes;
    Q.send_e = reinterpret_cast<PEntry*>(Q.sendbuf);
    Q.send_h = reinterpret_cast<uint4*>(Q.sendbuf + (size_t)UM * kmax * sizeof(PEntry));
    Q.send_l = reinterpret_cast<unsigned long long*>(Q.sendbuf + (size_t)UM * kmax * sizeof(PEntry) + (size_t)UM * sizeof(uint4));
    DA(Q.g_extcov, (uint64_t)U_pad * 4, 0); DA(Q.g_rfin, (uint64_t)U_pad * 4, 0);
    DA(Q.ub_x, ((uint64_t)n_x + 1) * 4, 0); DA(Q.x_local, ((uint64_t)n_x + 1) * 4, 0xFF); DA(Q.xparts, ((uint64_t)n_x + 1) * UW * 4, 0);
    DA(Q.xflags, ((uint64_t)n_x + 1) * 4, 0); DA(Q.xstage, (uint64_t)DIST_XCAP * 4, 0); DA(Q.n_xstage, 4, 0);
    DA(Q.cviol, (uint64_t)DIST_XCAP * 16, 0xFF); DA(Q.err, 4, 0);
    DA(xscan[d], ((uint64_t)n_x + 1) * 4, 0);
    uint32_t* m_xid = nullptr; uint32_t* xlen_local = nullptr;
    DA(m_xid, ((uint64_t)P.n_multi + 1) * 4, 0xFF); DA(xlen_local, ((uint64_t)n_x + 1) * 4, 0);
    Q.m_xid = m_xid;
    mark(4);
    if (P.n_multi) {
      pv_mlen_kernel<<<(P.n_multi + 255u) / 256u, 256, 0, st>>>(P.ucodes, P.n_single, P.n_multi, P.post_off, P.ub, nullptr);
      dist_xmap_kernel<<<(P.n_multi + 7u) / 8u, 256, 0, st>>>(A, X, d, m_xid, xlen_local);
    }
    if (n_x) {   // global length of every cross list (its first upper bound) and the units it touches
      MSSPE_NCCL_TRY(c, ds, N->AllReduce(xlen_local, Q.ub_x, n_x, ncclUint32, ncclSum, ds->comm, st));
      MSSPE_NCCL_TRY(c, ds, N->AllReduce(Q.xparts, Q.xparts, (size_t)n_x * UW, ncclUint32, ncclSum, ds->comm, st));   // disjoint bits per rank: sum = or
      if (n_x > xstage_rows) {   // global lengths in descending order, for the window clip
        uint64_t *ka2 = nullptr, *kb2 = nullptr; uint32_t *va2 = nullptr, *vb2 = nullptr; uint32_t* desc = nullptr;
        DA(ka2, (uint64_t)n_x * 8, -1); DA(kb2, (uint64_t)n_x * 8, -1); DA(va2, (uint64_t)n_x * 4, 0); DA(vb2, (uint64_t)n_x * 4, 0); DA(desc, (uint64_t)n_x * 4, -1);
        dist_lenkey_kernel<<<(n_x + 255u) / 256u, 256, 0, st>>>(Q.ub_x, n_x, ka2);
        rc = msspe_radix_sort_pairs(c, &ka2, &va2, &kb2, &vb2, n_x, 32, st);
        if (rc) return rc;
        dist_lenunkey_kernel<<<(n_x + 255u) / 256u, 256, 0, st>>>(ka2, n_x, desc);
        Q.xlen_desc = desc;
      }
    }
    // PartCtl: verify windows are bounded by the exchange record
    MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
    mark(5);
  }
  // the all-reduced buffer: rows of the staged cross lists | ties of local lists | reported partitions | local best records | limit flag
  X.xcap = max_nx <= xstage_rows ? std::max<uint32_t>(128u, (max_nx + 127u) & ~127u) : xstage_rows;
  // iterations verified ahead per round: as many as 256 KB of histogram rows per direction allow (62 ... 1020)
  X.wmax = std::min<uint32_t>(std::min<uint32_t>(1020u, std::max<uint32_t>(62u, 65536u / X.xcap > 4u ? 65536u / X.xcap - 4u : 0u)), std::max<uint32_t>(max_iter, 1u));
  X.xw = X.wmax + 4u;
  {
    PartCtl h0; memset(&h0, 0, sizeof h0); h0.wmax = X.wmax; h0.tq = T_INF; h0.tq_next = T_INF;
    for (int d = 0; d < 2; d++) MSSPE_CUDA_TRY(c, cudaMemcpyAsync(A.d[d].ctl, &h0, sizeof h0, cudaMemcpyHostToDevice, st));
    MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  }
  X.xb_off_mt = X.xcap * X.xw;
  X.xb_off_pb = X.xb_off_mt + X.xw;
  X.xb_off_vr = X.xb_off_pb + X.xcap * (uint32_t)world * (DIST_SL * DIST_PW);
  X.xb_off_err = X.xb_off_vr + (uint32_t)world * X.vr_words;
  X.xb_words = X.xb_off_err + 1u;
  DA(xbuf_all, 2ull * X.xb_words * 4, 0);
  for (int d = 0; d < 2; d++) X.x[d].xbuf = xbuf_all + (size_t)d * X.xb_words;
  if (U_loc) {
    dist_status_kernel<<<dim3((U_loc + 255u) / 256u, 2), 256, 0, st>>>(A);
    unsigned mx = 1;
    for (int d = 0; d < 2; d++) mx = std::max<unsigned>(mx, A.d[d].n_single);
    part_init_freq_kernel<<<dim3((mx + 255u) / 256u, 2), 256, 0, st>>>(A);
    if (G) part_init_live_kernel<<<dim3((unsigned)div_up_u64(G, 256), 2), 256, 0, st>>>(A, G);
  }
  const size_t ver_smem = (2 * ((size_t)max_iter + 2) + (U_loc + 31u) / 32u) * 4;
  if (ver_smem > c->smem_optin) { c->set_error("msspe_select_both_dist: max_iterations %u too large for the verify kernel's histogram", max_iter); return MSSPE_ERR_CAPACITY; }
  MSSPE_CUDA_TRY(c, cudaFuncSetAttribute(part_verify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ver_smem));
  const uint64_t unit_items = U_loc ? G / U_loc * c->slots : 0;
  int csize = unit_items >= (1u << 20) ? 8 : unit_items >= (1u << 18) ? 4 : unit_items >= (1u << 16) ? 2 : 1;
  if (const char* e = getenv("MSSPE_PART_CLUSTER")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) csize = v; }
  const uint32_t chunk0 = std::min<uint32_t>(kmax, (uint32_t)std::max<uint64_t>(4, (3ull * max_iter + 2ull * U_pad - 1) / (2ull * U_pad)));
  const uint32_t chunk = 8;
  const unsigned merge_grid = (unsigned)c->sm_count * 2u, ver_grid = (unsigned)c->sm_count * 2u;
  PartCtl* h = reinterpret_cast<PartCtl*>(c->h_ctl);
  uint32_t h_err[2] = {0, 0};
  uint32_t round = 0;
  const uint32_t BATCH = 4;
  mark(6);
  if (dbg) { cudaStreamSynchronize(st); t_setup = now_ms() - t_begin; }
  for (;;) {
    for (uint32_t b = 0; b < BATCH; b++, round++) {
      A.nsteps = round == 0 ? chunk0 : chunk;
      if (U_loc) { KPROF(c, KP_GREEDY_UNIT, st, 0) int rc2 = launch_extend(c, A, csize, st); if (rc2) return rc2; }
      { KPROF(c, KP_GREEDY_MERGE, st, 0) dist_pack_kernel<<<dim3(UM, 2), 64, 0, st>>>(A, X); }
      MSSPE_NCCL_TRY(c, ds, N->AllGather(send_all, recv_all, 2 * block_bytes, ncclUint8, ds->comm, st));
      { KPROF(c, KP_GREEDY_MERGE, st, 0) dist_unpack_kernel<<<dim3(U_pad, 2), 32, 0, st>>>(AV, X, 2 * block_bytes); }
      { KPROF(c, KP_GREEDY_MERGE, st, 0) part_gather_kernel<<<2, 1024, 0, st>>>(AV); }
      { KPROF(c, KP_GREEDY_MERGE, st, 0) part_merge_kernel<<<dim3(merge_grid, 2), 256, 0, st>>>(AV); }
      { KPROF(c, KP_GREEDY_MERGE, st, 0) part_plan_kernel<<<2, 1024, 0, st>>>(AV); }
      if (max_nx > X.xcap) { KPROF(c, KP_GREEDY_MERGE, st, 0) dist_clip_kernel<<<2, 32, 0, st>>>(AV, X); }
      if (U_loc) { KPROF(c, KP_GREEDY_MERGE, st, 0) dist_posback_kernel<<<dim3(U_loc, 2), 32, 0, st>>>(A, AV, X); }
      MSSPE_CUDA_TRY(c, cudaMemsetAsync(xbuf_all, 0, 2 * (size_t)X.xb_words * 4, st));
      if (max_multi) {
        { KPROF(c, KP_GREEDY_MERGE, st, 0) part_stage_kernel<<<dim3((max_multi + 255u) / 256u, 2), 256, 0, st>>>(A); }
        { KPROF(c, KP_GREEDY_VERIFY, st, 0) part_verify_kernel<<<dim3(ver_grid, 2), VER_T, ver_smem, st>>>(A); }
      }
      { KPROF(c, KP_GREEDY_MERGE, st, 0) dist_localbest_kernel<<<2, 256, 0, st>>>(A, X); }
      if (max_nx) {
        { KPROF(c, KP_GREEDY_MERGE, st, 0) dist_xflag_kernel<<<dim3((max_nx + 255u) / 256u, 2), 256, 0, st>>>(A, X); }
        for (int d = 0; d < 2; d++) {
          if (!X.x[d].n_x) continue;
          int rc2 = msspe_exclusive_scan_u32(c, X.x[d].xflags, xscan[d], X.x[d].n_x, nullptr, st);
          if (rc2) return rc2;
          dist_xscatter_kernel<<<(X.x[d].n_x + 255u) / 256u, 256, 0, st>>>(A, X, d, xscan[d]);
        }
        { KPROF(c, KP_GREEDY_VERIFY, st, 0) dist_xhist_kernel<<<dim3(ver_grid, 2), 128, (size_t)X.xw * 4, st>>>(A, X); }
      }
      MSSPE_NCCL_TRY(c, ds, N->AllReduce(xbuf_all, xbuf_all, 2 * (size_t)X.xb_words, ncclUint32, ncclSum, ds->comm, st));
      if (max_nx) { KPROF(c, KP_GREEDY_VERIFY, st, 0) dist_decide_kernel<<<dim3(X.xcap / 128, 2), 128, 0, st>>>(AV, X); }
      { KPROF(c, KP_GREEDY_MERGE, st, 0) dist_finalize_kernel<<<2, 1024, 0, st>>>(A, AV, X); }
    }
    MSSPE_CUDA_TRY(c, cudaGetLastError());
    for (int d = 0; d < 2; d++) {
      MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&h[d], A.d[d].ctl, sizeof(PartCtl), cudaMemcpyDeviceToHost, st));
      MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&h_err[d], X.x[d].err, 4, cudaMemcpyDeviceToHost, st));
    }
    MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
    if ((h_err[0] | h_err[1]) & (1u | 4u | 8u)) {   // bits every rank sees in the same round (2 = local: it arrives as 8 one round later)
      c->set_error("msspe_select_both_dist: limit of the multi-GPU loop reached (flags %u/%u: 1 = more than %u cross-rank lists staged in a round, "
                   "2 = a cross-rank list in more than 3 partitions of one rank or 24 in total)",
                   h_err[0], h_err[1], DIST_XCAP);
      *retry = ((h_err[0] | h_err[1]) & (2u | 4u)) == 0;    // only the number of rows was short
      cudaStreamSynchronize(st);
      return MSSPE_ERR_CAPACITY;
    }
    if (h[0].done && h[1].done) break;     // the control blocks are replicated: every rank leaves in the same batch
    if (round > 4u * max_iter + 64u) { c->set_error("msspe_select_both_dist: loop did not converge"); return MSSPE_ERR_STATE; }
  }
  {  // a limit hit on one rank in the very last round: agree on it before anybody reports success
    MSSPE_NCCL_TRY(c, ds, N->AllReduce(X.x[0].err, X.x[0].err, 1, ncclUint32, ncclMax, ds->comm, st));
    MSSPE_NCCL_TRY(c, ds, N->AllReduce(X.x[1].err, X.x[1].err, 1, ncclUint32, ncclMax, ds->comm, st));
    for (int d = 0; d < 2; d++) MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&h_err[d], X.x[d].err, 4, cudaMemcpyDeviceToHost, st));
    MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
    if (h_err[0] | h_err[1]) { c->set_error("msspe_select_both_dist: limit of the multi-GPU loop reached on some rank (flags %u/%u)", h_err[0], h_err[1]); return MSSPE_ERR_CAPACITY; }
  }
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[3], st));
  msspe_candidate* outs[2] = {out_fwd, out_rev};
  uint32_t* ns[2] = {n_fwd, n_rev};
  for (int d = 0; d < 2; d++) {
    const uint32_t n = h[d].n_out;
    if (n) MSSPE_CUDA_TRY(c, cudaMemcpyAsync(outs[d], c->dir[d].out, (size_t)n * sizeof(msspe_candidate), cudaMemcpyDeviceToHost, st));
    *ns[d] = n;
    c->timing.select_evals[d] = h[d].evals;
    c->kprof[KP_GREEDY_UNIT].bytes += h[d].work_bytes;
    c->timing.select_iterations[d] = h[d].iterations;
    c->timing.select_postings_read[d] = c->dir[d].n_records;
  }
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  float ms = 0.f;
  MSSPE_CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]));
  c->timing.select_ms[0] = c->timing.select_ms[1] = ms;
  if (dbg && rank == 0)
    fprintf(stderr, "[msspe] rank 0 set-up phases (both directions, ms): sizes %.3f | words all-gather + cross flags %.3f | partition view %.3f | cross words exchange + sort %.3f | "
                    "allocations %.3f | cross lengths all-reduce %.3f | init kernels %.3f\n", t_ph[0], t_ph[1], t_ph[2], t_ph[3], t_ph[4], t_ph[5], t_ph[6]);
  if (getenv("MSSPE_DEBUG_TIMERS"))
    for (int d = 0; d < 2; d++)
      fprintf(stderr, "[msspe] rank %d/%d partitioned greedy dir %d: %u winners, %u rounds, %u external winners, %u local multi lists, %u cross-rank lists, %.3f ms (set-up %.3f ms, wall %.3f ms; kmax %u, window %u, staged cap %u, buffer %u KB)\n",
              rank, world, d, h[d].n_out, h[d].rounds, h[d].rollbacks, A.d[d].n_multi, X.x[d].n_x, ms, t_setup, now_ms() - t_begin, X.kmax, X.wmax, X.xcap, 2 * X.xb_words / 256);
#undef DA
  return MSSPE_OK;
}
}  // namespace
