// select_part.cu -- MSSPE_SELECT_PARTITIONED: the greedy loop of find_candidates_kmers (od-msspe/src/main.rs:331-406)
// decomposed by partition.  No grid barrier, no per-iteration recount of the whole index.
//
// Why it is exact.  A k-mer whose postings all lie in ONE partition p (Segment.partition_no, main.rs:227) -- the rule in
// a pre-aligned alignment, where a word sits in the same columns of every genome -- only ever changes the live counts
// of k-mers of partition-p segments when it is chosen, and its partition_tie_score (main.rs:261-283) is
// 0.0 + 1/(partition_coverage[p] + 1) whichever of its postings are live.  So each partition ("unit") runs its OWN
// greedy sequence, key (frequency, smaller word), independently of every other one, and the loop's global selection
// order is the merge of those sequences by (frequency desc, partition_coverage asc, word asc) -- partition_coverage[p]
// being the number of winners p has already supplied -- cut by the stop rules of main.rs:353-390.  Lists that span
// several partitions couple units.  They are never candidates inside a unit; after every merge each of them is checked
// against the merged order: its live count at iteration t follows from the cover times of its postings, its score is
// the reference's sequential f32 sum over first-seen live partitions.  The earliest iteration t* at which one of them
// beats the merged winner is an EXTERNAL winner: everything before t* is final, the units it touches are rolled back
// to their state at t* and re-run, the rest is re-merged.  tests/_partitioned_model.py is this algorithm in Python,
// checked on the CPU against the restated reference loop (tests/test_partitioned_model.py); the GPU tests compare this
// file with reference-loop goldens at BASELINE sizes.
//
// Rounds of seven launches, enqueued in batches without host round trips (every kernel returns at once when its
// direction is done):
//   part_extend_kernel   one thread-block cluster per unit: roll-back of what changed since the external winner, then up to
//                        `nsteps` greedy steps on exact incremental counts (the unit's forward index); cover token per
//                        segment = the entry that covered it; partial arg-max / counts exchanged through DSMEM
//   part_gather_kernel   list of not-yet-final entries
//   part_merge_kernel    one warp per entry: its global position = entries of all units with a better key (binary
//                        search per unit); also the tie count and the live records (coverage evals) of that iteration
//   part_plan_kernel     horizon (first position an unfinished unit could still change), stop rules, window to verify
//   part_stage_kernel    multi-partition lists whose upper bound reaches the smallest winning frequency of the window
//   part_verify_kernel   one block per staged list: cover-time histogram -> live count per iteration, exact tie scores
//   part_finalize_kernel winners before the horizon / before t* become final; external winner applied, units flagged
#include "select_part.cuh"

int msspe_partition_view(msspe_ctx* c, int dir, cudaStream_t st) {
  DirIndex& D = c->dir[dir];
  if (D.pv_built && D.pv_dist) {   // a view built for the multi-GPU loop: rebuild from the index's own list_part
    msspe_dev_free(c, D.pv_ucode_off); msspe_dev_free(c, D.pv_ucodes); msspe_dev_free(c, D.pv_fwdl); msspe_dev_free(c, D.pv_useg_off); msspe_dev_free(c, D.pv_usegs);
    D.pv_ucode_off = D.pv_ucodes = D.pv_fwdl = D.pv_useg_off = D.pv_usegs = nullptr;
    D.pv_built = false; D.pv_dist = false;
  }
  if (D.pv_built) return MSSPE_OK;
  const uint32_t U = c->n_segments ? c->max_partition + 1u : 0u;
  const uint32_t nc = (uint32_t)D.n_codes;
  const uint64_t G = c->n_segments, GS = G * c->slots;
  const uint32_t uni = (c->uniform_parts && c->uniform_parts <= 65536u) ? c->uniform_parts : 0u;
  D.pv_units = U;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.pv_ucode_off, ((uint64_t)U + 2) * 4, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.pv_useg_off, ((uint64_t)U + 2) * 4, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.pv_ucodes, ((uint64_t)nc + 1) * 4, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.pv_usegs, (G + 1) * 4, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&D.pv_fwdl, (GS + 1) * 4, c->stream));
  uint64_t *ka = nullptr, *kb = nullptr; uint32_t *va = nullptr, *vb = nullptr; uint32_t* lid = nullptr; unsigned long long* d_tot = nullptr;
  const uint64_t nmax = std::max<uint64_t>(std::max<uint64_t>(nc, G), 1);
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&ka, nmax * 8, st)); MSSPE_CUDA_TRY(c, cudaMallocAsync(&kb, nmax * 8, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&va, nmax * 4, st)); MSSPE_CUDA_TRY(c, cudaMallocAsync(&vb, nmax * 4, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&lid, ((uint64_t)nc + 1) * 4, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&d_tot, 8, st));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(d_tot, 0, 8, st));
  uint32_t h_off[2] = {0, 0};
  if (nc) {  // codes by unit (stable: ascending code id = ascending word inside a unit), multi-partition lists last
    pv_key_kernel<<<(nc + 255u) / 256u, 256, 0, st>>>(D.list_part, nc, U, ka, va);
    c->timing.kernel_launches++;
    int rc = msspe_radix_sort_pairs(c, &ka, &va, &kb, &vb, nc, bits_for(U), st);
    if (rc) return rc;
    MSSPE_CUDA_TRY(c, cudaMemcpyAsync(D.pv_ucodes, va, (uint64_t)nc * 4, cudaMemcpyDeviceToDevice, st));
  }
  pv_bounds_kernel<<<(U + 1 + 255u) / 256u, 256, 0, st>>>(ka, nc, U, D.pv_ucode_off);
  c->timing.kernel_launches++;
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&h_off[0], D.pv_ucode_off + U, 4, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  D.pv_single = h_off[0]; D.pv_multi_n = nc - h_off[0];
  if (nc) {
    pv_lid_kernel<<<(nc + 255u) / 256u, 256, 0, st>>>(D.pv_ucodes, nc, D.pv_single, lid);
    c->timing.kernel_launches++;
  }
  if (GS) {
    pv_fwdl_kernel<<<(unsigned)div_up_u64(GS, 256), 256, 0, st>>>(D.fwd_ids, GS, lid, D.pv_fwdl);
    c->timing.kernel_launches++;
  }
  if (D.pv_multi_n) {
    pv_mlen_kernel<<<(D.pv_multi_n + 255u) / 256u, 256, 0, st>>>(D.pv_ucodes, D.pv_single, D.pv_multi_n, D.post_off, nullptr, d_tot);
    c->timing.kernel_launches++;
  }
  if (G) {   // segments by partition (stable: ascending segment index inside a partition)
    pv_segkey_kernel<<<(unsigned)div_up_u64(G, 256), 256, 0, st>>>(c->d_seg_part, uni, G, ka, va);
    c->timing.kernel_launches++;
    int rc = msspe_radix_sort_pairs(c, &ka, &va, &kb, &vb, G, bits_for(U), st);
    if (rc) return rc;
    MSSPE_CUDA_TRY(c, cudaMemcpyAsync(D.pv_usegs, va, G * 4, cudaMemcpyDeviceToDevice, st));
  }
  pv_bounds_kernel<<<(U + 1 + 255u) / 256u, 256, 0, st>>>(ka, (uint32_t)G, U, D.pv_useg_off);
  c->timing.kernel_launches++;
  unsigned long long tot = 0;
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&tot, d_tot, 8, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  D.pv_multi_postings = tot;
  MSSPE_CUDA_TRY(c, cudaFreeAsync(ka, st)); MSSPE_CUDA_TRY(c, cudaFreeAsync(kb, st));
  MSSPE_CUDA_TRY(c, cudaFreeAsync(va, st)); MSSPE_CUDA_TRY(c, cudaFreeAsync(vb, st));
  MSSPE_CUDA_TRY(c, cudaFreeAsync(lid, st)); MSSPE_CUDA_TRY(c, cudaFreeAsync(d_tot, st));
  D.pv_built = true;
  return MSSPE_OK;
}

// AUTO's test: the decomposition pays when (almost) every list lies in one partition; entries are U x max_iterations.
bool msspe_partitioned_applicable(msspe_ctx* c, int ndirs, const int* dirs, uint32_t max_iter) {
  if (!c->n_segments || !max_iter) return false;
  const uint64_t U = (uint64_t)c->max_partition + 1u;
  if (U > 4096u || U * max_iter > (1ull << 24)) return false;
  for (int i = 0; i < ndirs; i++) {
    if (msspe_partition_view(c, dirs[i], c->stream) != MSSPE_OK) return false;
    const DirIndex& D = c->dir[dirs[i]];
    if (D.pv_multi_postings * 8u > D.n_records) return false;   // more than 1/8 of the postings in multi-partition lists
  }
  return true;
}

int msspe_select_partitioned(msspe_ctx* c, int ndirs, const int* dirs, uint32_t max_iter, uint32_t mms, msspe_candidate** outs,
                             uint32_t** n_outs) {
  cudaStream_t st = c->stream;
  const uint64_t G = c->n_segments;
  const uint32_t U = G ? c->max_partition + 1u : 0u;
  for (int i = 0; i < ndirs; i++) { *n_outs[i] = 0; c->timing.select_evals[dirs[i]] = 0; c->timing.select_iterations[dirs[i]] = 0; c->timing.select_ms[dirs[i]] = 0.f;
                                    c->timing.select_postings_read[dirs[i]] = 0; c->timing.count_kernel_launches[dirs[i]] = 0; c->timing.count_kernel_ms[dirs[i]] = 0.f; }
  if (max_iter == 0) return MSSPE_OK;
  if (U == 0) { for (int i = 0; i < ndirs; i++) c->timing.select_iterations[dirs[i]] = 1; return MSSPE_OK; }
  const uint64_t CAP = max_iter;
  if (G >= 0x80000000ull) { c->set_error("msspe_select: the partitioned loop holds at most 2^31 segments"); return MSSPE_ERR_CAPACITY; }
  if ((uint64_t)U * CAP + 2ull * max_iter >= 0x7FFFFFF0ull) { c->set_error("msspe_select: %u partitions x %u iterations exceed the entry table of the partitioned loop", U, max_iter); return MSSPE_ERR_CAPACITY; }
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[2], st));
  PartArgs A{};
  A.ndirs = ndirs; A.U = U; A.CAP = (uint32_t)CAP; A.slots = c->slots; A.max_iter = max_iter; A.mms = mms;
  A.uniform_parts = (c->uniform_parts && c->uniform_parts <= 65536u) ? c->uniform_parts : 0u;
  A.seg_part = c->d_seg_part; A.max_ahead = 0xFFFFFFFFu;
  // the loop's scratch of all directions comes as THREE slabs (zero-filled, 0xFF-filled, unfilled): 3 allocations and 2
  // memsets per call instead of ~45 of each per direction (0.25 ms of host time in front of an idle GPU at cfg3)
  std::vector<void*> scratch;
  struct Carve { void** p; uint64_t off; int kind; };
  std::vector<Carve> carve;
  uint64_t slab_bytes[3] = {0, 0, 0};
  auto alloc = [&](void** p, uint64_t bytes, int fill) -> int {
    const int kind = fill == 0 ? 0 : (fill == 0xFF ? 1 : 2);
    carve.push_back({p, slab_bytes[kind], kind});
    slab_bytes[kind] += ((bytes ? bytes : 4) + 255u) & ~(uint64_t)255u;
    return MSSPE_OK;
  };
  auto commit = [&]() -> int {
    unsigned char* slab[3] = {nullptr, nullptr, nullptr};
    for (int kd = 0; kd < 3; kd++) {
      if (!slab_bytes[kd]) continue;
      MSSPE_CUDA_TRY(c, cudaMallocAsync((void**)&slab[kd], slab_bytes[kd], st));
      scratch.push_back(slab[kd]);
      if (kd < 2) MSSPE_CUDA_TRY(c, cudaMemsetAsync(slab[kd], kd == 0 ? 0 : 0xFF, slab_bytes[kd], st));
    }
    for (const Carve& cv : carve) *cv.p = slab[cv.kind] + cv.off;
    return MSSPE_OK;
  };
  uint32_t max_multi = 0;
  for (int i = 0; i < ndirs; i++) {
    int rc = msspe_partition_view(c, dirs[i], st);
    if (rc) return rc;
    DirIndex& X = c->dir[dirs[i]];
    if (X.out_capacity < max_iter) {
      msspe_dev_free(c, X.out); X.out = nullptr;
      MSSPE_CUDA_TRY(c, cudaMallocAsync(&X.out, (uint64_t)max_iter * sizeof(msspe_candidate), c->stream));
      X.out_capacity = max_iter;
    }
    PartDir& P = A.d[i];
    P.codes = X.codes; P.post_off = X.post_off; P.postings = X.postings; P.ucode_off = X.pv_ucode_off; P.ucodes = X.pv_ucodes;
    P.fwdl = X.pv_fwdl; P.useg_off = X.pv_useg_off; P.usegs = X.pv_usegs; P.n_single = X.pv_single; P.n_multi = X.pv_multi_n;
    P.out = X.out;
    max_multi = std::max(max_multi, P.n_multi);
    const uint64_t ne = (uint64_t)U * CAP;
#define PV_ALLOC(field, bytes, fill) { int rc2 = alloc((void**)&P.field, (bytes), (fill)); if (rc2) return rc2; }
    PV_ALLOC(pfreq, ((uint64_t)X.n_codes + 1) * 4, -1);
    PV_ALLOC(token, (G + 1) * 4, 0xFF);
    PV_ALLOC(ulive, (uint64_t)U * 8, 0);
    PV_ALLOC(entries, ne * sizeof(PEntry), -1);
    PV_ALLOC(pos, (ne + 2ull * max_iter + 2) * 4, 0xFF);
    PV_ALLOC(rfin, (uint64_t)U * 4, 0);
    PV_ALLOC(ulen, (uint64_t)U * 4, 0);
    PV_ALLOC(status, (uint64_t)U * 4, -1);
    PV_ALLOC(ext_cov, (uint64_t)U * 4, 0);
    PV_ALLOC(elist, ne * 4, -1);
    PV_ALLOC(order, ((uint64_t)max_iter + 2) * 4, 0);
    PV_ALLOC(tied, ((uint64_t)max_iter + 2) * 4, 0);
    PV_ALLOC(tot_live, ((uint64_t)max_iter + 2) * 8, 0);
    PV_ALLOC(mt, ((uint64_t)max_iter + 2) * 4, 0);
    PV_ALLOC(ub, ((uint64_t)P.n_multi + 1) * 4, 0);
    PV_ALLOC(stage, ((uint64_t)P.n_multi + 1) * 4, -1);
    PV_ALLOC(viol, ((uint64_t)P.n_multi + 1) * 16, -1);
    PV_ALLOC(touch, (uint64_t)U * 4, 0);
    PV_ALLOC(win_freq, ((uint64_t)max_iter + 2) * 4, 0);
    PV_ALLOC(win_cov, ((uint64_t)max_iter + 2) * 4, 0);
    PV_ALLOC(win_code, ((uint64_t)max_iter + 2) * 8, 0);
    PV_ALLOC(ctl, sizeof(PartCtl), 0);
#undef PV_ALLOC
  }
  { int rc2 = commit(); if (rc2) return rc2; }
  for (int i = 0; i < ndirs; i++) {
    PartDir& P = A.d[i];
    pv_status_kernel<<<(U + 255u) / 256u, 256, 0, st>>>(P.status, U);
    c->timing.kernel_launches++;
    if (P.n_multi) {
      pv_mlen_kernel<<<(P.n_multi + 255u) / 256u, 256, 0, st>>>(P.ucodes, P.n_single, P.n_multi, P.post_off, P.ub, nullptr);
      c->timing.kernel_launches++;
    }
  }
  const size_t ver_smem = (2 * ((size_t)max_iter + 2) + (U + 31u) / 32u) * 4;
  if (ver_smem > c->smem_optin) { c->set_error("msspe_select: max_iterations %u too large for the verify kernel's histogram", max_iter); return MSSPE_ERR_CAPACITY; }
  MSSPE_CUDA_TRY(c, cudaFuncSetAttribute(part_verify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ver_smem));
  uint32_t chunk0 = (uint32_t)std::max<uint64_t>(4, (3ull * max_iter + 2ull * U - 1) / (2ull * U));
  uint32_t chunk = 8;
  if (const char* e = getenv("MSSPE_PART_CHUNK0")) chunk0 = (uint32_t)std::max(1, atoi(e));
  if (const char* e = getenv("MSSPE_PART_CHUNK")) chunk = (uint32_t)std::max(1, atoi(e));
  const unsigned merge_grid = (unsigned)c->sm_count * 4u, ver_grid = (unsigned)c->sm_count * 2u;
  // CTAs per unit (one thread-block cluster): by the forward-index entries of a unit, i.e. the work of covering all of it
  const uint64_t unit_items = G / U * c->slots;
  int csize = unit_items >= (1u << 20) ? 8 : unit_items >= (1u << 18) ? 4 : unit_items >= (1u << 16) ? 2 : 1;
  if (const char* e = getenv("MSSPE_PART_CLUSTER")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) csize = v; }
  {
    unsigned mx = 1;
    for (int i = 0; i < ndirs; i++) mx = std::max<unsigned>(mx, A.d[i].n_single);
    part_init_freq_kernel<<<dim3((mx + 255u) / 256u, ndirs), 256, 0, st>>>(A);
    part_init_live_kernel<<<dim3((unsigned)div_up_u64(G, 256), ndirs), 256, 0, st>>>(A, G);
    c->timing.kernel_launches += 2;
  }
  PartCtl* h = reinterpret_cast<PartCtl*>(c->h_ctl);   // pinned staging: 2 x SelectCtl is larger than 2 x PartCtl
  static_assert(2 * sizeof(PartCtl) <= 2 * sizeof(SelectCtl), "pinned staging too small");
  uint32_t round = 0;
  const uint32_t BATCH = 4;
  for (;;) {
    for (uint32_t b = 0; b < BATCH; b++, round++) {
      A.nsteps = round == 0 ? chunk0 : chunk;
      {
        KPROF(c, KP_GREEDY_UNIT, st, 0)
        int rc2 = launch_extend(c, A, csize, st);
        if (rc2) return rc2;
      }
      { KPROF(c, KP_GREEDY_MERGE, st, 0) part_gather_kernel<<<ndirs, 1024, 0, st>>>(A); }
      { KPROF(c, KP_GREEDY_MERGE, st, 0) part_merge_kernel<<<dim3(merge_grid, ndirs), 256, 0, st>>>(A); }
      { KPROF(c, KP_GREEDY_MERGE, st, 0) part_plan_kernel<<<ndirs, 1024, 0, st>>>(A); }
      if (max_multi) {
        { KPROF(c, KP_GREEDY_MERGE, st, 0) part_stage_kernel<<<dim3((max_multi + 255u) / 256u, ndirs), 256, 0, st>>>(A); }
        { KPROF(c, KP_GREEDY_VERIFY, st, 0) part_verify_kernel<<<dim3(ver_grid, ndirs), VER_T, ver_smem, st>>>(A); }
      }
      { KPROF(c, KP_GREEDY_MERGE, st, 0) part_finalize_kernel<<<ndirs, 1024, 0, st>>>(A); }
    }
    MSSPE_CUDA_TRY(c, cudaGetLastError());
    for (int i = 0; i < ndirs; i++) MSSPE_CUDA_TRY(c, cudaMemcpyAsync(&h[i], A.d[i].ctl, sizeof(PartCtl), cudaMemcpyDeviceToHost, st));
    MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
    bool all = true;
    for (int i = 0; i < ndirs; i++) all = all && h[i].done;
    if (all) break;
    if (round > 4u * max_iter + 64u) { c->set_error("msspe_select: partitioned loop did not converge"); return MSSPE_ERR_STATE; }
  }
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[3], st));
  for (int i = 0; i < ndirs; i++) {
    const int d = dirs[i];
    const uint32_t n = h[i].n_out;
    if (n) MSSPE_CUDA_TRY(c, cudaMemcpyAsync(outs[i], c->dir[d].out, (size_t)n * sizeof(msspe_candidate), cudaMemcpyDeviceToHost, st));
    *n_outs[i] = n;
    c->timing.select_evals[d] = h[i].evals;
    c->kprof[KP_GREEDY_UNIT].bytes += h[i].work_bytes;
    c->timing.select_iterations[d] = h[i].iterations;
    c->timing.select_postings_read[d] = c->dir[d].n_records;   // the forward index is read about once (counts) + once (decrements)
  }
  for (void* p : scratch) MSSPE_CUDA_TRY(c, cudaFreeAsync(p, st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  float ms = 0.f;
  MSSPE_CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]));
  for (int i = 0; i < ndirs; i++) c->timing.select_ms[dirs[i]] = ms;
  if (getenv("MSSPE_DEBUG_TIMERS"))
    for (int i = 0; i < ndirs; i++)
      fprintf(stderr, "[msspe] partitioned greedy dir %d: %u winners, %u iterations, %u rounds, %u external winners, %u multi-partition lists (%llu postings), %.3f ms\n",
              dirs[i], h[i].n_out, h[i].iterations, h[i].rounds, h[i].rollbacks, A.d[i].n_multi, (unsigned long long)c->dir[dirs[i]].pv_multi_postings, ms);
  return MSSPE_OK;
}
