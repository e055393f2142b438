// ctx.cu -- context lifecycle, genome upload, timing.  Part of libodmsspe_b200.so (C ABI in
// include/od_msspe_b200.h).  No CPU fallback: every entry point needs a working CUDA device.
#include "engine.cuh"

static thread_local std::string g_create_error;
static void kprof_resolve(msspe_ctx* c);

extern "C" int msspe_abi_version(void) { return MSSPE_ABI_VERSION; }

extern "C" const char* msspe_last_error(const msspe_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

static int fail_create(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_create_error = buf;
  return 0;
}

extern "C" int msspe_create(const msspe_config* cfg, msspe_ctx** out) {
  if (!cfg || !out) { fail_create("msspe_create: null argument"); return MSSPE_ERR_INVALID; }
  *out = nullptr;
  // Argument checks that the reference turns into panics: main.rs:201-203 (overlap < search window),
  // step_by(0) at main.rs:178, windows(0) at main.rs:177.
  if (cfg->kmer_size < 1 || cfg->kmer_size > MSSPE_MAX_KMER) {
    fail_create("kmer_size %u outside 1..%d", cfg->kmer_size, MSSPE_MAX_KMER); return MSSPE_ERR_INVALID;
  }
  if (cfg->overlap_size < cfg->search_windows_size) {
    fail_create("Overlap windows size must be greater or equal than search windows size"); return MSSPE_ERR_INVALID;
  }
  if (cfg->overlap_size == 0 || cfg->window_size == 0) {
    fail_create("window_size and overlap_size must be > 0"); return MSSPE_ERR_INVALID;
  }
  if (cfg->search_windows_size > cfg->window_size) {
    fail_create("search_windows_size %u exceeds window_size %u", cfg->search_windows_size, cfg->window_size);
    return MSSPE_ERR_INVALID;
  }
  if (cfg->search_windows_size > 1024) {
    fail_create("search_windows_size %u > 1024 is not supported", cfg->search_windows_size); return MSSPE_ERR_INVALID;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    fail_create("no CUDA device available (%s); this engine has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return MSSPE_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { fail_create("device %d out of range (0..%d)", cfg->device, ndev - 1); return MSSPE_ERR_INVALID; }
  e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) { fail_create("cudaSetDevice(%d): %s", cfg->device, cudaGetErrorString(e)); return MSSPE_ERR_CUDA; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, cfg->device);
  if (e != cudaSuccess) { fail_create("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return MSSPE_ERR_CUDA; }
  if (prop.major < 10) {
    fail_create("device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device, prop.major, prop.minor);
    return MSSPE_ERR_CUDA;
  }
  msspe_ctx* c = new (std::nothrow) msspe_ctx();
  if (!c) { fail_create("out of host memory"); return MSSPE_ERR_NOMEM; }
  c->cfg = *cfg;
  c->device = cfg->device;
  {  // keep freed blocks in the stream-ordered pool instead of returning them to the driver at every sync
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, cfg->device) == cudaSuccess) {
      uint64_t thr = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  c->slots = cfg->search_windows_size >= cfg->kmer_size ? cfg->search_windows_size - cfg->kmer_size + 1 : 0;
  bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) == cudaSuccess &&
            cudaMallocHost(&c->h_ctl, 2 * sizeof(SelectCtl)) == cudaSuccess;
  for (int i = 0; ok && i < 8; i++) ok = cudaEventCreate(&c->ev[i]) == cudaSuccess;
  if (!ok) {
    fail_create("CUDA resource creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    msspe_destroy(c);
    return MSSPE_ERR_CUDA;
  }
  *out = c;
  return MSSPE_OK;
}

static void free_genomes(msspe_ctx* c) {
  if (c->d_bases && !c->bases_borrowed) msspe_dev_free(c, c->d_bases);
  c->d_bases = nullptr; c->bases_borrowed = false;
  if (c->d_offsets) msspe_dev_free(c, c->d_offsets);
  if (c->d_seg_base) msspe_dev_free(c, c->d_seg_base);
  c->d_offsets = c->d_seg_base = nullptr;
  c->loaded = false;
}

extern "C" void msspe_destroy(msspe_ctx* c) {
  msspe_join_reserve(c);
  if (c && c->h_stage) { cudaFreeHost(c->h_stage); c->h_stage = nullptr; c->h_stage_bytes = 0; }
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->stream2) cudaStreamSynchronize(c->stream2);
  msspe_free_index(c);
  free_genomes(c);
  msspe_thal_free_tables(c);
  msspe_dist_free(c);
  if (c->xd_edges) cudaFreeAsync(c->xd_edges, c->stream);
  if (c->xd_nostruct) cudaFreeAsync(c->xd_nostruct, c->stream);
  kprof_resolve(c);
  for (cudaEvent_t e : c->kprof_pool) cudaEventDestroy(e);
  for (int i = 0; i < 8; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->h_ctl) cudaFreeHost(c->h_ctl);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  delete c;
}

void msspe_join_reserve(msspe_ctx* c) {
  if (c && c->reserve_thread) { c->reserve_thread->join(); delete c->reserve_thread; c->reserve_thread = nullptr; }
}

extern "C" int msspe_reserve_pool(msspe_ctx* c, uint64_t bytes) {
  if (!c) return MSSPE_ERR_INVALID;
  msspe_join_reserve(c);
  if (bytes == 0) return MSSPE_OK;
  const int device = c->device;
  c->reserve_thread = new std::thread([device, bytes]() {
    if (cudaSetDevice(device) != cudaSuccess) return;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return;
    const uint64_t want = bytes < (uint64_t)(free_b / 5 * 4) ? bytes : (uint64_t)(free_b / 5 * 4);
    cudaStream_t s = nullptr;
    if (want == 0 || cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return;
    void* p = nullptr;
    // the pool keeps freed blocks (release threshold = max): allocate once, hand it back, the mapping stays
    if (cudaMallocAsync(&p, want, s) == cudaSuccess) cudaFreeAsync(p, s);
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    (void)cudaGetLastError();
  });
  return MSSPE_OK;
}

extern "C" int msspe_set_stream(msspe_ctx* c, void* cuda_stream) {
  if (!c) return MSSPE_ERR_INVALID;
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)cuda_stream;
  c->own_stream = false;
  return MSSPE_OK;
}

extern "C" int msspe_synchronize(msspe_ctx* c) {
  if (!c) return MSSPE_ERR_INVALID;
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(c->stream2));
  return MSSPE_OK;
}

extern "C" int msspe_get_timing(msspe_ctx* c, msspe_timing* out) {
  if (!c || !out) return MSSPE_ERR_INVALID;
  *out = c->timing;
  return MSSPE_OK;
}
static void kprof_resolve(msspe_ctx* c) {
  for (auto& p : c->kprof_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) c->kprof[p.cls].ms += ms;
    c->kprof_pool.push_back(p.e0); c->kprof_pool.push_back(p.e1);
  }
  c->kprof_pending.clear();
}
extern "C" int msspe_reset_timing(msspe_ctx* c) {
  if (!c) return MSSPE_ERR_INVALID;
  memset(&c->timing, 0, sizeof c->timing);
  kprof_resolve(c);
  for (int i = 0; i < KP_N; i++) c->kprof[i] = KProfClass();
  return MSSPE_OK;
}
cudaEvent_t msspe_kprof_begin(msspe_ctx* c, cudaStream_t st) {
  if (!c->profiling) return nullptr;
  cudaEvent_t e = nullptr;
  if (!c->kprof_pool.empty()) { e = c->kprof_pool.back(); c->kprof_pool.pop_back(); }
  else if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
  cudaEventRecord(e, st);
  return e;
}
void msspe_kprof_end(msspe_ctx* c, int cls, cudaEvent_t e0, cudaStream_t st, uint64_t alg_bytes) {
  c->kprof[cls].launches++; c->kprof[cls].bytes += alg_bytes;
  c->timing.kernel_launches++;
  if (!e0) return;
  cudaEvent_t e1 = nullptr;
  if (!c->kprof_pool.empty()) { e1 = c->kprof_pool.back(); c->kprof_pool.pop_back(); }
  else if (cudaEventCreate(&e1) != cudaSuccess) { c->kprof_pool.push_back(e0); return; }
  cudaEventRecord(e1, st);
  c->kprof_pending.push_back({cls, e0, e1});
}
extern "C" int msspe_get_kernel_profile(msspe_ctx* c, msspe_kernel_prof* out, uint32_t capacity, uint32_t* n) {
  if (!c || !n) return MSSPE_ERR_INVALID;
  static const char* names[KP_N] = {"encode_windows", "radix_hist", "radix_scatter", "scan", "build_csr", "partition_view", "greedy_unit (part_extend)",
                                    "greedy_merge (gather/merge/plan/stage/finalize)", "greedy_verify", "greedy_whole_index", "primer_thermo", "thal_dimer"};
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  kprof_resolve(c);
  uint32_t k = 0;
  for (int i = 0; i < KP_N; i++) {
    if (!c->kprof[i].launches) continue;
    if (out && k < capacity) {
      memset(&out[k], 0, sizeof out[k]);
      snprintf(out[k].name, sizeof out[k].name, "%s", names[i]);
      out[k].ms = (float)c->kprof[i].ms; out[k].launches = c->kprof[i].launches; out[k].alg_bytes = c->kprof[i].bytes;
    }
    k++;
  }
  *n = k;
  return (out && k > capacity) ? MSSPE_ERR_CAPACITY : MSSPE_OK;
}
extern "C" int msspe_set_profiling(msspe_ctx* c, int on) {
  if (!c) return MSSPE_ERR_INVALID;
  c->profiling = on != 0;
  return MSSPE_OK;
}

// Segment geometry of partitioning_sequence (main.rs:173-181): full windows only, starts 0,S,2S,...
static int plan_segments(msspe_ctx* c, const uint64_t* offsets, uint32_t n) {
  const uint64_t W = c->cfg.window_size, S = c->cfg.overlap_size;
  if (offsets[0] != 0) { c->set_error("offsets[0] must be 0"); return MSSPE_ERR_INVALID; }
  c->h_offsets.assign(offsets, offsets + n + 1);
  c->h_seg_base.assign(n + 1, 0);
  uint64_t g = 0, maxp = 0, minp = UINT64_MAX;
  for (uint32_t r = 0; r < n; r++) {
    if (offsets[r + 1] < offsets[r]) { c->set_error("offsets not monotone at record %u", r); return MSSPE_ERR_INVALID; }
    uint64_t L = offsets[r + 1] - offsets[r];
    uint64_t P = L >= W ? (L - W) / S + 1 : 0;
    if (P > 65536) {  // Segment.partition_no is `j as u16` (main.rs:227): beyond 65,536 windows the reference wraps silently
      c->set_error("record %u has %llu partitions; partition_no is u16 in the reference (main.rs:84,227) and would wrap: use a larger --overlap-size", r, (unsigned long long)P);
      return MSSPE_ERR_INVALID;
    }
    c->h_seg_base[r] = g;
    g += P;
    if (P > maxp) maxp = P;
    if (P < minp) minp = P;
  }
  c->h_seg_base[n] = g;
  if (g >= 0xFFFFFFFFull) { c->set_error("%llu segments exceed the u32 segment index of the reference (main.rs:250)", (unsigned long long)g); return MSSPE_ERR_CAPACITY; }
  if (g * (uint64_t)(c->slots ? c->slots : 1) >= 0xFFFFFFFFull) {
    c->set_error("%llu segments x %u slots exceed one GPU shard (record index is 32-bit); shard the genomes", (unsigned long long)g, c->slots);
    return MSSPE_ERR_CAPACITY;
  }
  c->n_segments = g;
  c->uniform_parts = (n > 0 && minp == maxp && maxp > 0 && maxp < 0xFFFFFFFFull) ? (uint32_t)maxp : 0u;
  // Segment.partition_no is `j as u16` (main.rs:227): the maximum of the wrapped values
  c->max_partition = maxp == 0 ? 0 : (maxp > 65536 ? 65535u : (uint32_t)(maxp - 1));
  c->n_records = n;
  return MSSPE_OK;
}

static int upload_plan(msspe_ctx* c) {
  uint32_t n = c->n_records;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&c->d_offsets, (n + 1) * sizeof(uint64_t), c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&c->d_seg_base, (n + 1) * sizeof(uint64_t), c->stream));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(c->d_offsets, c->h_offsets.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(c->d_seg_base, c->h_seg_base.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
  return MSSPE_OK;
}

// Plan the segments of n records and allocate their device buffer (fasta.cu streams the bases in chunk by chunk).
int msspe_load_begin(msspe_ctx* c, const uint64_t* offsets, uint32_t n) {
  if (n == 0) { c->set_error("No sequences found in the input file"); return MSSPE_ERR_INVALID; }  // main.rs:652-654
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  msspe_free_index(c);
  free_genomes(c);
  int rc = plan_segments(c, offsets, n);
  if (rc) return rc;
  c->bases_bytes = offsets[n];
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&c->d_bases, c->bases_bytes ? c->bases_bytes : 1, c->stream));
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[0], c->stream));
  return MSSPE_OK;
}
int msspe_load_finish(msspe_ctx* c) {
  int rc = upload_plan(c);
  if (rc) return rc;
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[1], c->stream));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  MSSPE_CUDA_TRY(c, cudaEventElapsedTime(&c->timing.h2d_ms, c->ev[0], c->ev[1]));
  c->loaded = true;
  return MSSPE_OK;
}

extern "C" int msspe_load_genomes(msspe_ctx* c, const uint8_t* bases, const uint64_t* offsets, uint32_t n) {
  if (!c) return MSSPE_ERR_INVALID;
  if (!offsets || (n && !bases && offsets[n] > 0)) { c->set_error("msspe_load_genomes: null argument"); return MSSPE_ERR_INVALID; }
  int rc = msspe_load_begin(c, offsets, n);
  if (rc) return rc;
  if (c->bases_bytes)
    MSSPE_CUDA_TRY(c, cudaMemcpyAsync(c->d_bases, bases, c->bases_bytes, cudaMemcpyHostToDevice, c->stream));
  return msspe_load_finish(c);
}

extern "C" int msspe_load_genomes_device(msspe_ctx* c, const uint8_t* d_bases, const uint64_t* offsets, uint32_t n) {
  if (!c) return MSSPE_ERR_INVALID;
  if (!offsets || !d_bases) { c->set_error("msspe_load_genomes_device: null argument"); return MSSPE_ERR_INVALID; }
  if (n == 0) { c->set_error("No sequences found in the input file"); return MSSPE_ERR_INVALID; }
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  msspe_free_index(c);
  free_genomes(c);
  int rc = plan_segments(c, offsets, n);
  if (rc) return rc;
  c->bases_bytes = offsets[n];
  c->d_bases = const_cast<uint8_t*>(d_bases);
  c->bases_borrowed = true;
  rc = upload_plan(c);
  if (rc) return rc;
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->timing.h2d_ms = 0.f;
  c->loaded = true;
  return MSSPE_OK;
}

extern "C" int msspe_segment_info(msspe_ctx* c, uint64_t* n_segments, uint32_t* max_partition, uint32_t* slots) {
  if (!c) return MSSPE_ERR_INVALID;
  if (!c->loaded) { c->set_error("msspe_segment_info: no genomes loaded"); return MSSPE_ERR_STATE; }
  if (n_segments) *n_segments = c->n_segments;
  if (max_partition) *max_partition = c->max_partition;
  if (slots) *slots = c->slots;
  return MSSPE_OK;
}
