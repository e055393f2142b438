// fasta.cu -- SURVEY section 8 row f-3: FASTA ingest.  Replaces to_records (od-msspe/src/main.rs:108-122: seq_io
// reader, id = header up to the first space, all sequence lines joined, to_uppercase, "U" -> "T") with a
// multi-threaded parse straight into one pinned buffer and, for msspe_load_fasta, a chunked host-to-device copy
// that overlaps the parse: the copy of chunk c is issued while the worker threads normalise chunk c+1.
// Host code only (no kernels); at 100,000 x 30 kb (3 GB) the parse + copy is what bounds an end-to-end run.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <thread>

#include "engine.cuh"
#include <chrono>

struct msspe_fasta {
  std::vector<std::string> names;
  std::vector<uint64_t> offsets;  // [n+1]
  uint8_t* bases = nullptr;       // normalised sequences, concatenated
  uint64_t n_bytes = 0;
  bool pinned = false;
};

namespace {

struct Mapped {
  const uint8_t* p = nullptr; size_t n = 0; bool mapped = false; std::vector<uint8_t> own;
  ~Mapped() { if (mapped && p) munmap(const_cast<uint8_t*>(p), n); }
};

void set_err(char* err, size_t len, const char* msg) { if (err && len) { strncpy(err, msg, len - 1); err[len - 1] = 0; } }

template <class F>
void parallel_for(uint32_t threads, uint64_t n, uint64_t grain, F&& f) {  // f(begin, end), dynamic chunks
  if (n == 0) return;
  std::atomic<uint64_t> next{0};
  auto work = [&] { for (;;) { const uint64_t b = next.fetch_add(grain); if (b >= n) return; f(b, std::min(n, b + grain)); } };
  const uint32_t t = (uint32_t)std::min<uint64_t>(threads, (n + grain - 1) / grain);
  std::vector<std::thread> pool;
  for (uint32_t i = 1; i < t; i++) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
}

struct Lut { uint8_t t[256]; Lut() { for (int i = 0; i < 256; i++) { int c = (i >= 'a' && i <= 'z') ? i - 32 : i; t[i] = (uint8_t)(c == 'U' ? 'T' : c); } } };
const Lut kLut;

// sequence lines of one record, [pos, end): calls line(begin, end) with the '\r' before a '\n' removed
template <class F>
inline void for_lines(const uint8_t* s, uint64_t pos, uint64_t end, F&& line) {
  while (pos < end) {
    const uint8_t* nl = (const uint8_t*)memchr(s + pos, '\n', end - pos);
    const uint64_t e = nl ? (uint64_t)(nl - s) : end;
    uint64_t le = e;
    if (le > pos && s[le - 1] == '\r') le--;
    line(pos, le);
    pos = e + 1;
  }
}

// The parse.  on_sized(total bytes) is called once the offsets are known (the caller allocates there and returns the
// destination), on_chunk(byte_begin, byte_end) as soon as that range of the destination is final, in ascending order.
int parse_fasta(const char* path, uint32_t threads, msspe_fasta* F, const std::function<uint8_t*(uint64_t)>& on_sized,
                const std::function<void(uint64_t, uint64_t)>& on_chunk, char* err, size_t err_len) {
  if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
  Mapped M;
  {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { set_err(err, err_len, "No such file or directory (os error 2)"); return MSSPE_ERR_IO; }
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); set_err(err, err_len, "cannot stat the input"); return MSSPE_ERR_IO; }
    M.n = (size_t)st.st_size;
    if (M.n) {
      void* p = mmap(nullptr, M.n, PROT_READ, MAP_PRIVATE, fd, 0);
      if (p != MAP_FAILED) { M.p = (const uint8_t*)p; M.mapped = true; madvise(p, M.n, MADV_WILLNEED); }
      else {  // pipes and the like: read it
        M.own.resize(M.n);
        size_t got = 0;
        while (got < M.n) { const ssize_t r = read(fd, M.own.data() + got, M.n - got); if (r <= 0) break; got += (size_t)r; }
        if (got != M.n) { close(fd); set_err(err, err_len, "short read"); return MSSPE_ERR_IO; }
        M.p = M.own.data();
      }
    }
    close(fd);
  }
  const uint8_t* s = M.p;
  const uint64_t n = M.n;
  const bool dbg = getenv("MSSPE_DEBUG_TIMERS") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!dbg) return;
    const auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[msspe] fasta: %s %.3f s\n", what, std::chrono::duration<double>(t - t_last).count());
    t_last = t;
  };
  // 1. record starts: '>' at the start of a line
  std::vector<uint64_t> starts;
  {
    const uint64_t grain = std::max<uint64_t>(1 << 20, n / (threads * 8ull) + 1);
    const uint64_t n_chunks = (n + grain - 1) / grain;
    std::vector<std::vector<uint64_t>> part(n_chunks);
    parallel_for(threads, n, grain, [&](uint64_t b, uint64_t e) {
      auto& v = part[b / grain];
      for (uint64_t p = b; p < e;) {
        const uint8_t* q = (const uint8_t*)memchr(s + p, '>', e - p);
        if (!q) break;
        const uint64_t at = (uint64_t)(q - s);
        if (at == 0 || s[at - 1] == '\n') v.push_back(at);
        p = at + 1;
      }
    });
    for (auto& v : part) starts.insert(starts.end(), v.begin(), v.end());
  }
  // anything but empty lines before the first header is seq_io's InvalidStart
  {
    const uint64_t lim = starts.empty() ? n : starts[0];
    bool bad = false;
    for_lines(s, 0, lim, [&](uint64_t b, uint64_t e) { if (e > b) bad = true; });
    if (bad) { set_err(err, err_len, "called `Result::unwrap()` on an `Err` value: InvalidStart (FASTA must begin with '>')"); return MSSPE_ERR_INVALID; }
  }
  lap("record starts");
  const uint64_t nrec = starts.size();
  if (nrec >= 0xFFFFFFFFull) { set_err(err, err_len, "too many records"); return MSSPE_ERR_CAPACITY; }
  F->names.assign(nrec, std::string());
  F->offsets.assign(nrec + 1, 0);
  std::vector<uint64_t> body(nrec);  // first byte after the header line
  // 2. names and lengths
  parallel_for(threads, nrec, 64, [&](uint64_t b, uint64_t e) {
    for (uint64_t i = b; i < e; i++) {
      const uint64_t h = starts[i], lim = i + 1 < nrec ? starts[i + 1] : n;
      const uint8_t* nl = (const uint8_t*)memchr(s + h, '\n', lim - h);
      const uint64_t he = nl ? (uint64_t)(nl - s) : lim;
      uint64_t le = he;
      if (le > h && s[le - 1] == '\r') le--;
      uint64_t t = h + 1;
      while (t < le && s[t] != ' ') t++;
      F->names[i].assign((const char*)s + h + 1, t - h - 1);
      body[i] = std::min(lim, he + 1);
      uint64_t len = 0;
      for_lines(s, body[i], lim, [&](uint64_t lb, uint64_t le2) { len += le2 - lb; });
      F->offsets[i + 1] = len;
    }
  });
  for (uint64_t i = 0; i < nrec; i++) F->offsets[i + 1] += F->offsets[i];
  F->n_bytes = F->offsets[nrec];
  lap("names and lengths");
  uint8_t* dst = on_sized(F->n_bytes);
  lap("allocation");
  if (!dst && F->n_bytes) { set_err(err, err_len, "out of host memory"); return MSSPE_ERR_NOMEM; }
  // 3. normalise, in chunks of ~32 MB of output; a finished chunk is handed on while the next one is being written
  const uint64_t chunk_bytes = 32ull << 20;
  uint64_t r0 = 0;
  std::thread copier;  // runs on_chunk for the previous chunk
  while (r0 < nrec) {
    uint64_t r1 = r0 + 1;
    while (r1 < nrec && F->offsets[r1] - F->offsets[r0] < chunk_bytes) r1++;
    parallel_for(threads, r1 - r0, 16, [&](uint64_t b, uint64_t e) {
      for (uint64_t i = r0 + b; i < r0 + e; i++) {
        uint8_t* o = dst + F->offsets[i];
        const uint64_t lim = i + 1 < nrec ? starts[i + 1] : n;
        for_lines(s, body[i], lim, [&](uint64_t lb, uint64_t le2) { for (uint64_t q = lb; q < le2; q++) *o++ = kLut.t[s[q]]; });
      }
    });
    if (copier.joinable()) copier.join();
    if (on_chunk) { const uint64_t b0 = F->offsets[r0], b1 = F->offsets[r1]; copier = std::thread([&on_chunk, b0, b1] { on_chunk(b0, b1); }); }
    r0 = r1;
  }
  if (copier.joinable()) copier.join();
  lap("normalise + copy");
  return MSSPE_OK;
}

// Pinning costs about as much per byte as parsing does (page-locking 3 GB: ~1 s), so only small inputs get a pinned
// buffer; a large one is pageable and its chunks go to the device through the driver's own staging buffers while the
// worker threads normalise the next chunk.  MSSPE_FASTA_PINNED=0/1 forces either.
uint8_t* host_alloc(msspe_fasta* F, uint64_t bytes) {
  const uint64_t a = bytes ? bytes : 1;
  void* p = nullptr;
  bool want_pinned = a <= (256ull << 20);
  if (const char* e = getenv("MSSPE_FASTA_PINNED")) want_pinned = atoi(e) != 0;
  if (want_pinned && cudaMallocHost(&p, a) == cudaSuccess) { F->pinned = true; }
  else { if (want_pinned) (void)cudaGetLastError(); p = malloc(a); F->pinned = false; }  // also: no CUDA context (tests of the parser)
  F->bases = (uint8_t*)p;
  return F->bases;
}

}  // namespace

extern "C" int msspe_fasta_open(const char* path, uint32_t n_threads, msspe_fasta** out, char* err, size_t err_len) {
  if (!path || !out) { set_err(err, err_len, "null argument"); return MSSPE_ERR_INVALID; }
  *out = nullptr;
  msspe_fasta* F = new (std::nothrow) msspe_fasta();
  if (!F) return MSSPE_ERR_NOMEM;
  const int rc = parse_fasta(path, n_threads, F, [F](uint64_t b) { return host_alloc(F, b); }, nullptr, err, err_len);
  if (rc != MSSPE_OK) { msspe_fasta_close(F); return rc; }
  *out = F;
  return MSSPE_OK;
}

extern "C" void msspe_fasta_close(msspe_fasta* F) {
  if (!F) return;
  if (F->bases) { if (F->pinned) cudaFreeHost(F->bases); else free(F->bases); }
  delete F;
}
extern "C" uint32_t msspe_fasta_records(const msspe_fasta* F) { return F ? (uint32_t)F->names.size() : 0; }
extern "C" const char* msspe_fasta_name(const msspe_fasta* F, uint32_t i) { return F && i < F->names.size() ? F->names[i].c_str() : nullptr; }
extern "C" const uint8_t* msspe_fasta_bases(const msspe_fasta* F) { return F ? F->bases : nullptr; }
extern "C" const uint64_t* msspe_fasta_offsets(const msspe_fasta* F) { return F ? F->offsets.data() : nullptr; }

// ctx.cu
int msspe_load_begin(msspe_ctx* c, const uint64_t* offsets, uint32_t n);  // plan + device allocation, returns d_bases in ctx
int msspe_load_finish(msspe_ctx* c);

extern "C" int msspe_load_fasta(msspe_ctx* c, const char* path, uint32_t n_threads, msspe_fasta** out) {
  if (!c) return MSSPE_ERR_INVALID;
  if (!path || !out) { c->set_error("msspe_load_fasta: null argument"); return MSSPE_ERR_INVALID; }
  *out = nullptr;
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  msspe_fasta* F = new (std::nothrow) msspe_fasta();
  if (!F) return MSSPE_ERR_NOMEM;
  char err[256] = {0};
  int dev_rc = MSSPE_OK;
  cudaStream_t st = c->stream;
  const int device = c->device;
  const int rc = parse_fasta(
      path, n_threads, F,
      [&](uint64_t bytes) -> uint8_t* {
        uint8_t* h = host_alloc(F, bytes);
        if (F->names.empty()) { dev_rc = MSSPE_ERR_INVALID; c->set_error("No sequences found in the input file"); return h; }  // main.rs:652-654
        dev_rc = msspe_load_begin(c, F->offsets.data(), (uint32_t)F->names.size());
        return h;
      },
      [&](uint64_t b0, uint64_t b1) {
        if (dev_rc != MSSPE_OK || b1 <= b0) return;
        cudaSetDevice(device);  // called from a helper thread
        if (cudaMemcpyAsync(c->d_bases + b0, F->bases + b0, b1 - b0, cudaMemcpyHostToDevice, st) != cudaSuccess) {
          c->set_error("msspe_load_fasta: host-to-device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
          dev_rc = MSSPE_ERR_CUDA;
        }
      },
      err, sizeof err);
  if (rc != MSSPE_OK) { c->set_error("%s", err); msspe_fasta_close(F); return rc; }
  if (dev_rc == MSSPE_OK) dev_rc = msspe_load_finish(c);
  if (dev_rc != MSSPE_OK) { msspe_fasta_close(F); return dev_rc; }
  *out = F;
  return MSSPE_OK;
}
