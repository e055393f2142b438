// od-msspe (B200 engine) -- drop-in replacement of the reference's `od-msspe` process for the pre-aligned path
// (--do-align=false): same command-line flags and environment fallbacks (od-msspe/src/config.rs:11-148), FASTA
// input (main.rs:108-122), CSV primer output (main.rs:836-857) and stdout coverage report (main.rs:575-593).
// Everything between them runs on the GPU through the C ABI of include/od_msspe_b200.h; this file is the host
// glue the north star keeps on the CPU: argument parsing, FASTA parsing, the Primer3 text round trips, the
// 5-line ntthal parser's bookkeeping (delta_g.rs:27-59), the conflict-graph vertex cover (main.rs:754-815).
// There is no CPU fallback: without a usable CUDA device the program exits with an error.
#include <sys/stat.h>
#include <thread>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/od_msspe_b200.h"

namespace {

const char* kVersion = "1.2.0";  // od-msspe/Cargo.toml:3

struct Args {  // config.rs:11-148 with constants.rs defaults
  std::string input, output;
  uint64_t kmer_size = 13, window_size = 500, overlap_size = 250, max_iterations = 1000, search_windows_size = 50;
  bool has_mms = false; uint64_t max_mismatch_segments = 0;
  float mv_conc = 50.0f, dv_conc = 3.0f, dntp_conc = 0.0f, dna_conc = 250.0f, annealing_temp = 25.0f, min_tm = 30.0f,
        max_tm = 60.0f, max_self_dimer_any_tm = 47.0f, max_self_dimer_end_tm = 47.0f, max_hairpin_tm = 24.0f,
        delta_g_threshold = -9000.0f, tm_stddev = 2.0f;
  std::string keep_all = "false", check_cross_dimers = "true", check_self_dimers = "true", check_hairpin = "true",
              disable_tm_stddev = "false", disable_min_max_tm = "false", do_align = "true";
  std::string ntthal = "ntthal", primer3 = "primer3_core";
};

[[noreturn]] void clap_error(const std::string& msg) {
  std::cerr << "error: " << msg << "\n\nUsage: od-msspe [OPTIONS] --input <INPUT> --output <OUTPUT>\n\nFor more information, try '--help'.\n";
  exit(2);
}
[[noreturn]] void panic(const std::string& msg) {  // Rust panics exit with status 101
  std::cerr << "thread 'main' panicked:\n" << msg << "\n";
  exit(101);
}
bool log_enabled() { static int v = -1; if (v < 0) v = (getenv("RUST_LOG") || getenv("MSSPE_LOG")) ? 1 : 0; return v == 1; }
double elapsed_s() { static const auto t0 = std::chrono::steady_clock::now(); return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
void log_info(const std::string& m) {  // env_logger prints a wall-clock stamp here; the elapsed time is more useful for a pipeline
  if (!log_enabled()) return;
  char b[32]; snprintf(b, sizeof b, "%8.3fs", elapsed_s());
  std::cerr << "[" << b << " INFO  od_msspe] " << m << "\n";
}

struct Opt { const char* flag; const char* env; int kind; void* dst; };  // kind: 0 u64, 1 f32, 2 bool-string, 3 string, 4 optional u64

uint64_t parse_u64(const std::string& flag, const std::string& v) {
  if (v.empty() || v.find_first_not_of("0123456789") != std::string::npos) clap_error("invalid value '" + v + "' for '--" + flag + "': invalid digit found in string");
  return strtoull(v.c_str(), nullptr, 10);
}
float parse_f32(const std::string& flag, const std::string& v) {
  char* end = nullptr;
  float f = strtof(v.c_str(), &end);
  if (v.empty() || *end != 0) clap_error("invalid value '" + v + "' for '--" + flag + "': invalid float literal");
  return f;
}
void assign(const Opt& o, const std::string& v, Args& a) {
  switch (o.kind) {
    case 0: *(uint64_t*)o.dst = parse_u64(o.flag, v); break;
    case 1: *(float*)o.dst = parse_f32(o.flag, v); break;
    case 2: if (v != "true" && v != "false") clap_error("invalid value '" + v + "' for '--" + std::string(o.flag) + "'\n  [possible values: true, false]");
            *(std::string*)o.dst = v; break;
    case 3: *(std::string*)o.dst = v; break;
    case 4: a.max_mismatch_segments = parse_u64(o.flag, v); a.has_mms = true; break;
  }
}

void print_help() {
  std::cout <<
      "Usage: od-msspe [OPTIONS] --input <INPUT> --output <OUTPUT>\n\nOptions:\n"
      "  -i, --input <INPUT>\n  -o, --output <OUTPUT>\n"
      "      --kmer-size <KMER_SIZE>                          [env: KMER_SIZE=] [default: 13]\n"
      "      --window-size <WINDOW_SIZE>                      [env: WINDOW_SIZE=] [default: 500]\n"
      "      --overlap-size <OVERLAP_SIZE>                    [env: OVERLAP_SIZE=] [default: 250]\n"
      "      --max-mismatch-segments <MAX_MISMATCH_SEGMENTS>  Stop when k-mer frequency drops below this value. Defaults to max(1, min(10, num_sequences / 50)) [env: MAX_MISMATCH_SEGMENTS=]\n"
      "      --max-iterations <MAX_ITERATIONS>                [env: MAX_ITERATIONS=] [default: 1000]\n"
      "      --search-windows-size <SEARCH_WINDOWS_SIZE>      [env: SEARCH_WINDOWS_SIZE=] [default: 50]\n"
      "      --mv-conc <MV_CONC>                              [env: MV_CONC=] [default: 50]\n"
      "      --dv-conc <DV_CONC>                              [env: DV_CONC=] [default: 3]\n"
      "      --dntp-conc <DNTP_CONC>                          [env: DNTP_CONC=] [default: 0]\n"
      "      --dna-conc <DNA_CONC>                            [env: DNA_CONC=] [default: 250]\n"
      "      --annealing-temp <ANNEALING_TEMP>                [env: ANNEALING_TEMP=] [default: 25]\n"
      "      --min-tm <MIN_TM>                                [env: MIN_TM=] [default: 30]\n"
      "      --max-tm <MAX_TM>                                [env: MAX_TM=] [default: 60]\n"
      "      --max-self-dimer-any-tm <MAX_SELF_DIMER_ANY_TM>  [env: MAX_SELF_DIMER_ANY_TM=] [default: 47]\n"
      "      --max-self-dimer-end-tm <MAX_SELF_DIMER_END_TM>  [env: MAX_SELF_DIMER_END_TM=] [default: 47]\n"
      "      --max-hairpin-tm <MAX_HAIRPIN_TM>                [env: MAX_HAIRPIN_TM=] [default: 24]\n"
      "      --delta-g-threshold <DELTA_G_THRESHOLD>          Threshold for dG, default is -9000.0 J/mol [env: DELTA_G_THRESHOLD=] [default: -9000]\n"
      "      --keep-all <KEEP_ALL>                            Ignores all filtering and does NOT remove any primers. [env: KEEP_ALL=] [default: false] [possible values: true, false]\n"
      "      --check-cross-dimers <CHECK_CROSS_DIMERS>        [env: CHECK_CROSS_DIMERS=] [default: true] [possible values: true, false]\n"
      "      --check-self-dimers <CHECK_SELF_DIMERS>          [env: CHECK_SELF_DIMERS=] [default: true] [possible values: true, false]\n"
      "      --check-hairpin <CHECK_HAIRPIN>                  [env: CHECK_HAIRPIN=] [default: true] [possible values: true, false]\n"
      "      --tm-stddev <TM_STDDEV>                          [env: TM_STDDEV=] [default: 2]\n"
      "      --disable-tm-stddev <DISABLE_TM_STDDEV>          [env: DISABLE_TM_STDDEV=] [default: false] [possible values: true, false]\n"
      "      --disable-min-max-tm <DISABLE_MIN_MAX_TM>        [env: DISABLE_MIN_MAX_TM=] [default: false] [possible values: true, false]\n"
      "      --do-align <DO_ALIGN>                            Does MAFFT multiple sequence alignment (NOT supported by the B200 engine: pass false) [env: DO_ALIGN=] [default: true] [possible values: true, false]\n"
      "      --ntthal <NTTHAL>                                accepted for compatibility; thal runs on the GPU [env: NTTHAL=] [default: ntthal]\n"
      "      --primer3 <PRIMER3>                              accepted for compatibility; Tm/thal run on the GPU [env: PRIMER3=] [default: primer3_core]\n"
      "  -h, --help                                           Print help\n  -V, --version                                        Print version\n";
}

Args parse_args(int argc, char** argv) {
  Args a;
  bool has_in = false, has_out = false;
  std::vector<Opt> opts = {
      {"input", nullptr, 3, &a.input}, {"output", nullptr, 3, &a.output},
      {"kmer-size", "KMER_SIZE", 0, &a.kmer_size}, {"window-size", "WINDOW_SIZE", 0, &a.window_size},
      {"overlap-size", "OVERLAP_SIZE", 0, &a.overlap_size}, {"max-mismatch-segments", "MAX_MISMATCH_SEGMENTS", 4, nullptr},
      {"max-iterations", "MAX_ITERATIONS", 0, &a.max_iterations}, {"search-windows-size", "SEARCH_WINDOWS_SIZE", 0, &a.search_windows_size},
      {"mv-conc", "MV_CONC", 1, &a.mv_conc}, {"dv-conc", "DV_CONC", 1, &a.dv_conc}, {"dntp-conc", "DNTP_CONC", 1, &a.dntp_conc},
      {"dna-conc", "DNA_CONC", 1, &a.dna_conc}, {"annealing-temp", "ANNEALING_TEMP", 1, &a.annealing_temp},
      {"min-tm", "MIN_TM", 1, &a.min_tm}, {"max-tm", "MAX_TM", 1, &a.max_tm},
      {"max-self-dimer-any-tm", "MAX_SELF_DIMER_ANY_TM", 1, &a.max_self_dimer_any_tm},
      {"max-self-dimer-end-tm", "MAX_SELF_DIMER_END_TM", 1, &a.max_self_dimer_end_tm},
      {"max-hairpin-tm", "MAX_HAIRPIN_TM", 1, &a.max_hairpin_tm}, {"delta-g-threshold", "DELTA_G_THRESHOLD", 1, &a.delta_g_threshold},
      {"keep-all", "KEEP_ALL", 2, &a.keep_all}, {"check-cross-dimers", "CHECK_CROSS_DIMERS", 2, &a.check_cross_dimers},
      {"check-self-dimers", "CHECK_SELF_DIMERS", 2, &a.check_self_dimers}, {"check-hairpin", "CHECK_HAIRPIN", 2, &a.check_hairpin},
      {"tm-stddev", "TM_STDDEV", 1, &a.tm_stddev}, {"disable-tm-stddev", "DISABLE_TM_STDDEV", 2, &a.disable_tm_stddev},
      {"disable-min-max-tm", "DISABLE_MIN_MAX_TM", 2, &a.disable_min_max_tm}, {"do-align", "DO_ALIGN", 2, &a.do_align},
      {"ntthal", "NTTHAL", 3, &a.ntthal}, {"primer3", "PRIMER3", 3, &a.primer3}};
  for (auto& o : opts)  // environment fallbacks first; the command line overrides them
    if (o.env) { const char* e = getenv(o.env); if (e && *e) assign(o, e, a); }
  for (int i = 1; i < argc; i++) {
    std::string s = argv[i];
    if (s == "-h" || s == "--help") { print_help(); exit(0); }
    if (s == "-V" || s == "--version") { std::cout << "od-msspe " << kVersion << "\n"; exit(0); }
    std::string name, val; bool has_val = false;
    if (s.rfind("--", 0) == 0) {
      size_t eq = s.find('=');
      name = s.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
      if (eq != std::string::npos) { val = s.substr(eq + 1); has_val = true; }
    } else if (s.size() >= 2 && s[0] == '-' && (s[1] == 'i' || s[1] == 'o')) {
      name = s[1] == 'i' ? "input" : "output";
      if (s.size() > 2) { val = s.substr(s[2] == '=' ? 3 : 2); has_val = true; }
    } else {
      clap_error("unexpected argument '" + s + "' found");
    }
    const Opt* found = nullptr;
    for (auto& o : opts) if (name == o.flag) found = &o;
    if (!found) clap_error("unexpected argument '--" + name + "' found");
    if (!has_val) {
      if (i + 1 >= argc) clap_error("a value is required for '--" + name + "' but none was supplied");
      val = argv[++i];
    }
    assign(*found, val, a);
    if (name == "input") has_in = true;
    if (name == "output") has_out = true;
  }
  if (!has_in || !has_out) clap_error("the following required arguments were not provided:" + std::string(has_in ? "" : "\n  --input <INPUT>") + (has_out ? "" : "\n  --output <OUTPUT>"));
  return a;
}

struct Record { std::string name; };  // SequenceRecord.name; the sequences live in the msspe_fasta handle / on the device

std::string decode(uint64_t code, unsigned k) { std::string s(k, 'A'); for (unsigned i = 0; i < k; i++) s[i] = "ACGT"[(code >> (2 * (k - 1 - i))) & 3]; return s; }
uint64_t revcomp_code(uint64_t code, unsigned k) { uint64_t r = 0; for (unsigned t = 0; t < k; t++) { r = (r << 2) | (3u - (code & 3u)); code >>= 2; } return r; }
float via_text(double v, const char* fmt) { char b[64]; snprintf(b, sizeof b, fmt, v); return strtof(b, nullptr); }
std::string fmt2(float x) { char b[64]; snprintf(b, sizeof b, "%.2f", (double)x); return b; }  // Rust {:.2} of an f32
std::string fmt1(float x) { if (std::isnan(x)) return "NaN"; if (std::isinf(x)) return x > 0 ? "inf" : "-inf"; char b[64]; snprintf(b, sizeof b, "%.1f", (double)x); return b; }

struct KmerStat { uint64_t code; std::string word; uint8_t direction; float gc_percent, mean, std, tm; bool tm_ok; float self_any_th, self_end_th, hairpin_th; bool runs; bool keep; };


#define CHECK(call) do { int rc_ = (call); if (rc_ != MSSPE_OK) { std::cerr << "od-msspe: " #call " failed (" << rc_ << "): " << msspe_last_error(ctx) << "\n"; exit(1); } } while (0)

// get_kmer_stats + filter_kmers (main.rs:408-516) for both directions through the ABI: msspe_kmer_stats_both does the
// device thermodynamics in one batch and the reference's text round trips / f32 statistics / strict comparisons.
void kmer_stats_both(msspe_ctx* ctx, const std::vector<msspe_candidate> cand[2], unsigned k, const Args& a, std::vector<KmerStat> out[2]) {
  std::vector<uint64_t> codes[2];
  std::vector<msspe_kmer_stat> st[2];
  for (int d = 0; d < 2; d++) {
    codes[d].resize(cand[d].size());
    for (size_t i = 0; i < cand[d].size(); i++) codes[d][i] = cand[d][i].code;
    st[d].resize(cand[d].size() ? cand[d].size() : 1);
  }
  msspe_filter_cfg fc{a.min_tm, a.max_tm, a.max_self_dimer_any_tm, a.max_self_dimer_end_tm, a.max_hairpin_tm, a.tm_stddev,
                      (uint8_t)(a.check_self_dimers == "true"), (uint8_t)(a.check_hairpin == "true"),
                      (uint8_t)(a.disable_tm_stddev == "true"), (uint8_t)(a.disable_min_max_tm == "true")};
  CHECK(msspe_kmer_stats_both(ctx, codes[0].data(), (uint32_t)codes[0].size(), codes[1].data(), (uint32_t)codes[1].size(), k, &fc, st[0].data(), st[1].data()));
  for (int d = 0; d < 2; d++) {
    out[d].resize(codes[d].size());
    for (size_t i = 0; i < codes[d].size(); i++) {
      KmerStat& s = out[d][i];
      const msspe_kmer_stat& t = st[d][i];
      s.code = codes[d][i]; s.word = decode(codes[d][i], k); s.direction = (uint8_t)d;
      s.gc_percent = t.gc_percent; s.mean = t.mean; s.std = t.std; s.tm = t.tm; s.tm_ok = t.tm_ok != 0;
      s.self_any_th = t.self_any_th; s.self_end_th = t.self_end_th; s.hairpin_th = t.hairpin_th; s.runs = t.runs != 0;
      s.keep = t.keep != 0;
    }
  }
}

std::vector<KmerStat> filter_kmers(const std::vector<KmerStat>& st) {
  std::vector<KmerStat> out;
  for (auto& k : st) if (k.keep) out.push_back(k);
  return out;
}

}  // namespace

int main(int argc, char** argv) {
  elapsed_s();
  Args a = parse_args(argc, argv);
  if (a.do_align == "true") {
    std::cerr << "od-msspe (B200 engine): MAFFT alignment is out of scope of this engine; align the input first and pass --do-align=false\n";
    return 2;
  }
  {  // the two input errors the reference raises before anything else, from the first non-empty line alone
    std::ifstream in(a.input, std::ios::binary);
    if (!in) { std::cerr << "Error: Os { code: 2, kind: NotFound, message: \"No such file or directory\" }\n"; return 1; }
    std::string line; bool any = false;
    while (std::getline(in, line)) {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      if (line.empty()) continue;
      any = true;
      if (line[0] != '>') panic("called `Result::unwrap()` on an `Err` value: InvalidStart (FASTA must begin with '>')");
      break;
    }
    if (!any) panic("No sequences found in the input file");                                    // main.rs:652-654
  }
  log_info("Aligning sequences...");
  log_info(".... SKIPPED.");
  if (a.overlap_size < a.search_windows_size) panic("Overlap windows size must be greater or equal than search windows size");  // main.rs:201-203
  if (a.overlap_size == 0) panic("assertion failed: step != 0");                                 // step_by(0)
  if (a.window_size == 0) panic("window size must be non-zero");                                  // windows(0)
  const unsigned k = (unsigned)a.kmer_size;

  // Cold start: creating the CUDA context (about 2 s on a fresh process) and parsing the FASTA (main.rs:108-122, host only)
  // are independent, so the context is created on a second thread while this one parses into pageable memory; the pool's
  // first touch for an input of this size follows on that thread's heels (msspe_reserve_pool).  MSSPE_SERIAL_START=1 restores
  // the one-after-the-other order (parse into pinned memory with chunked uploads, msspe_load_fasta).
  msspe_ctx* ctx = nullptr;
  msspe_config cfg{(uint32_t)a.kmer_size, (uint32_t)a.window_size, (uint32_t)a.overlap_size, (uint32_t)a.search_windows_size,
                   getenv("MSSPE_DEVICE") ? atoi(getenv("MSSPE_DEVICE")) : 0, 0};
  const uint32_t n_threads = getenv("MSSPE_THREADS") ? (uint32_t)atoi(getenv("MSSPE_THREADS")) : 0u;
  uint64_t reserve = 0;
  {
    struct stat sb;
    if (!getenv("MSSPE_NO_RESERVE") && stat(a.input.c_str(), &sb) == 0 && sb.st_size > (64 << 20))
      reserve = (uint64_t)sb.st_size * 13u;   // genomes + keys, postings and forward index of both directions
  }
  msspe_fasta* fasta = nullptr;
  if (getenv("MSSPE_SERIAL_START")) {
    if (int rc = msspe_create(&cfg, &ctx)) { std::cerr << "od-msspe: cannot start the GPU engine (" << rc << "): " << msspe_last_error(nullptr) << "\n"; return 1; }
    log_info("GPU engine ready");
    if (reserve) msspe_reserve_pool(ctx, reserve);
    // 1. to_records (main.rs:108-122) + upload: multi-threaded parse into pinned memory, chunked copies overlap the parse
    if (int rc = msspe_load_fasta(ctx, a.input.c_str(), n_threads, &fasta)) {
      const std::string msg = msspe_last_error(ctx);
      if (rc == MSSPE_ERR_INVALID) panic(msg);                                                      // InvalidStart, or main.rs:652-654
      std::cerr << "od-msspe: reading " << a.input << " failed (" << rc << "): " << msg << "\n";
      return 1;
    }
  } else {
    int rc_create = 0;
    std::string create_err;
    std::thread creator([&] {
      rc_create = msspe_create(&cfg, &ctx);
      if (rc_create) create_err = msspe_last_error(nullptr);
      else if (reserve) msspe_reserve_pool(ctx, reserve);
    });
    setenv("MSSPE_FASTA_PINNED", "0", 0);      // page-locking would wait for the context that is still being created
    char err[512] = {0};
    const int rc_parse = msspe_fasta_open(a.input.c_str(), n_threads, &fasta, err, sizeof err);
    creator.join();
    if (rc_create) { std::cerr << "od-msspe: cannot start the GPU engine (" << rc_create << "): " << create_err << "\n"; return 1; }
    log_info("GPU engine ready");
    if (rc_parse) {
      if (rc_parse == MSSPE_ERR_INVALID) panic(err);                                                // InvalidStart
      std::cerr << "od-msspe: reading " << a.input << " failed (" << rc_parse << "): " << err << "\n";
      return 1;
    }
    if (msspe_fasta_records(fasta) == 0) panic("No sequences found in the input file");           // main.rs:652-654
    CHECK(msspe_load_genomes(ctx, msspe_fasta_bases(fasta), msspe_fasta_offsets(fasta), msspe_fasta_records(fasta)));
  }
  std::vector<Record> records(msspe_fasta_records(fasta));
  for (size_t i = 0; i < records.size(); i++) records[i].name = msspe_fasta_name(fasta, (uint32_t)i);  // sequences stay in the handle
  const uint64_t mms = a.has_mms ? a.max_mismatch_segments
                                 : std::min<uint64_t>(10, std::max<uint64_t>(1, (records.size() + 49) / 50));  // main.rs:658-660
  log_info("max_mismatch_segments=" + std::to_string(mms) + " (auto-scaled from " + std::to_string(records.size()) + " sequences)");
  {  // `ntthal -path <cwd>/primer3_config/` (delta_g.rs:90): use that directory when present, else the embedded tables
    msspe_thal_raw_params* p = new msspe_thal_raw_params();
    char err[256] = {0};
    if (msspe_thal_params_from_dir("primer3_config", p, err, sizeof err) == MSSPE_OK) { CHECK(msspe_set_thal_params(ctx, p)); log_info("thermodynamic parameters: ./primer3_config/"); }
    else log_info("thermodynamic parameters: embedded tables");
    delete p;
  }
  // 2. segments + inverted index (get_segment_manager, main.rs:693)
  log_info("Read " + std::to_string(records.size()) + " sequences; extracting n-grams from each sequence segments...");
  CHECK(msspe_build_index(ctx));
  uint64_t G = 0; uint32_t maxp = 0, slots = 0;
  CHECK(msspe_segment_info(ctx, &G, &maxp, &slots));
  if (G == 0) panic("called `Option::unwrap()` on a `None` value");                               // main.rs:694-699
  log_info("Done, total partitions: " + std::to_string(maxp) + ", total segments: " + std::to_string(G));
  // 3. greedy selection, both directions (main.rs:709-714)
  log_info("Calculating frequencies of k-mer for all segments...");
  std::vector<msspe_candidate> cand[2];
  cand[0].resize(a.max_iterations ? a.max_iterations : 1); cand[1].resize(a.max_iterations ? a.max_iterations : 1);
  uint32_t nf = 0, nr = 0;
  CHECK(msspe_select_both(ctx, (uint32_t)a.max_iterations, (uint32_t)std::min<uint64_t>(mms, 0xFFFFFFFFull), MSSPE_SELECT_AUTO,
                          cand[0].data(), &nf, cand[1].data(), &nr));
  cand[0].resize(nf); cand[1].resize(nr);
  log_info("Done calculating, Total candidate k-mers: fwd: " + std::to_string(nf) + ", rev: " + std::to_string(nr));
  // 4. thermodynamic filters (main.rs:723-732)
  const bool keep_all = a.keep_all == "true";
  std::vector<KmerStat> stats[2], primers_dir[2];
  kmer_stats_both(ctx, cand, k, a, stats);
  for (int d = 0; d < 2; d++) primers_dir[d] = keep_all ? stats[d] : filter_kmers(stats[d]);
  // cross dimers (run_ntthal, main.rs:752 / delta_g.rs:61-153)
  std::vector<const KmerStat*> primers;
  for (int d = 0; d < 2; d++) for (auto& s : primers_dir[d]) primers.push_back(&s);
  const uint32_t n = (uint32_t)primers.size();
  const bool cross = a.check_cross_dimers == "true", self = a.check_self_dimers == "true";
  std::set<std::string> deleted;
  if (cross && n > 0) {
    auto r2 = [](float v) { char b[64]; snprintf(b, sizeof b, "%.2f", (double)v); return atof(b); };   // argv "{:.2}" (delta_g.rs:97-105)
    msspe_thal_cond cond{r2(a.mv_conc), r2(a.dv_conc), r2(a.dntp_conc), r2(a.dna_conc), r2(a.annealing_temp), 30, 0};
    std::vector<uint64_t> codes(n);
    for (uint32_t i = 0; i < n; i++) codes[i] = primers[i]->code;
    // input lines of format_ntthal_input (delta_g.rs:61-81): every ordered pair, a-major, minus the self/revcomp pairs
    std::vector<uint64_t> lines; lines.reserve((size_t)n * n);
    for (uint32_t i = 0; i < n; i++)
      for (uint32_t j = 0; j < n; j++) {
        if (!self && (codes[i] == codes[j] || revcomp_code(codes[j], k) == codes[i])) continue;
        lines.push_back((uint64_t)i * n + j);
      }
    const double limit = (double)a.delta_g_threshold + 1.0;  // superset; the exact "%g" -> f32 comparison happens below
    std::vector<msspe_dimer_edge> edges((size_t)n * n);
    std::vector<uint64_t> nos((size_t)n * n);
    uint64_t ne = 0, nn = 0;
    CHECK(msspe_cross_dimer(ctx, codes.data(), n, k, &cond, 0, n, limit, edges.data(), edges.size(), &ne, nos.data(), nos.size(), &nn));
    std::unordered_map<uint64_t, double> dg_of;
    for (uint64_t e = 0; e < ne; e++) dg_of[edges[e].pair] = edges[e].dg;
    std::unordered_set<uint64_t> nos_set(nos.begin(), nos.begin() + nn);
    // parse_ntthal_output (delta_g.rs:27-59): input line t reads output line 5t and skips four more.  The reference's ntthal
    // (Primer3 2.6.1) prints a 5-line block per pair and NOTHING for a pair without structure (its own stdout under
    // tools/a64emu, tests/golden/ntthal_emulated.json), so output block t belongs to the t-th pair THAT HAS a structure while
    // the parser credits it to input line t: after m structure-less pairs every edge lands m input lines early, and the last
    // m input lines read nothing.
    // conflict edges (main.rs:755-771) over the DISTINCT words (the reference keys its graph by word)
    std::unordered_map<std::string, uint32_t> node_of;
    std::vector<uint64_t> node_code; std::vector<const std::string*> node_word;
    std::vector<uint32_t> node_of_primer(n);
    for (uint32_t i = 0; i < n; i++) {
      auto ins = node_of.emplace(primers[i]->word, (uint32_t)node_code.size());
      if (ins.second) { node_code.push_back(primers[i]->code); node_word.push_back(&primers[i]->word); }
      node_of_primer[i] = ins.first->second;
    }
    std::vector<uint32_t> ea, eb;
    std::set<std::string> edge_ids;
    uint64_t m = 0;
    for (uint64_t u = 0; u < lines.size(); u++) {
      const uint64_t pu = lines[u];
      if (nos_set.count(pu)) { m++; continue; }
      auto it = dg_of.find(pu);
      if (it == dg_of.end()) continue;
      const float dg = via_text(it->second, "%g");                    // ntthal prints "%g"; the parser reads an f32
      if (!(dg < a.delta_g_threshold)) continue;
      const uint64_t t = u - m;
      const uint64_t pt = lines[t];
      const std::string& wa = primers[pt / n]->word; const std::string& wb = primers[pt % n]->word;
      if (!edge_ids.insert(wa + ":" + wb).second) continue;           // HashSet<Edge> keyed by id: first insert stays
      const float stored = strtof(fmt2(dg).c_str(), nullptr);         // attrs "dg" = format!("{:.2}", dg); get_dg() re-parses
      if (stored < a.delta_g_threshold) { ea.push_back(node_of_primer[pt / n]); eb.push_back(node_of_primer[pt % n]); }
    }
    // greedy vertex cover on the device (main.rs:776-798: most active conflicts, ties -> lexicographically greatest)
    std::vector<uint8_t> del(node_code.size(), 0);
    uint32_t n_del = 0;
    CHECK(msspe_vertex_cover(ctx, node_code.data(), (uint32_t)node_code.size(), ea.data(), eb.data(), ea.size(), del.data(), &n_del));
    for (size_t v = 0; v < del.size(); v++) if (del[v]) deleted.insert(*node_word[v]);
  }
  std::vector<KmerStat> final_dir[2];
  for (int d = 0; d < 2; d++) {
    if (keep_all) final_dir[d] = primers_dir[d];
    else for (auto& p : primers_dir[d]) if (!deleted.count(p.word)) final_dir[d].push_back(p);
  }
  log_info("Filtering out unmatched, primers left is fwd=" + std::to_string(final_dir[0].size()) + ", rev=" + std::to_string(final_dir[1].size()));
  // 5. coverage report (main.rs:518-594)
  {
    std::vector<uint64_t> fc, rc;
    for (auto& p : final_dir[0]) fc.push_back(p.code);
    for (auto& p : final_dir[1]) rc.push_back(p.code);
    // per-record / per-partition (covered, total) reduced on the device: O(records + partitions) comes back
    const uint32_t nrec = (uint32_t)records.size(), npart = maxp + 1;
    std::vector<uint32_t> rcv(nrec), rtot(nrec), pcv(npart), ptot(npart);
    uint64_t ncov64 = 0;
    CHECK(msspe_coverage_summary(ctx, fc.data(), (uint32_t)fc.size(), rc.data(), (uint32_t)rc.size(), rcv.data(), rtot.data(), nrec,
                                 pcv.data(), ptot.data(), npart, &ncov64));
    const size_t ncov = (size_t)ncov64;
    std::unordered_map<std::string, std::pair<size_t, size_t>> seq; std::map<uint16_t, std::pair<size_t, size_t>> ps;
    for (uint32_t r = 0; r < nrec; r++) if (rtot[r]) { auto& se = seq[records[r].name]; se.first += rcv[r]; se.second += rtot[r]; }
    for (uint32_t q = 0; q < npart; q++) if (ptot[q]) { auto& pe = ps[(uint16_t)q]; pe.first += pcv[q]; pe.second += ptot[q]; }
    float mn = INFINITY, mx = -INFINITY; size_t well = 0;
    for (auto& kv : seq) { const float c = (float)kv.second.first / (float)kv.second.second * 100.0f; mn = std::fmin(mn, c); mx = std::fmax(mx, c); if (c >= 80.0f) well++; }
    printf("\nCoverage report:\n");
    printf("  Segments:  %zu/%llu covered (%s%%)\n", ncov, (unsigned long long)G, fmt1(100.0f * (float)ncov / (float)G).c_str());
    printf("  Sequences: %zu/%zu at \xE2\x89\xA5" "80%% coverage (min %s%%, max %s%%)\n", well, seq.size(), fmt1(mn).c_str(), fmt1(mx).c_str());
    std::string unc; bool any = false;
    for (auto& kv : ps) if (kv.second.first == 0) { if (any) unc += ", "; unc += std::to_string(kv.first); any = true; }
    if (!any) printf("  All partitions have primer coverage\n"); else printf("  Uncovered partitions: [%s]\n", unc.c_str());
  }
  // 6. CSV (main.rs:836-857)
  log_info("Coverage report done; outputting primers...");
  std::ofstream out(a.output, std::ios::binary);
  if (!out) { std::cerr << "Error: cannot create " << a.output << "\n"; return 1; }
  out << "direction,name,primers,gc,avg,std,tm\n";
  for (int d = 0; d < 2; d++) {
    size_t idx = 0;
    for (auto& p : final_dir[d]) {
      const char* dir = p.direction == 0 ? "F" : "R";
      out << dir << ",Primer_" << idx << "_" << dir << "," << p.word << "," << fmt2(p.gc_percent / 100.0f) << "," << fmt2(p.mean) << ","
          << fmt2(p.std) << "," << fmt2(p.tm) << "\n";
      idx++;
    }
  }
  out.close();
  log_info("Done outputting primers");
  msspe_fasta_close(fasta);
  msspe_destroy(ctx);
  return 0;
}
