// kmer_oracle.cpp -- TEST INFRASTRUCTURE ONLY (oracle / CPU baseline "port").
//
// C++ restatement of the reference pipeline od-msspe/src/main.rs with the reference's own data structures
// and asymptotic behaviour: heap std::string per k-mer, hash maps keyed by (word, direction), a FULL
// recount of every live segment in every greedy iteration (main.rs:292-309), N^2 ordered thal pairs
// exchanged as TEXT and parsed line by line (delta_g.rs:27-59), string-keyed conflict sets.
// Single-threaded like the reference.  It exists to (1) check the CUDA engine at sizes where the pure
// Python oracle (kmer_oracle.py) is too slow, (2) be the timed CPU baseline in bench.py.
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library; the product
// (open-msspe-design_b200/) never does.  Validated against kmer_oracle.py and the reference's unit-test
// vectors in tests/test_oracle_kmer.py.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <cctype>
#include <unistd.h>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../include/od_msspe_b200.h"

extern "C" {
int oracle_thal(const char* o1, const char* o2, int type, const msspe_thal_cond* c, msspe_thal_out* out);
double oracle_oligotm(const char* s, double mv, double dv, double dntp, double dna_conc);
double oracle_gc_percent(const char* s);
}

namespace {

struct SequenceRecord { std::string name, sequence; };                      // main.rs:21-24
struct KmerRecord { std::string word; uint8_t direction; };                 // main.rs:26-30
struct Segment {                                                            // main.rs:82-87
  int record; uint16_t partition_no; size_t index; std::vector<KmerRecord> kmers[2];
};
struct KeyHash {
  size_t operator()(const std::pair<std::string, uint8_t>& k) const {
    return std::hash<std::string>()(k.first) * 31u + k.second;
  }
};
using Key = std::pair<std::string, uint8_t>;
using Mapping = std::unordered_map<Key, std::vector<uint32_t>, KeyHash>;

// main.rs:108-122 (seq_io id() = up to first space; full_seq joins lines; uppercase; U->T)
std::vector<SequenceRecord> to_records(const char* src, size_t len) {
  std::vector<SequenceRecord> recs;
  size_t pos = 0; bool have = false; SequenceRecord cur;
  while (pos < len) {
    size_t e = pos; while (e < len && src[e] != '\n') e++;
    size_t le = e; if (le > pos && src[le - 1] == '\r') le--;
    if (le > pos && src[pos] == '>') {
      if (have) recs.push_back(cur);
      cur = SequenceRecord(); have = true;
      size_t s = pos + 1, t = s; while (t < le && src[t] != ' ') t++;
      cur.name.assign(src + s, t - s);
    } else if (have) {
      for (size_t i = pos; i < le; i++) {
        char c = (char)toupper((unsigned char)src[i]); if (c == 'U') c = 'T';
        cur.sequence.push_back(c);
      }
    }
    pos = e + 1;
  }
  if (have) recs.push_back(cur);
  return recs;
}

std::string reverse_complement(const std::string& s) {                      // main.rs:148-161
  std::string r; r.reserve(s.size());
  for (auto it = s.rbegin(); it != s.rend(); ++it) {
    char c = *it;
    switch (c) { case 'A': c = 'T'; break; case 'T': c = 'A'; break; case 'U': c = 'A'; break;
                 case 'C': c = 'G'; break; case 'G': c = 'C'; break; default: break; }
    r.push_back(c);
  }
  return r;
}

std::vector<std::string> find_kmers(const std::string& seq, size_t k) {      // main.rs:163-171
  std::vector<std::string> out; std::unordered_set<std::string> seen;
  if (seq.size() < k) return out;
  for (size_t i = 0; i + k <= seq.size(); i++) {
    std::string w = seq.substr(i, k);
    bool ok = true;
    for (char c : w) if (!strchr("ATCGU", c)) { ok = false; break; }
    if (ok && seen.insert(w).second) out.push_back(w);
  }
  return out;
}

std::vector<std::string> partitioning_sequence(const std::string& s, size_t size, size_t step) {  // main.rs:173-181
  std::vector<std::string> out;
  for (size_t st = 0; st + size <= s.size(); st += step) out.push_back(s.substr(st, size));
  return out;
}

struct Manager { std::vector<SequenceRecord> records; std::vector<Segment> segments; };

// main.rs:196-235
void get_segment_manager(Manager& m, size_t W, size_t S, size_t w, size_t k) {
  for (size_t r = 0; r < m.records.size(); r++) {
    auto parts = partitioning_sequence(m.records[r].sequence, W, S);
    for (size_t j = 0; j < parts.size(); j++) {
      std::string start = parts[j].substr(0, w), end = parts[j].substr(parts[j].size() - w);
      Segment seg; seg.record = (int)r; seg.partition_no = (uint16_t)j; seg.index = m.segments.size();
      for (auto& x : find_kmers(start, k)) seg.kmers[0].push_back({x, 0});
      for (auto& x : find_kmers(end, k)) seg.kmers[1].push_back({reverse_complement(x), 1});
      m.segments.push_back(std::move(seg));
    }
  }
}

Mapping make_mapping(const std::vector<Segment>& segs) {                    // main.rs:237-255
  Mapping mp;
  for (auto& s : segs) for (int d = 0; d < 2; d++) for (auto& km : s.kmers[d]) {
    if (km.direction != d) continue;
    mp[{km.word, km.direction}].push_back((uint32_t)s.index);
  }
  return mp;
}

float partition_tie_score(const Key& k, const Mapping& mp, const std::vector<Segment>& segs,
                          const std::unordered_set<uint32_t>& ignored,
                          const std::unordered_map<uint16_t, size_t>& cov) {   // main.rs:261-283
  std::unordered_set<uint16_t> seen; float score = 0.0f;
  auto it = mp.find(k);
  if (it != mp.end()) for (uint32_t idx : it->second) {
    if (ignored.count(idx)) continue;
    uint16_t p = segs[idx].partition_no;
    if (seen.insert(p).second) {
      auto c = cov.find(p); size_t already = c == cov.end() ? 0 : c->second;
      score += 1.0f / ((float)already + 1.0f);
    }
  }
  return score;
}

struct Winner { bool some; std::string word; size_t freq; size_t n_tied; float score; };

Winner find_most_freq_kmer(const std::vector<Segment>& segs, uint8_t dir, const std::unordered_set<uint32_t>& ignored,
                           const Mapping& mp, const std::unordered_map<uint16_t, size_t>& cov, uint64_t* evals) {
  std::unordered_map<std::string, size_t> freq;                             // main.rs:292 (fresh map every call)
  for (size_t idx = 0; idx < segs.size(); idx++) {
    if (ignored.count((uint32_t)idx)) continue;
    for (auto& km : segs[idx].kmers[dir]) { freq[km.word] += 1; (*evals)++; }
  }
  Winner w{false, "", 0, 0, 0.f};
  if (freq.empty()) return w;
  size_t mx = 0; for (auto& kv : freq) mx = std::max(mx, kv.second);
  bool have = false;
  for (auto& kv : freq) {
    if (kv.second != mx) continue;
    w.n_tied++;
    float s = partition_tie_score({kv.first, dir}, mp, segs, ignored, cov);
    if (!have || s > w.score || (s == w.score && kv.first < w.word)) { w.word = kv.first; w.score = s; have = true; }
  }
  w.some = true; w.freq = mx;
  return w;
}

struct Cand { std::string word; size_t freq; size_t n_tied; float score; };

std::vector<Cand> find_candidates_kmers(const Manager& m, uint8_t dir, size_t max_iter, size_t mms, uint64_t* evals,
                                        const Mapping* prebuilt = nullptr) {
  std::vector<Cand> out; Mapping own; if (!prebuilt) own = make_mapping(m.segments);   // main.rs:331-406 (:337 builds the mapping)
  const Mapping& mp = prebuilt ? *prebuilt : own;
  std::unordered_set<uint32_t> ignored; std::unordered_map<uint16_t, size_t> cov;
  for (size_t it = 0; it < max_iter; it++) {
    Winner w = find_most_freq_kmer(m.segments, dir, ignored, mp, cov, evals);
    if (!w.some) break;
    if (w.freq == 1) break;
    out.push_back({w.word, w.freq, w.n_tied, w.score});
    std::unordered_set<uint16_t> newly;
    for (uint32_t idx : mp.at({w.word, dir})) { ignored.insert(idx); newly.insert(m.segments[idx].partition_no); }
    for (uint16_t p : newly) cov[p] += 1;
    if (w.freq < mms) break;
  }
  return out;
}

// ---- text round trips -------------------------------------------------------------------------------
float parse_f32(const char* s) { return strtof(s, nullptr); }
float via_text(double v, const char* fmt) { char b[64]; snprintf(b, sizeof b, fmt, v); return parse_f32(b); }

// Rust `{:.2}` of an f32: exact decimal expansion, ties-to-even on the exact value == glibc printf of the
// value widened to double (widening is exact).
std::string fmt2(float x) { char b[64]; snprintf(b, sizeof b, "%.2f", (double)x); return b; }
std::string fmt1(float x) {
  if (std::isnan(x)) return "NaN"; if (std::isinf(x)) return x > 0 ? "inf" : "-inf";
  char b[64]; snprintf(b, sizeof b, "%.1f", (double)x); return b;
}

struct PrimerInfo { std::string id; float tm, gc, self_any_th, self_end_th, hairpin_th; };
struct KmerStat { std::string word; uint8_t direction; float gc_percent, mean, std, tm; bool tm_ok;
                  float self_any_th, self_end_th, hairpin_th; bool runs; };

struct Config {
  uint64_t kmer_size, window_size, overlap_size, max_mismatch_segments /*0 = auto*/, max_iterations,
      search_windows_size;
  float mv_conc, dv_conc, dntp_conc, dna_conc, annealing_temp, min_tm, max_tm, max_self_dimer_any_tm,
      max_self_dimer_end_tm, max_hairpin_tm, delta_g_threshold, tm_stddev;
  int keep_all, check_cross_dimers, check_self_dimers, check_hairpin, disable_tm_stddev, disable_min_max_tm;
};

// ---- external-tool mode (tests/test_shims.py): like the reference, spawn the executables named by ORACLE_PRIMER3 /
// ORACLE_NTTHAL (the reference's --primer3 / --ntthal, config.rs:142-147), feed them the reference's input text and
// read their output with restatements of the reference's parsers.  Unset = the in-process oracle arithmetic.
std::string run_tool(const std::string& cmd, const std::string& input) {
  char in_path[] = "/tmp/oracle_tool_in_XXXXXX", out_path[] = "/tmp/oracle_tool_out_XXXXXX";
  int fi = mkstemp(in_path), fo = mkstemp(out_path);
  if (fi < 0 || fo < 0) return "";
  FILE* f = fdopen(fi, "wb"); fwrite(input.data(), 1, input.size(), f); fclose(f);
  close(fo);
  const std::string full = cmd + " < " + in_path + " > " + out_path;
  const int rc = system(full.c_str());
  std::string out;
  if (rc == 0) { FILE* g = fopen(out_path, "rb"); if (g) { char b[65536]; size_t n; while ((n = fread(b, 1, sizeof b, g)) > 0) out.append(b, n); fclose(g); } }
  remove(in_path); remove(out_path);
  return out;
}

// primer3_core check_primers emulation: primer.rs:125-166 (+ Primer3 defaults: mv 50, dv 1.5, dNTP 0.6,
// DNA 50 nM, thal at 37 C, maxLoop 30; "%.3f" for TM/GC, "%.2f" for *_TH, then parse::<f32>).
std::vector<PrimerInfo> check_primers(const std::vector<std::string>& primers, float min_tm, float max_tm) {
  std::vector<PrimerInfo> out;
  if (const char* exe = getenv("ORACLE_PRIMER3")) {
    std::string in;  // format_primer3_input, primer.rs:125-140
    char b[128];
    for (auto& p : primers) {
      in += "SEQUENCE_ID=" + p + "\nSEQUENCE_PRIMER=" + p + "\nPRIMER_TASK=check_primers\nPRIMER_MIN_SIZE=13\n";
      snprintf(b, sizeof b, "PRIMER_MIN_TM=%.2f\nPRIMER_MAX_TM=%.2f\nPRIMER_OPT_TM=%.2f\n", (double)min_tm, (double)max_tm, (double)max_tm);
      in += b; in += "PRIMER_PICK_ANYWAY=1\n=\n";
    }
    const std::string text = run_tool(exe, in);
    PrimerInfo cur{"", 0, 0, 0, 0, 0};  // parse_primer3_output, primer.rs:67-114
    size_t pos = 0;
    while (pos < text.size()) {
      size_t e = text.find('\n', pos); if (e == std::string::npos) e = text.size();
      const std::string line = text.substr(pos, e - pos); pos = e + 1;
      if (line == "=") { if (cur.id.empty()) break; out.push_back(cur); cur = PrimerInfo{"", 0, 0, 0, 0, 0}; continue; }
      const size_t q = line.find('='); if (q == std::string::npos) continue;
      const std::string k = line.substr(0, q); std::string v = line.substr(q + 1); { const size_t q2 = v.find('='); if (q2 != std::string::npos) v = v.substr(0, q2); }
      if (k == "SEQUENCE_ID") cur.id = v;
      else if (k == "PRIMER_LEFT_0_TM") cur.tm = parse_f32(v.c_str());
      else if (k == "PRIMER_LEFT_0_GC_PERCENT") cur.gc = parse_f32(v.c_str());
      else if (k == "PRIMER_LEFT_0_SELF_ANY_TH") cur.self_any_th = parse_f32(v.c_str());
      else if (k == "PRIMER_LEFT_0_SELF_END_TH") cur.self_end_th = parse_f32(v.c_str());
      else if (k == "PRIMER_LEFT_0_HAIRPIN_TH") cur.hairpin_th = parse_f32(v.c_str());
    }
    return out;
  }
  msspe_thal_cond c{50.0, 1.5, 0.6, 50.0, 37.0, 30, 0};
  for (auto& p : primers) {
    PrimerInfo pi; pi.id = p;
    pi.tm = via_text(oracle_oligotm(p.c_str(), 50.0, 1.5, 0.6, 50.0), "%.3f");
    pi.gc = via_text(oracle_gc_percent(p.c_str()), "%.3f");
    msspe_thal_out o;
    oracle_thal(p.c_str(), p.c_str(), MSSPE_THAL_ANY, &c, &o);
    pi.self_any_th = via_text(o.tm < 0.0 ? 0.0 : o.tm, "%.2f");
    oracle_thal(p.c_str(), p.c_str(), MSSPE_THAL_END1, &c, &o);
    pi.self_end_th = via_text(o.tm < 0.0 ? 0.0 : o.tm, "%.2f");
    oracle_thal(p.c_str(), p.c_str(), MSSPE_THAL_HAIRPIN, &c, &o);
    pi.hairpin_th = via_text(o.tm < 0.0 ? 0.0 : o.tm, "%.2f");
    out.push_back(pi);
  }
  return out;
}

// std-dev 0.1.0 `standard_deviation(&[f32])`: mean, then sqrt(sum((x-mean)^2)/(n-1)) in f32 (PARITY
// UNPINNED: the crate is not vendored; divisor n-1 per SURVEY.md a-11).
void get_tm_stat(const std::vector<PrimerInfo>& l, float* mean, float* sd) {    // main.rs:462-467
  float sum = 0.0f; for (auto& i : l) sum += i.tm;
  *mean = sum / (float)l.size();
  float m2 = 0.0f; for (auto& i : l) m2 += i.tm;   // the crate recomputes its own mean
  m2 = m2 / (float)l.size();
  float sq = 0.0f; for (auto& i : l) { float d = i.tm - m2; sq += d * d; }
  *sd = std::sqrt(sq / (float)(l.size() - 1));
}

bool is_run(const std::string& k) {                                           // main.rs:478-490
  int runs = 0; char last = ' ';
  for (char c : k) { if (c == last) runs++; else runs = 0; last = c; }
  return runs >= 5;
}

std::vector<KmerStat> get_kmer_stats(const std::vector<Cand>& recs, uint8_t dir, const Config& cfg) {  // main.rs:408-455
  std::vector<std::string> primers; for (auto& r : recs) primers.push_back(r.word);
  auto infos = check_primers(primers, cfg.min_tm, cfg.max_tm);
  std::unordered_map<std::string, const PrimerInfo*> mp;
  for (auto& i : infos) mp.emplace(i.id, &i);
  float mean = NAN, sd = NAN;
  if (!infos.empty()) get_tm_stat(infos, &mean, &sd);
  std::vector<KmerStat> out;
  for (auto& r : recs) {
    const PrimerInfo* pi = mp.at(r.word);
    KmerStat s{r.word, dir, pi->gc, mean, sd, pi->tm, std::fabs(pi->tm - mean) <= (cfg.tm_stddev * sd),
               pi->self_any_th, pi->self_end_th, pi->hairpin_th, is_run(r.word)};
    out.push_back(s);
  }
  return out;
}

std::vector<KmerStat> filter_kmers(const std::vector<KmerStat>& st, const Config& c) {   // main.rs:492-516
  std::vector<KmerStat> out;
  for (auto& k : st) {
    bool a = !c.check_self_dimers || (k.self_any_th < c.max_self_dimer_any_tm);
    bool e = !c.check_self_dimers || (k.self_end_th < c.max_self_dimer_end_tm);
    bool h = !c.check_hairpin || (k.hairpin_th < c.max_hairpin_tm);
    bool mm = c.disable_min_max_tm || (k.tm > c.min_tm && k.tm < c.max_tm);
    bool sd = c.disable_tm_stddev || k.tm_ok;
    if (a && e && h && mm && sd && !k.runs) out.push_back(k);
  }
  return out;
}

// delta_g.rs:61-81
std::string format_ntthal_input(const std::vector<std::string>& primers, const Config& c) {
  std::string out;
  for (auto& a : primers) for (auto& b : primers) {
    if (!c.check_self_dimers && (a == b || reverse_complement(b) == a)) continue;
    if (!c.check_cross_dimers) continue;
    out += a; out += ','; out += b; out += '\n';
  }
  while (!out.empty() && isspace((unsigned char)out.back())) out.pop_back();   // .trim()
  size_t s = 0; while (s < out.size() && isspace((unsigned char)out[s])) s++;
  return out.substr(s);
}

std::vector<std::string> lines_of(const std::string& s) {                   // str::lines()
  std::vector<std::string> v; size_t p = 0;
  while (p < s.size()) { size_t e = s.find('\n', p); if (e == std::string::npos) e = s.size();
    std::string l = s.substr(p, e - p); if (!l.empty() && l.back() == '\r') l.pop_back(); v.push_back(l); p = e + 1; }
  return v;
}

// `ntthal -a ANY -mv .. -dv .. -n .. -d .. -t .. -i` emulation: per input line the 5-line block, or nothing at all
// for a pair without structure (delta_g.rs:93-113; that is what the reference's own executable prints, see
// tests/golden/ntthal_emulated.json).  Only line 0 carries numbers; SEQ/STR lines are placeholders (< 14 tokens).
std::string run_ntthal_text(const std::string& input, const Config& c, uint64_t* n_pairs) {
  auto r2 = [](float v) { char b[64]; snprintf(b, sizeof b, "%.2f", (double)v); return atof(b); };  // {:.2} argv
  msspe_thal_cond cond{r2(c.mv_conc), r2(c.dv_conc), r2(c.dntp_conc), r2(c.dna_conc), r2(c.annealing_temp), 30, 0};
  std::string out;
  if (const char* exe = getenv("ORACLE_NTTHAL")) {  // argv of delta_g.rs:93-110 (-path only when ./primer3_config/ exists)
    char b[512];
    snprintf(b, sizeof b, "%s -a ANY -mv %.2f -dv %.2f -n %.2f -d %.2f -t %.2f %s-i", exe, (double)c.mv_conc, (double)c.dv_conc, (double)c.dntp_conc,
             (double)c.dna_conc, (double)c.annealing_temp, access("primer3_config", R_OK) == 0 ? "-path primer3_config/ " : "");
    for (auto& l : lines_of(input)) { (void)l; (*n_pairs)++; }
    return run_tool(b, input + "\n");
  }
  for (auto& l : lines_of(input)) {
    size_t comma = l.find(',');
    std::string a = l.substr(0, comma), b = l.substr(comma + 1);
    msspe_thal_out o; oracle_thal(a.c_str(), b.c_str(), MSSPE_THAL_ANY, &cond, &o);
    (*n_pairs)++;
    if (o.no_structure) continue;
    char buf[256];
    snprintf(buf, sizeof buf, "Calculated thermodynamical parameters for dimer:\tdS = %g\tdH = %g\tdG = %g\tt = %g\n",
             o.ds, o.dh, o.dg, o.tm);
    out += buf; out += "SEQ\t\nSEQ\t\nSTR\t\nSTR\t\n";
  }
  return out;
}

struct Edge { std::string a, b; float dg; };
// delta_g.rs:27-59 (faithful, including the unconditional 4-line skip)
std::vector<Edge> parse_ntthal_output(const std::string& input, const std::string& output, float thr) {
  std::vector<Edge> edges; std::set<std::string> ids;
  auto ol = lines_of(output); size_t op = 0;
  for (auto& il : lines_of(input)) {
    if (op < ol.size()) {
      const std::string& line = ol[op++];
      // split_whitespace().nth(13)
      std::vector<std::string> tok; size_t p = 0;
      while (p < line.size()) { while (p < line.size() && isspace((unsigned char)line[p])) p++;
        size_t e = p; while (e < line.size() && !isspace((unsigned char)line[e])) e++;
        if (e > p) tok.push_back(line.substr(p, e - p)); p = e; }
      if (tok.size() > 13) {
        float dg = parse_f32(tok[13].c_str());
        if (dg < thr) {
          size_t comma = il.find(',');
          std::string a = il.substr(0, comma), b = il.substr(comma + 1);
          std::string id = a + ":" + b;
          if (ids.insert(id).second) {                 // HashSet<Edge> keyed by id: first insert wins
            char t[64]; snprintf(t, sizeof t, "%.2f", (double)dg);    // attrs "dg" = format!("{:.2}")
            edges.push_back({a, b, parse_f32(t)});
          }
        }
      }
    }
    op += 4;                                           // output_lines.nth(3)
  }
  return edges;
}

struct Result {
  std::vector<Cand> cand[2]; std::vector<KmerStat> stats[2], filtered[2], final_[2];
  std::string csv, report; uint64_t evals[2] = {0, 0}; uint64_t n_pairs = 0; uint64_t n_segments = 0;
  double t_segments = 0, t_select[2] = {0, 0}, t_thermo = 0, t_dimer = 0;
};

double now_s() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

std::string coverage_report(const std::vector<KmerStat>& f, const std::vector<KmerStat>& r, const Manager& m) {  // main.rs:518-594
  std::unordered_set<std::string> sf, sr; for (auto& p : f) sf.insert(p.word); for (auto& p : r) sr.insert(p.word);
  std::unordered_set<size_t> covered;
  for (auto& s : m.segments) {
    bool fh = false, rh = false;
    for (auto& k : s.kmers[0]) if (sf.count(k.word)) { fh = true; break; }
    for (auto& k : s.kmers[1]) if (sr.count(k.word)) { rh = true; break; }
    if (fh || rh) covered.insert(s.index);
  }
  size_t total = m.segments.size();
  std::unordered_map<std::string, std::pair<size_t, size_t>> seq; std::map<uint16_t, std::pair<size_t, size_t>> part;
  for (auto& s : m.segments) {
    auto& se = seq[m.records[s.record].name]; se.second++;
    auto& pe = part[s.partition_no]; pe.second++;
    if (covered.count(s.index)) { se.first++; pe.first++; }
  }
  float mn = INFINITY, mx = -INFINITY; size_t well = 0;
  for (auto& kv : seq) { float c = (float)kv.second.first / (float)kv.second.second * 100.0f;
    mn = std::fmin(mn, c); mx = std::fmax(mx, c); if (c >= 80.0f) well++; }
  std::string out = "\nCoverage report:\n";
  char b[512];
  snprintf(b, sizeof b, "  Segments:  %zu/%zu covered (%s%%)\n", covered.size(), total,
           fmt1(100.0f * (float)covered.size() / (float)total).c_str()); out += b;
  snprintf(b, sizeof b, "  Sequences: %zu/%zu at \xE2\x89\xA5" "80%% coverage (min %s%%, max %s%%)\n", well, seq.size(),
           fmt1(mn).c_str(), fmt1(mx).c_str()); out += b;
  std::string unc; bool any = false;
  for (auto& kv : part) if (kv.second.first == 0) { if (any) unc += ", "; unc += std::to_string(kv.first); any = true; }
  if (!any) out += "  All partitions have primer coverage\n"; else out += "  Uncovered partitions: [" + unc + "]\n";
  return out;
}

Result* run_pipeline(const char* fasta, size_t len, Config cfg, int stop_after /*0 all,1 select,*/) {
  Result* R = new Result();
  Manager m; m.records = to_records(fasta, len);
  if (m.records.empty()) { delete R; return nullptr; }
  if (cfg.max_mismatch_segments == 0) {                                       // main.rs:658-660
    size_t n = m.records.size(); size_t v = (n + 49) / 50; cfg.max_mismatch_segments = std::min<size_t>(10, std::max<size_t>(1, v));
  }
  double t0 = now_s();
  get_segment_manager(m, cfg.window_size, cfg.overlap_size, cfg.search_windows_size, cfg.kmer_size);
  R->t_segments = now_s() - t0; R->n_segments = m.segments.size();
  for (int d = 0; d < 2; d++) {
    t0 = now_s();
    R->cand[d] = find_candidates_kmers(m, (uint8_t)d, cfg.max_iterations, cfg.max_mismatch_segments, &R->evals[d]);
    R->t_select[d] = now_s() - t0;
  }
  if (stop_after == 1) return R;
  t0 = now_s();
  for (int d = 0; d < 2; d++) {
    R->stats[d] = get_kmer_stats(R->cand[d], (uint8_t)d, cfg);
    R->filtered[d] = cfg.keep_all ? R->stats[d] : filter_kmers(R->stats[d], cfg);
  }
  R->t_thermo = now_s() - t0;
  std::vector<std::string> primers;
  for (int d = 0; d < 2; d++) for (auto& s : R->filtered[d]) primers.push_back(s.word);   // main.rs:739-743
  t0 = now_s();
  std::string input = format_ntthal_input(primers, cfg);
  std::string output = run_ntthal_text(input, cfg, &R->n_pairs);
  auto edges = parse_ntthal_output(input, output, cfg.delta_g_threshold);
  R->t_dimer = now_s() - t0;
  // conflicts (main.rs:754-771): every stored edge with dg < threshold links a<->b
  std::map<std::string, std::set<std::string>> conflicts;
  std::set<std::string> primer_set(primers.begin(), primers.end());
  for (auto& e : edges) if (e.dg < cfg.delta_g_threshold) { conflicts[e.a].insert(e.b); conflicts[e.b].insert(e.a); }
  std::set<std::string> deleted;                                             // main.rs:776-798
  for (;;) {
    bool have = false; std::string worst; size_t wc = 0;
    for (auto& kv : conflicts) {
      if (deleted.count(kv.first)) continue;
      size_t active = 0; for (auto& n : kv.second) if (!deleted.count(n)) active++;
      if (active == 0) continue;
      if (!have || active > wc || (active == wc && kv.first > worst)) { have = true; worst = kv.first; wc = active; }
    }
    if (!have) break;
    deleted.insert(worst);
  }
  for (int d = 0; d < 2; d++) {
    if (cfg.keep_all) R->final_[d] = R->filtered[d];
    else for (auto& p : R->filtered[d]) if (!deleted.count(p.word)) R->final_[d].push_back(p);
  }
  R->report = coverage_report(R->final_[0], R->final_[1], m);
  R->csv = "direction,name,primers,gc,avg,std,tm\n";                         // main.rs:836-857
  for (int d = 0; d < 2; d++) {
    size_t idx = 0;
    for (auto& p : R->final_[d]) {
      const char* dir = p.direction == 0 ? "F" : "R";
      R->csv += std::string(dir) + ",Primer_" + std::to_string(idx) + "_" + dir + "," + p.word + "," +
                fmt2(p.gc_percent / 100.0f) + "," + fmt2(p.mean) + "," + fmt2(p.std) + "," + fmt2(p.tm) + "\n";
      idx++;
    }
  }
  return R;
}

}  // namespace

extern "C" {

typedef Config oracle_config;

void* oracle_pipeline_run(const char* fasta, uint64_t len, const oracle_config* cfg, int stop_after) {
  return run_pipeline(fasta, len, *cfg, stop_after);
}
void oracle_pipeline_free(void* r) { delete (Result*)r; }
const char* oracle_pipeline_csv(void* r) { return ((Result*)r)->csv.c_str(); }
const char* oracle_pipeline_report(void* r) { return ((Result*)r)->report.c_str(); }
uint64_t oracle_pipeline_n_candidates(void* r, int dir) { return ((Result*)r)->cand[dir].size(); }
// stage: 0 candidates, 1 filtered, 2 final
uint64_t oracle_pipeline_count(void* r, int dir, int stage) {
  Result* R = (Result*)r;
  return stage == 0 ? R->cand[dir].size() : stage == 1 ? R->filtered[dir].size() : R->final_[dir].size();
}
const char* oracle_pipeline_word(void* r, int dir, int stage, uint64_t i) {
  Result* R = (Result*)r;
  return stage == 0 ? R->cand[dir][i].word.c_str() : stage == 1 ? R->filtered[dir][i].word.c_str() : R->final_[dir][i].word.c_str();
}
void oracle_pipeline_candidate(void* r, int dir, uint64_t i, uint64_t* freq, uint64_t* n_tied, float* score) {
  Result* R = (Result*)r; *freq = R->cand[dir][i].freq; *n_tied = R->cand[dir][i].n_tied; *score = R->cand[dir][i].score;
}
void oracle_pipeline_stat(void* r, int dir, uint64_t i, float* v /*tm,gc,any,end,hp,mean,std*/, int* flags /*tm_ok,runs*/) {
  auto& s = ((Result*)r)->stats[dir][i];
  v[0] = s.tm; v[1] = s.gc_percent; v[2] = s.self_any_th; v[3] = s.self_end_th; v[4] = s.hairpin_th; v[5] = s.mean; v[6] = s.std;
  flags[0] = s.tm_ok; flags[1] = s.runs;
}
void oracle_pipeline_timing(void* r, double* t /*segments, sel0, sel1, thermo, dimer*/, uint64_t* c /*evals0, evals1, pairs, segments*/) {
  Result* R = (Result*)r;
  t[0] = R->t_segments; t[1] = R->t_select[0]; t[2] = R->t_select[1]; t[3] = R->t_thermo; t[4] = R->t_dimer;
  c[0] = R->evals[0]; c[1] = R->evals[1]; c[2] = R->n_pairs; c[3] = R->n_segments;
}

// Dense slot table of Segment.kmers[dir] for K1 parity: codes[seg*slots + q] = 2-bit code of the k-mer that
// starts at position q of the head (dir 0) / tail (dir 1, reverse-complemented) search window if it is valid
// and the first occurrence inside that window, else UINT64_MAX.  Returns number of segments.
uint64_t oracle_segment_slots(const char* fasta, uint64_t len, uint64_t W, uint64_t S, uint64_t w, uint64_t k, int dir,
                              uint64_t* codes, uint64_t capacity, uint16_t* partition_no) {
  auto recs = to_records(fasta, len);
  uint64_t slots = w >= k ? w - k + 1 : 0, g = 0;
  auto enc = [&](const std::string& s) { uint64_t v = 0; for (char c : s) v = (v << 2) | (uint64_t)(strchr("ACGT", c) - "ACGT"); return v; };
  for (auto& r : recs) {
    auto parts = partitioning_sequence(r.sequence, W, S);
    for (size_t j = 0; j < parts.size(); j++, g++) {
      if ((g + 1) * slots > capacity) continue;
      std::string win = dir == 0 ? parts[j].substr(0, w) : parts[j].substr(parts[j].size() - w);
      std::unordered_set<std::string> seen;
      for (uint64_t q = 0; q < slots; q++) {
        std::string x = win.substr(q, k); bool ok = true;
        for (char c : x) if (!strchr("ATCGU", c)) { ok = false; break; }
        uint64_t code = UINT64_MAX;
        if (ok && seen.insert(x).second) code = enc(dir == 0 ? x : reverse_complement(x));
        codes[g * slots + q] = code;
      }
      if (partition_no) partition_no[g] = (uint16_t)j;
    }
  }
  return g;
}

// Greedy selection only (for select parity / CPU baseline of the k-mer stage).
uint64_t oracle_select(const char* fasta, uint64_t len, uint64_t W, uint64_t S, uint64_t w, uint64_t k, int dir,
                       uint64_t max_iter, uint64_t mms, uint64_t* codes, uint32_t* freqs, uint32_t* n_tied, float* scores,
                       uint64_t capacity, uint64_t* evals, double* seconds) {
  Manager m; m.records = to_records(fasta, len);
  get_segment_manager(m, W, S, w, k);
  uint64_t ev = 0; double t0 = now_s();
  auto c = find_candidates_kmers(m, (uint8_t)dir, max_iter, mms, &ev);
  if (seconds) *seconds = now_s() - t0;
  if (evals) *evals = ev;
  for (size_t i = 0; i < c.size() && i < capacity; i++) {
    uint64_t v = 0; for (char ch : c[i].word) v = (v << 2) | (uint64_t)(strchr("ACGT", ch) - "ACGT");
    codes[i] = v; freqs[i] = (uint32_t)c[i].freq; n_tied[i] = (uint32_t)c[i].n_tied; scores[i] = c[i].score;
  }
  return c.size();
}

// A segment manager kept across calls (bench.py's CPU arm: build once, untimed -- the GPU arm's genomes are resident too --
// then time find_candidates_kmers alone on the complete alignment for a bounded number of iterations).
struct KeptManager { Manager m; Mapping mp; };
void* oracle_manager_create(const char* fasta, uint64_t len, uint64_t W, uint64_t S, uint64_t w, uint64_t k) {
  KeptManager* km = new KeptManager();
  km->m.records = to_records(fasta, len);
  get_segment_manager(km->m, W, S, w, k);
  km->mp = make_mapping(km->m.segments);      // main.rs:337, once for all later calls (both directions share the map, as in the reference)
  return km;
}
uint64_t oracle_manager_segments(void* h) { return ((KeptManager*)h)->m.segments.size(); }
uint64_t oracle_manager_select(void* h, int dir, uint64_t max_iter, uint64_t mms, uint64_t* codes, uint32_t* freqs, uint64_t capacity,
                               uint64_t* evals, double* seconds) {
  KeptManager& km = *(KeptManager*)h;
  uint64_t ev = 0; double t0 = now_s();
  auto c = find_candidates_kmers(km.m, (uint8_t)dir, max_iter, mms, &ev, &km.mp);
  if (seconds) *seconds = now_s() - t0;
  if (evals) *evals = ev;
  for (size_t i = 0; i < c.size() && i < capacity; i++) {
    uint64_t v = 0; for (char ch : c[i].word) v = (v << 2) | (uint64_t)(strchr("ACGT", ch) - "ACGT");
    codes[i] = v; freqs[i] = (uint32_t)c[i].freq;
  }
  return c.size();
}
void oracle_manager_free(void* h) { delete (KeptManager*)h; }

}  // extern "C"
