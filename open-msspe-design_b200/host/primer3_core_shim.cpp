// primer3_core (B200 engine) -- SURVEY section 8 row f-4 / plugin seam #1 of section 8b: an executable that speaks
// the Boulder-IO subset od-msspe uses for `--primer3 <path>` (od-msspe/src/primer.rs:67-166), so that the UNMODIFIED
// Rust binary gets Tm / GC / self-dimer / hairpin numbers from the GPU:
//   stdin   records of TAG=VALUE lines ended by "=":  SEQUENCE_ID, SEQUENCE_PRIMER, PRIMER_TASK=check_primers,
//           PRIMER_MIN_SIZE, PRIMER_MIN_TM, PRIMER_MAX_TM, PRIMER_OPT_TM, PRIMER_PICK_ANYWAY=1     (primer.rs:125-140)
//   stdout  per record the input tags echoed, then PRIMER_LEFT_NUM_RETURNED=1 ... PRIMER_LEFT_0_TM (%.3f),
//           _GC_PERCENT (%.3f), _SELF_ANY_TH, _SELF_END_TH, _HAIRPIN_TH (%.2f) and "="              (primer.rs:36-66)
// The numbers are Primer3's defaults for this task (no salt tags are sent, primer.rs:125-140): oligotm
// SantaLucia-1998 with the SantaLucia salt correction, thal ANY / END1 / HAIRPIN at 37 C, each max(0, Tm).
// PRIMER_LEFT_0_END_STABILITY is not produced (od-msspe does not read it).  No CPU fallback.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "../../include/od_msspe_b200.h"

namespace {
struct Rec { std::vector<std::pair<std::string, std::string>> tags; std::string primer, task; double opt_tm = 60.0, opt_size = 20.0; };
bool encode(const std::string& w, uint64_t* code) {
  uint64_t c = 0;
  for (char ch : w) {
    int v;
    switch (ch) { case 'A': case 'a': v = 0; break; case 'C': case 'c': v = 1; break; case 'G': case 'g': v = 2; break;
                  case 'T': case 't': v = 3; break; default: return false; }
    c = (c << 2) | (uint64_t)v;
  }
  *code = c;
  return true;
}
}  // namespace

int main(int argc, char** argv) {
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    if (a == "--version" || a == "-version") { std::cout << "libprimer3 release 2.6.1 (B200 engine, check_primers only)\n"; return 0; }
  }
  std::vector<Rec> recs;
  Rec cur; bool open = false;
  std::string l;
  while (std::getline(std::cin, l)) {
    if (!l.empty() && l.back() == '\r') l.pop_back();
    if (l == "=") { if (open) recs.push_back(cur); cur = Rec(); open = false; continue; }
    if (l.empty()) continue;
    const size_t e = l.find('=');
    if (e == std::string::npos) { std::cerr << "primer3_core: line without '=': " << l << "\n"; return 255; }
    const std::string k = l.substr(0, e), v = l.substr(e + 1);
    open = true;
    cur.tags.push_back({k, v});
    if (k == "SEQUENCE_PRIMER") cur.primer = v;
    else if (k == "PRIMER_TASK") cur.task = v;
    else if (k == "PRIMER_OPT_TM") cur.opt_tm = atof(v.c_str());
    else if (k == "PRIMER_OPT_SIZE") cur.opt_size = atof(v.c_str());
  }
  if (open) recs.push_back(cur);
  if (recs.empty()) return 0;
  msspe_ctx* ctx = nullptr;
  msspe_config cfg{13, 500, 250, 50, getenv("MSSPE_DEVICE") ? atoi(getenv("MSSPE_DEVICE")) : 0, 0};
  if (int rc = msspe_create(&cfg, &ctx)) { std::cerr << "primer3_core: cannot start the GPU engine (" << rc << "): " << msspe_last_error(nullptr) << "\n"; return 255; }
  const size_t n = recs.size();
  std::vector<double> tm(n), gc(n), sa(n), se(n), hp(n);
  std::vector<std::string> error(n);
  std::map<size_t, std::vector<size_t>> by_len;
  for (size_t i = 0; i < n; i++) {
    uint64_t code;
    if (recs[i].task != "check_primers") error[i] = "the B200 engine implements PRIMER_TASK=check_primers only";
    else if (recs[i].primer.empty() || recs[i].primer.size() > MSSPE_MAX_OLIGO || !encode(recs[i].primer, &code)) error[i] = "Missing or unsupported SEQUENCE_PRIMER";
    else by_len[recs[i].primer.size()].push_back(i);
  }
  for (auto& kv : by_len) {
    const size_t m = kv.second.size();
    std::vector<uint64_t> codes(m);
    for (size_t q = 0; q < m; q++) encode(recs[kv.second[q]].primer, &codes[q]);
    std::vector<double> a(m), b(m), c(m), d(m), e(m);
    if (msspe_primer_thermo(ctx, codes.data(), (uint32_t)m, (uint32_t)kv.first, a.data(), b.data(), c.data(), d.data(), e.data()) != MSSPE_OK) {
      std::cerr << "primer3_core: " << msspe_last_error(ctx) << "\n"; return 255;
    }
    for (size_t q = 0; q < m; q++) { const size_t i = kv.second[q]; tm[i] = a[q]; gc[i] = b[q]; sa[i] = c[q]; se[i] = d[q]; hp[i] = e[q]; }
  }
  std::string text;
  char buf[256];
  for (size_t i = 0; i < n; i++) {
    for (auto& t : recs[i].tags) text += t.first + "=" + t.second + "\n";
    if (!error[i].empty()) { text += "PRIMER_ERROR=" + error[i] + "\n=\n"; continue; }
    const double size = (double)recs[i].primer.size();
    const double penalty = std::fabs(tm[i] - recs[i].opt_tm) + std::fabs(size - recs[i].opt_size);  // WT_TM_* = WT_SIZE_* = 1, others 0
    snprintf(buf, sizeof buf,
             "PRIMER_LEFT_NUM_RETURNED=1\nPRIMER_RIGHT_NUM_RETURNED=0\nPRIMER_INTERNAL_NUM_RETURNED=0\nPRIMER_PAIR_NUM_RETURNED=0\n"
             "PRIMER_LEFT_0_PENALTY=%f\n", penalty);
    text += buf;
    text += "PRIMER_LEFT_0_SEQUENCE=" + recs[i].primer + "\n";
    snprintf(buf, sizeof buf,
             "PRIMER_LEFT_0=0,%d\nPRIMER_LEFT_0_TM=%.3f\nPRIMER_LEFT_0_GC_PERCENT=%.3f\nPRIMER_LEFT_0_SELF_ANY_TH=%.2f\n"
             "PRIMER_LEFT_0_SELF_END_TH=%.2f\nPRIMER_LEFT_0_HAIRPIN_TH=%.2f\n=\n",
             (int)size, tm[i], gc[i], sa[i], se[i], hp[i]);
    text += buf;
  }
  fwrite(text.data(), 1, text.size(), stdout);
  msspe_destroy(ctx);
  return 0;
}
