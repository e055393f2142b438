// ntthal (B200 engine) -- SURVEY section 8 row f-4 / plugin seam #1 of section 8b: an executable that speaks the
// protocol od-msspe uses for `--ntthal <path>` (od-msspe/src/delta_g.rs:83-153), so that the UNMODIFIED Rust binary
// runs its cross-dimer stage on the GPU:
//   argv   -a ANY|END1 -mv x -dv x -n x -d x -t x [-maxloop n] [-path dir] -i        (delta_g.rs:93-110)
//          or a single pair: ... -s1 SEQ -s2 SEQ
//   stdin  one "a,b" line per ordered pair                                               (delta_g.rs:61-81)
//   stdout per pair either the 5-line block
//            Calculated thermodynamical parameters for dimer:\tdS = %g\tdH = %g\tdG = %g\tt = %g
//            SEQ\t..  SEQ\t..  STR\t..  STR\t..        (the drawn duplex)
//          or, for a pair without any structure, NOTHING: that is what the reference's Primer3 2.6.1 executable does
//          (its stdout under tools/a64emu, tests/golden/ntthal_emulated.json); the "No secondary structure could be
//          calculated" string in that binary belongs to HAIRPIN mode and goes to stderr.  delta_g.rs:27-59 reads these.
// All pairs of the run go to the device in one batch (msspe_thal_pairs_aligned); this file only parses text and draws.
// There is no CPU fallback: without a usable CUDA device the program exits with an error.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "../../include/od_msspe_b200.h"

namespace {

[[noreturn]] void die(const std::string& m) { std::cerr << "Error: " << m << "\n"; exit(1); }

bool encode(const std::string& w, uint64_t* code) {
  uint64_t c = 0;
  for (char ch : w) {
    int v;
    switch (ch) { case 'A': case 'a': v = 0; break; case 'C': case 'c': v = 1; break; case 'G': case 'g': v = 2; break;
                  case 'T': case 't': case 'U': case 'u': v = 3; break; default: return false; }
    c = (c << 2) | (uint64_t)v;
  }
  *code = c;
  return true;
}

// The duplex as ntthal draws it.  pairing[i] = 1-based position in the REVERSED second oligo that base i of the first
// one pairs with (0 = unpaired).  Rows: unpaired / paired bases of oligo 1, paired / unpaired bases of oligo 2 (3'->5').
// Left ends are padded with blanks, loops and right ends of the shorter strand with '-'.
void draw(const std::string& o1, const std::string& o2, const uint8_t* pairing, std::string rows[4]) {
  const std::string r2(o2.rbegin(), o2.rend());
  const int n1 = (int)o1.size(), n2 = (int)r2.size();
  std::vector<std::pair<int, int>> bp;  // (i, j) 0-based, ascending in both
  for (int i = 0; i < n1; i++) if (pairing[i]) bp.push_back({i, (int)pairing[i] - 1});
  for (auto& r : {0, 1, 2, 3}) rows[r].clear();
  int i = 0, j = 0;
  for (size_t t = 0; t < bp.size(); t++) {
    const int u1 = bp[t].first - i, u2 = bp[t].second - j, L = u1 > u2 ? u1 : u2;
    if (t == 0) {  // left ends: blanks in front of the shorter one
      rows[0] += std::string(L - u1, ' ') + o1.substr(i, u1);
      rows[3] += std::string(L - u2, ' ') + r2.substr(j, u2);
    } else {       // loop: '-' behind the shorter side
      rows[0] += o1.substr(i, u1) + std::string(L - u1, '-');
      rows[3] += r2.substr(j, u2) + std::string(L - u2, '-');
    }
    rows[1] += std::string(L, ' '); rows[2] += std::string(L, ' ');
    rows[0] += ' '; rows[3] += ' ';
    rows[1] += o1[bp[t].first]; rows[2] += r2[bp[t].second];
    i = bp[t].first + 1; j = bp[t].second + 1;
  }
  const int t1 = n1 - i, t2 = n2 - j, L = t1 > t2 ? t1 : t2;
  rows[0] += o1.substr(i, t1) + std::string(L - t1, '-');
  rows[3] += r2.substr(j, t2) + std::string(L - t2, '-');
  rows[1] += std::string(L, ' '); rows[2] += std::string(L, ' ');   // ntthal pads the two middle rows to the full width
}

}  // namespace

int main(int argc, char** argv) {
  std::string mode = "ANY", path, s1, s2;
  msspe_thal_cond cond{50.0, 0.0, 0.8, 50.0, 37.0, 30, 0};  // ntthal's own defaults
  bool interactive = false;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto val = [&](const char* what) -> std::string { if (i + 1 >= argc) die(std::string("missing value for ") + what); return argv[++i]; };
    if (a == "-a") mode = val("-a");
    else if (a == "-mv") cond.mv = atof(val("-mv").c_str());
    else if (a == "-dv") cond.dv = atof(val("-dv").c_str());
    else if (a == "-n") cond.dntp = atof(val("-n").c_str());
    else if (a == "-d") cond.dna_conc = atof(val("-d").c_str());
    else if (a == "-t") cond.temp_c = atof(val("-t").c_str());
    else if (a == "-maxloop") cond.max_loop = atoi(val("-maxloop").c_str());
    else if (a == "-path") path = val("-path");
    else if (a == "-s1") s1 = val("-s1");
    else if (a == "-s2") s2 = val("-s2");
    else if (a == "-i") interactive = true;
    else if (a == "-r") { /* "print only the numbers": the reference never passes it */ }
    else die("unknown option " + a);
  }
  int type;
  if (mode == "ANY") type = MSSPE_THAL_ANY;
  else if (mode == "END1") type = MSSPE_THAL_END1;
  else die("alignment type " + mode + " is not supported by the B200 engine (ANY, END1)");
  std::vector<std::pair<std::string, std::string>> lines;
  if (interactive) {
    std::string l;
    while (std::getline(std::cin, l)) {
      if (!l.empty() && l.back() == '\r') l.pop_back();
      if (l.empty()) continue;
      const size_t c = l.find(',');
      if (c == std::string::npos) die("input line without ',': " + l);
      lines.push_back({l.substr(0, c), l.substr(c + 1)});
    }
  } else {
    if (s1.empty() || s2.empty()) die("give -s1 and -s2, or -i");
    lines.push_back({s1, s2});
  }
  if (lines.empty()) return 0;
  msspe_ctx* ctx = nullptr;
  msspe_config cfg{13, 500, 250, 50, getenv("MSSPE_DEVICE") ? atoi(getenv("MSSPE_DEVICE")) : 0, 0};
  if (int rc = msspe_create(&cfg, &ctx)) die(std::string("cannot start the GPU engine (") + std::to_string(rc) + "): " + msspe_last_error(nullptr));
  if (!path.empty()) {
    msspe_thal_raw_params* p = new msspe_thal_raw_params();
    char err[256] = {0};
    if (msspe_thal_params_from_dir(path.c_str(), p, err, sizeof err) != MSSPE_OK) die(std::string("-path ") + path + ": " + err);
    if (msspe_set_thal_params(ctx, p) != MSSPE_OK) die(msspe_last_error(ctx));
    delete p;
  }
  // one device batch per oligo length (the reference sends k-mers of one length)
  std::map<size_t, std::vector<size_t>> by_len;
  for (size_t t = 0; t < lines.size(); t++) {
    if (lines[t].first.size() != lines[t].second.size() || lines[t].first.empty() || lines[t].first.size() > MSSPE_MAX_OLIGO)
      die("the B200 engine aligns oligos of equal length up to 32 nt: " + lines[t].first + "," + lines[t].second);
    by_len[lines[t].first.size()].push_back(t);
  }
  std::vector<msspe_thal_out> out(lines.size());
  std::vector<uint8_t> pairing(lines.size() * (size_t)MSSPE_MAX_OLIGO);
  for (auto& kv : by_len) {
    const size_t m = kv.second.size();
    std::vector<uint64_t> a(m), b(m);
    for (size_t q = 0; q < m; q++)
      if (!encode(lines[kv.second[q]].first, &a[q]) || !encode(lines[kv.second[q]].second, &b[q]))
        die("only A, C, G, T oligos: " + lines[kv.second[q]].first + "," + lines[kv.second[q]].second);
    std::vector<msspe_thal_out> o(m);
    std::vector<uint8_t> pr(m * (size_t)MSSPE_MAX_OLIGO);
    if (msspe_thal_pairs_aligned(ctx, a.data(), b.data(), m, (uint32_t)kv.first, type, &cond, o.data(), pr.data()) != MSSPE_OK) die(msspe_last_error(ctx));
    for (size_t q = 0; q < m; q++) {
      out[kv.second[q]] = o[q];
      memcpy(&pairing[kv.second[q] * (size_t)MSSPE_MAX_OLIGO], &pr[q * (size_t)MSSPE_MAX_OLIGO], MSSPE_MAX_OLIGO);
    }
  }
  std::string text;
  char buf[512];
  for (size_t t = 0; t < lines.size(); t++) {
    if (out[t].no_structure) continue;   // the reference executable prints nothing for such a pair
    snprintf(buf, sizeof buf, "Calculated thermodynamical parameters for dimer:\tdS = %g\tdH = %g\tdG = %g\tt = %g\n", out[t].ds, out[t].dh,
             out[t].dg, out[t].tm);
    text += buf;
    std::string rows[4];
    draw(lines[t].first, lines[t].second, &pairing[t * (size_t)MSSPE_MAX_OLIGO], rows);
    text += "SEQ\t" + rows[0] + "\nSEQ\t" + rows[1] + "\nSTR\t" + rows[2] + "\nSTR\t" + rows[3] + "\n";
  }
  fwrite(text.data(), 1, text.size(), stdout);
  msspe_destroy(ctx);
  return 0;
}
