// thal_params.cu -- parameter sources for the thermodynamic kernels (host side).
// Replaces `ntthal -path <cwd>/primer3_config/` (od-msspe/src/delta_g.rs:90,107-108): either the tables
// embedded at build time (generated from that directory by tools/gen_thal_params.py) or a directory read at
// run time; then the expansion into the 5-symbol (A,C,G,T,N) tables the kernels index.
#include <cmath>
#include <cstdlib>
#include <algorithm>

#include "engine.cuh"
#include "thal_tables.cuh"

static const msspe_thal_raw_params kEmbeddedParams =
#include "thal_params_data.inc"
    ;

extern "C" int msspe_thal_params_default(msspe_thal_raw_params* out) {
  if (!out) return MSSPE_ERR_INVALID;
  *out = kEmbeddedParams;
  return MSSPE_OK;
}

namespace {
bool read_values(const std::string& path, double* dst, int n, int skip_first_col, int cols) {
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return false;
  char tok[64];
  int got = 0, col = 0;
  while (got < n && fscanf(f, "%63s", tok) == 1) {
    if (skip_first_col && col == 0) { col = (col + 1) % cols; continue; }
    dst[got++] = strcmp(tok, "inf") == 0 ? (double)INFINITY : atof(tok);
    col = (col + 1) % cols;
  }
  fclose(f);
  return got == n;
}
int read_keyed(const std::string& path, char seqs[][8], double* vals, int cap) {
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return -1;
  char s[64], v[64];
  int n = 0;
  while (n < cap && fscanf(f, "%63s %63s", s, v) == 2) {
    memset(seqs[n], 0, 8);
    memcpy(seqs[n], s, strnlen(s, 7));
    vals[n] = strcmp(v, "inf") == 0 ? (double)INFINITY : atof(v);
    n++;
  }
  fclose(f);
  return n;
}
}  // namespace

extern "C" int msspe_thal_params_from_dir(const char* dir, msspe_thal_raw_params* p, char* err, size_t err_len) {
  if (!dir || !p) return MSSPE_ERR_INVALID;
  memset(p, 0, sizeof *p);
  std::string d(dir);
  if (!d.empty() && d.back() != '/') d += '/';
  struct { const char* fn; double* dst; int n; int skip; int cols; } plain[] = {
      {"stack.ds", p->stack_ds, 256, 0, 1},      {"stack.dh", p->stack_dh, 256, 0, 1},
      {"stackmm.ds", p->stackmm_ds, 256, 0, 1},  {"stackmm.dh", p->stackmm_dh, 256, 0, 1},
      {"dangle.ds", p->dangle_ds, 128, 0, 1},    {"dangle.dh", p->dangle_dh, 128, 0, 1},
      {"loops.ds", p->loops_ds, 90, 1, 4},       {"loops.dh", p->loops_dh, 90, 1, 4},
      {"tstack_tm_inf.ds", p->tstack_ds, 256, 0, 1}, {"tstack.dh", p->tstack_dh, 256, 0, 1},
      {"tstack2.ds", p->tstack2_ds, 256, 0, 1},  {"tstack2.dh", p->tstack2_dh, 256, 0, 1}};
  for (auto& t : plain)
    if (!read_values(d + t.fn, t.dst, t.n, t.skip, t.cols)) {
      if (err && err_len) snprintf(err, err_len, "cannot read %d values from %s%s", t.n, d.c_str(), t.fn);
      return MSSPE_ERR_IO;
    }
  p->n_triloop_ds = read_keyed(d + "triloop.ds", p->triloop_ds_seq, p->triloop_ds, 32);
  p->n_triloop_dh = read_keyed(d + "triloop.dh", p->triloop_dh_seq, p->triloop_dh, 32);
  p->n_tetraloop_ds = read_keyed(d + "tetraloop.ds", p->tetraloop_ds_seq, p->tetraloop_ds, 128);
  p->n_tetraloop_dh = read_keyed(d + "tetraloop.dh", p->tetraloop_dh_seq, p->tetraloop_dh, 128);
  if (p->n_triloop_ds < 0 || p->n_triloop_dh < 0 || p->n_tetraloop_ds < 0 || p->n_tetraloop_dh < 0) {
    if (err && err_len) snprintf(err, err_len, "cannot read triloop/tetraloop tables from %s", d.c_str());
    return MSSPE_ERR_IO;
  }
  return MSSPE_OK;
}

namespace {
int bidx(char c) {
  switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2;
               case 'T': case 't': case 'U': case 'u': return 3; default: return 4; }
}
void expand4(const double* ds, const double* dh, double* S, double* H, bool terminal_style) {
  int n = 0;
  for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) for (int k = 0; k < 5; k++) for (int l = 0; l < 5; l++) {
    const int x = THAL_IDX4(i, j, k, l);
    if (!terminal_style) {
      if (i == 4 || j == 4 || k == 4 || l == 4) { S[x] = -1.0; H[x] = INFINITY; continue; }
    } else {
      if (i == 4 || k == 4) { S[x] = -1.0; H[x] = INFINITY; continue; }
      if (j == 4 || l == 4) { S[x] = 0.00000000001; H[x] = 0.0; continue; }
    }
    double s = ds[n], h = dh[n];
    n++;
    if (!std::isfinite(s) || !std::isfinite(h)) { s = -1.0; h = INFINITY; }
    S[x] = s; H[x] = h;
  }
}
void sort_keyed(int n, char seqs[][8], const double* vals, int len, uint32_t* keys, double* out) {
  std::vector<std::pair<uint32_t, double>> v;
  for (int i = 0; i < n; i++) {
    uint32_t key = 0;
    for (int c = 0; c < len; c++) key = key * 5u + (uint32_t)bidx(seqs[i][c]);
    v.push_back({key, vals[i]});
  }
  std::stable_sort(v.begin(), v.end(), [](const std::pair<uint32_t, double>& a, const std::pair<uint32_t, double>& b) { return a.first < b.first; });
  for (int i = 0; i < n; i++) { keys[i] = v[i].first; out[i] = v[i].second; }
}
}  // namespace

void msspe_thal_expand(const msspe_thal_raw_params* p_in, ThalDeviceTables* T) {
  msspe_thal_raw_params* p = const_cast<msspe_thal_raw_params*>(p_in);
  memset(T, 0, sizeof *T);
  expand4(p->stack_ds, p->stack_dh, T->stackS, T->stackH, false);
  expand4(p->stackmm_ds, p->stackmm_dh, T->stackint2S, T->stackint2H, false);
  expand4(p->tstack_ds, p->tstack_dh, T->tstackS, T->tstackH, true);
  expand4(p->tstack2_ds, p->tstack2_dh, T->tstack2S, T->tstack2H, true);
  int n = 0;  // 64 "3' dangling" lines: loops i,j,k -> dangle3[i][k][j]; then 64 "5' dangling" -> dangle5[i][j][k]
  for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) for (int k = 0; k < 5; k++) {
    const int x = THAL_IDX3(i, k, j);
    if (i == 4 || j == 4 || k == 4) { T->dangle3S[x] = -1.0; T->dangle3H[x] = INFINITY; continue; }
    double s = p->dangle_ds[n], h = p->dangle_dh[n];
    n++;
    if (!std::isfinite(s) || !std::isfinite(h)) { s = -1.0; h = INFINITY; }
    T->dangle3S[x] = s; T->dangle3H[x] = h;
  }
  for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) for (int k = 0; k < 5; k++) {
    const int x = THAL_IDX3(i, j, k);
    if (i == 4 || j == 4 || k == 4) { T->dangle5S[x] = -1.0; T->dangle5H[x] = INFINITY; continue; }
    double s = p->dangle_ds[n], h = p->dangle_dh[n];
    n++;
    if (!std::isfinite(s) || !std::isfinite(h)) { s = -1.0; h = INFINITY; }
    T->dangle5S[x] = s; T->dangle5H[x] = h;
  }
  for (int k = 0; k < 30; k++) {
    T->interiorS[k] = p->loops_ds[3 * k]; T->bulgeS[k] = p->loops_ds[3 * k + 1]; T->hairpinS[k] = p->loops_ds[3 * k + 2];
    T->interiorH[k] = p->loops_dh[3 * k]; T->bulgeH[k] = p->loops_dh[3 * k + 1]; T->hairpinH[k] = p->loops_dh[3 * k + 2];
  }
  for (int i = 0; i < 25; i++) { T->atpS[i] = 0.00000000001; T->atpH[i] = 0.0; }
  T->atpS[0 * 5 + 3] = T->atpS[3 * 5 + 0] = 6.9;
  T->atpH[0 * 5 + 3] = T->atpH[3 * 5 + 0] = 2200.0;
  T->nTriS = p->n_triloop_ds; T->nTriH = p->n_triloop_dh; T->nTetraS = p->n_tetraloop_ds; T->nTetraH = p->n_tetraloop_dh;
  sort_keyed(T->nTriS, p->triloop_ds_seq, p->triloop_ds, 5, T->triKeyS, T->triS);
  sort_keyed(T->nTriH, p->triloop_dh_seq, p->triloop_dh, 5, T->triKeyH, T->triH);
  sort_keyed(T->nTetraS, p->tetraloop_ds_seq, p->tetraloop_ds, 6, T->tetraKeyS, T->tetraS);
  sort_keyed(T->nTetraH, p->tetraloop_dh_seq, p->tetraloop_dh, 6, T->tetraKeyH, T->tetraH);
}

// Host-only diagnostic of the C ABI: one expanded table under the name libprimer3's thal.c gives it.  A CPU test compares
// every one of them with the arrays compiled into the reference's Primer3 2.6.1 executables
// (tests/golden/primer3_2_6_1_compiled_in_tables.json), so the tables the kernels index are pinned, not only the raw files.
extern "C" int msspe_thal_expanded_table(const msspe_thal_raw_params* p, const char* name, double* out, uint32_t cap) {
  if (!p || !name || !out) return MSSPE_ERR_INVALID;
  std::vector<ThalDeviceTables> holder(1);
  ThalDeviceTables* T = holder.data();
  msspe_thal_expand(p, T);
  const struct { const char* name; const double* v; int n; } plain[] = {
      {"stackEntropies", T->stackS, 625},         {"stackEnthalpies", T->stackH, 625},
      {"stackint2Entropies", T->stackint2S, 625}, {"stackint2Enthalpies", T->stackint2H, 625},
      {"tstackEntropies", T->tstackS, 625},       {"tstackEnthalpies", T->tstackH, 625},
      {"tstack2Entropies", T->tstack2S, 625},     {"tstack2Enthalpies", T->tstack2H, 625},
      {"dangleEntropies3", T->dangle3S, 125},     {"dangleEnthalpies3", T->dangle3H, 125},
      {"dangleEntropies5", T->dangle5S, 125},     {"dangleEnthalpies5", T->dangle5H, 125},
      {"hairpinLoopEntropies", T->hairpinS, 30},  {"interiorLoopEntropies", T->interiorS, 30},
      {"bulgeLoopEntropies", T->bulgeS, 30},      {"hairpinLoopEnthalpies", T->hairpinH, 30},
      {"interiorLoopEnthalpies", T->interiorH, 30}, {"bulgeLoopEnthalpies", T->bulgeH, 30},
      {"atpS", T->atpS, 25},                      {"atpH", T->atpH, 25}};
  for (const auto& t : plain)
    if (strcmp(name, t.name) == 0) {
      if (cap < (uint32_t)t.n) return MSSPE_ERR_CAPACITY;
      memcpy(out, t.v, sizeof(double) * t.n);
      return t.n;
    }
  // loop bonus tables: (key, value) pairs in table order, key = base-5 digits of the loop (5 resp. 6 bases)
  const struct { const char* name; const uint32_t* key; const double* v; int n; } keyed[] = {
      {"defaultTriloopEntropies", T->triKeyS, T->triS, T->nTriS},
      {"defaultTriloopEnthalpies", T->triKeyH, T->triH, T->nTriH},
      {"defaultTetraloopEntropies", T->tetraKeyS, T->tetraS, T->nTetraS},
      {"defaultTetraloopEnthalpies", T->tetraKeyH, T->tetraH, T->nTetraH}};
  for (const auto& t : keyed)
    if (strcmp(name, t.name) == 0) {
      if (t.n < 0 || cap < 2u * (uint32_t)t.n) return MSSPE_ERR_CAPACITY;
      for (int i = 0; i < t.n; i++) { out[2 * i] = (double)t.key[i]; out[2 * i + 1] = t.v[i]; }
      return 2 * t.n;
    }
  return MSSPE_ERR_INVALID;
}
