/*
 * thal_oracle.c -- TEST INFRASTRUCTURE ONLY (oracle).  Never linked into, imported by, or executed
 * from the product path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm
 * use it.
 *
 * CPU restatement (plain scalar C, FP64) of the third-party arithmetic the reference reaches through
 * its two subprocesses:
 *   - `ntthal -a ANY ...`      reference od-msspe/src/delta_g.rs:93-113   -> thal dimer (ANY / END1)
 *   - `primer3_core` check_primers, od-msspe/src/primer.rs:125-166        -> oligotm Tm, GC%, thal ANY/END1
 *                                                                              self-dimer, thal HAIRPIN
 * The algorithm lives in Primer3 libprimer3 2.6.1 (src/thal.c, src/oligotm.c), which is NOT in
 * /root/reference (only Mach-O arm64 binaries of it are: od-msspe/bin/ntthal, bin/primer3_core).  This
 * file restates the published algorithm (SantaLucia & Hicks 2004 nearest-neighbour model as organised by
 * Primer3's thermodynamic alignment) over the parameter tables the reference
 * vendors (od-msspe/primer3_config/ *.ds, *.dh).
 *
 * Pinning: dimer ANY reproduces the five real ntthal outputs the reference keeps at
 * delta_g.rs:197-230, and Tm/GC/self-any reproduce primer.rs:238-250 (tests/test_oracle_thermo.py).
 * The parameter tables AS LOADED (every stack / dangle / loop / terminal-stack array, the AT penalty, the sorted
 * tri- and tetraloop bonus tables) are bit-identical to the arrays compiled into the reference's own Primer3 2.6.1
 * executables, read out of their Mach-O data by symbol name (oracle_thal_table below,
 * tests/golden/primer3_2_6_1_compiled_in_tables.json).
 * And to REFERENCE OUTPUT: stdout of those executables themselves, interpreted instruction by instruction under
 * tools/a64emu: 526 ntthal invocations over ANY / END1 / END2 / HAIRPIN (tests/golden/ntthal_emulated.json,
 * tests/test_ntthal_emulated_golden.py) and 337 primer3_core check_primers records (tests/golden/
 * primer3_core_emulated.json, tests/test_primer3_core_emulated_golden.py), all equal as printed.  That pin replaced
 * the hairpin-closing test of older Primer3 releases (melting temperatures) with 2.6.1's (free energies with the
 * right-end term, calc_hairpin below), and showed that ntthal prints nothing for a structure-less dimer.
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction, so the CUDA kernels compiled with -fmad=false
 * can be compared bit-for-bit).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

#include "../include/od_msspe_b200.h"

#define ORACLE_MAX_LEN 60

static const double R_GAS = 1.9872;
static const double ABS_ZERO = 273.15;
static const double T_KELVIN = 310.15;
static const double MIN_ENTROPY_CUTOFF = -2500.0;
static const double MIN_ENTROPY = -3224.0;
static const double SMALL_NON_ZERO = 0.000001;
static const double AT_S = 6.9, AT_H = 2200.0;
static const double ILAS = (-300 / 310.15);
static const double ILAH = 0.0;
#define MIN_HRPN_LOOP 3

/* ---- expanded tables (index 4 = N, the sequence-end sentinel) ---- */
static double stackS[5][5][5][5], stackH[5][5][5][5];
static double stackint2S[5][5][5][5], stackint2H[5][5][5][5];
static double tstackS[5][5][5][5], tstackH[5][5][5][5];
static double tstack2S[5][5][5][5], tstack2H[5][5][5][5];
static double dangle3S[5][5][5], dangle3H[5][5][5];
static double dangle5S[5][5][5], dangle5H[5][5][5];
static double interiorS[30], interiorH[30], bulgeS[30], bulgeH[30], hairpinS[30], hairpinH[30];
static double atpS[5][5], atpH[5][5];
struct loopent { unsigned char key[6]; double value; };
static struct loopent triS[32], triH[32], tetraS[128], tetraH[128];
static int nTriS, nTriH, nTetraS, nTetraH;
static int params_loaded = 0;

static int base_idx(int c) {
  switch (toupper(c)) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 4; }
}
static int bp(int a, int b) { return a + b == 3 && a < 4 && b < 4; }

static void joint4(const double* ds, const double* dh, double S[5][5][5][5], double H[5][5][5][5], int tstack_style) {
  int n = 0;
  for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) for (int k = 0; k < 5; k++) for (int l = 0; l < 5; l++) {
    if (!tstack_style) {
      if (i == 4 || j == 4 || k == 4 || l == 4) { S[i][j][k][l] = -1.0; H[i][j][k][l] = INFINITY; continue; }
    } else {
      if (i == 4 || k == 4) { S[i][j][k][l] = -1.0; H[i][j][k][l] = INFINITY; continue; }
      if (j == 4 || l == 4) { S[i][j][k][l] = 0.00000000001; H[i][j][k][l] = 0.0; continue; }
    }
    double s = ds[n], h = dh[n]; n++;
    if (!isfinite(s) || !isfinite(h)) { s = -1.0; h = INFINITY; }
    S[i][j][k][l] = s; H[i][j][k][l] = h;
  }
}

static int cmp_loop5(const void* a, const void* b) { return memcmp(a, b, 5); }
static int cmp_loop6(const void* a, const void* b) { return memcmp(a, b, 6); }

static void load_loops(int n, char seqs[][8], const double* vals, struct loopent* dst, int len) {
  for (int i = 0; i < n; i++) {
    memset(dst[i].key, 0, 6);
    for (int c = 0; c < len; c++) dst[i].key[c] = (unsigned char)base_idx(seqs[i][c]);
    dst[i].value = vals[i];
  }
  qsort(dst, n, sizeof(struct loopent), len == 5 ? cmp_loop5 : cmp_loop6);
}

void oracle_thal_set_params(const msspe_thal_raw_params* p_in) {
  msspe_thal_raw_params* p = (msspe_thal_raw_params*)p_in;
  joint4(p->stack_ds, p->stack_dh, stackS, stackH, 0);
  joint4(p->stackmm_ds, p->stackmm_dh, stackint2S, stackint2H, 0);
  joint4(p->tstack_ds, p->tstack_dh, tstackS, tstackH, 1);
  joint4(p->tstack2_ds, p->tstack2_dh, tstack2S, tstack2H, 1);
  /* dangle: first 64 lines "3' dangling" in loops i,j,k -> dangle3[i][k][j]; next 64 -> dangle5[i][j][k] */
  int n = 0;
  for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) for (int k = 0; k < 5; k++) {
    if (i == 4 || j == 4 || k == 4) { dangle3S[i][k][j] = -1.0; dangle3H[i][k][j] = INFINITY; continue; }
    double s = p->dangle_ds[n], h = p->dangle_dh[n]; n++;
    if (!isfinite(s) || !isfinite(h)) { s = -1.0; h = INFINITY; }
    dangle3S[i][k][j] = s; dangle3H[i][k][j] = h;
  }
  for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) for (int k = 0; k < 5; k++) {
    if (i == 4 || j == 4 || k == 4) { dangle5S[i][j][k] = -1.0; dangle5H[i][j][k] = INFINITY; continue; }
    double s = p->dangle_ds[n], h = p->dangle_dh[n]; n++;
    if (!isfinite(s) || !isfinite(h)) { s = -1.0; h = INFINITY; }
    dangle5S[i][j][k] = s; dangle5H[i][j][k] = h;
  }
  for (int k = 0; k < 30; k++) {
    interiorS[k] = p->loops_ds[3 * k]; bulgeS[k] = p->loops_ds[3 * k + 1]; hairpinS[k] = p->loops_ds[3 * k + 2];
    interiorH[k] = p->loops_dh[3 * k]; bulgeH[k] = p->loops_dh[3 * k + 1]; hairpinH[k] = p->loops_dh[3 * k + 2];
  }
  for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) { atpS[i][j] = 0.00000000001; atpH[i][j] = 0.0; }
  atpS[0][3] = atpS[3][0] = AT_S; atpH[0][3] = atpH[3][0] = AT_H;
  nTriS = p->n_triloop_ds; nTriH = p->n_triloop_dh; nTetraS = p->n_tetraloop_ds; nTetraH = p->n_tetraloop_dh;
  load_loops(nTriS, p->triloop_ds_seq, p->triloop_ds, triS, 5);
  load_loops(nTriH, p->triloop_dh_seq, p->triloop_dh, triH, 5);
  load_loops(nTetraS, p->tetraloop_ds_seq, p->tetraloop_ds, tetraS, 6);
  load_loops(nTetraH, p->tetraloop_dh_seq, p->tetraloop_dh, tetraH, 6);
  params_loaded = 1;
}

/* ---- per-call state (single-threaded oracle; one global working set like the scalar original) ---- */
typedef struct {
  int n1[ORACLE_MAX_LEN + 2], n2[ORACLE_MAX_LEN + 2];
  int len1, len2;
  double Sm[ORACLE_MAX_LEN + 2][ORACLE_MAX_LEN + 2], Hm[ORACLE_MAX_LEN + 2][ORACLE_MAX_LEN + 2];
  double send5[ORACLE_MAX_LEN + 2], hend5[ORACLE_MAX_LEN + 2];
  double dHi, dSi, RC;
  int maxLoop;
  long stat_cells, stat_loopcand;
} work_t;

static int fin(double x) { return isfinite(x); }
static int equal2(double a, double b) { if (!isfinite(a) || !isfinite(b)) return 0; return fabs(a - b) < 1e-5; }

static int symmetry(const char* s, int len) {
  if (len % 2 == 1) return 0;
  for (int i = 0; i < len / 2; i++) {
    int a = toupper(s[i]), e = toupper(s[len - 1 - i]);
    if ((a == 'A' && e != 'T') || (a == 'T' && e != 'A') || (e == 'A' && a != 'T') || (e == 'T' && a != 'A')) return 0;
    if ((a == 'C' && e != 'G') || (a == 'G' && e != 'C') || (e == 'C' && a != 'G') || (e == 'G' && a != 'C')) return 0;
  }
  return 1;
}

static double salt_corr(double mv, double dv, double dntp) {
  if (dv <= 0) dntp = dv;
  return 0.368 * (log((mv + 120 * (sqrt(fmax(0.0, dv - dntp)))) / 1000));
}

/* Shared tail of the terminal (left/right) end evaluation: choose between terminal mismatch (S1,H1),
 * an optional dangling-end variant (S2,H2) and the bare AT penalty. */
static void end_term(const work_t* w, double S1, double H1, int has_opt, double So, double Ho, int a, int b,
                     double* outS, double* outH) {
  double G1 = H1 - T_KELVIN * S1, T1 = -INFINITY, S2, H2, G2, T2;
  if (!fin(H1) || G1 > 0) { H1 = INFINITY; S1 = -1.0; G1 = 1.0; }
  if (has_opt) {
    S2 = So; H2 = Ho; G2 = H2 - T_KELVIN * S2;
    if (!fin(H2) || G2 > 0) { H2 = INFINITY; S2 = -1.0; G2 = 1.0; }
    T2 = (H2 + w->dHi) / (S2 + w->dSi + w->RC);
    if (fin(H1) && G1 < 0) {
      T1 = (H1 + w->dHi) / (S1 + w->dSi + w->RC);
      if (T1 < T2 && G2 < 0) { S1 = S2; H1 = H2; T1 = T2; }
    } else if (G2 < 0) { S1 = S2; H1 = H2; T1 = T2; }
  }
  S2 = atpS[a][b]; H2 = atpH[a][b];
  T2 = (H2 + w->dHi) / (S2 + w->dSi + w->RC);
  if (fin(H1)) {
    if (T1 < T2) { *outS = S2; *outH = H2; } else { *outS = S1; *outH = H1; }
  } else { *outS = S2; *outH = H2; }
}

static void RSH(const work_t* w, int i, int j, double* S, double* H) {
  const int* n1 = w->n1; const int* n2 = w->n2;
  int a = n1[i], b = n2[j];
  if (!bp(a, b)) { *S = -1.0; *H = INFINITY; return; }
  double S1 = atpS[a][b] + tstack2S[a][n1[i + 1]][b][n2[j + 1]];
  double H1 = atpH[a][b] + tstack2H[a][n1[i + 1]][b][n2[j + 1]];
  int has = 0; double So = 0, Ho = 0;
  if (!bp(n1[i + 1], n2[j + 1])) {
    double h3 = dangle3H[a][n1[i + 1]][b], h5 = dangle5H[a][b][n2[j + 1]];
    if (fin(h3) && fin(h5)) {
      So = atpS[a][b] + dangle3S[a][n1[i + 1]][b] + dangle5S[a][b][n2[j + 1]];
      Ho = atpH[a][b] + h3 + h5; has = 1;
    } else if (fin(h3)) { So = atpS[a][b] + dangle3S[a][n1[i + 1]][b]; Ho = atpH[a][b] + h3; has = 1; }
    else if (fin(h5)) { So = atpS[a][b] + dangle5S[a][b][n2[j + 1]]; Ho = atpH[a][b] + h5; has = 1; }
  }
  end_term(w, S1, H1, has, So, Ho, a, b, S, H);
}

static void LSH(const work_t* w, int i, int j, double* S, double* H) {
  const int* n1 = w->n1; const int* n2 = w->n2;
  int a = n1[i], b = n2[j];
  if (!bp(a, b)) { *S = -1.0; *H = INFINITY; return; }
  double S1 = atpS[a][b] + tstack2S[b][n2[j - 1]][a][n1[i - 1]];
  double H1 = atpH[a][b] + tstack2H[b][n2[j - 1]][a][n1[i - 1]];
  int has = 0; double So = 0, Ho = 0;
  if (!bp(n1[i - 1], n2[j - 1])) {
    double h3 = dangle3H[b][n2[j - 1]][a], h5 = dangle5H[b][a][n1[i - 1]];
    if (fin(h3) && fin(h5)) {
      So = atpS[a][b] + dangle3S[b][n2[j - 1]][a] + dangle5S[b][a][n1[i - 1]];
      Ho = atpH[a][b] + h3 + h5; has = 1;
    } else if (fin(h3)) { So = atpS[a][b] + dangle3S[b][n2[j - 1]][a]; Ho = atpH[a][b] + h3; has = 1; }
    else if (fin(h5)) { So = atpS[a][b] + dangle5S[b][a][n1[i - 1]]; Ho = atpH[a][b] + h5; has = 1; }
  }
  end_term(w, S1, H1, has, So, Ho, a, b, S, H);
}

/* Bulge / internal loop between the inner pair (i,j) and the closing pair (ii,jj) of a dimer. Leaves
 * (*S,*H) untouched when the candidate loses against the current value of cell (ii,jj). */
static void bulge_internal_dimer(work_t* w, int i, int j, int ii, int jj, double* outS, double* outH, int tb) {
  const int* n1 = w->n1; const int* n2 = w->n2;
  int l1 = ii - i - 1, l2 = jj - j - 1, ls = l1 + l2 - 1;
  double S = -1.0, H = INFINITY, rS, rH, G1, G2;
  if (l1 + l2 > w->maxLoop) return; /* cannot occur inside the d-loop bounds; kept as a guard */
  if ((l1 == 0 && l2 > 0) || (l2 == 0 && l1 > 0)) {
    if (l2 == 1 || l1 == 1) {
      H = bulgeH[ls] + stackH[n1[i]][n1[ii]][n2[j]][n2[jj]];
      S = bulgeS[ls] + stackS[n1[i]][n1[ii]][n2[j]][n2[jj]];
      if (H > 0 || S > 0) { H = INFINITY; S = -1.0; }
      H += w->Hm[i][j]; S += w->Sm[i][j];
      if (!fin(H)) { H = INFINITY; S = -1.0; }
    } else {
      H = bulgeH[ls] + atpH[n1[i]][n2[j]] + atpH[n1[ii]][n2[jj]];
      H += w->Hm[i][j];
      S = bulgeS[ls] + atpS[n1[i]][n2[j]] + atpS[n1[ii]][n2[jj]];
      S += w->Sm[i][j];
      if (!fin(H)) { H = INFINITY; S = -1.0; }
      if (H > 0 && S > 0) { H = INFINITY; S = -1.0; }
    }
  } else if (l1 == 1 && l2 == 1) {
    S = stackint2S[n1[i]][n1[i + 1]][n2[j]][n2[j + 1]] + stackint2S[n2[jj]][n2[jj - 1]][n1[ii]][n1[ii - 1]];
    S += w->Sm[i][j];
    H = stackint2H[n1[i]][n1[i + 1]][n2[j]][n2[j + 1]] + stackint2H[n2[jj]][n2[jj - 1]][n1[ii]][n1[ii - 1]];
    H += w->Hm[i][j];
    if (!fin(H)) { H = INFINITY; S = -1.0; }
    if (H > 0 && S > 0) { H = INFINITY; S = -1.0; }
  } else {
    H = interiorH[ls] + tstackH[n1[i]][n1[i + 1]][n2[j]][n2[j + 1]] + tstackH[n2[jj]][n2[jj - 1]][n1[ii]][n1[ii - 1]] +
        (ILAH * abs(l1 - l2));
    H += w->Hm[i][j];
    S = interiorS[ls] + tstackS[n1[i]][n1[i + 1]][n2[j]][n2[j + 1]] + tstackS[n2[jj]][n2[jj - 1]][n1[ii]][n1[ii - 1]] +
        (ILAS * abs(l1 - l2));
    S += w->Sm[i][j];
    if (!fin(H)) { H = INFINITY; S = -1.0; }
    if (H > 0 && S > 0) { H = INFINITY; S = -1.0; }
  }
  RSH(w, ii, jj, &rS, &rH);
  G1 = H + rH - T_KELVIN * (S + rS);
  G2 = w->Hm[ii][jj] + rH - T_KELVIN * (w->Sm[ii][jj] + rS);
  if (G1 < G2 || tb) { *outS = S; *outH = H; }
}

static void fill_dimer(work_t* w) {
  const int* n1 = w->n1; const int* n2 = w->n2;
  for (int i = 1; i <= w->len1; i++)
    for (int j = 1; j <= w->len2; j++) {
      if (bp(n1[i], n2[j])) { w->Hm[i][j] = 0.0; w->Sm[i][j] = MIN_ENTROPY; }
      else { w->Hm[i][j] = INFINITY; w->Sm[i][j] = -1.0; }
    }
  for (int i = 1; i <= w->len1; i++)
    for (int j = 1; j <= w->len2; j++) {
      if (!fin(w->Hm[i][j])) continue;
      w->stat_cells++;
      double s = -1.0, h = INFINITY;
      LSH(w, i, j, &s, &h);
      if (fin(h)) { w->Sm[i][j] = s; w->Hm[i][j] = h; }
      if (i > 1 && j > 1) {
        /* stack onto (i-1,j-1) versus the current value, compared by Tm */
        double rS, rH, S0, H0, S1, H1, T0, T1;
        RSH(w, i, j, &rS, &rH);
        S0 = w->Sm[i][j]; H0 = w->Hm[i][j];
        T0 = (H0 + w->dHi + rH) / (S0 + w->dSi + rS + w->RC);
        double stH = stackH[n1[i - 1]][n1[i]][n2[j - 1]][n2[j]];
        if (fin(w->Hm[i - 1][j - 1]) && fin(stH)) {
          S1 = w->Sm[i - 1][j - 1] + stackS[n1[i - 1]][n1[i]][n2[j - 1]][n2[j]];
          H1 = w->Hm[i - 1][j - 1] + stH;
          T1 = (H1 + w->dHi + rH) / (S1 + w->dSi + rS + w->RC);
        } else {
          S1 = -1.0; H1 = INFINITY;
          T1 = (H1 + w->dHi) / (S1 + w->dSi + w->RC);
        }
        if (S1 < MIN_ENTROPY_CUTOFF) { S1 = MIN_ENTROPY; H1 = 0.0; }
        if (S0 < MIN_ENTROPY_CUTOFF) { S0 = MIN_ENTROPY; H0 = 0.0; }
        if (T1 > T0) { w->Sm[i][j] = S1; w->Hm[i][j] = H1; }
        else if (T0 >= T1) { w->Sm[i][j] = S0; w->Hm[i][j] = H0; }
        for (int d = 3; d <= w->maxLoop + 2; d++) {
          int ii = i - 1, jj = -ii - d + (j + i);
          if (jj < 1) { ii -= abs(jj - 1); jj = 1; }
          for (; ii > 0 && jj < j; --ii, ++jj) {
            if (!fin(w->Hm[ii][jj])) continue;
            w->stat_loopcand++;
            s = -1.0; h = INFINITY;
            bulge_internal_dimer(w, ii, jj, i, j, &s, &h, 0);
            if (s < MIN_ENTROPY_CUTOFF) { s = MIN_ENTROPY; h = 0.0; }
            if (fin(h)) { w->Hm[i][j] = h; w->Sm[i][j] = s; }
          }
        }
      }
    }
}

static void traceback_dimer(work_t* w, int i, int j, int* ps1, int* ps2) {
  const int* n1 = w->n1; const int* n2 = w->n2;
  ps1[i - 1] = j; ps2[j - 1] = i;
  for (int guard = 0; guard < 4 * ORACLE_MAX_LEN; guard++) {
    double s = -1.0, h = INFINITY;
    LSH(w, i, j, &s, &h);
    if (equal2(w->Sm[i][j], s) && equal2(w->Hm[i][j], h)) break;
    int done = 0;
    if (i > 1 && j > 1 &&
        equal2(w->Sm[i][j], stackS[n1[i - 1]][n1[i]][n2[j - 1]][n2[j]] + w->Sm[i - 1][j - 1]) &&
        equal2(w->Hm[i][j], stackH[n1[i - 1]][n1[i]][n2[j - 1]][n2[j]] + w->Hm[i - 1][j - 1])) {
      i = i - 1; j = j - 1; ps1[i - 1] = j; ps2[j - 1] = i; done = 1;
    }
    for (int d = 3; !done && d <= w->maxLoop + 2; ++d) {
      int ii = i - 1, jj = -ii - d + (j + i);
      if (jj < 1) { ii -= abs(jj - 1); jj = 1; }
      for (; !done && ii > 0 && jj < j; --ii, ++jj) {
        s = -1.0; h = INFINITY;
        bulge_internal_dimer(w, ii, jj, i, j, &s, &h, 1);
        if (equal2(w->Sm[i][j], s) && equal2(w->Hm[i][j], h)) {
          i = ii; j = jj; ps1[i - 1] = j; ps2[j - 1] = i; done = 1; break;
        }
      }
    }
    if (!done) break; /* the scalar original would spin; unreachable for consistent matrices */
  }
}

static void setup_seq(work_t* w, const char* o1, const char* o2, int reverse2) {
  w->len1 = (int)strlen(o1); w->len2 = (int)strlen(o2);
  for (int i = 1; i <= w->len1; i++) w->n1[i] = base_idx(o1[i - 1]);
  for (int j = 1; j <= w->len2; j++) w->n2[j] = base_idx(reverse2 ? o2[w->len2 - j] : o2[j - 1]);
  w->n1[0] = w->n1[w->len1 + 1] = w->n2[0] = w->n2[w->len2 + 1] = 4;
}

static int last_ps1[ORACLE_MAX_LEN];   /* pairing of the last dimer traceback: partner (1-based, in the reversed second oligo) of base i */
/* type: MSSPE_THAL_ANY or MSSPE_THAL_END1 */
static void thal_dimer(const char* o1, const char* o2, const msspe_thal_cond* c, int type, msspe_thal_out* out,
                       long* stats) {
  static work_t W; work_t* w = &W;
  memset(out, 0, sizeof(*out));
  w->stat_cells = w->stat_loopcand = 0;
  setup_seq(w, o1, o2, 1);
  w->dHi = 200; w->dSi = -5.7; w->maxLoop = c->max_loop;
  if (symmetry(o1, w->len1) && symmetry(o2, w->len2)) w->RC = R_GAS * log(c->dna_conc / 1000000000.0);
  else w->RC = R_GAS * log(c->dna_conc / 4000000000.0);
  double saltCorrection = salt_corr(c->mv, c->dv, c->dntp);
  fill_dimer(w);
  int bestI = 0, bestJ = 0; double bestG = INFINITY, rS, rH, G1;
  if (type == MSSPE_THAL_ANY) {
    for (int i = 1; i <= w->len1; i++)
      for (int j = 1; j <= w->len2; j++) {
        RSH(w, i, j, &rS, &rH);
        rS = rS + SMALL_NON_ZERO; rH = rH + SMALL_NON_ZERO;
        G1 = (w->Hm[i][j] + rH + w->dHi) - T_KELVIN * (w->Sm[i][j] + rS + w->dSi);
        if (G1 < bestG) { bestG = G1; bestI = i; bestJ = j; }
      }
  } else {
    bestI = w->len1; int i = w->len1;
    for (int j = 1; j <= w->len2; j++) {
      RSH(w, i, j, &rS, &rH);
      rS = rS + SMALL_NON_ZERO; rH = rH + SMALL_NON_ZERO;
      G1 = (w->Hm[i][j] + rH + w->dHi) - T_KELVIN * (w->Sm[i][j] + rS + w->dSi);
      if (G1 < bestG) { bestG = G1; bestJ = j; }
    }
  }
  if (!fin(bestG)) bestI = bestJ = 1;
  RSH(w, bestI, bestJ, &rS, &rH);
  double dH = w->Hm[bestI][bestJ] + rH + w->dHi;
  double dS = w->Sm[bestI][bestJ] + rS + w->dSi;
  if (stats) { stats[0] = w->stat_cells; stats[1] = w->stat_loopcand; }
  if (!fin(w->Hm[bestI][bestJ])) { out->no_structure = 1; out->tm = 0.0; return; }
  int ps1[ORACLE_MAX_LEN], ps2[ORACLE_MAX_LEN];
  for (int i = 0; i < w->len1; i++) ps1[i] = 0;
  for (int j = 0; j < w->len2; j++) ps2[j] = 0;
  traceback_dimer(w, bestI, bestJ, ps1, ps2);
  for (int i = 0; i < ORACLE_MAX_LEN; i++) last_ps1[i] = i < w->len1 ? ps1[i] : 0;
  int N = 0;
  for (int i = 0; i < w->len1; i++) if (ps1[i] > 0) ++N;
  for (int j = 0; j < w->len2; j++) if (ps2[j] > 0) ++N;
  out->n_bp = N / 2;
  N = (N / 2) - 1;
  double t = (dH / (dS + (N * saltCorrection) + w->RC)) - ABS_ZERO;
  double t_user = c->temp_c + ABS_ZERO;
  out->dg = dH - (t_user * (dS + (N * saltCorrection)));
  out->ds = dS + (N * saltCorrection);
  out->dh = dH;
  out->tm = t;
}

/* ------------------------------------------------------------------------------------------------
 * Hairpin (monomer): restated from the published structure of Primer3's unimolecular thermodynamic alignment and
 * pinned to 160 HAIRPIN outputs of the reference's ntthal executable and 337 HAIRPIN_TH values of its primer3_core
 * (tools/a64emu; see the header).
 * ------------------------------------------------------------------------------------------------ */
static double Ss2(const work_t* w, int i, int j) {
  if (i >= j) return -1.0;
  if (i == w->len1 || j == w->len2 + 1) return -1.0;
  return stackS[w->n1[i]][w->n1[i + 1]][w->n2[j]][w->n2[j - 1]];
}
static double Hs2(const work_t* w, int i, int j) {
  if (i >= j) return INFINITY;
  if (i == w->len1 || j == w->len2 + 1) return INFINITY;
  double h = stackH[w->n1[i]][w->n1[i + 1]][w->n2[j]][w->n2[j - 1]];
  return fin(h) ? h : INFINITY;
}

static void calc_hairpin(work_t* w, int i, int j, double* S, double* H, int tb) {
  const int* n1 = w->n1;
  int loopSize = j - i - 1;
  if (loopSize < MIN_HRPN_LOOP) { *S = -1.0; *H = INFINITY; return; }
  if (i <= w->len1 && w->len2 < j) { *S = -1.0; *H = INFINITY; return; }
  if (loopSize <= 30) { *H = hairpinH[loopSize - 1]; *S = hairpinS[loopSize - 1]; }
  else { *H = hairpinH[29]; *S = hairpinS[29]; }
  if (loopSize > 3) {
    *H += tstack2H[n1[i]][n1[i + 1]][n1[j]][n1[j - 1]];
    *S += tstack2S[n1[i]][n1[i + 1]][n1[j]][n1[j - 1]];
  } else if (loopSize == 3) {
    *H += atpH[n1[i]][n1[j]];
    *S += atpS[n1[i]][n1[j]];
  }
  if (loopSize == 3) {
    unsigned char key[6]; for (int c = 0; c < 5; c++) key[c] = (unsigned char)n1[i + c];
    struct loopent* e;
    if (nTriH && (e = bsearch(key, triH, nTriH, sizeof(struct loopent), cmp_loop5))) *H += e->value;
    if (nTriS && (e = bsearch(key, triS, nTriS, sizeof(struct loopent), cmp_loop5))) *S += e->value;
  } else if (loopSize == 4) {
    unsigned char key[6]; for (int c = 0; c < 6; c++) key[c] = (unsigned char)n1[i + c];
    struct loopent* e;
    if (nTetraH && (e = bsearch(key, tetraH, nTetraH, sizeof(struct loopent), cmp_loop6))) *H += e->value;
    if (nTetraS && (e = bsearch(key, tetraS, nTetraS, sizeof(struct loopent), cmp_loop6))) *S += e->value;
  }
  if (!fin(*H)) { *H = INFINITY; *S = -1.0; }
  if (*H > 0 && *S > 0 && (!(w->Hm[i][j] > 0) || !(w->Sm[i][j] > 0))) { *H = INFINITY; *S = -1.0; }
  double rs, rh;
  RSH(w, i, j, &rs, &rh);
  double G1 = *H + rh - T_KELVIN * (*S + rs);
  double G2 = w->Hm[i][j] + rh - T_KELVIN * (w->Sm[i][j] + rs);
  if (G2 < G1 && tb == 0) { *S = w->Sm[i][j]; *H = w->Hm[i][j]; }
}

/* loop between closing pair (i,j) and inner pair (ii,jj), i<ii<jj<j */
static void bulge_internal_mono(work_t* w, int i, int j, int ii, int jj, double* outS, double* outH, int tb) {
  const int* n1 = w->n1; const int* n2 = w->n2;
  int l1 = ii - i - 1, l2 = j - jj - 1, ls;
  double S, H, T1, T2;
  if (l1 + l2 > w->maxLoop) { *outS = -1.0; *outH = INFINITY; return; }
  ls = l1 + l2 - 1;
  if ((l1 == 0 && l2 > 0) || (l2 == 0 && l1 > 0)) {
    if (l2 == 1 || l1 == 1) {
      H = bulgeH[ls] + stackH[n1[i]][n1[ii]][n2[j]][n2[jj]];
      S = bulgeS[ls] + stackS[n1[i]][n1[ii]][n2[j]][n2[jj]];
    } else {
      H = bulgeH[ls] + atpH[n1[i]][n2[j]] + atpH[n1[ii]][n2[jj]];
      S = bulgeS[ls] + atpS[n1[i]][n2[j]] + atpS[n1[ii]][n2[jj]];
    }
    if (tb != 1) { H += w->Hm[ii][jj]; S += w->Sm[ii][jj]; }
    if (!fin(H)) { H = INFINITY; S = -1.0; }
    T1 = (H + w->dHi) / ((S + w->dSi) + w->RC);
    T2 = (w->Hm[i][j] + w->dHi) / ((w->Sm[i][j]) + w->dSi + w->RC);
    if ((T1 > T2) || ((tb && T1 >= T2) || tb == 1)) { *outS = S; *outH = H; }
  } else if (l1 == 1 && l2 == 1) {
    S = stackint2S[n1[i]][n1[i + 1]][n2[j]][n2[j - 1]] + stackint2S[n2[jj]][n2[jj + 1]][n1[ii]][n1[ii - 1]];
    if (tb != 1) S += w->Sm[ii][jj];
    H = stackint2H[n1[i]][n1[i + 1]][n2[j]][n2[j - 1]] + stackint2H[n2[jj]][n2[jj + 1]][n1[ii]][n1[ii - 1]];
    if (tb != 1) H += w->Hm[ii][jj];
    if (!fin(H)) { H = INFINITY; S = -1.0; }
    T1 = (H + w->dHi) / ((S + w->dSi) + w->RC);
    T2 = (w->Hm[i][j] + w->dHi) / ((w->Sm[i][j]) + w->dSi + w->RC);
    if ((T1 - T2 >= 0.000001) || tb) {
      if ((T1 > T2) || ((tb && T1 >= T2) || tb == 1)) { *outS = S; *outH = H; }
    }
  } else {
    H = interiorH[ls] + tstackH[n1[i]][n1[i + 1]][n2[j]][n2[j - 1]] + tstackH[n2[jj]][n2[jj + 1]][n1[ii]][n1[ii - 1]] +
        (ILAH * abs(l1 - l2));
    if (tb != 1) H += w->Hm[ii][jj];
    S = interiorS[ls] + tstackS[n1[i]][n1[i + 1]][n2[j]][n2[j - 1]] + tstackS[n2[jj]][n2[jj + 1]][n1[ii]][n1[ii - 1]] +
        (ILAS * abs(l1 - l2));
    if (tb != 1) S += w->Sm[ii][jj];
    if (!fin(H)) { H = INFINITY; S = -1.0; }
    T1 = (H + w->dHi) / ((S + w->dSi) + w->RC);
    T2 = (w->Hm[i][j] + w->dHi) / ((w->Sm[i][j]) + w->dSi + w->RC);
    if ((T1 > T2) || ((tb && T1 >= T2) || (tb == 1))) { *outS = S; *outH = H; }
  }
}

static void CBI(work_t* w, int i, int j, double* S, double* H, int tb) {
  for (int d = j - i - 3; d >= MIN_HRPN_LOOP + 1 && d >= j - i - 2 - w->maxLoop; --d)
    for (int ii = i + 1; ii < j - d && ii <= w->len1; ++ii) {
      int jj = d + ii;
      if (tb == 0) { *S = -1.0; *H = INFINITY; }
      if (fin(w->Hm[ii][jj]) && fin(w->Hm[i][j])) {
        bulge_internal_mono(w, i, j, ii, jj, S, H, tb);
        if (fin(*H)) {
          if (*S < MIN_ENTROPY_CUTOFF) { *S = MIN_ENTROPY; *H = 0.0; }
          if (tb == 0) { w->Hm[i][j] = *H; w->Sm[i][j] = *S; }
        }
      }
    }
}

static void fill_mono(work_t* w) {
  const int* n1 = w->n1;
  for (int i = 1; i <= w->len1; ++i)
    for (int j = i; j <= w->len2; ++j) {
      if (j - i < MIN_HRPN_LOOP + 1 || !bp(n1[i], n1[j])) { w->Hm[i][j] = INFINITY; w->Sm[i][j] = -1.0; }
      else { w->Hm[i][j] = 0.0; w->Sm[i][j] = MIN_ENTROPY; }
    }
  for (int j = 2; j <= w->len2; ++j)
    for (int i = j - MIN_HRPN_LOOP - 1; i >= 1; --i) {
      if (!fin(w->Hm[i][j])) continue;
      /* stack onto (i+1,j-1) vs current, by Tm */
      double S0 = w->Sm[i][j], H0 = w->Hm[i][j], S1, H1, T0, T1;
      T0 = (H0 + w->dHi) / (S0 + w->dSi + w->RC);
      S1 = w->Sm[i + 1][j - 1] + Ss2(w, i, j);
      H1 = w->Hm[i + 1][j - 1] + Hs2(w, i, j);
      T1 = (H1 + w->dHi) / (S1 + w->dSi + w->RC);
      if (S1 < MIN_ENTROPY_CUTOFF) { S1 = MIN_ENTROPY; H1 = 0.0; }
      if (S0 < MIN_ENTROPY_CUTOFF) { S0 = MIN_ENTROPY; H0 = 0.0; }
      if (T1 > T0) { w->Sm[i][j] = S1; w->Hm[i][j] = H1; } else { w->Sm[i][j] = S0; w->Hm[i][j] = H0; }
      double s = -1.0, h = INFINITY;
      CBI(w, i, j, &s, &h, 0);
      s = -1.0; h = INFINITY;
      calc_hairpin(w, i, j, &s, &h, 0);
      if (fin(h)) {
        if (s < MIN_ENTROPY_CUTOFF) { s = MIN_ENTROPY; h = 0.0; }
        w->Sm[i][j] = s; w->Hm[i][j] = h;
      }
    }
}

/* the four ways the 5' exterior fragment [1..i] can end in a helix closing at i (or i-1) */
static void end5_variant(const work_t* w, int i, int variant, double* outH, double* outS) {
  const int* n1 = w->n1;
  double H_max = INFINITY, S_max = -1.0, max_tm = -INFINITY;
  int kmax = (variant == 1) ? i - MIN_HRPN_LOOP - 2 : (variant == 4 ? i - MIN_HRPN_LOOP - 4 : i - MIN_HRPN_LOOP - 3);
  for (int k = 0; k <= kmax; ++k) {
    double T1 = (w->hend5[k] + w->dHi) / (w->send5[k] + w->dSi + w->RC);
    double T2 = (0 + w->dHi) / (0 + w->dSi + w->RC);
    double eH, eS;
    switch (variant) {
      case 1: eH = atpH[n1[k + 1]][n1[i]] + w->Hm[k + 1][i]; eS = atpS[n1[k + 1]][n1[i]] + w->Sm[k + 1][i]; break;
      case 2: /* 5' dangle: Hd5(i, k+2) */
        eH = atpH[n1[k + 2]][n1[i]] + dangle5H[n1[i]][n1[k + 2]][n1[k + 1]] + w->Hm[k + 2][i];
        eS = atpS[n1[k + 2]][n1[i]] + dangle5S[n1[i]][n1[k + 2]][n1[k + 1]] + w->Sm[k + 2][i]; break;
      case 3: /* 3' dangle: Hd3(i-1, k+1) */
        eH = atpH[n1[k + 1]][n1[i - 1]] + dangle3H[n1[i - 1]][n1[i]][n1[k + 1]] + w->Hm[k + 1][i - 1];
        eS = atpS[n1[k + 1]][n1[i - 1]] + dangle3S[n1[i - 1]][n1[i]][n1[k + 1]] + w->Sm[k + 1][i - 1]; break;
      default: /* terminal mismatch: Htstack(i-1, k+2) */
        eH = atpH[n1[k + 2]][n1[i - 1]] + tstack2H[n1[i - 1]][n1[i]][n1[k + 2]][n1[k + 1]] + w->Hm[k + 2][i - 1];
        eS = atpS[n1[k + 2]][n1[i - 1]] + tstack2S[n1[i - 1]][n1[i]][n1[k + 2]][n1[k + 1]] + w->Sm[k + 2][i - 1]; break;
    }
    double H, S;
    if (T1 >= T2) { H = w->hend5[k] + eH; S = w->send5[k] + eS; }
    else { H = 0 + eH; S = 0 + eS; }
    if (!fin(H) || H > 0 || S > 0) { H = INFINITY; S = -1.0; }
    T1 = (H + w->dHi) / (S + w->dSi + w->RC);
    if (max_tm < T1) {
      if (S > MIN_ENTROPY_CUTOFF) { H_max = H; S_max = S; max_tm = T1; }
    }
  }
  *outH = H_max; *outS = S_max;
}

static void calc_terminal_bp(work_t* w, double temp) {
  w->send5[0] = w->send5[1] = -1.0;
  w->hend5[0] = w->hend5[1] = INFINITY;
  for (int i = 2; i <= w->len1; i++) { w->send5[i] = MIN_ENTROPY; w->hend5[i] = 0; }
  for (int i = 2; i <= w->len1; ++i) {
    double eh[5], es[5], T[6];
    T[1] = (w->hend5[i - 1] + w->dHi) / (w->send5[i - 1] + w->dSi + w->RC);
    for (int v = 1; v <= 4; v++) {
      end5_variant(w, i, v, &eh[v], &es[v]);
      T[v + 1] = (eh[v] + w->dHi) / (es[v] + w->dSi + w->RC);
    }
    int max;
    if (T[1] > T[2] && T[1] > T[3] && T[1] > T[4] && T[1] > T[5]) max = 1;
    else if (T[2] > T[3] && T[2] > T[4] && T[2] > T[5]) max = 2;
    else if (T[3] > T[4] && T[3] > T[5]) max = 3;
    else if (T[4] > T[5]) max = 4;
    else max = 5;
    if (max == 1) { w->send5[i] = w->send5[i - 1]; w->hend5[i] = w->hend5[i - 1]; }
    else {
      int v = max - 1;
      double G = eh[v] - (temp * (es[v]));
      if (G < 0.0) { w->send5[i] = es[v]; w->hend5[i] = eh[v]; }
      else { w->send5[i] = w->send5[i - 1]; w->hend5[i] = w->hend5[i - 1]; }
    }
  }
}

static void traceback_mono(work_t* w, int* bpv) {
  const int* n1 = w->n1;
  struct { int i, j, m; } stack[4 * ORACLE_MAX_LEN]; int sp = 0;
#define PUSH(a, b, c) do { if (sp < 4 * ORACLE_MAX_LEN) { stack[sp].i = (a); stack[sp].j = (b); stack[sp].m = (c); sp++; } } while (0)
  PUSH(w->len1, 0, 1);
  while (sp > 0) {
    sp--;
    int i = stack[sp].i, j = stack[sp].j, m = stack[sp].m;
    if (m == 1) {
      while (i >= 1 && equal2(w->send5[i], w->send5[i - 1]) && equal2(w->hend5[i], w->hend5[i - 1])) --i;
      if (i == 0) continue;
      double eh, es; int handled = 0;
      for (int v = 1; v <= 4 && !handled; v++) {
        end5_variant(w, i, v, &eh, &es);
        if (!(equal2(w->send5[i], es) && equal2(w->hend5[i], eh))) continue;
        handled = 1;
        int kmax = (v == 1) ? i - MIN_HRPN_LOOP - 2 : (v == 4 ? i - MIN_HRPN_LOOP - 4 : i - MIN_HRPN_LOOP - 3);
        for (int k = 0; k <= kmax; ++k) {
          double xS, xH; int pi, pj;
          switch (v) {
            case 1: xS = atpS[n1[k + 1]][n1[i]] + w->Sm[k + 1][i]; xH = atpH[n1[k + 1]][n1[i]] + w->Hm[k + 1][i]; pi = k + 1; pj = i; break;
            case 2: xS = atpS[n1[k + 2]][n1[i]] + dangle5S[n1[i]][n1[k + 2]][n1[k + 1]] + w->Sm[k + 2][i];
                    xH = atpH[n1[k + 2]][n1[i]] + dangle5H[n1[i]][n1[k + 2]][n1[k + 1]] + w->Hm[k + 2][i]; pi = k + 2; pj = i; break;
            case 3: xS = atpS[n1[k + 1]][n1[i - 1]] + dangle3S[n1[i - 1]][n1[i]][n1[k + 1]] + w->Sm[k + 1][i - 1];
                    xH = atpH[n1[k + 1]][n1[i - 1]] + dangle3H[n1[i - 1]][n1[i]][n1[k + 1]] + w->Hm[k + 1][i - 1]; pi = k + 1; pj = i - 1; break;
            default: xS = atpS[n1[k + 2]][n1[i - 1]] + tstack2S[n1[i - 1]][n1[i]][n1[k + 2]][n1[k + 1]] + w->Sm[k + 2][i - 1];
                     xH = atpH[n1[k + 2]][n1[i - 1]] + tstack2H[n1[i - 1]][n1[i]][n1[k + 2]][n1[k + 1]] + w->Hm[k + 2][i - 1]; pi = k + 2; pj = i - 1; break;
          }
          if (equal2(w->send5[i], xS) && equal2(w->hend5[i], xH)) { PUSH(pi, pj, 0); break; }
          else if (equal2(w->send5[i], w->send5[k] + xS) && equal2(w->hend5[i], w->hend5[k] + xH)) {
            PUSH(pi, pj, 0); PUSH(k, 0, 1); break;
          }
        }
      }
    } else {
      bpv[i - 1] = j; bpv[j - 1] = i;
      double s1 = -1.0, h1 = INFINITY, s2 = -1.0, h2 = INFINITY;
      calc_hairpin(w, i, j, &s1, &h1, 1);
      CBI(w, i, j, &s2, &h2, 2);
      if (equal2(w->Sm[i][j], Ss2(w, i, j) + w->Sm[i + 1][j - 1]) && equal2(w->Hm[i][j], Hs2(w, i, j) + w->Hm[i + 1][j - 1])) {
        PUSH(i + 1, j - 1, 0);
      } else if (equal2(w->Sm[i][j], s1) && equal2(w->Hm[i][j], h1)) {
        /* hairpin loop closes here */
      } else if (equal2(w->Sm[i][j], s2) && equal2(w->Hm[i][j], h2)) {
        int done = 0;
        for (int d = j - i - 3; d >= MIN_HRPN_LOOP + 1 && d >= j - i - 2 - w->maxLoop && !done; --d)
          for (int ii = i + 1; ii < j - d; ++ii) {
            int jj = d + ii; double es = -1.0, eh = INFINITY;
            bulge_internal_mono(w, i, j, ii, jj, &es, &eh, 1);
            if (equal2(w->Sm[i][j], es + w->Sm[ii][jj]) && equal2(w->Hm[i][j], eh + w->Hm[ii][jj])) {
              PUSH(ii, jj, 0); ++done; break;
            }
          }
      }
    }
  }
#undef PUSH
}

static void thal_hairpin(const char* o1, const msspe_thal_cond* c, msspe_thal_out* out) {
  static work_t W; work_t* w = &W;
  memset(out, 0, sizeof(*out));
  setup_seq(w, o1, o1, 0);
  w->dHi = 0.0; w->dSi = -0.00000000001; w->RC = 0; w->maxLoop = c->max_loop;
  double saltCorrection = salt_corr(c->mv, c->dv, c->dntp);
  double temp = c->temp_c + ABS_ZERO;
  fill_mono(w);
  calc_terminal_bp(w, temp);
  double mh = w->hend5[w->len1], ms = w->send5[w->len1];
  if (!fin(mh) || !fin(ms)) { out->no_structure = 1; out->tm = 0.0; return; }
  int bpv[ORACLE_MAX_LEN + 2];
  for (int k = 0; k < w->len1; ++k) bpv[k] = 0;
  traceback_mono(w, bpv);
  for (int k = 0; k < ORACLE_MAX_LEN; ++k) last_ps1[k] = k < w->len1 ? bpv[k] : 0;   /* oracle_thal_last_pairing: partner position, 1-based */
  int N = 0;
  for (int i = 1; i < w->len1; ++i) if (bpv[i - 1] > 0) N++;
  out->n_bp = N / 2;
  double t = (mh / (ms + (((N / 2) - 1) * saltCorrection))) - ABS_ZERO;
  out->dg = mh - (temp * (ms + (((N / 2) - 1) * saltCorrection)));
  out->ds = ms + (((N / 2) - 1) * saltCorrection);
  out->dh = mh;
  out->tm = t;
}

/* ---- parameter sources: the data generated from the reference's primer3_config (tools/gen_thal_params.py)
 * or a primer3_config directory read at run time ---- */
static const msspe_thal_raw_params EMBEDDED_PARAMS =
#include "../open-msspe-design_b200/csrc/thal_params_data.inc"
;

void oracle_thal_load_embedded(void) { oracle_thal_set_params(&EMBEDDED_PARAMS); }

static int read_table(const char* dir, const char* fn, double* dst, int n) {
  char path[1024]; snprintf(path, sizeof path, "%s/%s", dir, fn);
  FILE* f = fopen(path, "r"); if (!f) return -1;
  char tok[64]; int i = 0;
  while (i < n && fscanf(f, "%63s", tok) == 1) dst[i++] = (strcmp(tok, "inf") == 0) ? INFINITY : atof(tok);
  fclose(f); return i == n ? 0 : -1;
}
static int read_loops(const char* dir, const char* fn, double* dst) {
  char path[1024]; snprintf(path, sizeof path, "%s/%s", dir, fn);
  FILE* f = fopen(path, "r"); if (!f) return -1;
  char t[4][64]; int r = 0;
  while (r < 30 && fscanf(f, "%63s %63s %63s %63s", t[0], t[1], t[2], t[3]) == 4) {
    for (int c = 0; c < 3; c++) dst[3 * r + c] = (strcmp(t[c + 1], "inf") == 0) ? INFINITY : atof(t[c + 1]);
    r++;
  }
  fclose(f); return r == 30 ? 0 : -1;
}
static int read_nloop(const char* dir, const char* fn, char seqs[][8], double* vals, int cap) {
  char path[1024]; snprintf(path, sizeof path, "%s/%s", dir, fn);
  FILE* f = fopen(path, "r"); if (!f) return -1;
  char s[64], v[64]; int n = 0;
  while (n < cap && fscanf(f, "%63s %63s", s, v) == 2) { strncpy(seqs[n], s, 7); seqs[n][7] = 0; vals[n] = atof(v); n++; }
  fclose(f); return n;
}
int oracle_thal_load_dir(const char* dir) {
  static msspe_thal_raw_params p; memset(&p, 0, sizeof p);
  if (read_table(dir, "stack.ds", p.stack_ds, 256) || read_table(dir, "stack.dh", p.stack_dh, 256) ||
      read_table(dir, "stackmm.ds", p.stackmm_ds, 256) || read_table(dir, "stackmm.dh", p.stackmm_dh, 256) ||
      read_table(dir, "dangle.ds", p.dangle_ds, 128) || read_table(dir, "dangle.dh", p.dangle_dh, 128) ||
      read_loops(dir, "loops.ds", p.loops_ds) || read_loops(dir, "loops.dh", p.loops_dh) ||
      read_table(dir, "tstack_tm_inf.ds", p.tstack_ds, 256) || read_table(dir, "tstack.dh", p.tstack_dh, 256) ||
      read_table(dir, "tstack2.ds", p.tstack2_ds, 256) || read_table(dir, "tstack2.dh", p.tstack2_dh, 256)) return -1;
  if ((p.n_triloop_ds = read_nloop(dir, "triloop.ds", p.triloop_ds_seq, p.triloop_ds, 32)) < 0) return -1;
  if ((p.n_triloop_dh = read_nloop(dir, "triloop.dh", p.triloop_dh_seq, p.triloop_dh, 32)) < 0) return -1;
  if ((p.n_tetraloop_ds = read_nloop(dir, "tetraloop.ds", p.tetraloop_ds_seq, p.tetraloop_ds, 128)) < 0) return -1;
  if ((p.n_tetraloop_dh = read_nloop(dir, "tetraloop.dh", p.tetraloop_dh_seq, p.tetraloop_dh, 128)) < 0) return -1;
  oracle_thal_set_params(&p);
  return 0;
}

/* ---- test hook: the loaded tables under the names libprimer3's thal.c gives them, so that they can be compared with the
 * arrays compiled into the reference's Primer3 2.6.1 executables (tests/golden/primer3_2_6_1_compiled_in_tables.json,
 * tools/extract_primer3_compiled_in_tables.py).  Returns the number of doubles written, -1 for an unknown name. ---- */
int oracle_thal_table(const char* name, double* out, int cap) {
  static const struct { const char* name; const double* p; int n; } T[] = {
      {"stackEntropies", &stackS[0][0][0][0], 625}, {"stackEnthalpies", &stackH[0][0][0][0], 625},
      {"stackint2Entropies", &stackint2S[0][0][0][0], 625}, {"stackint2Enthalpies", &stackint2H[0][0][0][0], 625},
      {"tstackEntropies", &tstackS[0][0][0][0], 625}, {"tstackEnthalpies", &tstackH[0][0][0][0], 625},
      {"tstack2Entropies", &tstack2S[0][0][0][0], 625}, {"tstack2Enthalpies", &tstack2H[0][0][0][0], 625},
      {"dangleEntropies3", &dangle3S[0][0][0], 125}, {"dangleEnthalpies3", &dangle3H[0][0][0], 125},
      {"dangleEntropies5", &dangle5S[0][0][0], 125}, {"dangleEnthalpies5", &dangle5H[0][0][0], 125},
      {"hairpinLoopEntropies", hairpinS, 30}, {"interiorLoopEntropies", interiorS, 30}, {"bulgeLoopEntropies", bulgeS, 30},
      {"hairpinLoopEnthalpies", hairpinH, 30}, {"interiorLoopEnthalpies", interiorH, 30}, {"bulgeLoopEnthalpies", bulgeH, 30},
      {"atpS", &atpS[0][0], 25}, {"atpH", &atpH[0][0], 25}};
  if (!params_loaded) return -1;
  for (unsigned i = 0; i < sizeof T / sizeof T[0]; i++)
    if (strcmp(name, T[i].name) == 0) {
      if (cap < T[i].n) return -1;
      memcpy(out, T[i].p, sizeof(double) * T[i].n);
      return T[i].n;
    }
  return -1;
}
/* The sorted tri/tetraloop bonus tables: keys[8 * i .. ] = the loop as base indices (0..3), vals[i] its bonus. */
int oracle_thal_loop_table(const char* name, unsigned char* keys, double* vals, int cap) {
  const struct loopent* t; int n, len;
  if (!params_loaded) return -1;
  if (strcmp(name, "defaultTriloopEntropies") == 0) { t = triS; n = nTriS; len = 5; }
  else if (strcmp(name, "defaultTriloopEnthalpies") == 0) { t = triH; n = nTriH; len = 5; }
  else if (strcmp(name, "defaultTetraloopEntropies") == 0) { t = tetraS; n = nTetraS; len = 6; }
  else if (strcmp(name, "defaultTetraloopEnthalpies") == 0) { t = tetraH; n = nTetraH; len = 6; }
  else return -1;
  if (cap < n) return -1;
  for (int i = 0; i < n; i++) { memset(keys + 8 * i, 0xff, 8); memcpy(keys + 8 * i, t[i].key, len); vals[i] = t[i].value; }
  return n;
}

/* ---- exported ---- */
int oracle_thal(const char* o1, const char* o2, int type, const msspe_thal_cond* c, msspe_thal_out* out) {
  if (!params_loaded) return -1;
  if (strlen(o1) > ORACLE_MAX_LEN || (o2 && strlen(o2) > ORACLE_MAX_LEN)) return -2;
  if (type == MSSPE_THAL_HAIRPIN) thal_hairpin(o1, c, out);
  else if (type == 3) thal_dimer(o2, o1, c, MSSPE_THAL_END1, out, NULL);  /* thal_end2 = END1 with the oligos exchanged (not on od-msspe's path; kept so the whole ntthal fixture applies) */
  else thal_dimer(o1, o2, c, type, out, NULL);
  return 0;
}

/* the traced duplex of the most recent oracle_thal dimer call, or the traced fold of the most recent hairpin call (partner position
 * within the same oligo), i.e. what ntthal draws; for the drawing tests */
int oracle_thal_last_pairing(int* ps1, int cap) {
  int n = cap < ORACLE_MAX_LEN ? cap : ORACLE_MAX_LEN;
  for (int i = 0; i < n; i++) ps1[i] = last_ps1[i];
  return n;
}

int oracle_thal_stats(const char* o1, const char* o2, const msspe_thal_cond* c, long* stats) {
  msspe_thal_out out;
  if (!params_loaded) return -1;
  thal_dimer(o1, o2, c, MSSPE_THAL_ANY, &out, stats);
  return 0;
}

/* oligotm with SantaLucia-1998 parameters and SantaLucia salt correction, as primer3_core applies it for
 * PRIMER_LEFT_0_TM when no salt tags are given (reference primer.rs:125-140).  SURVEY.md Appendix C. */
double oracle_oligotm(const char* s, double mv, double dv, double dntp, double dna_conc) {
  int len = (int)strlen(s);
  double dh = 0, ds = 0;
  int sym = symmetry(s, len);
  if (sym) ds += -1.4;
  int f = base_idx(s[0]), l = base_idx(s[len - 1]);
  /* terminal corrections */
  if (f == 0 || f == 3) { ds += 4.1; dh += 2300; } else { ds += -2.8; dh += 100; }
  if (l == 0 || l == 3) { ds += 4.1; dh += 2300; } else { ds += -2.8; dh += 100; }
  for (int i = 0; i + 1 < len; i++) {
    int a = base_idx(s[i]), b = base_idx(s[i + 1]);
    if (a > 3 || b > 3) return -999999.9999;
    /* Watson-Crick stack 5'-ab-3' / 3'-a'b'-5' */
    dh += stackH[a][b][3 - a][3 - b];
    ds += stackS[a][b][3 - a][3 - b];
  }
  if (dv == 0) dntp = 0;
  if (dv < dntp) dv = dntp;
  double K = mv + 120 * sqrt(dv - dntp);
  ds = ds + 0.368 * (len - 1) * log(K / 1000.0);
  double tm;
  if (sym) tm = dh / (ds + 1.987 * log(dna_conc / 1000000000.0)) - 273.15;
  else tm = dh / (ds + 1.987 * log(dna_conc / 4000000000.0)) - 273.15;
  return tm;
}

double oracle_gc_percent(const char* s) {
  int len = (int)strlen(s), gc = 0;
  for (int i = 0; i < len; i++) { int b = base_idx(s[i]); if (b == 1 || b == 2) gc++; }
  return 100.0 * gc / len;
}
