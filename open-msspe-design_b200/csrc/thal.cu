// thal.cu -- K4/K5/K6: nearest-neighbour thermodynamics on the device.
//   K5 thal_dimer_kernel : Primer3's thermodynamic alignment of two oligos (ntthal -a ANY / END1), the work the
//                          reference sends to an `ntthal` subprocess pair by pair (od-msspe/src/delta_g.rs:83-153)
//                          and to primer3_core for SELF_ANY_TH / SELF_END_TH (primer.rs:125-166).
//   K6 thal_mono_kernel  : hairpin (ntthal -a HAIRPIN / PRIMER_LEFT_0_HAIRPIN_TH).
//   K4 oligotm_kernel    : PRIMER_LEFT_0_TM (oligotm, SantaLucia 1998 + SantaLucia salt correction at Primer3's
//                          default concentrations, primer.rs:125-140 passes no salt tags) and GC percent.
// All arithmetic is FP64 and this file is compiled with -fmad=false so every expression rounds exactly like
// scalar C without contraction; the tensor cores are not involved (nothing here is a contraction).
//
// Dimer kernels (SM-issue bound, not HBM bound: 16 B in, <= 48 B out per pair), three exact forms chosen per call
// (launch_dimer; MSSPE_THAL_KERNEL = thread | flat | legacy forces one; tests/test_gpu_thermo.py runs all of them):
//   * thal_dimer_thread_kernel  one THREAD per ordered pair -- large batches of oligos <= 16 nt (the all-ordered-pairs
//     matrix of delta_g.rs:61-81): paired cells compact in an L2-resident scratch, one (S,H) table with a branch-free
//     candidate, one flat candidate loop per row, matrix columns visited in base-composition order (column_order);
//   * thal_dimer_flat_kernel    one warp per pair, all loop candidates of a row in one flat index space -- oligos > 16 nt;
//   * thal_dimer_kernel         one warp per pair, lanes over the inner rows of a cell's candidates -- small batches.
// Common to all: tables in shared memory, among them the LEFT/RIGHT end-of-duplex terms pre-tabulated on the host over
// their 2x2 base context for this run's RC (~60 % of the scalar algorithm's work when recomputed per use); rows of the DP
// matrix in order; every minimum taken as (dG, scan-order key) so that the result equals Primer3's sequential scan; then
// best-cell selection, traceback (needed for the salt correction: N paired bases) and dS/dH/dG/Tm.
// Hairpin: thal_mono_group_kernel (16 lanes per oligo, oligos <= 16 nt) / thal_mono_kernel (one thread per oligo).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "engine.cuh"
#include "thal_tables.cuh"

void msspe_thal_expand(const msspe_thal_raw_params* p, ThalDeviceTables* T);

namespace {

constexpr double kR = 1.9872, kAbsZero = 273.15, kTK = 310.15, kMinEntropyCutoff = -2500.0, kMinEntropy = -3224.0;
constexpr double kSmallNonZero = 0.000001, kDHi = 200.0, kDSi = -5.7;
#define K_ILAS (-300 / 310.15)
#define K_ILAH 0.0
constexpr int DIMER_THREADS = 128;

// ---------------------------------------------------------------- host: per-run end-of-duplex tables
struct EndSel { double S, H; };
EndSel end_choice(const ThalDeviceTables& T, double RC, double S1, double H1, bool has_opt, double So, double Ho, int a, int b) {
  double G1 = H1 - kTK * S1, T1 = -INFINITY, S2, H2, G2, T2;
  if (!std::isfinite(H1) || G1 > 0) { H1 = INFINITY; S1 = -1.0; G1 = 1.0; }
  if (has_opt) {
    S2 = So; H2 = Ho; G2 = H2 - kTK * S2;
    if (!std::isfinite(H2) || G2 > 0) { H2 = INFINITY; S2 = -1.0; G2 = 1.0; }
    T2 = (H2 + kDHi) / (S2 + kDSi + RC);
    if (std::isfinite(H1) && G1 < 0) {
      T1 = (H1 + kDHi) / (S1 + kDSi + RC);
      if (T1 < T2 && G2 < 0) { S1 = S2; H1 = H2; T1 = T2; }
    } else if (G2 < 0) { S1 = S2; H1 = H2; T1 = T2; }
  }
  S2 = T.atpS[a * 5 + b]; H2 = T.atpH[a * 5 + b];
  T2 = (H2 + kDHi) / (S2 + kDSi + RC);
  if (std::isfinite(H1)) { if (T1 < T2) return {S2, H2}; return {S1, H1}; }
  return {S2, H2};
}

void build_dimer_consts(const ThalDeviceTables& T, const msspe_thal_cond& c, ThalDimerConsts* K) {
  memset(K, 0, sizeof *K);
  K->RC[0] = kR * log(c.dna_conc / 4000000000.0);
  K->RC[1] = kR * log(c.dna_conc / 1000000000.0);
  double dntp = c.dntp;
  if (c.dv <= 0) dntp = c.dv;
  K->saltCorr = 0.368 * (log((c.mv + 120 * (sqrt(fmax(0.0, c.dv - dntp)))) / 1000));
  K->t_user_K = c.temp_c + kAbsZero;
  K->maxLoop = c.max_loop;
  for (int sym = 0; sym < 2; sym++)
    for (int a = 0; a < 4; a++) {
      const int b = 3 - a;
      for (int x = 0; x < 5; x++)      // neighbour on strand 1 (i+1 for RIGHT, i-1 for LEFT)
        for (int y = 0; y < 5; y++) {  // neighbour on strand 2 (j+1 for RIGHT, j-1 for LEFT)
          const int idx = (a * 5 + x) * 5 + y;
          const bool nb_pair = (x + y == 3) && x < 4 && y < 4;
          {  // RIGHT end: tstack2[a][x][b][y], dangle3[a][x][b], dangle5[a][b][y]
            double S1 = T.atpS[a * 5 + b] + T.tstack2S[THAL_IDX4(a, x, b, y)];
            double H1 = T.atpH[a * 5 + b] + T.tstack2H[THAL_IDX4(a, x, b, y)];
            bool has = false; double So = 0, Ho = 0;
            if (!nb_pair) {
              const double h3 = T.dangle3H[THAL_IDX3(a, x, b)], h5 = T.dangle5H[THAL_IDX3(a, b, y)];
              if (std::isfinite(h3) && std::isfinite(h5)) {
                So = T.atpS[a * 5 + b] + T.dangle3S[THAL_IDX3(a, x, b)] + T.dangle5S[THAL_IDX3(a, b, y)];
                Ho = T.atpH[a * 5 + b] + h3 + h5; has = true;
              } else if (std::isfinite(h3)) { So = T.atpS[a * 5 + b] + T.dangle3S[THAL_IDX3(a, x, b)]; Ho = T.atpH[a * 5 + b] + h3; has = true; }
              else if (std::isfinite(h5)) { So = T.atpS[a * 5 + b] + T.dangle5S[THAL_IDX3(a, b, y)]; Ho = T.atpH[a * 5 + b] + h5; has = true; }
            }
            EndSel e = end_choice(T, K->RC[sym], S1, H1, has, So, Ho, a, b);
            K->rshS[sym][idx] = e.S; K->rshH[sym][idx] = e.H;
          }
          {  // LEFT end: tstack2[b][y][a][x], dangle3[b][y][a], dangle5[b][a][x]
            double S1 = T.atpS[a * 5 + b] + T.tstack2S[THAL_IDX4(b, y, a, x)];
            double H1 = T.atpH[a * 5 + b] + T.tstack2H[THAL_IDX4(b, y, a, x)];
            bool has = false; double So = 0, Ho = 0;
            if (!nb_pair) {
              const double h3 = T.dangle3H[THAL_IDX3(b, y, a)], h5 = T.dangle5H[THAL_IDX3(b, a, x)];
              if (std::isfinite(h3) && std::isfinite(h5)) {
                So = T.atpS[a * 5 + b] + T.dangle3S[THAL_IDX3(b, y, a)] + T.dangle5S[THAL_IDX3(b, a, x)];
                Ho = T.atpH[a * 5 + b] + h3 + h5; has = true;
              } else if (std::isfinite(h3)) { So = T.atpS[a * 5 + b] + T.dangle3S[THAL_IDX3(b, y, a)]; Ho = T.atpH[a * 5 + b] + h3; has = true; }
              else if (std::isfinite(h5)) { So = T.atpS[a * 5 + b] + T.dangle5S[THAL_IDX3(b, a, x)]; Ho = T.atpH[a * 5 + b] + h5; has = true; }
            }
            EndSel e = end_choice(T, K->RC[sym], S1, H1, has, So, Ho, a, b);
            K->lshS[sym][idx] = e.S; K->lshH[sym][idx] = e.H;
          }
        }
    }
  // t0[sym][li * 25 + rn]: the quotient of the stack test for a cell that starts a duplex (thal_dimer_flat_kernel), with the
  // kernel's expression: li = (a, n1[i-1], n2[j-1]) indexes lsh, a * 25 + rn = (a, n1[i+1], n2[j+1]) indexes rsh
  for (int sym = 0; sym < 2; sym++)
    for (int li = 0; li < 100; li++)
      for (int rn = 0; rn < 25; rn++) {
        const int ri = (li / 25) * 25 + rn;
        const double H0 = K->lshH[sym][li], S0 = K->lshS[sym][li], rH = K->rshH[sym][ri], rS = K->rshS[sym][ri], RC = K->RC[sym];
        K->t0[sym][li * 25 + rn] = (H0 + kDHi + rH) / (S0 + kDSi + rS + RC);
      }
}

// ---------------------------------------------------------------- device: dimer
struct DimerArgs {
  const uint64_t* a; const uint64_t* b;  // pair list: a[p], b[p].  matrix: a = b = codes[n]
  unsigned long long n_pairs;
  uint32_t n, row_begin;
  int matrix, k, type;
  const ThalDeviceTables* T; const ThalDimerConsts* C;
  msspe_thal_out* out;
  double dg_limit;
  msspe_dimer_edge* edges; unsigned long long edge_cap; unsigned long long* n_edges;
  uint64_t* nostruct; unsigned long long nostruct_cap; unsigned long long* n_nostruct;
  int dbg;  // diagnostic: 1 = skip loop candidates, 2 = skip traceback, 4 = skip fill entirely
  double2* scratch;  // thal_dimer_thread_kernel: k*k cells per resident thread
  const uint32_t* colperm;  // matrix mode, optional: the order in which a row visits the columns (column_order())
  uint8_t* pairing;  // optional [n_pairs][MSSPE_MAX_OLIGO], zeroed: partner (1-based, in the reversed second oligo) of base i
};

struct DimerShared {  // per block
  double stackS[256], stackH[256], int2S[256], int2H[256], tstS[256], tstH[256];
  double lshS[2][100], lshH[2][100], rshS[2][100], rshH[2][100];
  double interiorS[30], interiorH[30], bulgeS[30], bulgeH[30];
  double atpS[16], atpH[16];
};

__device__ __forceinline__ int i4(int a, int b, int c, int d) { return (a << 6) | (b << 4) | (c << 2) | d; }
__device__ __forceinline__ bool eq2(double a, double b) { return isfinite(a) && isfinite(b) && fabs(a - b) < 1e-5; }

struct PairView {  // per group, in shared memory
  double2* cell;       // [k*k] (S,H)
  uint32_t* rowmask;   // [k+2]: bit (j-1) set <=> n1[i] pairs with n2[j]
  uint8_t* n1; uint8_t* n2;  // [k+2]
  int k;
};

// What a loop candidate needs to know about its closing pair (i,j): computed once per cell, not per candidate.
struct Closing {
  uint32_t a, b;             // n1[i], n2[j]
  double atpS, atpH;         // AT penalty of the closing pair
  double int2S, int2H;       // 1x1 mismatch term on the closing side: stackint2[n2[j]][n2[j-1]][n1[i]][n1[i-1]]
  double tstS, tstH;         // terminal mismatch term on the closing side
};

__device__ __forceinline__ Closing make_closing(const DimerShared& sh, const PairView& pv, int i, int j) {
  Closing c;
  c.a = pv.n1[i]; c.b = pv.n2[j];
  const int x = i4(pv.n2[j], pv.n2[j - 1] & 3, pv.n1[i], pv.n1[i - 1] & 3);  // i, j >= 2 where it is used
  c.atpS = sh.atpS[c.a * 4 + c.b]; c.atpH = sh.atpH[c.a * 4 + c.b];
  c.int2S = sh.int2S[x]; c.int2H = sh.int2H[x];
  c.tstS = sh.tstS[x]; c.tstH = sh.tstH[x];
  return c;
}

// (S,H) of the bulge / internal loop closed by (i,j) with inner pair (ii,jj), including the inner cell's value.
// H = +inf marks "not possible".  ca/cb are the 2-bit codes: n1[t] = (ca >> 2(k-t)) & 3, n2[t] = (cb >> 2(t-1)) & 3.
__device__ __forceinline__ void loop_candidate(const DimerShared& sh, const double2* __restrict__ cell, int k, uint64_t ca, uint64_t cb,
                                               const Closing& cl, int i, int j, int ii, int jj, double* outS, double* outH) {
  const int l1 = i - ii - 1, l2 = j - jj - 1, ls = l1 + l2 - 1;
  const double2 inner = cell[(ii - 1) * k + (jj - 1)];
  const uint32_t a_in = (uint32_t)(ca >> (2 * (k - ii))) & 3u, b_in = (uint32_t)(cb >> (2 * (jj - 1))) & 3u;
  double S, H;
  if (l1 == 0 || l2 == 0) {      // bulge (l1 + l2 >= 1 guaranteed by the caller)
    if (l1 + l2 == 1) {          // size 1: the flanking pairs still stack
      const int x = i4(a_in, cl.a, b_in, cl.b);
      H = sh.bulgeH[ls] + sh.stackH[x];
      S = sh.bulgeS[ls] + sh.stackS[x];
      if (H > 0 || S > 0) { H = INFINITY; S = -1.0; }
      H += inner.y; S += inner.x;
      if (!isfinite(H)) { H = INFINITY; S = -1.0; }
    } else {
      H = sh.bulgeH[ls] + sh.atpH[a_in * 4 + b_in] + cl.atpH;
      H += inner.y;
      S = sh.bulgeS[ls] + sh.atpS[a_in * 4 + b_in] + cl.atpS;
      S += inner.x;
      if (!isfinite(H)) { H = INFINITY; S = -1.0; }
      if (H > 0 && S > 0) { H = INFINITY; S = -1.0; }
    }
  } else {
    // inner-side context: n1[ii], n1[ii+1], n2[jj], n2[jj+1]  (ii+1 <= i-1 and jj+1 <= j-1 here)
    const uint32_t a_nx = (uint32_t)(ca >> (2 * (k - ii - 1))) & 3u, b_nx = (uint32_t)(cb >> (2 * jj)) & 3u;
    const int x = i4(a_in, a_nx, b_in, b_nx);
    if (l1 == 1 && l2 == 1) {
      S = sh.int2S[x] + cl.int2S;
      S += inner.x;
      H = sh.int2H[x] + cl.int2H;
      H += inner.y;
    } else {
      const int asym = l1 > l2 ? l1 - l2 : l2 - l1;
      H = sh.interiorH[ls] + sh.tstH[x] + cl.tstH + (K_ILAH * asym);
      H += inner.y;
      S = sh.interiorS[ls] + sh.tstS[x] + cl.tstS + (K_ILAS * asym);
      S += inner.x;
    }
    if (!isfinite(H)) { H = INFINITY; S = -1.0; }
    if (H > 0 && S > 0) { H = INFINITY; S = -1.0; }
  }
  *outS = S; *outH = H;
}

__device__ __forceinline__ uint64_t revcomp_code(uint64_t code, int k) {
  uint64_t r = 0;
  for (int t = 0; t < k; t++) { r = (r << 2) | (3u - (code & 3u)); code >>= 2; }
  return r;
}

template <int SUB>
__global__ void __launch_bounds__(DIMER_THREADS)
thal_dimer_kernel(const DimerArgs A) {
  constexpr int GROUP = 32;  // one warp per ordered pair
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  DimerShared& sh = *reinterpret_cast<DimerShared*>(dyn_smem);
  const int k = A.k;
  const int tid = threadIdx.x;
  {  // stage tables
    const ThalDeviceTables* T = A.T;
    for (int x = tid; x < 256; x += DIMER_THREADS) {
      const int a = x >> 6, b = (x >> 4) & 3, c = (x >> 2) & 3, d = x & 3;
      const int g = THAL_IDX4(a, b, c, d);
      sh.stackS[x] = T->stackS[g]; sh.stackH[x] = T->stackH[g];
      sh.int2S[x] = T->stackint2S[g]; sh.int2H[x] = T->stackint2H[g];
      sh.tstS[x] = T->tstackS[g]; sh.tstH[x] = T->tstackH[g];
    }
    for (int x = tid; x < 200; x += DIMER_THREADS) {
      (&sh.lshS[0][0])[x] = (&A.C->lshS[0][0])[x]; (&sh.lshH[0][0])[x] = (&A.C->lshH[0][0])[x];
      (&sh.rshS[0][0])[x] = (&A.C->rshS[0][0])[x]; (&sh.rshH[0][0])[x] = (&A.C->rshH[0][0])[x];
    }
    for (int x = tid; x < 30; x += DIMER_THREADS) {
      sh.interiorS[x] = T->interiorS[x]; sh.interiorH[x] = T->interiorH[x];
      sh.bulgeS[x] = T->bulgeS[x]; sh.bulgeH[x] = T->bulgeH[x];
    }
    for (int x = tid; x < 16; x += DIMER_THREADS) { sh.atpS[x] = T->atpS[(x >> 2) * 5 + (x & 3)]; sh.atpH[x] = T->atpH[(x >> 2) * 5 + (x & 3)]; }
  }
  __syncthreads();
  constexpr int GROUPS = DIMER_THREADS / GROUP;
  const int grp = tid / GROUP, gl = tid % GROUP, lane = tid & 31;
  const unsigned gmask = GROUP == 32 ? 0xffffffffu : (0xFFFFu << (lane & 16));
  const size_t cell_bytes = (size_t)k * k * sizeof(double2);
  const size_t grp_bytes = (cell_bytes + (size_t)(k + 2) * 4 + 2 * (size_t)(k + 2) + 15) & ~(size_t)15;
  unsigned char* gbase = dyn_smem + ((sizeof(DimerShared) + 15) & ~(size_t)15) + grp * grp_bytes;
  PairView pv;
  pv.cell = reinterpret_cast<double2*>(gbase);
  pv.rowmask = reinterpret_cast<uint32_t*>(gbase + cell_bytes);
  pv.n1 = reinterpret_cast<uint8_t*>(pv.rowmask + (k + 2));
  pv.n2 = pv.n1 + (k + 2);
  pv.k = k;
  const int maxLoop = A.C->maxLoop;
  const double saltCorr = A.C->saltCorr, t_user = A.C->t_user_K;

  for (unsigned long long p = (unsigned long long)blockIdx.x * GROUPS + grp; p < A.n_pairs; p += (unsigned long long)gridDim.x * GROUPS) {
    uint64_t ca, cb;
    uint32_t col = 0;
    if (A.matrix) { col = A.colperm ? A.colperm[p % A.n] : (uint32_t)(p % A.n); ca = A.a[A.row_begin + p / A.n]; cb = A.b[col]; }
    else { ca = A.a[p]; cb = A.b[p]; }
    __syncwarp(gmask);
    // numSeq1 = oligo1 5'->3'; numSeq2 = oligo2 REVERSED (not complemented); N sentinels at both ends
    for (int t = gl; t < k; t += GROUP) {
      pv.n1[t + 1] = (uint8_t)((ca >> (2 * (k - 1 - t))) & 3u);
      pv.n2[t + 1] = (uint8_t)((cb >> (2 * t)) & 3u);
    }
    if (gl == 0) { pv.n1[0] = 4; pv.n1[k + 1] = 4; pv.n2[0] = 4; pv.n2[k + 1] = 4; }
    __syncwarp(gmask);
    {  // rowmask[i] = positions j whose base pairs with n1[i]
      uint32_t cm[4] = {0u, 0u, 0u, 0u};
      for (int j = 1; j <= k; j++) cm[pv.n2[j]] |= 1u << (j - 1);
      for (int i = gl + 1; i <= k; i += GROUP) pv.rowmask[i] = cm[3 - pv.n1[i]];
      if (gl == 0) { pv.rowmask[0] = 0; pv.rowmask[k + 1] = 0; }
    }
    const int sym = ((k & 1) == 0 && revcomp_code(ca, k) == ca && revcomp_code(cb, k) == cb) ? 1 : 0;
    const double RC = A.C->RC[sym];
    const double* lshS = sh.lshS[sym]; const double* lshH = sh.lshH[sym];
    const double* rshS = sh.rshS[sym]; const double* rshH = sh.rshH[sym];
    __syncwarp(gmask);

    // ---------------- fill ----------------
    for (int i = 1; i <= ((A.dbg & 4) ? 0 : k); i++) {
      const uint32_t rm = pv.rowmask[i];
      const int a = pv.n1[i];
      for (int j = gl + 1; j <= k; j += GROUP) {
        if (!((rm >> (j - 1)) & 1u)) continue;
        const int li = (a * 5 + pv.n1[i - 1]) * 5 + pv.n2[j - 1];
        double S = lshS[li], H = lshH[li];
        if (i > 1 && j > 1) {
          const int ri = (a * 5 + pv.n1[i + 1]) * 5 + pv.n2[j + 1];
          const double rS = rshS[ri], rH = rshH[ri];
          double S0 = S, H0 = H, S1, H1, T1;
          const double T0 = (H0 + kDHi + rH) / (S0 + kDSi + rS + RC);
          const int si = i4(pv.n1[i - 1], a, pv.n2[j - 1], pv.n2[j]);
          const double stH = sh.stackH[si];
          const bool prev_bp = (pv.rowmask[i - 1] >> (j - 2)) & 1u;
          if (prev_bp && isfinite(stH)) {
            const double2 pc = pv.cell[(i - 2) * k + (j - 2)];
            S1 = pc.x + sh.stackS[si];
            H1 = pc.y + stH;
            T1 = (H1 + kDHi + rH) / (S1 + kDSi + rS + RC);
          } else {
            S1 = -1.0; H1 = INFINITY;
            T1 = (H1 + kDHi) / (S1 + kDSi + RC);
          }
          if (S1 < kMinEntropyCutoff) { S1 = kMinEntropy; H1 = 0.0; }
          if (S0 < kMinEntropyCutoff) { S0 = kMinEntropy; H0 = 0.0; }
          if (T1 > T0) { S = S1; H = H1; } else if (T0 >= T1) { S = S0; H = H0; }
        }
        pv.cell[(i - 1) * k + (j - 1)] = make_double2(S, H);
      }
      __syncwarp(gmask);
      if (i > 1 && !(A.dbg & 1)) {
        // CELLS paired cells of this row are processed at once, SUB lanes each (lane = inner row of the candidates)
        constexpr int CELLS = 32 / SUB;
        const int slot = lane / SUB, sl = lane % SUB;
        for (uint32_t bj = rm & ~1u; bj;) {
          uint32_t t = bj;
          int j = 0;
#pragma unroll
          for (int q = 0; q < CELLS; q++) { const int f = t ? __ffs(t) : 0; if (q == slot) j = f; t &= t - 1u; }
          bj = t;  // uniform: every lane strips the same CELLS lowest bits
          const bool has = j != 0;
          if (!has) j = 2;
          const int ri = (a * 5 + pv.n1[i + 1]) * 5 + pv.n2[j + 1];
          const double rS = rshS[ri], rH = rshH[ri];
          const double2 cur = pv.cell[(i - 1) * k + (j - 1)];
          const double Gcur = cur.y + rH - kTK * (cur.x + rS);
          const Closing cl = make_closing(sh, pv, i, j);
          double bG = INFINITY, bS = -1.0, bH = INFINITY;
          int bkey = 0x7fffffff;
          bool clamp = false;
          if (has) {
            for (int l1 = sl; l1 <= i - 2; l1 += SUB) {
              const int ii = i - 1 - l1;
              uint32_t cand = pv.rowmask[ii] & ((1u << (j - 1)) - 1u);
              if (l1 == 0) cand &= ~(1u << (j - 2));
              while (cand) {
                const int jj = __ffs(cand);
                cand &= cand - 1u;
                const int l2 = j - jj - 1;
                if (l1 + l2 > maxLoop) continue;
                double S, H;
                loop_candidate(sh, pv.cell, k, ca, cb, cl, i, j, ii, jj, &S, &H);
                if (isfinite(H)) {
                  if (S < kMinEntropyCutoff) clamp = true;
                  const double G1 = H + rH - kTK * (S + rS);
                  const int key = (l1 + l2) * 64 + l1;
                  if (G1 < bG || (G1 == bG && key < bkey)) { bG = G1; bkey = key; bS = S; bH = H; }
                }
              }
            }
          }
          if (__any_sync(0xffffffffu, clamp)) {
            // An entropy below the cutoff re-bases the cell mid-scan: replay the scan sequentially, one lane per cell.
            if (has && sl == 0) {
              double cS = cur.x, cH = cur.y;
              for (int d = 3; d <= maxLoop + 2; d++) {
                int ii = i - 1, jj = -ii - d + (j + i);
                if (jj < 1) { ii -= (1 - jj); jj = 1; }
                for (; ii > 0 && jj < j; --ii, ++jj) {
                  if (!((pv.rowmask[ii] >> (jj - 1)) & 1u)) continue;
                  double S, H;
                  loop_candidate(sh, pv.cell, k, ca, cb, cl, i, j, ii, jj, &S, &H);
                  const double G1 = H + rH - kTK * (S + rS), G2 = cH + rH - kTK * (cS + rS);
                  if (!(G1 < G2)) { S = -1.0; H = INFINITY; }
                  if (S < kMinEntropyCutoff) { S = kMinEntropy; H = 0.0; }
                  if (isfinite(H)) { cS = S; cH = H; }
                }
              }
              pv.cell[(i - 1) * k + (j - 1)] = make_double2(cS, cH);
            }
          } else {
            double mG = bG;
#pragma unroll
            for (int o = SUB / 2; o > 0; o >>= 1) mG = fmin(mG, __shfl_xor_sync(0xffffffffu, mG, o, SUB));
            int mk = (bG == mG && mG < Gcur) ? bkey : 0x7fffffff;
#pragma unroll
            for (int o = SUB / 2; o > 0; o >>= 1) mk = min(mk, __shfl_xor_sync(0xffffffffu, mk, o, SUB));
            if (has && mG < Gcur && bG == mG && bkey == mk) pv.cell[(i - 1) * k + (j - 1)] = make_double2(bS, bH);
          }
        }
        __syncwarp(gmask);
      }
    }

    // ---------------- best terminal pair ----------------
    double bG = INFINITY; int bkey = 0x7fffffff;
    {
      const int i_lo = A.type == MSSPE_THAL_ANY ? 1 : k;
      for (int i = i_lo + gl; i <= k; i += GROUP) {
        const int a = pv.n1[i];
        for (uint32_t bj = pv.rowmask[i]; bj; bj &= bj - 1u) {
          const int j = __ffs(bj);
          const int ri = (a * 5 + pv.n1[i + 1]) * 5 + pv.n2[j + 1];
          double rS = rshS[ri], rH = rshH[ri];
          rS = rS + kSmallNonZero; rH = rH + kSmallNonZero;
          const double2 c = pv.cell[(i - 1) * k + (j - 1)];
          const double G1 = (c.y + rH + kDHi) - kTK * (c.x + rS + kDSi);
          const int key = i * 64 + j;
          if (G1 < bG || (G1 == bG && key < bkey)) { bG = G1; bkey = key; }
        }
      }
      double mG = bG;
#pragma unroll
      for (int o = GROUP / 2; o > 0; o >>= 1) mG = fmin(mG, __shfl_xor_sync(gmask, mG, o, GROUP));
      int mk = (bG == mG && isfinite(bG)) ? bkey : 0x7fffffff;
#pragma unroll
      for (int o = GROUP / 2; o > 0; o >>= 1) mk = min(mk, __shfl_xor_sync(gmask, mk, o, GROUP));
      bG = mG; bkey = mk;
    }
    const bool none = !isfinite(bG);  // no base pair anywhere (for END1: none in the last row)
    int bi = none ? (A.type == MSSPE_THAL_ANY ? 1 : k) : (bkey >> 6);
    int bjx = none ? 1 : (bkey & 63);
    if (none && A.type != MSSPE_THAL_ANY) { bi = 1; bjx = 1; }  // `if (!isFinite(bestG)) bestI = bestJ = 1`
    const bool has_struct = (pv.rowmask[bi] >> (bjx - 1)) & 1u;
    msspe_thal_out res;
    res.ds = 0; res.dh = 0; res.dg = 0; res.tm = 0; res.no_structure = 1; res.n_bp = 0;
    if (has_struct) {
      const int ri = (pv.n1[bi] * 5 + pv.n1[bi + 1]) * 5 + pv.n2[bjx + 1];
      const double2 bc = pv.cell[(bi - 1) * k + (bjx - 1)];
      const double dH = bc.y + rshH[ri] + kDHi;
      const double dS = bc.x + rshS[ri] + kDSi;
      // ---------------- traceback: count paired positions ----------------
      int i = bi, j = bjx, pairs = 1;
      if (A.pairing && gl == 0) A.pairing[p * MSSPE_MAX_OLIGO + (i - 1)] = (uint8_t)j;
      for (int guard = 0; guard < ((A.dbg & 2) ? 0 : 2 * k + 2); guard++) {
        const int li = (pv.n1[i] * 5 + pv.n1[i - 1]) * 5 + pv.n2[j - 1];
        const double2 c = pv.cell[(i - 1) * k + (j - 1)];
        if (eq2(c.x, lshS[li]) && eq2(c.y, lshH[li])) break;
        if (i > 1 && j > 1 && ((pv.rowmask[i - 1] >> (j - 2)) & 1u)) {
          const int si = i4(pv.n1[i - 1], pv.n1[i], pv.n2[j - 1], pv.n2[j]);
          const double2 pc = pv.cell[(i - 2) * k + (j - 2)];
          if (eq2(c.x, sh.stackS[si] + pc.x) && eq2(c.y, sh.stackH[si] + pc.y)) {
            i--; j--; pairs++;
            if (A.pairing && gl == 0) A.pairing[p * MSSPE_MAX_OLIGO + (i - 1)] = (uint8_t)j;
            continue;
          }
        }
        int key = 0x7fffffff;
        const Closing cl = make_closing(sh, pv, i > 1 ? i : 2, j > 1 ? j : 2);  // only used when i, j >= 2
        for (int l1 = gl; l1 <= i - 2; l1 += GROUP) {
          const int ii = i - 1 - l1;
          uint32_t cand = j >= 2 ? (pv.rowmask[ii] & ((1u << (j - 1)) - 1u)) : 0u;
          if (l1 == 0 && j >= 2) cand &= ~(1u << (j - 2));
          while (cand) {
            const int jj = __ffs(cand);
            cand &= cand - 1u;
            const int l2 = j - jj - 1;
            if (l1 + l2 > maxLoop) continue;
            double S, H;
            loop_candidate(sh, pv.cell, k, ca, cb, cl, i, j, ii, jj, &S, &H);
            if (eq2(c.x, S) && eq2(c.y, H)) key = min(key, (l1 + l2) * 64 + l1);
          }
        }
#pragma unroll
        for (int o = GROUP / 2; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(gmask, key, o, GROUP));
        if (key == 0x7fffffff) break;
        const int l1 = key & 63, l2 = (key >> 6) - l1;
        i = i - 1 - l1; j = j - 1 - l2; pairs++;
        if (A.pairing && gl == 0) A.pairing[p * MSSPE_MAX_OLIGO + (i - 1)] = (uint8_t)j;
      }
      const int N = pairs - 1;  // (#paired bases in both strands)/2 - 1
      const double t = (dH / (dS + (N * saltCorr) + RC)) - kAbsZero;
      res.dg = dH - (t_user * (dS + (N * saltCorr)));
      res.ds = dS + (N * saltCorr);
      res.dh = dH;
      res.tm = t;
      res.no_structure = 0;
      res.n_bp = pairs;
    }
    if (gl == 0) {
      if (A.out) A.out[p] = res;
      if (A.matrix) {
        const unsigned long long pair = (unsigned long long)(A.row_begin + p / A.n) * A.n + col;
        if (res.no_structure) {
          const unsigned long long at = atomicAdd(A.n_nostruct, 1ull);
          if (at < A.nostruct_cap) A.nostruct[at] = pair;
        } else if (res.dg < A.dg_limit) {
          const unsigned long long at = atomicAdd(A.n_edges, 1ull);
          if (at < A.edge_cap) { A.edges[at].pair = pair; A.edges[at].dg = res.dg; }
        }
      }
    }
  }
}

// ---------------------------------------------------------------- device: dimer, flat candidate enumeration
// Same algorithm and the same results as thal_dimer_kernel; what changes is how the bulge / internal-loop candidates of a
// row are spread over the warp, and how many instructions one candidate costs.
//   * thal_dimer_kernel gives every paired cell 8 (16) lanes that stride over the inner ROWS and loop over the paired
//     columns of each: 5.5 of 32 lanes are active in that loop, 12,000 warp instructions per 13-mer pair (ncu).
//   * Here the paired cells of the rows above row i are listed once per row in COLUMN-major order (desc[]); the
//     candidates of cell (i, j) are then exactly the first pcol[j - 1] entries of that list, so the candidates of ALL paired
//     cells of the row form one flat index space (j-major, oend[] = running ends) that the 32 lanes take 32 at a time: an
//     item finds its cell by a binary search in oend[].  18 rounds per 13-mer pair instead of 57.
//   * The four candidate kinds (bulge of one base, longer bulge, 1x1 mismatch, general interior loop) are ONE arithmetic
//     sequence over a single (S,H) table: three entries chosen by integer selects, zeros where a kind has fewer terms, so
//     the lanes of a round do not diverge by kind.  The sums keep the scalar code's association.
//   * The minimum per cell: the lanes of one cell are consecutive, so each lane passes ITS cell's lane range as the member
//     mask of three integer REDUX (ordered bits of dG high word, low word, scan-order key) -- all cells of the round at once.
//   * The first of the two quotients of the stack test depends only on the 2x2x2 base context: tabulated on the host
//     (ThalDimerConsts::t0) with the same expression, so a row pays for one FP64 division, not two.
constexpr int FT_STACK = 0, FT_INT2 = 256, FT_TST = 512, FT_ATP = 768, FT_BULGE = 784, FT_INTERIOR = 814, FT_ZERO = 844, FT_N = 845;
struct FlatShared {   // per block
  double2 tab[FT_N];  // (S,H): stack | 1x1 mismatch | terminal mismatch (256 each, index i4) | AT penalty (16) | bulge, interior (30 each) | zero
  double2 lsh[2][100], rsh[2][100];
};
struct FlatView {   // per warp, in shared memory
  double2* cell;                                          // [k*k] (S,H)
  double *bG, *bS, *bH;                                   // [k+2] running best candidate of the row's cells
  uint32_t *rowmask, *colmask, *pcol, *oend; int* bKey;   // [k+2]
  uint32_t* desc;                                         // [k*k] (inner context << 16) | (ii << 8) | jj
  uint8_t *n1, *n2, *clx, *rix, *actx, *bctx;             // [k+2]
};
__host__ __device__ inline size_t flat_group_bytes(int k) {
  const size_t b = (size_t)k * k * 16 + 3 * (size_t)(k + 2) * 8 + 5 * (size_t)(k + 2) * 4 + (size_t)k * k * 4 + 6 * (size_t)(k + 2);
  return (b + 15) & ~(size_t)15;
}

// (S,H) of the bulge / internal loop closed by (i,j) with inner pair (ii,jj), including the inner cell's value; H = +inf
// marks "not possible".  xin = i4(n1[ii], n1[ii+1], n2[jj], n2[jj+1]), xcl = i4(n2[j], n2[j-1], n1[i], n1[i-1]).
__device__ __forceinline__ void flat_candidate(const double2* __restrict__ tab, const double2 inner, uint32_t xin, uint32_t xcl, int l1, int l2,
                                               double* outS, double* outH) {
  const int ls = l1 + l2 - 1;
  const bool bulge = (l1 == 0) | (l2 == 0);
  const bool b1 = bulge & (ls == 0);          // one bulged base: the flanking pairs still stack
  const bool one = (l1 == 1) & (l2 == 1);
  const uint32_t a_in = xin >> 6, b_in = (xin >> 2) & 3u, a_cl = (xcl >> 2) & 3u, b_cl = xcl >> 6;
  const uint32_t x1 = (a_in << 6) | (a_cl << 4) | (b_in << 2) | b_cl;
  const uint32_t iA = bulge ? FT_BULGE + ls : (one ? FT_INT2 + xin : FT_INTERIOR + ls);
  const uint32_t iB = b1 ? FT_STACK + x1 : (bulge ? FT_ATP + a_in * 4 + b_in : (one ? FT_INT2 + xcl : FT_TST + xin));
  const uint32_t iC = (b1 | one) ? FT_ZERO : (bulge ? FT_ATP + a_cl * 4 + b_cl : FT_TST + xcl);
  const double2 tA = tab[iA], tB = tab[iB], tC = tab[iC];
  const int asym = l1 > l2 ? l1 - l2 : l2 - l1;
  const bool gen = !bulge && !one;
  const double dH = gen ? (K_ILAH * asym) : 0.0, dS = gen ? (K_ILAS * asym) : 0.0;
  double H = tA.y + tB.y + tC.y + dH;
  double S = tA.x + tB.x + tC.x + dS;
  if (b1 && (H > 0 || S > 0)) { H = INFINITY; S = -1.0; }
  H += inner.y; S += inner.x;
  if (!isfinite(H)) { H = INFINITY; S = -1.0; }
  if (!b1 && H > 0 && S > 0) { H = INFINITY; S = -1.0; }
  *outS = S; *outH = H;
}

__global__ void __launch_bounds__(DIMER_THREADS, 5)
thal_dimer_flat_kernel(const DimerArgs A) {
  constexpr int GROUP = 32;  // one warp per ordered pair
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  FlatShared& sh = *reinterpret_cast<FlatShared*>(dyn_smem);
  const int k = A.k;
  const int tid = threadIdx.x;
  {  // stage tables
    const ThalDeviceTables* T = A.T;
    for (int x = tid; x < 256; x += DIMER_THREADS) {
      const int a = x >> 6, b = (x >> 4) & 3, c = (x >> 2) & 3, d = x & 3;
      const int g = THAL_IDX4(a, b, c, d);
      sh.tab[FT_STACK + x] = make_double2(T->stackS[g], T->stackH[g]);
      sh.tab[FT_INT2 + x] = make_double2(T->stackint2S[g], T->stackint2H[g]);
      sh.tab[FT_TST + x] = make_double2(T->tstackS[g], T->tstackH[g]);
    }
    for (int x = tid; x < 200; x += DIMER_THREADS) {
      (&sh.lsh[0][0])[x] = make_double2((&A.C->lshS[0][0])[x], (&A.C->lshH[0][0])[x]);
      (&sh.rsh[0][0])[x] = make_double2((&A.C->rshS[0][0])[x], (&A.C->rshH[0][0])[x]);
    }
    for (int x = tid; x < 30; x += DIMER_THREADS) {
      sh.tab[FT_INTERIOR + x] = make_double2(T->interiorS[x], T->interiorH[x]);
      sh.tab[FT_BULGE + x] = make_double2(T->bulgeS[x], T->bulgeH[x]);
    }
    for (int x = tid; x < 16; x += DIMER_THREADS) sh.tab[FT_ATP + x] = make_double2(T->atpS[(x >> 2) * 5 + (x & 3)], T->atpH[(x >> 2) * 5 + (x & 3)]);
    if (tid == 0) sh.tab[FT_ZERO] = make_double2(0.0, 0.0);
  }
  __syncthreads();
  const double2* __restrict__ tab = sh.tab;
  constexpr int GROUPS = DIMER_THREADS / GROUP;
  const int grp = tid / GROUP, gl = tid % GROUP, lane = tid & 31;
  const unsigned gmask = 0xffffffffu;
  FlatView fv;
  {
    unsigned char* gb = dyn_smem + ((sizeof(FlatShared) + 15) & ~(size_t)15) + grp * flat_group_bytes(k);
    const size_t n8 = (size_t)(k + 2) * 8, n4 = (size_t)(k + 2) * 4, n1b = (size_t)(k + 2);
    fv.cell = reinterpret_cast<double2*>(gb); gb += (size_t)k * k * 16;
    fv.bG = reinterpret_cast<double*>(gb); gb += n8;
    fv.bS = reinterpret_cast<double*>(gb); gb += n8;
    fv.bH = reinterpret_cast<double*>(gb); gb += n8;
    fv.rowmask = reinterpret_cast<uint32_t*>(gb); gb += n4;
    fv.colmask = reinterpret_cast<uint32_t*>(gb); gb += n4;
    fv.pcol = reinterpret_cast<uint32_t*>(gb); gb += n4;
    fv.oend = reinterpret_cast<uint32_t*>(gb); gb += n4;
    fv.bKey = reinterpret_cast<int*>(gb); gb += n4;
    fv.desc = reinterpret_cast<uint32_t*>(gb); gb += (size_t)k * k * 4;
    fv.n1 = gb; gb += n1b; fv.n2 = gb; gb += n1b; fv.clx = gb; gb += n1b; fv.rix = gb; gb += n1b; fv.actx = gb; gb += n1b; fv.bctx = gb;
  }
  double2* const cell = fv.cell;
  const uint8_t* const n1 = fv.n1; const uint8_t* const n2 = fv.n2;
  const int maxLoop = A.C->maxLoop;
  const double saltCorr = A.C->saltCorr, t_user = A.C->t_user_K;

  for (unsigned long long p = (unsigned long long)blockIdx.x * GROUPS + grp; p < A.n_pairs; p += (unsigned long long)gridDim.x * GROUPS) {
    uint64_t ca, cb;
    uint32_t col = 0;
    if (A.matrix) { col = A.colperm ? A.colperm[p % A.n] : (uint32_t)(p % A.n); ca = A.a[A.row_begin + p / A.n]; cb = A.b[col]; }
    else { ca = A.a[p]; cb = A.b[p]; }
    __syncwarp(gmask);
    // numSeq1 = oligo1 5'->3'; numSeq2 = oligo2 REVERSED (not complemented); N sentinels at both ends
    for (int t = gl; t < k; t += GROUP) {
      fv.n1[t + 1] = (uint8_t)((ca >> (2 * (k - 1 - t))) & 3u);
      fv.n2[t + 1] = (uint8_t)((cb >> (2 * t)) & 3u);
    }
    if (gl == 0) { fv.n1[0] = 4; fv.n1[k + 1] = 4; fv.n2[0] = 4; fv.n2[k + 1] = 4; }
    __syncwarp(gmask);
    {  // rowmask[i] = columns j whose base pairs with n1[i]; colmask[j] = rows i whose base pairs with n2[j]; neighbour contexts
      uint32_t cm0 = 0u, cm1 = 0u, cm2 = 0u, cm3 = 0u, rm0 = 0u, rm1 = 0u, rm2 = 0u, rm3 = 0u;   // positions of A, C, G, T in oligo 2 (reversed) / oligo 1
      for (int j = 1; j <= k; j++) {
        const uint32_t bit = 1u << (j - 1), y = n2[j], x = n1[j];
        cm0 |= y == 0u ? bit : 0u; cm1 |= y == 1u ? bit : 0u; cm2 |= y == 2u ? bit : 0u; cm3 |= y == 3u ? bit : 0u;
        rm0 |= x == 0u ? bit : 0u; rm1 |= x == 1u ? bit : 0u; rm2 |= x == 2u ? bit : 0u; rm3 |= x == 3u ? bit : 0u;
      }
      for (int i = gl + 1; i <= k; i += GROUP) {
        const uint32_t x = n1[i], y = n2[i];
        fv.rowmask[i] = x == 0u ? cm3 : x == 1u ? cm2 : x == 2u ? cm1 : cm0;
        fv.colmask[i] = y == 0u ? rm3 : y == 1u ? rm2 : y == 2u ? rm1 : rm0;
        fv.actx[i] = (uint8_t)((x << 6) | ((n1[i + 1] & 3) << 4));
        fv.bctx[i] = (uint8_t)((y << 2) | (n2[i + 1] & 3));
      }
      if (gl == 0) { fv.rowmask[0] = 0; fv.rowmask[k + 1] = 0; fv.colmask[0] = 0; fv.colmask[k + 1] = 0; }
    }
    const int sym = ((k & 1) == 0 && revcomp_code(ca, k) == ca && revcomp_code(cb, k) == cb) ? 1 : 0;
    const double RC = A.C->RC[sym];
    const double2* __restrict__ lsh = sh.lsh[sym];
    const double2* __restrict__ rsh = sh.rsh[sym];
    const double* __restrict__ t0tab = A.C->t0[sym];
    __syncwarp(gmask);

    // ---------------- fill ----------------
    for (int i = 1; i <= ((A.dbg & 4) ? 0 : k); i++) {
      const uint32_t rm = fv.rowmask[i];
      const int a = n1[i];
      for (int j = gl + 1; j <= k; j += GROUP) {
        if (!((rm >> (j - 1)) & 1u)) continue;
        const int li = (a * 5 + n1[i - 1]) * 5 + n2[j - 1];
        const double2 l = lsh[li];
        double S = l.x, H = l.y;
        if (i > 1 && j > 1) {
          const int rn = n1[i + 1] * 5 + n2[j + 1];
          const int ri = a * 25 + rn;
          const double2 r = rsh[ri];
          const double rS = r.x, rH = r.y;
          double S0 = S, H0 = H, S1, H1, T1;
          const double T0 = __ldg(&t0tab[li * 25 + rn]);   // (H0 + kDHi + rH) / (S0 + kDSi + rS + RC), tabulated on the host
          const int si = i4(n1[i - 1], a, n2[j - 1], n2[j]);
          const double2 st = tab[FT_STACK + si];
          const bool prev_bp = (fv.rowmask[i - 1] >> (j - 2)) & 1u;
          if (prev_bp && isfinite(st.y)) {
            const double2 pc = cell[(i - 2) * k + (j - 2)];
            S1 = pc.x + st.x;
            H1 = pc.y + st.y;
            T1 = (H1 + kDHi + rH) / (S1 + kDSi + rS + RC);
          } else {
            S1 = -1.0; H1 = INFINITY;
            T1 = (H1 + kDHi) / (S1 + kDSi + RC);
          }
          if (S1 < kMinEntropyCutoff) { S1 = kMinEntropy; H1 = 0.0; }
          if (S0 < kMinEntropyCutoff) { S0 = kMinEntropy; H0 = 0.0; }
          if (T1 > T0) { S = S1; H = H1; } else if (T0 >= T1) { S = S0; H = H0; }
          fv.clx[j] = (uint8_t)i4(n2[j], n2[j - 1] & 3, a, n1[i - 1] & 3);
          fv.rix[j] = (uint8_t)ri;
          fv.bG[j] = INFINITY; fv.bKey[j] = 0x7fffffff;
        }
        cell[(i - 1) * k + (j - 1)] = make_double2(S, H);
      }
      __syncwarp(gmask);
      const uint32_t bjm = rm & ~1u;   // paired columns j >= 2 of this row
      if (i > 1 && bjm && !(A.dbg & 1)) {
        // the paired cells of rows 1 .. i-1, column-major: lane c lists column c + 1 behind the columns before it;
        // oend[c] = end of the candidates of column c + 1 in the row's flat (j-major) index space
        uint32_t T;
        {
          const uint32_t cm = lane < k ? (fv.colmask[lane + 1] & ((1u << (i - 1)) - 1u)) : 0u;
          const uint32_t cc = __popc(cm);
          uint32_t inc = cc;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
          uint32_t r = inc - cc;                    // cells in columns 1 .. lane  ==  candidates of a cell in column lane + 1
          uint32_t oe = ((bjm >> lane) & 1u) ? r : 0u;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, oe, o); if (lane >= o) oe += t; }
          if (lane < k) { fv.pcol[lane] = r; fv.oend[lane] = oe; }
          T = __shfl_sync(0xffffffffu, oe, 31);
          const uint32_t bx = lane < k ? fv.bctx[lane + 1] : 0u;
          for (uint32_t m = cm; m; m &= m - 1u) {
            const uint32_t ii = __ffs(m);
            fv.desc[r++] = ((uint32_t)(fv.actx[ii] | bx) << 16) | (ii << 8) | (uint32_t)(lane + 1);
          }
        }
        __syncwarp(gmask);
        bool clamp = false;
        for (uint32_t base = 0; base < T; base += 32) {
          const uint32_t t = base + lane;
          double S = -1.0, H = INFINITY, G1 = INFINITY;
          int key = 0x7fffffff, j = 0;
          bool fin = false;
          unsigned segmask = 1u << lane;
          if (t < T) {
            uint32_t c = 0;         // number of columns whose candidates end at or before item t  ==  the item's column - 1
#pragma unroll
            for (int s = 16; s; s >>= 1) { const uint32_t m = c + s; if (m <= (uint32_t)k && fv.oend[m - 1] <= t) c = m; }
            j = (int)c + 1;
            const uint32_t cnt = fv.pcol[c], end = fv.oend[c], beg = end - cnt;
            const uint32_t lb = beg > base ? beg - base : 0u, le = end - base;     // the cell's lanes in this round: [lb, le)
            segmask = (le >= 32u ? 0xffffffffu : ((1u << le) - 1u)) & ~((1u << lb) - 1u);
            const uint32_t d = fv.desc[t - beg];
            const int ii = (d >> 8) & 0xff, jj = d & 0xff;
            const int l1 = i - ii - 1, l2 = j - jj - 1;
            if (l1 + l2 >= 1 && l1 + l2 <= maxLoop) {
              flat_candidate(tab, cell[(ii - 1) * k + (jj - 1)], d >> 16, fv.clx[j], l1, l2, &S, &H);
              if (isfinite(H)) {
                if (S < kMinEntropyCutoff) clamp = true;
                const double2 r = rsh[fv.rix[j]];
                G1 = H + r.y - kTK * (S + r.x);
                key = (l1 + l2) * 64 + l1;
                fin = true;
              }
            }
          }
          // minimum per cell: first dG (unsigned order of the ordered bits == order of the doubles), then the scan-order key
          const unsigned long long gbits = (unsigned long long)__double_as_longlong(G1);
          const unsigned long long ob = gbits ^ ((gbits >> 63) ? ~0ull : 0x8000000000000000ull);
          const uint32_t oh = fin ? (uint32_t)(ob >> 32) : 0xffffffffu;
          const uint32_t mh = __reduce_min_sync(segmask, oh);
          const bool c1 = fin && oh == mh;
          const uint32_t ol = c1 ? (uint32_t)ob : 0xffffffffu;
          const uint32_t ml = __reduce_min_sync(segmask, ol);
          const bool c2 = c1 && ol == ml;
          const uint32_t mk = __reduce_min_sync(segmask, c2 ? (uint32_t)key : 0xffffffffu);
          if (c2 && (uint32_t)key == mk) {
            const double bG = fv.bG[j];
            if (G1 < bG || (G1 == bG && key < fv.bKey[j])) { fv.bG[j] = G1; fv.bS[j] = S; fv.bH[j] = H; fv.bKey[j] = key; }
          }
          __syncwarp(gmask);
        }
        const bool cellmine = lane < k && ((bjm >> lane) & 1u);
        const int j = lane + 1;
        if (__any_sync(0xffffffffu, clamp)) {
          // An entropy below the cutoff re-bases the cell mid-scan: replay the scan sequentially, one lane per cell.
          if (cellmine) {
            const double2 r = rsh[fv.rix[j]];
            const double rS = r.x, rH = r.y;
            const double2 cur = cell[(i - 1) * k + (j - 1)];
            const uint32_t xcl = fv.clx[j];
            double cS = cur.x, cH = cur.y;
            for (int d = 3; d <= maxLoop + 2; d++) {
              int ii = i - 1, jj = -ii - d + (j + i);
              if (jj < 1) { ii -= (1 - jj); jj = 1; }
              for (; ii > 0 && jj < j; --ii, ++jj) {
                if (!((fv.rowmask[ii] >> (jj - 1)) & 1u)) continue;
                double S, H;
                flat_candidate(tab, cell[(ii - 1) * k + (jj - 1)], (uint32_t)(fv.actx[ii] | fv.bctx[jj]), xcl, i - ii - 1, j - jj - 1, &S, &H);
                const double G1 = H + rH - kTK * (S + rS), G2 = cH + rH - kTK * (cS + rS);
                if (!(G1 < G2)) { S = -1.0; H = INFINITY; }
                if (S < kMinEntropyCutoff) { S = kMinEntropy; H = 0.0; }
                if (isfinite(H)) { cS = S; cH = H; }
              }
            }
            cell[(i - 1) * k + (j - 1)] = make_double2(cS, cH);
          }
        } else if (cellmine) {
          const double2 r = rsh[fv.rix[j]];
          const double2 cur = cell[(i - 1) * k + (j - 1)];
          const double Gcur = cur.y + r.y - kTK * (cur.x + r.x);
          if (fv.bG[j] < Gcur) cell[(i - 1) * k + (j - 1)] = make_double2(fv.bS[j], fv.bH[j]);
        }
        __syncwarp(gmask);
      }
    }

    // ---------------- best terminal pair ----------------
    double bG = INFINITY; int bkey = 0x7fffffff;
    {
      const int i_lo = A.type == MSSPE_THAL_ANY ? 1 : k;
      for (int i = i_lo + gl; i <= k; i += GROUP) {
        const int a = n1[i];
        for (uint32_t bj = fv.rowmask[i]; bj; bj &= bj - 1u) {
          const int j = __ffs(bj);
          const int ri = (a * 5 + n1[i + 1]) * 5 + n2[j + 1];
          const double2 r = rsh[ri];
          const double rS = r.x + kSmallNonZero, rH = r.y + kSmallNonZero;
          const double2 c = cell[(i - 1) * k + (j - 1)];
          const double G1 = (c.y + rH + kDHi) - kTK * (c.x + rS + kDSi);
          const int key = i * 64 + j;
          if (G1 < bG || (G1 == bG && key < bkey)) { bG = G1; bkey = key; }
        }
      }
      double mG = bG;
#pragma unroll
      for (int o = GROUP / 2; o > 0; o >>= 1) mG = fmin(mG, __shfl_xor_sync(gmask, mG, o, GROUP));
      int mk = (bG == mG && isfinite(bG)) ? bkey : 0x7fffffff;
      mk = (int)__reduce_min_sync(gmask, (uint32_t)mk);
      bG = mG; bkey = mk;
    }
    const bool none = !isfinite(bG);  // no base pair anywhere (for END1: none in the last row)
    int bi = none ? (A.type == MSSPE_THAL_ANY ? 1 : k) : (bkey >> 6);
    int bjx = none ? 1 : (bkey & 63);
    if (none && A.type != MSSPE_THAL_ANY) { bi = 1; bjx = 1; }  // `if (!isFinite(bestG)) bestI = bestJ = 1`
    const bool has_struct = (fv.rowmask[bi] >> (bjx - 1)) & 1u;
    msspe_thal_out res;
    res.ds = 0; res.dh = 0; res.dg = 0; res.tm = 0; res.no_structure = 1; res.n_bp = 0;
    if (has_struct) {
      const int ri = (n1[bi] * 5 + n1[bi + 1]) * 5 + n2[bjx + 1];
      const double2 bc = cell[(bi - 1) * k + (bjx - 1)];
      const double dH = bc.y + rsh[ri].y + kDHi;
      const double dS = bc.x + rsh[ri].x + kDSi;
      // ---------------- traceback: count paired positions ----------------
      int i = bi, j = bjx, pairs = 1;
      if (A.pairing && gl == 0) A.pairing[p * MSSPE_MAX_OLIGO + (i - 1)] = (uint8_t)j;
      for (int guard = 0; guard < ((A.dbg & 2) ? 0 : 2 * k + 2); guard++) {
        const int li = (n1[i] * 5 + n1[i - 1]) * 5 + n2[j - 1];
        const double2 c = cell[(i - 1) * k + (j - 1)];
        if (eq2(c.x, lsh[li].x) && eq2(c.y, lsh[li].y)) break;
        if (i > 1 && j > 1 && ((fv.rowmask[i - 1] >> (j - 2)) & 1u)) {
          const double2 st = tab[FT_STACK + i4(n1[i - 1], n1[i], n2[j - 1], n2[j])];
          const double2 pc = cell[(i - 2) * k + (j - 2)];
          if (eq2(c.x, st.x + pc.x) && eq2(c.y, st.y + pc.y)) {
            i--; j--; pairs++;
            if (A.pairing && gl == 0) A.pairing[p * MSSPE_MAX_OLIGO + (i - 1)] = (uint8_t)j;
            continue;
          }
        }
        int key = 0x7fffffff;
        const uint32_t xcl = (i > 1 && j > 1) ? (uint32_t)i4(n2[j], n2[j - 1] & 3, n1[i], n1[i - 1] & 3) : 0u;   // only used when i, j >= 2
        for (int l1 = gl; l1 <= i - 2; l1 += GROUP) {
          const int ii = i - 1 - l1;
          uint32_t cand = j >= 2 ? (fv.rowmask[ii] & ((1u << (j - 1)) - 1u)) : 0u;
          if (l1 == 0 && j >= 2) cand &= ~(1u << (j - 2));
          while (cand) {
            const int jj = __ffs(cand);
            cand &= cand - 1u;
            const int l2 = j - jj - 1;
            if (l1 + l2 > maxLoop) continue;
            double S, H;
            flat_candidate(tab, cell[(ii - 1) * k + (jj - 1)], (uint32_t)(fv.actx[ii] | fv.bctx[jj]), xcl, l1, l2, &S, &H);
            if (eq2(c.x, S) && eq2(c.y, H)) key = min(key, (l1 + l2) * 64 + l1);
          }
        }
        key = (int)__reduce_min_sync(gmask, (uint32_t)key);
        if (key == 0x7fffffff) break;
        const int l1 = key & 63, l2 = (key >> 6) - l1;
        i = i - 1 - l1; j = j - 1 - l2; pairs++;
        if (A.pairing && gl == 0) A.pairing[p * MSSPE_MAX_OLIGO + (i - 1)] = (uint8_t)j;
      }
      const int N = pairs - 1;  // (#paired bases in both strands)/2 - 1
      const double t = (dH / (dS + (N * saltCorr) + RC)) - kAbsZero;
      res.dg = dH - (t_user * (dS + (N * saltCorr)));
      res.ds = dS + (N * saltCorr);
      res.dh = dH;
      res.tm = t;
      res.no_structure = 0;
      res.n_bp = pairs;
    }
    if (gl == 0) {
      if (A.out) A.out[p] = res;
      if (A.matrix) {
        const unsigned long long pair = (unsigned long long)(A.row_begin + p / A.n) * A.n + col;
        if (res.no_structure) {
          const unsigned long long at = atomicAdd(A.n_nostruct, 1ull);
          if (at < A.nostruct_cap) A.nostruct[at] = pair;
        } else if (res.dg < A.dg_limit) {
          const unsigned long long at = atomicAdd(A.n_edges, 1ull);
          if (at < A.edge_cap) { A.edges[at].pair = pair; A.edges[at].dg = res.dg; }
        }
      }
    }
  }
}

// ---------------------------------------------------------------- device: dimer, one THREAD per ordered pair (oligos <= 16 nt)
// The same algorithm and results once more, mapped the other way round: the warp kernels above spread ONE pair over 32
// lanes and pay ~12,000 warp instructions per 13-mer pair, most of them bookkeeping that does not shrink with the work
// (ncu: profiles/r2s3_thal_flat*.txt).  A 13-mer pair has only ~42 paired cells and ~380 loop candidates, so here every
// lane walks its own pair: the DP matrix of a pair lives in that thread's local memory (k*k (S,H) cells, L1/L2-resident),
// the tables in shared memory (one (S,H) table, flat_candidate), the pairing masks are four registers.  Lanes diverge
// only in the trip counts of the bit loops; the scalar scan order of Primer3 is reproduced by the (dG, key) minimum as in
// the warp kernels.
constexpr int TK_MAX = 16;
struct ThreadShared : FlatShared {   // per block: the tables + what phase (A) of a row parks for the flat candidate loop
  double gcur[TK_MAX][DIMER_THREADS];     // dG of the q-th paired cell of the row as it stands after (A)
  uint16_t meta[TK_MAX][DIMER_THREADS];   // closing context i4(n2[j], n2[j-1], n1[i], n1[i-1]) | rsh index << 8
  uint16_t rmask[TK_MAX + 1][DIMER_THREADS];   // rowmask of this thread's pair, rows 1 .. k (k <= 16 columns)
};
__global__ void __launch_bounds__(DIMER_THREADS)
thal_dimer_thread_kernel(const DimerArgs A) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  ThreadShared& sh = *reinterpret_cast<ThreadShared*>(dyn_smem);
  const int k = A.k;
  const int tid = threadIdx.x;
  {  // stage tables
    const ThalDeviceTables* T = A.T;
    for (int x = tid; x < 256; x += DIMER_THREADS) {
      const int a = x >> 6, b = (x >> 4) & 3, c = (x >> 2) & 3, d = x & 3;
      const int g = THAL_IDX4(a, b, c, d);
      sh.tab[FT_STACK + x] = make_double2(T->stackS[g], T->stackH[g]);
      sh.tab[FT_INT2 + x] = make_double2(T->stackint2S[g], T->stackint2H[g]);
      sh.tab[FT_TST + x] = make_double2(T->tstackS[g], T->tstackH[g]);
    }
    for (int x = tid; x < 200; x += DIMER_THREADS) {
      (&sh.lsh[0][0])[x] = make_double2((&A.C->lshS[0][0])[x], (&A.C->lshH[0][0])[x]);
      (&sh.rsh[0][0])[x] = make_double2((&A.C->rshS[0][0])[x], (&A.C->rshH[0][0])[x]);
    }
    for (int x = tid; x < 30; x += DIMER_THREADS) {
      sh.tab[FT_INTERIOR + x] = make_double2(T->interiorS[x], T->interiorH[x]);
      sh.tab[FT_BULGE + x] = make_double2(T->bulgeS[x], T->bulgeH[x]);
    }
    for (int x = tid; x < 16; x += DIMER_THREADS) sh.tab[FT_ATP + x] = make_double2(T->atpS[(x >> 2) * 5 + (x & 3)], T->atpH[(x >> 2) * 5 + (x & 3)]);
    if (tid == 0) sh.tab[FT_ZERO] = make_double2(0.0, 0.0);
  }
  __syncthreads();
  const double2* __restrict__ tab = sh.tab;
  const int maxLoop = A.C->maxLoop;
  const double saltCorr = A.C->saltCorr, t_user = A.C->t_user_K;
  // (S,H) of the PAIRED cells only, row-major in the order the fill visits them, in this thread's own contiguous piece of
  // a global scratch buffer: ~42 cells = 670 B per 13-mer pair, so the pieces of all resident threads stay in L2 (thread-
  // local arrays are interleaved by 4-byte words: one divergent 16-byte access touched 4 sectors and the k*k array of every
  // resident thread was 4x the L2 -- 53 KB of DRAM traffic per pair, measured).  slot(i,j) = cells of rows < i + paired
  // columns of row i below j.
  double2* const cell = A.scratch + ((size_t)blockIdx.x * DIMER_THREADS + tid) * (size_t)(k * k);

  for (unsigned long long p = (unsigned long long)blockIdx.x * DIMER_THREADS + tid; p < A.n_pairs; p += (unsigned long long)gridDim.x * DIMER_THREADS) {
    uint64_t ca, cb;
    uint32_t col = 0;
    if (A.matrix) { col = A.colperm ? A.colperm[p % A.n] : (uint32_t)(p % A.n); ca = A.a[A.row_begin + p / A.n]; cb = A.b[col]; }
    else { ca = A.a[p]; cb = A.b[p]; }
    // numSeq1 = oligo1 5'->3': n1(i) = (ca >> 2(k-i)) & 3; numSeq2 = oligo2 REVERSED: n2(j) = (cb >> 2(j-1)) & 3; N (4) outside 1..k
    auto n1 = [&](int i) -> uint32_t { return (i < 1 || i > k) ? 4u : (uint32_t)(ca >> (2 * (k - i))) & 3u; };
    auto n2 = [&](int j) -> uint32_t { return (j < 1 || j > k) ? 4u : (uint32_t)(cb >> (2 * (j - 1))) & 3u; };
    uint32_t cm0 = 0u, cm1 = 0u, cm2 = 0u, cm3 = 0u;   // columns of A, C, G, T in the reversed second oligo
    for (int j = 1; j <= k; j++) {
      const uint32_t bit = 1u << (j - 1), y = (uint32_t)(cb >> (2 * (j - 1))) & 3u;
      cm0 |= y == 0u ? bit : 0u; cm1 |= y == 1u ? bit : 0u; cm2 |= y == 2u ? bit : 0u; cm3 |= y == 3u ? bit : 0u;
    }
    for (int i = 1; i <= k; i++) {                     // columns j that pair with row i, parked in this thread's column of smem
      const uint32_t x = (uint32_t)(ca >> (2 * (k - i))) & 3u;
      sh.rmask[i][tid] = (uint16_t)(x == 0u ? cm3 : x == 1u ? cm2 : x == 2u ? cm1 : cm0);
    }
    auto rowmask = [&](int i) -> uint32_t { return sh.rmask[i][tid]; };
    // context of an inner pair (ii,jj), ii, jj <= k - 1: i4(n1[ii], n1[ii+1], n2[jj], n2[jj+1])
    auto ctx_in = [&](int ii, int jj) -> uint32_t {
      const uint32_t av = (uint32_t)(ca >> (2 * (k - ii - 1))) & 15u;          // n1[ii] << 2 | n1[ii+1]
      const uint32_t bv = (uint32_t)(cb >> (2 * (jj - 1))) & 15u;              // n2[jj] | n2[jj+1] << 2
      return (av << 4) | ((bv & 3u) << 2) | (bv >> 2);
    };
    auto slot_of = [&](int i, int j) -> int {         // general form (tail); the fill keeps the row bases incrementally
      int base = 0;
      for (int r = 1; r < i; r++) base += __popc(rowmask(r));
      return base + __popc(rowmask(i) & ((1u << (j - 1)) - 1u));
    };
    const int sym = ((k & 1) == 0 && revcomp_code(ca, k) == ca && revcomp_code(cb, k) == cb) ? 1 : 0;
    const double RC = A.C->RC[sym];
    const double2* __restrict__ lsh = sh.lsh[sym];
    const double2* __restrict__ rsh = sh.rsh[sym];
    const double* __restrict__ t0tab = A.C->t0[sym];

    // ---------------- fill ----------------
    // Per row: (A) the end / stack terms of the row's paired cells; (B) ONE flat loop over all bulge / internal-loop
    // candidates of all cells of the row: a lane whose inner row or whose cell is exhausted moves on inside the same
    // iteration, so the lanes of a warp (32 different pairs) diverge by their candidate COUNT per row only -- with the plain
    // cell / inner-row / column nest 7 of 32 lanes were active (ncu).  What the loop needs per cell (dG of the cell as it
    // stands, closing context, end-table index) is parked in shared memory by (A).
    int rb_i = 0, rb_prev = 0;                        // cells in rows < i, cells in rows < i - 1
    for (int i = 1; i <= k; i++) {
      const uint32_t a = n1(i), am = n1(i - 1), ap = n1(i + 1);
      const uint32_t rm_i = rowmask(i), rm_prev = i > 1 ? rowmask(i - 1) : 0u;
      int q = 0;
      for (uint32_t bj = rm_i; bj; bj &= bj - 1u, q++) {
        const int j = __ffs(bj);
        const uint32_t b = n2(j), bm = n2(j - 1), bp = n2(j + 1);
        const int li = (a * 5 + am) * 5 + bm;
        const double2 l = lsh[li];
        double S = l.x, H = l.y;
        if (i > 1 && j > 1) {
          const int rn = ap * 5 + bp;
          const double2 r = rsh[a * 25 + rn];
          const double rS = r.x, rH = r.y;
          double S0 = S, H0 = H, S1, H1, T1;
          const double T0 = __ldg(&t0tab[li * 25 + rn]);   // (H0 + kDHi + rH) / (S0 + kDSi + rS + RC), tabulated on the host
          const double2 st = tab[FT_STACK + i4(am, a, bm, b)];
          const bool prev_bp = (rm_prev >> (j - 2)) & 1u;
          if (prev_bp && isfinite(st.y)) {
            const double2 pc = cell[rb_prev + __popc(rm_prev & ((1u << (j - 2)) - 1u))];
            S1 = pc.x + st.x;
            H1 = pc.y + st.y;
            T1 = (H1 + kDHi + rH) / (S1 + kDSi + rS + RC);
          } else {
            S1 = -1.0; H1 = INFINITY;
            T1 = (H1 + kDHi) / (S1 + kDSi + RC);
          }
          if (S1 < kMinEntropyCutoff) { S1 = kMinEntropy; H1 = 0.0; }
          if (S0 < kMinEntropyCutoff) { S0 = kMinEntropy; H0 = 0.0; }
          if (T1 > T0) { S = S1; H = H1; } else if (T0 >= T1) { S = S0; H = H0; }
          sh.gcur[q][tid] = H + rH - kTK * (S + rS);
          sh.meta[q][tid] = (uint16_t)(i4(b, bm & 3, a, am & 3) | ((a * 25 + rn) << 8));
        }
        cell[rb_i + q] = make_double2(S, H);
      }
      if (i > 1 && (rm_i & ~1u) && !(A.dbg & 1)) {
        uint32_t bjr = rm_i & ~1u, clampbits = 0u;
        int qq = (int)(rm_i & 1u) - 1;                 // index of the current cell among the row's paired cells
        int j = 0, ii = 1, rb_ii = 0;
        uint32_t cand = 0u, rm_ii = 0u, below = 0u, xcl = 0u;
        double Gc = 0.0, rS = 0.0, rH = 0.0, bG = INFINITY, bS = -1.0, bH = INFINITY;
        int bkey = 0x7fffffff;
        for (;;) {
          while (cand == 0u) {
            if (j > 0 && ii > 1) { ii--; rm_ii = rowmask(ii); rb_ii -= __popc(rm_ii); cand = rm_ii & below; continue; }
            if (j > 0 && bG < Gc && !((clampbits >> (j - 1)) & 1u)) cell[rb_i + qq] = make_double2(bS, bH);
            if (bjr == 0u) { j = -1; break; }
            j = __ffs(bjr); bjr &= bjr - 1u; qq++;
            Gc = sh.gcur[qq][tid];
            const uint32_t m = sh.meta[qq][tid];
            xcl = m & 0xffu;
            const double2 r = rsh[m >> 8];
            rS = r.x; rH = r.y;
            bG = INFINITY; bS = -1.0; bH = INFINITY; bkey = 0x7fffffff;
            below = (1u << (j - 1)) - 1u;
            ii = i - 1; rm_ii = rm_prev; rb_ii = rb_prev;
            cand = rm_prev & below & ~(1u << (j - 2));
          }
          if (j < 0) break;
          const int jj = __ffs(cand);
          cand &= cand - 1u;
          const int l1 = i - ii - 1, l2 = j - jj - 1;
          if (l1 + l2 <= maxLoop) {
            double cS, cH;
            flat_candidate(tab, cell[rb_ii + __popc(rm_ii & ((1u << (jj - 1)) - 1u))], ctx_in(ii, jj), xcl, l1, l2, &cS, &cH);
            if (isfinite(cH)) {
              if (cS < kMinEntropyCutoff) clampbits |= 1u << (j - 1);
              const double G1 = cH + rH - kTK * (cS + rS);
              const int key = (l1 + l2) * 64 + l1;
              const bool better = (G1 < bG) | ((G1 == bG) & (key < bkey));
              bG = better ? G1 : bG; bkey = better ? key : bkey; bS = better ? cS : bS; bH = better ? cH : bH;
            }
          }
        }
        // An entropy below the cutoff re-bases the cell mid-scan: replay those cells' scans in the scalar order.
        for (; clampbits; clampbits &= clampbits - 1u) {
          const int jc = __ffs(clampbits);
          const int sl = rb_i + __popc(rm_i & ((1u << (jc - 1)) - 1u));
          const uint32_t bc = n2(jc), bmc = n2(jc - 1);
          const uint32_t xc = (uint32_t)i4(bc, bmc & 3, a, am & 3);
          const double2 r = rsh[a * 25 + ap * 5 + n2(jc + 1)];
          double cS = cell[sl].x, cH = cell[sl].y;
          for (int d = 3; d <= maxLoop + 2; d++) {
            int i2 = i - 1, j2 = -i2 - d + (jc + i);
            if (j2 < 1) { i2 -= (1 - j2); j2 = 1; }
            for (; i2 > 0 && j2 < jc; --i2, ++j2) {
              if (!((rowmask(i2) >> (j2 - 1)) & 1u)) continue;
              double xS, xH;
              flat_candidate(tab, cell[slot_of(i2, j2)], ctx_in(i2, j2), xc, i - i2 - 1, jc - j2 - 1, &xS, &xH);
              const double G1 = xH + r.y - kTK * (xS + r.x), G2 = cH + r.y - kTK * (cS + r.x);
              if (!(G1 < G2)) { xS = -1.0; xH = INFINITY; }
              if (xS < kMinEntropyCutoff) { xS = kMinEntropy; xH = 0.0; }
              if (isfinite(xH)) { cS = xS; cH = xH; }
            }
          }
          cell[sl] = make_double2(cS, cH);
        }
      }
      rb_prev = rb_i; rb_i += __popc(rm_i);
    }

    // ---------------- best terminal pair ----------------
    double bG = INFINITY; int bkey = 0x7fffffff;
    const int i_first = A.type == MSSPE_THAL_ANY ? 1 : k;
    int sl = slot_of(i_first, 1);                     // cells are stored in exactly this order
    for (int i = i_first; i <= k; i++) {
      const uint32_t a = n1(i), ap = n1(i + 1);
      for (uint32_t bj = rowmask(i); bj; bj &= bj - 1u) {
        const int j = __ffs(bj);
        const double2 r = rsh[(a * 5 + ap) * 5 + n2(j + 1)];
        const double rS = r.x + kSmallNonZero, rH = r.y + kSmallNonZero;
        const double2 c = cell[sl++];
        const double G1 = (c.y + rH + kDHi) - kTK * (c.x + rS + kDSi);
        const int key = i * 64 + j;
        if (G1 < bG || (G1 == bG && key < bkey)) { bG = G1; bkey = key; }
      }
    }
    const bool none = !isfinite(bG);  // no base pair anywhere (for END1: none in the last row)
    int bi = none ? (A.type == MSSPE_THAL_ANY ? 1 : k) : (bkey >> 6);
    int bjx = none ? 1 : (bkey & 63);
    if (none && A.type != MSSPE_THAL_ANY) { bi = 1; bjx = 1; }  // `if (!isFinite(bestG)) bestI = bestJ = 1`
    const bool has_struct = (rowmask(bi) >> (bjx - 1)) & 1u;
    msspe_thal_out res;
    res.ds = 0; res.dh = 0; res.dg = 0; res.tm = 0; res.no_structure = 1; res.n_bp = 0;
    if (has_struct) {
      const double2 rb = rsh[(n1(bi) * 5 + n1(bi + 1)) * 5 + n2(bjx + 1)];
      const double2 bc = cell[slot_of(bi, bjx)];
      const double dH = bc.y + rb.y + kDHi;
      const double dS = bc.x + rb.x + kDSi;
      // ---------------- traceback: count paired positions ----------------
      int i = bi, j = bjx, pairs = 1;
      if (A.pairing) A.pairing[p * MSSPE_MAX_OLIGO + (i - 1)] = (uint8_t)j;
      for (int guard = 0; guard < ((A.dbg & 2) ? 0 : 2 * k + 2); guard++) {
        const double2 l = lsh[(n1(i) * 5 + n1(i - 1)) * 5 + n2(j - 1)];
        const double2 c = cell[slot_of(i, j)];
        if (eq2(c.x, l.x) && eq2(c.y, l.y)) break;
        if (i > 1 && j > 1 && ((rowmask(i - 1) >> (j - 2)) & 1u)) {
          const double2 st = tab[FT_STACK + i4(n1(i - 1), n1(i), n2(j - 1), n2(j))];
          const double2 pc = cell[slot_of(i - 1, j - 1)];
          if (eq2(c.x, st.x + pc.x) && eq2(c.y, st.y + pc.y)) {
            i--; j--; pairs++;
            if (A.pairing) A.pairing[p * MSSPE_MAX_OLIGO + (i - 1)] = (uint8_t)j;
            continue;
          }
        }
        int key = 0x7fffffff;
        if (i > 1 && j > 1) {
          const uint32_t xcl = (uint32_t)i4(n2(j), n2(j - 1) & 3, n1(i), n1(i - 1) & 3);
          for (int ii = i - 1; ii >= 1; ii--) {
            uint32_t cand = rowmask(ii) & ((1u << (j - 1)) - 1u);
            if (ii == i - 1) cand &= ~(1u << (j - 2));
            const int l1 = i - ii - 1;
            while (cand) {
              const int jj = __ffs(cand);
              cand &= cand - 1u;
              const int l2 = j - jj - 1;
              if (l1 + l2 > maxLoop) continue;
              double cS, cH;
              flat_candidate(tab, cell[slot_of(ii, jj)], ctx_in(ii, jj), xcl, l1, l2, &cS, &cH);
              if (eq2(c.x, cS) && eq2(c.y, cH)) key = min(key, (l1 + l2) * 64 + l1);
            }
          }
        }
        if (key == 0x7fffffff) break;
        const int l1 = key & 63, l2 = (key >> 6) - l1;
        i = i - 1 - l1; j = j - 1 - l2; pairs++;
        if (A.pairing) A.pairing[p * MSSPE_MAX_OLIGO + (i - 1)] = (uint8_t)j;
      }
      const int N = pairs - 1;  // (#paired bases in both strands)/2 - 1
      const double t = (dH / (dS + (N * saltCorr) + RC)) - kAbsZero;
      res.dg = dH - (t_user * (dS + (N * saltCorr)));
      res.ds = dS + (N * saltCorr);
      res.dh = dH;
      res.tm = t;
      res.no_structure = 0;
      res.n_bp = pairs;
    }
    if (A.out) A.out[p] = res;
    if (A.matrix) {
      const unsigned long long pair = (unsigned long long)(A.row_begin + p / A.n) * A.n + col;
      if (res.no_structure) {
        const unsigned long long at = atomicAdd(A.n_nostruct, 1ull);
        if (at < A.nostruct_cap) A.nostruct[at] = pair;
      } else if (res.dg < A.dg_limit) {
        const unsigned long long at = atomicAdd(A.n_edges, 1ull);
        if (at < A.edge_cap) { A.edges[at].pair = pair; A.edges[at].dg = res.dg; }
      }
    }
  }
}

// ---------------------------------------------------------------- device: monomer (hairpin), one thread per oligo
constexpr int MONO_MAX = MSSPE_MAX_OLIGO;
constexpr int MIN_HRPN_LOOP = 3;

// M = largest oligo the scratch holds: 16 for the 13-16-mers of the pipeline (5.6 KB per thread: 2000 primers fit one wave
// of shared-memory blocks), 32 for the general entry points
template <int M>
struct MonoWorkT {
  int n1[M + 2];
  int len;
  double Sm[M + 2][M + 2], Hm[M + 2][M + 2];
  double send5[M + 2], hend5[M + 2];
  int maxLoop;
};
using MonoWork = MonoWorkT<MONO_MAX>;
// dplx_init_H = 0, dplx_init_S = -1e-11, RC = 0 for unimolecular folding
#define M_DHI 0.0
#define M_DSI (-0.00000000001)
#define M_RC 0.0

__device__ __forceinline__ bool m_bp(int a, int b) { return a + b == 3 && a < 4 && b < 4; }

template <class WK>
__device__ double m_Ss(const ThalDeviceTables* T, const WK& w, int i, int j) {
  if (i >= j) return -1.0;
  if (i == w.len || j == w.len + 1) return -1.0;
  return T->stackS[THAL_IDX4(w.n1[i], w.n1[i + 1], w.n1[j], w.n1[j - 1])];
}
template <class WK>
__device__ double m_Hs(const ThalDeviceTables* T, const WK& w, int i, int j) {
  if (i >= j) return INFINITY;
  if (i == w.len || j == w.len + 1) return INFINITY;
  const double h = T->stackH[THAL_IDX4(w.n1[i], w.n1[i + 1], w.n1[j], w.n1[j - 1])];
  return isfinite(h) ? h : (double)INFINITY;
}

__device__ bool m_find(const uint32_t* keys, const double* vals, int n, uint32_t key, double* v) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] == key) { *v = vals[mid]; return true; }
    if (keys[mid] < key) lo = mid + 1; else hi = mid - 1;
  }
  return false;
}

// RSH(i, j) of thal.c in the unimolecular setting (second strand = the oligo itself, unreversed; dplx_init_H = 0,
// dplx_init_S = -1e-11, RC = 0): the right-end term that Primer3 2.6.1 adds to both sides of the hairpin-closing comparison.
template <class WK>
__device__ void m_rsh(const ThalDeviceTables* T, const WK& w, int i, int j, double* S, double* H) {
  const int a = w.n1[i], b = w.n1[j], a1 = w.n1[i + 1], b1 = w.n1[j + 1];
  if (!m_bp(a, b)) { *S = -1.0; *H = INFINITY; return; }
  const double aS = T->atpS[a * 5 + b], aH = T->atpH[a * 5 + b];
  double S1 = aS + T->tstack2S[THAL_IDX4(a, a1, b, b1)];
  double H1 = aH + T->tstack2H[THAL_IDX4(a, a1, b, b1)];
  bool has = false;
  double S2 = 0.0, H2 = 0.0;
  if (!m_bp(a1, b1)) {
    const double h3 = T->dangle3H[THAL_IDX3(a, a1, b)], h5 = T->dangle5H[THAL_IDX3(a, b, b1)];
    if (isfinite(h3) && isfinite(h5)) {
      S2 = aS + T->dangle3S[THAL_IDX3(a, a1, b)] + T->dangle5S[THAL_IDX3(a, b, b1)]; H2 = aH + h3 + h5; has = true;
    } else if (isfinite(h3)) { S2 = aS + T->dangle3S[THAL_IDX3(a, a1, b)]; H2 = aH + h3; has = true; }
    else if (isfinite(h5)) { S2 = aS + T->dangle5S[THAL_IDX3(a, b, b1)]; H2 = aH + h5; has = true; }
  }
  double G1 = H1 - kTK * S1, T1 = -INFINITY, G2, T2;
  if (!isfinite(H1) || G1 > 0) { H1 = INFINITY; S1 = -1.0; G1 = 1.0; }
  if (has) {
    G2 = H2 - kTK * S2;
    if (!isfinite(H2) || G2 > 0) { H2 = INFINITY; S2 = -1.0; G2 = 1.0; }
    T2 = (H2 + M_DHI) / (S2 + M_DSI + M_RC);
    if (isfinite(H1) && G1 < 0) {
      T1 = (H1 + M_DHI) / (S1 + M_DSI + M_RC);
      if (T1 < T2 && G2 < 0) { S1 = S2; H1 = H2; T1 = T2; }
    } else if (G2 < 0) { S1 = S2; H1 = H2; T1 = T2; }
  }
  T2 = (aH + M_DHI) / (aS + M_DSI + M_RC);
  if (isfinite(H1) && !(T1 < T2)) { *S = S1; *H = H1; } else { *S = aS; *H = aH; }
}

template <class WK>
__device__ void m_hairpin_loop(const ThalDeviceTables* T, const WK& w, int i, int j, double* S, double* H, int tb) {
  const int* n1 = w.n1;
  const int loopSize = j - i - 1;
  if (loopSize < MIN_HRPN_LOOP) { *S = -1.0; *H = INFINITY; return; }
  if (i <= w.len && w.len < j) { *S = -1.0; *H = INFINITY; return; }
  if (loopSize <= 30) { *H = T->hairpinH[loopSize - 1]; *S = T->hairpinS[loopSize - 1]; }
  else { *H = T->hairpinH[29]; *S = T->hairpinS[29]; }
  if (loopSize > 3) {
    *H += T->tstack2H[THAL_IDX4(n1[i], n1[i + 1], n1[j], n1[j - 1])];
    *S += T->tstack2S[THAL_IDX4(n1[i], n1[i + 1], n1[j], n1[j - 1])];
  } else if (loopSize == 3) {
    *H += T->atpH[n1[i] * 5 + n1[j]];
    *S += T->atpS[n1[i] * 5 + n1[j]];
  }
  if (loopSize == 3) {
    uint32_t key = 0;
    for (int c = 0; c < 5; c++) key = key * 5u + (uint32_t)n1[i + c];
    double v;
    if (T->nTriH && m_find(T->triKeyH, T->triH, T->nTriH, key, &v)) *H += v;
    if (T->nTriS && m_find(T->triKeyS, T->triS, T->nTriS, key, &v)) *S += v;
  } else if (loopSize == 4) {
    uint32_t key = 0;
    for (int c = 0; c < 6; c++) key = key * 5u + (uint32_t)n1[i + c];
    double v;
    if (T->nTetraH && m_find(T->tetraKeyH, T->tetraH, T->nTetraH, key, &v)) *H += v;
    if (T->nTetraS && m_find(T->tetraKeyS, T->tetraS, T->nTetraS, key, &v)) *S += v;
  }
  if (!isfinite(*H)) { *H = INFINITY; *S = -1.0; }
  if (*H > 0 && *S > 0 && (!(w.Hm[i][j] > 0) || !(w.Sm[i][j] > 0))) { *H = INFINITY; *S = -1.0; }
  // Primer3 2.6.1 compares free energies at 37 C with the right-end term on both sides (calc_hairpin calls RSH in the
  // reference's ntthal executable; tests/golden/ntthal_emulated.json), not melting temperatures as older releases did
  if (tb == 0) {
    double rs, rh;
    m_rsh(T, w, i, j, &rs, &rh);
    const double G1 = *H + rh - kTK * (*S + rs);
    const double G2 = w.Hm[i][j] + rh - kTK * (w.Sm[i][j] + rs);
    if (G2 < G1) { *S = w.Sm[i][j]; *H = w.Hm[i][j]; }
  }
}

template <class WK>
__device__ void m_bulge_internal(const ThalDeviceTables* T, const WK& w, int i, int j, int ii, int jj, double* outS,
                                 double* outH, int tb) {
  const int* n1 = w.n1;
  const int l1 = ii - i - 1, l2 = j - jj - 1;
  double S, H, T1, T2;
  if (l1 + l2 > w.maxLoop) { *outS = -1.0; *outH = INFINITY; return; }
  const int ls = l1 + l2 - 1;
  if ((l1 == 0 && l2 > 0) || (l2 == 0 && l1 > 0)) {
    if (l2 == 1 || l1 == 1) {
      H = T->bulgeH[ls] + T->stackH[THAL_IDX4(n1[i], n1[ii], n1[j], n1[jj])];
      S = T->bulgeS[ls] + T->stackS[THAL_IDX4(n1[i], n1[ii], n1[j], n1[jj])];
    } else {
      H = T->bulgeH[ls] + T->atpH[n1[i] * 5 + n1[j]] + T->atpH[n1[ii] * 5 + n1[jj]];
      S = T->bulgeS[ls] + T->atpS[n1[i] * 5 + n1[j]] + T->atpS[n1[ii] * 5 + n1[jj]];
    }
    if (tb != 1) { H += w.Hm[ii][jj]; S += w.Sm[ii][jj]; }
    if (!isfinite(H)) { H = INFINITY; S = -1.0; }
    T1 = (H + M_DHI) / ((S + M_DSI) + M_RC);
    T2 = (w.Hm[i][j] + M_DHI) / ((w.Sm[i][j]) + M_DSI + M_RC);
    if ((T1 > T2) || ((tb && T1 >= T2) || tb == 1)) { *outS = S; *outH = H; }
  } else if (l1 == 1 && l2 == 1) {
    S = T->stackint2S[THAL_IDX4(n1[i], n1[i + 1], n1[j], n1[j - 1])] + T->stackint2S[THAL_IDX4(n1[jj], n1[jj + 1], n1[ii], n1[ii - 1])];
    if (tb != 1) S += w.Sm[ii][jj];
    H = T->stackint2H[THAL_IDX4(n1[i], n1[i + 1], n1[j], n1[j - 1])] + T->stackint2H[THAL_IDX4(n1[jj], n1[jj + 1], n1[ii], n1[ii - 1])];
    if (tb != 1) H += w.Hm[ii][jj];
    if (!isfinite(H)) { H = INFINITY; S = -1.0; }
    T1 = (H + M_DHI) / ((S + M_DSI) + M_RC);
    T2 = (w.Hm[i][j] + M_DHI) / ((w.Sm[i][j]) + M_DSI + M_RC);
    if ((T1 - T2 >= 0.000001) || tb) {
      if ((T1 > T2) || ((tb && T1 >= T2) || tb == 1)) { *outS = S; *outH = H; }
    }
  } else {
    const int asym = l1 > l2 ? l1 - l2 : l2 - l1;
    H = T->interiorH[ls] + T->tstackH[THAL_IDX4(n1[i], n1[i + 1], n1[j], n1[j - 1])] +
        T->tstackH[THAL_IDX4(n1[jj], n1[jj + 1], n1[ii], n1[ii - 1])] + (K_ILAH * asym);
    if (tb != 1) H += w.Hm[ii][jj];
    S = T->interiorS[ls] + T->tstackS[THAL_IDX4(n1[i], n1[i + 1], n1[j], n1[j - 1])] +
        T->tstackS[THAL_IDX4(n1[jj], n1[jj + 1], n1[ii], n1[ii - 1])] + (K_ILAS * asym);
    if (tb != 1) S += w.Sm[ii][jj];
    if (!isfinite(H)) { H = INFINITY; S = -1.0; }
    T1 = (H + M_DHI) / ((S + M_DSI) + M_RC);
    T2 = (w.Hm[i][j] + M_DHI) / ((w.Sm[i][j]) + M_DSI + M_RC);
    if ((T1 > T2) || ((tb && T1 >= T2) || (tb == 1))) { *outS = S; *outH = H; }
  }
}

template <class WK>
__device__ void m_loops(const ThalDeviceTables* T, WK& w, int i, int j, double* S, double* H, int tb) {
  for (int d = j - i - 3; d >= MIN_HRPN_LOOP + 1 && d >= j - i - 2 - w.maxLoop; --d)
    for (int ii = i + 1; ii < j - d && ii <= w.len; ++ii) {
      const int jj = d + ii;
      if (tb == 0) { *S = -1.0; *H = INFINITY; }
      if (isfinite(w.Hm[ii][jj]) && isfinite(w.Hm[i][j])) {
        m_bulge_internal(T, w, i, j, ii, jj, S, H, tb);
        if (isfinite(*H)) {
          if (*S < kMinEntropyCutoff) { *S = kMinEntropy; *H = 0.0; }
          if (tb == 0) { w.Hm[i][j] = *H; w.Sm[i][j] = *S; }
        }
      }
    }
}

// exterior-fragment recursion: variant 1 = plain pair (k+1,i); 2 = 5' dangle; 3 = 3' dangle; 4 = terminal mismatch
template <class WK>
__device__ void m_end5_term(const ThalDeviceTables* T, const WK& w, int i, int k, int variant, double* eS, double* eH) {
  const int* n1 = w.n1;
  switch (variant) {
    case 1: *eH = T->atpH[n1[k + 1] * 5 + n1[i]] + w.Hm[k + 1][i]; *eS = T->atpS[n1[k + 1] * 5 + n1[i]] + w.Sm[k + 1][i]; break;
    case 2: *eH = T->atpH[n1[k + 2] * 5 + n1[i]] + T->dangle5H[THAL_IDX3(n1[i], n1[k + 2], n1[k + 1])] + w.Hm[k + 2][i];
            *eS = T->atpS[n1[k + 2] * 5 + n1[i]] + T->dangle5S[THAL_IDX3(n1[i], n1[k + 2], n1[k + 1])] + w.Sm[k + 2][i]; break;
    case 3: *eH = T->atpH[n1[k + 1] * 5 + n1[i - 1]] + T->dangle3H[THAL_IDX3(n1[i - 1], n1[i], n1[k + 1])] + w.Hm[k + 1][i - 1];
            *eS = T->atpS[n1[k + 1] * 5 + n1[i - 1]] + T->dangle3S[THAL_IDX3(n1[i - 1], n1[i], n1[k + 1])] + w.Sm[k + 1][i - 1]; break;
    default: *eH = T->atpH[n1[k + 2] * 5 + n1[i - 1]] + T->tstack2H[THAL_IDX4(n1[i - 1], n1[i], n1[k + 2], n1[k + 1])] + w.Hm[k + 2][i - 1];
             *eS = T->atpS[n1[k + 2] * 5 + n1[i - 1]] + T->tstack2S[THAL_IDX4(n1[i - 1], n1[i], n1[k + 2], n1[k + 1])] + w.Sm[k + 2][i - 1]; break;
  }
}
__device__ __forceinline__ int m_end5_kmax(int i, int variant) {
  return variant == 1 ? i - MIN_HRPN_LOOP - 2 : (variant == 4 ? i - MIN_HRPN_LOOP - 4 : i - MIN_HRPN_LOOP - 3);
}
template <class WK>
__device__ void m_end5(const ThalDeviceTables* T, const WK& w, int i, int variant, double* outH, double* outS) {
  double H_max = INFINITY, S_max = -1.0, max_tm = -INFINITY;
  const int kmax = m_end5_kmax(i, variant);
  for (int k = 0; k <= kmax; ++k) {
    double T1 = (w.hend5[k] + M_DHI) / (w.send5[k] + M_DSI + M_RC);
    const double T2 = (0 + M_DHI) / (0 + M_DSI + M_RC);
    double eH, eS, H, S;
    m_end5_term(T, w, i, k, variant, &eS, &eH);
    if (T1 >= T2) { H = w.hend5[k] + eH; S = w.send5[k] + eS; } else { H = 0 + eH; S = 0 + eS; }
    if (!isfinite(H) || H > 0 || S > 0) { H = INFINITY; S = -1.0; }
    T1 = (H + M_DHI) / (S + M_DSI + M_RC);
    if (max_tm < T1) {
      if (S > kMinEntropyCutoff) { H_max = H; S_max = S; max_tm = T1; }
    }
  }
  *outH = H_max; *outS = S_max;
}

constexpr int MONO_GW = 16;   // lanes per oligo in the group kernel (oligos <= 20 nt have at most 16 rows per column)
template <class WK>
__device__ void mono_run(const ThalDeviceTables* T, WK& w, double saltCorr, double temp_K, msspe_thal_out* out) {
  const int* n1 = w.n1;
  const int len = w.len;
  out->ds = 0; out->dh = 0; out->dg = 0; out->tm = 0; out->no_structure = 1; out->n_bp = 0;
  for (int i = 1; i <= len; ++i)
    for (int j = i; j <= len; ++j) {
      if (j - i < MIN_HRPN_LOOP + 1 || !m_bp(n1[i], n1[j])) { w.Hm[i][j] = INFINITY; w.Sm[i][j] = -1.0; }
      else { w.Hm[i][j] = 0.0; w.Sm[i][j] = kMinEntropy; }
    }
  for (int j = 2; j <= len; ++j)
    for (int i = j - MIN_HRPN_LOOP - 1; i >= 1; --i) {
      if (!isfinite(w.Hm[i][j])) continue;
      double S0 = w.Sm[i][j], H0 = w.Hm[i][j];
      const double T0 = (H0 + M_DHI) / (S0 + M_DSI + M_RC);
      double S1 = w.Sm[i + 1][j - 1] + m_Ss(T, w, i, j);
      double H1 = w.Hm[i + 1][j - 1] + m_Hs(T, w, i, j);
      const double T1 = (H1 + M_DHI) / (S1 + M_DSI + M_RC);
      if (S1 < kMinEntropyCutoff) { S1 = kMinEntropy; H1 = 0.0; }
      if (S0 < kMinEntropyCutoff) { S0 = kMinEntropy; H0 = 0.0; }
      if (T1 > T0) { w.Sm[i][j] = S1; w.Hm[i][j] = H1; } else { w.Sm[i][j] = S0; w.Hm[i][j] = H0; }
      double s = -1.0, h = INFINITY;
      m_loops(T, w, i, j, &s, &h, 0);
      s = -1.0; h = INFINITY;
      m_hairpin_loop(T, w, i, j, &s, &h, 0);
      if (isfinite(h)) {
        if (s < kMinEntropyCutoff) { s = kMinEntropy; h = 0.0; }
        w.Sm[i][j] = s; w.Hm[i][j] = h;
      }
    }
  // exterior fragments
  w.send5[0] = w.send5[1] = -1.0;
  w.hend5[0] = w.hend5[1] = INFINITY;
  for (int i = 2; i <= len; i++) { w.send5[i] = kMinEntropy; w.hend5[i] = 0; }
  for (int i = 2; i <= len; ++i) {
    double eh[5], es[5], Tm[6];
    Tm[1] = (w.hend5[i - 1] + M_DHI) / (w.send5[i - 1] + M_DSI + M_RC);
    for (int v = 1; v <= 4; v++) {
      m_end5(T, w, i, v, &eh[v], &es[v]);
      Tm[v + 1] = (eh[v] + M_DHI) / (es[v] + M_DSI + M_RC);
    }
    int mx;
    if (Tm[1] > Tm[2] && Tm[1] > Tm[3] && Tm[1] > Tm[4] && Tm[1] > Tm[5]) mx = 1;
    else if (Tm[2] > Tm[3] && Tm[2] > Tm[4] && Tm[2] > Tm[5]) mx = 2;
    else if (Tm[3] > Tm[4] && Tm[3] > Tm[5]) mx = 3;
    else if (Tm[4] > Tm[5]) mx = 4;
    else mx = 5;
    if (mx == 1) { w.send5[i] = w.send5[i - 1]; w.hend5[i] = w.hend5[i - 1]; }
    else {
      const int v = mx - 1;
      const double G = eh[v] - (temp_K * (es[v]));
      if (G < 0.0) { w.send5[i] = es[v]; w.hend5[i] = eh[v]; }
      else { w.send5[i] = w.send5[i - 1]; w.hend5[i] = w.hend5[i - 1]; }
    }
  }
  const double mh = w.hend5[len], ms = w.send5[len];
  if (!isfinite(mh) || !isfinite(ms)) return;
  // traceback with an explicit stack
  int bpv[MONO_MAX + 2];
  for (int t = 0; t < len; ++t) bpv[t] = 0;
  struct { short i, j, m; } stk[4 * MONO_MAX];
  int sp = 0;
#define M_PUSH(a_, b_, c_) do { if (sp < 4 * MONO_MAX) { stk[sp].i = (short)(a_); stk[sp].j = (short)(b_); stk[sp].m = (short)(c_); sp++; } } while (0)
  M_PUSH(len, 0, 1);
  while (sp > 0) {
    sp--;
    int i = stk[sp].i;
    const int j = stk[sp].j, m = stk[sp].m;
    if (m == 1) {
      while (i >= 1 && eq2(w.send5[i], w.send5[i - 1]) && eq2(w.hend5[i], w.hend5[i - 1])) --i;
      if (i == 0) continue;
      bool handled = false;
      for (int v = 1; v <= 4 && !handled; v++) {
        double eh, es;
        m_end5(T, w, i, v, &eh, &es);
        if (!(eq2(w.send5[i], es) && eq2(w.hend5[i], eh))) continue;
        handled = true;
        const int kmax = m_end5_kmax(i, v);
        for (int k = 0; k <= kmax; ++k) {
          double xS, xH;
          m_end5_term(T, w, i, k, v, &xS, &xH);
          const int pi = (v == 1 || v == 3) ? k + 1 : k + 2;
          const int pj = (v <= 2) ? i : i - 1;
          if (eq2(w.send5[i], xS) && eq2(w.hend5[i], xH)) { M_PUSH(pi, pj, 0); break; }
          else if (eq2(w.send5[i], w.send5[k] + xS) && eq2(w.hend5[i], w.hend5[k] + xH)) { M_PUSH(pi, pj, 0); M_PUSH(k, 0, 1); break; }
        }
      }
    } else {
      bpv[i - 1] = j; bpv[j - 1] = i;
      double s1 = -1.0, h1 = INFINITY, s2 = -1.0, h2 = INFINITY;
      m_hairpin_loop(T, w, i, j, &s1, &h1, 1);
      m_loops(T, w, i, j, &s2, &h2, 2);
      if (eq2(w.Sm[i][j], m_Ss(T, w, i, j) + w.Sm[i + 1][j - 1]) && eq2(w.Hm[i][j], m_Hs(T, w, i, j) + w.Hm[i + 1][j - 1])) {
        M_PUSH(i + 1, j - 1, 0);
      } else if (eq2(w.Sm[i][j], s1) && eq2(w.Hm[i][j], h1)) {
        // hairpin loop closes here
      } else if (eq2(w.Sm[i][j], s2) && eq2(w.Hm[i][j], h2)) {
        bool done = false;
        for (int d = j - i - 3; d >= MIN_HRPN_LOOP + 1 && d >= j - i - 2 - w.maxLoop && !done; --d)
          for (int ii = i + 1; ii < j - d; ++ii) {
            const int jj = d + ii;
            double es = -1.0, eh = INFINITY;
            m_bulge_internal(T, w, i, j, ii, jj, &es, &eh, 1);
            if (eq2(w.Sm[i][j], es + w.Sm[ii][jj]) && eq2(w.Hm[i][j], eh + w.Hm[ii][jj])) { M_PUSH(ii, jj, 0); done = true; break; }
          }
      }
    }
  }
#undef M_PUSH
  int N = 0;
  for (int i = 1; i < len; ++i) if (bpv[i - 1] > 0) N++;
  out->n_bp = N / 2;
  const double t = (mh / (ms + (((N / 2) - 1) * saltCorr))) - kAbsZero;
  out->dg = mh - (temp_K * (ms + (((N / 2) - 1) * saltCorr)));
  out->ds = ms + (((N / 2) - 1) * saltCorr);
  out->dh = mh;
  out->tm = t;
  out->no_structure = 0;
}

template <class WK>
__device__ void mono_run_group(const ThalDeviceTables* T, WK& w, double saltCorr, double temp_K, msspe_thal_out* out, const int gl, const unsigned gmask) {
  const int* n1 = w.n1;
  const int len = w.len;
  // MONO_GW lanes share one oligo: the cells of a column (all i for one j) depend on earlier columns only, so lane gl takes
  // row i = j - 4 - gl of every column; the exterior recursion gives its four variants to four lanes; the traceback is lane 0's
  out->ds = 0; out->dh = 0; out->dg = 0; out->tm = 0; out->no_structure = 1; out->n_bp = 0;
  for (int i = 1 + gl; i <= len; i += MONO_GW)
    for (int j = i; j <= len; ++j) {
      if (j - i < MIN_HRPN_LOOP + 1 || !m_bp(n1[i], n1[j])) { w.Hm[i][j] = INFINITY; w.Sm[i][j] = -1.0; }
      else { w.Hm[i][j] = 0.0; w.Sm[i][j] = kMinEntropy; }
    }
  __syncwarp(gmask);
  for (int j = 2; j <= len; ++j) {
    for (int i = j - MIN_HRPN_LOOP - 1 - gl; i >= 1; i -= MONO_GW) {
      if (!isfinite(w.Hm[i][j])) continue;
      double S0 = w.Sm[i][j], H0 = w.Hm[i][j];
      const double T0 = (H0 + M_DHI) / (S0 + M_DSI + M_RC);
      double S1 = w.Sm[i + 1][j - 1] + m_Ss(T, w, i, j);
      double H1 = w.Hm[i + 1][j - 1] + m_Hs(T, w, i, j);
      const double T1 = (H1 + M_DHI) / (S1 + M_DSI + M_RC);
      if (S1 < kMinEntropyCutoff) { S1 = kMinEntropy; H1 = 0.0; }
      if (S0 < kMinEntropyCutoff) { S0 = kMinEntropy; H0 = 0.0; }
      if (T1 > T0) { w.Sm[i][j] = S1; w.Hm[i][j] = H1; } else { w.Sm[i][j] = S0; w.Hm[i][j] = H0; }
      double s = -1.0, h = INFINITY;
      m_loops(T, w, i, j, &s, &h, 0);
      s = -1.0; h = INFINITY;
      m_hairpin_loop(T, w, i, j, &s, &h, 0);
      if (isfinite(h)) {
        if (s < kMinEntropyCutoff) { s = kMinEntropy; h = 0.0; }
        w.Sm[i][j] = s; w.Hm[i][j] = h;
      }
    }
    __syncwarp(gmask);
  }
  // exterior fragments
  if (gl == 0) {
    w.send5[0] = w.send5[1] = -1.0;
    w.hend5[0] = w.hend5[1] = INFINITY;
    for (int i = 2; i <= len; i++) { w.send5[i] = kMinEntropy; w.hend5[i] = 0; }
  }
  __syncwarp(gmask);
  for (int i = 2; i <= len; ++i) {
    double eh[5], es[5], Tm[6];
    {
      double vh = 0.0, vs = 0.0;
      if (gl < 4) m_end5(T, w, i, gl + 1, &vh, &vs);          // variant gl + 1 on lane gl
      for (int v = 1; v <= 4; v++) {
        eh[v] = __shfl_sync(gmask, vh, v - 1, MONO_GW);
        es[v] = __shfl_sync(gmask, vs, v - 1, MONO_GW);
      }
    }
    if (gl == 0) {
    Tm[1] = (w.hend5[i - 1] + M_DHI) / (w.send5[i - 1] + M_DSI + M_RC);
    for (int v = 1; v <= 4; v++) Tm[v + 1] = (eh[v] + M_DHI) / (es[v] + M_DSI + M_RC);
    int mx;
    if (Tm[1] > Tm[2] && Tm[1] > Tm[3] && Tm[1] > Tm[4] && Tm[1] > Tm[5]) mx = 1;
    else if (Tm[2] > Tm[3] && Tm[2] > Tm[4] && Tm[2] > Tm[5]) mx = 2;
    else if (Tm[3] > Tm[4] && Tm[3] > Tm[5]) mx = 3;
    else if (Tm[4] > Tm[5]) mx = 4;
    else mx = 5;
    if (mx == 1) { w.send5[i] = w.send5[i - 1]; w.hend5[i] = w.hend5[i - 1]; }
    else {
      const int v = mx - 1;
      const double G = eh[v] - (temp_K * (es[v]));
      if (G < 0.0) { w.send5[i] = es[v]; w.hend5[i] = eh[v]; }
      else { w.send5[i] = w.send5[i - 1]; w.hend5[i] = w.hend5[i - 1]; }
    }
    }
    __syncwarp(gmask);
  }
  if (gl != 0) return;                       // traceback and result: lane 0
  const double mh = w.hend5[len], ms = w.send5[len];
  if (!isfinite(mh) || !isfinite(ms)) return;
  // traceback with an explicit stack
  int bpv[MONO_MAX + 2];
  for (int t = 0; t < len; ++t) bpv[t] = 0;
  struct { short i, j, m; } stk[4 * MONO_MAX];
  int sp = 0;
#define M_PUSH(a_, b_, c_) do { if (sp < 4 * MONO_MAX) { stk[sp].i = (short)(a_); stk[sp].j = (short)(b_); stk[sp].m = (short)(c_); sp++; } } while (0)
  M_PUSH(len, 0, 1);
  while (sp > 0) {
    sp--;
    int i = stk[sp].i;
    const int j = stk[sp].j, m = stk[sp].m;
    if (m == 1) {
      while (i >= 1 && eq2(w.send5[i], w.send5[i - 1]) && eq2(w.hend5[i], w.hend5[i - 1])) --i;
      if (i == 0) continue;
      bool handled = false;
      for (int v = 1; v <= 4 && !handled; v++) {
        double eh, es;
        m_end5(T, w, i, v, &eh, &es);
        if (!(eq2(w.send5[i], es) && eq2(w.hend5[i], eh))) continue;
        handled = true;
        const int kmax = m_end5_kmax(i, v);
        for (int k = 0; k <= kmax; ++k) {
          double xS, xH;
          m_end5_term(T, w, i, k, v, &xS, &xH);
          const int pi = (v == 1 || v == 3) ? k + 1 : k + 2;
          const int pj = (v <= 2) ? i : i - 1;
          if (eq2(w.send5[i], xS) && eq2(w.hend5[i], xH)) { M_PUSH(pi, pj, 0); break; }
          else if (eq2(w.send5[i], w.send5[k] + xS) && eq2(w.hend5[i], w.hend5[k] + xH)) { M_PUSH(pi, pj, 0); M_PUSH(k, 0, 1); break; }
        }
      }
    } else {
      bpv[i - 1] = j; bpv[j - 1] = i;
      double s1 = -1.0, h1 = INFINITY, s2 = -1.0, h2 = INFINITY;
      m_hairpin_loop(T, w, i, j, &s1, &h1, 1);
      m_loops(T, w, i, j, &s2, &h2, 2);
      if (eq2(w.Sm[i][j], m_Ss(T, w, i, j) + w.Sm[i + 1][j - 1]) && eq2(w.Hm[i][j], m_Hs(T, w, i, j) + w.Hm[i + 1][j - 1])) {
        M_PUSH(i + 1, j - 1, 0);
      } else if (eq2(w.Sm[i][j], s1) && eq2(w.Hm[i][j], h1)) {
        // hairpin loop closes here
      } else if (eq2(w.Sm[i][j], s2) && eq2(w.Hm[i][j], h2)) {
        bool done = false;
        for (int d = j - i - 3; d >= MIN_HRPN_LOOP + 1 && d >= j - i - 2 - w.maxLoop && !done; --d)
          for (int ii = i + 1; ii < j - d; ++ii) {
            const int jj = d + ii;
            double es = -1.0, eh = INFINITY;
            m_bulge_internal(T, w, i, j, ii, jj, &es, &eh, 1);
            if (eq2(w.Sm[i][j], es + w.Sm[ii][jj]) && eq2(w.Hm[i][j], eh + w.Hm[ii][jj])) { M_PUSH(ii, jj, 0); done = true; break; }
          }
      }
    }
  }
#undef M_PUSH
  int N = 0;
  for (int i = 1; i < len; ++i) if (bpv[i - 1] > 0) N++;
  out->n_bp = N / 2;
  const double t = (mh / (ms + (((N / 2) - 1) * saltCorr))) - kAbsZero;
  out->dg = mh - (temp_K * (ms + (((N / 2) - 1) * saltCorr)));
  out->ds = ms + (((N / 2) - 1) * saltCorr);
  out->dh = mh;
  out->tm = t;
  out->no_structure = 0;
}

template <int M>
__global__ void __launch_bounds__(64)
thal_mono_kernel(const uint64_t* __restrict__ codes, uint32_t n, int k, const ThalDeviceTables* T, double saltCorr, double temp_K,
                 int maxLoop, MonoWorkT<M>* work, msspe_thal_out* out) {
  // One thread per oligo (<= 2000 per run: latency, not throughput).  The DP matrices of the thread live in SHARED
  // memory when the launch provides it (MONO_SMEM_THREADS threads per block, one MonoWork each): the scalar recursion
  // is a chain of dependent matrix reads, ~30 cycles each from shared memory against ~600 from a per-thread global scratch.
  extern __shared__ __align__(16) unsigned char mono_smem[];
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  MonoWorkT<M>& w = work ? work[t] : reinterpret_cast<MonoWorkT<M>*>(mono_smem)[threadIdx.x];
  const uint64_t c = codes[t];
  w.len = k; w.maxLoop = maxLoop;
  for (int x = 0; x < k; x++) w.n1[x + 1] = (int)((c >> (2 * (k - 1 - x))) & 3u);
  w.n1[0] = 4; w.n1[k + 1] = 4;
  msspe_thal_out r;
  mono_run(T, w, saltCorr, temp_K, &r);
  out[t] = r;
}

// ---------------------------------------------------------------- device: oligotm + GC
// MONO_GW lanes per oligo (mono_run_group), the DP scratch of the group in shared memory: 4 oligos per 64-thread block
template <int M>
__global__ void __launch_bounds__(64)
thal_mono_group_kernel(const uint64_t* __restrict__ codes, uint32_t n, int k, const ThalDeviceTables* T, double saltCorr, double temp_K,
                       int maxLoop, msspe_thal_out* out) {
  extern __shared__ __align__(16) unsigned char mono_smem[];
  const int grp = threadIdx.x / MONO_GW, gl = threadIdx.x % MONO_GW;
  const uint32_t t = blockIdx.x * (64 / MONO_GW) + grp;
  if (t >= n) return;                                  // whole groups leave together
  const unsigned gmask = (MONO_GW == 32 ? 0xffffffffu : ((1u << MONO_GW) - 1u)) << ((threadIdx.x & 31) / MONO_GW * MONO_GW);
  MonoWorkT<M>& w = reinterpret_cast<MonoWorkT<M>*>(mono_smem)[grp];
  const uint64_t c = codes[t];
  if (gl == 0) {
    w.len = k; w.maxLoop = maxLoop;
    for (int x = 0; x < k; x++) w.n1[x + 1] = (int)((c >> (2 * (k - 1 - x))) & 3u);
    w.n1[0] = 4; w.n1[k + 1] = 4;
  }
  __syncwarp(gmask);
  msspe_thal_out r;
  mono_run_group(T, w, saltCorr, temp_K, &r, gl, gmask);
  if (gl == 0) out[t] = r;
}

struct OligoTmConsts { double salt_term; double conc_term[2]; };  // 0.368*(len-1)*ln(K/1000); 1.987*ln(C/4e9 | C/1e9)

__global__ void oligotm_kernel(const uint64_t* __restrict__ codes, uint32_t n, int k, const ThalDeviceTables* T, OligoTmConsts K,
                               double* __restrict__ tm, double* __restrict__ gc) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint64_t c = codes[t];
  double dh = 0, ds = 0;
  const int sym = ((k & 1) == 0 && revcomp_code(c, k) == c) ? 1 : 0;
  if (sym) ds += -1.4;
  const int f = (int)((c >> (2 * (k - 1))) & 3u), l = (int)(c & 3u);
  if (f == 0 || f == 3) { ds += 4.1; dh += 2300; } else { ds += -2.8; dh += 100; }
  if (l == 0 || l == 3) { ds += 4.1; dh += 2300; } else { ds += -2.8; dh += 100; }
  int ngc = 0;
  for (int i = 0; i < k; i++) {
    const int a = (int)((c >> (2 * (k - 1 - i))) & 3u);
    if (a == 1 || a == 2) ngc++;
    if (i + 1 < k) {
      const int b = (int)((c >> (2 * (k - 2 - i))) & 3u);
      dh += T->stackH[THAL_IDX4(a, b, 3 - a, 3 - b)];
      ds += T->stackS[THAL_IDX4(a, b, 3 - a, 3 - b)];
    }
  }
  ds = ds + K.salt_term;
  tm[t] = dh / (ds + K.conc_term[sym]) - 273.15;
  gc[t] = 100.0 * ngc / k;
}

// ---------------------------------------------------------------- host side
int ensure_tables(msspe_ctx* c) {
  if (c->d_thal) return MSSPE_OK;
  if (!c->raw_set) { msspe_thal_params_default(&c->raw); c->raw_set = true; }
  return msspe_thal_upload_tables(c);
}

// primer3_core is started without any parameter path (primer.rs:125-140, 151-160): its thal tables and oligotm's
// SantaLucia table are the ones compiled into Primer3.  So the primer3_core stand-ins (msspe_primer_thermo, msspe_kmer_stats*)
// always use the embedded tables; msspe_set_thal_params / `-path` (delta_g.rs:90) reaches only the ntthal stand-ins.
int ensure_p3_tables(msspe_ctx* c) {
  if (c->d_thal_p3) return MSSPE_OK;
  if (!c->raw_p3) { c->raw_p3 = new msspe_thal_raw_params(); msspe_thal_params_default(c->raw_p3); }
  ThalDeviceTables* h = new ThalDeviceTables();
  msspe_thal_expand(c->raw_p3, h);
  cudaError_t e = cudaMallocAsync(&c->d_thal_p3, sizeof(ThalDeviceTables), c->stream);
  if (e == cudaSuccess) e = cudaMemcpy(c->d_thal_p3, h, sizeof(ThalDeviceTables), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {   // the conditions of the primer3_core stand-ins never change: their per-run constants are built once
    // Primer3 defaults (primer.rs:125-140 sends no salt tags): mv 50, dv 1.5, dNTP 0.6, DNA 50 nM; thal at 37 C, maxLoop 30
    const msspe_thal_cond p3{50.0, 1.5, 0.6, 50.0, 37.0, 30, 0};
    if (!c->h_p3_consts) c->h_p3_consts = new ThalDimerConsts();
    build_dimer_consts(*h, p3, c->h_p3_consts);
    e = cudaMallocAsync((void**)&c->d_p3_consts, sizeof(ThalDimerConsts), c->stream);
    if (e == cudaSuccess) e = cudaMemcpy(c->d_p3_consts, c->h_p3_consts, sizeof(ThalDimerConsts), cudaMemcpyHostToDevice);
  }
  delete h;
  if (e != cudaSuccess) { c->set_error("upload Primer3 default tables: %s", cudaGetErrorString(e)); return MSSPE_ERR_CUDA; }
  return MSSPE_OK;
}

struct DeviceBuf {  // stream-ordered scratch, returned to the pool when the call ends
  void* p = nullptr;
  cudaStream_t st = nullptr;
  ~DeviceBuf() { if (p) cudaFreeAsync(p, st); }
};

// Matrix mode hands 32 consecutive columns of one row to the 32 lanes of a warp (thal_dimer_thread_kernel), and a warp is
// as slow as its busiest lane.  The work of a pair follows its pairing pattern: a row of the DP matrix has one paired cell
// per occurrence of the complementary base in the second oligo.  Visiting the columns in the order of their base
// composition (then of their code) gives the lanes of a warp second oligos with the same number of A, C, G and T, i.e. the
// same number of paired cells in every row.  The results carry the ORIGINAL column index.
int column_order(msspe_ctx* c, const uint64_t* codes, uint32_t n, uint32_t k, cudaStream_t st, DeviceBuf* dperm) {
  std::vector<uint64_t> key(n);
  for (uint32_t i = 0; i < n; i++) {
    uint32_t cnt[4] = {0, 0, 0, 0};
    for (uint32_t t = 0; t < k; t++) cnt[(codes[i] >> (2 * t)) & 3u]++;
    key[i] = ((uint64_t)cnt[0] << 18 | (uint64_t)cnt[1] << 12 | (uint64_t)cnt[2] << 6 | cnt[3]) << 32 | i;
  }
  static const int order_mode = getenv("MSSPE_THAL_ORDER") ? atoi(getenv("MSSPE_THAL_ORDER")) : 2;
  if (order_mode == 1) std::sort(key.begin(), key.end());       // composition, then position in the pool
  else {                                                        // 2: composition, then code from the 3' end; 3: from the 5' end
    std::sort(key.begin(), key.end(), [&](uint64_t x, uint64_t y) {
      if ((x >> 32) != (y >> 32)) return (x >> 32) < (y >> 32);
      uint64_t cx = codes[(uint32_t)x], cy = codes[(uint32_t)y];
      if (order_mode == 2) {   // reverse the 2-bit groups: the second oligo enters the DP matrix reversed, column 1 = its 3' end
        uint64_t rx = 0, ry = 0;
        for (uint32_t t = 0; t < k; t++) { rx = (rx << 2) | ((cx >> (2 * t)) & 3u); ry = (ry << 2) | ((cy >> (2 * t)) & 3u); }
        cx = rx; cy = ry;
      }
      return cx != cy ? cx < cy : (uint32_t)x < (uint32_t)y;
    });
  }
  std::vector<uint32_t> perm(n);
  for (uint32_t i = 0; i < n; i++) perm[i] = (uint32_t)key[i];
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dperm->st = st, dperm->p), (size_t)n * 4, st));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(dperm->p, perm.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));   // perm goes out of scope
  return MSSPE_OK;
}

int launch_dimer_warp(msspe_ctx* c, DimerArgs& A, cudaStream_t st) {   // thal_dimer_kernel: one warp per pair (small batches, latency)
  const int k = A.k;
  const int sub = k <= 16 ? 8 : 16;
  const int groups = DIMER_THREADS / 32;
  const size_t grp_bytes = ((size_t)k * k * 16 + (size_t)(k + 2) * 4 + 2 * (size_t)(k + 2) + 15) & ~(size_t)15;
  const size_t smem = ((sizeof(DimerShared) + 15) & ~(size_t)15) + groups * grp_bytes;
  if (smem > c->smem_optin) { c->set_error("thal dimer: %zu B shared memory needed, device offers %zu", smem, c->smem_optin); return MSSPE_ERR_CAPACITY; }
  int per_sm = 1;
  void (*kern)(const DimerArgs) = sub == 8 ? thal_dimer_kernel<8> : thal_dimer_kernel<16>;
  MSSPE_CUDA_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MSSPE_CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DIMER_THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  const unsigned long long blocks_needed = (A.n_pairs + groups - 1) / groups, resident = (unsigned long long)c->sm_count * per_sm;
  const unsigned grid = (unsigned)(blocks_needed < resident ? blocks_needed : resident);
  {
    KPROF(c, KP_DIMER, st, A.n_pairs * 16)
    kern<<<grid, DIMER_THREADS, smem, st>>>(A);
  }
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  return MSSPE_OK;
}

int launch_dimer(msspe_ctx* c, DimerArgs& A, cudaStream_t st) {
  if (A.n_pairs == 0) return MSSPE_OK;
  const int k = A.k;
  // MSSPE_THAL_KERNEL = thread (default for oligos <= 16 nt) | flat (default beyond) | legacy: the same results three ways
  const char* which = getenv("MSSPE_THAL_KERNEL");
  const bool legacy = which && !strcmp(which, "legacy");
  // one thread per pair needs a few pairs per lane of the whole GPU to pay (a pair is ~70,000 dependent thread
  // instructions: 2000 self-dimer pairs took 1.2 ms that way against 0.13 ms with a warp per pair)
  const bool many = A.n_pairs >= (unsigned long long)c->sm_count * DIMER_THREADS * 8ull;
  const bool thread = !legacy && k <= TK_MAX && !(which && !strcmp(which, "flat")) && (many || (which && !strcmp(which, "thread")));
  if (!thread && !(which && !strcmp(which, "flat")) && k <= 16) return launch_dimer_warp(c, A, st);
  const int sub = k <= 16 ? 8 : 16;  // legacy: lanes per cell in flight (inner rows <= 2*sub)
  const int groups = thread ? DIMER_THREADS : DIMER_THREADS / 32;   // pairs per block and sweep
  const size_t cell_bytes = (size_t)k * k * 16;
  const size_t grp_bytes = thread ? 0 : legacy ? ((cell_bytes + (size_t)(k + 2) * 4 + 2 * (size_t)(k + 2) + 15) & ~(size_t)15) : flat_group_bytes(k);
  const size_t smem = (((legacy ? sizeof(DimerShared) : thread ? sizeof(ThreadShared) : sizeof(FlatShared)) + 15) & ~(size_t)15) + (thread ? 0 : groups * grp_bytes);
  if (smem > c->smem_optin) { c->set_error("thal dimer: %zu B shared memory needed, device offers %zu", smem, c->smem_optin); return MSSPE_ERR_CAPACITY; }
  int per_sm = 1;
  unsigned long long blocks_needed = (A.n_pairs + groups - 1) / groups;
  void (*kern)(const DimerArgs) = thread ? thal_dimer_thread_kernel : !legacy ? thal_dimer_flat_kernel : (sub == 8 ? thal_dimer_kernel<8> : thal_dimer_kernel<16>);
  MSSPE_CUDA_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MSSPE_CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DIMER_THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  if (getenv("MSSPE_DEBUG_TIMERS")) fprintf(stderr, "[msspe] thal dimer kernel %s: %zu B shared memory per block, %d blocks per SM\n", thread ? "thread" : legacy ? "legacy" : "flat", smem, per_sm);
  const unsigned long long resident = (unsigned long long)c->sm_count * per_sm;
  const unsigned grid = (unsigned)(blocks_needed < resident ? blocks_needed : resident);
  DeviceBuf scratch;
  if (thread) {
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&(scratch.st = st, scratch.p), (size_t)grid * DIMER_THREADS * (size_t)(k * k) * sizeof(double2), st));
    A.scratch = (double2*)scratch.p;
  }
  {
    KPROF(c, KP_DIMER, st, A.n_pairs * 16)
    kern<<<grid, DIMER_THREADS, smem, st>>>(A);
  }
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  return MSSPE_OK;   // the scratch returns to the pool in stream order
}

// Hairpin launch: the per-thread DP scratch in shared memory (as many threads per block as fit) when that keeps the
// whole batch in one wave of blocks, else the global scratch `work` with 64-thread blocks.
template <int M>
int launch_mono_m(msspe_ctx* c, const uint64_t* d_codes, uint32_t n, int k, const ThalDimerConsts& K, void* work, msspe_thal_out* out,
                  cudaStream_t st, const ThalDeviceTables* T) {
  int per_block = (int)((c->smem_optin - 1024) / sizeof(MonoWorkT<M>));
  if (per_block > 64) per_block = 64;    // __launch_bounds__(64)
  const bool smem_ok = per_block >= 1 && !getenv("MSSPE_MONO_GLOBAL") && (uint64_t)n <= (uint64_t)per_block * (uint64_t)c->sm_count;  // one wave (measured: 0.49 vs 0.65 ms at 600, slower beyond one wave)
  if (M == 16 && k <= 16 && !getenv("MSSPE_MONO_THREAD")) {   // 16 lanes per oligo, 4 oligos per block
    const size_t smem = 4 * sizeof(MonoWorkT<M>);
    MSSPE_CUDA_TRY(c, cudaFuncSetAttribute(thal_mono_group_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { KPROF(c, KP_THERMO, st, (uint64_t)n * 56) thal_mono_group_kernel<M><<<(n + 3) / 4, 64, smem, st>>>(d_codes, n, k, T, K.saltCorr, K.t_user_K, K.maxLoop, out); }
  } else if (smem_ok) {
    // spread the batch over all SMs: fewer threads per block than fit, so that every SM gets a block
    int tpb = (int)((n + (uint32_t)c->sm_count - 1) / (uint32_t)c->sm_count);
    if (tpb < 1) tpb = 1;
    if (tpb > per_block) tpb = per_block;
    const size_t smem = (size_t)tpb * sizeof(MonoWorkT<M>);
    MSSPE_CUDA_TRY(c, cudaFuncSetAttribute(thal_mono_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { KPROF(c, KP_THERMO, st, (uint64_t)n * 56) thal_mono_kernel<M><<<(n + tpb - 1) / tpb, tpb, smem, st>>>(d_codes, n, k, T, K.saltCorr, K.t_user_K, K.maxLoop, nullptr, out); }
  } else {
    { KPROF(c, KP_THERMO, st, (uint64_t)n * 56) thal_mono_kernel<M><<<(n + 63) / 64, 64, 0, st>>>(d_codes, n, k, T, K.saltCorr, K.t_user_K, K.maxLoop, (MonoWorkT<M>*)work, out); }
  }
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  return MSSPE_OK;
}
// `work` holds n x sizeof(MonoWork) bytes (the largest scratch)
int launch_mono(msspe_ctx* c, const uint64_t* d_codes, uint32_t n, int k, const ThalDimerConsts& K, MonoWork* work, msspe_thal_out* out,
                cudaStream_t st, const ThalDeviceTables* T) {
  return k <= 16 ? launch_mono_m<16>(c, d_codes, n, k, K, work, out, st, T) : launch_mono_m<MONO_MAX>(c, d_codes, n, k, K, work, out, st, T);
}

int check_thal_args(msspe_ctx* c, uint32_t oligo_len, const msspe_thal_cond* cond) {
  if (oligo_len < 1 || oligo_len > MSSPE_MAX_OLIGO) { c->set_error("oligo length %u outside 1..%d", oligo_len, MSSPE_MAX_OLIGO); return MSSPE_ERR_INVALID; }
  if (cond && (cond->max_loop < 0 || cond->max_loop > 30)) { c->set_error("max_loop %d outside 0..30", cond->max_loop); return MSSPE_ERR_INVALID; }
  return MSSPE_OK;
}

}  // namespace

int msspe_thal_upload_tables(msspe_ctx* c) {
  ThalDeviceTables* h = new ThalDeviceTables();
  msspe_thal_expand(&c->raw, h);
  if (!c->d_thal) {
    cudaError_t e = cudaMallocAsync(&c->d_thal, sizeof(ThalDeviceTables), c->stream);
    if (e != cudaSuccess) { delete h; c->set_error("cudaMalloc thal tables: %s", cudaGetErrorString(e)); return MSSPE_ERR_CUDA; }
  }
  cudaError_t e = cudaMemcpy(c->d_thal, h, sizeof(ThalDeviceTables), cudaMemcpyHostToDevice);
  delete h;
  if (e != cudaSuccess) { c->set_error("upload thal tables: %s", cudaGetErrorString(e)); return MSSPE_ERR_CUDA; }
  return MSSPE_OK;
}

void msspe_thal_free_tables(msspe_ctx* c) {
  if (c->d_thal) msspe_dev_free(c, c->d_thal);
  c->d_thal = nullptr;
  if (c->d_thal_p3) msspe_dev_free(c, c->d_thal_p3);
  c->d_thal_p3 = nullptr;
  if (c->d_p3_consts) msspe_dev_free(c, c->d_p3_consts);
  c->d_p3_consts = nullptr;
  delete c->h_p3_consts; c->h_p3_consts = nullptr;
  delete c->raw_p3; c->raw_p3 = nullptr;
}

extern "C" int msspe_set_thal_params(msspe_ctx* c, const msspe_thal_raw_params* p) {
  if (!c || !p) return MSSPE_ERR_INVALID;
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  c->raw = *p;
  c->raw_set = true;
  return msspe_thal_upload_tables(c);
}

static int thal_pairs_impl(msspe_ctx* c, const uint64_t* a, const uint64_t* b, uint64_t n_pairs, uint32_t oligo_len,
                           int32_t type, const msspe_thal_cond* cond, msspe_thal_out* out, uint8_t* pairing) {
  if (!c) return MSSPE_ERR_INVALID;
  if (!cond || (n_pairs && (!a || !out))) { c->set_error("msspe_thal_pairs: null argument"); return MSSPE_ERR_INVALID; }
  if (type != MSSPE_THAL_ANY && type != MSSPE_THAL_END1 && type != MSSPE_THAL_HAIRPIN) { c->set_error("msspe_thal_pairs: unsupported type %d", type); return MSSPE_ERR_INVALID; }
  if (type != MSSPE_THAL_HAIRPIN && n_pairs && !b) { c->set_error("msspe_thal_pairs: null argument"); return MSSPE_ERR_INVALID; }
  int rc = check_thal_args(c, oligo_len, cond);
  if (rc) return rc;
  if (n_pairs == 0) return MSSPE_OK;
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  rc = ensure_tables(c);
  if (rc) return rc;
  cudaStream_t st = c->stream;
  ThalDeviceTables* hT = new ThalDeviceTables();
  msspe_thal_expand(&c->raw, hT);
  ThalDimerConsts K;
  build_dimer_consts(*hT, *cond, &K);
  delete hT;
  DeviceBuf da, db, dout, dK, dwork, dpair;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&(da.st = c->stream, da.p), n_pairs * 8, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dout.st = c->stream, dout.p), n_pairs * sizeof(msspe_thal_out), c->stream));
  if (pairing) {
    if (type == MSSPE_THAL_HAIRPIN) { c->set_error("msspe_thal_pairs_aligned: dimer types only"); return MSSPE_ERR_INVALID; }
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dpair.st = c->stream, dpair.p), n_pairs * MSSPE_MAX_OLIGO, c->stream));
    MSSPE_CUDA_TRY(c, cudaMemsetAsync(dpair.p, 0, n_pairs * MSSPE_MAX_OLIGO, st));
  }
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(da.p, a, n_pairs * 8, cudaMemcpyHostToDevice, st));
  if (type == MSSPE_THAL_HAIRPIN) {
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dwork.st = c->stream, dwork.p), n_pairs * sizeof(MonoWork), c->stream));
    rc = launch_mono(c, (const uint64_t*)da.p, (uint32_t)n_pairs, (int)oligo_len, K, (MonoWork*)dwork.p, (msspe_thal_out*)dout.p, st, c->d_thal);
    if (rc) return rc;
  } else {
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&(db.st = c->stream, db.p), n_pairs * 8, c->stream));
    MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dK.st = c->stream, dK.p), sizeof(ThalDimerConsts), c->stream));
    MSSPE_CUDA_TRY(c, cudaMemcpyAsync(db.p, b, n_pairs * 8, cudaMemcpyHostToDevice, st));
    MSSPE_CUDA_TRY(c, cudaMemcpyAsync(dK.p, &K, sizeof K, cudaMemcpyHostToDevice, st));
    DimerArgs A{};
    A.a = (const uint64_t*)da.p; A.b = (const uint64_t*)db.p; A.n_pairs = n_pairs; A.matrix = 0; A.k = (int)oligo_len; A.type = type;
    A.T = c->d_thal; A.C = (const ThalDimerConsts*)dK.p; A.out = (msspe_thal_out*)dout.p; A.pairing = (uint8_t*)dpair.p;
    rc = launch_dimer(c, A, st);
    if (rc) return rc;
  }
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(out, dout.p, n_pairs * sizeof(msspe_thal_out), cudaMemcpyDeviceToHost, st));
  if (pairing) MSSPE_CUDA_TRY(c, cudaMemcpyAsync(pairing, dpair.p, n_pairs * MSSPE_MAX_OLIGO, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  return MSSPE_OK;
}

extern "C" int msspe_thal_pairs(msspe_ctx* c, const uint64_t* a, const uint64_t* b, uint64_t n_pairs, uint32_t oligo_len,
                                int32_t type, const msspe_thal_cond* cond, msspe_thal_out* out) {
  return thal_pairs_impl(c, a, b, n_pairs, oligo_len, type, cond, out, nullptr);
}

extern "C" int msspe_thal_pairs_aligned(msspe_ctx* c, const uint64_t* a, const uint64_t* b, uint64_t n_pairs, uint32_t oligo_len,
                                        int32_t type, const msspe_thal_cond* cond, msspe_thal_out* out, uint8_t* pairing) {
  if (c && !pairing) { c->set_error("msspe_thal_pairs_aligned: null argument"); return MSSPE_ERR_INVALID; }
  return thal_pairs_impl(c, a, b, n_pairs, oligo_len, type, cond, out, pairing);
}

extern "C" int msspe_primer_thermo(msspe_ctx* c, const uint64_t* codes, uint32_t n, uint32_t oligo_len, double* tm, double* gc,
                                   double* self_any, double* self_end, double* hairpin) {
  if (!c) return MSSPE_ERR_INVALID;
  if (n && (!codes || !tm || !gc || !self_any || !self_end || !hairpin)) { c->set_error("msspe_primer_thermo: null argument"); return MSSPE_ERR_INVALID; }
  int rc = check_thal_args(c, oligo_len, nullptr);
  if (rc) return rc;
  if (n == 0) return MSSPE_OK;
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  rc = ensure_p3_tables(c);
  if (rc) return rc;
  cudaStream_t st = c->stream;
  // Primer3 defaults (primer.rs:125-140 sends no salt tags): mv 50, dv 1.5, dNTP 0.6, DNA 50 nM; thal at 37 C, maxLoop 30
  const msspe_thal_cond p3{50.0, 1.5, 0.6, 50.0, 37.0, 30, 0};
  const ThalDimerConsts& K = *c->h_p3_consts;      // built by ensure_p3_tables, resident on the device
  OligoTmConsts OK;
  {
    double dv = p3.dv, dntp = p3.dntp;
    if (dv == 0) dntp = 0;
    if (dv < dntp) dv = dntp;
    const double Ksalt = p3.mv + 120 * sqrt(dv - dntp);
    OK.salt_term = 0.368 * ((int)oligo_len - 1) * log(Ksalt / 1000.0);
    OK.conc_term[0] = 1.987 * log(p3.dna_conc / 4000000000.0);
    OK.conc_term[1] = 1.987 * log(p3.dna_conc / 1000000000.0);
  }
  DeviceBuf dcodes, dtm, dgc, dout, dwork;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dcodes.st = c->stream, dcodes.p), (size_t)n * 8, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dtm.st = c->stream, dtm.p), (size_t)n * 8, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dgc.st = c->stream, dgc.p), (size_t)n * 8, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dout.st = c->stream, dout.p), (size_t)n * 3 * sizeof(msspe_thal_out), c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dwork.st = c->stream, dwork.p), (size_t)n * sizeof(MonoWork), c->stream));
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[6], st));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(dcodes.p, codes, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  msspe_thal_out* o3 = (msspe_thal_out*)dout.p;
  // the hairpin kernel (one thread per primer: latency, a fraction of the SMs) runs on the second stream beside oligotm and
  // the two self-dimer launches: 0.5 ms || 0.35 ms instead of their sum
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev_fork, st));
  MSSPE_CUDA_TRY(c, cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
  rc = launch_mono(c, (const uint64_t*)dcodes.p, n, (int)oligo_len, K, (MonoWork*)dwork.p, o3 + (size_t)2 * n, c->stream2, c->d_thal_p3);
  if (rc) return rc;
  oligotm_kernel<<<(n + 127) / 128, 128, 0, st>>>((const uint64_t*)dcodes.p, n, (int)oligo_len, c->d_thal_p3, OK, (double*)dtm.p, (double*)dgc.p);
  c->timing.kernel_launches++;
  for (int pass = 0; pass < 2; pass++) {  // SELF_ANY_TH, SELF_END_TH: thal(s, s)
    DimerArgs A{};
    A.a = (const uint64_t*)dcodes.p; A.b = (const uint64_t*)dcodes.p; A.n_pairs = n; A.matrix = 0; A.k = (int)oligo_len;
    A.type = pass == 0 ? MSSPE_THAL_ANY : MSSPE_THAL_END1;
    A.T = c->d_thal_p3; A.C = c->d_p3_consts; A.out = o3 + (size_t)pass * n;
    rc = launch_dimer(c, A, st);
    if (rc) return rc;
  }
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev_join, c->stream2));
  MSSPE_CUDA_TRY(c, cudaStreamWaitEvent(st, c->ev_join, 0));
  MSSPE_CUDA_TRY(c, cudaGetLastError());
  // results come back through the context's pinned staging buffer (a device-to-pageable copy goes through the driver's own
  // staging, chunk by chunk, with the host blocked)
  const size_t out_bytes = (size_t)n * 3 * sizeof(msspe_thal_out), need = out_bytes + (size_t)n * 16;
  if (c->h_stage_bytes < need) {
    if (c->h_stage) cudaFreeHost(c->h_stage);
    c->h_stage = nullptr; c->h_stage_bytes = 0;
    MSSPE_CUDA_TRY(c, cudaMallocHost(&c->h_stage, need));
    c->h_stage_bytes = need;
  }
  const msspe_thal_out* h = reinterpret_cast<const msspe_thal_out*>(c->h_stage);
  double* h_tm = reinterpret_cast<double*>((unsigned char*)c->h_stage + out_bytes);
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(h_tm, dtm.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(h_tm + n, dgc.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(c->h_stage, dout.p, out_bytes, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[7], st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  MSSPE_CUDA_TRY(c, cudaEventElapsedTime(&c->timing.thermo_ms, c->ev[6], c->ev[7]));
  memcpy(tm, h_tm, (size_t)n * 8);
  memcpy(gc, h_tm + n, (size_t)n * 8);
  for (uint32_t i = 0; i < n; i++) {  // align_thermod: negative or absent structure temperatures report 0
    const double a = h[i].tm, e = h[(size_t)n + i].tm, hp = h[(size_t)2 * n + i].tm;
    self_any[i] = a < 0.0 ? 0.0 : a;
    self_end[i] = e < 0.0 ? 0.0 : e;
    hairpin[i] = hp < 0.0 ? 0.0 : hp;
  }
  return MSSPE_OK;
}

// The compacted lists of a cross-dimer call stay on the device (ctx-owned, valid until the next such call): what a rank
// of the row-tiled matrix hands to ONE all_gather on device buffers (msspe_b200/distributed.py) instead of a host round trip.
extern "C" int msspe_cross_dimer_device(msspe_ctx* c, const uint64_t* codes, uint32_t n, uint32_t oligo_len, const msspe_thal_cond* cond,
                                        uint32_t row_begin, uint32_t row_end, double dg_limit, uint64_t edge_capacity, uint64_t nostruct_capacity,
                                        const msspe_dimer_edge** d_edges, uint64_t* n_edges, const uint64_t** d_nostruct, uint64_t* n_nostruct) {
  if (!c) return MSSPE_ERR_INVALID;
  if (!cond || !n_edges || !n_nostruct || !d_edges || !d_nostruct || (n && !codes) || row_begin > row_end || row_end > n) { c->set_error("msspe_cross_dimer_device: bad argument"); return MSSPE_ERR_INVALID; }
  int rc = check_thal_args(c, oligo_len, cond);
  if (rc) return rc;
  *n_edges = 0; *n_nostruct = 0; *d_edges = nullptr; *d_nostruct = nullptr;
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  if (c->xd_edges) { cudaFreeAsync(c->xd_edges, st); c->xd_edges = nullptr; }
  if (c->xd_nostruct) { cudaFreeAsync(c->xd_nostruct, st); c->xd_nostruct = nullptr; }
  const unsigned long long n_pairs = (unsigned long long)(row_end - row_begin) * n;
  if (n_pairs == 0) return MSSPE_OK;
  rc = ensure_tables(c);
  if (rc) return rc;
  ThalDeviceTables* hT = new ThalDeviceTables();
  msspe_thal_expand(&c->raw, hT);
  ThalDimerConsts K;
  build_dimer_consts(*hT, *cond, &K);
  delete hT;
  DeviceBuf dcodes, dK, dcnt;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dcodes.st = st, dcodes.p), (size_t)n * 8, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dK.st = st, dK.p), sizeof K, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dcnt.st = st, dcnt.p), 16, st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&c->xd_edges, (size_t)(edge_capacity ? edge_capacity : 1) * sizeof(msspe_dimer_edge), st));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&c->xd_nostruct, (size_t)(nostruct_capacity ? nostruct_capacity : 1) * 8, st));
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[6], st));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(dcodes.p, codes, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(dK.p, &K, sizeof K, cudaMemcpyHostToDevice, st));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(dcnt.p, 0, 16, st));
  DimerArgs A{};
  A.a = (const uint64_t*)dcodes.p; A.b = (const uint64_t*)dcodes.p; A.n_pairs = n_pairs; A.n = n; A.row_begin = row_begin; A.matrix = 1;
  A.k = (int)oligo_len; A.type = MSSPE_THAL_ANY; A.T = c->d_thal; A.C = (const ThalDimerConsts*)dK.p; A.out = nullptr;
  A.dg_limit = dg_limit; A.edges = c->xd_edges; A.edge_cap = edge_capacity; A.n_edges = (unsigned long long*)dcnt.p;
  A.nostruct = c->xd_nostruct; A.nostruct_cap = nostruct_capacity; A.n_nostruct = (unsigned long long*)dcnt.p + 1;
  A.dbg = 0;
  DeviceBuf dperm;
  if (n >= 64 && !getenv("MSSPE_THAL_NO_ORDER")) { rc = column_order(c, codes, n, oligo_len, st, &dperm); if (rc) return rc; A.colperm = (const uint32_t*)dperm.p; }
  rc = launch_dimer(c, A, st);
  if (rc) return rc;
  unsigned long long cnt[2] = {0, 0};
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(cnt, dcnt.p, 16, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[7], st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  MSSPE_CUDA_TRY(c, cudaEventElapsedTime(&c->timing.dimer_ms, c->ev[6], c->ev[7]));
  *n_edges = cnt[0]; *n_nostruct = cnt[1];
  if (cnt[0] > edge_capacity || cnt[1] > nostruct_capacity) {
    c->set_error("msspe_cross_dimer_device: %llu edges / %llu structure-less pairs exceed the capacity (%llu / %llu)", cnt[0], cnt[1],
                 (unsigned long long)edge_capacity, (unsigned long long)nostruct_capacity);
    return MSSPE_ERR_CAPACITY;
  }
  *d_edges = c->xd_edges; *d_nostruct = c->xd_nostruct;
  return MSSPE_OK;
}

extern "C" int msspe_cross_dimer(msspe_ctx* c, const uint64_t* codes, uint32_t n, uint32_t oligo_len, const msspe_thal_cond* cond,
                                 uint32_t row_begin, uint32_t row_end, double dg_limit, msspe_dimer_edge* edges,
                                 uint64_t edge_capacity, uint64_t* n_edges, uint64_t* nostruct_pairs, uint64_t nostruct_capacity,
                                 uint64_t* n_nostruct) {
  if (!c) return MSSPE_ERR_INVALID;
  if (!cond || !n_edges || !n_nostruct || (n && !codes) || row_begin > row_end || row_end > n) { c->set_error("msspe_cross_dimer: bad argument"); return MSSPE_ERR_INVALID; }
  int rc = check_thal_args(c, oligo_len, cond);
  if (rc) return rc;
  *n_edges = 0; *n_nostruct = 0;
  const unsigned long long n_pairs = (unsigned long long)(row_end - row_begin) * n;
  if (n_pairs == 0) return MSSPE_OK;
  MSSPE_CUDA_TRY(c, cudaSetDevice(c->device));
  rc = ensure_tables(c);
  if (rc) return rc;
  cudaStream_t st = c->stream;
  ThalDeviceTables* hT = new ThalDeviceTables();
  msspe_thal_expand(&c->raw, hT);
  ThalDimerConsts K;
  build_dimer_consts(*hT, *cond, &K);
  delete hT;
  DeviceBuf dcodes, dK, dedges, dnos, dcnt;
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dcodes.st = c->stream, dcodes.p), (size_t)n * 8, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dK.st = c->stream, dK.p), sizeof K, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dedges.st = c->stream, dedges.p), (size_t)(edge_capacity ? edge_capacity : 1) * sizeof(msspe_dimer_edge), c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dnos.st = c->stream, dnos.p), (size_t)(nostruct_capacity ? nostruct_capacity : 1) * 8, c->stream));
  MSSPE_CUDA_TRY(c, cudaMallocAsync(&(dcnt.st = c->stream, dcnt.p), 16, c->stream));
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[6], st));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(dcodes.p, codes, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(dK.p, &K, sizeof K, cudaMemcpyHostToDevice, st));
  MSSPE_CUDA_TRY(c, cudaMemsetAsync(dcnt.p, 0, 16, st));
  DimerArgs A{};
  A.a = (const uint64_t*)dcodes.p; A.b = (const uint64_t*)dcodes.p; A.n_pairs = n_pairs; A.n = n; A.row_begin = row_begin; A.matrix = 1;
  A.k = (int)oligo_len; A.type = MSSPE_THAL_ANY; A.T = c->d_thal; A.C = (const ThalDimerConsts*)dK.p; A.out = nullptr;
  A.dg_limit = dg_limit; A.edges = (msspe_dimer_edge*)dedges.p; A.edge_cap = edge_capacity; A.n_edges = (unsigned long long*)dcnt.p;
  A.nostruct = (uint64_t*)dnos.p; A.nostruct_cap = nostruct_capacity; A.n_nostruct = (unsigned long long*)dcnt.p + 1;
  A.dbg = getenv("MSSPE_THAL_DBG") ? atoi(getenv("MSSPE_THAL_DBG")) : 0;
  DeviceBuf dperm;
  if (n >= 64 && !getenv("MSSPE_THAL_NO_ORDER")) { rc = column_order(c, codes, n, oligo_len, st, &dperm); if (rc) return rc; A.colperm = (const uint32_t*)dperm.p; }
  rc = launch_dimer(c, A, st);
  if (rc) return rc;
  unsigned long long cnt[2] = {0, 0};
  MSSPE_CUDA_TRY(c, cudaMemcpyAsync(cnt, dcnt.p, 16, cudaMemcpyDeviceToHost, st));
  MSSPE_CUDA_TRY(c, cudaEventRecord(c->ev[7], st));
  MSSPE_CUDA_TRY(c, cudaStreamSynchronize(st));
  MSSPE_CUDA_TRY(c, cudaEventElapsedTime(&c->timing.dimer_ms, c->ev[6], c->ev[7]));
  *n_edges = cnt[0]; *n_nostruct = cnt[1];
  if (cnt[0] > edge_capacity || cnt[1] > nostruct_capacity) {
    c->set_error("msspe_cross_dimer: %llu edges / %llu structure-less pairs exceed the caller's capacity (%llu / %llu)", cnt[0], cnt[1],
                 (unsigned long long)edge_capacity, (unsigned long long)nostruct_capacity);
    return MSSPE_ERR_CAPACITY;
  }
  if (cnt[0]) {
    MSSPE_CUDA_TRY(c, cudaMemcpy(edges, dedges.p, cnt[0] * sizeof(msspe_dimer_edge), cudaMemcpyDeviceToHost));
    std::sort(edges, edges + cnt[0], [](const msspe_dimer_edge& x, const msspe_dimer_edge& y) { return x.pair < y.pair; });
  }
  if (cnt[1]) {
    MSSPE_CUDA_TRY(c, cudaMemcpy(nostruct_pairs, dnos.p, cnt[1] * 8, cudaMemcpyDeviceToHost));
    std::sort(nostruct_pairs, nostruct_pairs + cnt[1]);
  }
  return MSSPE_OK;
}
