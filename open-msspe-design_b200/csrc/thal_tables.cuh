// thal_tables.cuh -- nearest-neighbour parameter tables as the thermodynamic kernels consume them.
// Static part (independent of salt/concentration): expanded on the host from msspe_thal_raw_params with the
// joint-infinity rule (an "inf" in either the .ds or the .dh file makes the entry (S=-1, H=+inf)) and the
// sequence-end sentinel N=4; layout per table documented in od-msspe/primer3_config/interpretations/*.
#pragma once
#include <cstdint>

#define THAL_IDX4(a, b, c, d) ((((a)*5 + (b)) * 5 + (c)) * 5 + (d))
#define THAL_IDX3(a, b, c) (((a)*5 + (b)) * 5 + (c))

struct ThalDeviceTables {
  double stackS[625], stackH[625];
  double stackint2S[625], stackint2H[625];
  double tstackS[625], tstackH[625];
  double tstack2S[625], tstack2H[625];
  double dangle3S[125], dangle3H[125];
  double dangle5S[125], dangle5H[125];
  double interiorS[30], interiorH[30], bulgeS[30], bulgeH[30], hairpinS[30], hairpinH[30];
  double atpS[25], atpH[25];
  int nTriS, nTriH, nTetraS, nTetraH;
  // loop bonus tables sorted by key; key = base-5 digits of the loop sequence (5 resp. 6 bases)
  uint32_t triKeyS[32], triKeyH[32], tetraKeyS[128], tetraKeyH[128];
  double triS[32], triH[32], tetraS[128], tetraH[128];
};

// Per-call constants of a dimer run (depend on msspe_thal_cond): the two end-of-duplex tables are tabulated on
// the host because they depend only on a 2x2 base context and on RC (symmetric / non-symmetric pair).
//   lsh[sym][a][a_prev][b_prev], rsh[sym][a][a_next][b_next]  with b = 3 - a  (a base pair is required)
struct ThalDimerConsts {
  double lshS[2][100], lshH[2][100];
  double rshS[2][100], rshH[2][100];
  double RC[2];          // [0] non-symmetric, [1] both oligos self-complementary
  double saltCorr;
  double t_user_K;       // cond.temp_c + 273.15
  int maxLoop;
  int pad;
  double t0[2][2500];    // [sym][li * 25 + rn]: (lshH + 200 + rshH) / (lshS - 5.7 + rshS + RC), the first quotient of the stack test
};
