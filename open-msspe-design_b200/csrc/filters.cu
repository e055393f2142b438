// filters.cu -- get_kmer_stats / get_tm_stat / tm_in_threshold / is_run / filter_kmers of the reference
// (od-msspe/src/main.rs:408-516) behind the C ABI.  The five thermodynamic numbers come from the device
// (msspe_primer_thermo); what follows is <= 1000 values per direction of strictly ordered f32 arithmetic and
// Primer3's text formatting, which is host work by nature and must match the reference bit for bit.
#include <charconv>
#include <cmath>
#include <cstdlib>

#include "engine.cuh"

namespace {
// Primer3 prints "%.3f" / "%.2f", parse_primer3_output reads an f32 (primer.rs:67-114).  std::to_chars(fixed, prec)
// and std::from_chars are correctly rounded like glibc's printf / strtof (checked on 6 M values incl. exact ties:
// identical bits) and four times faster, which matters at 5 numbers x 2000 candidates per run.
float via_text(double v, int prec) {
  char b[64];
  const auto r = std::to_chars(b, b + sizeof b, v, std::chars_format::fixed, prec);
  float f = 0.0f;
  std::from_chars(b, r.ptr, f);
  return f;
}
}  // namespace

namespace {
// Everything after the device numbers, for the candidates of ONE direction (statistics are per direction).
void stats_of_direction(const uint64_t* codes, uint32_t n, uint32_t oligo_len, const msspe_filter_cfg* cfg, const double* tm, const double* gc,
                        const double* sa, const double* se, const double* hp, msspe_kmer_stat* out) {
  if (n == 0) return;
  for (uint32_t i = 0; i < n; i++) {
    msspe_kmer_stat& s = out[i];
    s.code = codes[i];
    s.tm = via_text(tm[i], 3);
    s.gc_percent = via_text(gc[i], 3);
    s.self_any_th = via_text(sa[i], 2);
    s.self_end_th = via_text(se[i], 2);
    s.hairpin_th = via_text(hp[i], 2);
  }
  // get_tm_stat (main.rs:462-467): f32 sum in candidate order; std-dev 0.1.0 = sqrt(sum((x-mean)^2)/(n-1))
  float sum = 0.0f;
  for (uint32_t i = 0; i < n; i++) sum += out[i].tm;
  const float mean = sum / (float)n;
  float sq = 0.0f;
  for (uint32_t i = 0; i < n; i++) { const float d = out[i].tm - mean; sq += d * d; }
  const float sd = std::sqrt(sq / (float)((double)n - 1.0));
  for (uint32_t i = 0; i < n; i++) {
    msspe_kmer_stat& s = out[i];
    s.mean = mean; s.std = sd;
    s.tm_ok = std::fabs(s.tm - mean) <= (cfg->tm_stddev * sd) ? 1 : 0;      // tm_in_threshold (main.rs:469-471)
    {  // is_run (main.rs:478-490): true iff the last six bases are equal
      int runs = 0; uint32_t last = 0xFFu;
      for (uint32_t t = 0; t < oligo_len; t++) {
        const uint32_t b = (uint32_t)(codes[i] >> (2 * (oligo_len - 1 - t))) & 3u;
        if (b == last) runs++; else runs = 0;
        last = b;
      }
      s.runs = runs >= 5 ? 1 : 0;
    }
    const bool p_any = !cfg->check_self_dimers || (s.self_any_th < cfg->max_self_dimer_any_tm);   // filter_kmers (main.rs:492-516)
    const bool p_end = !cfg->check_self_dimers || (s.self_end_th < cfg->max_self_dimer_end_tm);
    const bool p_hp = !cfg->check_hairpin || (s.hairpin_th < cfg->max_hairpin_tm);
    const bool p_mm = cfg->disable_min_max_tm || (s.tm > cfg->min_tm && s.tm < cfg->max_tm);
    const bool p_sd = cfg->disable_tm_stddev || s.tm_ok;
    s.keep = (p_any && p_end && p_hp && p_mm && p_sd && !s.runs) ? 1 : 0;
    s.reserved = 0;
  }
}
}  // namespace

extern "C" int msspe_kmer_stats(msspe_ctx* c, const uint64_t* codes, uint32_t n, uint32_t oligo_len, const msspe_filter_cfg* cfg,
                                msspe_kmer_stat* out) {
  if (!c) return MSSPE_ERR_INVALID;
  if (!cfg || (n && (!codes || !out))) { c->set_error("msspe_kmer_stats: null argument"); return MSSPE_ERR_INVALID; }
  if (n == 0) return MSSPE_OK;
  std::vector<double> tm(n), gc(n), sa(n), se(n), hp(n);
  int rc = msspe_primer_thermo(c, codes, n, oligo_len, tm.data(), gc.data(), sa.data(), se.data(), hp.data());
  if (rc) return rc;
  stats_of_direction(codes, n, oligo_len, cfg, tm.data(), gc.data(), sa.data(), se.data(), hp.data(), out);
  return MSSPE_OK;
}

// Both directions (main.rs:723-724 calls get_kmer_stats twice): ONE device batch for all candidates, then the
// per-direction statistics and verdicts exactly as two msspe_kmer_stats calls would give them.
extern "C" int msspe_kmer_stats_both(msspe_ctx* c, const uint64_t* fwd_codes, uint32_t n_fwd, const uint64_t* rev_codes, uint32_t n_rev,
                                     uint32_t oligo_len, const msspe_filter_cfg* cfg, msspe_kmer_stat* out_fwd, msspe_kmer_stat* out_rev) {
  if (!c) return MSSPE_ERR_INVALID;
  if (!cfg || (n_fwd && (!fwd_codes || !out_fwd)) || (n_rev && (!rev_codes || !out_rev))) { c->set_error("msspe_kmer_stats_both: null argument"); return MSSPE_ERR_INVALID; }
  const uint32_t n = n_fwd + n_rev;
  if (n == 0) return MSSPE_OK;
  std::vector<uint64_t> codes(n);
  for (uint32_t i = 0; i < n_fwd; i++) codes[i] = fwd_codes[i];
  for (uint32_t i = 0; i < n_rev; i++) codes[n_fwd + i] = rev_codes[i];
  std::vector<double> tm(n), gc(n), sa(n), se(n), hp(n);
  int rc = msspe_primer_thermo(c, codes.data(), n, oligo_len, tm.data(), gc.data(), sa.data(), se.data(), hp.data());
  if (rc) return rc;
  stats_of_direction(fwd_codes, n_fwd, oligo_len, cfg, tm.data(), gc.data(), sa.data(), se.data(), hp.data(), out_fwd);
  stats_of_direction(rev_codes, n_rev, oligo_len, cfg, tm.data() + n_fwd, gc.data() + n_fwd, sa.data() + n_fwd, se.data() + n_fwd, hp.data() + n_fwd, out_rev);
  return MSSPE_OK;
}
