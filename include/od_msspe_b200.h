/*
 * od_msspe_b200.h -- C ABI of the B200-native engine for the open-msspe-design hot path.
 *
 * This is the drop-in boundary.  Every entry point below replaces one call
 * the reference's `main` makes into its own k-mer engine or into a Primer3 subprocess; the reference
 * location each one stands in for is cited as od-msspe/src/<file>:<line>.  A maintainer binds these
 * from Rust with a plain `extern "C"` block; tests bind them with ctypes.
 *
 * Conventions
 *   - plain C types only; caller allocates every output buffer; the context owns all device memory.
 *   - return value 0 = MSSPE_OK, negative = error; text via msspe_last_error().  Nothing aborts/throws.
 *   - there is NO CPU fallback: every compute entry point fails with MSSPE_ERR_CUDA when no sm_100
 *     device / driver is usable.
 *   - a context is single-caller (not re-entrant), like the single-threaded reference.
 *   - k-mers cross the boundary as 2-bit big-endian codes (A=0 C=1 G=2 T=3, first base in the most
 *     significant used bits), which preserves the reference's lexicographic String order
 *     (main.rs:320-324) for words of equal length.
 */
#ifndef OD_MSSPE_B200_H
#define OD_MSSPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSSPE_ABI_VERSION 2

#define MSSPE_OK 0
#define MSSPE_ERR_INVALID (-1)     /* bad argument (the reference would panic!/clap-error) */
#define MSSPE_ERR_CUDA (-2)        /* CUDA runtime/driver failure, or no device */
#define MSSPE_ERR_NOMEM (-3)
#define MSSPE_ERR_STATE (-4)       /* call order violated (e.g. select before build) */
#define MSSPE_ERR_IO (-5)
#define MSSPE_ERR_CAPACITY (-6)    /* caller buffer too small; required size reported */

#define MSSPE_DIR_FWD 0            /* constants.rs:22 SEQ_DIR_FWD */
#define MSSPE_DIR_REV 1            /* constants.rs:23 SEQ_DIR_REV */

#define MSSPE_NO_KMER UINT64_MAX   /* sentinel slot: invalid (non-ACGT) or duplicate within the window */
#define MSSPE_MAX_KMER 31         /* 2-bit codes in a u64 with UINT64_MAX free as the sentinel */
#define MSSPE_MAX_OLIGO 32         /* thal kernels: oligo length limit (reference uses 13..15) */

typedef struct msspe_ctx msspe_ctx;

/* PartitioningOption, main.rs:189-194, plus device selection. */
typedef struct {
  uint32_t kmer_size;           /* --kmer-size            (config.rs:20)  1..31 */
  uint32_t window_size;         /* --window-size          (config.rs:23)  segment length W */
  uint32_t overlap_size;        /* --overlap-size         (config.rs:26)  used as the STEP (main.rs:178) */
  uint32_t search_windows_size; /* --search-windows-size  (config.rs:39)  head/tail width w */
  int32_t device;               /* CUDA device ordinal */
  uint32_t flags;               /* reserved, 0 */
} msspe_config;

/* One greedy winner, main.rs:47-51 KmerFrequency (+ diagnostics the reference only logs). */
typedef struct {
  uint64_t code;       /* 2-bit code of the winning word */
  uint32_t freq;       /* KmerFrequency.frequency: live segments containing it when chosen */
  uint32_t n_tied;     /* how many k-mers shared that frequency (diagnostic) */
  float tie_score;     /* partition_tie_score of the winner, main.rs:261-283 (diagnostic) */
  uint32_t reserved;
} msspe_candidate;

/* How the device-resident greedy loop recomputes frequencies each iteration. Results are identical. */
#define MSSPE_SELECT_RECOUNT 0     /* re-stream every posting each iteration = main.rs:292-309 */
#define MSSPE_SELECT_INCREMENTAL 1 /* decrement counts of k-mers in newly covered segments */
#define MSSPE_SELECT_AUTO 2        /* whichever is faster for the input: partitioned when few lists span several partitions,
                                      else incremental from 2^21 postings per direction on, else recount */
#define MSSPE_SELECT_PARTITIONED 3 /* every partition runs its own greedy sequence (a k-mer whose postings lie in one partition only
                                      changes counts inside it); the reference's global order is their merge by (frequency,
                                      partition_coverage, word); lists spanning several partitions are checked against the merged
                                      order and, where one would have won, inserted with a roll-back of the partitions it touches */
#define MSSPE_SELECT_BATCHED 0x100 /* OR-able: one launch per phase instead of the persistent cooperative kernel */

/* NtthalOptions, delta_g.rs:18-25 (without the threshold), and Primer3's thal_args. */
typedef struct {
  double mv;        /* -mv   monovalent cations, mM */
  double dv;        /* -dv   divalent cations, mM */
  double dntp;      /* -n    dNTP, mM */
  double dna_conc;  /* -d    oligo concentration, nM */
  double temp_c;    /* -t    temperature for dG, Celsius */
  int32_t max_loop; /* -maxloop, Primer3 default 30 */
  int32_t reserved;
} msspe_thal_cond;

#define MSSPE_THAL_ANY 1      /* ntthal -a ANY   (delta_g.rs:95) */
#define MSSPE_THAL_END1 2     /* ntthal -a END1 */
#define MSSPE_THAL_HAIRPIN 4  /* ntthal -a HAIRPIN */

/* Result of one thermodynamic alignment; the numbers ntthal prints (delta_g.rs:27-59 reads dG). */
typedef struct {
  double ds;            /* dS incl. salt correction, cal/(K mol) */
  double dh;            /* dH, cal/mol */
  double dg;            /* dG at cond.temp_c, cal/mol */
  double tm;            /* melting temperature, Celsius */
  int32_t no_structure; /* 1 = no structure: ntthal prints nothing for such a dimer (a stderr message for a hairpin) */
  int32_t n_bp;         /* paired bases counted by the traceback */
} msspe_thal_out;

/* Nearest-neighbour tables in the file order of a primer3_config directory (delta_g.rs:90,107).
 * "inf" entries are INFINITY.  loops_*: 30 rows of (interior, bulge, hairpin). */
typedef struct {
  double stack_ds[256], stack_dh[256];
  double stackmm_ds[256], stackmm_dh[256];
  double dangle_ds[128], dangle_dh[128];
  double loops_ds[90], loops_dh[90];
  double tstack_ds[256], tstack_dh[256];   /* tstack_tm_inf.ds + tstack.dh */
  double tstack2_ds[256], tstack2_dh[256];
  int32_t n_triloop_ds;  char triloop_ds_seq[32][8];   double triloop_ds[32];
  int32_t n_triloop_dh;  char triloop_dh_seq[32][8];   double triloop_dh[32];
  int32_t n_tetraloop_ds; char tetraloop_ds_seq[128][8]; double tetraloop_ds[128];
  int32_t n_tetraloop_dh; char tetraloop_dh_seq[128][8]; double tetraloop_dh[128];
} msspe_thal_raw_params;

/* A cross-dimer pair the host must look at: pair = a * n + b (row-major, a-major like delta_g.rs:64-78). */
typedef struct {
  uint64_t pair;
  double dg;           /* raw FP64 dG; the host applies ntthal's "%g" -> f32 round trip */
} msspe_dimer_edge;

/* Per-stage device times of the last call of each kind, milliseconds (CUDA events on the ctx stream). */
typedef struct {
  float h2d_ms;        /* genome upload */
  float encode_ms;     /* K1 window slice + 2-bit encode + per-window dedup */
  float index_ms;      /* K2 radix sort + CSR inverted index + forward index */
  float select_ms[2];  /* K3 greedy loop, per direction */
  float thermo_ms;     /* K4 oligotm + K5 self-dimers + K6 hairpin */
  float dimer_ms;      /* K5 all-pairs dimer */
  uint64_t select_evals[2];        /* sum over iterations of live (segment,k-mer) records = main.rs:302-307 executions */
  uint64_t select_postings_read[2];/* physical postings streamed by the count kernel */
  uint32_t select_iterations[2];
  uint32_t kernel_launches;        /* launches of this library's kernels since the last reset */
  float count_kernel_ms[2];        /* total device time inside the count/score kernel (when profiling is on) */
  uint32_t count_kernel_launches[2];
} msspe_timing;

/* ---- lifecycle ---------------------------------------------------------------------------- */
int msspe_abi_version(void);
/* Replaces: process start-up of main.rs:596-629 for this path.  Panics of main.rs:201-203 / step_by(0)
 * become MSSPE_ERR_INVALID. */
int msspe_create(const msspe_config* cfg, msspe_ctx** out);
void msspe_destroy(msspe_ctx* ctx);
const char* msspe_last_error(const msspe_ctx* ctx); /* ctx may be NULL: error of the last failed create */
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream) instead of the ctx's own stream. */
int msspe_set_stream(msspe_ctx* ctx, void* cuda_stream);
/* Cold start (no reference counterpart: the reference has no device): map `bytes` of device memory into the context's
 * stream-ordered pool in the BACKGROUND (returns at once; the first index build waits for it).  A process that knows the
 * size of its input calls this right after msspe_create so that the driver's first-touch cost of the pool (2.4 s at the
 * complete configs[4] input) overlaps the FASTA parse (main.rs:108-122) instead of the first build.  The amount is capped
 * at 80 % of the free device memory; 0 is a no-op. */
int msspe_reserve_pool(msspe_ctx* ctx, uint64_t bytes);
int msspe_synchronize(msspe_ctx* ctx);
int msspe_get_timing(msspe_ctx* ctx, msspe_timing* out);
int msspe_reset_timing(msspe_ctx* ctx);
int msspe_set_profiling(msspe_ctx* ctx, int on); /* per-launch CUDA-event timing of this library's kernels */
/* Per kernel class since the last msspe_reset_timing: launches and algorithmic bytes (DESIGN.md states the per-unit figure
 * of each class) always; device time (CUDA events around every launch, on the launching stream) while profiling is on. */
typedef struct {
  char name[56];
  float ms;
  uint32_t launches;
  uint64_t alg_bytes;
} msspe_kernel_prof;
int msspe_get_kernel_profile(msspe_ctx* ctx, msspe_kernel_prof* out, uint32_t capacity, uint32_t* n);

/* ---- (a) segments and inverted index ------------------------------------------------------- */
/* Replaces get_segment_manager, main.rs:196-235 (input side).  `bases` = the records' sequences
 * concatenated (already upper-cased with U->T as to_records does, main.rs:108-122; the engine also
 * accepts lower case and U), offsets[n+1] delimit them.  Host memory; copied to the device. */
int msspe_load_genomes(msspe_ctx* ctx, const uint8_t* bases, const uint64_t* offsets, uint32_t n_records);
/* Same, but `d_bases` is already resident in device memory on ctx's device (offsets stay host). */
int msspe_load_genomes_device(msspe_ctx* ctx, const uint8_t* d_bases, const uint64_t* offsets, uint32_t n_records);
/* FASTA ingest.  Replaces to_records, main.rs:108-122 (seq_io reader: id = header up to the first space, sequence
 * lines joined, upper-cased, U -> T) with a multi-threaded parse into one pinned buffer; n_threads 0 = all cores.
 * msspe_fasta_open only parses (no device needed); msspe_load_fasta parses AND loads: the host-to-device copy of each
 * 32 MB chunk is issued while the worker threads normalise the next one.  A FASTA that does not start with '>' gives
 * MSSPE_ERR_INVALID (seq_io InvalidStart), a missing file MSSPE_ERR_IO.  The handle owns names/offsets/bases. */
typedef struct msspe_fasta msspe_fasta;
int msspe_fasta_open(const char* path, uint32_t n_threads, msspe_fasta** out, char* err, size_t err_len);
void msspe_fasta_close(msspe_fasta* f);
uint32_t msspe_fasta_records(const msspe_fasta* f);
const char* msspe_fasta_name(const msspe_fasta* f, uint32_t i);  /* SequenceRecord.name */
const uint8_t* msspe_fasta_bases(const msspe_fasta* f);         /* SequenceRecord.sequence of all records, concatenated */
const uint64_t* msspe_fasta_offsets(const msspe_fasta* f);      /* [records + 1] */
int msspe_load_fasta(msspe_ctx* ctx, const char* path, uint32_t n_threads, msspe_fasta** out);
/* K1 + K2: slice windows, encode, dedup per (segment,direction), sort, build CSR postings + forward
 * index for both directions.  Replaces main.rs:205-232 and make_kmer_segments_windows_mapping :237-255. */
int msspe_build_index(msspe_ctx* ctx);
/* n_segments = SegmentManager.segments.len(); max_partition = main.rs:694-699 (0 if no segment). */
int msspe_segment_info(msspe_ctx* ctx, uint64_t* n_segments, uint32_t* max_partition, uint32_t* slots_per_window);
/* Segment.kmers[dir] as a dense [segment][slot] table, MSSPE_NO_KMER for invalid/duplicate slots. */
int msspe_get_segment_kmers(msspe_ctx* ctx, uint8_t dir, uint64_t* codes, uint64_t capacity);
/* The inverted index of main.rs:237-255: n_codes distinct words (ascending), offsets[n_codes+1], postings
 * (ascending segment index inside each list).  Pass NULL buffers to query sizes only. */
int msspe_get_index(msspe_ctx* ctx, uint8_t dir, uint64_t* n_codes, uint64_t* n_postings,
                    uint64_t* codes, uint64_t* offsets, uint32_t* postings);

/* ---- (b)(c) greedy selection ---------------------------------------------------------------- */
/* Replaces find_candidates_kmers, main.rs:331-406, for one direction.  out has capacity max_iterations. */
int msspe_select(msspe_ctx* ctx, uint8_t dir, uint32_t max_iterations, uint32_t max_mismatch_segments,
                 uint32_t mode, msspe_candidate* out, uint32_t* n_out);
/* Both directions (main.rs:709-714) overlapped on the device. */
int msspe_select_both(msspe_ctx* ctx, uint32_t max_iterations, uint32_t max_mismatch_segments, uint32_t mode,
                      msspe_candidate* out_fwd, uint32_t* n_fwd, msspe_candidate* out_rev, uint32_t* n_rev);
/* Device side of print_coverage_report, main.rs:518-537: covered[g] = 1 iff segment g holds a selected
 * fwd k-mer in its head set or a selected rev k-mer in its tail set.  partition_no[g] as main.rs:227. */
int msspe_coverage(msspe_ctx* ctx, const uint64_t* fwd_codes, uint32_t n_fwd, const uint64_t* rev_codes,
                   uint32_t n_rev, uint8_t* covered, uint16_t* partition_no, uint32_t* record_of_segment,
                   uint64_t capacity);

/* print_coverage_report's aggregation (main.rs:518-574) as device reductions, so that a 12-million-segment job copies
 * back O(records + partitions) instead of O(segments): rec_covered/rec_total[r] = covered / all segments of record r
 * (SequenceRecord order of msspe_load_genomes), part_covered/part_total[p] likewise per Segment.partition_no,
 * *n_covered = covered segments.  Capacities: n_records >= loaded records, n_part > max_partition. */
int msspe_coverage_summary(msspe_ctx* ctx, const uint64_t* fwd_codes, uint32_t n_fwd, const uint64_t* rev_codes,
                           uint32_t n_rev, uint32_t* rec_covered, uint32_t* rec_total, uint32_t n_records,
                           uint32_t* part_covered, uint32_t* part_total, uint32_t n_part, uint64_t* n_covered);

/* ---- conflict graph + greedy vertex cover ----------------------------------------------------------------------
 * Replaces main.rs:754-798 (and graphdb.rs as its container).  codes[n] = the DISTINCT primer words (the reference
 * keys its graph by word: a word selected in both directions is one node), edges (edge_a[e], edge_b[e]) = node
 * indices of every conflict edge, i.e. every stored edge with dG below the threshold (a == b allowed: self
 * conflict).  deleted[v] = 1 for the primers the loop removes: repeatedly the live primer with the most live
 * neighbours, ties -> lexicographically greatest word.  Device: n x n adjacency bit matrix, popcount degrees, one
 * persistent kernel for the whole loop.  n <= 65536. */
int msspe_vertex_cover(msspe_ctx* ctx, const uint64_t* codes, uint32_t n, const uint32_t* edge_a, const uint32_t* edge_b,
                       uint64_t n_edges, uint8_t* deleted, uint32_t* n_deleted);

/* ---- (e) genome-sharded selection: per-rank primitives ----------------------------------------------------
 * One process per GPU holds a contiguous block of the records (hence of the global segment order).  The loop of
 * main.rs:331-406 is then driven above the ABI (msspe_b200/distributed.py): local recount -> all-reduce(sum) of the
 * per-k-mer counts -> identical arg-max on every rank -> tie scores from the per-rank first-seen partition
 * positions (rank order = segment order, main.rs:268-281) -> every rank marks its own postings of the winner.
 * These entry points are the device work of one iteration; ids are LOCAL code ids (index into msspe_get_index's
 * codes[]), MSSPE_NO_LOCAL_ID = the k-mer does not occur on this rank. */
#define MSSPE_NO_LOCAL_ID 0xFFFFFFFFu
int msspe_shard_begin(msspe_ctx* ctx, uint8_t dir);
/* device pointers (valid until the next build): ascending codes[n_codes] (u64) and freq[n_codes] (u32) */
int msspe_shard_buffers(msspe_ctx* ctx, uint8_t dir, const uint64_t** d_codes, const uint32_t** d_freq, uint64_t* n_codes);
/* main.rs:292-309 on this rank's segments: freq[] <- live segments per local k-mer; *live = their sum */
int msspe_shard_count(msspe_ctx* ctx, uint8_t dir, uint64_t* live);
/* first_pos[t * n_part + p] = offset of the first live posting of local_ids[t] in partition p, 0xFFFFFFFF if none */
int msspe_shard_firstpos(msspe_ctx* ctx, uint8_t dir, const uint32_t* local_ids, uint32_t n, uint32_t n_part, uint32_t* first_pos);
/* main.rs:371-378 for the winner on this rank: mark all its postings covered; part_flags[p] = 1 for every partition
 * any of its postings lies in (n_part entries) */
int msspe_shard_apply(msspe_ctx* ctx, uint8_t dir, uint32_t local_id, uint32_t n_part, uint8_t* part_flags);

/* ---- (d) thermodynamics --------------------------------------------------------------------- */
/* Replaces `-path <cwd>/primer3_config/` (delta_g.rs:90).  Default = tables embedded at build time. */
int msspe_thal_params_default(msspe_thal_raw_params* out);
int msspe_thal_params_from_dir(const char* dir, msspe_thal_raw_params* out, char* err, size_t err_len);
int msspe_set_thal_params(msspe_ctx* ctx, const msspe_thal_raw_params* p);
/* Host-only diagnostic (no device, no ctx): one table as the kernels index it -- the expansion of `p` to the 5-symbol
 * (A,C,G,T,N) arrays with Primer3's load rules -- under the name libprimer3 2.6.1's thal.c gives that array
 * ("stackEntropies", "tstack2Enthalpies", "dangleEntropies3", "hairpinLoopEntropies", "atpS", ...; for
 * "default{Tri,Tetra}loop{Entropies,Enthalpies}" `out` receives (key, value) pairs, key = the loop's base-5 digits).
 * Returns the number of doubles written or a negative MSSPE_ERR_*.  What it is for: tests compare these arrays with the
 * ones compiled into the Primer3 executables the reference spawns (primer.rs:125-140, delta_g.rs:90-108). */
int msspe_thal_expanded_table(const msspe_thal_raw_params* p, const char* name, double* out, uint32_t cap);
/* Replaces check_primers, primer.rs:143-166 (one primer3_core run): per primer the five numbers
 * parse_primer3_output reads (primer.rs:67-114), as raw FP64 before Primer3's "%.3f"/"%.2f" printing:
 * tm = oligotm at Primer3 defaults, gc = percent, self_any/self_end/hairpin = max(0, thal Tm). */
int msspe_primer_thermo(msspe_ctx* ctx, const uint64_t* codes, uint32_t n, uint32_t oligo_len, double* tm,
                        double* gc, double* self_any, double* self_end, double* hairpin);
/* One row of get_kmer_stats (main.rs:408-455): the five numbers AFTER Primer3's text round trip ("%.3f" for TM and
 * GC_PERCENT, "%.2f" for the *_TH values, each parsed back as f32 like primer.rs:67-114), the per-direction Tm mean
 * and standard deviation (main.rs:462-467; std-dev 0.1.0 = sample deviation), tm_in_threshold (:469-471), is_run
 * (:478-490) and the verdict of filter_kmers (:492-516). */
typedef struct {
  uint64_t code;
  float tm, gc_percent, self_any_th, self_end_th, hairpin_th, mean, std;
  uint8_t tm_ok, runs, keep, reserved;
} msspe_kmer_stat;

/* The ProgramConfig / PrimerConfig fields filter_kmers reads (config.rs:150-177). */
typedef struct {
  float min_tm, max_tm, max_self_dimer_any_tm, max_self_dimer_end_tm, max_hairpin_tm, tm_stddev;
  uint8_t check_self_dimers, check_hairpin, disable_tm_stddev, disable_min_max_tm;
} msspe_filter_cfg;

/* Replaces get_kmer_stats + filter_kmers for the candidates of ONE direction, in selection order: the thermodynamics
 * run on the device (msspe_primer_thermo), the printing/rounding, the sequential f32 mean / deviation and the strict
 * f32 comparisons on the host exactly as the reference does them. */
int msspe_kmer_stats(msspe_ctx* ctx, const uint64_t* codes, uint32_t n, uint32_t oligo_len, const msspe_filter_cfg* cfg,
                     msspe_kmer_stat* out);

/* The two get_kmer_stats calls of main.rs:723-724 at once: one device batch for the candidates of both directions,
 * per-direction statistics and verdicts identical to two msspe_kmer_stats calls. */
int msspe_kmer_stats_both(msspe_ctx* ctx, const uint64_t* fwd_codes, uint32_t n_fwd, const uint64_t* rev_codes, uint32_t n_rev,
                          uint32_t oligo_len, const msspe_filter_cfg* cfg, msspe_kmer_stat* out_fwd, msspe_kmer_stat* out_rev);

/* Arbitrary pair list through thal (type = MSSPE_THAL_*); a[i], b[i] are 2-bit codes of length oligo_len.
 * For HAIRPIN b is ignored.  One ntthal invocation per pair in the reference (delta_g.rs:93-113). */
int msspe_thal_pairs(msspe_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t n_pairs, uint32_t oligo_len,
                     int32_t type, const msspe_thal_cond* cond, msspe_thal_out* out);
/* msspe_thal_pairs for the dimer types (ANY, END1) plus the base pairs of every structure, which is what ntthal's
 * SEQ/STR drawing (the four lines the parser of delta_g.rs:55 skips) is made from:
 * pairing[p * MSSPE_MAX_OLIGO + i] = 1-based position, in the REVERSED second oligo (3'->5', as ntthal draws it),
 * of the base that base i (0-based) of the first oligo pairs with; 0 = unpaired. */
int msspe_thal_pairs_aligned(msspe_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t n_pairs, uint32_t oligo_len,
                             int32_t type, const msspe_thal_cond* cond, msspe_thal_out* out, uint8_t* pairing);
/* Replaces run_ntthal, delta_g.rs:83-153, for rows [row_begin,row_end) of the n x n ordered-pair matrix
 * (all pairs incl. self, a-major, delta_g.rs:64-78).  Emits every pair with dG < dg_limit (caller passes
 * threshold + margin and finishes the "%g"/f32 comparison, delta_g.rs:33-36) and every pair without a
 * structure (needed to emulate the 5-line parser, delta_g.rs:31-56).  Lists are sorted by pair index. */
int msspe_cross_dimer(msspe_ctx* ctx, const uint64_t* codes, uint32_t n, uint32_t oligo_len,
                      const msspe_thal_cond* cond, uint32_t row_begin, uint32_t row_end, double dg_limit,
                      msspe_dimer_edge* edges, uint64_t edge_capacity, uint64_t* n_edges,
                      uint64_t* nostruct_pairs, uint64_t nostruct_capacity, uint64_t* n_nostruct);

/* msspe_cross_dimer with the two compacted lists left on the DEVICE (unsorted; ctx-owned buffers, valid until the next
 * cross-dimer call on this ctx): the per-rank half of the row-tiled N x N matrix of SURVEY 8(e) -- the ranks exchange the
 * counts and merge the lists with one all_gather on device buffers instead of a host round trip per rank. */
int msspe_cross_dimer_device(msspe_ctx* ctx, const uint64_t* codes, uint32_t n, uint32_t oligo_len,
                             const msspe_thal_cond* cond, uint32_t row_begin, uint32_t row_end, double dg_limit,
                             uint64_t edge_capacity, uint64_t nostruct_capacity, const msspe_dimer_edge** d_edges,
                             uint64_t* n_edges, const uint64_t** d_nostruct, uint64_t* n_nostruct);

/* ---- (e) ONE design job over several GPUs: partitions (alignment columns) sharded over the ranks -------------------------
 * One process per GPU.  Every rank's ctx is loaded with the columns of a contiguous range of partitions of EVERY genome
 * (partition p of a genome = columns [p * overlap_size, p * overlap_size + window_size), main.rs:173-181; rank ranges in
 * rank order, together all partitions) and builds its own index (no communication).  msspe_select_both_dist then
 * replaces find_candidates_kmers (main.rs:331-406, both directions) for the WHOLE input: collective over the ranks,
 * every rank receives the complete candidate lists, bit-identical to msspe_select_both on one GPU holding all
 * columns.  Communication: NCCL (bound with dlopen from the host process), an all-gather of the ranks' not-yet-final
 * per-partition winners and one all-reduce per round of the greedy loop's rounds -- nothing per iteration.
 * msspe_dist_unique_id (rank 0; broadcast the 128 bytes with any out-of-band channel) + msspe_dist_init (all ranks,
 * collective) create the communicator.  MSSPE_ERR_CAPACITY for the documented limits of this path. */
int msspe_dist_unique_id(uint8_t* out128);
int msspe_dist_init(msspe_ctx* ctx, const uint8_t* id128, int rank, int world);
int msspe_select_both_dist(msspe_ctx* ctx, uint32_t max_iterations, uint32_t max_mismatch_segments,
                           msspe_candidate* out_fwd, uint32_t* n_fwd, msspe_candidate* out_rev, uint32_t* n_rev);

#ifdef __cplusplus
}
#endif
#endif /* OD_MSSPE_B200_H */
