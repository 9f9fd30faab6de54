/* oracle/kuq_oracle.h — CPU restatement of KrakenUniq's per-read classification hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker.  Nothing under krakenuniq_b200/ links or calls it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors, so this restatement is pinned
 * against the reference itself, compiled unmodified by oracle/build_ref.sh into oracle/_ref/:
 *   - function by function through oracle/_ref/libkuref.so (tests/test_oracle_vs_reference.py), and
 *   - end to end against `oracle/_ref/classify` outputs committed under tests/golden/ (tests/test_golden.py).
 *
 * Every function cites the reference file:line (relative to /root/reference/) it follows.
 */
#ifndef KUQ_ORACLE_H
#define KUQ_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define KUQO_AMBIG 0xFFFFFFFFu /* per-k-mer code for an ambiguous window (classify.cpp:920-923 "A:") */

/* ---- bit arithmetic ------------------------------------------------------------------------------------ */
uint64_t kuqo_fmix64(uint64_t key);                                   /* hyperloglogplus.cpp:830-838 */
uint64_t kuqo_revcomp(uint64_t kmer, unsigned n);                     /* krakendb.cpp:218-225 */
uint64_t kuqo_canonical(uint64_t kmer, unsigned n);                   /* krakendb.cpp:238-246 */
uint64_t kuqo_bin_key(uint64_t kmer, unsigned k, unsigned nt, int idx_type); /* krakendb.cpp:200-215 */
/* KmerScanner::next_kmer / ambig_kmer, krakenutil.cpp:239-282, under the guard of classify.cpp:913.
 * Returns the number of windows; kmers/ambig need len-k+1 (+1 for a trailing-'\r' read) entries. */
uint32_t kuqo_scan(const char *seq, size_t len, unsigned k, uint64_t *kmers, uint8_t *ambig);

/* ---- database view over database.kdb / database.idx images ------------------------------------------- */
typedef struct {
  const uint8_t *pairs;    /* first record (after the Jellyfish header) */
  uint64_t key_ct;
  unsigned k, key_bits, key_len, pair_sz;
  const uint64_t *offsets; /* 4^nt + 1 cumulative record offsets */
  unsigned nt;
  int idx_type;            /* 1 = KRAKIDX, 2 = KRAKIX2 */
} kuqo_db;
/* krakendb.cpp:60-78,177,534-544.  Returns 0, or a negative code for a malformed image. */
int kuqo_db_open(kuqo_db *db, const void *kdb_image, uint64_t kdb_bytes, const void *idx_image, uint64_t idx_bytes);
/* Stateless exact-match lookup == KrakenDB::kmer_query (krakendb.cpp:250-325).  Returns 1 and *taxon when the
 * key is present (the stored value may be 0), 0 on a miss. */
int kuqo_kmer_query(const kuqo_db *db, uint64_t canon, uint32_t *taxon);

/* ---- taxonomy: Parent_map as built by TaxonomyDB::getParentMap (taxdb.hpp:383-398) ---------------------- */
typedef struct kuqo_parent_map kuqo_parent_map;
kuqo_parent_map *kuqo_parent_map_new(const uint32_t *taxid, const uint32_t *parent, uint32_t n);
void kuqo_parent_map_free(kuqo_parent_map *pm);
uint32_t kuqo_lca(const kuqo_parent_map *pm, uint32_t a, uint32_t b);                 /* krakenutil.cpp:90-118 */
uint32_t kuqo_resolve_tree(const kuqo_parent_map *pm, const uint32_t *hit_taxa,
                           const uint32_t *hit_counts, uint32_t n_hits);              /* krakenutil.cpp:149-200 */

/* ---- HyperLogLogPlusMinus<uint64_t>, default ctor p=12 sparse (readcounts.hpp:40) ---------------------- */
typedef struct kuqo_hll kuqo_hll;
kuqo_hll *kuqo_hll_new(void);
void kuqo_hll_free(kuqo_hll *h);
void kuqo_hll_insert(kuqo_hll *h, uint64_t item);            /* hyperloglogplus.cpp:485-523 */
void kuqo_hll_merge(kuqo_hll *dst, const kuqo_hll *src);     /* hyperloglogplus.cpp:586-665 */
uint64_t kuqo_hll_cardinality(const kuqo_hll *h);            /* ertlCardinality, :722-753 */
int kuqo_hll_is_sparse(const kuqo_hll *h);
uint64_t kuqo_hll_n_observed(const kuqo_hll *h);
uint32_t kuqo_hll_sparse_size(const kuqo_hll *h);
/* Dense p=12 registers implied by the sketch (its M when dense; the fold of its sparse codes otherwise,
 * addToRegisters :559-577).  regs must hold 4096 bytes. */
void kuqo_hll_registers(const kuqo_hll *h, uint8_t *regs);
/* Ertl estimate straight from 4096 dense registers + n_observed (:730-752): used to check the product's host
 * estimator independently of any sketch object. */
uint64_t kuqo_ertl_dense(const uint8_t *regs, uint64_t n_observed);
uint32_t kuqo_encode_hash32(uint64_t hash);                  /* encodeHashIn32Bit(h,25,12), :181-204 */

/* ---- one read (classify_sequence, classify.cpp:897-1012, non-quick, single DB) ------------------------- */
/* codes_out[i] = taxon of window i (0 = miss) or KUQO_AMBIG; returns the call. *n_windows = #windows. */
uint32_t kuqo_classify_read(const kuqo_db *db, const kuqo_parent_map *pm, const char *seq, size_t len,
                            uint32_t *codes_out, uint32_t *n_windows);
/* hitlist_string, classify.cpp:826-861 ("0:0" when there are no windows, :994-995).  Returns strlen. */
size_t kuqo_hitlist_string(const uint32_t *codes, uint32_t n, char *buf, size_t cap);

/* ---- whole run (process_file work-unit loop, classify.cpp:487-564; chunked rule :663-791) --------------- */
typedef struct kuqo_run kuqo_run;
/* mode 0 = preload/mmap (per-work-unit sketches merged into the global map, :525,542-544);
 * mode 1 = chunked -x (every k-mer inserted straight into the global sketch, :719). */
kuqo_run *kuqo_run_new(const kuqo_db *db, const kuqo_parent_map *pm, uint64_t work_unit_size, int mode);
void kuqo_run_free(kuqo_run *r);
/* Classify n_reads reads (concatenated bases, offsets[n_reads+1]); appends to the run's global counts.
 * calls_out[n_reads]; codes_out (may be NULL) receives the per-window codes back to back, code_offsets_out
 * [n_reads+1] their prefix sums.  Work units continue across calls exactly as one input file would. */
int kuqo_run_classify(kuqo_run *r, const char *bases, const uint64_t *offsets, uint32_t n_reads,
                      uint32_t *calls_out, uint32_t *codes_out, uint64_t *code_offsets_out);
/* flush the last partial work unit into the global map (end of file) */
void kuqo_run_finish(kuqo_run *r);
uint32_t kuqo_run_n_taxa(const kuqo_run *r);
/* per-taxon results sorted by taxid: reads, kmers, Ertl estimate (uniqueKmerCount, readcounts.hpp:121-124),
 * sparse flag, and (regs != NULL) 4096 implied dense registers per taxon. */
void kuqo_run_counts(const kuqo_run *r, uint32_t *taxid, uint64_t *n_reads, uint64_t *n_kmers,
                     uint64_t *unique_est, uint8_t *is_sparse, uint8_t *regs);

/* further databases, tried in order after the first one for every k-mer (classify.cpp:928-936); same k required */
int kuqo_run_add_db(kuqo_run *r, const kuqo_db *db);
/* classifyExact (EXACT_COUNTING, classify.cpp:46-49; readcounts.hpp:71-81,127-130): unique counts become exact set
 * sizes (per taxon) / union sizes (per clade).  Call before the first kuqo_run_classify. */
void kuqo_run_set_exact(kuqo_run *r, int on);
/* quick mode (-q -m min_hits; classify.cpp:943-944,963-964 preloaded, :701-702,737-738 with -x); the codes of a read
 * then cover the windows actually visited; "Q:hits" = min(hits among them, min_hits).  0 = off. */
void kuqo_run_set_quick(kuqo_run *r, uint32_t min_hits);
/* clade roll-up (TaxReport ctor, taxdb.hpp:956-973): sums the listed taxa's ReadCounts; returns unique estimate */
uint64_t kuqo_run_clade(const kuqo_run *r, const uint32_t *taxa, uint32_t n, uint64_t *n_reads, uint64_t *n_kmers);

/* ---- database build ------------------------------------------------------------------ */
/* db_sort (db_sort.cpp:41-116, make_index krakendb.cpp:118-148): unsorted Jellyfish-style image → database.kdb
 * image (jdb_bytes bytes) + KRAKIX2 index image (8 + 8 * (4^nt + 1) bytes).  zero_vals = db_sort -z. */
int kuqo_db_sort(const void *jdb_image, uint64_t jdb_bytes, uint32_t nt, int zero_vals, void *kdb_out, void *idx_out);
/* set_lcas for one library sequence (set_lcas.cpp:429-476): value = lca(taxid, value) for every k-mer of the
 * sequence that the database holds; `db` must be open on a WRITABLE image.  Returns #k-mers not in the database. */
uint64_t kuqo_set_lcas_sequence(kuqo_db *db, const kuqo_parent_map *pm, const char *seq, size_t len, uint32_t taxid);
/* flags: 1 = set_lcas -T (values 32630 'synthetic construct' / 81077 'artificial sequences' stick; a sequence with
 * such a taxid overwrites instead of taking the LCA, :462-474), 2 = -R (the value is reset to 0, :458-459) */
uint64_t kuqo_set_lcas_sequence_flags(kuqo_db *db, const kuqo_parent_map *pm, const char *seq, size_t len, uint32_t taxid,
                                      uint32_t flags);

#ifdef __cplusplus
}
#endif
#endif
