/* oracle/kuq_oracle.c — CPU restatement of KrakenUniq's per-read classification hot path (plain C).
 *
 * TEST INFRASTRUCTURE ONLY — see kuq_oracle.h.  Parity status: PINNED against the compiled reference
 * (oracle/_ref, built by oracle/build_ref.sh) by tests/test_oracle_vs_reference.py and tests/test_golden.py.
 *
 * Citations are file:line under /root/reference/src/.  Where the reference leans on std::unordered_map/set the
 * restatement uses small open-addressing tables; every result the reference derives from those containers is
 * iteration-order independent, which the pin tests confirm.
 */
#include "kuq_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================================================= */
/* bit arithmetic                                                                                          */
/* ======================================================================================================= */

/* murmurhash3_finalizer, hyperloglogplus.cpp:830-838 (note the +1 so that key 0 does not hash to 0) */
uint64_t kuqo_fmix64(uint64_t key) {
  key += 1;
  key ^= key >> 33;
  key *= 0xff51afd7ed558ccdULL;
  key ^= key >> 33;
  key *= 0xc4ceb9fe1a85ec53ULL;
  key ^= key >> 33;
  return key;
}

/* KrakenDB::reverse_complement(kmer, n), krakendb.cpp:218-225: reverse the 2-bit groups of the whole word,
 * complement, shift the n significant groups back down. */
uint64_t kuqo_revcomp(uint64_t kmer, unsigned n) {
  kmer = ((kmer >> 2) & 0x3333333333333333ULL) | ((kmer & 0x3333333333333333ULL) << 2);
  kmer = ((kmer >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((kmer & 0x0F0F0F0F0F0F0F0FULL) << 4);
  kmer = ((kmer >> 8) & 0x00FF00FF00FF00FFULL) | ((kmer & 0x00FF00FF00FF00FFULL) << 8);
  kmer = ((kmer >> 16) & 0x0000FFFF0000FFFFULL) | ((kmer & 0x0000FFFF0000FFFFULL) << 16);
  kmer = (kmer >> 32) | (kmer << 32);
  return (~kmer) >> (64 - 2 * n);
}

/* KrakenDB::canonical_representation, krakendb.cpp:238-246 */
uint64_t kuqo_canonical(uint64_t kmer, unsigned n) {
  uint64_t rc = kuqo_revcomp(kmer, n);
  return kmer < rc ? kmer : rc;
}

/* KrakenDB::bin_key(kmer), krakendb.cpp:200-215.  idx_type 1 (KRAKIDX) → no scrambling; 2 (KRAKIX2) → XOR with
 * INDEX2_XOR_MASK (:45) truncated to 2*nt bits.  `1 << (nt*2)` is an int shift in the reference (:204), so
 * nt <= 15 is the supported range; we do the same arithmetic in 64 bits for nt <= 15. */
uint64_t kuqo_bin_key(uint64_t kmer, unsigned k, unsigned nt, int idx_type) {
  uint64_t xor_mask = idx_type == 1 ? 0 : 0xe37e28c4271b5a2dULL;
  uint64_t mask = (1ull << (nt * 2)) - 1;
  xor_mask &= mask;
  uint64_t min_bin_key = ~0ull;
  for (unsigned i = 0; i < k - nt + 1; i++) {
    uint64_t t = xor_mask ^ kuqo_canonical(kmer & mask, nt);
    if (t < min_bin_key) min_bin_key = t;
    kmer >>= 2;
  }
  return min_bin_key;
}

/* KmerScanner (krakenutil.cpp:205-282) driven as classify_sequence does (classify.cpp:913-918).
 * std::string guarantees seq[len] == '\0', which the scanner reads (as an ambiguous base) when the sequence
 * ends in skipped characters ('\r' of a CRLF file: SURVEY App. A9).  Reads further past the end are undefined
 * behaviour in the reference; we define them as '\0' too. */
uint32_t kuqo_scan(const char *seq, size_t len, unsigned k, uint64_t *kmers, uint8_t *ambig_out) {
  if (len < k) return 0; /* classify.cpp:913 */
  const uint64_t kmer_mask = ~0ull >> (64 - 2 * k);       /* krakenutil.cpp:233-234 */
  const uint32_t mini_kmer_mask = ~0u >> (32 - k);        /* :235-236 */
  size_t curr_pos = 0;
  const size_t pos2 = len;
  int64_t loaded_nt = 0;
  uint64_t kmer = 0;
  uint32_t ambig = 0;
  uint32_t n = 0;
  for (;;) {                                              /* one iteration == one next_kmer() call, :239-278 */
    int skip_pos = 0;
    if (curr_pos >= pos2) break;
    if (loaded_nt) loaded_nt--;
    while (loaded_nt < (int64_t)k) {
      if (skip_pos) {
        skip_pos = 0;
      } else {
        loaded_nt++;
        kmer <<= 2;
        ambig <<= 1;
      }
      char c = curr_pos < len ? seq[curr_pos] : '\0';
      curr_pos++;
      switch (c) {
        case 'A': case 'a': break;
        case 'C': case 'c': kmer |= 1; break;
        case 'G': case 'g': kmer |= 2; break;
        case 'T': case 't': kmer |= 3; break;
        case '\n': case '\r':
          --loaded_nt;
          skip_pos = 1;
          continue;
        default: ambig |= 1; break;
      }
      kmer &= kmer_mask;
      ambig &= mini_kmer_mask;
    }
    kmers[n] = kmer;
    ambig_out[n] = ambig ? 1 : 0;                          /* ambig_kmer(), :280-282 */
    n++;
  }
  return n;
}

/* ======================================================================================================= */
/* database                                                                                                */
/* ======================================================================================================= */

int kuqo_db_open(kuqo_db *db, const void *kdb_image, uint64_t kdb_bytes, const void *idx_image, uint64_t idx_bytes) {
  const uint8_t *p = (const uint8_t *)kdb_image;
  if (!p || kdb_bytes < 56 || memcmp(p, "JFLISTDN", 8) != 0) return -1;       /* krakendb.cpp:32,67 */
  uint64_t key_bits, val_len, key_ct;
  memcpy(&key_bits, p + 8, 8);                                                  /* :70 */
  memcpy(&val_len, p + 16, 8);                                                  /* :71 */
  memcpy(&key_ct, p + 48, 8);                                                   /* :72 */
  if (val_len != 4) return -2;                                                  /* :73-74 */
  db->key_bits = (unsigned)key_bits;
  db->k = (unsigned)(key_bits / 2);                                             /* :75 */
  db->key_len = (unsigned)(key_bits / 8 + !!(key_bits % 8));                    /* :76 */
  db->pair_sz = db->key_len + 4;
  db->key_ct = key_ct;
  uint64_t header = 72 + 2 * (4 + 8 * key_bits);                                /* :177 */
  if (kdb_bytes < header + key_ct * db->pair_sz) return -3;
  db->pairs = p + header;
  const uint8_t *q = (const uint8_t *)idx_image;
  if (!q || idx_bytes < 8) return -4;
  if (memcmp(q, "KRAKIDX", 7) == 0) db->idx_type = 1;                           /* :36,536 */
  else if (memcmp(q, "KRAKIX2", 7) == 0) db->idx_type = 2;                      /* :41,538-540 */
  else return -5;
  db->nt = q[7];                                                                /* :543 */
  if (db->nt < 1 || db->nt > 15) return -6;
  if (idx_bytes < 8 + 8 * ((1ull << (2 * db->nt)) + 1)) return -7;
  db->offsets = (const uint64_t *)(q + 8);
  return 0;
}

static inline uint64_t db_key_at(const kuqo_db *db, uint64_t pos) {
  uint64_t key = 0;
  memcpy(&key, db->pairs + db->pair_sz * pos, db->key_len);                     /* krakendb.cpp:283 */
  if (db->key_bits < 64) key &= (1ull << db->key_bits) - 1;                     /* :284 */
  return key;
}

/* KrakenDB::kmer_query, krakendb.cpp:250-321, in its stateless form (:322-325).  The cached-range variant the
 * read loop uses returns the same answer for every key, because a key is stored only in the bin of its own
 * bin_key (db_sort.cpp:98-104): a stale range can only miss, and a miss re-runs this search (:304-319). */
int kuqo_kmer_query(const kuqo_db *db, uint64_t kmer, uint32_t *taxon) {
  uint64_t b = kuqo_bin_key(kmer, db->k, db->nt, db->idx_type);
  int64_t min = (int64_t)db->offsets[b];
  int64_t max = (int64_t)db->offsets[b + 1] - 1;
  while (min + 15 <= max) {                                                     /* :280 */
    int64_t mid = min + (max - min) / 2;
    uint64_t c = db_key_at(db, (uint64_t)mid);
    if (kmer > c) min = mid + 1;
    else if (kmer < c) max = mid - 1;
    else { memcpy(taxon, db->pairs + db->pair_sz * (uint64_t)mid + db->key_len, 4); return 1; }
  }
  for (int64_t mid = min; mid <= max; mid++) {                                  /* :293-299 */
    if (db_key_at(db, (uint64_t)mid) == kmer) {
      memcpy(taxon, db->pairs + db->pair_sz * (uint64_t)mid + db->key_len, 4);
      return 1;
    }
  }
  return 0;
}

/* ======================================================================================================= */
/* taxonomy                                                                                                */
/* ======================================================================================================= */

struct kuqo_parent_map {
  uint32_t cap;        /* power of two */
  uint32_t *key;       /* taxid */
  uint32_t *val;       /* parent */
  uint8_t *used;
};

static inline uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

kuqo_parent_map *kuqo_parent_map_new(const uint32_t *taxid, const uint32_t *parent, uint32_t n) {
  kuqo_parent_map *pm = (kuqo_parent_map *)calloc(1, sizeof(*pm));
  uint32_t cap = 16;
  while (cap < 2 * (uint64_t)n + 2) cap <<= 1;
  pm->cap = cap;
  pm->key = (uint32_t *)calloc(cap, 4);
  pm->val = (uint32_t *)calloc(cap, 4);
  pm->used = (uint8_t *)calloc(cap, 1);
  for (uint32_t i = 0; i < n; i++) {
    uint32_t s = hash32(taxid[i]) & (cap - 1);
    while (pm->used[s] && pm->key[s] != taxid[i]) s = (s + 1) & (cap - 1);
    pm->used[s] = 1; pm->key[s] = taxid[i]; pm->val[s] = parent[i];             /* later entry wins, like map[k]=v */
  }
  return pm;
}
void kuqo_parent_map_free(kuqo_parent_map *pm) {
  if (!pm) return;
  free(pm->key); free(pm->val); free(pm->used); free(pm);
}
/* parent_map.find(): returns 1 and *parent when present */
static inline int pm_find(const kuqo_parent_map *pm, uint32_t t, uint32_t *parent) {
  uint32_t s = hash32(t) & (pm->cap - 1);
  while (pm->used[s]) {
    if (pm->key[s] == t) { *parent = pm->val[s]; return 1; }
    s = (s + 1) & (pm->cap - 1);
  }
  return 0;
}

/* lca(), krakenutil.cpp:90-118.  (The "No parent for" stderr chatter is not restated.) */
uint32_t kuqo_lca(const kuqo_parent_map *pm, uint32_t a, uint32_t b) {
  if (a == 0 || b == 0) return a ? a : b;
  uint32_t path[4096];
  uint32_t n = 0, p;
  while (a > 1) {
    if (n < 4096) path[n++] = a;
    if (!pm_find(pm, a, &p)) break;
    a = p;
  }
  while (b > 1) {
    for (uint32_t i = 0; i < n; i++) if (path[i] == b) return b;
    if (!pm_find(pm, b, &p)) break;
    b = p;
  }
  return 1;
}

/* resolve_tree(), krakenutil.cpp:149-200.  hit_taxa are the distinct keys of hit_counts (all nonzero,
 * classify.cpp:941-942).  The tie set is a std::set: folded with lca in ascending taxid order (:190-196). */
uint32_t kuqo_resolve_tree(const kuqo_parent_map *pm, const uint32_t *hit_taxa, const uint32_t *hit_counts,
                           uint32_t n_hits) {
  uint32_t max_taxon = 0, max_score = 0;
  uint32_t *ties = (uint32_t *)malloc(4 * (size_t)(n_hits + 1));
  uint32_t n_ties = 0;
  for (uint32_t i = 0; i < n_hits; i++) {
    uint32_t taxon = hit_taxa[i], node = taxon, score = 0, p;
    while (node > 0) {
      for (uint32_t j = 0; j < n_hits; j++) if (hit_taxa[j] == node) { score += hit_counts[j]; break; }
      if (!pm_find(pm, node, &p)) break;                  /* :166-169 */
      if (p == node) break;                               /* :170-172 */
      node = p;
    }
    if (score > max_score) {                              /* :179-183 */
      n_ties = 0; max_score = score; max_taxon = taxon;
    } else if (score == max_score) {                      /* :184-188 */
      if (n_ties == 0) ties[n_ties++] = max_taxon;
      ties[n_ties++] = taxon;
    }
  }
  if (n_ties) {                                           /* :191-197, set iteration = ascending order */
    for (uint32_t i = 1; i < n_ties; i++) {               /* insertion sort + dedup-free (keys are distinct) */
      uint32_t v = ties[i]; uint32_t j = i;
      while (j > 0 && ties[j - 1] > v) { ties[j] = ties[j - 1]; j--; }
      ties[j] = v;
    }
    max_taxon = ties[0];
    for (uint32_t i = 1; i < n_ties; i++) max_taxon = kuqo_lca(pm, max_taxon, ties[i]);
  }
  free(ties);
  return max_taxon;
}

/* ======================================================================================================= */
/* HyperLogLogPlusMinus<uint64_t>(precision = 12, sparse = true)                                           */
/* ======================================================================================================= */
#define HLL_P 12u
#define HLL_M 4096u
#define HLL_PPRIME 25u
#define HLL_MPRIME (1u << 25)

struct kuqo_hll {
  int sparse;
  uint64_t n_observed;
  uint8_t *M;            /* HLL_M registers once dense */
  uint32_t *set;         /* sparse list: open-addressing set of encoded hashes (0 = empty; codes are never 0) */
  uint32_t set_cap, set_n;
};

kuqo_hll *kuqo_hll_new(void) {
  kuqo_hll *h = (kuqo_hll *)calloc(1, sizeof(*h));
  h->sparse = 1;
  return h;
}
void kuqo_hll_free(kuqo_hll *h) {
  if (!h) return;
  free(h->M); free(h->set); free(h);
}
static void set_insert_raw(uint32_t *tab, uint32_t cap, uint32_t v) {
  uint32_t s = hash32(v) & (cap - 1);
  while (tab[s]) s = (s + 1) & (cap - 1);
  tab[s] = v;
}
static void set_insert(kuqo_hll *h, uint32_t v) {
  if (h->set_cap == 0 || 2 * (h->set_n + 1) > h->set_cap) {
    uint32_t ncap = h->set_cap ? h->set_cap * 2 : 64;
    uint32_t *nt = (uint32_t *)calloc(ncap, 4);
    for (uint32_t i = 0; i < h->set_cap; i++) if (h->set[i]) set_insert_raw(nt, ncap, h->set[i]);
    free(h->set);
    h->set = nt; h->set_cap = ncap;
  }
  uint32_t s = hash32(v) & (h->set_cap - 1);
  while (h->set[s]) {
    if (h->set[s] == v) return;
    s = (s + 1) & (h->set_cap - 1);
  }
  h->set[s] = v;
  h->set_n++;
}

static inline unsigned clz32_max(uint32_t x, unsigned max) { return x == 0 ? max : (unsigned)__builtin_clz(x); }   /* :68-71 */
static inline unsigned clz64_max(uint64_t x, unsigned max) { return x == 0 ? max : (unsigned)__builtin_clzll(x); } /* :73-76 */

/* encodeHashIn32Bit(hash, pPrime=25, p=12), hyperloglogplus.cpp:181-204 */
uint32_t kuqo_encode_hash32(uint64_t hash) {
  uint32_t idx = (uint32_t)((hash >> (64 - HLL_PPRIME)) << (32 - HLL_PPRIME));   /* :183 */
  if ((uint32_t)(idx << HLL_P) == 0) {                                           /* :191 */
    unsigned additional_rank = clz64_max(hash << HLL_PPRIME, 64 - HLL_PPRIME) + 1; /* getRank(h, pPrime), :140-147 */
    return idx | (uint32_t)(additional_rank << 1) | 1;                           /* :198 */
  }
  return idx;                                                                    /* :202 */
}
/* getEncodedRank(enc, pPrime=25, p=12), :152-161 */
static inline unsigned encoded_rank(uint32_t e) {
  if (e & 1) return (HLL_PPRIME - HLL_P) + ((e >> 1) & 0x3F);                    /* extractBits(e,7,1), :87-105 */
  return clz32_max(e << HLL_P, 32 - HLL_P) + 1;                                  /* getRank(uint32), :130-137 */
}
/* addToRegisters, :559-577 */
static void fold_code(uint8_t *M, uint32_t e) {
  uint32_t idx = e >> (32 - HLL_P);                                              /* getIndex(uint32), :121-124 */
  unsigned r = encoded_rank(e);
  if (r > M[idx]) M[idx] = (uint8_t)r;
}
/* switchToNormalRepresentation, :541-556 */
static void to_dense(kuqo_hll *h) {
  if (!h->sparse) return;
  h->sparse = 0;
  h->M = (uint8_t *)calloc(HLL_M, 1);
  for (uint32_t i = 0; i < h->set_cap; i++) if (h->set[i]) fold_code(h->M, h->set[i]);
  free(h->set); h->set = NULL; h->set_cap = 0; h->set_n = 0;
}

/* insert(), :485-523 */
void kuqo_hll_insert(kuqo_hll *h, uint64_t item) {
  ++h->n_observed;
  uint64_t hash = kuqo_fmix64(item);
  if (h->sparse && h->set_n + 1 > HLL_M / 4) to_dense(h);                        /* :496-498 */
  if (h->sparse) {
    set_insert(h, kuqo_encode_hash32(hash));                                     /* :499-503 (unordered_set insert) */
  } else {
    uint32_t idx = (uint32_t)(hash >> (64 - HLL_P));                             /* :514 */
    unsigned rank = clz64_max(hash << HLL_P, 64 - HLL_P) + 1;                    /* :516, getRank(uint64) :140-147 */
    if (rank > h->M[idx]) h->M[idx] = (uint8_t)rank;                             /* :519-521 */
  }
}

/* merge(), :586-665 (both overloads have the same observable effect on *dst) */
void kuqo_hll_merge(kuqo_hll *d, const kuqo_hll *o) {
  if (o->n_observed == 0) return;                                                /* :590-591 */
  if (d->n_observed == 0) {                                                      /* :593-597: adopt */
    d->n_observed = o->n_observed;
    d->sparse = o->sparse;
    free(d->set); d->set = NULL; d->set_cap = 0; d->set_n = 0;
    free(d->M); d->M = NULL;
    if (o->sparse) {
      for (uint32_t i = 0; i < o->set_cap; i++) if (o->set[i]) set_insert(d, o->set[i]);
    } else {
      d->M = (uint8_t *)malloc(HLL_M);
      memcpy(d->M, o->M, HLL_M);
    }
    return;
  }
  d->n_observed += o->n_observed;                                                /* :599 */
  if (d->sparse && o->sparse) {                                                  /* :600-603: union, NO size check */
    for (uint32_t i = 0; i < o->set_cap; i++) if (o->set[i]) set_insert(d, o->set[i]);
  } else if (o->sparse) {                                                        /* :604-606 */
    for (uint32_t i = 0; i < o->set_cap; i++) if (o->set[i]) fold_code(d->M, o->set[i]);
  } else if (d->sparse) {                                                        /* :608-612 */
    d->sparse = 0;
    d->M = (uint8_t *)malloc(HLL_M);
    memcpy(d->M, o->M, HLL_M);
    for (uint32_t i = 0; i < d->set_cap; i++) if (d->set[i]) fold_code(d->M, d->set[i]);
    free(d->set); d->set = NULL; d->set_cap = 0; d->set_n = 0;
  } else {                                                                       /* :614-620 */
    for (uint32_t i = 0; i < HLL_M; i++) if (o->M[i] > d->M[i]) d->M[i] = o->M[i];
  }
}

/* sigma(), :373-387 */
static double ertl_sigma(double x) {
  if (x == 1.0) return INFINITY;
  double prev, sigma_x = x, y = 1.0;
  do {
    prev = sigma_x;
    x *= x;
    sigma_x += x * y;
    y += y;
  } while (sigma_x != prev);
  return sigma_x;
}
/* tau(), :408-422 */
static double ertl_tau(double x) {
  if (x == 0.0 || x == 1.0) return 0.0;
  double prev, y = 1.0, tau_x = 1 - x;
  do {
    prev = tau_x;
    x = sqrt(x);
    y /= 2.0;
    tau_x -= pow(1 - x, 2) * y;
  } while (tau_x != prev);
  return tau_x / 3.0;
}
/* the estimator body of ertlCardinality(), :738-752, given the histogram C[0..q+1] */
static uint64_t ertl_from_hist(const int *C, size_t q, size_t m, uint64_t n_observed) {
  double est_denominator = m * ertl_tau(1.0 - (double)C[q + 1] / (double)m);
  for (int k = (int)q; k >= 1; --k) {
    est_denominator += C[k];
    est_denominator *= 0.5;
  }
  est_denominator += m * ertl_sigma((double)C[0] / (double)m);
  double m_sq_alpha_inf = (m / (2.0 * log(2))) * m;
  double est = m_sq_alpha_inf / est_denominator;
  return ((double)n_observed < est) ? n_observed : (uint64_t)round(est);
}
uint64_t kuqo_ertl_dense(const uint8_t *regs, uint64_t n_observed) {
  int C[64 + 2];
  memset(C, 0, sizeof(C));
  for (uint32_t i = 0; i < HLL_M; i++) ++C[regs[i]];                             /* registerHistogram, :337-354 */
  return ertl_from_hist(C, 64 - HLL_P, HLL_M, n_observed);
}
/* ertlCardinality(), :722-753 */
uint64_t kuqo_hll_cardinality(const kuqo_hll *h) {
  int C[64 + 2];
  memset(C, 0, sizeof(C));
  if (h->sparse) {
    /* sparseRegisterHistogram, :356-366: q = 64 - pPrime = 39, m = 2^25, ranks stay p=12-relative.  The
     * reference's C has q+2 = 41 slots and writes out of bounds for rank > 40 (p ~ 2^-40 per insert); we
     * keep 66 slots and, like the reference's arithmetic, never read beyond C[q+1]. */
    size_t m = HLL_MPRIME;
    for (uint32_t i = 0; i < h->set_cap; i++) if (h->set[i]) { ++C[encoded_rank(h->set[i])]; --m; }
    C[0] = (int)m;
    return ertl_from_hist(C, 64 - HLL_PPRIME, HLL_MPRIME, h->n_observed);
  }
  for (uint32_t i = 0; i < HLL_M; i++) ++C[h->M[i]];
  return ertl_from_hist(C, 64 - HLL_P, HLL_M, h->n_observed);
}
int kuqo_hll_is_sparse(const kuqo_hll *h) { return h->sparse; }
uint64_t kuqo_hll_n_observed(const kuqo_hll *h) { return h->n_observed; }
uint32_t kuqo_hll_sparse_size(const kuqo_hll *h) { return h->sparse ? h->set_n : 0; }
void kuqo_hll_registers(const kuqo_hll *h, uint8_t *regs) {
  if (!h->sparse) { memcpy(regs, h->M, HLL_M); return; }
  memset(regs, 0, HLL_M);
  for (uint32_t i = 0; i < h->set_cap; i++) if (h->set[i]) fold_code(regs, h->set[i]);
}

/* ======================================================================================================= */
/* per-taxon counters: unordered_map<uint32_t, ReadCounts<HLL>> (classify.cpp:78, readcounts.hpp:32-105)    */
/* ======================================================================================================= */
typedef struct {
  uint32_t taxid;
  uint64_t n_reads, n_kmers;
  kuqo_hll *hll;
} readcounts;
/* classifyExact (EXACT_COUNTING, classify.cpp:46-49): the container of ReadCounts is a set of canonical k-mers
 * (khset64_t) instead of a sketch; add_kmer inserts (readcounts.hpp:71-74), += is the set union (:76-81), the unique
 * count is the set size (:127-130).  All per-taxon sets of a run live in one set of (taxon, k-mer) pairs. */
typedef struct {
  uint32_t *taxon;
  uint64_t *kmer;
  uint8_t *used;
  uint64_t n, cap;
} pairset;
typedef struct {
  readcounts *e;
  uint32_t n, cap_e;
  uint32_t *slot;      /* open addressing: index+1 into e, 0 = empty */
  uint32_t cap_s;
  pairset *exact;      /* not NULL: exact counting; shared by the work units' maps and the global one */
} countmap;

static uint64_t pair_hash(uint32_t taxon, uint64_t kmer) { return kuqo_fmix64(kmer ^ ((uint64_t)taxon << 32) ^ taxon); }
static int ps_insert(pairset *p, uint32_t taxon, uint64_t kmer) {          /* 1 = new */
  if (p->cap == 0 || 2 * (p->n + 1) > p->cap) {
    uint64_t ncap = p->cap ? p->cap * 2 : 1024;
    uint32_t *nt = (uint32_t *)malloc(4 * ncap);
    uint64_t *nk = (uint64_t *)malloc(8 * ncap);
    uint8_t *nu = (uint8_t *)calloc(ncap, 1);
    for (uint64_t i = 0; i < p->cap; i++) {
      if (!p->used[i]) continue;
      uint64_t s = pair_hash(p->taxon[i], p->kmer[i]) & (ncap - 1);
      while (nu[s]) s = (s + 1) & (ncap - 1);
      nu[s] = 1; nt[s] = p->taxon[i]; nk[s] = p->kmer[i];
    }
    free(p->taxon); free(p->kmer); free(p->used);
    p->taxon = nt; p->kmer = nk; p->used = nu; p->cap = ncap;
  }
  uint64_t s = pair_hash(taxon, kmer) & (p->cap - 1);
  while (p->used[s]) {
    if (p->taxon[s] == taxon && p->kmer[s] == kmer) return 0;
    s = (s + 1) & (p->cap - 1);
  }
  p->used[s] = 1; p->taxon[s] = taxon; p->kmer[s] = kmer; p->n++;
  return 1;
}
static void ps_free(pairset *p) { if (p) { free(p->taxon); free(p->kmer); free(p->used); free(p); } }

static void cm_init(countmap *m) { memset(m, 0, sizeof(*m)); }
static void cm_clear(countmap *m) {
  pairset *keep = m->exact;
  for (uint32_t i = 0; i < m->n; i++) kuqo_hll_free(m->e[i].hll);
  free(m->e); free(m->slot);
  memset(m, 0, sizeof(*m));
  m->exact = keep;
}
static readcounts *cm_get(countmap *m, uint32_t taxid) {               /* operator[] */
  if (m->cap_s == 0 || 2 * (m->n + 1) > m->cap_s) {
    uint32_t ncap = m->cap_s ? m->cap_s * 2 : 64;
    uint32_t *ns = (uint32_t *)calloc(ncap, 4);
    for (uint32_t i = 0; i < m->n; i++) {
      uint32_t s = hash32(m->e[i].taxid) & (ncap - 1);
      while (ns[s]) s = (s + 1) & (ncap - 1);
      ns[s] = i + 1;
    }
    free(m->slot); m->slot = ns; m->cap_s = ncap;
  }
  uint32_t s = hash32(taxid) & (m->cap_s - 1);
  while (m->slot[s]) {
    if (m->e[m->slot[s] - 1].taxid == taxid) return &m->e[m->slot[s] - 1];
    s = (s + 1) & (m->cap_s - 1);
  }
  if (m->n == m->cap_e) {
    m->cap_e = m->cap_e ? m->cap_e * 2 : 32;
    m->e = (readcounts *)realloc(m->e, sizeof(readcounts) * m->cap_e);
  }
  readcounts *rc = &m->e[m->n];
  rc->taxid = taxid; rc->n_reads = 0; rc->n_kmers = 0; rc->hll = kuqo_hll_new();
  m->slot[s] = ++m->n;
  return rc;
}

/* ======================================================================================================= */
/* one read                                                                                                */
/* ======================================================================================================= */

/* classify_sequence(), classify.cpp:897-1012 (Quick_mode off, one database, Map_UIDs off).  `counts` receives
 * add_kmer (:939) for every non-ambiguous window — taxon 0 for misses — and incrementReadCount (:968). */
static uint32_t classify_read_multi(const kuqo_db *const *dbs, uint32_t n_db, const kuqo_parent_map *pm, const char *seq,
                                    size_t len, uint32_t *codes, uint32_t *n_windows, countmap *counts,
                                    uint32_t quick_min, int quick_stop);
static uint32_t classify_read_into(const kuqo_db *db, const kuqo_parent_map *pm, const char *seq, size_t len,
                                   uint32_t *codes, uint32_t *n_windows, countmap *counts) {
  return classify_read_multi(&db, 1, pm, seq, len, codes, n_windows, counts, 0, 0);
}
/* several databases: the first one that holds the key decides, even if it stores taxon 0 (classify.cpp:928-936).
 * Quick mode (-q, -m quick_min): the preloaded path leaves the k-mer loop at the quick_min-th hit (:943-944,
 * quick_stop = 1) and calls that hit's taxon (:963-964); the -x path looks at every k-mer, stops COUNTING hits at
 * quick_min (:701-702) and calls the taxon of the read's last non-ambiguous k-mer (:705-721,737-738). */
static uint32_t classify_read_multi(const kuqo_db *const *dbs, uint32_t n_db, const kuqo_parent_map *pm, const char *seq,
                                    size_t len, uint32_t *codes, uint32_t *n_windows, countmap *counts,
                                    uint32_t quick_min, int quick_stop) {
  const kuqo_db *db = dbs[0];
  uint32_t n = 0, quick_hits = 0, last_taxon = 0;
  uint32_t *hit_taxa = NULL, *hit_cnt = NULL, n_hits = 0;
  if (len >= db->k) {
    size_t cap = len - db->k + 2;
    uint64_t *kmers = (uint64_t *)malloc(8 * cap);
    uint8_t *amb = (uint8_t *)malloc(cap);
    hit_taxa = (uint32_t *)malloc(4 * cap);
    hit_cnt = (uint32_t *)malloc(4 * cap);
    n = kuqo_scan(seq, len, db->k, kmers, amb);
    for (uint32_t i = 0; i < n; i++) {
      uint32_t taxon = 0;
      if (amb[i]) {                                                         /* :920-923 */
        codes[i] = KUQO_AMBIG;
        continue;
      }
      uint64_t canon = kuqo_canonical(kmers[i], db->k);                     /* :925 */
      uint32_t v;
      for (uint32_t d = 0; d < n_db; d++)                                   /* :928-936 */
        if (kuqo_kmer_query(dbs[d], canon, &v)) { taxon = v; break; }
      if (counts) {
        readcounts *rc = cm_get(counts, taxon);                             /* :939 */
        ++rc->n_kmers;
        if (counts->exact) ps_insert(counts->exact, taxon, canon);
        else kuqo_hll_insert(rc->hll, canon);
      }
      last_taxon = taxon;
      if (taxon) {                                                          /* :941-942 */
        uint32_t j = 0;
        for (; j < n_hits; j++) if (hit_taxa[j] == taxon) break;
        if (j == n_hits) { hit_taxa[n_hits] = taxon; hit_cnt[n_hits] = 0; n_hits++; }
        hit_cnt[j]++;
        if (quick_min && quick_hits < quick_min) quick_hits++;
        if (quick_min && quick_stop && quick_hits >= quick_min) {           /* :943-944: leaves before :947 */
          codes[i] = taxon;
          n = i + 1;
          break;
        }
      }
      codes[i] = taxon;                                                     /* :947 */
    }
    free(kmers); free(amb);
  }
  uint32_t call = quick_min ? (quick_hits >= quick_min ? last_taxon : 0)    /* :963-964 */
                            : kuqo_resolve_tree(pm, hit_taxa, hit_cnt, n_hits);   /* :965 */
  free(hit_taxa); free(hit_cnt);
  if (counts) cm_get(counts, call)->n_reads++;                              /* :968 */
  *n_windows = n;
  return call;
}

uint32_t kuqo_classify_read(const kuqo_db *db, const kuqo_parent_map *pm, const char *seq, size_t len,
                            uint32_t *codes_out, uint32_t *n_windows) {
  return classify_read_into(db, pm, seq, len, codes_out, n_windows, NULL);
}

/* hitlist_string(), classify.cpp:826-861; "0:0" for an empty list (:994-995) */
size_t kuqo_hitlist_string(const uint32_t *codes, uint32_t n, char *buf, size_t cap) {
  size_t w = 0;
  if (n == 0) return (size_t)snprintf(buf, cap, "0:0");
  int64_t last_code = codes[0] == KUQO_AMBIG ? -1 : (int64_t)codes[0];
  int code_count = 1;
  for (uint32_t i = 1; i < n; i++) {
    int64_t code = codes[i] == KUQO_AMBIG ? -1 : (int64_t)codes[i];
    if (code == last_code) {
      code_count++;
    } else {
      if (last_code >= 0) w += (size_t)snprintf(buf + w, w < cap ? cap - w : 0, "%lld:%d ", (long long)last_code, code_count);
      else w += (size_t)snprintf(buf + w, w < cap ? cap - w : 0, "A:%d ", code_count);
      code_count = 1;
      last_code = code;
    }
  }
  if (last_code >= 0) w += (size_t)snprintf(buf + w, w < cap ? cap - w : 0, "%lld:%d", (long long)last_code, code_count);
  else w += (size_t)snprintf(buf + w, w < cap ? cap - w : 0, "A:%d", code_count);
  return w;
}

/* ======================================================================================================= */
/* whole run                                                                                               */
/* ======================================================================================================= */
struct kuqo_run {
  const kuqo_db *dbs[16];
  uint32_t n_db;
  const kuqo_db *db;
  const kuqo_parent_map *pm;
  uint64_t unit_size;
  int mode;
  countmap global;      /* taxon_counts, classify.cpp:78 */
  countmap local;       /* my_taxon_counts of the open work unit, :525 */
  uint64_t unit_nt;     /* total_nt of the open work unit, :508,519 */
  uint32_t quick_min;   /* -q: Minimum_hit_count (-m, default 1); 0 = quick mode off */
};

kuqo_run *kuqo_run_new(const kuqo_db *db, const kuqo_parent_map *pm, uint64_t work_unit_size, int mode) {
  kuqo_run *r = (kuqo_run *)calloc(1, sizeof(*r));
  r->db = db; r->dbs[0] = db; r->n_db = 1; r->pm = pm; r->unit_size = work_unit_size ? work_unit_size : 500000; r->mode = mode;
  cm_init(&r->global); cm_init(&r->local);
  return r;
}
void kuqo_run_free(kuqo_run *r) {
  if (!r) return;
  ps_free(r->global.exact);
  cm_clear(&r->global); cm_clear(&r->local); free(r);
}
/* critical(write_output) merge, classify.cpp:542-544: taxon_counts[t] += move(local[t])
 * (ReadCounts::operator+=, readcounts.hpp:83-88 → HLL merge) */
static void flush_unit(kuqo_run *r) {
  for (uint32_t i = 0; i < r->local.n; i++) {
    readcounts *src = &r->local.e[i];
    readcounts *dst = cm_get(&r->global, src->taxid);
    dst->n_reads += src->n_reads;
    dst->n_kmers += src->n_kmers;
    kuqo_hll_merge(dst->hll, src->hll);
  }
  cm_clear(&r->local);
  cm_init(&r->local);
  r->local.exact = r->global.exact;
  r->unit_nt = 0;
}
int kuqo_run_classify(kuqo_run *r, const char *bases, const uint64_t *offsets, uint32_t n_reads,
                      uint32_t *calls_out, uint32_t *codes_out, uint64_t *code_offsets_out) {
  uint64_t code_pos = 0;
  uint32_t *scratch = NULL;
  size_t scratch_cap = 0;
  for (uint32_t i = 0; i < n_reads; i++) {
    const char *seq = bases + offsets[i];
    size_t len = (size_t)(offsets[i + 1] - offsets[i]);
    if (len + 2 > scratch_cap) { scratch_cap = len + 2; scratch = (uint32_t *)realloc(scratch, 4 * scratch_cap); }
    uint32_t nw = 0;
    /* preload: work unit's private map (:525,530-535); chunked: the global map directly (:719,747) */
    countmap *target = r->mode == 0 ? &r->local : &r->global;
    uint32_t call = classify_read_multi(r->dbs, r->n_db, r->pm, seq, len, scratch, &nw, target, r->quick_min, r->mode == 0);
    calls_out[i] = call;
    if (code_offsets_out) code_offsets_out[i] = code_pos;
    if (codes_out) memcpy(codes_out + code_pos, scratch, 4 * (size_t)nw);
    code_pos += nw;
    if (r->mode == 0) {
      r->unit_nt += len;                                   /* :519 */
      if (r->unit_nt >= r->unit_size) flush_unit(r);       /* loop condition :514 closes the unit */
    }
  }
  if (code_offsets_out) code_offsets_out[n_reads] = code_pos;
  free(scratch);
  return 0;
}
void kuqo_run_finish(kuqo_run *r) {
  if (r->mode == 0 && r->local.n) flush_unit(r);
}
uint32_t kuqo_run_n_taxa(const kuqo_run *r) { return r->global.n; }
static int cmp_rc(const void *a, const void *b) {
  uint32_t x = ((const readcounts *)a)->taxid, y = ((const readcounts *)b)->taxid;
  return x < y ? -1 : x > y;
}
void kuqo_run_counts(const kuqo_run *r, uint32_t *taxid, uint64_t *n_reads, uint64_t *n_kmers,
                     uint64_t *unique_est, uint8_t *is_sparse, uint8_t *regs) {
  readcounts *tmp = (readcounts *)malloc(sizeof(readcounts) * (r->global.n + 1));
  memcpy(tmp, r->global.e, sizeof(readcounts) * r->global.n);
  qsort(tmp, r->global.n, sizeof(readcounts), cmp_rc);
  for (uint32_t i = 0; i < r->global.n; i++) {
    taxid[i] = tmp[i].taxid;
    n_reads[i] = tmp[i].n_reads;
    n_kmers[i] = tmp[i].n_kmers;
    unique_est[i] = kuqo_hll_cardinality(tmp[i].hll);      /* uniqueKmerCount, readcounts.hpp:121-124 */
    if (r->global.exact) {                                 /* set size, readcounts.hpp:127-130 */
      const pairset *ps = r->global.exact;
      uint64_t c = 0;
      for (uint64_t s = 0; s < ps->cap; s++) c += ps->used[s] && ps->taxon[s] == tmp[i].taxid;
      unique_est[i] = c;
    }
    is_sparse[i] = (uint8_t)tmp[i].hll->sparse;
    if (regs) kuqo_hll_registers(tmp[i].hll, regs + (size_t)HLL_M * i);
  }
  free(tmp);
}

/* Clade roll-up of TaxReport's constructor (taxdb.hpp:956-973): ReadCounts of the listed taxa summed with
 * operator+= (readcounts.hpp:76-81 → HLL merge :627-665).  Returns the clade's uniqueKmerCount. */
uint64_t kuqo_run_clade(const kuqo_run *r, const uint32_t *taxa, uint32_t n, uint64_t *n_reads, uint64_t *n_kmers) {
  kuqo_hll *acc = kuqo_hll_new();
  uint64_t reads = 0, kmers = 0;
  for (uint32_t i = 0; i < n; i++) {
    for (uint32_t j = 0; j < r->global.n; j++) {
      if (r->global.e[j].taxid != taxa[i]) continue;
      reads += r->global.e[j].n_reads;
      kmers += r->global.e[j].n_kmers;
      kuqo_hll_merge(acc, r->global.e[j].hll);
      break;
    }
  }
  uint64_t u = kuqo_hll_cardinality(acc);
  kuqo_hll_free(acc);
  if (r->global.exact) {                                   /* union of the members' k-mer sets */
    const pairset *ps = r->global.exact;
    pairset *un = (pairset *)calloc(1, sizeof(*un));
    for (uint64_t s = 0; s < ps->cap; s++) {
      if (!ps->used[s]) continue;
      for (uint32_t i = 0; i < n; i++)
        if (taxa[i] == ps->taxon[s]) { ps_insert(un, 0, ps->kmer[s]); break; }
    }
    u = un->n;
    ps_free(un);
  }
  if (n_reads) *n_reads = reads;
  if (n_kmers) *n_kmers = kmers;
  return u;
}

/* classify -d db1 -d db2 ...: databases are tried in command-line order for every k-mer (classify.cpp:928-936) */
int kuqo_run_add_db(kuqo_run *r, const kuqo_db *db) {
  if (r->n_db >= 16 || db->k != r->dbs[0]->k) return -1;      /* all databases must share k (:203-210) */
  r->dbs[r->n_db++] = db;
  return 0;
}

/* classify -q [-m min_hits]; min_hits = 0 switches quick mode off */
void kuqo_run_set_quick(kuqo_run *r, uint32_t min_hits) { r->quick_min = min_hits; }

/* classifyExact: exact distinct k-mer counts per taxon / clade instead of the HLL estimate (classify.cpp:46-49) */
void kuqo_run_set_exact(kuqo_run *r, int on) {
  if (on && !r->global.exact) r->global.exact = (pairset *)calloc(1, sizeof(pairset));
  if (!on) { ps_free(r->global.exact); r->global.exact = NULL; }
  r->local.exact = r->global.exact;
}

/* ======================================================================================================= */
/* database build: db_sort and set_lcas                                                  */
/* ======================================================================================================= */

/* db_sort.cpp:41-78 (main) with make_index (krakendb.cpp:118-148) and bin_and_sort_data (db_sort.cpp:80-116):
 * records go to the bin of bin_key(key, nt) — the key as stored, no canonicalisation — in input order, then every
 * bin is sorted by key; the header is copied; the index is always written as KRAKIX2.  Keys are distinct in a
 * Jellyfish dump, so qsort's instability cannot show.  kdb_out: jdb_bytes, idx_out: 8 + 8 * (4^nt + 1) bytes. */
typedef struct { uint64_t key; uint32_t pos; } sort_item;
static int cmp_item(const void *a, const void *b) {
  uint64_t x = ((const sort_item *)a)->key, y = ((const sort_item *)b)->key;
  return x < y ? -1 : x > y;
}
int kuqo_db_sort(const void *jdb_image, uint64_t jdb_bytes, uint32_t nt, int zero_vals, void *kdb_out, void *idx_out) {
  const uint8_t *p = (const uint8_t *)jdb_image;
  if (!p || jdb_bytes < 56 || memcmp(p, "JFLISTDN", 8) != 0 || nt < 1 || nt > 15) return -1;
  uint64_t key_bits, val_len, key_ct;
  memcpy(&key_bits, p + 8, 8); memcpy(&val_len, p + 16, 8); memcpy(&key_ct, p + 48, 8);
  if (val_len != 4) return -2;
  const unsigned k = (unsigned)(key_bits / 2), key_len = (unsigned)(key_bits / 8 + !!(key_bits % 8));
  const uint64_t pair_sz = key_len + 4, header = 72 + 2 * (4 + 8 * key_bits);
  if (jdb_bytes < header + key_ct * pair_sz) return -3;
  const uint8_t *src = p + header;
  const uint64_t entries = 1ull << (2 * nt);
  uint64_t *offsets = (uint64_t *)calloc(entries + 1, 8);
  uint64_t *bins = (uint64_t *)malloc(8 * (key_ct ? key_ct : 1));
  for (uint64_t i = 0; i < key_ct; i++) {                                      /* make_index, :126-134 */
    uint64_t kmer = 0;
    memcpy(&kmer, src + i * pair_sz, key_len);
    bins[i] = kuqo_bin_key(kmer, k, nt, 2);
    offsets[bins[i] + 1]++;
  }
  for (uint64_t i = 1; i <= entries; i++) offsets[i] += offsets[i - 1];        /* :136-139 */
  uint8_t *idx = (uint8_t *)idx_out;
  memcpy(idx, "KRAKIX2", 7);                                                   /* :144-147 */
  idx[7] = (uint8_t)nt;
  memcpy(idx + 8, offsets, 8 * (entries + 1));
  uint8_t *out = (uint8_t *)kdb_out;
  memcpy(out, p, header);                                                      /* db_sort.cpp:56-58,71 */
  uint8_t *data = out + header;
  uint64_t *pos = (uint64_t *)malloc(8 * entries);
  memcpy(pos, offsets, 8 * entries);
  for (uint64_t i = 0; i < key_ct; i++) {                                      /* :96-107 */
    uint8_t *dst = data + pair_sz * pos[bins[i]]++;
    memcpy(dst, src + i * pair_sz, pair_sz);
    if (zero_vals) memset(dst + key_len, 0, 4);
  }
  sort_item *tmp = NULL;
  uint64_t tmp_cap = 0;
  uint8_t *buf = NULL;
  for (uint64_t b = 0; b < entries; b++) {                                     /* :110-115, pair_cmp :118-128 */
    const uint64_t n = offsets[b + 1] - offsets[b];
    if (n < 2) continue;
    if (n > tmp_cap) { tmp_cap = n; tmp = (sort_item *)realloc(tmp, sizeof(sort_item) * n); buf = (uint8_t *)realloc(buf, pair_sz * n); }
    uint8_t *base = data + offsets[b] * pair_sz;
    for (uint64_t i = 0; i < n; i++) { tmp[i].key = 0; memcpy(&tmp[i].key, base + i * pair_sz, key_len); tmp[i].pos = (uint32_t)i; }
    qsort(tmp, n, sizeof(sort_item), cmp_item);
    for (uint64_t i = 0; i < n; i++) memcpy(buf + i * pair_sz, base + (uint64_t)tmp[i].pos * pair_sz, pair_sz);
    memcpy(base, buf, n * pair_sz);
  }
  free(tmp); free(buf); free(pos); free(bins); free(offsets);
  return 0;
}

/* set_lcas() for one library sequence, set_lcas.cpp:429-476 (taxid mode, no forced contaminants): every
 * unambiguous k-mer found in the database gets value = lca(taxid, value).  process_single_file cuts the sequence
 * into pieces of SKIP_LEN that overlap by k-1 (:363-364), which visits every k-mer exactly once, like one scan.
 * `db` must have been opened on a writable image.  Returns the number of k-mers that were not in the database
 * (the reference stops at the first one unless -x, :441-443). */
uint64_t kuqo_set_lcas_sequence(kuqo_db *db, const kuqo_parent_map *pm, const char *seq, size_t len, uint32_t taxid) {
  return kuqo_set_lcas_sequence_flags(db, pm, seq, len, taxid, 0);
}
/* flags: 1 = -T (Force_contaminant_taxid, :462-474: values 32630 / 81077 stick, a sequence with one of these taxids
 * overwrites instead of taking the LCA), 2 = -R (Reset_taxid, :458-459: the value becomes 0) */
uint64_t kuqo_set_lcas_sequence_flags(kuqo_db *db, const kuqo_parent_map *pm, const char *seq, size_t len, uint32_t taxid,
                                      uint32_t flags) {
  const uint32_t C1 = 32630, C2 = 81077;                                        /* set_lcas.cpp:88-89 */
  uint64_t missing = 0;
  if (len < db->k) return 0;
  size_t cap = len - db->k + 2;
  uint64_t *kmers = (uint64_t *)malloc(8 * cap);
  uint8_t *amb = (uint8_t *)malloc(cap);
  uint32_t n = kuqo_scan(seq, len, db->k, kmers, amb);
  for (uint32_t i = 0; i < n; i++) {
    if (amb[i]) continue;                                                       /* :434-435 */
    const uint64_t kmer = kuqo_canonical(kmers[i], db->k);
    const uint64_t b = kuqo_bin_key(kmer, db->k, db->nt, db->idx_type);
    int64_t lo = (int64_t)db->offsets[b], hi = (int64_t)db->offsets[b + 1] - 1, at = -1;
    while (lo <= hi) {
      int64_t mid = lo + (hi - lo) / 2;
      uint64_t c = db_key_at(db, (uint64_t)mid);
      if (kmer > c) lo = mid + 1; else if (kmer < c) hi = mid - 1; else { at = mid; break; }
    }
    if (at < 0) { missing++; continue; }                                        /* :439-447 */
    uint8_t *val = (uint8_t *)(uintptr_t)(db->pairs + db->pair_sz * (uint64_t)at + db->key_len);
    uint32_t v;
    memcpy(&v, val, 4);
    if (flags & 2u) v = 0;                                                      /* :458-459 */
    else if (!(flags & 1u)) v = kuqo_lca(pm, taxid, v);                         /* :461 */
    else if (v == C1 || v == C2) { /* keep, :465-466 */ }
    else if (taxid == C1 || taxid == C2) v = taxid;                             /* :467-470 */
    else v = kuqo_lca(pm, taxid, v);                                            /* :472 */
    memcpy(val, &v, 4);
  }
  free(kmers); free(amb);
  return missing;
}
