/*
  gcsa_oracle.h -- CPU restatement of the GCSA2 query path.  TEST INFRASTRUCTURE ONLY.

  This is the parity oracle for gcsa2_b200.  Only tests/, __graft_entry__.smoke()
  and bench.py's cpu_baseline / --impl reference legs may load it; the product
  (gcsa2_b200/) never does.

  Every function cites the reference file:line (relative to jltsiren/gcsa2) whose
  behaviour it restates.  The rank/select/access arithmetic of the reference lives
  in the external sdsl-lite fork (vgteam), which is NOT in the reference tree and
  not installed here; it is restated from the mathematical definitions
  (rank1(i) = ones in [0,i), select1(k) = position of the k-th one, k >= 1).
  PARITY PINNING: pinned against the reference itself.  oracle/_ref/libgcsa2_ref.so is the
  reference's own, unmodified sources compiled against the SDSL shim of oracle/sdsl_shim/ (the
  real sdsl-lite fork is unavailable); tests/test_reference.py checks that every function below
  returns exactly what the reference's own method returns, on indexes the reference's own
  constructor built and its own verifyIndex() accepted.  Also pinned on the paper's Figure 3
  worked example (tests/golden/kat1_fig3.json) and the verifyIndex predicates
  (src/algorithms.cpp:101-295).  What stays unpinned is SDSL itself (not in the tree): rank,
  select and access have unique mathematical answers, and the shim implements those.
*/
#ifndef GCSA_ORACLE_H
#define GCSA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_SIGMA 7
#define ORACLE_UNKNOWN (~(uint64_t)0)

/* Plain bit vector with an interleaved 512-bit rank directory (count word + 8 data
   words per block), the shape of sdsl::bit_vector_il<512> [SDSL, not in tree]. */
typedef struct {
  uint64_t  n_bits;
  uint64_t  n_blocks;
  uint64_t  ones;
  uint64_t* il;          /* n_blocks * 9 words */
} oracle_bv;

typedef struct {
  uint64_t path_nodes, edge_count, order;
  uint64_t sigma, fast_chars;
  uint64_t C[ORACLE_SIGMA + 1];
  uint8_t  char2comp[256];
  oracle_bv bwt[ORACLE_SIGMA];   /* fast_bwt[1..4] and sparse_bwt[0,5,6], all as plain bits */
  oracle_bv edges;
  oracle_bv sampled_paths;
  oracle_bv samples;             /* select_1 */
  uint64_t  sample_count;
  uint64_t* stored_samples;      /* unpacked copy */
  oracle_bv extra_filter;        /* SadaSparse::filter */
  oracle_bv extra_values;        /* SadaSparse::values */
  oracle_bv redundant;           /* SadaCount::data */
} oracle_gcsa;

typedef struct {
  uint64_t size, branching, levels, values;
  uint64_t* offsets;             /* levels + 1 */
  uint8_t*  data;                /* values */
} oracle_lcp;

typedef struct { uint64_t sp, ep, left_lcp, right_lcp, node_lcp; } oracle_stnode;

/* Flat description handed over by the caller (same shape as gcsa_flat_index in
   include/gcsa2_b200.h, restated here so the oracle does not include product headers). */
typedef struct {
  uint64_t path_nodes, edge_count, order, sigma, fast_chars;
  uint64_t C[ORACLE_SIGMA + 1];
  uint8_t  char2comp[256];
  const uint64_t* bwt[ORACLE_SIGMA];
  const uint64_t* edges;
  const uint64_t* sampled_paths;
  uint64_t sample_count;
  const uint64_t* stored_samples;
  const uint64_t* samples;
  const uint64_t* extra_filter;
  uint64_t extra_values_len;
  const uint64_t* extra_values;
  uint64_t redundant_len;
  const uint64_t* redundant;
} oracle_flat;

oracle_gcsa* oracle_gcsa_create(const oracle_flat* flat);
void         oracle_gcsa_destroy(oracle_gcsa* g);

oracle_lcp*  oracle_lcp_create(uint64_t size, uint64_t branching, uint64_t levels,
                               const uint64_t* offsets, const uint8_t* data);
void         oracle_lcp_destroy(oracle_lcp* l);

/* rank / select / access primitives (exposed so tests can check them against naive loops) */
uint64_t oracle_bv_rank(const oracle_bv* v, uint64_t i);
uint64_t oracle_bv_select(const oracle_bv* v, uint64_t k);
int      oracle_bv_get(const oracle_bv* v, uint64_t i);
oracle_bv* oracle_bv_create(const uint64_t* words, uint64_t n_bits);
void       oracle_bv_destroy(oracle_bv* v);

/* GCSA queries */
void     oracle_find(const oracle_gcsa* g, const uint8_t* pattern, uint64_t len, uint64_t* sp, uint64_t* ep);
void     oracle_char_range(const oracle_gcsa* g, uint64_t comp, uint64_t* sp, uint64_t* ep);
void     oracle_lf_range(const oracle_gcsa* g, uint64_t sp, uint64_t ep, uint64_t comp, uint64_t* osp, uint64_t* oep);
uint64_t oracle_lf_node(const oracle_gcsa* g, uint64_t path_node);
void     oracle_lf_fast(const oracle_gcsa* g, uint64_t sp, uint64_t ep, uint64_t* out /* 2*sigma */);
void     oracle_lf_all(const oracle_gcsa* g, uint64_t sp, uint64_t ep, uint64_t* out /* 2*sigma */);
uint64_t oracle_count(const oracle_gcsa* g, uint64_t sp, uint64_t ep);
/* locate: results are malloc'ed into *out (caller frees with oracle_free), count returned */
uint64_t oracle_locate_node(const oracle_gcsa* g, uint64_t path_node, uint64_t** out);
uint64_t oracle_locate_range(const oracle_gcsa* g, uint64_t sp, uint64_t ep, uint64_t** out);
uint64_t oracle_locate_max(const oracle_gcsa* g, uint64_t sp, uint64_t ep, uint64_t max_positions, uint64_t** out);
void     oracle_free(void* p);

/* Statistics variant of find(): executed LF steps and distinct 64-byte rank probes
   (SURVEY.md section 8(d) accounting: two probes of one vector with equal i>>9 count once;
   edges probes are not issued when the edge-space range is empty). */
void oracle_find_stats(const oracle_gcsa* g, const uint8_t* pattern, uint64_t len,
                       uint64_t* sp, uint64_t* ep, uint64_t* steps, uint64_t* probes);

/* Batch drivers (the loop of benchmark/query_gcsa.cpp:88-103, optionally OpenMP over queries
   the way src/algorithms.cpp:113 parallelises its callers).  Returns seconds (omp_get_wtime). */
double oracle_find_batch(const oracle_gcsa* g, const uint8_t* chars, const uint64_t* offsets,
                         uint64_t n, uint64_t* sp, uint64_t* ep, int threads);
double oracle_find_batch_stats(const oracle_gcsa* g, const uint8_t* chars, const uint64_t* offsets,
                         uint64_t n, uint64_t* sp, uint64_t* ep, int threads,
                         uint64_t* total_steps, uint64_t* total_probes);
double oracle_count_batch(const oracle_gcsa* g, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                          uint64_t* out, int threads);
/* locate batch: two passes; out_offsets has n+1 entries; *values malloc'ed */
double oracle_locate_batch(const oracle_gcsa* g, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                           uint64_t* out_offsets, uint64_t** values, int threads);
int    oracle_max_threads(void);
/* countKMers(index, k), src/algorithms.cpp:387-421 */
uint64_t oracle_count_kmers(const oracle_gcsa* g, uint64_t k, int include_Ns, int threads);

/* KMerComparisonState, src/algorithms.cpp:425-460 (the record compareKMers writes to .left / .right) */
typedef struct oracle_kmer_cmp { uint64_t left_sp, left_ep, right_sp, right_ep, k, kmer[3]; } oracle_kmer_cmp;
/* compareKMers, src/algorithms.cpp:535-616; result[3] = shared, left only, right only */
void oracle_compare_kmers(const oracle_gcsa* left, const oracle_gcsa* right, uint64_t k, int include_Ns, int threads,
                          uint64_t* result, oracle_kmer_cmp** left_kmers, oracle_kmer_cmp** right_kmers);

/* LCP queries */
void     oracle_lcp_parent(const oracle_lcp* l, uint64_t sp, uint64_t ep, oracle_stnode* out);
uint64_t oracle_lcp_depth(const oracle_lcp* l, uint64_t sp, uint64_t ep);
void     oracle_lcp_psv(const oracle_lcp* l, uint64_t pos, uint64_t* rpos, uint64_t* rval);
void     oracle_lcp_psev(const oracle_lcp* l, uint64_t pos, uint64_t* rpos, uint64_t* rval);
void     oracle_lcp_nsv(const oracle_lcp* l, uint64_t pos, uint64_t* rpos, uint64_t* rval);
void     oracle_lcp_nsev(const oracle_lcp* l, uint64_t pos, uint64_t* rpos, uint64_t* rval);
void     oracle_lcp_rmq(const oracle_lcp* l, uint64_t sp, uint64_t ep, uint64_t* rpos, uint64_t* rval);
double   oracle_parent_batch(const oracle_lcp* l, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                             oracle_stnode* out, int threads);

/* MEM-style scan (config 5; this repository's driver over LF + parent, defined in gcsa_oracle.c):
   matches of pattern i are the 4-tuples (start, length, sp, ep) at matches[4 * out_offsets[i] ...). */
double   oracle_mem_batch(const oracle_gcsa* g, const oracle_lcp* l, const uint8_t* chars, const uint64_t* offsets,
                          uint64_t n, uint64_t* out_offsets, uint64_t** matches, int threads);

/* std::mt19937_64 and Thomas Wang's hash, exposed for known-answer tests */
typedef struct { uint64_t mt[312]; int idx; } oracle_mt64;
void     oracle_mt64_seed(oracle_mt64* r, uint64_t seed);
uint64_t oracle_mt64_next(oracle_mt64* r);
uint64_t oracle_wang_hash_64(uint64_t key);

#ifdef __cplusplus
}
#endif
#endif
