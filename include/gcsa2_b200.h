/*
  gcsa2_b200.h -- C ABI of the B200 batched backward-search engine for GCSA2 indexes.

  This is the drop-in boundary.  The reference (jltsiren/gcsa2) exposes the query path as the
  C++ class gcsa::GCSA / gcsa::LCPArray with header-inline methods; a foreign-function binding
  of that path would bind exactly the operations below, one batch entry point per method.  Each
  entry point cites the reference interface it replaces (file:line in the reference tree).

  Conventions
    * plain pointers and sizes, no C++ or torch types; status return: 0 = ok, < 0 = GCSA_B200_ERR_*;
      no exceptions cross the boundary; gcsa_b200_last_error() gives the message of the calling
      thread's last failure.
    * "*_batch" entry points take DEVICE pointers and are stream-ordered on `stream`
      (a cudaStream_t passed as void*; NULL = the legacy default stream).  They never synchronise.
    * "*_host" entry points take HOST pointers; they copy in, run the same kernels, copy out and
      return when the results are in the caller's buffers.
    * ranges are closed [sp, ep] pairs of path-node ranks, empty iff sp + 1 > ep + 1
      (include/gcsa/utils.h:84-117); empty results of find()/LF() are returned uncanonicalised
      exactly as the reference returns them (include/gcsa/gcsa.h:160).
    * handles are immutable after creation; any number of host threads may issue queries on one
      handle concurrently (the reference's query methods are const, src/algorithms.cpp:113).
    * the library fails loudly (GCSA_B200_ERR_CUDA) when there is no usable CUDA device; there is
      no CPU fallback.
*/
#ifndef GCSA2_B200_H
#define GCSA2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCSA_B200_SIGMA 7              /* $ A C G T N #  (src/support.cpp:92) */
#define GCSA_B200_FAST_CHARS 4         /* comps 1..4 (include/gcsa/support.h:108) */
#define GCSA_B200_UNKNOWN (~(uint64_t)0)   /* STNode::UNKNOWN, include/gcsa/lcp.h:46 */

#define GCSA_B200_OK 0
#define GCSA_B200_ERR_INVALID      (-1)   /* bad argument */
#define GCSA_B200_ERR_CUDA         (-2)   /* CUDA runtime failure / no device */
#define GCSA_B200_ERR_NOMEM        (-3)
#define GCSA_B200_ERR_CAPACITY     (-4)   /* output buffer too small (locate) */
#define GCSA_B200_ERR_INCONSISTENT (-5)   /* builder: input violates the GCSA invariants */

/* ---------------------------------------------------------------------------------------------
   The index as plain host arrays (members of gcsa::GCSA, include/gcsa/gcsa.h:214-240).
   Every bit vector is little-endian 64-bit words: bit i = (words[i >> 6] >> (i & 63)) & 1.
   --------------------------------------------------------------------------------------------- */
typedef struct gcsa_flat_index {
  uint64_t path_nodes;                 /* GCSAHeader::path_nodes, include/gcsa/files.h:135-156 */
  uint64_t edge_count;                 /* GCSAHeader::edges */
  uint64_t order;                      /* GCSAHeader::order */
  uint64_t sigma, fast_chars;          /* Alphabet::sigma, fast_chars (support.h:149-151) */
  uint64_t C[GCSA_B200_SIGMA + 1];     /* Alphabet::C */
  uint8_t  char2comp[256];             /* Alphabet::char2comp */
  const uint64_t* bwt[GCSA_B200_SIGMA];/* fast_bwt[1..4], sparse_bwt[0,5,6]: path_nodes bits each */
  const uint64_t* edges;               /* edge_count bits, 1 = last outgoing edge of a node */
  const uint64_t* sampled_paths;       /* path_nodes bits */
  uint64_t sample_count;
  const uint64_t* stored_samples;      /* sample_count values (int_vector<0> unpacked) */
  const uint64_t* samples;             /* sample_count bits, 1 = last sample of a node */
  const uint64_t* extra_filter;        /* SadaSparse::filter, path_nodes bits (support.h:319) */
  uint64_t extra_values_len;
  const uint64_t* extra_values;        /* SadaSparse::values (support.h:323) */
  uint64_t redundant_len;
  const uint64_t* redundant;           /* SadaCount::data (support.h:252) */
} gcsa_flat_index;

/* gcsa::LCPArray members (include/gcsa/lcp.h:182-190): levels of a k-ary range-minimum tree,
   level 0 = the LCP array, concatenated; one byte per value (values are <= 255). */
typedef struct gcsa_flat_lcp {
  uint64_t size, branching, levels;
  const uint64_t* offsets;             /* levels + 1 */
  const uint8_t*  data;                /* offsets[levels] */
} gcsa_flat_lcp;

/* STNode, include/gcsa/lcp.h:40-79 */
typedef struct gcsa_b200_stnode { uint64_t sp, ep, left_lcp, right_lcp, node_lcp; } gcsa_b200_stnode;

typedef struct gcsa_b200_options {
  int      kmer_table_k;   /* 0 = none; else a lookup table of find() results for all ACGT strings of
                              this length is built at creation (4^k * 8 bytes) and used to skip the
                              first k backward steps.  -1 = engine default. */
  int      two_step;       /* 1 = also build (-1 = decide by index size, 0 = never) the two-step blocks (16 sectors per 87 path nodes, 5.9 B per
                              node): two backward steps per probe for pairs of ACGT characters. */
  int      walk_table;     /* locate(): precomputed per-node tables.  1 = build the locate table (8 bytes per path node:
                              the start position itself for nodes with one position, else the sampled node and the
                              step count, so that locating a node is one load), falling back to the walk table if
                              that fails; 2 = walk table only (4 bytes per node below 2^31 nodes, else 8: LF(node) or,
                              for sampled nodes, the index of their samples -- one load per LF step); 0 = neither
                              (bit vectors only); -1 = what fits: each table must fit in a fraction of the free
                              device memory. */
  int      jump_table;     /* find(): one 8-byte entry per path node holding the characters and the end of the unary
                              backward path starting there (up to 16 steps: every node on it has exactly one
                              predecessor character), so that a singleton range advances that many characters
                              with one load.  0 = build if it fits, 1 = build, -1 = do not.  Exact: a pattern that
                              leaves the path or ends inside it is continued with single steps.  An 8-byte entry
                              has room for 16 characters up to 2^27 path nodes (13 at 3 G nodes); beyond that the
                              long table gets 16-byte entries when they fit (2 = force 16-byte entries). */
  int      fused_table;    /* find(): k-mer table entries of 16 bytes instead of 8 -- next to the result of the k-mer, the
                              jump-table entry of its path node when the result is a single node, so that the table
                              lookup and the first jump are ONE load (a 32-mer with k = 16 is one 16-byte probe).
                              0 = when there is a jump table and 4^k * 16 bytes fit comfortably, 1 = always (needs
                              the jump table), -1 = never.  Exact: the same answers as the two separate loads. */
  int      reserved[3];
} gcsa_b200_options;

typedef struct gcsa_b200_info {
  uint64_t path_nodes, edge_count, order, sample_count;
  uint64_t device_bytes;               /* HBM used by the index */
  int      kmer_table_k;
  int      device;
  int      sm_count;
  int      two_step;                   /* 1 if the two-step blocks are in use */
  int      jump_k;                     /* longest path of the jump table (0 = no table) */
  int      fused_table;                /* 1 if the k-mer table holds fused 16-byte entries */
} gcsa_b200_info;

/* Per-batch statistics of find(): filled by gcsa_b200_find_stats_host (measurement only). */
typedef struct gcsa_b200_find_stats {
  uint64_t queries, found, total_length, lf_steps, sector_probes, table_hits;
} gcsa_b200_find_stats;

typedef struct gcsa_b200_index gcsa_b200_index;
typedef struct gcsa_b200_lcp   gcsa_b200_lcp;

const char* gcsa_b200_last_error(void);
const char* gcsa_b200_version(void);
int gcsa_b200_device_count(void);

/* Replaces GCSA::load() + the SDSL rank/select supports (src/gcsa.cpp:140-216, 726-738):
   uploads the arrays and builds the device rank dictionary.  options may be NULL. */
int  gcsa_b200_index_create(const gcsa_flat_index* host, int device, const gcsa_b200_options* options,
                            gcsa_b200_index** out);
void gcsa_b200_index_destroy(gcsa_b200_index* index);
int  gcsa_b200_index_info(const gcsa_b200_index* index, gcsa_b200_info* info);

/* GCSA::find(begin, end) / find(Container) / find(Element*, length), include/gcsa/gcsa.h:96-122.
   Pattern i is chars[offsets[i] .. offsets[i+1]); characters are raw bytes mapped through
   char2comp like the reference does (gcsa.h:102,106). */
int gcsa_b200_find_batch(const gcsa_b200_index* index, const uint8_t* d_chars, const uint64_t* d_offsets,
                         uint64_t n, uint64_t* d_sp, uint64_t* d_ep, void* stream);
int gcsa_b200_find_host(const gcsa_b200_index* index, const uint8_t* chars, const uint64_t* offsets,
                        uint64_t n, uint64_t* sp, uint64_t* ep);
/* The k-mer form of the same call: n patterns of one length stored back to back (pattern i is
   chars[i * pattern_length .. (i + 1) * pattern_length)), no offsets array to move. */
int gcsa_b200_find_fixed_batch(const gcsa_b200_index* index, const uint8_t* d_chars, uint64_t pattern_length,
                               uint64_t n, uint64_t* d_sp, uint64_t* d_ep, void* stream);
int gcsa_b200_find_fixed_host(const gcsa_b200_index* index, const uint8_t* chars, uint64_t pattern_length,
                              uint64_t n, uint64_t* sp, uint64_t* ep);
/* Same answers plus executed-work counters (slower; for the roofline accounting only). */
int gcsa_b200_find_stats_host(const gcsa_b200_index* index, const uint8_t* chars, const uint64_t* offsets,
                              uint64_t n, uint64_t* sp, uint64_t* ep, gcsa_b200_find_stats* stats);
/* The same for the k-mer form (counts what the kernels of gcsa_b200_find_fixed_* execute). */
int gcsa_b200_find_fixed_stats_host(const gcsa_b200_index* index, const uint8_t* chars, uint64_t pattern_length,
                                    uint64_t n, uint64_t* sp, uint64_t* ep, gcsa_b200_find_stats* stats);

/* GCSA::charRange(comp), include/gcsa/gcsa.h:150-153 (host-side, O(1), no device work). */
int gcsa_b200_char_range(const gcsa_b200_index* index, uint64_t comp, uint64_t* sp, uint64_t* ep);

/* GCSA::LF(range_type, comp_type), include/gcsa/gcsa.h:155-162. */
int gcsa_b200_lf_batch(const gcsa_b200_index* index, const uint64_t* d_sp, const uint64_t* d_ep,
                       const uint8_t* d_comp, uint64_t n, uint64_t* d_sp_out, uint64_t* d_ep_out, void* stream);
int gcsa_b200_lf_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep,
                      const uint8_t* comp, uint64_t n, uint64_t* sp_out, uint64_t* ep_out);

/* GCSA::LF(size_type path_node), include/gcsa/gcsa.h:165-183. */
int gcsa_b200_lf_node_batch(const gcsa_b200_index* index, const uint64_t* d_nodes, uint64_t n,
                            uint64_t* d_out, void* stream);
int gcsa_b200_lf_node_host(const gcsa_b200_index* index, const uint64_t* nodes, uint64_t n, uint64_t* out);

/* GCSA::LF_fast / LF_all, src/gcsa.cpp:742-798.  out holds sigma (sp, ep) pairs per input range,
   indexed by comp: out[(i * sigma + comp) * 2 + {0,1}]; slots the reference does not write are
   (1, 0) = Range::empty_range().  all_chars = 0: LF_fast (comps 1..fast_chars);
   all_chars = 1: LF_all (comps 1..sigma-2). */
int gcsa_b200_lf_multi_batch(const gcsa_b200_index* index, const uint64_t* d_sp, const uint64_t* d_ep,
                             uint64_t n, int all_chars, uint64_t* d_out, void* stream);
int gcsa_b200_lf_multi_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep,
                            uint64_t n, int all_chars, uint64_t* out);

/* GCSA::count(range_type), src/gcsa.cpp:802-809. */
int gcsa_b200_count_batch(const gcsa_b200_index* index, const uint64_t* d_sp, const uint64_t* d_ep,
                          uint64_t n, uint64_t* d_out, void* stream);
int gcsa_b200_count_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep,
                         uint64_t n, uint64_t* out);

/* GCSA::locate(range_type, results, append = false, sort = true), src/gcsa.cpp:827-842, as a CSR:
   the sorted distinct node_type values of range i are values[out_offsets[i] .. out_offsets[i+1]).
   *_host allocates *values with malloc (release with gcsa_b200_free).
   *_batch writes into caller-owned device buffers; if capacity is too small it returns
   GCSA_B200_ERR_CAPACITY after writing the needed size to *needed (a host pointer); the call
   synchronises `stream` once to learn the sizes. */
int gcsa_b200_locate_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                          uint64_t* out_offsets, uint64_t** values);
int gcsa_b200_locate_batch(const gcsa_b200_index* index, const uint64_t* d_sp, const uint64_t* d_ep, uint64_t n,
                           uint64_t* d_out_offsets, uint64_t* d_values, uint64_t capacity, uint64_t* needed,
                           void* stream);
/* The same CSR into caller-owned host buffers (use pinned memory: the copies then run at PCIe speed and are
   pipelined with the kernels chunk by chunk).  values holds `capacity` entries; *needed receives the number of
   values of the whole batch; GCSA_B200_ERR_CAPACITY if they did not fit (out_offsets is complete either way, so
   a second call with out_offsets[n] entries succeeds). */
int gcsa_b200_locate_into_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                               uint64_t* out_offsets, uint64_t* values, uint64_t capacity, uint64_t* needed);
/* The sort = false form of the same call (src/gcsa.cpp:840): every value locateInternal() emits,
   duplicates included, in the reference's order (path nodes ascending, samples in stored order). */
int gcsa_b200_locate_raw_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                              uint64_t* out_offsets, uint64_t** values);
/* GCSA::locate(range_type, max_positions, results), src/gcsa.cpp:844-878 (host buffers). */
int gcsa_b200_locate_max_host(const gcsa_b200_index* index, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                              uint64_t max_positions, uint64_t* out_offsets, uint64_t** values);
void gcsa_b200_free(void* p);

/* One process, several GPUs.  The reference's callers parallelise over independent queries with OpenMP threads
   inside one process (src/algorithms.cpp:113, 409; vg does the same); these are the entry points such a caller
   uses when the box has more than one GPU: `indexes` are `count` handles of the SAME index, one per device
   (gcsa_b200_index_create once per device); the batch is cut into `count` contiguous blocks, one host thread per
   handle drives its device, results land in place in the caller's arrays.  Same arguments and results as the
   single-handle calls; there is no exchange between the devices. */
int gcsa_b200_find_fixed_host_multi(const gcsa_b200_index* const* indexes, int count, const uint8_t* chars,
                                    uint64_t pattern_length, uint64_t n, uint64_t* sp, uint64_t* ep);
int gcsa_b200_find_host_multi(const gcsa_b200_index* const* indexes, int count, const uint8_t* chars,
                              const uint64_t* offsets, uint64_t n, uint64_t* sp, uint64_t* ep);
int gcsa_b200_locate_into_host_multi(const gcsa_b200_index* const* indexes, int count, const uint64_t* sp,
                                     const uint64_t* ep, uint64_t n, uint64_t* out_offsets, uint64_t* values,
                                     uint64_t capacity, uint64_t* needed);

/* countKMers(index, k, parameters), src/algorithms.cpp:387-421 (declared include/gcsa/algorithms.h:80-89):
   the number of distinct k-mers over the bases (include_Ns != 0: bases and N).  If ranges is not NULL it
   receives the path ranges of those k-mers (ordered by the reversed k-mer: the trie grows leftwards),
   malloc'ed as sp[0..count) followed by
   ep[0..count) (release with gcsa_b200_free). */
int gcsa_b200_count_kmers(const gcsa_b200_index* index, uint64_t k, int include_Ns, uint64_t* result, uint64_t** ranges);

/* KMerComparisonState, src/algorithms.cpp:425-460: the record compareKMers() writes to <output>.left /
   <output>.right -- both ranges, k, and the kmer packed 3 bits per comp in reading order from the right
   end (comp i of the backward walk at bits [3i, 3i+3)). */
typedef struct gcsa_b200_kmer_state { uint64_t left_sp, left_ep, right_sp, right_ep, k, kmer[3]; } gcsa_b200_kmer_state;

/* compareKMers(left, right, k, parameters), src/algorithms.cpp:535-616 (declared include/gcsa/algorithms.h:91-97):
   result[0..3) = kmers in both indexes, only in left, only in right; k <= 64.  If left_kmers / right_kmers are
   not NULL they receive the unique kmers (result[1] / result[2] records, malloc'ed, release with gcsa_b200_free;
   order unspecified -- the reference's depends on thread scheduling).  Both handles must be on one device.
   Like gcsa_b200_count_kmers this does not compare k with order() (parameters.force = true). */
int gcsa_b200_compare_kmers(const gcsa_b200_index* left, const gcsa_b200_index* right, uint64_t k, int include_Ns,
                            uint64_t* result, gcsa_b200_kmer_state** left_kmers, gcsa_b200_kmer_state** right_kmers);
/* The same with KMerSearchParameters::output set (algorithms.h:59-71): the unique kmers go to <output>.left and
   <output>.right as the reference writes them (raw 64-byte KMerComparisonState records, src/algorithms.cpp:606-607). */
int gcsa_b200_compare_kmers_to_files(const gcsa_b200_index* left, const gcsa_b200_index* right, uint64_t k, int include_Ns,
                                     const char* output, uint64_t* result);

/* verifyIndex(index, lcp, kmers, kmer_length), src/algorithms.cpp:101-295 (declared include/gcsa/algorithms.h:40-55),
   batched: every distinct kmer label of the construction input (keys / from = the KMer records,
   include/gcsa/support.h:475-497) is searched with find(); parent() must equal the first different range
   obtained by dropping characters from the right end and depth() must agree; count() must equal the number
   of distinct start nodes; locate() must return exactly those; locate(range, 10) must return min(10, n) of
   them (checked through gcsa_b200_locate_max_host for the labels with more than 10 start nodes: for the others
   locate(range, 10) is locate(range) by definition, src/gcsa.cpp:860-875).  Device-resident: the records are
   uploaded once and sorted, grouped, turned into patterns and compared on the GPU.  lcp may be NULL (the parent / depth checks are skipped, like `lcp == 0` in the reference).
   Returns 0 when the verification ran; the index is correct iff report->failures == 0. */
typedef struct gcsa_b200_verify_report {
  uint64_t unique;                      /* distinct labels queried */
  uint64_t failures;                    /* sum of the stage counters below */
  uint64_t find_failures, parent_failures, depth_failures, count_failures, locate_failures, random_locate_failures;
  double   seconds;                     /* wall clock of the whole verification */
  double   engine_seconds;              /* of which inside the engine's entry points (the rest is host work) */
} gcsa_b200_verify_report;
int gcsa_b200_verify_index(const gcsa_b200_index* index, const gcsa_b200_lcp* lcp, const uint64_t* keys,
                           const uint64_t* from, uint64_t n, int kmer_length, gcsa_b200_verify_report* report);
/* The same for an index built with a NodeMapping (the `mapping` argument of verifyIndex, algorithms.h:54): the
   expected occurrences are the mapped start nodes.  Arguments as for gcsa_b200_build_from_kmers_mapped. */
int gcsa_b200_verify_index_mapped(const gcsa_b200_index* index, const gcsa_b200_lcp* lcp, const uint64_t* keys,
                                  const uint64_t* from, uint64_t n, int kmer_length, uint64_t mapping_first_node,
                                  const uint64_t* mapping_ids, uint64_t mapping_size, gcsa_b200_verify_report* report);

/* LCPArray, include/gcsa/lcp.h:90-194; load() at src/lcp.cpp:116-143. */
int  gcsa_b200_lcp_create(const gcsa_flat_lcp* host, int device, gcsa_b200_lcp** out);
void gcsa_b200_lcp_destroy(gcsa_b200_lcp* lcp);
/* LCPArray::parent(range_type), src/lcp.cpp:276-301. */
int gcsa_b200_parent_batch(const gcsa_b200_lcp* lcp, const uint64_t* d_sp, const uint64_t* d_ep, uint64_t n,
                           gcsa_b200_stnode* d_out, void* stream);
int gcsa_b200_parent_host(const gcsa_b200_lcp* lcp, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                          gcsa_b200_stnode* out);
/* LCPArray::depth(range_type), src/lcp.cpp:319-325. */
int gcsa_b200_depth_batch(const gcsa_b200_lcp* lcp, const uint64_t* d_sp, const uint64_t* d_ep, uint64_t n,
                          uint64_t* d_out, void* stream);
int gcsa_b200_depth_host(const gcsa_b200_lcp* lcp, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                         uint64_t* out);
/* LCPArray::psv / psev / nsv / nsev, src/lcp.cpp:372-438: which = 0 psv, 1 psev, 2 nsv, 3 nsev.
   out_pos / out_val = the (res, LCP[res]) pair, or notFound() = (values, values). */
int gcsa_b200_lcp_sv_host(const gcsa_b200_lcp* lcp, int which, const uint64_t* pos, uint64_t n,
                          uint64_t* out_pos, uint64_t* out_val);
/* LCPArray::rmq(sp, ep), src/lcp.cpp:448-513. */
int gcsa_b200_lcp_rmq_host(const gcsa_b200_lcp* lcp, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                           uint64_t* out_pos, uint64_t* out_val);

/* MEM-style scan (BASELINE.json configs[4]): the driver loop over GCSA::LF (gcsa.h:155-162) and
   LCPArray::parent (src/lcp.cpp:276-301) in the scheme cited at paper/paper.tex:340, 606 -- extend the
   match to the left while possible, else report it and move to the suffix-tree parent.  The reference
   ships the two building blocks but no driver; this is the engine's.  Matches of pattern i are the
   4-tuples (start, length, sp, ep) at matches[4 * out_offsets[i] .. 4 * out_offsets[i + 1]).
   *_batch: capacity in matches; GCSA_B200_ERR_CAPACITY + *needed if too small (one stream sync). */
int gcsa_b200_mem_batch(const gcsa_b200_index* index, const gcsa_b200_lcp* lcp, const uint8_t* d_chars,
                        const uint64_t* d_offsets, uint64_t n, uint64_t* d_out_offsets, uint64_t* d_matches,
                        uint64_t capacity, uint64_t* needed, void* stream);
int gcsa_b200_mem_host(const gcsa_b200_index* index, const gcsa_b200_lcp* lcp, const uint8_t* chars,
                       const uint64_t* offsets, uint64_t n, uint64_t* out_offsets, uint64_t** matches);

/* ---------------------------------------------------------------------------------------------
   Host-side construction (CPU; gcsa2_b200/csrc/builder.cpp).  Replaces, for in-memory inputs,
   GCSA::GCSA(InputGraph&, ConstructionParameters) (src/gcsa.cpp:447-724) and
   LCPArray::LCPArray(InputGraph&) level 0 (src/lcp.cpp:204-272).
   --------------------------------------------------------------------------------------------- */
typedef struct gcsa_b200_built {
  gcsa_flat_index index;               /* arrays owned by the library: gcsa_b200_built_free */
  uint64_t lcp_size;
  uint8_t* lcp;                        /* LCP of adjacent path-node labels, in characters */
} gcsa_b200_built;

/* keys / from / to: the KMer records of the reference (include/gcsa/support.h:475-497):
   key = label (3 bits per character, first character most significant) << 16 | predecessor
   comp mask << 8 | successor comp mask; from / to = node_type; to = ~0 for kmers that are not
   extended.  Returns 0, or GCSA_B200_ERR_INCONSISTENT (arrays still returned) / other error;
   GCSA_B200_ERR_INVALID if the input has, or a doubling step would create, 2^32 - 1 paths or more (this in-memory
   builder numbers paths with 32 bits; gcsa_b200_build_linear handles 3 Gbp linear references). */
int  gcsa_b200_build_from_kmers(const uint64_t* keys, const uint64_t* from, const uint64_t* to, uint64_t n,
                                int kmer_length, int doubling_steps, uint64_t sample_period,
                                gcsa_b200_built* result);
/* The same with a NodeMapping (include/gcsa/support.h:167-222, InputGraph::mapping): node ids in
   [mapping_first_node, mapping_first_node + mapping_size) of the input graph are REPORTED as mapping_ids[id - first]
   -- the index is built for the input graph (with duplicated nodes, say) and locate() answers in the ids of the
   original graph.  As in the reference the mapping is applied after the final merge (src/gcsa.cpp:428-443, 600): it
   changes the samples and the counting structures, not the path nodes.  A .gcsa file carries the mapped values, so
   loading one needs no mapping.  mapping_size = 0: identity (gcsa_b200_build_from_kmers). */
int  gcsa_b200_build_from_kmers_mapped(const uint64_t* keys, const uint64_t* from, const uint64_t* to, uint64_t n,
                                       int kmer_length, int doubling_steps, uint64_t sample_period,
                                       uint64_t mapping_first_node, const uint64_t* mapping_ids, uint64_t mapping_size,
                                       gcsa_b200_built* result);
void gcsa_b200_built_free(gcsa_b200_built* result);

/* The same construction for a graph that is ONE PATH (# -> s[0] -> ... -> s[length-1] -> $), on the device
   (gcsa2_b200/csrc/linear_builder.cu): what GCSA::GCSA(InputGraph&, ...) (src/gcsa.cpp:447-724) produces for the
   kmers of a linear reference has a closed form -- the path nodes are the distinct length-K prefixes of the
   suffixes, K = kmer_length << doubling_steps -- so the index is built by radix-sorting suffixes instead of
   doubling paths.  sequence: comp values 1..5 (A C G T N), a host pointer or (sequence_on_device != 0) a device
   pointer on `device`.  Node ids as vg assigns them to a chopped path: the source is node 1, base i is node
   2 + i / node_length at offset i % node_length (node_length <= 1024), the sink is the next free id.
   Bit-identical to gcsa_b200_build_from_kmers on the kmers of that graph; length + 2 < 2^32 - 1.
   result as for gcsa_b200_build_from_kmers (release with gcsa_b200_built_free). */
int  gcsa_b200_build_linear(const uint8_t* sequence, uint64_t length, int sequence_on_device, uint64_t node_length,
                            int kmer_length, int doubling_steps, uint64_t sample_period, int device,
                            gcsa_b200_built* result);

/* ---------------------------------------------------------------------------------------------
   Index files of the reference (host; gcsa2_b200/csrc/gcsa_file.cpp).
   gcsa_b200_load_gcsa_file replaces GCSA::load / sdsl::load_from_file(index, name)
   (src/gcsa.cpp:182-216, src/build_gcsa.cpp:148-160): header check (tag 0x6C5A6C5A, version 3,
   flags 0; src/files.cpp:540-544), then the members in the reference's order, decoded to plain
   arrays; the SDSL rank/select supports in the file are skipped.  result->lcp stays NULL; release
   with gcsa_b200_built_free.  gcsa_b200_write_gcsa_file replaces GCSA::serialize
   (src/gcsa.cpp:140-180).  The *_lcp_* pair does the same for LCPArray::load / serialize
   (src/lcp.cpp:116-143; tag 0x6C5A7C94, version 1); release with gcsa_b200_flat_lcp_free.
   The reader validates sizes, cumulative counts and the end of file and returns
   GCSA_B200_ERR_INVALID with a message instead of guessing.
   --------------------------------------------------------------------------------------------- */
int  gcsa_b200_load_gcsa_file(const char* path, gcsa_b200_built* result);
int  gcsa_b200_write_gcsa_file(const gcsa_flat_index* index, const char* path);
int  gcsa_b200_load_lcp_file(const char* path, gcsa_flat_lcp* result);
int  gcsa_b200_write_lcp_file(const gcsa_flat_lcp* lcp, const char* path);
void gcsa_b200_flat_lcp_free(gcsa_flat_lcp* lcp);

/* A graph of single-character nodes in CSR form, for the synthetic inputs of the benchmarks. */
typedef struct gcsa_b200_graph {
  uint64_t nodes;
  const uint8_t*  comp;                /* comp value of each node's character */
  const uint64_t* value;               /* node_type of each node (include/gcsa/support.h:443-471) */
  const uint64_t* succ_offsets;        /* nodes + 1 */
  const uint64_t* succ;                /* successor node indexes */
  uint64_t sink;                       /* index of the '$' node */
  uint64_t n_sources;
  const uint64_t* sources;             /* nodes that follow the sink through the technical edge */
} gcsa_b200_graph;

typedef struct gcsa_b200_kmers { uint64_t n; uint64_t* key; uint64_t* from; uint64_t* to; } gcsa_b200_kmers;

/* The construction input of the reference from its own files (host; gcsa2_b200/csrc/kmer_file.cpp): kmer files as vg
   writes them -- binary .graph (sections of GraphFileHeader + KMer records, readBinary, src/files.cpp:127-167) or
   text .gcsa2 (five tab-separated columns, readText, src/files.cpp:86-124; char2comp = NULL: the default alphabet) --
   concatenated like InputGraph does for several files (src/files.cpp:308-345); *kmer_length receives their common kmer
   length.  Release with gcsa_b200_kmers_free.  gcsa_b200_load_node_mapping replaces NodeMapping::load
   (src/support.cpp:335-342) for build_gcsa's mapping file; release *ids with gcsa_b200_free. */
int  gcsa_b200_read_kmer_files(const char* const* paths, int count, int binary, const uint8_t* char2comp,
                               gcsa_b200_kmers* result, int* kmer_length);
int  gcsa_b200_load_node_mapping(const char* path, uint64_t* first_node, uint64_t** ids, uint64_t* size);

int  gcsa_b200_enumerate_kmers(const gcsa_b200_graph* graph, int kmer_length, gcsa_b200_kmers* result);
void gcsa_b200_kmers_free(gcsa_b200_kmers* result);
void gcsa_b200_default_char2comp(uint8_t* table256);

#ifdef __cplusplus
}
#endif
#endif /* GCSA2_B200_H */
