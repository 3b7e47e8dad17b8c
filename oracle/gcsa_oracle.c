/*
  gcsa_oracle.c -- CPU restatement of the GCSA2 query path.  TEST INFRASTRUCTURE ONLY
  (see gcsa_oracle.h for the rules and the parity-pinning statement).

  Citations are file:line in jltsiren/gcsa2 (the read-only reference).
*/
#include "gcsa_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <omp.h>

/* ------------------------------------------------------------------------------------------
   Range helpers -- include/gcsa/utils.h:84-117
   ------------------------------------------------------------------------------------------ */

static inline int range_empty(uint64_t sp, uint64_t ep) { return (sp + 1 > ep + 1); }   /* utils.h:93-101 */
static inline uint64_t range_length(uint64_t sp, uint64_t ep) { return ep + 1 - sp; }    /* utils.h:88-91 */

/* ------------------------------------------------------------------------------------------
   Bit vector with rank / select / access.
   Replaces sdsl::bit_vector_il<512>::rank_1_type, operator[] and bit_vector::select_1_type
   (include/gcsa/gcsa.h:46,219,227,231,236) and, semantically, sd_vector rank/select
   (gcsa.h:223; support.h:320,324).  [SDSL, not in tree: restated from the definitions.]
   ------------------------------------------------------------------------------------------ */

static void bv_init(oracle_bv* v, const uint64_t* words, uint64_t n_bits)
{
  v->n_bits = n_bits;
  v->n_blocks = n_bits / 512 + 1;          /* rank(n_bits) must be answerable */
  v->il = (uint64_t*)calloc(v->n_blocks * 9, sizeof(uint64_t));
  uint64_t n_words = (n_bits + 63) / 64, cum = 0;
  for(uint64_t b = 0; b < v->n_blocks; b++)
  {
    v->il[b * 9] = cum;
    for(uint64_t w = 0; w < 8; w++)
    {
      uint64_t idx = b * 8 + w, word = 0;
      if(idx < n_words)
      {
        word = (words != NULL ? words[idx] : 0);
        uint64_t rem = n_bits - idx * 64;
        if(rem < 64) { word &= (((uint64_t)1 << rem) - 1); }   /* ignore garbage past the end */
      }
      v->il[b * 9 + 1 + w] = word;
      cum += (uint64_t)__builtin_popcountll(word);
    }
  }
  v->ones = cum;
}

static void bv_clear(oracle_bv* v) { free(v->il); v->il = NULL; }

oracle_bv* oracle_bv_create(const uint64_t* words, uint64_t n_bits)
{
  oracle_bv* v = (oracle_bv*)calloc(1, sizeof(oracle_bv));
  bv_init(v, words, n_bits);
  return v;
}

void oracle_bv_destroy(oracle_bv* v) { if(v) { bv_clear(v); free(v); } }

/* rank1(i): number of ones in [0, i), valid for 0 <= i <= n_bits (paper/paper.tex:133-137). */
uint64_t oracle_bv_rank(const oracle_bv* v, uint64_t i)
{
  const uint64_t* blk = v->il + (i >> 9) * 9;
  uint64_t res = blk[0];
  uint64_t full = (i & 511) >> 6;
  for(uint64_t w = 0; w < full; w++) { res += (uint64_t)__builtin_popcountll(blk[1 + w]); }
  uint64_t rem = i & 63;
  if(rem) { res += (uint64_t)__builtin_popcountll(blk[1 + full] & (((uint64_t)1 << rem) - 1)); }
  return res;
}

int oracle_bv_get(const oracle_bv* v, uint64_t i)
{
  const uint64_t* blk = v->il + (i >> 9) * 9;
  return (int)((blk[1 + ((i & 511) >> 6)] >> (i & 63)) & 1);
}

/* select1(k): position of the k-th one, k >= 1 (k <= ones is the caller's job). */
uint64_t oracle_bv_select(const oracle_bv* v, uint64_t k)
{
  uint64_t lo = 0, hi = v->n_blocks - 1;       /* last block with cum < k */
  while(lo < hi)
  {
    uint64_t mid = lo + (hi - lo + 1) / 2;
    if(v->il[mid * 9] < k) { lo = mid; } else { hi = mid - 1; }
  }
  const uint64_t* blk = v->il + lo * 9;
  uint64_t need = k - blk[0];
  for(uint64_t w = 0; w < 8; w++)
  {
    uint64_t word = blk[1 + w];
    uint64_t c = (uint64_t)__builtin_popcountll(word);
    if(need <= c)
    {
      for(uint64_t j = 1; j < need; j++) { word &= word - 1; }
      return lo * 512 + w * 64 + (uint64_t)__builtin_ctzll(word);
    }
    need -= c;
  }
  return v->n_bits;   /* k > ones: not reached for valid input */
}

/* ------------------------------------------------------------------------------------------
   Index life cycle
   ------------------------------------------------------------------------------------------ */

oracle_gcsa* oracle_gcsa_create(const oracle_flat* f)
{
  oracle_gcsa* g = (oracle_gcsa*)calloc(1, sizeof(oracle_gcsa));
  g->path_nodes = f->path_nodes; g->edge_count = f->edge_count; g->order = f->order;
  g->sigma = f->sigma; g->fast_chars = f->fast_chars;
  memcpy(g->C, f->C, sizeof(g->C));
  memcpy(g->char2comp, f->char2comp, 256);
  for(int c = 0; c < ORACLE_SIGMA; c++) { bv_init(&g->bwt[c], f->bwt[c], f->path_nodes); }
  bv_init(&g->edges, f->edges, f->edge_count);
  bv_init(&g->sampled_paths, f->sampled_paths, f->path_nodes);
  bv_init(&g->samples, f->samples, f->sample_count);
  g->sample_count = f->sample_count;
  g->stored_samples = (uint64_t*)malloc((f->sample_count + 1) * sizeof(uint64_t));
  if(f->sample_count) { memcpy(g->stored_samples, f->stored_samples, f->sample_count * sizeof(uint64_t)); }
  bv_init(&g->extra_filter, f->extra_filter, f->path_nodes);
  bv_init(&g->extra_values, f->extra_values, f->extra_values_len);
  bv_init(&g->redundant, f->redundant, f->redundant_len);
  return g;
}

void oracle_gcsa_destroy(oracle_gcsa* g)
{
  if(!g) { return; }
  for(int c = 0; c < ORACLE_SIGMA; c++) { bv_clear(&g->bwt[c]); }
  bv_clear(&g->edges); bv_clear(&g->sampled_paths); bv_clear(&g->samples);
  bv_clear(&g->extra_filter); bv_clear(&g->extra_values); bv_clear(&g->redundant);
  free(g->stored_samples);
  free(g);
}

void oracle_free(void* p) { free(p); }

/* ------------------------------------------------------------------------------------------
   GCSA low-level interface
   ------------------------------------------------------------------------------------------ */

/* gcsa.h:262-266: alpha.C[comp] + rank[comp](i).  fast_rank (comp 1..fast_chars) and sparse_rank
   (comp 0, 5, 6) differ only in the encoding of the bit vector, not in the answer. */
static inline uint64_t lf_pos(const oracle_gcsa* g, uint64_t i, uint64_t comp)
{
  return g->C[comp] + oracle_bv_rank(&g->bwt[comp], i);
}

/* gcsa.h:253-258 pathNodeRange: map an outgoing-edge range to a path-node range. */
static inline void path_node_range(const oracle_gcsa* g, uint64_t* sp, uint64_t* ep)
{
  *sp = oracle_bv_rank(&g->edges, *sp);
  *ep = oracle_bv_rank(&g->edges, *ep);
}

/* gcsa.h:150-153 + utils.h:414-419: charRange(comp) = pathNodeRange(C[comp], C[comp+1]-1).
   The reference has no guard for C[comp+1] == 0 (the subtraction wraps and rank is then
   undefined); the oracle and the engine both define that case as the empty range (0, ~0). */
void oracle_char_range(const oracle_gcsa* g, uint64_t comp, uint64_t* sp, uint64_t* ep)
{
  if(comp >= g->sigma || g->C[comp + 1] == 0) { *sp = 0; *ep = ~(uint64_t)0; return; }
  *sp = g->C[comp]; *ep = g->C[comp + 1] - 1;
  path_node_range(g, sp, ep);
}

/* gcsa.h:155-162 LF(range, comp) with gcsa.h:268-274: the empty edge-space range is returned
   as it is (NOT canonicalised). */
void oracle_lf_range(const oracle_gcsa* g, uint64_t sp, uint64_t ep, uint64_t comp, uint64_t* osp, uint64_t* oep)
{
  uint64_t f = lf_pos(g, sp, comp);
  uint64_t s = lf_pos(g, ep + 1, comp) - 1;
  if(range_empty(f, s)) { *osp = f; *oep = s; return; }
  path_node_range(g, &f, &s);
  *osp = f; *oep = s;
}

/* gcsa.h:165-183 LF(path_node): follow the first edge backwards, fast characters first,
   then the sparse ones above fast_chars, finally comp 0. */
uint64_t oracle_lf_node(const oracle_gcsa* g, uint64_t i)
{
  for(uint64_t comp = 1; comp <= g->fast_chars; comp++)
  {
    if(oracle_bv_get(&g->bwt[comp], i)) { return oracle_bv_rank(&g->edges, lf_pos(g, i, comp)); }
  }
  for(uint64_t comp = g->fast_chars + 1; comp < g->sigma; comp++)
  {
    if(oracle_bv_get(&g->bwt[comp], i)) { return oracle_bv_rank(&g->edges, lf_pos(g, i, comp)); }
  }
  return oracle_bv_rank(&g->edges, lf_pos(g, i, 0));
}

/* src/gcsa.cpp:742-767 LF_fast: results[comp] for 1 <= comp <= fast_chars; others untouched
   (the caller's vector keeps its old contents there; we write (1,0) to all slots first and
   only slots 1..fast_chars are compared by the tests). */
void oracle_lf_fast(const oracle_gcsa* g, uint64_t sp, uint64_t ep, uint64_t* out)
{
  for(uint64_t c = 0; c < g->sigma; c++) { out[2 * c] = 1; out[2 * c + 1] = 0; }   /* Range::empty_range() utils.h:113-116 */
  if(range_empty(sp, ep)) { return; }
  if(sp == ep)
  {
    for(uint64_t comp = 1; comp <= g->fast_chars; comp++)
    {
      if(oracle_bv_get(&g->bwt[comp], sp))
      {
        out[2 * comp] = out[2 * comp + 1] = oracle_bv_rank(&g->edges, lf_pos(g, sp, comp));
      }
    }
  }
  else
  {
    for(uint64_t comp = 1; comp <= g->fast_chars; comp++)
    {
      uint64_t f = lf_pos(g, sp, comp), s = lf_pos(g, ep + 1, comp) - 1;
      if(!range_empty(f, s)) { path_node_range(g, &f, &s); }
      out[2 * comp] = f; out[2 * comp + 1] = s;
    }
  }
}

/* src/gcsa.cpp:769-798 LF_all: 1 <= comp < sigma - 1. */
void oracle_lf_all(const oracle_gcsa* g, uint64_t sp, uint64_t ep, uint64_t* out)
{
  for(uint64_t c = 0; c < g->sigma; c++) { out[2 * c] = 1; out[2 * c + 1] = 0; }
  if(range_empty(sp, ep)) { return; }
  if(sp == ep)
  {
    for(uint64_t comp = 1; comp + 1 < g->sigma; comp++)
    {
      if(oracle_bv_get(&g->bwt[comp], sp))
      {
        out[2 * comp] = out[2 * comp + 1] = oracle_bv_rank(&g->edges, lf_pos(g, sp, comp));
      }
    }
  }
  else
  {
    for(uint64_t comp = 1; comp + 1 < g->sigma; comp++)
    {
      oracle_lf_range(g, sp, ep, comp, &out[2 * comp], &out[2 * comp + 1]);
    }
  }
}

/* gcsa.h:96-110 find(begin, end). */
void oracle_find(const oracle_gcsa* g, const uint8_t* pattern, uint64_t len, uint64_t* sp, uint64_t* ep)
{
  if(len == 0 || g->path_nodes == 0) { *sp = 0; *ep = g->path_nodes - 1; return; }
  uint64_t pos = len - 1;
  uint64_t f, s;
  oracle_char_range(g, g->char2comp[pattern[pos]], &f, &s);
  while(!range_empty(f, s) && pos != 0)
  {
    pos--;
    oracle_lf_range(g, f, s, g->char2comp[pattern[pos]], &f, &s);
  }
  *sp = f; *ep = s;
}

void oracle_find_stats(const oracle_gcsa* g, const uint8_t* pattern, uint64_t len,
                       uint64_t* sp, uint64_t* ep, uint64_t* steps, uint64_t* probes)
{
  *steps = 0; *probes = 0;
  if(len == 0 || g->path_nodes == 0) { *sp = 0; *ep = g->path_nodes - 1; return; }
  uint64_t pos = len - 1;
  uint64_t f, s;
  uint64_t comp = g->char2comp[pattern[pos]];
  oracle_char_range(g, comp, &f, &s);
  if(comp < g->sigma && g->C[comp + 1] != 0)
  {
    *probes += ((g->C[comp] >> 9) == ((g->C[comp + 1] - 1) >> 9) ? 1 : 2);
  }
  while(!range_empty(f, s) && pos != 0)
  {
    pos--;
    comp = g->char2comp[pattern[pos]];
    *steps += 1;
    *probes += ((f >> 9) == ((s + 1) >> 9) ? 1 : 2);
    uint64_t ef = lf_pos(g, f, comp), es = lf_pos(g, s + 1, comp) - 1;
    if(range_empty(ef, es)) { f = ef; s = es; break; }
    *probes += ((ef >> 9) == (es >> 9) ? 1 : 2);
    f = oracle_bv_rank(&g->edges, ef); s = oracle_bv_rank(&g->edges, es);
  }
  *sp = f; *ep = s;
}

/* ------------------------------------------------------------------------------------------
   count() -- src/gcsa.cpp:802-809, support.h:255-258 (SadaCount), support.h:329-335 (SadaSparse)
   ------------------------------------------------------------------------------------------ */

static inline uint64_t sada_sparse_count(const oracle_gcsa* g, uint64_t sp, uint64_t ep)
{
  sp = oracle_bv_rank(&g->extra_filter, sp);
  ep = oracle_bv_rank(&g->extra_filter, ep + 1);
  if(ep <= sp) { return 0; }
  return (oracle_bv_select(&g->extra_values, ep) + 1) - (sp > 0 ? oracle_bv_select(&g->extra_values, sp) + 1 : 0);
}

static inline uint64_t sada_count(const oracle_gcsa* g, uint64_t sp, uint64_t ep)
{
  return (oracle_bv_select(&g->redundant, ep + 1) - ep) - (sp > 0 ? oracle_bv_select(&g->redundant, sp) + 1 - sp : 0);
}

uint64_t oracle_count(const oracle_gcsa* g, uint64_t sp, uint64_t ep)
{
  if(range_empty(sp, ep) || ep >= g->path_nodes) { return 0; }
  uint64_t res = sada_sparse_count(g, sp, ep) + range_length(sp, ep);
  if(ep > sp) { res -= sada_count(g, sp, ep - 1); }
  return res;
}

/* ------------------------------------------------------------------------------------------
   locate() -- src/gcsa.cpp:813-896
   ------------------------------------------------------------------------------------------ */

typedef struct { uint64_t* data; uint64_t size, cap; } u64vec;

static void vec_push(u64vec* v, uint64_t x)
{
  if(v->size == v->cap)
  {
    v->cap = (v->cap ? v->cap * 2 : 16);
    v->data = (uint64_t*)realloc(v->data, v->cap * sizeof(uint64_t));
  }
  v->data[v->size++] = x;
}

static int cmp_u64(const void* a, const void* b)
{
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return (x < y ? -1 : (x > y ? 1 : 0));
}

/* utils.h:350-357 removeDuplicates: sort + unique. */
static void remove_duplicates(u64vec* v)
{
  if(v->size == 0) { return; }
  qsort(v->data, v->size, sizeof(uint64_t), cmp_u64);
  uint64_t out = 1;
  for(uint64_t i = 1; i < v->size; i++)
  {
    if(v->data[i] != v->data[out - 1]) { v->data[out++] = v->data[i]; }
  }
  v->size = out;
}

/* gcsa.h:202-206 firstSample. */
static inline uint64_t first_sample(const oracle_gcsa* g, uint64_t path_node)
{
  uint64_t r = oracle_bv_rank(&g->sampled_paths, path_node);
  return (r > 0 ? oracle_bv_select(&g->samples, r) + 1 : 0);
}

/* src/gcsa.cpp:880-896 locateInternal. */
static void locate_internal(const oracle_gcsa* g, uint64_t path_node, u64vec* results)
{
  uint64_t steps = 0;
  while(!oracle_bv_get(&g->sampled_paths, path_node))       /* gcsa.h:191 sampled() */
  {
    path_node = oracle_lf_node(g, path_node);
    steps++;
  }
  uint64_t sample = first_sample(g, path_node);
  do
  {
    vec_push(results, g->stored_samples[sample] + steps); sample++;   /* gcsa.h:210 sample() */
  }
  while(!oracle_bv_get(&g->samples, sample - 1));                      /* gcsa.h:208 lastSample() */
}

/* src/gcsa.cpp:813-825 locate(path_node, results, append = false, sort = true). */
uint64_t oracle_locate_node(const oracle_gcsa* g, uint64_t path_node, uint64_t** out)
{
  u64vec v = { NULL, 0, 0 };
  if(path_node < g->path_nodes) { locate_internal(g, path_node, &v); }
  remove_duplicates(&v);
  *out = v.data;
  return v.size;
}

/* src/gcsa.cpp:827-842 locate(range, results, append = false, sort = true). */
static void locate_range_vec(const oracle_gcsa* g, uint64_t sp, uint64_t ep, u64vec* v)
{
  v->size = 0;
  if(range_empty(sp, ep) || ep >= g->path_nodes) { return; }
  for(uint64_t i = sp; i <= ep; i++) { locate_internal(g, i, v); }
  remove_duplicates(v);
}

uint64_t oracle_locate_range(const oracle_gcsa* g, uint64_t sp, uint64_t ep, uint64_t** out)
{
  u64vec v = { NULL, 0, 0 };
  locate_range_vec(g, sp, ep, &v);
  *out = v.data;
  return v.size;
}

/* std::mt19937_64 (ISO C++ [rand.predef]: w=64, n=312, m=156, r=31, a=0xB5026F5AA96619E9,
   u=29, d=0x5555555555555555, s=17, b=0x71D67FFFEDA60000, t=37, c=0xFFF7EEE000000000, l=43,
   f=6364136223846793005), used by src/gcsa.cpp:853. */
void oracle_mt64_seed(oracle_mt64* r, uint64_t seed)
{
  r->mt[0] = seed;
  for(int i = 1; i < 312; i++)
  {
    r->mt[i] = 6364136223846793005ULL * (r->mt[i - 1] ^ (r->mt[i - 1] >> 62)) + (uint64_t)i;
  }
  r->idx = 312;
}

uint64_t oracle_mt64_next(oracle_mt64* r)
{
  if(r->idx >= 312)
  {
    for(int i = 0; i < 312; i++)
    {
      uint64_t x = (r->mt[i] & 0xFFFFFFFF80000000ULL) | (r->mt[(i + 1) % 312] & 0x7FFFFFFFULL);
      uint64_t xa = x >> 1;
      if(x & 1) { xa ^= 0xB5026F5AA96619E9ULL; }
      r->mt[i] = r->mt[(i + 156) % 312] ^ xa;
    }
    r->idx = 0;
  }
  uint64_t y = r->mt[r->idx++];
  y ^= (y >> 29) & 0x5555555555555555ULL;
  y ^= (y << 17) & 0x71D67FFFEDA60000ULL;
  y ^= (y << 37) & 0xFFF7EEE000000000ULL;
  y ^= (y >> 43);
  return y;
}

/* utils.h:185-196 */
uint64_t oracle_wang_hash_64(uint64_t key)
{
  key = (~key) + (key << 21);
  key = key ^ (key >> 24);
  key = (key + (key << 3)) + (key << 8);
  key = key ^ (key >> 14);
  key = (key + (key << 2)) + (key << 4);
  key = key ^ (key >> 28);
  key = key + (key << 31);
  return key;
}

/* A small open-addressing set standing in for std::unordered_set<node_type> (gcsa.cpp:860);
   only membership and size matter because the results are re-sorted (gcsa.cpp:868-877). */
typedef struct { uint64_t* slots; uint8_t* used; uint64_t cap, size; } u64set;

static void set_init(u64set* s, uint64_t cap)
{
  s->cap = 16; while(s->cap < cap * 2) { s->cap *= 2; }
  s->slots = (uint64_t*)malloc(s->cap * sizeof(uint64_t));
  s->used = (uint8_t*)calloc(s->cap, 1);
  s->size = 0;
}

static void set_insert(u64set* s, uint64_t x);

static void set_grow(u64set* s)
{
  u64set n; set_init(&n, s->cap);
  for(uint64_t i = 0; i < s->cap; i++) { if(s->used[i]) { set_insert(&n, s->slots[i]); } }
  free(s->slots); free(s->used);
  *s = n;
}

static void set_insert(u64set* s, uint64_t x)
{
  if((s->size + 1) * 2 > s->cap) { set_grow(s); }
  uint64_t h = oracle_wang_hash_64(x) & (s->cap - 1);
  while(s->used[h])
  {
    if(s->slots[h] == x) { return; }
    h = (h + 1) & (s->cap - 1);
  }
  s->used[h] = 1; s->slots[h] = x; s->size++;
}

/* src/gcsa.cpp:844-878 locate(range, max_positions, results). */
uint64_t oracle_locate_max(const oracle_gcsa* g, uint64_t sp, uint64_t ep, uint64_t max_positions, uint64_t** out)
{
  u64vec results = { NULL, 0, 0 };
  *out = NULL;

  uint64_t total_positions = oracle_count(g, sp, ep);
  if(total_positions <= 0) { return 0; }
  if(max_positions > total_positions) { max_positions = total_positions; }

  oracle_mt64 rng; oracle_mt64_seed(&rng, sp ^ ep);
  if(max_positions >= total_positions / 2)
  {
    locate_range_vec(g, sp, ep, &results);
  }
  else
  {
    /* The reference loops until enough distinct values were seen; that never ends when count()
       overestimates the distinct values of a range that is not a suffix-tree node.  Oracle and
       engine both give up after 16 * length + 1024 draws and locate the whole range instead. */
    u64set found; set_init(&found, 16);
    uint64_t draws = 0, max_draws = 16 * range_length(sp, ep) + 1024;
    while(found.size < max_positions)
    {
      if(draws++ >= max_draws)
      {
        locate_range_vec(g, sp, ep, &results);
        for(uint64_t i = 0; i < results.size; i++) { set_insert(&found, results.data[i]); }
        results.size = 0;
        break;
      }
      uint64_t pos = sp + oracle_mt64_next(&rng) % range_length(sp, ep);
      locate_internal(g, pos, &results);
      for(uint64_t i = 0; i < results.size; i++) { set_insert(&found, results.data[i]); }
      results.size = 0;
    }
    for(uint64_t i = 0; i < found.cap; i++) { if(found.used[i]) { vec_push(&results, found.slots[i]); } }
    free(found.slots); free(found.used);
  }

  if(results.size > max_positions)
  {
    /* utils.h:359-370 deterministicShuffle: sort first, then swap from the back. */
    qsort(results.data, results.size, sizeof(uint64_t), cmp_u64);
    for(uint64_t i = results.size; i > 0; i--)
    {
      uint64_t j = oracle_mt64_next(&rng) % i;
      uint64_t tmp = results.data[i - 1]; results.data[i - 1] = results.data[j]; results.data[j] = tmp;
    }
    results.size = max_positions;
  }
  if(results.size) { qsort(results.data, results.size, sizeof(uint64_t), cmp_u64); }
  *out = results.data;
  return results.size;
}

/* ------------------------------------------------------------------------------------------
   Batch drivers (benchmark/query_gcsa.cpp:88-103, 152-167; src/algorithms.cpp:113)
   ------------------------------------------------------------------------------------------ */

int oracle_max_threads(void) { return omp_get_max_threads(); }

double oracle_find_batch(const oracle_gcsa* g, const uint8_t* chars, const uint64_t* offsets,
                         uint64_t n, uint64_t* sp, uint64_t* ep, int threads)
{
  if(threads < 1) { threads = 1; }
  double start = omp_get_wtime();
  #pragma omp parallel for schedule(dynamic, 4096) num_threads(threads)
  for(uint64_t i = 0; i < n; i++)
  {
    oracle_find(g, chars + offsets[i], offsets[i + 1] - offsets[i], &sp[i], &ep[i]);
  }
  return omp_get_wtime() - start;
}

double oracle_find_batch_stats(const oracle_gcsa* g, const uint8_t* chars, const uint64_t* offsets,
                         uint64_t n, uint64_t* sp, uint64_t* ep, int threads,
                         uint64_t* total_steps, uint64_t* total_probes)
{
  if(threads < 1) { threads = 1; }
  uint64_t ts = 0, tp = 0;
  double start = omp_get_wtime();
  #pragma omp parallel for schedule(dynamic, 4096) num_threads(threads) reduction(+:ts,tp)
  for(uint64_t i = 0; i < n; i++)
  {
    uint64_t st, pr;
    oracle_find_stats(g, chars + offsets[i], offsets[i + 1] - offsets[i], &sp[i], &ep[i], &st, &pr);
    ts += st; tp += pr;
  }
  *total_steps = ts; *total_probes = tp;
  return omp_get_wtime() - start;
}

double oracle_count_batch(const oracle_gcsa* g, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                          uint64_t* out, int threads)
{
  if(threads < 1) { threads = 1; }
  double start = omp_get_wtime();
  #pragma omp parallel for schedule(dynamic, 4096) num_threads(threads)
  for(uint64_t i = 0; i < n; i++) { out[i] = oracle_count(g, sp[i], ep[i]); }
  return omp_get_wtime() - start;
}

double oracle_locate_batch(const oracle_gcsa* g, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                           uint64_t* out_offsets, uint64_t** values, int threads)
{
  if(threads < 1) { threads = 1; }
  uint64_t** per = (uint64_t**)calloc(n + 1, sizeof(uint64_t*));
  uint64_t* cnt = (uint64_t*)calloc(n + 1, sizeof(uint64_t));
  double start = omp_get_wtime();
  #pragma omp parallel for schedule(dynamic, 256) num_threads(threads)
  for(uint64_t i = 0; i < n; i++) { cnt[i] = oracle_locate_range(g, sp[i], ep[i], &per[i]); }
  double secs = omp_get_wtime() - start;
  out_offsets[0] = 0;
  for(uint64_t i = 0; i < n; i++) { out_offsets[i + 1] = out_offsets[i] + cnt[i]; }
  uint64_t* vals = (uint64_t*)malloc((out_offsets[n] + 1) * sizeof(uint64_t));
  for(uint64_t i = 0; i < n; i++)
  {
    if(cnt[i]) { memcpy(vals + out_offsets[i], per[i], cnt[i] * sizeof(uint64_t)); }
    free(per[i]);
  }
  free(per); free(cnt);
  *values = vals;
  return secs;
}

/* ------------------------------------------------------------------------------------------
   countKMers -- src/algorithms.cpp:327-421 (KMerSearchState, KMerCounter, processSubtree)
   ------------------------------------------------------------------------------------------ */

typedef struct { uint64_t sp, ep, k; } kmer_state;

/* processSubtree (algorithms.cpp:364-385) with the KMerCounter / KMerSeedCollector handlers: report
   states at depth `depth`, expand shallower ones with LF_fast (bases) or LF_all (bases + N). */
static uint64_t process_subtree(const oracle_gcsa* g, kmer_state start, uint64_t depth, int include_Ns,
                                kmer_state** collect, uint64_t* collected, uint64_t* collect_cap)
{
  uint64_t count = 0, size = 0, cap = 64;
  kmer_state* stack = (kmer_state*)malloc(cap * sizeof(kmer_state));
  stack[size++] = start;
  uint64_t pred[2 * ORACLE_SIGMA];
  uint64_t limit = (include_Ns ? g->sigma : g->fast_chars + 2);
  while(size > 0)
  {
    kmer_state curr = stack[--size];
    if(range_empty(curr.sp, curr.ep)) { continue; }
    if(curr.k == depth)
    {
      count++;
      if(collect != NULL)
      {
        if(*collected == *collect_cap) { *collect_cap = (*collect_cap ? *collect_cap * 2 : 64); *collect = (kmer_state*)realloc(*collect, *collect_cap * sizeof(kmer_state)); }
        (*collect)[(*collected)++] = curr;
      }
    }
    if(curr.k < depth)
    {
      if(include_Ns) { oracle_lf_all(g, curr.sp, curr.ep, pred); } else { oracle_lf_fast(g, curr.sp, curr.ep, pred); }
      for(uint64_t comp = 1; comp + 1 < limit; comp++)
      {
        if(size == cap) { cap *= 2; stack = (kmer_state*)realloc(stack, cap * sizeof(kmer_state)); }
        kmer_state next = { pred[2 * comp], pred[2 * comp + 1], curr.k + 1 };
        stack[size++] = next;
      }
    }
  }
  free(stack);
  return count;
}

/* countKMers (algorithms.cpp:387-421): seeds of length min(k, 5) on one thread, then OpenMP over seeds. */
uint64_t oracle_count_kmers(const oracle_gcsa* g, uint64_t k, int include_Ns, int threads)
{
  if(k == 0) { return 1; }
  if(threads < 1) { threads = 1; }
  kmer_state* seeds = NULL; uint64_t n_seeds = 0, cap = 0;
  kmer_state root = { 0, g->path_nodes - 1, 0 };
  process_subtree(g, root, (k < 5 ? k : 5), include_Ns, &seeds, &n_seeds, &cap);
  uint64_t result = 0;
  #pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+:result)
  for(uint64_t i = 0; i < n_seeds; i++)
  {
    result += process_subtree(g, seeds[i], k, include_Ns, NULL, NULL, NULL);
  }
  free(seeds);
  return result;
}

/* ------------------------------------------------------------------------------------------
   compareKMers -- src/algorithms.cpp:425-616 (KMerComparisonState, KMerSymmetricDifference,
   processSubtrees): the trie of both indexes walked in lockstep; a state lives while either
   side is non-empty; at depth k it is shared, left-only or right-only.
   ------------------------------------------------------------------------------------------ */

/* KMerComparisonState::set, algorithms.cpp:451-457: comp i of the kmer (read right to left), 3 bits each */
static void cmp_set(uint64_t* kmer, uint64_t i, uint64_t comp)
{
  uint64_t offset = (i * 3) >> 6, bit = (i * 3) & 63;
  kmer[offset] |= comp << bit;
  if(bit > 61) { kmer[offset + 1] |= comp >> (64 - bit); }      /* the reference ORs 0 when nothing overflows */
}

typedef struct { oracle_kmer_cmp* left; oracle_kmer_cmp* right; uint64_t n_left, n_right, cap_left, cap_right; } cmp_lists;

static void cmp_push(oracle_kmer_cmp** list, uint64_t* n, uint64_t* cap, const oracle_kmer_cmp* state)
{
  if(*n == *cap) { *cap = (*cap ? *cap * 2 : 64); *list = (oracle_kmer_cmp*)realloc(*list, *cap * sizeof(oracle_kmer_cmp)); }
  (*list)[(*n)++] = *state;
}

/* processSubtrees (algorithms.cpp:505-533).  collect != NULL: KMerSeedCollector (report = keep the state);
   else KMerSymmetricDifference (algorithms.cpp:488-500) into counts[3] and, if lists != NULL, the unique kmers. */
static void process_subtrees(const oracle_gcsa* left, const oracle_gcsa* right, oracle_kmer_cmp start, uint64_t depth, int include_Ns,
                             oracle_kmer_cmp** collect, uint64_t* collected, uint64_t* collect_cap, uint64_t* counts, cmp_lists* lists)
{
  uint64_t size = 0, cap = 64;
  oracle_kmer_cmp* stack = (oracle_kmer_cmp*)malloc(cap * sizeof(oracle_kmer_cmp));
  stack[size++] = start;
  uint64_t lpred[2 * ORACLE_SIGMA], rpred[2 * ORACLE_SIGMA];
  uint64_t limit = (include_Ns ? left->sigma : left->fast_chars + 2);
  while(size > 0)
  {
    oracle_kmer_cmp curr = stack[--size];
    int lempty = range_empty(curr.left_sp, curr.left_ep), rempty = range_empty(curr.right_sp, curr.right_ep);
    if(lempty && rempty) { continue; }
    if(curr.k == depth)
    {
      if(collect != NULL) { cmp_push(collect, collected, collect_cap, &curr); }
      else
      {
        uint64_t llen = curr.left_ep + 1 - curr.left_sp, rlen = curr.right_ep + 1 - curr.right_sp;    /* Range::length */
        if(llen > 0 && rlen > 0) { counts[0]++; }
        else if(llen > 0 && rlen == 0) { counts[1]++; if(lists) { cmp_push(&lists->left, &lists->n_left, &lists->cap_left, &curr); } }
        else if(llen == 0 && rlen > 0) { counts[2]++; if(lists) { cmp_push(&lists->right, &lists->n_right, &lists->cap_right, &curr); } }
      }
    }
    if(curr.k < depth)
    {
      if(include_Ns) { oracle_lf_all(left, curr.left_sp, curr.left_ep, lpred); oracle_lf_all(right, curr.right_sp, curr.right_ep, rpred); }
      else { oracle_lf_fast(left, curr.left_sp, curr.left_ep, lpred); oracle_lf_fast(right, curr.right_sp, curr.right_ep, rpred); }
      for(uint64_t comp = 1; comp + 1 < limit; comp++)
      {
        if(size == cap) { cap *= 2; stack = (oracle_kmer_cmp*)realloc(stack, cap * sizeof(oracle_kmer_cmp)); }
        oracle_kmer_cmp next = curr;
        next.left_sp = lpred[2 * comp]; next.left_ep = lpred[2 * comp + 1];
        next.right_sp = rpred[2 * comp]; next.right_ep = rpred[2 * comp + 1];
        next.k = curr.k + 1;
        cmp_set(next.kmer, curr.k, comp);
        stack[size++] = next;
      }
    }
  }
  free(stack);
}

/* compareKMers (algorithms.cpp:535-616): result = (shared, left only, right only); the unique kmers
   (what the reference writes to <output>.left / .right) are returned malloc'ed if asked for. */
void oracle_compare_kmers(const oracle_gcsa* left, const oracle_gcsa* right, uint64_t k, int include_Ns, int threads,
                          uint64_t* result, oracle_kmer_cmp** left_kmers, oracle_kmer_cmp** right_kmers)
{
  result[0] = result[1] = result[2] = 0;
  if(left_kmers) { *left_kmers = NULL; }
  if(right_kmers) { *right_kmers = NULL; }
  if(k == 0) { result[0] = 1; return; }
  if(k > 64) { return; }                                           /* KMerComparisonState::MAX_K */
  if(threads < 1) { threads = 1; }
  oracle_kmer_cmp* seeds = NULL; uint64_t n_seeds = 0, cap = 0;
  oracle_kmer_cmp root; memset(&root, 0, sizeof(root));
  root.left_ep = left->path_nodes - 1; root.right_ep = right->path_nodes - 1;
  process_subtrees(left, right, root, (k < 5 ? k : 5), include_Ns, &seeds, &n_seeds, &cap, NULL, NULL);
  cmp_lists all; memset(&all, 0, sizeof(all));
  int want = (left_kmers != NULL || right_kmers != NULL);
  #pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for(uint64_t i = 0; i < n_seeds; i++)
  {
    uint64_t counts[3] = { 0, 0, 0 };
    cmp_lists mine; memset(&mine, 0, sizeof(mine));
    process_subtrees(left, right, seeds[i], k, include_Ns, NULL, NULL, NULL, counts, want ? &mine : NULL);
    #pragma omp critical
    {
      result[0] += counts[0]; result[1] += counts[1]; result[2] += counts[2];
      for(uint64_t j = 0; j < mine.n_left; j++) { cmp_push(&all.left, &all.n_left, &all.cap_left, &mine.left[j]); }
      for(uint64_t j = 0; j < mine.n_right; j++) { cmp_push(&all.right, &all.n_right, &all.cap_right, &mine.right[j]); }
    }
    free(mine.left); free(mine.right);
  }
  free(seeds);
  if(left_kmers) { *left_kmers = all.left; } else { free(all.left); }
  if(right_kmers) { *right_kmers = all.right; } else { free(all.right); }
}

/* ------------------------------------------------------------------------------------------
   LCPArray -- include/gcsa/lcp.h:137-178, src/lcp.cpp:152-200 (tree arithmetic), 276-519
   ------------------------------------------------------------------------------------------ */

oracle_lcp* oracle_lcp_create(uint64_t size, uint64_t branching, uint64_t levels,
                              const uint64_t* offsets, const uint8_t* data)
{
  oracle_lcp* l = (oracle_lcp*)calloc(1, sizeof(oracle_lcp));
  l->size = size; l->branching = branching; l->levels = levels;
  l->offsets = (uint64_t*)malloc((levels + 1) * sizeof(uint64_t));
  memcpy(l->offsets, offsets, (levels + 1) * sizeof(uint64_t));
  l->values = offsets[levels];
  l->data = (uint8_t*)malloc(l->values + 1);
  memcpy(l->data, data, l->values);
  return l;
}

void oracle_lcp_destroy(oracle_lcp* l) { if(l) { free(l->offsets); free(l->data); free(l); } }

static inline uint64_t rmt_root(const oracle_lcp* l) { return l->values - 1; }                          /* lcp.cpp:152-156 */
static inline uint64_t rmt_parent(const oracle_lcp* l, uint64_t node, uint64_t level)                    /* lcp.cpp:158-162 */
{ return l->offsets[level + 1] + (node - l->offsets[level]) / l->branching; }
static inline uint64_t rmt_first_sibling(const oracle_lcp* l, uint64_t node, uint64_t level)            /* lcp.cpp:170-174 */
{ return node - (node - l->offsets[level]) % l->branching; }
static inline uint64_t rmt_last_sibling(const oracle_lcp* l, uint64_t first_child, uint64_t level)      /* lcp.cpp:176-180 */
{ uint64_t a = l->offsets[level + 1], b = first_child + l->branching; return (a < b ? a : b) - 1; }
static inline uint64_t rmt_first_child(const oracle_lcp* l, uint64_t node, uint64_t level)              /* lcp.cpp:182-186 */
{ return l->offsets[level - 1] + (node - l->offsets[level]) * l->branching; }
static inline uint64_t rmt_last_child(const oracle_lcp* l, uint64_t node, uint64_t level)               /* lcp.cpp:188-192 */
{ return rmt_last_sibling(l, rmt_first_child(l, node, level), level - 1); }
static inline uint64_t rmt_level(const oracle_lcp* l, uint64_t node)                                     /* lcp.cpp:194-200 */
{ uint64_t level = 0; while(l->offsets[level + 1] <= node) { level++; } return level; }

typedef struct { uint64_t first, second; } pair64;
static inline pair64 not_found(const oracle_lcp* l) { pair64 p = { l->values, l->values }; return p; }  /* lcp.h:178 */
static inline int cmp_less(uint64_t a, uint64_t b, int or_equal) { return (or_equal ? a <= b : a < b); }

/* lcp.cpp:333-343: last value comp-smaller than val in [from, to). */
static pair64 psv_scan(const oracle_lcp* l, uint64_t from, uint64_t to, uint64_t val, int or_equal)
{
  while(to > from)
  {
    to--;
    if(cmp_less(l->data[to], val, or_equal)) { pair64 p = { to, l->data[to] }; return p; }
  }
  return not_found(l);
}

/* lcp.cpp:345-370 */
static pair64 psv_impl(const oracle_lcp* l, uint64_t to, int or_equal)
{
  if(to == 0 || to >= l->size) { return not_found(l); }
  uint64_t level = 0, val = l->data[to];
  pair64 res = not_found(l);
  while(to != rmt_root(l))
  {
    res = psv_scan(l, rmt_first_sibling(l, to, level), to, val, or_equal);
    if(res.first < l->values) { break; }
    to = rmt_parent(l, to, level); level++;
  }
  if(res.first >= l->values) { return res; }
  while(level > 0)
  {
    uint64_t from = rmt_first_child(l, res.first, level); level--;
    res = psv_scan(l, from, rmt_last_sibling(l, from, level) + 1, val, or_equal);
  }
  return res;
}

/* lcp.cpp:389-399: first value comp-smaller than val in [from, to]. */
static pair64 nsv_scan(const oracle_lcp* l, uint64_t from, uint64_t to, uint64_t val, int or_equal)
{
  for(uint64_t i = from; i <= to; i++)
  {
    if(cmp_less(l->data[i], val, or_equal)) { pair64 p = { i, l->data[i] }; return p; }
  }
  return not_found(l);
}

/* lcp.cpp:401-426 */
static pair64 nsv_impl(const oracle_lcp* l, uint64_t from, int or_equal)
{
  if(from + 1 >= l->size) { return not_found(l); }
  uint64_t level = 0, val = l->data[from];
  pair64 res = not_found(l);
  while(from != rmt_root(l))
  {
    res = nsv_scan(l, from + 1, rmt_last_sibling(l, from, level), val, or_equal);
    if(res.first < l->values) { break; }
    from = rmt_parent(l, from, level); level++;
  }
  if(res.first >= l->values) { return res; }
  while(level > 0)
  {
    from = rmt_first_child(l, res.first, level); level--;
    res = nsv_scan(l, from, rmt_last_sibling(l, from, level), val, or_equal);
  }
  return res;
}

void oracle_lcp_psv(const oracle_lcp* l, uint64_t pos, uint64_t* rpos, uint64_t* rval)
{ pair64 p = psv_impl(l, pos, 0); *rpos = p.first; *rval = p.second; }
void oracle_lcp_psev(const oracle_lcp* l, uint64_t pos, uint64_t* rpos, uint64_t* rval)
{ pair64 p = psv_impl(l, pos, 1); *rpos = p.first; *rval = p.second; }
void oracle_lcp_nsv(const oracle_lcp* l, uint64_t pos, uint64_t* rpos, uint64_t* rval)
{ pair64 p = nsv_impl(l, pos, 0); *rpos = p.first; *rval = p.second; }
void oracle_lcp_nsev(const oracle_lcp* l, uint64_t pos, uint64_t* rpos, uint64_t* rval)
{ pair64 p = nsv_impl(l, pos, 1); *rpos = p.first; *rval = p.second; }

static inline void update_res(const oracle_lcp* l, pair64* res, uint64_t i)                              /* lcp.cpp:442-446 */
{ if(l->data[i] < res->second) { res->first = i; res->second = l->data[i]; } }

/* lcp.cpp:448-513 rmq(sp, ep): leftmost minimum. */
static pair64 rmq_impl(const oracle_lcp* l, uint64_t sp, uint64_t ep)
{
  if(sp > ep || ep >= l->size) { return not_found(l); }
  if(sp == ep) { pair64 p = { sp, l->data[sp] }; return p; }

  pair64 res = { l->values, l->size };
  uint64_t level = 0, left = sp, right = ep;
  pair64* tail = NULL; uint64_t tail_size = 0, tail_cap = 0;
  while(1)
  {
    uint64_t left_par = rmt_parent(l, left, level), right_par = rmt_parent(l, right, level);
    if(left_par == right_par)
    {
      for(uint64_t i = left; i <= right; i++) { update_res(l, &res, i); }
      break;
    }

    uint64_t left_child = rmt_first_child(l, left_par, level + 1);
    if(left != left_child)
    {
      uint64_t last_child = rmt_last_sibling(l, left_child, level);
      for(uint64_t i = left; i <= last_child; i++) { update_res(l, &res, i); }
      left_par++;
    }

    uint64_t right_child = rmt_last_child(l, right_par, level + 1);
    if(right != right_child)
    {
      uint64_t first_child = rmt_first_sibling(l, right_child, level);
      for(uint64_t i = right; ; i--)
      {
        if(tail_size == tail_cap)
        {
          tail_cap = (tail_cap ? tail_cap * 2 : 64);
          tail = (pair64*)realloc(tail, tail_cap * sizeof(pair64));
        }
        tail[tail_size].first = i; tail[tail_size].second = l->data[i]; tail_size++;
        if(i == first_child) { break; }
      }
      right_par--;
    }

    if(left_par >= right_par)
    {
      if(left_par == right_par) { update_res(l, &res, left_par); }
      break;
    }
    left = left_par; right = right_par; level++;
  }

  while(tail_size > 0)
  {
    pair64 temp = tail[--tail_size];
    if(temp.second < res.second) { res = temp; }
  }
  free(tail);

  /* The reference reads past offsets[] when nothing was smaller than size() (only possible if
     an LCP value >= size()); stop instead of descending. */
  if(res.first >= l->values) { return res; }

  level = rmt_level(l, res.first);
  while(level > 0)
  {
    res.first = rmt_first_child(l, res.first, level); level--;
    while(l->data[res.first] != res.second) { res.first++; }
  }
  return res;
}

void oracle_lcp_rmq(const oracle_lcp* l, uint64_t sp, uint64_t ep, uint64_t* rpos, uint64_t* rval)
{ pair64 p = rmq_impl(l, sp, ep); *rpos = p.first; *rval = p.second; }

/* lcp.cpp:276-301 parent(range) via lcp.h:163-175 nodeFor(range); root at lcp.h:137. */
void oracle_lcp_parent(const oracle_lcp* l, uint64_t sp, uint64_t ep, oracle_stnode* out)
{
  uint64_t left_lcp = l->data[sp];
  uint64_t right_lcp = (ep + 1 < l->size ? l->data[ep + 1] : 0);
  if(sp == 0 && ep == l->size - 1)
  {
    out->sp = 0; out->ep = l->size - 1; out->left_lcp = 0; out->right_lcp = 0; out->node_lcp = 0;
    return;
  }
  uint64_t node_lcp = (left_lcp > right_lcp ? left_lcp : right_lcp);
  pair64 left = { sp, left_lcp }, right = { ep + 1, right_lcp }, nf = not_found(l);
  if(left_lcp == node_lcp)
  {
    left = psv_impl(l, sp, 0);
    if(left.first == nf.first && left.second == nf.second) { left.first = 0; left.second = 0; }
  }
  if(right_lcp == node_lcp)
  {
    right = nsv_impl(l, ep + 1, 0);
    if(right.first == nf.first && right.second == nf.second) { right.first = l->size; right.second = 0; }
  }
  out->sp = left.first; out->ep = right.first - 1;
  out->left_lcp = left.second; out->right_lcp = right.second; out->node_lcp = node_lcp;
}

/* lcp.cpp:319-325 depth(range). */
uint64_t oracle_lcp_depth(const oracle_lcp* l, uint64_t sp, uint64_t ep)
{
  if(range_length(sp, ep) <= 1) { return ORACLE_UNKNOWN; }
  pair64 res = rmq_impl(l, sp + 1, ep), nf = not_found(l);
  return ((res.first == nf.first && res.second == nf.second) ? ORACLE_UNKNOWN : res.second);
}

double oracle_parent_batch(const oracle_lcp* l, const uint64_t* sp, const uint64_t* ep, uint64_t n,
                           oracle_stnode* out, int threads)
{
  if(threads < 1) { threads = 1; }
  double start = omp_get_wtime();
  #pragma omp parallel for schedule(dynamic, 4096) num_threads(threads)
  for(uint64_t i = 0; i < n; i++) { oracle_lcp_parent(l, sp[i], ep[i], &out[i]); }
  return omp_get_wtime() - start;
}

/* ------------------------------------------------------------------------------------------
   MEM-style scan (BASELINE.json configs[4]).  The reference has no such driver: it supplies the
   building blocks, LF() and LCPArray::parent(), for the scheme of Ohlebusch et al. cited at
   paper/paper.tex:340, 606.  The loop below is this repository's definition of it; the device
   kernel mirrors it step by step.

     i = m; range = root; depth = 0; extended = false
     while i > 0:
       next = LF(range, comp(P[i-1]))                          (gcsa.h:155-162)
       if next is not empty:  range = next; depth++; i--; extended = true; continue
       if depth == 0:         i--; continue                    (character that occurs nowhere)
       if extended:           emit (i, depth, range); extended = false
       node = parent(range); range = node.range(); depth = node.lcp()      (lcp.cpp:276-301)
     if depth > 0 and extended: emit (0, depth, range)

   A match (start, length, sp, ep) says P[start, start + length) occurs with path range [sp, ep]
   and cannot be extended to the left.
   ------------------------------------------------------------------------------------------ */

static uint64_t mem_scan(const oracle_gcsa* g, const oracle_lcp* l, const uint8_t* pattern, uint64_t len, uint64_t* out)
{
  uint64_t emitted = 0;
  uint64_t i = len, sp = 0, ep = g->path_nodes - 1, depth = 0;
  int extended = 0;
  if(g->path_nodes == 0) { return 0; }
  while(i > 0)
  {
    uint64_t nsp, nep;
    oracle_lf_range(g, sp, ep, g->char2comp[pattern[i - 1]], &nsp, &nep);
    if(!range_empty(nsp, nep)) { sp = nsp; ep = nep; depth++; i--; extended = 1; continue; }
    if(depth == 0) { i--; continue; }
    if(extended)
    {
      if(out) { out[4 * emitted] = i; out[4 * emitted + 1] = depth; out[4 * emitted + 2] = sp; out[4 * emitted + 3] = ep; }
      emitted++; extended = 0;
    }
    oracle_stnode node;
    oracle_lcp_parent(l, sp, ep, &node);
    sp = node.sp; ep = node.ep; depth = node.node_lcp;
  }
  if(depth > 0 && extended)
  {
    if(out) { out[4 * emitted] = 0; out[4 * emitted + 1] = depth; out[4 * emitted + 2] = sp; out[4 * emitted + 3] = ep; }
    emitted++;
  }
  return emitted;
}

double oracle_mem_batch(const oracle_gcsa* g, const oracle_lcp* l, const uint8_t* chars, const uint64_t* offsets,
                        uint64_t n, uint64_t* out_offsets, uint64_t** matches, int threads)
{
  if(threads < 1) { threads = 1; }
  double start = omp_get_wtime();
  uint64_t* cnt = (uint64_t*)calloc(n + 1, sizeof(uint64_t));
  #pragma omp parallel for schedule(dynamic, 1024) num_threads(threads)
  for(uint64_t i = 0; i < n; i++) { cnt[i] = mem_scan(g, l, chars + offsets[i], offsets[i + 1] - offsets[i], NULL); }
  out_offsets[0] = 0;
  for(uint64_t i = 0; i < n; i++) { out_offsets[i + 1] = out_offsets[i] + cnt[i]; }
  uint64_t* vals = (uint64_t*)malloc((4 * out_offsets[n] + 1) * sizeof(uint64_t));
  #pragma omp parallel for schedule(dynamic, 1024) num_threads(threads)
  for(uint64_t i = 0; i < n; i++) { mem_scan(g, l, chars + offsets[i], offsets[i + 1] - offsets[i], vals + 4 * out_offsets[i]); }
  free(cnt);
  *matches = vals;
  return omp_get_wtime() - start;
}
