/*
  builder.cpp -- in-memory GCSA construction on the host (C++17 + OpenMP).

  Index construction stays on the CPU (BASELINE.json north_star).  This is not a port of the
  reference's disk-based constructor; it produces the same structure -- the maximally pruned
  order-K de Bruijn graph of the input kmers and the arrays of gcsa::GCSA / gcsa::LCPArray --
  by an in-memory prefix doubling over dense label ranks:

    kmers (key, from, to)                         reference input: include/gcsa/support.h:475-497
      -> paths with a sequence of kmer ranks      (src/path_graph.cpp:56-105)
      -> per doubling step: prune (merge the paths of a label that all start from one node,
         src/path_graph.cpp:892-948) and extend (join path [from,to) with every path starting
         at `to`, src/path_graph.cpp:953-1032)
      -> merge equal labels into nodes, then maximal subtrees whose labels share one set of
         start nodes (src/path_graph.cpp:577-609, 1154-1226)
      -> emit BWT bits, edges, samples, counting structures, LCP (src/gcsa.cpp:568-704).

  Differences in mechanism (not in result): labels are compared through dense ranks that are
  re-assigned after every step (one 64-bit key sort per step instead of a multi-file merge);
  the predecessor node of (node, char) is found from the LCP array (see assignEdges) instead
  of intersecting label ranges through a kmer de Bruijn graph (src/gcsa.cpp:355-388).

  C ABI at the bottom (declared in include/gcsa2_b200.h).
*/
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>
#include <parallel/algorithm>
#include <omp.h>

#include "../../include/gcsa2_b200.h"

namespace {

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t  u8;

constexpr u32 SORTED = ~(u32)0;     // PathNode::sorted(): to == ~0
constexpr int SIGMA = 7;
constexpr int SINK_COMP = 0;

struct Paths
{
  int M = 0;                          // ranks per label (1 << steps)
  std::vector<u32> from, to, rank;    // dense ids of start / continuation positions; label rank
  std::vector<u8>  preds, len;
  std::vector<u32> labels;            // size() * M kmer ranks (1-based, 0 = padding)
  size_t size() const { return from.size(); }
  void reserve(size_t n) { from.reserve(n); to.reserve(n); rank.reserve(n); preds.reserve(n); len.reserve(n); labels.reserve(n * M); }
  void resize(size_t n) { from.resize(n); to.resize(n); rank.resize(n); preds.resize(n); len.resize(n); labels.resize(n * M); }
};

struct KeyIdx { u64 key; u32 idx; };
inline bool operator<(const KeyIdx& a, const KeyIdx& b) { return (a.key < b.key) || (a.key == b.key && a.idx < b.idx); }

template<class It> void psort(It a, It b) { __gnu_parallel::sort(a, b); }

inline void setBit(std::vector<u64>& v, u64 i) { v[i >> 6] |= (u64)1 << (i & 63); }
inline size_t wordsFor(u64 bits) { return (size_t)((bits + 63) / 64) + 1; }

// Copy path src[i] to dst[j].
inline void copyPath(const Paths& src, size_t i, Paths& dst, size_t j)
{
  dst.from[j] = src.from[i]; dst.to[j] = src.to[i]; dst.rank[j] = src.rank[i];
  dst.preds[j] = src.preds[i]; dst.len[j] = src.len[i];
  std::memcpy(&dst.labels[j * dst.M], &src.labels[i * src.M], sizeof(u32) * src.M);
}

/*
  prune: within the paths of one label, if all start from the same node the label is unique to
  that node: keep one path, OR the predecessor sets, stop extending it
  (src/path_graph.cpp:899-911, 611-634).  Otherwise keep the distinct (from, to) pairs.
*/
void prune(Paths& paths)
{
  size_t n = paths.size();
  std::vector<KeyIdx> order(n);
  #pragma omp parallel for
  for(size_t i = 0; i < n; i++) { order[i].key = ((u64)paths.rank[i] << 32) | paths.from[i]; order[i].idx = (u32)i; }
  psort(order.begin(), order.end());

  // Reorder the paths by (label, from) with a parallel gather so that the scan below is sequential.
  {
    Paths sorted; sorted.M = paths.M; sorted.resize(n);
    #pragma omp parallel for
    for(size_t i = 0; i < n; i++) { copyPath(paths, order[i].idx, sorted, i); }
    paths = std::move(sorted);
  }
  std::vector<KeyIdx>().swap(order);

  Paths out; out.M = paths.M; out.reserve(n);
  std::vector<u32> tos;
  for(size_t i = 0; i < n; )
  {
    size_t j = i;
    u32 rank = paths.rank[i];
    while(j < n && paths.rank[j] == rank) { j++; }
    bool same_from = (paths.from[i] == paths.from[j - 1]);
    if(same_from)
    {
      u8 preds = 0;
      for(size_t t = i; t < j; t++) { preds |= paths.preds[t]; }
      size_t pos = out.size(); out.resize(pos + 1);
      copyPath(paths, i, out, pos);
      out.preds[pos] = preds; out.to[pos] = SORTED;
    }
    else
    {
      for(size_t a = i; a < j; )
      {
        size_t b = a;
        while(b < j && paths.from[b] == paths.from[a]) { b++; }
        tos.clear();
        u8 preds = 0;
        for(size_t t = a; t < b; t++) { tos.push_back(paths.to[t]); preds |= paths.preds[t]; }
        std::sort(tos.begin(), tos.end());
        tos.erase(std::unique(tos.begin(), tos.end()), tos.end());
        for(u32 t : tos)
        {
          size_t pos = out.size(); out.resize(pos + 1);
          copyPath(paths, a, out, pos);
          out.to[pos] = t; out.preds[pos] = preds;
        }
        a = b;
      }
    }
    i = j;
  }
  paths = std::move(out);
}

/*
  extend: a path that can still be extended is joined with every path starting at its `to`
  (src/path_graph.cpp:992-1003, PathNode(left, right) at :83-105); the result inherits the
  right part's `to`, so joining with a sorted path yields a sorted path.  New ranks come from
  one sort of (rank(left), rank(right)).
*/
// Path numbers and label ranks are 32-bit (KeyIdx::idx, Paths::rank): returns false if the step would create more
// paths than that (the reference's disk-based builder has no such limit; this in-memory one is for fixtures).
bool extend(Paths& paths, u32 positions)
{
  size_t n = paths.size();
  // Bucket the paths by start position.
  std::vector<u64> start(positions + 2, 0);
  for(size_t i = 0; i < n; i++) { start[paths.from[i] + 1]++; }
  for(size_t i = 0; i + 1 < start.size(); i++) { start[i + 1] += start[i]; }
  std::vector<u32> bucket(n);
  {
    std::vector<u64> cursor(start.begin(), start.end() - 1);
    for(size_t i = 0; i < n; i++) { bucket[cursor[paths.from[i]]++] = (u32)i; }
  }

  std::vector<u64> out_pos(n + 1, 0);
  #pragma omp parallel for
  for(size_t i = 0; i < n; i++)
  {
    out_pos[i + 1] = (paths.to[i] == SORTED ? 1 : start[paths.to[i] + 1] - start[paths.to[i]]);
  }
  for(size_t i = 0; i < n; i++) { out_pos[i + 1] += out_pos[i]; }
  size_t total = out_pos[n];
  if(total >= (size_t)SORTED) { return false; }

  Paths out; out.M = paths.M; out.resize(total);
  std::vector<KeyIdx> order(total);
  const int M = paths.M;
  #pragma omp parallel for schedule(dynamic, 65536)
  for(size_t i = 0; i < n; i++)
  {
    size_t pos = out_pos[i];
    if(paths.to[i] == SORTED)
    {
      copyPath(paths, i, out, pos);
      order[pos].key = (u64)paths.rank[i] << 32; order[pos].idx = (u32)pos;
      continue;
    }
    for(u64 b = start[paths.to[i]]; b < start[paths.to[i] + 1]; b++, pos++)
    {
      u32 q = bucket[b];
      out.from[pos] = paths.from[i]; out.to[pos] = paths.to[q];
      out.preds[pos] = paths.preds[i];
      int la = paths.len[i], lb = paths.len[q];
      if(la + lb > M) { lb = M - la; }
      out.len[pos] = (u8)(la + lb);
      u32* dst = &out.labels[pos * M];
      std::memcpy(dst, &paths.labels[i * M], sizeof(u32) * la);
      std::memcpy(dst + la, &paths.labels[(size_t)q * M], sizeof(u32) * lb);
      for(int t = la + lb; t < M; t++) { dst[t] = 0; }
      order[pos].key = ((u64)paths.rank[i] << 32) | paths.rank[q]; order[pos].idx = (u32)pos;
    }
  }
  paths = Paths();   // release

  psort(order.begin(), order.end());
  u32 rank = 0;
  for(size_t i = 0; i < total; i++)
  {
    if(i == 0 || order[i].key != order[i - 1].key) { rank++; }
    out.rank[order[i].idx] = rank;
  }
  paths = std::move(out);
  return true;
}

struct Builder
{
  int k = 0, steps = 0, M = 1, K = 0;
  u64 sample_period = 64;
  std::vector<u64> positions;         // sorted distinct node_type values (dense id -> value)
  std::vector<u64> kmer_labels;       // sorted distinct kmer labels (3 bits per char)

  // LCP of two kmer labels in characters.
  int kmerLcp(u64 a, u64 b) const
  {
    u64 x = a ^ b;
    if(x == 0) { return k; }
    int high = 63 - __builtin_clzll(x);        // highest differing bit
    int ch = high / 3;                         // character index from the end
    return k - 1 - ch;
  }

  int labelLcp(const u32* a, int la, const u32* b, int lb) const
  {
    int m = std::min(la, lb), j = 0;
    while(j < m && a[j] == b[j]) { j++; }
    if(j == m) { return std::min(K, m * k); }
    return j * k + kmerLcp(kmer_labels[a[j] - 1], kmer_labels[b[j] - 1]);
  }
};

struct Timer
{
  double t0 = omp_get_wtime(); bool on = (std::getenv("GCSA_B200_VERBOSE") != nullptr);
  void lap(const char* what) { if(on) { double t = omp_get_wtime(); std::fprintf(stderr, "[builder] %-28s %8.3f s\n", what, t - t0); t0 = t; } }
};

struct Output
{
  u64 path_nodes = 0, edge_count = 0, order = 0;
  u64 C[SIGMA + 1] = {};
  std::vector<u64> bwt[SIGMA];
  std::vector<u64> edges, sampled_paths, stored_samples, samples, extra_filter, extra_values, redundant;
  u64 sample_count = 0, extra_values_len = 0, redundant_len = 0;
  std::vector<u8> lcp;
  int status = 0;
};

/*
  Build everything.  Returns 0 on success, a negative GCSA_B200_ERR_* code otherwise.
*/
// NodeMapping, include/gcsa/support.h:167-222: node ids in [first, first + size) are reported as mapping[id - first].
struct Mapping
{
  u64 first = 0; const u64* ids = nullptr; u64 size = 0;
  bool empty() const { return size == 0; }
  // Node::map, src/support.cpp:604-612: the id of a node_type (value >> 11) is mapped, offset and orientation stay
  u64 operator()(u64 value) const
  {
    u64 id = value >> 11;
    if(id < first || id - first >= size) { return value; }
    return (ids[id - first] << 11) | (value & 0x7FF);
  }
};

int buildIndex(const u64* keys, const u64* from, const u64* to, u64 n, int k, int steps, u64 sample_period, const Mapping& mapping, Output& out)
{
  if(n == 0 || k < 1 || k > 16 || steps < 0 || steps > 4) { return GCSA_B200_ERR_INVALID; }
  if(n >= (u64)SORTED) { return GCSA_B200_ERR_INVALID; }            // 32-bit path numbers (see extend)
  Builder B; B.k = k; B.steps = steps; B.M = 1 << steps; B.K = k << steps; B.sample_period = (sample_period ? sample_period : 64);
  if(B.K > 255) { return GCSA_B200_ERR_INVALID; }    // LCP values are bytes (include/gcsa/support.h:44)
  const int M = B.M;

  Timer timer;
  // Dense ids for positions (from / to values).
  B.positions.assign(from, from + n);
  psort(B.positions.begin(), B.positions.end());
  B.positions.erase(std::unique(B.positions.begin(), B.positions.end()), B.positions.end());
  if(B.positions.size() >= (size_t)SORTED) { return GCSA_B200_ERR_INVALID; }
  auto posId = [&](u64 v) -> u32 {
    auto it = std::lower_bound(B.positions.begin(), B.positions.end(), v);
    return (it != B.positions.end() && *it == v ? (u32)(it - B.positions.begin()) : SORTED);
  };

  // Kmer ranks (Key::label = key >> 16, include/gcsa/support.h:376-403).
  B.kmer_labels.resize(n);
  #pragma omp parallel for
  for(u64 i = 0; i < n; i++) { B.kmer_labels[i] = keys[i] >> 16; }
  psort(B.kmer_labels.begin(), B.kmer_labels.end());
  B.kmer_labels.erase(std::unique(B.kmer_labels.begin(), B.kmer_labels.end()), B.kmer_labels.end());

  Paths paths; paths.M = M; paths.resize(n);
  int bad = 0;
  #pragma omp parallel for reduction(+:bad)
  for(u64 i = 0; i < n; i++)
  {
    u64 label = keys[i] >> 16;
    u32 r = (u32)(std::lower_bound(B.kmer_labels.begin(), B.kmer_labels.end(), label) - B.kmer_labels.begin()) + 1;
    paths.from[i] = posId(from[i]);
    // A kmer whose last character is the endmarker is not extended (src/files.cpp:272-282).
    bool ends = ((label & 7) == SINK_COMP);
    paths.to[i] = (ends || to[i] == ~(u64)0 ? SORTED : posId(to[i]));
    if(!ends && to[i] != ~(u64)0 && paths.to[i] == SORTED) { bad++; }   // continuation with no kmers
    paths.rank[i] = r; paths.preds[i] = (u8)((keys[i] >> 8) & 0xFF); paths.len[i] = 1;
    u32* lab = &paths.labels[i * M];
    lab[0] = r; for(int t = 1; t < M; t++) { lab[t] = 0; }
  }
  if(bad) { return GCSA_B200_ERR_INVALID; }

  timer.lap("kmer ranks");
  for(int step = 0; step < steps; step++)
  {
    prune(paths); timer.lap("prune");
    // Nothing left to extend: the remaining steps would only copy the paths.
    bool all_sorted = true;
    for(size_t i = 0; i < paths.size(); i++) { if(paths.to[i] != SORTED) { all_sorted = false; break; } }
    if(all_sorted) { break; }
    if(!extend(paths, (u32)B.positions.size())) { return GCSA_B200_ERR_INVALID; }
    timer.lap("extend");
  }

  // ---- merge: equal labels -> groups; maximal subtrees with one start set -> nodes ----
  size_t P = paths.size();
  std::vector<KeyIdx> order(P);
  #pragma omp parallel for
  for(size_t i = 0; i < P; i++) { order[i].key = ((u64)paths.rank[i] << 32) | paths.from[i]; order[i].idx = (u32)i; }
  psort(order.begin(), order.end());

  timer.lap("final sort");
  std::vector<u64> gstart;            // first index in `order` of each label group
  for(size_t i = 0; i < P; i++)
  {
    if(i == 0 || (order[i].key >> 32) != (order[i - 1].key >> 32)) { gstart.push_back(i); }
  }
  size_t G = gstart.size();
  gstart.push_back(P);

  // Distinct start positions per group (sorted: the sort key ends with from).
  std::vector<u64> gfrom_start(G + 1, 0);
  std::vector<u32> gfrom; gfrom.reserve(P);
  std::vector<u8>  gpreds(G, 0);
  for(size_t g = 0; g < G; g++)
  {
    gfrom_start[g] = gfrom.size();
    u8 preds = 0;
    for(u64 t = gstart[g]; t < gstart[g + 1]; t++)
    {
      u32 f = (u32)order[t].key;
      if(gfrom.size() == gfrom_start[g] || gfrom.back() != f) { gfrom.push_back(f); }
      preds |= paths.preds[order[t].idx];
    }
    gpreds[g] = preds;
  }
  gfrom_start[G] = gfrom.size();
  auto sameSet = [&](size_t a, size_t b) -> bool {
    u64 la = gfrom_start[a + 1] - gfrom_start[a], lb = gfrom_start[b + 1] - gfrom_start[b];
    if(la != lb) { return false; }
    return std::equal(gfrom.begin() + gfrom_start[a], gfrom.begin() + gfrom_start[a + 1], gfrom.begin() + gfrom_start[b]);
  };

  timer.lap("groups");
  // left_lcp[g]: LCP (characters) of label g-1 and label g; left_lcp[0] = left_lcp[G] = 0.
  std::vector<int> left_lcp(G + 1, 0);
  #pragma omp parallel for
  for(size_t g = 1; g < G; g++)
  {
    u32 a = order[gstart[g - 1]].idx, b = order[gstart[g]].idx;
    left_lcp[g] = B.labelLcp(&paths.labels[(size_t)a * M], paths.len[a], &paths.labels[(size_t)b * M], paths.len[b]);
  }

  timer.lap("lcp");
  // PathGraphMerger::extendRange (src/path_graph.cpp:577-609) over the label groups.
  std::vector<u64> nfirst, nlast;     // node -> first / last group
  for(size_t i = 0; i < G; )
  {
    size_t range_to = i;
    int range_left = left_lcp[i];
    int parent = std::min(B.K, (int)paths.len[order[gstart[i]].idx] * k);    // range_lcp of one label
    size_t curr = i + 1;
    while(curr < G)
    {
      if(!sameSet(curr, i)) { break; }
      parent = std::min(parent, left_lcp[curr]);
      if(parent <= range_left) { break; }
      int next_right = (curr + 1 < G ? left_lcp[curr + 1] : 0);
      if(next_right >= parent) { curr++; continue; }
      range_to = curr; curr++;
    }
    nfirst.push_back(i); nlast.push_back(range_to);
    i = range_to + 1;
  }
  size_t N = nfirst.size();

  /*
    The values of a node: the start positions of (any of) its labels, reported through the NodeMapping
    (MergedGraphReader::fromNodes, src/gcsa.cpp:428-443: map, then sort and remove duplicates).  The mapping is
    applied only here, after merging, as in the reference (support.h:178-183): everything emitted below --
    samples, SadaSparse / SadaCount counters -- is about mapped values.  value_ids are dense ids into `values`
    (sorted distinct mapped positions; without a mapping: B.positions and the ids of the group).
  */
  std::vector<u64> values;
  std::vector<u64> nv_start(N + 1, 0);
  std::vector<u32> nv_id;
  if(mapping.empty())
  {
    values = B.positions;
    nv_id.reserve(N);
    for(size_t v = 0; v < N; v++)
    {
      u64 g = nfirst[v];
      nv_start[v] = nv_id.size();
      nv_id.insert(nv_id.end(), gfrom.begin() + gfrom_start[g], gfrom.begin() + gfrom_start[g + 1]);
    }
    nv_start[N] = nv_id.size();
  }
  else
  {
    std::vector<u64> mapped(B.positions.size());
    #pragma omp parallel for
    for(size_t i = 0; i < mapped.size(); i++) { mapped[i] = mapping(B.positions[i]); }
    values = mapped;
    psort(values.begin(), values.end());
    values.erase(std::unique(values.begin(), values.end()), values.end());
    std::vector<u32> mapped_id(mapped.size());
    #pragma omp parallel for
    for(size_t i = 0; i < mapped.size(); i++) { mapped_id[i] = (u32)(std::lower_bound(values.begin(), values.end(), mapped[i]) - values.begin()); }
    std::vector<u32> tmp;
    for(size_t v = 0; v < N; v++)
    {
      u64 g = nfirst[v];
      nv_start[v] = nv_id.size();
      tmp.clear();
      for(u64 t = gfrom_start[g]; t < gfrom_start[g + 1]; t++) { tmp.push_back(mapped_id[gfrom[t]]); }
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      nv_id.insert(nv_id.end(), tmp.begin(), tmp.end());
    }
    nv_start[N] = nv_id.size();
  }

  out.path_nodes = N; out.order = B.K;
  out.lcp.assign(N, 0);
  for(size_t v = 0; v < N; v++) { out.lcp[v] = (u8)(v == 0 ? 0 : left_lcp[nfirst[v]]); }

  timer.lap("merge");
  // Per-node predecessor sets, first characters, values.
  std::vector<u8> npreds(N, 0), nchar(N, 0);
  #pragma omp parallel for
  for(size_t v = 0; v < N; v++)
  {
    u8 preds = 0;
    for(u64 g = nfirst[v]; g <= nlast[v]; g++) { preds |= gpreds[g]; }
    npreds[v] = preds;
    u32 first_rank = paths.labels[(size_t)order[gstart[nfirst[v]]].idx * M];
    nchar[v] = (u8)((B.kmer_labels[first_rank - 1] >> (3 * (k - 1))) & 7);
  }
  auto nodeValues = [&](size_t v, std::vector<u64>& res) {
    res.clear();
    for(u64 t = nv_start[v]; t < nv_start[v + 1]; t++) { res.push_back(values[nv_id[t]]); }
  };   // already sorted and distinct: ids are assigned in value order

  timer.lap("node info");
  // ---- BWT bits and C (src/gcsa.cpp:573-588, 666) ----
  for(int c = 0; c < SIGMA; c++) { out.bwt[c].assign(wordsFor(N), 0); }
  u64 counts[SIGMA] = {};
  for(size_t v = 0; v < N; v++)
  {
    for(int c = 0; c < SIGMA; c++) { if(npreds[v] & (1 << c)) { setBit(out.bwt[c], v); counts[c]++; } }
  }
  out.C[0] = 0;
  for(int c = 0; c < SIGMA; c++) { out.C[c + 1] = out.C[c] + counts[c]; }
  out.edge_count = out.C[SIGMA];

  /*
    assignEdges.  The edges labelled c, in order of their target node i, are the outgoing edges
    of the nodes starting with c, in node order (src/gcsa.cpp:576-587 finds the source by
    intersecting label ranges; the map is monotone: stay or advance by one).  Let mu be the LCP
    of c.X for two consecutive targets (1 + the minimum LCP between them, capped at K) and
    lambda the LCP between the current source j and the next node starting with c.  Both
    targets leave from j iff their common prefix is longer than anything j shares with its
    neighbour, i.e. iff mu > lambda; otherwise the second one belongs to the next source.
  */
  timer.lap("bwt");
  std::vector<u64> cfirst(SIGMA + 1, N);      // first node whose label starts with a character >= c
  for(size_t v = N; v-- > 0; ) { cfirst[nchar[v]] = v; }
  for(int c = SIGMA - 1; c >= 0; c--) { if(cfirst[c] == N) { cfirst[c] = cfirst[c + 1]; } }
  std::vector<u32> outdeg(N, 0);
  std::vector<u64> pred_of(N, ~(u64)0);          // source of the node's only edge (valid if indegree == 1)
  std::vector<u8>  indeg(N, 0);
  bool consistent = true;
  for(int c = 0; c < SIGMA; c++)
  {
    u64 j = cfirst[c], jend = cfirst[c + 1];
    int runmin = 1 << 30;
    bool first = true;
    for(size_t v = 0; v < N; v++)
    {
      if(v > 0) { runmin = std::min(runmin, (int)out.lcp[v]); }
      if(!(npreds[v] & (1 << c))) { continue; }
      if(j >= jend) { consistent = false; break; }
      if(!first)
      {
        int mu = std::min(B.K, 1 + runmin);
        if(j + 1 < jend && mu <= (int)out.lcp[j + 1]) { j++; }
      }
      first = false; runmin = 1 << 30;
      outdeg[j]++; indeg[v]++; pred_of[v] = j;
    }
  }
  for(size_t v = 0; v < N; v++) { if(outdeg[v] == 0) { consistent = false; } }
  if(!consistent) { out.status = GCSA_B200_ERR_INCONSISTENT; }

  out.edges.assign(wordsFor(out.edge_count), 0);
  {
    u64 total = 0;
    for(size_t v = 0; v < N; v++) { total += outdeg[v]; if(total > 0) { setBit(out.edges, total - 1); } }
  }

  timer.lap("edges");
  // ---- samples (src/gcsa.cpp:621-658) ----
  out.sampled_paths.assign(wordsFor(N), 0);
  std::vector<u64> cur, prev;
  std::vector<u64> sample_last;
  for(size_t v = 0; v < N; v++)
  {
    nodeValues(v, cur);
    bool sample = (indeg[v] > 1) || (npreds[v] & (1 << SINK_COMP));
    for(size_t t = 0; t < cur.size() && !sample; t++) { if(cur[t] % B.sample_period == 0) { sample = true; } }
    if(!sample)
    {
      if(pred_of[v] == ~(u64)0) { sample = true; }
      else
      {
        nodeValues(pred_of[v], prev);
        if(prev.size() != cur.size()) { sample = true; }
        else { for(size_t t = 0; t < cur.size(); t++) { if(cur[t] != prev[t] + 1) { sample = true; break; } } }
      }
    }
    if(sample)
    {
      setBit(out.sampled_paths, v);
      for(u64 x : cur) { out.stored_samples.push_back(x); }
      sample_last.push_back(out.stored_samples.size() - 1);
    }
  }
  out.sample_count = out.stored_samples.size();
  out.samples.assign(wordsFor(out.sample_count), 0);
  for(u64 p : sample_last) { setBit(out.samples, p); }

  timer.lap("samples");
  // ---- counting structures (src/gcsa.cpp:590-619, 668-672; support.h:264-279, 341-364) ----
  {
    std::vector<u32> redundant(N > 0 ? N - 1 : 0, 0);
    std::vector<u64> prev_occ(values.size(), 0);
    std::vector<u64> node_lcp, first_time, last_time;
    out.extra_filter.assign(wordsFor(N), 0);
    u64 occ_total = 0;
    std::vector<u64> extra_ones;
    for(size_t v = 0; v < N; v++)
    {
      u64 nvals = nv_start[v + 1] - nv_start[v];
      if(nvals > 1) { setBit(out.extra_filter, v); occ_total += nvals - 1; extra_ones.push_back(occ_total - 1); }
      u64 curr_lcp = (u64)out.lcp[v] + (v > 0 ? 1 : 0);
      while(!node_lcp.empty() && node_lcp.back() > curr_lcp) { node_lcp.pop_back(); first_time.pop_back(); last_time.pop_back(); }
      if(!node_lcp.empty() && node_lcp.back() == curr_lcp) { last_time.back() = v; }
      else { node_lcp.push_back(curr_lcp); first_time.push_back(v); last_time.push_back(v); }
      for(u64 t = nv_start[v]; t < nv_start[v + 1]; t++)
      {
        u32 id = nv_id[t];
        if(prev_occ[id] > 0)
        {
          size_t pos = std::lower_bound(last_time.begin(), last_time.end(), prev_occ[id]) - last_time.begin();
          redundant[first_time[pos] - 1]++;
        }
        prev_occ[id] = v + 1;
      }
    }
    out.extra_values_len = occ_total;
    out.extra_values.assign(wordsFor(occ_total), 0);
    for(u64 p : extra_ones) { setBit(out.extra_values, p); }
    u64 red_total = 0;
    for(u32 r : redundant) { red_total += r; }
    out.redundant_len = redundant.size() + red_total;
    out.redundant.assign(wordsFor(out.redundant_len), 0);
    u64 tail = 0;
    for(u32 r : redundant) { tail += (u64)r + 1; setBit(out.redundant, tail - 1); }
  }
  timer.lap("counting");
  return out.status;
}

//------------------------------------------------------------------------------

/*
  Kmer enumeration for a graph of single-character nodes (the role vg plays for the reference):
  one record per (walk of k characters, successor of its last node); walks that reach the sink
  are padded with the endmarker and not extended.
*/
struct KmerSink
{
  std::vector<u64> key, from, to;
};

void enumerate(const gcsa_b200_graph* g, int k, KmerSink& sink)
{
  u64 n = g->nodes;
  // predecessor masks
  std::vector<u8> pmask(n, 0);
  for(u64 u = 0; u < n; u++)
  {
    if(u == g->sink) { continue; }
    for(u64 e = g->succ_offsets[u]; e < g->succ_offsets[u + 1]; e++) { pmask[g->succ[e]] |= (u8)(1 << g->comp[u]); }
  }
  for(u64 s = 0; s < g->n_sources; s++) { pmask[g->sources[s]] |= (u8)(1 << SINK_COMP); }

  int threads = omp_get_max_threads();
  std::vector<KmerSink> local(threads);
  #pragma omp parallel
  {
    KmerSink& mine = local[omp_get_thread_num()];
    struct Frame { u64 node; u64 label; int depth; };
    std::vector<Frame> stack;
    #pragma omp for schedule(static)
    for(u64 v = 0; v < n; v++)
    {
      stack.clear();
      stack.push_back({ v, (u64)g->comp[v], 1 });
      while(!stack.empty())
      {
        Frame f = stack.back(); stack.pop_back();
        bool at_sink = (f.node == g->sink);
        if(f.depth == k || at_sink)
        {
          u64 label = f.label << (3 * (k - f.depth));     // pad with '$' (comp 0)
          u64 key = (label << 16) | ((u64)pmask[v] << 8);
          if(at_sink || f.depth < k)
          {
            // A kmer that reached the endmarker is followed by '$' again; the all-'$' kmer of the sink is
            // followed by '#' (the technical edge sink -> source).  Every kmer needs a successor
            // (include/gcsa/dbg.h:70-72).
            u64 succ = (label == 0 ? (u64)1 << 6 : (u64)1 << SINK_COMP);
            mine.key.push_back(key | succ); mine.from.push_back(g->value[v]); mine.to.push_back(~(u64)0);
          }
          else
          {
            for(u64 e = g->succ_offsets[f.node]; e < g->succ_offsets[f.node + 1]; e++)
            {
              mine.key.push_back(key | (u64)(1 << g->comp[g->succ[e]]));
              mine.from.push_back(g->value[v]); mine.to.push_back(g->value[g->succ[e]]);
            }
          }
          continue;
        }
        for(u64 e = g->succ_offsets[f.node]; e < g->succ_offsets[f.node + 1]; e++)
        {
          u64 w = g->succ[e];
          stack.push_back({ w, (f.label << 3) | g->comp[w], f.depth + 1 });
        }
      }
    }
  }
  size_t total = 0;
  for(auto& l : local) { total += l.key.size(); }
  sink.key.reserve(total); sink.from.reserve(total); sink.to.reserve(total);
  for(auto& l : local)
  {
    sink.key.insert(sink.key.end(), l.key.begin(), l.key.end());
    sink.from.insert(sink.from.end(), l.from.begin(), l.from.end());
    sink.to.insert(sink.to.end(), l.to.begin(), l.to.end());
    l = KmerSink();
  }
}

template<class T> T* release(std::vector<T>& v)
{
  T* p = (T*)std::malloc(sizeof(T) * (v.size() + 1));
  if(!v.empty()) { std::memcpy(p, v.data(), sizeof(T) * v.size()); }
  std::vector<T>().swap(v);
  return p;
}

} // namespace

//------------------------------------------------------------------------------

extern "C" {

int gcsa_b200_build_from_kmers(const uint64_t* keys, const uint64_t* from, const uint64_t* to, uint64_t n,
                               int kmer_length, int doubling_steps, uint64_t sample_period,
                               gcsa_b200_built* result)
{
  return gcsa_b200_build_from_kmers_mapped(keys, from, to, n, kmer_length, doubling_steps, sample_period, 0, nullptr, 0, result);
}

int gcsa_b200_build_from_kmers_mapped(const uint64_t* keys, const uint64_t* from, const uint64_t* to, uint64_t n,
                                      int kmer_length, int doubling_steps, uint64_t sample_period,
                                      uint64_t mapping_first_node, const uint64_t* mapping_ids, uint64_t mapping_size,
                                      gcsa_b200_built* result)
{
  if(result == nullptr || (mapping_size > 0 && mapping_ids == nullptr)) { return GCSA_B200_ERR_INVALID; }
  std::memset(result, 0, sizeof(*result));
  Output out;
  Mapping mapping; mapping.first = mapping_first_node; mapping.ids = mapping_ids; mapping.size = mapping_size;
  int status = buildIndex(keys, from, to, n, kmer_length, doubling_steps, sample_period, mapping, out);
  if(status != 0 && status != GCSA_B200_ERR_INCONSISTENT) { return status; }

  gcsa_flat_index& f = result->index;
  f.path_nodes = out.path_nodes; f.edge_count = out.edge_count; f.order = out.order;
  f.sigma = SIGMA; f.fast_chars = 4;
  for(int c = 0; c <= SIGMA; c++) { f.C[c] = out.C[c]; }
  gcsa_b200_default_char2comp(f.char2comp);
  for(int c = 0; c < SIGMA; c++) { f.bwt[c] = release(out.bwt[c]); }
  f.edges = release(out.edges);
  f.sampled_paths = release(out.sampled_paths);
  f.sample_count = out.sample_count;
  f.stored_samples = release(out.stored_samples);
  f.samples = release(out.samples);
  f.extra_filter = release(out.extra_filter);
  f.extra_values_len = out.extra_values_len; f.extra_values = release(out.extra_values);
  f.redundant_len = out.redundant_len; f.redundant = release(out.redundant);
  result->lcp_size = out.lcp.size();
  result->lcp = release(out.lcp);
  return status;
}

void gcsa_b200_built_free(gcsa_b200_built* result)
{
  if(result == nullptr) { return; }
  gcsa_flat_index& f = result->index;
  for(int c = 0; c < SIGMA; c++) { std::free((void*)f.bwt[c]); }
  std::free((void*)f.edges); std::free((void*)f.sampled_paths); std::free((void*)f.stored_samples);
  std::free((void*)f.samples); std::free((void*)f.extra_filter); std::free((void*)f.extra_values);
  std::free((void*)f.redundant); std::free((void*)result->lcp);
  std::memset(result, 0, sizeof(*result));
}

int gcsa_b200_enumerate_kmers(const gcsa_b200_graph* graph, int kmer_length, gcsa_b200_kmers* result)
{
  if(graph == nullptr || result == nullptr || kmer_length < 1 || kmer_length > 16) { return GCSA_B200_ERR_INVALID; }
  KmerSink sink;
  enumerate(graph, kmer_length, sink);
  result->n = sink.key.size();
  result->key = release(sink.key); result->from = release(sink.from); result->to = release(sink.to);
  return 0;
}

void gcsa_b200_kmers_free(gcsa_b200_kmers* result)
{
  if(result == nullptr) { return; }
  std::free(result->key); std::free(result->from); std::free(result->to);
  std::memset(result, 0, sizeof(*result));
}

void gcsa_b200_default_char2comp(uint8_t* table)
{
  // src/support.cpp:69-92
  std::memset(table, 5, 256);
  table[0] = 0; table[(unsigned char)'$'] = 0; table[(unsigned char)'#'] = 6;
  const char* acgt = "ACGT";
  for(int i = 0; i < 4; i++) { table[(unsigned char)acgt[i]] = (uint8_t)(i + 1); table[(unsigned char)(acgt[i] + 32)] = (uint8_t)(i + 1); }
}

} // extern "C"
