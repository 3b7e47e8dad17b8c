/*
  verify.cpp -- verifyIndex() of the reference (src/algorithms.cpp:101-295) as a batched driver.

  The reference queries the index with every distinct kmer label of the construction input, one
  OpenMP thread per chunk of labels, one query at a time.  Here the same six predicates are
  evaluated stage by stage over whole batches of labels through the C ABI, so every query runs in
  the CUDA engine: find -> parent / depth -> count -> locate -> locate(range, 10).  Host code only
  sorts the kmers, builds the patterns and compares the answers.

  Differences from the reference, on purpose: no NodeMapping (the identity mapping is assumed; the
  mapping only exists for indexes built with duplicated nodes, which are out of scope) and the
  failures are counted per stage instead of being printed.
*/
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>
#include <parallel/algorithm>

#include "../../include/gcsa2_b200.h"
#include "internal.h"

namespace
{

typedef uint64_t u64;
constexpr u64 RANDOM_LOCATE_SIZE = 10;        // algorithms.cpp:92
constexpr u64 CHUNK_LABELS = 4u << 20;        // labels per batch (bounds the host buffers)

inline bool rangeEmpty(u64 sp, u64 ep) { return (sp + 1 > ep + 1); }

// Wall-clock time spent inside the engine's entry points (the rest of verifyIndex is host work).
struct EngineClock
{
  double seconds = 0.0;
  template<class F> int operator()(F&& call)
  {
    auto t0 = std::chrono::steady_clock::now();
    int rc = call();
    seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return rc;
  }
};

struct Malloced
{
  u64* p = nullptr;
  ~Malloced() { gcsa_b200_free(p); }
};

} // namespace

extern "C" int gcsa_b200_verify_index(const gcsa_b200_index* index, const gcsa_b200_lcp* lcp, const uint64_t* keys,
                                      const uint64_t* from, uint64_t n, int kmer_length, gcsa_b200_verify_report* report)
{
  return gcsa_b200_verify_index_mapped(index, lcp, keys, from, n, kmer_length, 0, nullptr, 0, report);
}

extern "C" int gcsa_b200_verify_index_mapped(const gcsa_b200_index* index, const gcsa_b200_lcp* lcp, const uint64_t* keys,
                                             const uint64_t* from, uint64_t n, int kmer_length, uint64_t mapping_first_node,
                                             const uint64_t* mapping_ids, uint64_t mapping_size, gcsa_b200_verify_report* report)
{
  if(mapping_size > 0 && mapping_ids == nullptr)
  {
    gcsa_b200_internal_set_error("verify_index: null mapping");
    return GCSA_B200_ERR_INVALID;
  }
  // Node::map (src/support.cpp:604-612) of a start position: the expected occurrences are mapped ones (algorithms.cpp:185-187)
  auto mapped = [&](u64 value) -> u64
  {
    u64 id = value >> 11;
    if(mapping_size == 0 || id < mapping_first_node || id - mapping_first_node >= mapping_size) { return value; }
    return (mapping_ids[id - mapping_first_node] << 11) | (value & 0x7FF);
  };
  if(index == nullptr || report == nullptr || (n > 0 && (keys == nullptr || from == nullptr)) || kmer_length < 1 || kmer_length > 16)
  {
    gcsa_b200_internal_set_error("verify_index: bad argument");
    return GCSA_B200_ERR_INVALID;
  }
  std::memset(report, 0, sizeof(*report));
  auto started = std::chrono::steady_clock::now();
  EngineClock engine;
  // GCSA_B200_VERIFY_DEBUG: host and engine seconds per stage on stderr
  const bool debug = (std::getenv("GCSA_B200_VERIFY_DEBUG") != nullptr);
  auto lap_start = started; double lap_engine = 0.0;
  auto lap = [&](const char* what)
  {
    if(!debug) { return; }
    auto now = std::chrono::steady_clock::now();
    double total = std::chrono::duration<double>(now - lap_start).count();
    std::fprintf(stderr, "verify: %-28s host %.3f s, engine %.3f s\n", what, total - (engine.seconds - lap_engine), engine.seconds - lap_engine);
    lap_start = now; lap_engine = engine.seconds;
  };
  const u64 k = (u64)kmer_length;
  const char* comp2char = "$ACGTN#";           // src/support.cpp:92

  // Sort by (label, from): parallelQuickSort(kmers) + the label groups of algorithms.cpp:106-125.
  std::vector<std::pair<u64, u64>> recs(n);
  #pragma omp parallel for schedule(static)
  for(u64 i = 0; i < n; i++) { recs[i] = std::make_pair(keys[i] >> 16, mapped(from[i])); }     // Key::label, support.h:403
  __gnu_parallel::sort(recs.begin(), recs.end());
  std::vector<u64> group_start;
  for(u64 i = 0; i < n; i++) { if(i == 0 || recs[i].first != recs[i - 1].first) { group_start.push_back(i); } }
  group_start.push_back(n);
  const u64 unique = group_start.size() - 1;
  report->unique = unique;
  lap("sort + groups");

  for(u64 base = 0; base < unique; base += CHUNK_LABELS)
  {
    const u64 m = std::min(CHUNK_LABELS, unique - base);

    // Patterns: Key::decode (support.cpp:539-553), cut after the first endmarker (algorithms.cpp:127-129).
    std::vector<u64> offsets(m + 1, 0), exp_offsets(m + 1, 0);
    #pragma omp parallel for schedule(static)
    for(u64 g = 0; g < m; g++)
    {
      u64 label = recs[group_start[base + g]].first, len = 0;
      for(u64 i = 0; i < k; i++) { len++; if(((label >> (3 * (k - 1 - i))) & 7) == 0) { break; } }
      offsets[g + 1] = len;
      u64 distinct = 0;
      for(u64 j = group_start[base + g]; j < group_start[base + g + 1]; j++)
      {
        if(j == group_start[base + g] || recs[j].second != recs[j - 1].second) { distinct++; }
      }
      exp_offsets[g + 1] = distinct;
    }
    for(u64 g = 0; g < m; g++) { offsets[g + 1] += offsets[g]; exp_offsets[g + 1] += exp_offsets[g]; }
    std::vector<uint8_t> chars(offsets[m] + 1);
    // Expected occurrences: distinct `from` values per label (algorithms.cpp:184-187); sorted already.
    std::vector<u64> expected(exp_offsets[m]);
    #pragma omp parallel for schedule(static)
    for(u64 g = 0; g < m; g++)
    {
      u64 label = recs[group_start[base + g]].first;
      uint8_t* out = chars.data() + offsets[g];
      for(u64 i = 0; i < offsets[g + 1] - offsets[g]; i++)
      {
        u64 comp = (label >> (3 * (k - 1 - i))) & 7;
        out[i] = (uint8_t)comp2char[comp < 7 ? comp : 5];
      }
      u64* e = expected.data() + exp_offsets[g];
      for(u64 j = group_start[base + g]; j < group_start[base + g + 1]; j++)
      {
        if(j == group_start[base + g] || recs[j].second != recs[j - 1].second) { *e++ = recs[j].second; }
      }
    }

    lap("patterns + expected");
    // find() -- algorithms.cpp:131-143
    std::vector<u64> sp(m + 1), ep(m + 1);
    int rc = engine([&] { return gcsa_b200_find_host(index, chars.data(), offsets.data(), m, sp.data(), ep.data()); });
    if(rc != 0) { return rc; }
    std::vector<uint8_t> alive(m, 1);
    for(u64 g = 0; g < m; g++)
    {
      if(rangeEmpty(sp[g], ep[g])) { alive[g] = 0; report->find_failures++; }
    }

    lap("find");
    // parent() and depth() -- algorithms.cpp:145-181
    if(lcp != nullptr)
    {
      std::vector<u64> ids;
      for(u64 g = 0; g < m; g++) { if(alive[g]) { ids.push_back(g); } }
      std::vector<u64> a(ids.size() + 1), b(ids.size() + 1);
      for(u64 i = 0; i < ids.size(); i++) { a[i] = sp[ids[i]]; b[i] = ep[ids[i]]; }
      std::vector<gcsa_b200_stnode> parents(ids.size() + 1);
      rc = engine([&] { return gcsa_b200_parent_host(lcp, a.data(), b.data(), ids.size(), parents.data()); });
      if(rc != 0) { return rc; }

      // query_res: drop characters from the right end until the range changes, all labels of a round in one batch
      std::vector<u64> qsp(m), qep(m), qlen(m);
      std::vector<u64> todo = ids;
      for(u64 g : ids) { qsp[g] = sp[g]; qep[g] = ep[g]; qlen[g] = offsets[g + 1] - offsets[g]; }
      while(!todo.empty())
      {
        std::vector<uint8_t> sub; std::vector<u64> sub_offsets(1, 0);
        for(u64 g : todo)
        {
          qlen[g]--;
          sub.insert(sub.end(), chars.begin() + offsets[g], chars.begin() + offsets[g] + qlen[g]);
          sub_offsets.push_back(sub.size());
        }
        sub.push_back(0);
        std::vector<u64> s(todo.size() + 1), e(todo.size() + 1);
        rc = engine([&] { return gcsa_b200_find_host(index, sub.data(), sub_offsets.data(), todo.size(), s.data(), e.data()); });
        if(rc != 0) { return rc; }
        std::vector<u64> again;
        for(u64 i = 0; i < todo.size(); i++)
        {
          u64 g = todo[i];
          qsp[g] = s[i]; qep[g] = e[i];
          if(s[i] == sp[g] && e[i] == ep[g] && qlen[g] > 0) { again.push_back(g); }
        }
        todo.swap(again);
      }

      std::vector<u64> depth_ids;
      for(u64 i = 0; i < ids.size(); i++)
      {
        u64 g = ids[i];
        const gcsa_b200_stnode& p = parents[i];
        if(p.sp != qsp[g] || p.ep != qep[g] || p.node_lcp != qlen[g]) { alive[g] = 0; report->parent_failures++; }
        else { depth_ids.push_back(i); }
      }
      std::vector<u64> da(depth_ids.size() + 1), db(depth_ids.size() + 1), depth(depth_ids.size() + 1);
      for(u64 i = 0; i < depth_ids.size(); i++) { da[i] = parents[depth_ids[i]].sp; db[i] = parents[depth_ids[i]].ep; }
      rc = engine([&] { return gcsa_b200_depth_host(lcp, da.data(), db.data(), depth_ids.size(), depth.data()); });
      if(rc != 0) { return rc; }
      for(u64 i = 0; i < depth_ids.size(); i++)
      {
        if(depth[i] != parents[depth_ids[i]].node_lcp) { alive[ids[depth_ids[i]]] = 0; report->depth_failures++; }
      }
    }

    lap("parent + depth");
    // count() -- algorithms.cpp:183-200
    std::vector<u64> ids;
    for(u64 g = 0; g < m; g++) { if(alive[g]) { ids.push_back(g); } }
    std::vector<u64> a(ids.size() + 1), b(ids.size() + 1), counts(ids.size() + 1);
    for(u64 i = 0; i < ids.size(); i++) { a[i] = sp[ids[i]]; b[i] = ep[ids[i]]; }
    rc = engine([&] { return gcsa_b200_count_host(index, a.data(), b.data(), ids.size(), counts.data()); });
    if(rc != 0) { return rc; }
    {
      std::vector<u64> keep;
      for(u64 i = 0; i < ids.size(); i++)
      {
        u64 g = ids[i];
        if(counts[i] != exp_offsets[g + 1] - exp_offsets[g]) { alive[g] = 0; report->count_failures++; }
        else { keep.push_back(g); }
      }
      ids.swap(keep);
    }
    for(u64 i = 0; i < ids.size(); i++) { a[i] = sp[ids[i]]; b[i] = ep[ids[i]]; }

    lap("count");
    // locate() -- algorithms.cpp:202-234
    std::vector<u64> loc_offsets(ids.size() + 1, 0);
    Malloced located;
    rc = engine([&] { return gcsa_b200_locate_host(index, a.data(), b.data(), ids.size(), loc_offsets.data(), &located.p); });
    if(rc != 0) { return rc; }
    std::vector<u64> random_ids;                 // positions in ids whose locate() was right
    for(u64 i = 0; i < ids.size(); i++)
    {
      u64 g = ids[i], got = loc_offsets[i + 1] - loc_offsets[i], want = exp_offsets[g + 1] - exp_offsets[g];
      bool same = (got == want);
      if(same && want > 0) { same = std::equal(expected.begin() + exp_offsets[g], expected.begin() + exp_offsets[g + 1], located.p + loc_offsets[i]); }
      if(!same) { report->locate_failures++; }
      if(got == want) { random_ids.push_back(i); }     // the reference still tries the random locate after a value mismatch
    }

    lap("locate");
    // locate(range, 10) -- algorithms.cpp:236-274
    std::vector<u64> ra(random_ids.size() + 1), rb(random_ids.size() + 1), rnd_offsets(random_ids.size() + 1, 0);
    for(u64 i = 0; i < random_ids.size(); i++) { ra[i] = a[random_ids[i]]; rb[i] = b[random_ids[i]]; }
    Malloced randoms;
    rc = engine([&] { return gcsa_b200_locate_max_host(index, ra.data(), rb.data(), random_ids.size(), RANDOM_LOCATE_SIZE, rnd_offsets.data(), &randoms.p); });
    if(rc != 0) { return rc; }
    for(u64 i = 0; i < random_ids.size(); i++)
    {
      u64 q = random_ids[i];
      const u64* occs = located.p + loc_offsets[q]; u64 n_occs = loc_offsets[q + 1] - loc_offsets[q];
      const u64* rnd = randoms.p + rnd_offsets[i]; u64 n_rnd = rnd_offsets[i + 1] - rnd_offsets[i];
      if(n_rnd != std::min(RANDOM_LOCATE_SIZE, n_occs)) { report->random_locate_failures++; continue; }
      bool subset = true;
      for(u64 x = 0, y = 0; x < n_rnd; x++)
      {
        while(y + 1 < n_occs && occs[y] < rnd[x]) { y++; }
        if(y >= n_occs || rnd[x] != occs[y]) { subset = false; break; }
        y++;
      }
      if(!subset) { report->random_locate_failures++; }
    }
  }

  lap("locate(range, 10)");
  report->failures = report->find_failures + report->parent_failures + report->depth_failures + report->count_failures +
                     report->locate_failures + report->random_locate_failures;
  report->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - started).count();
  report->engine_seconds = engine.seconds;
  return GCSA_B200_OK;
}
