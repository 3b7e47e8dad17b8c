/*
  facade_test.cpp -- the C++ facade (include/gcsa2_b200.hpp) used the way a caller of gcsa::GCSA
  would use it: the loop of benchmark/query_gcsa.cpp:88-167 and the predicates of verifyIndex
  (src/algorithms.cpp:131-274) against a naive scan of the text.  Needs a GPU to run.
*/
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "gcsa2_b200.hpp"

using namespace gcsa_b200;

static std::uint64_t rng_state = 12345;
static std::uint64_t next_random() { rng_state = rng_state * 6364136223846793005ULL + 1442695040888963407ULL; return rng_state >> 33; }

#define REQUIRE(cond) do { if(!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while(0)

int main()
{
  const size_type L = 5000, node_len = 32;
  const char* acgt = "ACGT";
  std::string text(L, 'A');
  for(size_type i = 0; i < L; i++) { text[i] = acgt[next_random() % 4]; }
  for(size_type i = 100; i < 400; i++) { text[i + 1000] = text[i]; }      // a 300 bp repeat: multi-value nodes

  // graph: # text $ (nodes 0, 1..L, L+1)
  std::vector<std::uint8_t> comp(L + 2);
  std::vector<std::uint64_t> value(L + 2), succ_offsets(L + 3), succ(L + 1), sources(1, 0);
  std::uint8_t c2c[256]; gcsa_b200_default_char2comp(c2c);
  comp[0] = 6; comp[L + 1] = 0; value[0] = Node::encode(1, 0); value[L + 1] = Node::encode(2 + (L + node_len - 1) / node_len, 0);
  for(size_type p = 0; p < L; p++) { comp[p + 1] = c2c[(unsigned char)text[p]]; value[p + 1] = Node::encode(2 + p / node_len, p % node_len); }
  for(size_type i = 0; i <= L + 2; i++) { succ_offsets[i] = (i < L + 1 ? i : L + 1); }
  for(size_type i = 0; i <= L; i++) { succ[i] = i + 1; }
  gcsa_b200_graph graph = { L + 2, comp.data(), value.data(), succ_offsets.data(), succ.data(), L + 1, 1, sources.data() };

  gcsa_b200_kmers kmers = {};
  REQUIRE(gcsa_b200_enumerate_kmers(&graph, 8, &kmers) == 0);
  gcsa_b200_built built = {};
  REQUIRE(gcsa_b200_build_from_kmers(kmers.key, kmers.from, kmers.to, kmers.n, 8, 2, 64, &built) == 0);
  REQUIRE(built.index.order == 32);

  // LCP tree (src/lcp.cpp:224-258), branching 4
  const size_type branching = 4;
  std::vector<std::uint64_t> offsets(1, 0);
  std::vector<std::uint8_t> data(built.lcp, built.lcp + built.lcp_size);
  size_type level_size = built.lcp_size, level_start = 0;
  offsets.push_back(level_size);
  while(level_size > 1)
  {
    size_type next_size = (level_size + branching - 1) / branching;
    for(size_type i = 0; i < next_size; i++)
    {
      std::uint8_t m = 255;
      for(size_type j = i * branching; j < std::min(level_size, (i + 1) * branching); j++) { m = std::min(m, data[level_start + j]); }
      data.push_back(m);
    }
    level_start += level_size; level_size = next_size; offsets.push_back(offsets.back() + level_size);
  }
  gcsa_flat_lcp flat_lcp = { built.lcp_size, branching, offsets.size() - 1, offsets.data(), data.data() };

  GCSA index(built.index, 0, 4);
  LCPArray lcp(flat_lcp, 0);
  REQUIRE(index.size() == built.index.path_nodes && index.order() == 32);
  REQUIRE(index.find(std::string("")) == range_type(0, index.size() - 1));

  std::vector<std::string> patterns;
  for(int i = 0; i < 300; i++)
  {
    size_type len = 1 + next_random() % 32, start = next_random() % (L - len);
    std::string p = text.substr(start, len);
    if(i % 5 == 0) { p[next_random() % len] = acgt[next_random() % 4]; }
    patterns.push_back(p);
  }
  patterns.push_back(text.substr(150, 32));          // inside the repeat: two occurrences

  std::vector<range_type> batch;
  index.find(patterns, batch);

  // one process, several replicas (here two handles on device 0): the same answers, results in place
  {
    GCSA second(built.index, 0, 4);
    std::vector<const GCSA*> replicas = { &index, &second };
    std::vector<range_type> from_replicas;
    gcsa_b200::find(replicas, patterns, from_replicas);
    REQUIRE(from_replicas == batch);
    std::vector<size_type> offsets_one, offsets_two;
    std::vector<node_type> values_one, values_two;
    index.locate(batch, offsets_one, values_one);
    gcsa_b200::locate(replicas, batch, offsets_two, values_two);
    REQUIRE(offsets_one == offsets_two && values_one == values_two);
  }
  size_type found = 0;
  for(size_type i = 0; i < patterns.size(); i++)
  {
    const std::string& p = patterns[i];
    range_type range = index.find(p);                                    // query_gcsa.cpp:94
    REQUIRE(range == batch[i]);
    REQUIRE(range == index.find(p.c_str(), p.size()));
    std::vector<node_type> expected;
    for(size_type pos = text.find(p); pos != std::string::npos; pos = text.find(p, pos + 1)) { expected.push_back(value[pos + 1]); }
    std::sort(expected.begin(), expected.end());
    REQUIRE(Range::empty(range) == expected.empty());
    if(Range::empty(range)) { REQUIRE(index.count(range) == 0); continue; }
    found++;
    REQUIRE(index.count(range) == expected.size());                      // algorithms.cpp:183-200
    std::vector<node_type> occs;
    index.locate(range, occs);                                           // algorithms.cpp:202-234
    REQUIRE(occs == expected);
    std::vector<node_type> some;
    index.locate(range, 10, some);                                       // algorithms.cpp:236-274
    REQUIRE(some.size() == std::min<size_type>(10, occs.size()));
    for(node_type x : some) { REQUIRE(std::binary_search(occs.begin(), occs.end(), x)); }
    // append = true, sort = false keeps what was there and adds raw values (gcsa.cpp:827-842)
    std::vector<node_type> appended(1, 7);
    index.locate(range, appended, true, false);
    REQUIRE(appended[0] == 7 && appended.size() >= 1 + occs.size());
    index.locate(range, appended, true, true);
    REQUIRE(appended.size() == occs.size() + (std::binary_search(occs.begin(), occs.end(), (node_type)7) ? 0 : 1));
    // parent / depth (algorithms.cpp:146-181)
    if(range != range_type(0, index.size() - 1))
    {
      STNode parent = lcp.parent(range);
      range_type query = range; size_type end = p.size();
      while(query == range) { end--; query = index.find(p.begin(), p.begin() + end); }
      REQUIRE(parent == query && parent.lcp() == end);
      REQUIRE(lcp.depth(parent.range()) == parent.lcp());
    }
    // LF: one more character to the left equals find of the longer pattern
    for(comp_type c = 1; c <= 4; c++)
    {
      std::string longer = std::string(1, acgt[c - 1]) + p;
      if(longer.size() <= 32) { REQUIRE(index.LF(range, c) == index.find(longer)); }
    }
  }
  REQUIRE(found > 200);

  // The same index through the reference's file formats: GCSA::serialize / load, LCPArray::serialize / load
  // (src/gcsa.cpp:140-216, src/lcp.cpp:116-143; build_gcsa.cpp:148-178).
  {
    std::string base = std::string("/tmp/gcsa2_b200_facade_test_") + std::to_string((unsigned long)next_random());
    REQUIRE(gcsa_b200_write_gcsa_file(&built.index, (base + ".gcsa").c_str()) == 0);
    REQUIRE(gcsa_b200_write_lcp_file(&flat_lcp, (base + ".lcp").c_str()) == 0);
    GCSA from_file(base + ".gcsa", 0, 4);
    LCPArray lcp_from_file(base + ".lcp", 0);
    REQUIRE(from_file.size() == index.size() && from_file.order() == index.order() && lcp_from_file.size() == lcp.size());
    std::vector<range_type> again;
    from_file.find(patterns, again);
    REQUIRE(again == batch);
    for(size_type i = 0; i < 50; i++)
    {
      if(Range::empty(batch[i])) { continue; }
      std::vector<node_type> a, b;
      index.locate(batch[i], a); from_file.locate(batch[i], b);
      REQUIRE(a == b && from_file.count(batch[i]) == index.count(batch[i]));
      REQUIRE(lcp_from_file.parent(batch[i]) == lcp.parent(batch[i]));
    }
    std::remove((base + ".gcsa").c_str()); std::remove((base + ".lcp").c_str());
    bool threw = false;
    try { GCSA missing(base + ".gcsa", 0, 4); } catch(const std::runtime_error&) { threw = true; }
    REQUIRE(threw);                                                        // GCSA::load throws, gcsa.cpp:188-193
  }
  std::vector<range_type> fast(GCSA_B200_SIGMA), all(GCSA_B200_SIGMA);
  index.LF_fast(batch[0], fast); index.LF_all(batch[0], all);
  for(comp_type c = 1; c <= 4; c++) { REQUIRE(fast[c] == all[c]); }

  gcsa_b200_built_free(&built);
  gcsa_b200_kmers_free(&kmers);
  std::printf("facade_test OK: %zu patterns, %zu found\n", patterns.size(), (size_t)found);
  return 0;
}
