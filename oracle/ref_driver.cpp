/*
  ref_driver.cpp -- C entry points over the REFERENCE's own classes.  TEST INFRASTRUCTURE ONLY.

  Compiled by oracle/Makefile together with the reference's unmodified sources
  (/root/reference/src/*.cpp, where they lie) and the SDSL shim (oracle/sdsl_shim/) into
  oracle/_ref/libgcsa2_ref.so.  Everything that answers a query here is reference code:
  gcsa::GCSA::GCSA(InputGraph&, ...) builds the index from a binary .graph kmer file,
  gcsa::GCSA::find / count / locate / LF and gcsa::LCPArray::parent / depth answer the queries,
  gcsa::verifyIndex and gcsa::countKMers are the reference's own.  This file only moves arguments.
*/
#include <gcsa/gcsa.h>
#include <gcsa/lcp.h>
#include <gcsa/algorithms.h>

#include <cstring>
#include <functional>
#include <memory>
#include <omp.h>

using namespace gcsa;

struct RefIndex
{
  std::unique_ptr<InputGraph> graph;
  GCSA index;
  LCPArray lcp;
  bool has_lcp = false;
};

namespace
{

std::vector<std::uint64_t> plainBits(std::uint64_t n, const std::function<bool(std::uint64_t)>& get)
{
  std::vector<std::uint64_t> words(n / 64 + 2, 0);
  for(std::uint64_t i = 0; i < n; i++) { if(get(i)) { words[i >> 6] |= (std::uint64_t)1 << (i & 63); } }
  return words;
}

void copyOut(const std::vector<std::uint64_t>& words, std::uint64_t* out, std::uint64_t n_bits)
{
  if(out != nullptr) { std::memcpy(out, words.data(), ((n_bits + 63) / 64) * sizeof(std::uint64_t)); }
}

} // namespace

extern "C" {

/* GCSA(InputGraph&, ConstructionParameters) + LCPArray(InputGraph&, ...) on a binary kmer file. */
RefIndex* ref_build_from(const char* graph_file, int binary, const char* mapping_file, int doubling_steps, std::uint64_t sample_period,
                         std::uint64_t lcp_branching, const char* temp_dir);

RefIndex* ref_build(const char* graph_file, int doubling_steps, std::uint64_t sample_period, std::uint64_t lcp_branching, const char* temp_dir)
{
  return ref_build_from(graph_file, 1, nullptr, doubling_steps, sample_period, lcp_branching, temp_dir);
}

/* The same from a binary (.graph) or text (.gcsa2) kmer file, optionally with a NodeMapping file (build_gcsa's mapping option):
   InputGraph(files, binary, parameters, alphabet, mapping_name), src/files.cpp:286-361. */
RefIndex* ref_build_from(const char* graph_file, int binary, const char* mapping_file, int doubling_steps, std::uint64_t sample_period,
                         std::uint64_t lcp_branching, const char* temp_dir)
{
  Verbosity::set(Verbosity::SILENT);
  if(temp_dir != nullptr) { TempFile::setDirectory(temp_dir); }
  ConstructionParameters parameters;
  parameters.setSteps(doubling_steps);
  parameters.setSamplePeriod(sample_period);
  parameters.setLCPBranching(lcp_branching);
  RefIndex* r = new RefIndex();
  std::vector<std::string> files(1, graph_file);
  r->graph.reset(new InputGraph(files, binary != 0, parameters, Alphabet(), std::string(mapping_file != nullptr ? mapping_file : "")));
  GCSA built(*(r->graph), parameters);
  r->index.swap(built);
  LCPArray lcp(*(r->graph), parameters);
  r->lcp.swap(lcp);
  r->has_lcp = true;
  return r;
}

void ref_destroy(RefIndex* r) { delete r; }

/* verifyIndex(index, &lcp, graph), src/algorithms.cpp:85-99 -> 1 if the reference accepts its own index. */
int ref_verify(RefIndex* r)
{
  std::streambuf* old = std::cout.rdbuf(nullptr);          // it reports on stdout
  bool ok = verifyIndex(r->index, (r->has_lcp ? &(r->lcp) : nullptr), *(r->graph));
  std::cout.rdbuf(old);
  return ok ? 1 : 0;
}

/* header: path_nodes, edges, order, sample_count, extra_values_len, redundant_len, lcp size, lcp branching, lcp levels, lcp values */
void ref_sizes(const RefIndex* r, std::uint64_t* out)
{
  const GCSA& g = r->index;
  out[0] = g.size(); out[1] = g.edgeCount(); out[2] = g.order(); out[3] = g.sampleCount();
  out[4] = g.extra_pointers.values.size(); out[5] = g.redundant_pointers.data.size();
  out[6] = r->lcp.size(); out[7] = r->lcp.branching(); out[8] = r->lcp.levels(); out[9] = r->lcp.values();
}

/* The members of gcsa::GCSA (include/gcsa/gcsa.h:214-240) as plain bit vectors / arrays. */
void ref_export(const RefIndex* r, std::uint64_t* C, std::uint64_t** bwt, std::uint64_t* edges, std::uint64_t* sampled_paths,
                std::uint64_t* stored_samples, std::uint64_t* samples, std::uint64_t* extra_filter, std::uint64_t* extra_values,
                std::uint64_t* redundant, std::uint64_t* lcp_offsets, std::uint8_t* lcp_data)
{
  const GCSA& g = r->index;
  std::uint64_t N = g.size();
  for(size_type c = 0; c <= g.alpha.sigma; c++) { C[c] = g.alpha.C[c]; }
  for(size_type c = 0; c < g.alpha.sigma; c++)
  {
    bool fast = (c >= 1 && c <= g.alpha.fast_chars);
    copyOut(plainBits(N, [&](std::uint64_t i) { return fast ? (bool)g.fast_bwt[c][i] : (bool)g.sparse_bwt[c][i]; }), bwt[c], N);
  }
  copyOut(plainBits(g.edges.size(), [&](std::uint64_t i) { return (bool)g.edges[i]; }), edges, g.edges.size());
  copyOut(plainBits(N, [&](std::uint64_t i) { return (bool)g.sampled_paths[i]; }), sampled_paths, N);
  for(size_type i = 0; i < g.sampleCount(); i++) { stored_samples[i] = g.stored_samples[i]; }
  copyOut(plainBits(g.samples.size(), [&](std::uint64_t i) { return (bool)g.samples[i]; }), samples, g.samples.size());
  copyOut(plainBits(N, [&](std::uint64_t i) { return (bool)g.extra_pointers.filter[i]; }), extra_filter, N);
  copyOut(plainBits(g.extra_pointers.values.size(), [&](std::uint64_t i) { return (bool)g.extra_pointers.values[i]; }), extra_values, g.extra_pointers.values.size());
  copyOut(plainBits(g.redundant_pointers.data.size(), [&](std::uint64_t i) { return (bool)g.redundant_pointers.data[i]; }), redundant, g.redundant_pointers.data.size());
  for(size_type i = 0; i <= r->lcp.levels(); i++) { lcp_offsets[i] = r->lcp.offsets[i]; }
  for(size_type i = 0; i < r->lcp.values(); i++) { lcp_data[i] = (std::uint8_t)r->lcp.data[i]; }
}

/* Load plain arrays into a reference GCSA object (public members, gcsa.h:214-240) so that the
   reference's query code can be timed on an index built elsewhere. */
RefIndex* ref_from_flat(std::uint64_t path_nodes, std::uint64_t edge_count, std::uint64_t order, const std::uint64_t* C,
                        const std::uint64_t* const* bwt, const std::uint64_t* edges, const std::uint64_t* sampled_paths,
                        std::uint64_t sample_count, const std::uint64_t* stored_samples, const std::uint64_t* samples)
{
  RefIndex* r = new RefIndex();
  GCSA& g = r->index;
  g.header.path_nodes = path_nodes; g.header.edges = edge_count; g.header.order = order;
  sdsl::int_vector<64> counts(g.alpha.sigma, 0);
  for(size_type c = 0; c < g.alpha.sigma; c++) { counts[c] = C[c + 1] - C[c]; }
  g.alpha = Alphabet(counts);
  auto bits = [](const std::uint64_t* words, std::uint64_t n) {
    sdsl::bit_vector v(n, 0);
    std::memcpy(v.data(), words, ((n + 63) / 64) * sizeof(std::uint64_t));
    return v;
  };
  g.fast_bwt.resize(g.alpha.sigma); g.fast_rank.resize(g.alpha.sigma);
  g.sparse_bwt.resize(g.alpha.sigma); g.sparse_rank.resize(g.alpha.sigma);
  for(size_type c = 0; c < g.alpha.sigma; c++)
  {
    if(c >= 1 && c <= g.alpha.fast_chars) { g.fast_bwt[c] = bits(bwt[c], path_nodes); }
    else { g.sparse_bwt[c] = GCSA::sparse_vector(bits(bwt[c], path_nodes)); }
  }
  g.edges = bits(edges, edge_count);
  g.sampled_paths = bits(sampled_paths, path_nodes);
  std::uint64_t max_sample = 0;
  for(std::uint64_t i = 0; i < sample_count; i++) { max_sample = std::max(max_sample, stored_samples[i]); }
  g.stored_samples = sdsl::int_vector<0>(sample_count, 0, bit_length(max_sample));
  for(std::uint64_t i = 0; i < sample_count; i++) { g.stored_samples[i] = stored_samples[i]; }
  g.samples = bits(samples, sample_count);
  for(size_type c = 0; c < g.alpha.sigma; c++)        // GCSA::initSupport(), gcsa.cpp:726-738
  {
    sdsl::util::init_support(g.fast_rank[c], &(g.fast_bwt[c]));
    sdsl::util::init_support(g.sparse_rank[c], &(g.sparse_bwt[c]));
  }
  sdsl::util::init_support(g.edge_rank, &(g.edges));
  sdsl::util::init_support(g.sampled_path_rank, &(g.sampled_paths));
  sdsl::util::init_support(g.sample_select, &(g.samples));
  return r;
}

/* sdsl::store_to_file(index, name) / load_from_file, the way build_gcsa does it (src/build_gcsa.cpp:148-178):
   GCSA::serialize / GCSA::load and LCPArray::serialize / load are the reference's own. */
int ref_store(const RefIndex* r, const char* gcsa_file, const char* lcp_file)
{
  if(gcsa_file != nullptr && !sdsl::store_to_file(r->index, gcsa_file)) { return 0; }
  if(lcp_file != nullptr && r->has_lcp && !sdsl::store_to_file(r->lcp, lcp_file)) { return 0; }
  return 1;
}

RefIndex* ref_load(const char* gcsa_file, const char* lcp_file)
{
  RefIndex* r = new RefIndex();
  try
  {
    if(!sdsl::load_from_file(r->index, gcsa_file)) { delete r; return nullptr; }
    if(lcp_file != nullptr)
    {
      if(!sdsl::load_from_file(r->lcp, lcp_file)) { delete r; return nullptr; }
      r->has_lcp = true;
    }
  }
  catch(const std::exception&) { delete r; return nullptr; }
  return r;
}

int ref_max_threads(void) { return omp_get_max_threads(); }

/* The loop of benchmark/query_gcsa.cpp:88-103 over GCSA::find, OpenMP like src/algorithms.cpp:113. */
double ref_find_batch(const RefIndex* r, const std::uint8_t* chars, const std::uint64_t* offsets, std::uint64_t n,
                      std::uint64_t* sp, std::uint64_t* ep, int threads)
{
  if(threads < 1) { threads = 1; }
  double start = omp_get_wtime();
  #pragma omp parallel for schedule(dynamic, 4096) num_threads(threads)
  for(std::uint64_t i = 0; i < n; i++)
  {
    range_type range = r->index.find(chars + offsets[i], chars + offsets[i + 1]);
    sp[i] = range.first; ep[i] = range.second;
  }
  return omp_get_wtime() - start;
}

void ref_lf_batch(const RefIndex* r, const std::uint64_t* sp, const std::uint64_t* ep, const std::uint8_t* comp, std::uint64_t n,
                  std::uint64_t* osp, std::uint64_t* oep)
{
  for(std::uint64_t i = 0; i < n; i++)
  {
    range_type range = r->index.LF(range_type(sp[i], ep[i]), comp[i]);
    osp[i] = range.first; oep[i] = range.second;
  }
}

void ref_lf_node_batch(const RefIndex* r, const std::uint64_t* nodes, std::uint64_t n, std::uint64_t* out)
{
  for(std::uint64_t i = 0; i < n; i++) { out[i] = r->index.LF(nodes[i]); }
}

void ref_count_batch(const RefIndex* r, const std::uint64_t* sp, const std::uint64_t* ep, std::uint64_t n, std::uint64_t* out)
{
  for(std::uint64_t i = 0; i < n; i++) { out[i] = r->index.count(range_type(sp[i], ep[i])); }
}

/* GCSA::locate(range, results) / locate(range, max_positions, results) as a CSR; values malloc'ed. */
double ref_locate_batch(const RefIndex* r, const std::uint64_t* sp, const std::uint64_t* ep, std::uint64_t n, std::uint64_t max_positions,
                        std::uint64_t* out_offsets, std::uint64_t** values, int threads)
{
  if(threads < 1) { threads = 1; }
  std::vector<std::vector<node_type>> results(n);
  double start = omp_get_wtime();
  #pragma omp parallel for schedule(dynamic, 256) num_threads(threads)
  for(std::uint64_t i = 0; i < n; i++)
  {
    if(max_positions == 0) { r->index.locate(range_type(sp[i], ep[i]), results[i]); }
    else { r->index.locate(range_type(sp[i], ep[i]), max_positions, results[i]); }
  }
  double seconds = omp_get_wtime() - start;
  out_offsets[0] = 0;
  for(std::uint64_t i = 0; i < n; i++) { out_offsets[i + 1] = out_offsets[i] + results[i].size(); }
  std::uint64_t* vals = (std::uint64_t*)std::malloc((out_offsets[n] + 1) * sizeof(std::uint64_t));
  for(std::uint64_t i = 0; i < n; i++) { std::copy(results[i].begin(), results[i].end(), vals + out_offsets[i]); }
  *values = vals;
  return seconds;
}

void ref_free(void* p) { std::free(p); }

void ref_parent_batch(const RefIndex* r, const std::uint64_t* sp, const std::uint64_t* ep, std::uint64_t n, std::uint64_t* out /* 5 per range */)
{
  for(std::uint64_t i = 0; i < n; i++)
  {
    STNode node = r->lcp.parent(range_type(sp[i], ep[i]));
    out[5 * i] = node.sp; out[5 * i + 1] = node.ep; out[5 * i + 2] = node.left_lcp; out[5 * i + 3] = node.right_lcp; out[5 * i + 4] = node.node_lcp;
  }
}

void ref_depth_batch(const RefIndex* r, const std::uint64_t* sp, const std::uint64_t* ep, std::uint64_t n, std::uint64_t* out)
{
  for(std::uint64_t i = 0; i < n; i++) { out[i] = r->lcp.depth(range_type(sp[i], ep[i])); }
}

/* which: 0 psv, 1 psev, 2 nsv, 3 nsev, 4 rmq(a, b) */
void ref_lcp_query(const RefIndex* r, int which, const std::uint64_t* a, const std::uint64_t* b, std::uint64_t n, std::uint64_t* opos, std::uint64_t* oval)
{
  for(std::uint64_t i = 0; i < n; i++)
  {
    range_type res;
    switch(which)
    {
      case 0: res = r->lcp.psv(a[i]); break;
      case 1: res = r->lcp.psev(a[i]); break;
      case 2: res = r->lcp.nsv(a[i]); break;
      case 3: res = r->lcp.nsev(a[i]); break;
      default: res = r->lcp.rmq(a[i], b[i]); break;
    }
    opos[i] = res.first; oval[i] = res.second;
  }
}

std::uint64_t ref_count_kmers(const RefIndex* r, std::uint64_t k, int include_Ns)
{
  KMerSearchParameters parameters;
  parameters.include_Ns = (include_Ns != 0);
  parameters.force = true;
  return countKMers(r->index, k, parameters);
}

/* compareKMers(left, right, k, parameters), src/algorithms.cpp:535-616 */
void ref_compare_kmers(const RefIndex* left, const RefIndex* right, std::uint64_t k, int include_Ns, std::uint64_t* result)
{
  KMerSearchParameters parameters;
  parameters.include_Ns = (include_Ns != 0);
  parameters.force = true;
  std::array<size_type, 3> res = compareKMers(left->index, right->index, k, parameters);
  result[0] = res[0]; result[1] = res[1]; result[2] = res[2];
}

} // extern "C"
