/*
  gcsa2_b200.hpp -- C++ facade over the C ABI with the signatures of the reference's classes.

  gcsa_b200::GCSA     mirrors gcsa::GCSA      (reference include/gcsa/gcsa.h:40-275)
  gcsa_b200::LCPArray mirrors gcsa::LCPArray  (reference include/gcsa/lcp.h:90-194)

  Same method names, argument meaning and result conventions: ranges are closed pairs, empty
  results of find()/LF() are returned uncanonicalised, locate() honours append/sort exactly like
  src/gcsa.cpp:813-842, count()/locate() treat ranges past the index as empty, queries never
  throw; only construction throws (the reference's load() throws std::runtime_error,
  src/gcsa.cpp:188-193).  All query methods are const and re-entrant: every call uses its own
  CUDA stream, so they may be called concurrently from OpenMP threads like the reference's
  (src/algorithms.cpp:113, 409).

  A single-pattern call is a batch of one: correct, but the GPU earns its keep on the batch
  overloads (find(patterns, results), count(ranges, results), locate(ranges, offsets, values)).
  Header-only; link with libgcsa2_b200.so.
*/
#ifndef GCSA2_B200_HPP
#define GCSA2_B200_HPP

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "gcsa2_b200.h"

namespace gcsa_b200
{

typedef std::uint64_t size_type;
typedef std::uint8_t  comp_type;
typedef std::uint64_t node_type;
typedef std::pair<size_type, size_type> range_type;      // include/gcsa/utils.h:84-117

struct Range
{
  static size_type length(range_type range) { return range.second + 1 - range.first; }
  static bool empty(range_type range) { return (range.first + 1 > range.second + 1); }
  static bool empty(size_type sp, size_type ep) { return (sp + 1 > ep + 1); }
  static range_type empty_range() { return range_type(1, 0); }
};

struct Node                                                // include/gcsa/support.h:443-471
{
  constexpr static size_type OFFSET_BITS = 10;
  constexpr static size_type ID_OFFSET = OFFSET_BITS + 1;
  constexpr static size_type ORIENTATION_MASK = static_cast<size_type>(1) << OFFSET_BITS;
  constexpr static size_type OFFSET_MASK = ORIENTATION_MASK - 1;
  static node_type encode(size_type id, size_type offset, bool rc = false) { return (id << ID_OFFSET) | offset | (rc ? ORIENTATION_MASK : 0); }
  static size_type id(node_type node) { return node >> ID_OFFSET; }
  static bool rc(node_type node) { return node & ORIENTATION_MASK; }
  static size_type offset(node_type node) { return node & OFFSET_MASK; }
};

struct STNode                                              // include/gcsa/lcp.h:40-79
{
  size_type sp, ep, left_lcp, right_lcp, node_lcp;
  constexpr static size_type UNKNOWN = ~(size_type)0;
  STNode() : sp(0), ep(0), left_lcp(0), right_lcp(0), node_lcp(0) {}
  STNode(size_type s, size_type e, size_type l, size_type r, size_type d) : sp(s), ep(e), left_lcp(l), right_lcp(r), node_lcp(d) {}
  range_type range() const { return range_type(sp, ep); }
  size_type lcp() const { return node_lcp; }
  bool operator==(const STNode& n) const { return sp == n.sp && ep == n.ep; }
  bool operator==(range_type r) const { return sp == r.first && ep == r.second; }
  bool operator!=(const STNode& n) const { return !(*this == n); }
  bool operator!=(range_type r) const { return !(*this == r); }
};

inline void check(int rc, const char* what)
{
  if(rc != 0) { throw std::runtime_error(std::string(what) + ": " + gcsa_b200_last_error()); }
}

//------------------------------------------------------------------------------

class GCSA
{
public:
  typedef gcsa_b200::size_type size_type;

  GCSA() : handle(nullptr) {}
  explicit GCSA(const gcsa_flat_index& flat, int device = 0, int kmer_table_k = 12) : handle(nullptr)
  {
    gcsa_b200_options options = {};
    options.kmer_table_k = kmer_table_k; options.two_step = -1; options.walk_table = -1;
    check(gcsa_b200_index_create(&flat, device, &options, &handle), "GCSA::GCSA()");
    gcsa_b200_index_info(handle, &info);
    for(int i = 0; i < 256; i++) { char2comp[i] = flat.char2comp[i]; }
  }
  // From a .gcsa file of the reference: sdsl::load_from_file(index, name) -> GCSA::load()
  // (src/build_gcsa.cpp:148-153, src/gcsa.cpp:182-216); throws std::runtime_error on a bad file
  // like the reference does (gcsa.cpp:188-193).
  explicit GCSA(const std::string& gcsa_file, int device = 0, int kmer_table_k = 12) : handle(nullptr)
  {
    gcsa_b200_built loaded;
    check(gcsa_b200_load_gcsa_file(gcsa_file.c_str(), &loaded), "GCSA::load()");
    gcsa_b200_options options = {};
    options.kmer_table_k = kmer_table_k; options.two_step = -1; options.walk_table = -1;
    int rc = gcsa_b200_index_create(&loaded.index, device, &options, &handle);
    for(int i = 0; i < 256; i++) { char2comp[i] = loaded.index.char2comp[i]; }
    gcsa_b200_built_free(&loaded);
    check(rc, "GCSA::load()");
    gcsa_b200_index_info(handle, &info);
  }
  ~GCSA() { gcsa_b200_index_destroy(handle); }
  GCSA(const GCSA&) = delete;
  GCSA& operator=(const GCSA&) = delete;
  GCSA(GCSA&& other) noexcept : handle(other.handle), info(other.info) { std::copy(other.char2comp, other.char2comp + 256, char2comp); other.handle = nullptr; }

  // ---- high-level interface (gcsa.h:96-128) ----
  template<class Iterator>
  range_type find(Iterator begin, Iterator end) const
  {
    std::vector<std::uint8_t> buffer;
    for(Iterator it = begin; it != end; ++it) { buffer.push_back(static_cast<std::uint8_t>(*it)); }
    return this->find(buffer.data(), buffer.size());
  }

  template<class Container>
  range_type find(const Container& pattern) const { return this->find(pattern.begin(), pattern.end()); }

  template<class Element>
  range_type find(const Element* pattern, size_type length) const
  {
    std::vector<std::uint8_t> buffer(length + 1);
    for(size_type i = 0; i < length; i++) { buffer[i] = static_cast<std::uint8_t>(pattern[i]); }
    std::uint64_t offsets[2] = { 0, length }, sp = 0, ep = 0;
    gcsa_b200_find_host(handle, buffer.data(), offsets, 1, &sp, &ep);
    return range_type(sp, ep);
  }

  // batch: results[i] = find(patterns[i])
  void find(const std::vector<std::string>& patterns, std::vector<range_type>& results) const
  {
    std::vector<std::uint64_t> offsets(patterns.size() + 1, 0);
    for(size_type i = 0; i < patterns.size(); i++) { offsets[i + 1] = offsets[i] + patterns[i].size(); }
    std::vector<std::uint8_t> chars(offsets.back() + 1);
    for(size_type i = 0; i < patterns.size(); i++) { std::copy(patterns[i].begin(), patterns[i].end(), chars.begin() + offsets[i]); }
    std::vector<std::uint64_t> sp(patterns.size() + 1), ep(patterns.size() + 1);
    gcsa_b200_find_host(handle, chars.data(), offsets.data(), patterns.size(), sp.data(), ep.data());
    results.resize(patterns.size());
    for(size_type i = 0; i < patterns.size(); i++) { results[i] = range_type(sp[i], ep[i]); }
  }

  size_type count(range_type range) const
  {
    std::uint64_t out = 0;
    gcsa_b200_count_host(handle, &range.first, &range.second, 1, &out);
    return out;
  }

  void count(const std::vector<range_type>& ranges, std::vector<size_type>& results) const
  {
    std::vector<std::uint64_t> sp, ep; split(ranges, sp, ep);
    results.assign(ranges.size() + 1, 0);
    gcsa_b200_count_host(handle, sp.data(), ep.data(), ranges.size(), results.data());
    results.resize(ranges.size());
  }

  void locate(size_type path, std::vector<node_type>& results, bool append = false, bool sort = true) const
  {
    if(!append) { results.clear(); }
    if(path >= this->size()) { if(sort) { removeDuplicates(results); } return; }          // gcsa.cpp:817
    this->locateInto(range_type(path, path), results, sort && !append);
    if(sort && append) { removeDuplicates(results); }
  }

  void locate(range_type range, std::vector<node_type>& results, bool append = false, bool sort = true) const
  {
    if(!append) { results.clear(); }
    if(Range::empty(range) || range.second >= this->size()) { if(sort) { removeDuplicates(results); } return; }   // gcsa.cpp:831
    this->locateInto(range, results, sort && !append);
    if(sort && append) { removeDuplicates(results); }
  }

  void locate(range_type range, size_type max_positions, std::vector<node_type>& results) const
  {
    results.clear();
    std::uint64_t offsets[2] = { 0, 0 }; std::uint64_t* values = nullptr;
    if(gcsa_b200_locate_max_host(handle, &range.first, &range.second, 1, max_positions, offsets, &values) == 0)
    {
      results.assign(values, values + offsets[1]);
    }
    gcsa_b200_free(values);
  }

  // batch: CSR of sorted distinct values, values[offsets[i] .. offsets[i+1]) for ranges[i]
  void locate(const std::vector<range_type>& ranges, std::vector<size_type>& offsets, std::vector<node_type>& values) const
  {
    std::vector<std::uint64_t> sp, ep; split(ranges, sp, ep);
    offsets.assign(ranges.size() + 1, 0);
    std::uint64_t* out = nullptr;
    if(gcsa_b200_locate_host(handle, sp.data(), ep.data(), ranges.size(), offsets.data(), &out) == 0)
    {
      values.assign(out, out + offsets.back());
    }
    gcsa_b200_free(out);
  }

  // ---- low-level interface (gcsa.h:137-210) ----
  size_type size() const { return info.path_nodes; }
  bool empty() const { return (this->size() == 0); }
  size_type edgeCount() const { return info.edge_count; }
  size_type order() const { return info.order; }
  size_type sampleCount() const { return info.sample_count; }

  range_type charRange(comp_type comp) const
  {
    std::uint64_t sp = 0, ep = 0;
    gcsa_b200_char_range(handle, comp, &sp, &ep);
    return range_type(sp, ep);
  }

  range_type LF(range_type range, comp_type comp) const
  {
    std::uint64_t sp = 0, ep = 0;
    gcsa_b200_lf_host(handle, &range.first, &range.second, &comp, 1, &sp, &ep);
    return range_type(sp, ep);
  }

  size_type LF(size_type path_node) const
  {
    std::uint64_t out = 0;
    gcsa_b200_lf_node_host(handle, &path_node, 1, &out);
    return out;
  }

  // results must have at least sigma elements, like in the reference (gcsa.cpp:742-798).
  void LF_fast(range_type range, std::vector<range_type>& results) const { this->lfMulti(range, results, 0, GCSA_B200_FAST_CHARS); }
  void LF_all(range_type range, std::vector<range_type>& results) const { this->lfMulti(range, results, 1, GCSA_B200_SIGMA - 2); }

  std::uint8_t      char2comp[256];          // Alphabet::char2comp (alpha.char2comp in the reference)
  gcsa_b200_index*  handle;
  gcsa_b200_info    info;

private:
  static void removeDuplicates(std::vector<node_type>& vec)              // utils.h:350-357
  {
    std::sort(vec.begin(), vec.end());
    vec.resize(std::unique(vec.begin(), vec.end()) - vec.begin());
  }

  static void split(const std::vector<range_type>& ranges, std::vector<std::uint64_t>& sp, std::vector<std::uint64_t>& ep)
  {
    sp.resize(ranges.size() + 1); ep.resize(ranges.size() + 1);
    for(size_type i = 0; i < ranges.size(); i++) { sp[i] = ranges[i].first; ep[i] = ranges[i].second; }
  }

  void locateInto(range_type range, std::vector<node_type>& results, bool sorted) const
  {
    std::uint64_t offsets[2] = { 0, 0 }; std::uint64_t* values = nullptr;
    int rc = (sorted ? gcsa_b200_locate_host(handle, &range.first, &range.second, 1, offsets, &values)
                     : gcsa_b200_locate_raw_host(handle, &range.first, &range.second, 1, offsets, &values));
    if(rc == 0) { results.insert(results.end(), values, values + offsets[1]); }
    gcsa_b200_free(values);
  }

  void lfMulti(range_type range, std::vector<range_type>& results, int all_chars, size_type last) const
  {
    std::uint64_t out[GCSA_B200_SIGMA * 2];
    if(gcsa_b200_lf_multi_host(handle, &range.first, &range.second, 1, all_chars, out) != 0) { return; }
    for(size_type comp = 1; comp <= last && comp < results.size(); comp++) { results[comp] = range_type(out[2 * comp], out[2 * comp + 1]); }
  }
};

//------------------------------------------------------------------------------

//------------------------------------------------------------------------------
// One process, several GPUs: `replicas` are GCSA objects of the same index, one per device.  The batch is cut into
// contiguous blocks, one per replica, each driven by its own host thread (gcsa_b200_*_multi); results as for the
// member functions of the same name.  This is the form an OpenMP-over-queries caller (src/algorithms.cpp:113) uses
// on a box with several GPUs.
//------------------------------------------------------------------------------

inline void find(const std::vector<const GCSA*>& replicas, const std::vector<std::string>& patterns, std::vector<range_type>& results)
{
  std::vector<const gcsa_b200_index*> handles;
  for(const GCSA* r : replicas) { handles.push_back(r->handle); }
  std::vector<std::uint64_t> offsets(patterns.size() + 1, 0);
  for(size_type i = 0; i < patterns.size(); i++) { offsets[i + 1] = offsets[i] + patterns[i].size(); }
  std::vector<std::uint8_t> chars(offsets.back() + 1);
  for(size_type i = 0; i < patterns.size(); i++) { std::copy(patterns[i].begin(), patterns[i].end(), chars.begin() + offsets[i]); }
  std::vector<std::uint64_t> sp(patterns.size() + 1), ep(patterns.size() + 1);
  gcsa_b200_find_host_multi(handles.data(), (int)handles.size(), chars.data(), offsets.data(), patterns.size(), sp.data(), ep.data());
  results.resize(patterns.size());
  for(size_type i = 0; i < patterns.size(); i++) { results[i] = range_type(sp[i], ep[i]); }
}

// CSR of sorted distinct values, values[offsets[i] .. offsets[i+1]) for ranges[i]
inline void locate(const std::vector<const GCSA*>& replicas, const std::vector<range_type>& ranges, std::vector<size_type>& offsets, std::vector<node_type>& values)
{
  std::vector<const gcsa_b200_index*> handles;
  for(const GCSA* r : replicas) { handles.push_back(r->handle); }
  std::vector<std::uint64_t> sp(ranges.size() + 1), ep(ranges.size() + 1);
  for(size_type i = 0; i < ranges.size(); i++) { sp[i] = ranges[i].first; ep[i] = ranges[i].second; }
  offsets.assign(ranges.size() + 1, 0);
  std::uint64_t needed = 0;
  gcsa_b200_locate_into_host_multi(handles.data(), (int)handles.size(), sp.data(), ep.data(), ranges.size(), offsets.data(), nullptr, 0, &needed);
  values.assign(needed + 1, 0);
  if(gcsa_b200_locate_into_host_multi(handles.data(), (int)handles.size(), sp.data(), ep.data(), ranges.size(), offsets.data(), values.data(), needed, &needed) != 0)
  {
    needed = 0; offsets.assign(ranges.size() + 1, 0);
  }
  values.resize(needed);
}

class LCPArray
{
public:
  typedef gcsa_b200::size_type size_type;
  typedef STNode               node_type;

  LCPArray() : handle(nullptr), size_(0), values_(0), levels_(0), branching_(0) {}
  explicit LCPArray(const gcsa_flat_lcp& flat, int device = 0) : handle(nullptr)
  {
    check(gcsa_b200_lcp_create(&flat, device, &handle), "LCPArray::LCPArray()");
    size_ = flat.size; levels_ = flat.levels; branching_ = flat.branching; values_ = flat.offsets[flat.levels];
  }
  // From a .lcp file of the reference: LCPArray::load(), src/lcp.cpp:128-143.
  explicit LCPArray(const std::string& lcp_file, int device = 0) : handle(nullptr)
  {
    gcsa_flat_lcp flat;
    check(gcsa_b200_load_lcp_file(lcp_file.c_str(), &flat), "LCPArray::load()");
    int rc = gcsa_b200_lcp_create(&flat, device, &handle);
    size_ = flat.size; levels_ = flat.levels; branching_ = flat.branching; values_ = flat.offsets[flat.levels];
    gcsa_b200_flat_lcp_free(&flat);
    check(rc, "LCPArray::load()");
  }
  ~LCPArray() { gcsa_b200_lcp_destroy(handle); }
  LCPArray(const LCPArray&) = delete;
  LCPArray& operator=(const LCPArray&) = delete;

  size_type size() const { return size_; }
  size_type values() const { return values_; }
  size_type levels() const { return levels_; }
  size_type branching() const { return branching_; }

  node_type root() const { return node_type(0, this->size() - 1, 0, 0, 0); }                 // lcp.h:137
  range_type notFound() const { return range_type(this->values(), this->values()); }        // lcp.h:178

  node_type parent(range_type range) const                                                   // lcp.cpp:297-301
  {
    gcsa_b200_stnode out = {};
    gcsa_b200_parent_host(handle, &range.first, &range.second, 1, &out);
    return node_type(out.sp, out.ep, out.left_lcp, out.right_lcp, out.node_lcp);
  }
  node_type parent(const node_type& node) const { return (node == this->root() ? this->root() : this->parent(node.range())); }

  void parent(const std::vector<range_type>& ranges, std::vector<node_type>& results) const
  {
    std::vector<std::uint64_t> sp(ranges.size() + 1), ep(ranges.size() + 1);
    for(size_type i = 0; i < ranges.size(); i++) { sp[i] = ranges[i].first; ep[i] = ranges[i].second; }
    std::vector<gcsa_b200_stnode> out(ranges.size() + 1);
    gcsa_b200_parent_host(handle, sp.data(), ep.data(), ranges.size(), out.data());
    results.resize(ranges.size());
    for(size_type i = 0; i < ranges.size(); i++) { results[i] = node_type(out[i].sp, out[i].ep, out[i].left_lcp, out[i].right_lcp, out[i].node_lcp); }
  }

  size_type depth(range_type range) const                                                    // lcp.cpp:319-325
  {
    std::uint64_t out = 0;
    gcsa_b200_depth_host(handle, &range.first, &range.second, 1, &out);
    return out;
  }
  size_type depth(const node_type& node) const { return (node.lcp() != node_type::UNKNOWN ? node.lcp() : this->depth(node.range())); }
  size_type depth(node_type& node) const { if(node.lcp() == node_type::UNKNOWN) { node.node_lcp = this->depth(node.range()); } return node.lcp(); }

  range_type psv(size_type pos) const { return this->sv(0, pos); }
  range_type psev(size_type pos) const { return this->sv(1, pos); }
  range_type nsv(size_type pos) const { return this->sv(2, pos); }
  range_type nsev(size_type pos) const { return this->sv(3, pos); }

  range_type rmq(size_type sp, size_type ep) const
  {
    std::uint64_t a = 0, b = 0;
    gcsa_b200_lcp_rmq_host(handle, &sp, &ep, 1, &a, &b);
    return range_type(a, b);
  }
  range_type rmq(range_type range) const { return this->rmq(range.first, range.second); }

  gcsa_b200_lcp* handle;

private:
  range_type sv(int which, size_type pos) const
  {
    std::uint64_t a = 0, b = 0;
    gcsa_b200_lcp_sv_host(handle, which, &pos, 1, &a, &b);
    return range_type(a, b);
  }
  size_type size_, values_, levels_, branching_;
};

} // namespace gcsa_b200

#endif // GCSA2_B200_HPP
