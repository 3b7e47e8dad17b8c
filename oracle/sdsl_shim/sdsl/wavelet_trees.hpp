/*
  A minimal SDSL-compatible shim: TEST INFRASTRUCTURE ONLY (oracle/).

  The reference (jltsiren/gcsa2) includes exactly one third-party header,
  <sdsl/wavelet_trees.hpp> (include/gcsa/utils.h:36), from the vgteam fork of sdsl-lite, which is
  neither vendored nor installed here.  This file provides, written from scratch, just enough of
  the sdsl:: interface for the reference's sources to compile UNCHANGED from where they lie, so
  that its own constructor, find(), locate(), count() and LCPArray::parent() can be run as the
  parity reference (oracle/_ref/, built by oracle/Makefile).  It is not SDSL: the data structures
  are the simplest ones with the same semantics (plain bit vectors with interleaved rank blocks,
  sorted position lists), serialization uses its own byte layout, and nothing here is tuned.
*/
#ifndef GCSA2_B200_SDSL_SHIM_HPP
#define GCSA2_B200_SDSL_SHIM_HPP

#include <algorithm>
#include <array>
#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <deque>
#include <queue>
#include <set>
#include <stack>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <map>
#include <sstream>
#include <string>
#include <typeinfo>
#include <utility>
#include <vector>
#include <unistd.h>

namespace sdsl
{

typedef std::uint64_t size_type_shim;

//------------------------------------------------------------------------------
// bits
//------------------------------------------------------------------------------

namespace shim_detail
{
  struct LtCnt
  {
    std::uint8_t t[256];
    constexpr LtCnt() : t() { for(int i = 0; i < 256; i++) { int c = 0; for(int b = 0; b < 8; b++) { c += (i >> b) & 1; } t[i] = (std::uint8_t)c; } }
    constexpr std::uint8_t operator[](std::size_t i) const { return t[i]; }
  };
  struct LoSet
  {
    std::uint64_t t[65];
    constexpr LoSet() : t() { for(int i = 0; i < 64; i++) { t[i] = ((std::uint64_t)1 << i) - 1; } t[64] = ~(std::uint64_t)0; }
    constexpr std::uint64_t operator[](std::size_t i) const { return t[i]; }
  };
}

struct bits
{
  static constexpr shim_detail::LtCnt lt_cnt{};      // popcount of a byte
  static constexpr shim_detail::LoSet lo_set{};      // lo_set[i] = i low bits set
  static std::uint32_t hi(std::uint64_t x) { return (x == 0 ? 0 : 63 - __builtin_clzll(x)); }
  static std::uint32_t lo(std::uint64_t x) { return (x == 0 ? 0 : __builtin_ctzll(x)); }
  static std::uint64_t cnt(std::uint64_t x) { return __builtin_popcountll(x); }
};

struct rank_support_shim;
struct select_support_shim;

//------------------------------------------------------------------------------
// structure tree (size reporting): no-ops
//------------------------------------------------------------------------------

class structure_tree_node {};

struct structure_tree
{
  template<class... Args> static structure_tree_node* add_child(Args&&...) { return nullptr; }
  template<class... Args> static void add_size(Args&&...) {}
};

template<class T>
std::uint64_t write_member(const T& t, std::ostream& out, structure_tree_node* = nullptr, std::string = "")
{
  out.write(reinterpret_cast<const char*>(&t), sizeof(T));
  return sizeof(T);
}

inline std::uint64_t write_member(const std::string& t, std::ostream& out, structure_tree_node* = nullptr, std::string = "")
{
  std::uint64_t n = t.size();
  out.write(reinterpret_cast<const char*>(&n), sizeof(n));
  out.write(t.data(), n);
  return sizeof(n) + n;
}

template<class T> void read_member(T& t, std::istream& in) { in.read(reinterpret_cast<char*>(&t), sizeof(T)); }

inline void read_member(std::string& t, std::istream& in)
{
  std::uint64_t n = 0;
  in.read(reinterpret_cast<char*>(&n), sizeof(n));
  t.resize(n);
  in.read(&t[0], n);
}

//------------------------------------------------------------------------------
// int_vector<W>: W = 0 dynamic width, bit-packed little-endian in 64-bit words
//------------------------------------------------------------------------------

template<std::uint8_t W> class int_vector;

template<class Vector>
class int_vector_reference
{
public:
  int_vector_reference(Vector* v, std::uint64_t i) : vec(v), idx(i) {}
  operator std::uint64_t() const { return vec->get(idx); }
  int_vector_reference& operator=(std::uint64_t x) { vec->set(idx, x); return *this; }
  int_vector_reference& operator=(const int_vector_reference& x) { vec->set(idx, (std::uint64_t)x); return *this; }
  int_vector_reference& operator++() { vec->set(idx, vec->get(idx) + 1); return *this; }
  std::uint64_t operator++(int) { std::uint64_t old = vec->get(idx); vec->set(idx, old + 1); return old; }
  int_vector_reference& operator--() { vec->set(idx, vec->get(idx) - 1); return *this; }
  int_vector_reference& operator+=(std::uint64_t x) { vec->set(idx, vec->get(idx) + x); return *this; }
  int_vector_reference& operator-=(std::uint64_t x) { vec->set(idx, vec->get(idx) - x); return *this; }
  int_vector_reference& operator|=(std::uint64_t x) { vec->set(idx, vec->get(idx) | x); return *this; }
private:
  Vector* vec; std::uint64_t idx;
};

template<class Vector>
class int_vector_const_iterator
{
public:
  typedef std::random_access_iterator_tag iterator_category;
  typedef std::uint64_t value_type;
  typedef std::ptrdiff_t difference_type;
  typedef const std::uint64_t* pointer;
  typedef std::uint64_t reference;
  int_vector_const_iterator() : vec(nullptr), idx(0) {}
  int_vector_const_iterator(const Vector* v, std::uint64_t i) : vec(v), idx(i) {}
  std::uint64_t operator*() const { return vec->get(idx); }
  std::uint64_t operator[](difference_type d) const { return vec->get(idx + d); }
  int_vector_const_iterator& operator++() { idx++; return *this; }
  int_vector_const_iterator operator++(int) { int_vector_const_iterator t = *this; idx++; return t; }
  int_vector_const_iterator& operator--() { idx--; return *this; }
  int_vector_const_iterator& operator+=(difference_type d) { idx += d; return *this; }
  int_vector_const_iterator& operator-=(difference_type d) { idx -= d; return *this; }
  int_vector_const_iterator operator+(difference_type d) const { return int_vector_const_iterator(vec, idx + d); }
  int_vector_const_iterator operator-(difference_type d) const { return int_vector_const_iterator(vec, idx - d); }
  difference_type operator-(const int_vector_const_iterator& o) const { return (difference_type)idx - (difference_type)o.idx; }
  bool operator==(const int_vector_const_iterator& o) const { return idx == o.idx; }
  bool operator!=(const int_vector_const_iterator& o) const { return idx != o.idx; }
  bool operator<(const int_vector_const_iterator& o) const { return idx < o.idx; }
private:
  const Vector* vec; std::uint64_t idx;
};

template<std::uint8_t W>
class int_vector
{
public:
  typedef std::uint64_t value_type;
  typedef std::uint64_t size_type;
  typedef int_vector_reference<int_vector> reference;
  typedef int_vector_const_iterator<int_vector> const_iterator;
  typedef const_iterator iterator;
  // rank / select support types are attached below for W == 1
  typedef struct rank_support_shim   rank_1_type;
  typedef struct select_support_shim select_1_type;

  int_vector() : n(0), w(W == 0 ? 64 : W) {}
  int_vector(size_type size, value_type value = 0, std::uint8_t width = W) : n(size), w(W == 0 ? (width == 0 ? 64 : width) : W)
  {
    words.assign(word_count(n, w), 0);
    if(value != 0) { for(size_type i = 0; i < n; i++) { set(i, value); } }
  }
  int_vector(std::initializer_list<value_type> init) : n(init.size()), w(W == 0 ? 64 : W)
  {
    words.assign(word_count(n, w), 0);
    size_type i = 0;
    for(value_type x : init) { set(i++, x); }
  }

  size_type size() const { return n; }
  bool empty() const { return n == 0; }
  std::uint8_t width() const { return w; }
  void width(std::uint8_t new_width) { if(W == 0) { w = new_width; } }
  size_type bit_size() const { return n * w; }
  size_type capacity() const { return words.size() * 64; }
  const std::uint64_t* data() const { return words.data(); }
  std::uint64_t* data() { return words.data(); }

  void resize(size_type size) { n = size; words.resize(word_count(n, w), 0); }
  void bit_resize(size_type bitsize) { n = bitsize / w; words.resize((bitsize + 63) / 64 + 1, 0); }
  void swap(int_vector& o) { std::swap(n, o.n); std::swap(w, o.w); words.swap(o.words); }

  value_type get(size_type i) const
  {
    if(W == 64) { return words[i]; }
    size_type bit = i * w, word = bit >> 6, off = bit & 63;
    std::uint64_t x = words[word] >> off;
    if(off + w > 64) { x |= words[word + 1] << (64 - off); }
    return (w == 64 ? x : x & (((std::uint64_t)1 << w) - 1));
  }
  void set(size_type i, value_type x)
  {
    if(W == 64) { words[i] = x; return; }
    std::uint64_t mask = (w == 64 ? ~(std::uint64_t)0 : (((std::uint64_t)1 << w) - 1));
    x &= mask;
    size_type bit = i * w, word = bit >> 6, off = bit & 63;
    words[word] = (words[word] & ~(mask << off)) | (x << off);
    if(off + w > 64)
    {
      std::uint8_t done = 64 - off;
      words[word + 1] = (words[word + 1] & ~(mask >> done)) | (x >> done);
    }
  }

  reference operator[](size_type i) { return reference(this, i); }
  value_type operator[](size_type i) const { return get(i); }
  const_iterator begin() const { return const_iterator(this, 0); }
  const_iterator end() const { return const_iterator(this, n); }

  bool operator==(const int_vector& o) const
  {
    if(n != o.n || w != o.w) { return false; }
    for(size_type i = 0; i < n; i++) { if(get(i) != o.get(i)) { return false; } }
    return true;
  }
  bool operator!=(const int_vector& o) const { return !(*this == o); }

  // On-disk layout of sdsl-lite: size in bits, the width byte only for int_vector<0>, then
  // ceil(bits / 64) words.
  size_type serialize(std::ostream& out, structure_tree_node* = nullptr, std::string = "") const
  {
    size_type bytes = 0, bit_len = n * w, count = (bit_len + 63) / 64;
    bytes += write_member(bit_len, out);
    if(W == 0) { bytes += write_member(w, out); }
    out.write(reinterpret_cast<const char*>(words.data()), count * sizeof(std::uint64_t));
    return bytes + count * sizeof(std::uint64_t);
  }
  void load(std::istream& in)
  {
    size_type bit_len = 0;
    read_member(bit_len, in);
    if(W == 0) { read_member(w, in); }
    n = bit_len / w;
    words.assign(word_count(n, w), 0);
    in.read(reinterpret_cast<char*>(words.data()), ((bit_len + 63) / 64) * sizeof(std::uint64_t));
  }

private:
  static size_type word_count(size_type size, std::uint8_t width) { return (size * width + 63) / 64 + 1; }
  size_type n; std::uint8_t w;
  std::vector<std::uint64_t> words;
};

typedef int_vector<1> bit_vector;

//------------------------------------------------------------------------------
// rank / select over plain bits (shared by bit_vector, bit_vector_il, sd_vector)
//------------------------------------------------------------------------------

namespace shim_detail
{

// Interleaved 512-bit blocks: [cumulative count, 8 data words].
struct RankedBits
{
  std::uint64_t n_bits = 0, ones = 0;
  std::vector<std::uint64_t> il;

  void build(const std::uint64_t* words, std::uint64_t bits)
  {
    n_bits = bits;
    std::uint64_t blocks = bits / 512 + 1, n_words = (bits + 63) / 64, cum = 0;
    il.assign(blocks * 9, 0);
    for(std::uint64_t b = 0; b < blocks; b++)
    {
      il[b * 9] = cum;
      for(std::uint64_t k = 0; k < 8; k++)
      {
        std::uint64_t idx = b * 8 + k, word = 0;
        if(idx < n_words)
        {
          word = words[idx];
          std::uint64_t rem = bits - idx * 64;
          if(rem < 64) { word &= (((std::uint64_t)1 << rem) - 1); }
        }
        il[b * 9 + 1 + k] = word;
        cum += __builtin_popcountll(word);
      }
    }
    ones = cum;
  }

  bool get(std::uint64_t i) const { return (il[(i >> 9) * 9 + 1 + ((i & 511) >> 6)] >> (i & 63)) & 1; }

  std::uint64_t rank(std::uint64_t i) const
  {
    const std::uint64_t* blk = il.data() + (i >> 9) * 9;
    std::uint64_t res = blk[0], full = (i & 511) >> 6, rem = i & 63;
    for(std::uint64_t k = 0; k < full; k++) { res += __builtin_popcountll(blk[1 + k]); }
    if(rem) { res += __builtin_popcountll(blk[1 + full] & (((std::uint64_t)1 << rem) - 1)); }
    return res;
  }

  std::uint64_t select(std::uint64_t k) const
  {
    std::uint64_t lo = 0, hi = il.size() / 9 - 1;
    while(lo < hi)
    {
      std::uint64_t mid = lo + (hi - lo + 1) / 2;
      if(il[mid * 9] < k) { lo = mid; } else { hi = mid - 1; }
    }
    const std::uint64_t* blk = il.data() + lo * 9;
    std::uint64_t need = k - blk[0];
    for(std::uint64_t j = 0; j < 8; j++)
    {
      std::uint64_t word = blk[1 + j], c = __builtin_popcountll(word);
      if(need <= c)
      {
        for(std::uint64_t t = 1; t < need; t++) { word &= word - 1; }
        return lo * 512 + j * 64 + __builtin_ctzll(word);
      }
      need -= c;
    }
    return n_bits;
  }
};

// select_support_mcl on disk: the number of arguments; if there are any, the positions of every
// 4096th argument, a bit per superblock (1 = short) unless all are short, and per superblock
// either all positions (long, and always the last partial one) or every 64th relative position.
inline void writeVarVector(std::ostream& out, const std::vector<std::uint64_t>& values, std::uint64_t count, std::uint8_t width)
{
  std::uint64_t bit_len = count * width;
  std::vector<std::uint64_t> packed((bit_len + 63) / 64 + 1, 0);
  for(std::uint64_t i = 0; i < values.size(); i++)
  {
    std::uint64_t bit = i * width, word = bit >> 6, off = bit & 63;
    packed[word] |= values[i] << off;
    if(off + width > 64) { packed[word + 1] |= values[i] >> (64 - off); }
  }
  write_member(bit_len, out); write_member(width, out);
  out.write(reinterpret_cast<const char*>(packed.data()), ((bit_len + 63) / 64) * sizeof(std::uint64_t));
}

inline void skipVarVector(std::istream& in)
{
  std::uint64_t bit_len = 0; std::uint8_t width = 0;
  read_member(bit_len, in); read_member(width, in);
  in.seekg(((bit_len + 63) / 64) * sizeof(std::uint64_t), std::ios::cur);
}

inline std::uint64_t topBit(std::uint64_t x) { return (x == 0 ? 0 : 63 - __builtin_clzll(x)); }

inline void writeSelectDirectory(std::ostream& out, const std::uint64_t* words, std::uint64_t n, bool ones)
{
  std::vector<std::uint64_t> pos;
  for(std::uint64_t i = 0; i < n; i++)
  {
    bool bit = (words[i >> 6] >> (i & 63)) & 1;
    if(bit == ones) { pos.push_back(i); }
  }
  std::uint64_t total = pos.size();
  write_member(total, out);
  if(total == 0) { return; }
  std::uint64_t logn = topBit(((n + 63) / 64) * 64) + 1, limit = logn * logn * logn * logn;
  std::uint64_t superblocks = (total + 4095) / 4096;
  std::vector<std::uint64_t> firsts(superblocks), kind(superblocks / 64 + 1, 0);
  bool mixed = false;
  for(std::uint64_t b = 0; b < superblocks; b++)
  {
    std::uint64_t from = b * 4096, to = std::min(from + 4096, total);
    firsts[b] = pos[from];
    bool is_short = (to - from == 4096 && pos[to - 1] - pos[from] <= limit);
    if(is_short) { kind[b >> 6] |= (std::uint64_t)1 << (b & 63); } else { mixed = true; }
  }
  writeVarVector(out, firsts, superblocks, (std::uint8_t)logn);
  std::uint64_t kind_bits = (mixed ? superblocks : 0);
  write_member(kind_bits, out);
  out.write(reinterpret_cast<const char*>(kind.data()), ((kind_bits + 63) / 64) * sizeof(std::uint64_t));
  for(std::uint64_t b = 0; b < superblocks; b++)
  {
    std::uint64_t from = b * 4096, to = std::min(from + 4096, total);
    std::vector<std::uint64_t> values;
    if((kind[b >> 6] >> (b & 63)) & 1)
    {
      for(std::uint64_t j = from; j < to; j += 64) { values.push_back(pos[j] - pos[from]); }
      writeVarVector(out, values, 64, (std::uint8_t)(topBit(pos[to - 1] - pos[from]) + 1));
    }
    else
    {
      values.assign(pos.begin() + from, pos.begin() + to);
      std::uint64_t widest = (to - from == 4096 ? pos[to - 1] : n - 1);
      writeVarVector(out, values, 4096, (std::uint8_t)(topBit(widest) + 1));
    }
  }
}

inline void skipSelectDirectory(std::istream& in)
{
  std::uint64_t total = 0;
  read_member(total, in);
  if(total == 0) { return; }
  skipVarVector(in);
  std::uint64_t kind_bits = 0;
  read_member(kind_bits, in);
  in.seekg(((kind_bits + 63) / 64) * sizeof(std::uint64_t), std::ios::cur);
  for(std::uint64_t b = 0; b < (total + 4095) / 4096; b++) { skipVarVector(in); }
}

} // namespace shim_detail

// Supports for bit_vector: they index the vector they were initialised with.
struct rank_support_shim
{
  shim_detail::RankedBits bits;
  rank_support_shim() {}
  explicit rank_support_shim(const bit_vector* v) { set_vector(v); }
  void set_vector(const bit_vector* v) { if(v != nullptr) { bits.build(v->data(), v->size()); } }
  std::uint64_t operator()(std::uint64_t i) const { return bits.rank(i); }
  std::uint64_t rank(std::uint64_t i) const { return bits.rank(i); }
  std::uint64_t size() const { return bits.n_bits; }
  void swap(rank_support_shim& o) { std::swap(bits, o.bits); }
  std::uint64_t serialize(std::ostream&, structure_tree_node* = nullptr, std::string = "") const { return 0; }
  void load(std::istream&, const bit_vector* v = nullptr) { set_vector(v); }
};

struct select_support_shim
{
  shim_detail::RankedBits bits;
  select_support_shim() {}
  explicit select_support_shim(const bit_vector* v) { set_vector(v); }
  void set_vector(const bit_vector* v) { if(v != nullptr) { bits.build(v->data(), v->size()); } }
  std::uint64_t operator()(std::uint64_t k) const { return bits.select(k); }
  std::uint64_t select(std::uint64_t k) const { return bits.select(k); }
  void swap(select_support_shim& o) { std::swap(bits, o.bits); }
  std::uint64_t serialize(std::ostream& out, structure_tree_node* = nullptr, std::string = "") const
  {
    std::vector<std::uint64_t> plain((bits.n_bits + 63) / 64 + 1, 0);
    for(std::uint64_t i = 0; i < bits.n_bits; i++) { if(bits.get(i)) { plain[i >> 6] |= (std::uint64_t)1 << (i & 63); } }
    shim_detail::writeSelectDirectory(out, plain.data(), bits.n_bits, true);
    return 0;
  }
  void load(std::istream& in, const bit_vector* v = nullptr) { shim_detail::skipSelectDirectory(in); set_vector(v); }
};

//------------------------------------------------------------------------------
// bit_vector_il<>: the interleaved vector answers rank itself; the support only points at it
//------------------------------------------------------------------------------

template<std::uint32_t B = 512>
class bit_vector_il
{
public:
  typedef std::uint64_t size_type;
  typedef bool value_type;

  struct rank_1_type
  {
    const bit_vector_il* vec = nullptr;
    rank_1_type() {}
    explicit rank_1_type(const bit_vector_il* v) : vec(v) {}
    void set_vector(const bit_vector_il* v) { vec = v; }
    std::uint64_t operator()(std::uint64_t i) const { return vec->bits.rank(i); }
    std::uint64_t rank(std::uint64_t i) const { return vec->bits.rank(i); }
    void swap(rank_1_type&) {}
    std::uint64_t serialize(std::ostream&, structure_tree_node* = nullptr, std::string = "") const { return 0; }
    void load(std::istream&, const bit_vector_il* v = nullptr) { vec = v; }
  };

  bit_vector_il() {}
  bit_vector_il(const bit_vector& v) { bits.build(v.data(), v.size()); }
  bit_vector_il& operator=(const bit_vector& v) { bits.build(v.data(), v.size()); return *this; }

  size_type size() const { return bits.n_bits; }
  bool operator[](size_type i) const { return bits.get(i); }
  void swap(bit_vector_il& o) { std::swap(bits, o.bits); }

  // On-disk layout of sdsl-lite's bit_vector_il: size, number of words, number of blocks, log of
  // the block size, the interleaved words ([count, 8 data words] per block, data cut at
  // (size + 64) / 64 words, then the total), and the select samples over the block counts.
  size_type serialize(std::ostream& out, structure_tree_node* = nullptr, std::string = "") const
  {
    std::uint64_t zero = 0;
    if(bits.il.empty())
    {
      for(int i = 0; i < 6; i++) { write_member(zero, out); }
      return 6 * sizeof(std::uint64_t);
    }
    std::uint64_t n = bits.n_bits, blocks = (n + B) / B, data_words = (n + 64) / 64, mem = data_words + blocks + 1, shift = bits::hi(B);
    std::vector<std::uint64_t> data;
    data.reserve(mem);
    for(std::uint64_t i = 0; i < data_words; i++)
    {
      if(i % 8 == 0) { data.push_back(bits.il[(i / 8) * 9]); }
      data.push_back(bits.il[(i / 8) * 9 + 1 + i % 8]);
    }
    data.push_back(bits.ones);
    std::uint64_t n_samples = (blocks > 2048 ? 1024 : std::max<std::uint64_t>(1, (std::uint64_t)1 << bits::hi(blocks)));
    std::vector<std::uint64_t> samples(n_samples, 0);
    {
      std::vector<std::pair<std::uint64_t, std::uint64_t>> todo(1, std::make_pair((std::uint64_t)0, blocks));
      for(std::uint64_t head = 0, idx = 0; head < todo.size() && idx < n_samples; head++)
      {
        std::uint64_t lb = todo[head].first, rb = todo[head].second, mid = lb + (rb - lb) / 2;
        samples[idx++] = (mid * 9 < data.size() ? data[mid * 9] : bits.ones);
        todo.push_back(std::make_pair(lb, mid)); todo.push_back(std::make_pair(mid + 1, rb));
      }
    }
    write_member(n, out); write_member(mem, out); write_member(blocks, out); write_member(shift, out);
    std::uint64_t data_bits = data.size() * 64, sample_bits = samples.size() * 64;
    write_member(data_bits, out);
    out.write(reinterpret_cast<const char*>(data.data()), data.size() * sizeof(std::uint64_t));
    write_member(sample_bits, out);
    out.write(reinterpret_cast<const char*>(samples.data()), samples.size() * sizeof(std::uint64_t));
    return (6 + data.size() + samples.size()) * sizeof(std::uint64_t);
  }
  void load(std::istream& in)
  {
    std::uint64_t n = 0, mem = 0, blocks = 0, shift = 0, data_bits = 0, sample_bits = 0;
    read_member(n, in); read_member(mem, in); read_member(blocks, in); read_member(shift, in);
    read_member(data_bits, in);
    std::vector<std::uint64_t> data(data_bits / 64 + 1, 0);
    in.read(reinterpret_cast<char*>(data.data()), (data_bits / 64) * sizeof(std::uint64_t));
    read_member(sample_bits, in);
    in.seekg((sample_bits / 64) * sizeof(std::uint64_t), std::ios::cur);
    if(data_bits == 0) { bits = shim_detail::RankedBits(); return; }
    std::vector<std::uint64_t> plain((n + 63) / 64 + 1, 0);
    for(std::uint64_t i = 0; i < (n + 63) / 64; i++) { plain[i] = data[i + i / 8 + 1]; }
    bits.build(plain.data(), n);
  }

  shim_detail::RankedBits bits;
};

//------------------------------------------------------------------------------
// sd_vector<>: a sorted list of positions (a multiset when built that way)
//------------------------------------------------------------------------------

class sd_vector_builder
{
public:
  sd_vector_builder() : n(0), m(0), multiset(false) {}
  sd_vector_builder(std::uint64_t size, std::uint64_t ones, bool is_multiset = false) : n(size), m(ones), multiset(is_multiset) { positions.reserve(ones); }
  void set(std::uint64_t i) { positions.push_back(i); }
  void set_unsafe(std::uint64_t i) { positions.push_back(i); }
  std::uint64_t size() const { return n; }
  std::uint64_t capacity() const { return m; }
  std::uint64_t items() const { return positions.size(); }
  std::uint64_t n, m; bool multiset;
  std::vector<std::uint64_t> positions;
};

template<class A = void, class B = void, class C = void>
class sd_vector
{
public:
  typedef std::uint64_t size_type;
  typedef bool value_type;

  struct rank_1_type
  {
    const sd_vector* vec = nullptr;
    rank_1_type() {}
    explicit rank_1_type(const sd_vector* v) : vec(v) {}
    void set_vector(const sd_vector* v) { vec = v; }
    std::uint64_t operator()(std::uint64_t i) const { return vec->rank(i); }
    std::uint64_t rank(std::uint64_t i) const { return vec->rank(i); }
    void swap(rank_1_type&) {}
    std::uint64_t serialize(std::ostream&, structure_tree_node* = nullptr, std::string = "") const { return 0; }
    void load(std::istream&, const sd_vector* v = nullptr) { vec = v; }
  };

  struct select_1_type
  {
    const sd_vector* vec = nullptr;
    select_1_type() {}
    explicit select_1_type(const sd_vector* v) : vec(v) {}
    void set_vector(const sd_vector* v) { vec = v; }
    std::uint64_t operator()(std::uint64_t k) const { return vec->positions[k - 1]; }
    std::uint64_t select(std::uint64_t k) const { return vec->positions[k - 1]; }
    void swap(select_1_type&) {}
    std::uint64_t serialize(std::ostream&, structure_tree_node* = nullptr, std::string = "") const { return 0; }
    void load(std::istream&, const sd_vector* v = nullptr) { vec = v; }
  };

  // (rank, position) pairs, as returned by successor()
  struct one_iterator
  {
    std::pair<std::uint64_t, std::uint64_t> value;
    const std::pair<std::uint64_t, std::uint64_t>* operator->() const { return &value; }
    const std::pair<std::uint64_t, std::uint64_t>& operator*() const { return value; }
  };

  sd_vector() : n(0) {}
  sd_vector(const bit_vector& v) : n(v.size())
  {
    for(std::uint64_t w = 0; w * 64 < n; w++)
    {
      std::uint64_t word = v.data()[w];
      std::uint64_t rem = n - w * 64;
      if(rem < 64) { word &= (((std::uint64_t)1 << rem) - 1); }
      while(word) { positions.push_back(w * 64 + __builtin_ctzll(word)); word &= word - 1; }
    }
  }
  template<class Iterator>
  sd_vector(Iterator begin, Iterator end) : n(0)
  {
    for(Iterator it = begin; it != end; ++it) { positions.push_back(*it); }
    if(!positions.empty()) { n = positions.back() + 1; }
  }
  sd_vector(sd_vector_builder& builder) : n(builder.n) { positions.swap(builder.positions); }

  size_type size() const { return n; }
  size_type ones() const { return positions.size(); }
  bool operator[](size_type i) const { return std::binary_search(positions.begin(), positions.end(), i); }
  std::uint64_t rank(std::uint64_t i) const { return std::lower_bound(positions.begin(), positions.end(), i) - positions.begin(); }

  // First one at a position >= i: (its rank, its position); (ones(), size()) if there is none.
  one_iterator successor(std::uint64_t i) const
  {
    one_iterator res;
    std::uint64_t r = this->rank(i);
    res.value = (r < positions.size() ? std::make_pair(r, positions[r]) : std::make_pair((std::uint64_t)positions.size(), n));
    return res;
  }

  void swap(sd_vector& o) { std::swap(n, o.n); positions.swap(o.positions); }

  // On-disk layout of sdsl-lite's sd_vector (Elias-Fano): size, width of the low parts, the low
  // parts, the unary-coded high parts, and the select directories for the ones and zeros of those.
  size_type serialize(std::ostream& out, structure_tree_node* = nullptr, std::string = "") const
  {
    std::uint64_t m = positions.size(), zero = 0;
    if(n == 0 && m == 0)
    {
      std::uint8_t wl = 0, width = 64;
      write_member(zero, out); write_member(wl, out);
      write_member(zero, out); write_member(width, out);   // low
      write_member(zero, out);                             // high
      write_member(zero, out); write_member(zero, out);    // directories
      return 0;
    }
    std::uint64_t logm = bits::hi(m) + 1, logn = bits::hi(n) + 1;
    if(logm == logn) { logm--; }
    std::uint8_t wl = (std::uint8_t)(logn - logm);
    std::uint64_t high_bits = m + ((std::uint64_t)1 << logm);
    std::vector<std::uint64_t> low(m), high(high_bits / 64 + 2, 0);
    for(std::uint64_t i = 0; i < m; i++)
    {
      low[i] = positions[i] & (((std::uint64_t)1 << wl) - 1);
      std::uint64_t h = (positions[i] >> wl) + i;
      high[h >> 6] |= (std::uint64_t)1 << (h & 63);
    }
    write_member(n, out); write_member(wl, out);
    shim_detail::writeVarVector(out, low, m, wl);
    write_member(high_bits, out);
    out.write(reinterpret_cast<const char*>(high.data()), ((high_bits + 63) / 64) * sizeof(std::uint64_t));
    shim_detail::writeSelectDirectory(out, high.data(), high_bits, true);
    shim_detail::writeSelectDirectory(out, high.data(), high_bits, false);
    return 0;
  }
  void load(std::istream& in)
  {
    std::uint8_t wl = 0, width = 0;
    std::uint64_t low_bits = 0, high_bits = 0;
    read_member(n, in); read_member(wl, in);
    read_member(low_bits, in); read_member(width, in);
    std::vector<std::uint64_t> low((low_bits + 63) / 64 + 1, 0);
    in.read(reinterpret_cast<char*>(low.data()), ((low_bits + 63) / 64) * sizeof(std::uint64_t));
    read_member(high_bits, in);
    std::vector<std::uint64_t> high((high_bits + 63) / 64 + 1, 0);
    in.read(reinterpret_cast<char*>(high.data()), ((high_bits + 63) / 64) * sizeof(std::uint64_t));
    shim_detail::skipSelectDirectory(in); shim_detail::skipSelectDirectory(in);
    positions.clear();
    for(std::uint64_t p = 0, i = 0; p < high_bits; p++)
    {
      if(!((high[p >> 6] >> (p & 63)) & 1)) { continue; }
      std::uint64_t bit = i * wl, word = bit >> 6, off = bit & 63, x = 0;
      if(wl > 0)
      {
        x = low[word] >> off;
        if(off + wl > 64) { x |= low[word + 1] << (64 - off); }
        x &= (((std::uint64_t)1 << wl) - 1);
      }
      positions.push_back(((p - i) << wl) | x);
      i++;
    }
  }

  std::uint64_t n;
  std::vector<std::uint64_t> positions;
};

//------------------------------------------------------------------------------
// RAM file system, int_vector_buffer, store / load
//------------------------------------------------------------------------------

struct ram_fs
{
  static std::map<std::string, std::string>& files() { static std::map<std::string, std::string> f; return f; }
  static void remove(const std::string& name) { files().erase(name); }
};

inline std::string ram_file_name(const std::string& name) { return "@" + name; }
inline bool is_ram_file(const std::string& name) { return !name.empty() && name[0] == '@'; }

template<class T>
bool store_to_file(const T& object, const std::string& name)
{
  if(is_ram_file(name))
  {
    std::ostringstream out;
    object.serialize(out);
    ram_fs::files()[name] = out.str();
    return true;
  }
  std::ofstream out(name, std::ios_base::binary);
  if(!out) { return false; }
  object.serialize(out);
  return (bool)out;
}

template<class T>
bool load_from_file(T& object, const std::string& name)
{
  if(is_ram_file(name))
  {
    auto it = ram_fs::files().find(name);
    if(it == ram_fs::files().end()) { return false; }
    std::istringstream in(it->second);
    object.load(in);
    return true;
  }
  std::ifstream in(name, std::ios_base::binary);
  if(!in) { return false; }
  object.load(in);
  return (bool)in;
}

template<std::uint8_t W>
class int_vector_buffer
{
public:
  explicit int_vector_buffer(const std::string& name) { load_from_file(data, name); }
  std::uint64_t size() const { return data.size(); }
  std::uint64_t operator[](std::uint64_t i) const { return data[i]; }
  int_vector<W> data;
};

//------------------------------------------------------------------------------
// wt_blcd<>: only range minimum (quantile_freq(wt, l, r, 0)) is asked of it
//------------------------------------------------------------------------------

template<class A = void, class B = void, class C = void, class D = void>
class wt_blcd
{
public:
  typedef std::uint64_t size_type;
  typedef std::uint8_t value_type;

  wt_blcd() {}
  wt_blcd(int_vector_buffer<8>& buffer, size_type size)
  {
    levels.emplace_back(size);
    for(size_type i = 0; i < size; i++) { levels[0][i] = (std::uint8_t)buffer[i]; }
    for(size_type span = 1; 2 * span <= size; span *= 2)
    {
      const std::vector<std::uint8_t>& prev = levels.back();
      std::vector<std::uint8_t> next(size - 2 * span + 1);
      for(size_type i = 0; i < next.size(); i++) { next[i] = std::min(prev[i], prev[i + span]); }
      levels.push_back(std::move(next));
    }
  }
  size_type size() const { return (levels.empty() ? 0 : levels[0].size()); }
  std::uint8_t operator[](size_type i) const { return levels[0][i]; }
  void swap(wt_blcd& o) { levels.swap(o.levels); }

  // minimum of [l, r] (closed)
  std::uint8_t range_min(size_type l, size_type r) const
  {
    size_type len = r - l + 1;
    size_type k = 63 - __builtin_clzll(len);
    return std::min(levels[k][l], levels[k][r + 1 - ((size_type)1 << k)]);
  }

  size_type serialize(std::ostream&, structure_tree_node* = nullptr, std::string = "") const { return 0; }
  void load(std::istream&) {}

  std::vector<std::vector<std::uint8_t>> levels;   // sparse table
};

// q-th smallest value of wt[l, r] and its frequency; only q = 0 (the minimum) is supported.
template<class WT>
std::pair<std::uint64_t, std::uint64_t> quantile_freq(const WT& wt, std::uint64_t l, std::uint64_t r, std::uint64_t /* q = 0 */)
{
  return std::make_pair((std::uint64_t)wt.range_min(l, r), (std::uint64_t)1);
}

//------------------------------------------------------------------------------
// util
//------------------------------------------------------------------------------

namespace util
{

template<class T> void clear(T& x) { T empty; x.swap(empty); }
template<class T> void clear(std::vector<T>& x) { std::vector<T>().swap(x); }

template<class Support, class Vector> void init_support(Support& support, const Vector* vec) { Support temp(vec); support.swap(temp); support.set_vector(vec); }
template<class Support, class Vector> void swap_support(Support& a, Support& b, const Vector* va, const Vector* vb) { a.swap(b); a.set_vector(va); b.set_vector(vb); }

template<class T> std::string class_name(const T&) { return typeid(T).name(); }
template<class T> std::string to_string(const T& x) { std::ostringstream s; s << x; return s.str(); }
inline std::uint64_t pid() { return (std::uint64_t)getpid(); }

// Smallest width that holds every value.
template<std::uint8_t W>
void bit_compress(int_vector<W>& v)
{
  std::uint64_t max = 0;
  for(std::uint64_t i = 0; i < v.size(); i++) { max = std::max<std::uint64_t>(max, v[i]); }
  std::uint8_t width = (std::uint8_t)(bits::hi(max) + 1);
  int_vector<W> temp(v.size(), 0, width);
  for(std::uint64_t i = 0; i < v.size(); i++) { temp[i] = (std::uint64_t)v[i]; }
  v.swap(temp);
}

template<class T> void assign(T& a, T& b) { a.swap(b); }

} // namespace util

template<class T> std::uint64_t size_in_bytes(const T& object)
{
  std::ostringstream out;
  return object.serialize(out);
}

} // namespace sdsl

#endif // GCSA2_B200_SDSL_SHIM_HPP
