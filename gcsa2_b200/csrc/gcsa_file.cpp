/*
  gcsa_file.cpp -- reads and writes the reference's index files (.gcsa, .lcp) as flat arrays.

  Replaces, on the host, GCSA::load / GCSA::serialize (src/gcsa.cpp:140-216) and
  LCPArray::load / serialize (src/lcp.cpp:116-143).  The field order is the reference's; the byte
  layout of each SDSL member (int_vector, bit_vector_il<512>, sd_vector, select_support_mcl) is not
  in the reference tree -- it is restated here from the published sdsl-lite sources (v2.1.1 line,
  which the vgteam fork keeps) and has NOT been checked against a real SDSL build in this
  environment.  To compensate the reader validates everything that can be validated (cumulative
  counts of the interleaved blocks, Elias-Fano monotonicity, sizes against the header, C[] against
  the popcounts, exact end of file) and fails loudly instead of guessing.

  The rank / select supports are skipped on load (the engine builds its own device rank
  dictionary) and regenerated on write.
*/
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gcsa2_b200.h"
#include "internal.h"

namespace
{

typedef uint64_t u64;
constexpr int SIGMA = GCSA_B200_SIGMA;
constexpr uint32_t GCSA_TAG = 0x6C5A6C5A, GCSA_VERSION = 3;     // include/gcsa/files.h:143-144, utils.h:161-166
constexpr uint32_t LCP_TAG = 0x6C5A7C94, LCP_VERSION = 1;       // include/gcsa/files.h:178-179
constexpr u64 IL_BLOCK = 512;                                   // bit_vector_il<> default block size (gcsa.h:46)
constexpr u64 MCL_SUPERBLOCK = 4096;

inline u64 wordsFor(u64 bits) { return (bits + 63) / 64; }
inline u64 bitLength(u64 x) { return (x == 0 ? 1 : 64 - __builtin_clzll(x)); }
inline u64 hi(u64 x) { return (x == 0 ? 0 : 63 - __builtin_clzll(x)); }      // sdsl::bits::hi
inline bool getBit(const std::vector<u64>& v, u64 i) { return (v[i >> 6] >> (i & 63)) & 1; }
inline void setBit(std::vector<u64>& v, u64 i) { v[i >> 6] |= (u64)1 << (i & 63); }

struct Failure { std::string what; };
[[noreturn]] void bad(const std::string& what) { throw Failure{what}; }

//------------------------------------------------------------------------------
// Streams
//------------------------------------------------------------------------------

struct In
{
  FILE* f;
  explicit In(const char* path) : f(std::fopen(path, "rb")) { if(f == nullptr) { bad(std::string("cannot open ") + path); } }
  ~In() { if(f != nullptr) { std::fclose(f); } }
  void read(void* p, size_t bytes) { if(bytes > 0 && std::fread(p, 1, bytes, f) != bytes) { bad("unexpected end of file"); } }
  template<class T> T get() { T t; read(&t, sizeof(T)); return t; }
  void skip(u64 bytes) { if(std::fseek(f, (long)bytes, SEEK_CUR) != 0) { bad("seek failed"); } }
  bool atEnd() { int c = std::fgetc(f); if(c == EOF) { return true; } std::ungetc(c, f); return false; }
};

struct Out
{
  FILE* f;
  explicit Out(const char* path) : f(std::fopen(path, "wb")) { if(f == nullptr) { bad(std::string("cannot create ") + path); } }
  ~Out() { if(f != nullptr) { std::fclose(f); } }
  void write(const void* p, size_t bytes) { if(bytes > 0 && std::fwrite(p, 1, bytes, f) != bytes) { bad("write failed"); } }
  template<class T> void put(const T& t) { write(&t, sizeof(T)); }
  void close() { if(std::fclose(f) != 0) { f = nullptr; bad("close failed"); } f = nullptr; }
};

//------------------------------------------------------------------------------
// sdsl::int_vector<w>: u64 size in bits, (w == 0: u8 width), ceil(bits / 64) words.
//------------------------------------------------------------------------------

struct IntVector
{
  u64 bits = 0; uint8_t width = 64;
  std::vector<u64> words;                // one spare word at the end
  u64 size() const { return bits / width; }
  u64 get(u64 i) const
  {
    u64 bit = i * width, word = bit >> 6, off = bit & 63;
    u64 x = words[word] >> off;
    if(off + width > 64) { x |= words[word + 1] << (64 - off); }
    return (width == 64 ? x : x & (((u64)1 << width) - 1));
  }
  void set(u64 i, u64 x)
  {
    u64 bit = i * width, word = bit >> 6, off = bit & 63;
    words[word] |= x << off;
    if(off + width > 64) { words[word + 1] |= x >> (64 - off); }
  }
  static IntVector zeros(u64 n, uint8_t w) { IntVector v; v.width = w; v.bits = n * w; v.words.assign(wordsFor(v.bits) + 1, 0); return v; }
};

IntVector readIntVector(In& in, uint8_t fixed_width)
{
  IntVector v;
  v.bits = in.get<u64>();
  v.width = (fixed_width == 0 ? in.get<uint8_t>() : fixed_width);
  if(v.width == 0 || v.width > 64) { bad("int_vector: width out of range"); }
  if(v.bits > ((u64)1 << 48)) { bad("int_vector: implausible size"); }
  v.words.assign(wordsFor(v.bits) + 1, 0);
  in.read(v.words.data(), wordsFor(v.bits) * sizeof(u64));
  return v;
}

void skipIntVector(In& in, uint8_t fixed_width)
{
  u64 bits = in.get<u64>();
  if(fixed_width == 0) { uint8_t w = in.get<uint8_t>(); if(w == 0 || w > 64) { bad("int_vector: width out of range"); } }
  if(bits > ((u64)1 << 48)) { bad("int_vector: implausible size"); }
  in.skip(wordsFor(bits) * sizeof(u64));
}

void writeIntVector(Out& out, const IntVector& v, bool dynamic_width)
{
  out.put<u64>(v.bits);
  if(dynamic_width) { out.put<uint8_t>(v.width); }
  out.write(v.words.data(), wordsFor(v.bits) * sizeof(u64));
}

// A plain bit vector (sdsl::bit_vector = int_vector<1>) as words with one spare word.
struct Bits { u64 n = 0; std::vector<u64> words; };

Bits readBitVector(In& in)
{
  IntVector v = readIntVector(in, 1);
  Bits b; b.n = v.bits; b.words.swap(v.words);
  if(b.n & 63) { b.words[b.n >> 6] &= ((u64)1 << (b.n & 63)) - 1; }
  return b;
}

void writeBitVector(Out& out, const u64* words, u64 n)
{
  out.put<u64>(n);
  out.write(words, wordsFor(n) * sizeof(u64));
}

u64 popcount(const Bits& b)
{
  u64 r = 0;
  for(u64 i = 0; i < wordsFor(b.n); i++) { r += __builtin_popcountll(b.words[i]); }
  return r;
}

//------------------------------------------------------------------------------
// sdsl::select_support_mcl<b, 1>: u64 arg_cnt; if nonzero: int_vector<0> superblock, bit_vector
// mini_or_long, then per superblock of 4096 args one int_vector<0> (explicit positions for a long
// superblock, every 64th position relative to the first for a short one).
//------------------------------------------------------------------------------

void skipSelect(In& in, u64 expected_args)
{
  u64 args = in.get<u64>();
  if(args != expected_args) { bad("select_support_mcl: argument count does not match its vector"); }
  if(args == 0) { return; }
  u64 sb = (args + MCL_SUPERBLOCK - 1) / MCL_SUPERBLOCK;
  skipIntVector(in, 0);
  Bits mini_or_long = readBitVector(in);
  if(mini_or_long.n != 0 && mini_or_long.n != sb) { bad("select_support_mcl: bad superblock directory"); }
  for(u64 i = 0; i < sb; i++) { skipIntVector(in, 0); }
}

// bit == true: select_1, else select_0 over the n bits of `words`.
void writeSelect(Out& out, const u64* words, u64 n, bool bit)
{
  std::vector<u64> args;
  for(u64 w = 0; w < wordsFor(n); w++)
  {
    u64 word = (bit ? words[w] : ~words[w]);
    if((w + 1) * 64 > n) { u64 rem = n - w * 64; word &= (rem == 64 ? ~(u64)0 : (((u64)1 << rem) - 1)); }
    while(word) { args.push_back(w * 64 + __builtin_ctzll(word)); word &= word - 1; }
  }
  out.put<u64>(args.size());
  if(args.empty()) { return; }

  u64 capacity = wordsFor(n) * 64, logn = hi(capacity) + 1, logn4 = logn * logn * logn * logn;
  u64 sb = (args.size() + MCL_SUPERBLOCK - 1) / MCL_SUPERBLOCK;
  IntVector superblock = IntVector::zeros(sb, (uint8_t)logn);
  std::vector<IntVector> blocks(sb);
  std::vector<u64> is_mini(wordsFor(sb) + 1, 0);
  bool any_long = false;
  for(u64 s = 0; s < sb; s++)
  {
    u64 first = s * MCL_SUPERBLOCK, count = std::min<u64>(MCL_SUPERBLOCK, args.size() - first);
    superblock.set(s, args[first]);
    u64 diff = args[first + count - 1] - args[first];
    if(count < MCL_SUPERBLOCK)                 // the last, partial superblock is always stored explicitly
    {
      blocks[s] = IntVector::zeros(MCL_SUPERBLOCK, (uint8_t)(hi(n - 1) + 1));
      for(u64 j = 0; j < count; j++) { blocks[s].set(j, args[first + j]); }
      any_long = true;
    }
    else if(diff > logn4)
    {
      blocks[s] = IntVector::zeros(MCL_SUPERBLOCK, (uint8_t)(hi(args[first + count - 1]) + 1));
      for(u64 j = 0; j < count; j++) { blocks[s].set(j, args[first + j]); }
      any_long = true;
    }
    else
    {
      blocks[s] = IntVector::zeros(64, (uint8_t)(hi(diff) + 1));
      for(u64 j = 0; j < MCL_SUPERBLOCK; j += 64) { blocks[s].set(j / 64, args[first + j] - args[first]); }
      setBit(is_mini, s);
    }
  }
  writeIntVector(out, superblock, true);
  writeBitVector(out, is_mini.data(), any_long ? sb : 0);
  for(u64 s = 0; s < sb; s++) { writeIntVector(out, blocks[s], true); }
}

//------------------------------------------------------------------------------
// sdsl::bit_vector_il<512>: u64 size, block_num, superblocks, block_shift; int_vector<64> data =
// [cumulative count, 8 data words] per 512-bit block plus the final count; int_vector<64>
// rank_samples (used only by select on the interleaved vector, which GCSA never calls).
//------------------------------------------------------------------------------

Bits readInterleaved(In& in)
{
  Bits b;
  b.n = in.get<u64>();
  in.get<u64>();                                   // block_num (derived)
  in.get<u64>();                                   // superblocks (derived)
  u64 shift = in.get<u64>();
  IntVector data = readIntVector(in, 64);
  skipIntVector(in, 64);
  b.words.assign(wordsFor(b.n) + 1, 0);
  if(b.n == 0) { return b; }
  if(shift != hi(IL_BLOCK)) { bad("bit_vector_il: block size is not 512"); }
  u64 n_words = wordsFor(b.n), cum = 0, per = IL_BLOCK / 64;
  if(data.size() < n_words + (n_words + per - 1) / per) { bad("bit_vector_il: data too short"); }
  for(u64 i = 0; i < n_words; i++)
  {
    u64 pos = i + i / per + 1;
    if(i % per == 0 && data.words[pos - 1] != cum) { bad("bit_vector_il: cumulative counts do not match the data (layout mismatch)"); }
    b.words[i] = data.words[pos];
    cum += __builtin_popcountll(b.words[i]);
  }
  if(b.n & 63)
  {
    if(b.words[b.n >> 6] >> (b.n & 63)) { bad("bit_vector_il: bits set beyond the end"); }
  }
  return b;
}

void writeInterleaved(Out& out, const u64* words, u64 n)
{
  if(words == nullptr || n == 0)                   // default-constructed vector (the unused comps)
  {
    for(int i = 0; i < 4; i++) { out.put<u64>(0); }
    out.put<u64>(0); out.put<u64>(0);              // empty data, empty rank_samples
    return;
  }
  u64 per = IL_BLOCK / 64, superblocks = (n + IL_BLOCK) / IL_BLOCK, blocks = (n + 64) / 64, mem = blocks + superblocks + 1;
  IntVector data = IntVector::zeros(mem, 64);
  u64 j = 0, cum = 0, n_words = wordsFor(n);
  for(u64 i = 0; i < blocks; i++)
  {
    if(i % per == 0) { data.words[j++] = cum; }
    u64 word = (i < n_words ? words[i] : 0);
    if(i == (n >> 6) && (n & 63)) { word &= ((u64)1 << (n & 63)) - 1; }
    data.words[j++] = word;
    cum += __builtin_popcountll(word);
  }
  data.words[j] = cum;
  // rank samples: breadth-first layout of a binary search over the block counts
  u64 samples_n = (superblocks > 2048 ? 1024 : std::max<u64>(1, (u64)1 << hi(superblocks)));
  IntVector samples = IntVector::zeros(samples_n, 64);
  {
    std::vector<std::pair<u64, u64>> queue(1, std::make_pair((u64)0, superblocks));
    u64 idx = 0;
    for(u64 head = 0; head < queue.size() && idx < samples_n; head++)
    {
      u64 lb = queue[head].first, rb = queue[head].second, mid = lb + (rb - lb) / 2;
      u64 pos = mid * per + mid;
      samples.words[idx++] = (pos < mem ? data.words[pos] : cum);
      queue.push_back(std::make_pair(lb, mid)); queue.push_back(std::make_pair(mid + 1, rb));
    }
  }
  out.put<u64>(n); out.put<u64>(mem); out.put<u64>(superblocks); out.put<u64>(hi(IL_BLOCK));
  writeIntVector(out, data, false);
  writeIntVector(out, samples, false);
}

//------------------------------------------------------------------------------
// sdsl::sd_vector<>: u64 size, u8 wl, int_vector<0> low, bit_vector high, select_1 and select_0
// supports of high.  Element i = ((position of the i-th one of high) - i) << wl | low[i].
//------------------------------------------------------------------------------

Bits readSparse(In& in)
{
  Bits b;
  b.n = in.get<u64>();
  uint8_t wl = in.get<uint8_t>();
  IntVector low = readIntVector(in, 0);
  Bits high = readBitVector(in);
  u64 m = popcount(high);
  skipSelect(in, m);
  skipSelect(in, high.n - m);
  b.words.assign(wordsFor(b.n) + 1, 0);
  if(m == 0) { return b; }
  if(wl >= 64 || low.width != wl || low.size() != m) { bad("sd_vector: low part does not match the high part"); }
  u64 i = 0, prev = 0;
  for(u64 w = 0; w < wordsFor(high.n); w++)
  {
    u64 word = high.words[w];
    while(word)
    {
      u64 p = w * 64 + __builtin_ctzll(word); word &= word - 1;
      u64 value = ((p - i) << wl) | low.get(i);
      if(value >= b.n || (i > 0 && value <= prev)) { bad("sd_vector: positions are not increasing inside the vector"); }
      setBit(b.words, value);
      prev = value; i++;
    }
  }
  return b;
}

void writeSparse(Out& out, const u64* words, u64 n)
{
  std::vector<u64> ones;
  if(words != nullptr)
  {
    for(u64 w = 0; w < wordsFor(n); w++)
    {
      u64 word = words[w];
      if((w + 1) * 64 > n) { u64 rem = n - w * 64; word &= (rem == 64 ? ~(u64)0 : (((u64)1 << rem) - 1)); }
      while(word) { ones.push_back(w * 64 + __builtin_ctzll(word)); word &= word - 1; }
    }
  }
  if(words == nullptr || n == 0)                   // default-constructed
  {
    out.put<u64>(0); out.put<uint8_t>(0);
    IntVector low; low.bits = 0; low.width = 64; low.words.assign(1, 0);
    writeIntVector(out, low, true);
    out.put<u64>(0);                               // high
    out.put<u64>(0); out.put<u64>(0);              // selects
    return;
  }
  u64 m = ones.size(), logm = hi(m) + 1, logn = hi(n) + 1;
  if(logm == logn) { logm--; }
  uint8_t wl = (uint8_t)(logn - logm);
  IntVector low = IntVector::zeros(m, wl);
  u64 high_n = m + ((u64)1 << logm);
  std::vector<u64> high(wordsFor(high_n) + 1, 0);
  for(u64 i = 0; i < m; i++)
  {
    low.set(i, ones[i] & (((u64)1 << wl) - 1));
    setBit(high, (ones[i] >> wl) + i);
  }
  out.put<u64>(n); out.put<uint8_t>(wl);
  writeIntVector(out, low, true);
  writeBitVector(out, high.data(), high_n);
  writeSelect(out, high.data(), high_n, true);
  writeSelect(out, high.data(), high_n, false);
}

u64* release(std::vector<u64>& v)
{
  u64* p = (u64*)std::malloc(sizeof(u64) * (v.size() + 1));
  if(p == nullptr) { bad("out of memory"); }
  if(!v.empty()) { std::memcpy(p, v.data(), sizeof(u64) * v.size()); }
  p[v.size()] = 0;
  std::vector<u64>().swap(v);
  return p;
}

void loadGCSA(const char* path, gcsa_b200_built* result)
{
  In in(path);
  gcsa_flat_index& f = result->index;

  // GCSAHeader, src/files.cpp:513-538
  uint32_t tag = in.get<uint32_t>(), version = in.get<uint32_t>();
  f.path_nodes = in.get<u64>(); f.edge_count = in.get<u64>(); f.order = in.get<u64>();
  u64 flags = in.get<u64>();
  if(tag != GCSA_TAG) { bad("not a GCSA file (tag mismatch)"); }
  if(version != GCSA_VERSION || flags != 0) { bad("unsupported GCSA file version " + std::to_string(version) + " (expected 3, flags 0)"); }

  // Alphabet, src/support.cpp:228-250
  IntVector char2comp = readIntVector(in, 8), comp2char = readIntVector(in, 8), C = readIntVector(in, 64);
  f.sigma = in.get<u64>(); f.fast_chars = in.get<u64>();
  if(f.sigma != (u64)SIGMA || f.fast_chars != GCSA_B200_FAST_CHARS || char2comp.size() != 256 || comp2char.size() != f.sigma || C.size() != f.sigma + 1)
  {
    bad("alphabet is not the GCSA2 default shape (sigma 7, 4 fast characters)");
  }
  for(u64 i = 0; i < 256; i++) { f.char2comp[i] = (uint8_t)char2comp.get(i); if(f.char2comp[i] >= SIGMA) { bad("char2comp out of range"); } }
  for(int c = 0; c <= SIGMA; c++) { f.C[c] = C.get(c); }

  // fast_bwt[sigma], (fast_rank: no bytes), sparse_bwt[sigma], (sparse_rank: no bytes), src/gcsa.cpp:149-165
  std::vector<Bits> fast(SIGMA), sparse(SIGMA);
  for(int c = 0; c < SIGMA; c++) { fast[c] = readInterleaved(in); }
  for(int c = 0; c < SIGMA; c++) { sparse[c] = readSparse(in); }
  for(int c = 0; c < SIGMA; c++)
  {
    bool is_fast = (c >= 1 && c <= GCSA_B200_FAST_CHARS);
    Bits& b = (is_fast ? fast[c] : sparse[c]);
    if(b.n != f.path_nodes) { bad("BWT vector " + std::to_string(c) + " has the wrong length"); }
    if(popcount(b) != f.C[c + 1] - f.C[c]) { bad("C[] does not match BWT vector " + std::to_string(c)); }
    f.bwt[c] = release(b.words);
  }
  if(f.C[0] != 0 || f.C[SIGMA] != f.edge_count) { bad("C[] does not match the edge count"); }

  Bits edges = readInterleaved(in), sampled = readInterleaved(in);
  if(edges.n != f.edge_count || popcount(edges) != f.path_nodes) { bad("edges vector does not match the header"); }
  if(sampled.n != f.path_nodes) { bad("sampled_paths has the wrong length"); }
  u64 sampled_nodes = popcount(sampled);
  f.edges = release(edges.words); f.sampled_paths = release(sampled.words);

  IntVector stored = readIntVector(in, 0);
  Bits samples = readBitVector(in);
  if(samples.n != stored.size() || popcount(samples) != sampled_nodes) { bad("samples do not match sampled_paths"); }
  skipSelect(in, sampled_nodes);
  f.sample_count = stored.size();
  std::vector<u64> values(f.sample_count);
  for(u64 i = 0; i < f.sample_count; i++) { values[i] = stored.get(i); }
  f.stored_samples = release(values);
  f.samples = release(samples.words);

  // SadaSparse (filter, values; their supports are empty), SadaCount (data + select), src/support.cpp:400-418, 492-516
  Bits filter = readSparse(in), extra = readSparse(in);
  if(filter.n != f.path_nodes || popcount(filter) != popcount(extra)) { bad("extra_pointers do not match the path nodes"); }
  f.extra_filter = release(filter.words);
  f.extra_values_len = extra.n; f.extra_values = release(extra.words);
  Bits redundant = readBitVector(in);
  u64 red_ones = popcount(redundant);
  skipSelect(in, red_ones);
  if(f.path_nodes > 0 && red_ones != f.path_nodes - 1) { bad("redundant_pointers do not match the path nodes"); }
  f.redundant_len = redundant.n; f.redundant = release(redundant.words);

  if(!in.atEnd()) { bad("trailing bytes after the index"); }
}

void writeGCSA(const gcsa_flat_index* f, const char* path)
{
  if(f->sigma != (u64)SIGMA || f->fast_chars != GCSA_B200_FAST_CHARS) { bad("alphabet is not the GCSA2 default shape"); }
  Out out(path);
  out.put<uint32_t>(GCSA_TAG); out.put<uint32_t>(GCSA_VERSION);
  out.put<u64>(f->path_nodes); out.put<u64>(f->edge_count); out.put<u64>(f->order); out.put<u64>(0);

  IntVector char2comp = IntVector::zeros(256, 8), comp2char = IntVector::zeros(SIGMA, 8), C = IntVector::zeros(SIGMA + 1, 64);
  for(u64 i = 0; i < 256; i++) { char2comp.set(i, f->char2comp[i]); }
  const char* chars = "$ACGTN#";                  // src/support.cpp:92
  for(int c = 0; c < SIGMA; c++) { comp2char.set(c, (unsigned char)chars[c]); }
  for(int c = 0; c <= SIGMA; c++) { C.set(c, f->C[c]); }
  writeIntVector(out, char2comp, false); writeIntVector(out, comp2char, false); writeIntVector(out, C, false);
  out.put<u64>(f->sigma); out.put<u64>(f->fast_chars);

  for(int c = 0; c < SIGMA; c++)
  {
    bool is_fast = (c >= 1 && c <= GCSA_B200_FAST_CHARS);
    writeInterleaved(out, is_fast ? f->bwt[c] : nullptr, is_fast ? f->path_nodes : 0);
  }
  for(int c = 0; c < SIGMA; c++)
  {
    bool is_fast = (c >= 1 && c <= GCSA_B200_FAST_CHARS);
    writeSparse(out, is_fast ? nullptr : f->bwt[c], is_fast ? 0 : f->path_nodes);
  }
  writeInterleaved(out, f->edges, f->edge_count);
  writeInterleaved(out, f->sampled_paths, f->path_nodes);

  u64 max_sample = 0;
  for(u64 i = 0; i < f->sample_count; i++) { max_sample = std::max(max_sample, f->stored_samples[i]); }
  IntVector stored = IntVector::zeros(f->sample_count, (uint8_t)bitLength(max_sample));   // src/gcsa.cpp:702-703
  for(u64 i = 0; i < f->sample_count; i++) { stored.set(i, f->stored_samples[i]); }
  writeIntVector(out, stored, true);
  writeBitVector(out, f->samples, f->sample_count);
  writeSelect(out, f->samples, f->sample_count, true);

  writeSparse(out, f->extra_filter, f->path_nodes);
  writeSparse(out, f->extra_values, f->extra_values_len);
  writeBitVector(out, f->redundant, f->redundant_len);
  writeSelect(out, f->redundant, f->redundant_len, true);
  out.close();
}

void loadLCP(const char* path, gcsa_flat_lcp* lcp)
{
  In in(path);
  // LCPHeader, src/files.cpp:581-604
  uint32_t tag = in.get<uint32_t>(), version = in.get<uint32_t>();
  lcp->size = in.get<u64>(); lcp->branching = in.get<u64>();
  u64 flags = in.get<u64>();
  if(tag != LCP_TAG) { bad("not an LCP file (tag mismatch)"); }
  if(version != LCP_VERSION || flags != 0) { bad("unsupported LCP file version " + std::to_string(version)); }
  if(lcp->branching < 2) { bad("LCP branching factor below 2"); }
  IntVector data = readIntVector(in, 0), offsets = readIntVector(in, 64);
  if(!in.atEnd()) { bad("trailing bytes after the LCP array"); }
  if(data.width > 8) { bad("LCP values wider than a byte"); }
  if(offsets.size() < 1) { bad("LCP offsets missing"); }
  lcp->levels = offsets.size() - 1;
  // src/lcp.cpp:224-258: level 0 has `size` values, each further level ceil(previous / branching)
  u64 expect = 0, level_size = lcp->size;
  for(u64 l = 0; l < lcp->levels; l++)
  {
    if(offsets.get(l) != expect) { bad("LCP level offsets do not match the array size and branching factor"); }
    expect += level_size;
    level_size = (level_size + lcp->branching - 1) / lcp->branching;
  }
  if(offsets.get(lcp->levels) != expect || data.size() != expect) { bad("LCP data does not match the level offsets"); }
  u64* offs = (u64*)std::malloc(sizeof(u64) * (lcp->levels + 1));
  uint8_t* bytes = (uint8_t*)std::malloc(data.size() + 1);
  if(offs == nullptr || bytes == nullptr) { std::free(offs); std::free(bytes); bad("out of memory"); }
  for(u64 l = 0; l <= lcp->levels; l++) { offs[l] = offsets.get(l); }
  for(u64 i = 0; i < data.size(); i++) { bytes[i] = (uint8_t)data.get(i); }
  lcp->offsets = offs; lcp->data = bytes;
}

void writeLCP(const gcsa_flat_lcp* lcp, const char* path)
{
  Out out(path);
  out.put<uint32_t>(LCP_TAG); out.put<uint32_t>(LCP_VERSION);
  out.put<u64>(lcp->size); out.put<u64>(lcp->branching); out.put<u64>(0);
  u64 total = lcp->offsets[lcp->levels];
  uint8_t max_value = 0;
  for(u64 i = 0; i < total; i++) { max_value = std::max(max_value, lcp->data[i]); }
  IntVector data = IntVector::zeros(total, (uint8_t)bitLength(max_value));       // bit_compress, src/lcp.cpp:259
  for(u64 i = 0; i < total; i++) { data.set(i, lcp->data[i]); }
  IntVector offsets = IntVector::zeros(lcp->levels + 1, 64);
  for(u64 l = 0; l <= lcp->levels; l++) { offsets.set(l, lcp->offsets[l]); }
  writeIntVector(out, data, true);
  writeIntVector(out, offsets, false);
  out.close();
}

template<class Body> int guarded(const char* where, Body body)
{
  try { body(); return GCSA_B200_OK; }
  catch(const Failure& e) { gcsa_b200_internal_set_error((std::string(where) + ": " + e.what).c_str()); return GCSA_B200_ERR_INVALID; }
  catch(const std::bad_alloc&) { gcsa_b200_internal_set_error((std::string(where) + ": out of memory").c_str()); return GCSA_B200_ERR_NOMEM; }
}

} // namespace

extern "C" {

int gcsa_b200_load_gcsa_file(const char* path, gcsa_b200_built* result)
{
  if(path == nullptr || result == nullptr) { return GCSA_B200_ERR_INVALID; }
  std::memset(result, 0, sizeof(*result));
  int rc = guarded("gcsa_b200_load_gcsa_file", [&]() { loadGCSA(path, result); });
  if(rc != GCSA_B200_OK) { gcsa_b200_built_free(result); }
  return rc;
}

int gcsa_b200_write_gcsa_file(const gcsa_flat_index* index, const char* path)
{
  if(path == nullptr || index == nullptr) { return GCSA_B200_ERR_INVALID; }
  return guarded("gcsa_b200_write_gcsa_file", [&]() { writeGCSA(index, path); });
}

int gcsa_b200_load_lcp_file(const char* path, gcsa_flat_lcp* result)
{
  if(path == nullptr || result == nullptr) { return GCSA_B200_ERR_INVALID; }
  std::memset(result, 0, sizeof(*result));
  return guarded("gcsa_b200_load_lcp_file", [&]() { loadLCP(path, result); });
}

void gcsa_b200_flat_lcp_free(gcsa_flat_lcp* lcp)
{
  if(lcp == nullptr) { return; }
  std::free((void*)lcp->offsets); std::free((void*)lcp->data);
  std::memset(lcp, 0, sizeof(*lcp));
}

int gcsa_b200_write_lcp_file(const gcsa_flat_lcp* lcp, const char* path)
{
  if(path == nullptr || lcp == nullptr || lcp->offsets == nullptr) { return GCSA_B200_ERR_INVALID; }
  return guarded("gcsa_b200_write_lcp_file", [&]() { writeLCP(lcp, path); });
}

} // extern "C"
