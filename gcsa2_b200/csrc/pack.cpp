/*
  pack.cpp -- host-side 2-bit packing of fixed-length ACGT patterns for the host entry points of
  find().  PCIe is the bottleneck of gcsa_b200_find_fixed_host (32 pattern bytes in, 16 result bytes
  out per query); with enough host cores it is cheaper to pack the patterns to 2 bits per character
  before they cross the bus.  Patterns with any other character are left to the byte path.

  Layout (what find_kernel<.., PACKED = true> reads): ceil(length / 32) 64-bit words per pattern,
  character p at bits [2 (p % 32), 2 (p % 32) + 2) of word p / 32, value = comp - 1.
*/
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <omp.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "internal.h"

namespace
{

typedef uint64_t u64;

// One pattern through the lookup table; returns false on a character outside the table.
inline bool packScalar(const uint8_t* p, u64 length, const uint8_t* code, u64* out)
{
  uint8_t bad = 0;
  for(u64 w = 0; w * 32 < length; w++)
  {
    u64 word = 0, m = (length - w * 32 < 32 ? length - w * 32 : 32);
    for(u64 i = 0; i < m; i++)
    {
      uint8_t c = code[p[w * 32 + i]];
      bad |= c;
      word |= (u64)(c & 3) << (2 * i);
    }
    out[w] = word;
  }
  return (bad & 0x80) == 0;
}

#if defined(__x86_64__)
// 32 characters of the default alphabet (A, C, G, T in either case) -> one word.
// Codes: bits 1-2 of the byte are A 0, C 1, T 2, G 3; pext gathers them (16 bits per 8 characters) and the xor
// with the own high bit swaps G and T: comp - 1.  Validity: a nibble lookup (low nibble of the upper-cased byte
// -> the high nibble it must have), accumulated by the caller; one movemask per block instead of one per pattern.
__attribute__((target("avx2,bmi2")))
inline void pack32(const uint8_t* p, u64* out, __m256i& all_valid)
{
  const __m256i v = _mm256_loadu_si256((const __m256i*)p);
  const __m256i x = _mm256_and_si256(v, _mm256_set1_epi8((char)0xDF));
  const __m256i lut = _mm256_setr_epi8(-1, 4, -1, 4, 5, -1, -1, 4, -1, -1, -1, -1, -1, -1, -1, -1,
                                       -1, 4, -1, 4, 5, -1, -1, 4, -1, -1, -1, -1, -1, -1, -1, -1);
  const __m256i want = _mm256_shuffle_epi8(lut, _mm256_and_si256(x, _mm256_set1_epi8(0x0F)));
  const __m256i high = _mm256_and_si256(_mm256_srli_epi16(x, 4), _mm256_set1_epi8(0x0F));
  all_valid = _mm256_and_si256(all_valid, _mm256_cmpeq_epi8(want, high));
  // plain 8-byte loads for the scalar side (the laundered pointer keeps the compiler from carving them out of the
  // vector register, which costs more port-5 work than the loads do)
  const uint8_t* p2 = p;
  asm("" : "+r"(p2));
  u64 w[4];
  std::memcpy(w, p2, 32);
  const u64 M = 0x0606060606060606ull;
  u64 t = _pext_u64(w[0], M) | (_pext_u64(w[1], M) << 16) | (_pext_u64(w[2], M) << 32) | (_pext_u64(w[3], M) << 48);
  *out = t ^ ((t >> 1) & 0x5555555555555555ull);
}

__attribute__((target("avx2,bmi2")))
bool packRangeAvx2(const uint8_t* chars, u64 first, u64 last, u64 length, const uint8_t* code, u64* out)
{
  const u64 words = (length + 31) / 32, full = length / 32;
  bool ok = true;
  __m256i all_valid = _mm256_set1_epi8(-1);
  if(length == 32)
  {
    for(u64 q = first; q < last; q++)
    {
      _mm_prefetch((const char*)(chars + 32 * q + 1024), _MM_HINT_NTA);
      u64 w;
      pack32(chars + 32 * q, &w, all_valid);
      _mm_stream_si64((long long*)(out + q), (long long)w);      // the staging buffer is read next by the DMA engine, not by this core
    }
  }
  else
  {
    for(u64 q = first; q < last; q++)
    {
      const uint8_t* p = chars + q * length;
      u64* o = out + q * words;
      for(u64 w = 0; w < full; w++) { pack32(p + 32 * w, o + w, all_valid); }
      if(full < words) { ok &= packScalar(p + 32 * full, length - 32 * full, code, o + full); }
    }
  }
  _mm_sfence();
  return ok && ((uint32_t)_mm256_movemask_epi8(all_valid) == 0xFFFFFFFFu);
}

// Streaming stores send the packed words past the caches to memory; plain stores leave them in the last-level cache,
// where the copy engine's reads can find them (the staging ring of the host pipeline is a few MB).
// GCSA_B200_PACK_STREAM=0 selects plain stores.
static const bool g_stream_stores = []() { const char* e = std::getenv("GCSA_B200_PACK_STREAM"); return !(e != nullptr && std::atoi(e) == 0); }();

// The same with 512-bit registers: 64 characters (two words) per step.  Bits 1 and 2 of every byte come out as two
// 64-bit masks (vptestmb), comp - 1 = (b2, b1 ^ b2), and pdep interleaves the two masks into the 2-bit codes; validity
// is one byte shuffle (low nibble -> the only upper-cased letter with that nibble) and one compare into a mask.
__attribute__((target("avx512f,avx512bw,bmi2")))
inline void pack64(const uint8_t* p, u64* out, __mmask64& all_valid)
{
  const __m512i v = _mm512_loadu_si512((const void*)p);
  const __m512i x = _mm512_and_si512(v, _mm512_set1_epi8((char)0xDF));
  const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8(-1, 0x41, -1, 0x43, 0x54, -1, -1, 0x47, -1, -1, -1, -1, -1, -1, -1, -1));
  const __m512i want = _mm512_shuffle_epi8(lut, _mm512_and_si512(x, _mm512_set1_epi8(0x0F)));
  all_valid &= _mm512_cmpeq_epi8_mask(want, x);
  const u64 b1 = _mm512_test_epi8_mask(v, _mm512_set1_epi8(0x02)), b2 = _mm512_test_epi8_mask(v, _mm512_set1_epi8(0x04));
  const u64 lo = b1 ^ b2, hi = b2;
  const u64 EVEN = 0x5555555555555555ull, ODD = 0xAAAAAAAAAAAAAAAAull;
  const u64 w0 = _pdep_u64(lo & 0xFFFFFFFFull, EVEN) | _pdep_u64(hi & 0xFFFFFFFFull, ODD);
  const u64 w1 = _pdep_u64(lo >> 32, EVEN) | _pdep_u64(hi >> 32, ODD);
  if(g_stream_stores) { _mm_stream_si64((long long*)out, (long long)w0); _mm_stream_si64((long long*)(out + 1), (long long)w1); }
  else { out[0] = w0; out[1] = w1; }
}

// Patterns whose length is a multiple of 32: the batch is one stream of 32-character words.
__attribute__((target("avx512f,avx512bw,bmi2")))
bool packStreamAvx512(const uint8_t* chars, u64 first_word, u64 last_word, u64* out)
{
  __mmask64 all_valid = ~(__mmask64)0;
  u64 w = first_word;
  for(; w + 8 <= last_word; w += 8)
  {
    pack64(chars + 32 * w, out + w, all_valid); pack64(chars + 32 * w + 64, out + w + 2, all_valid);
    pack64(chars + 32 * w + 128, out + w + 4, all_valid); pack64(chars + 32 * w + 192, out + w + 6, all_valid);
  }
  for(; w + 2 <= last_word; w += 2) { pack64(chars + 32 * w, out + w, all_valid); }
  bool ok = (all_valid == ~(__mmask64)0);
  if(w < last_word)
  {
    __m256i valid256 = _mm256_set1_epi8(-1);
    pack32(chars + 32 * w, out + w, valid256);
    ok = ok && ((uint32_t)_mm256_movemask_epi8(valid256) == 0xFFFFFFFFu);
  }
  _mm_sfence();
  return ok;
}
#endif

bool packRangeScalar(const uint8_t* chars, u64 first, u64 last, u64 length, const uint8_t* code, u64* out)
{
  const u64 words = (length + 31) / 32;
  bool ok = true;
  for(u64 q = first; q < last; q++) { ok &= packScalar(chars + q * length, length, code, out + q * words); }
  return ok;
}

} // namespace

// Patterns [first, last) of the batch, on the calling thread: the widest path the CPU and the alphabet allow.
extern "C" int gcsa_b200_internal_pack_range(const uint8_t* chars, uint64_t first, uint64_t last, uint64_t length, const uint8_t* code,
                                             int default_alphabet, uint64_t* out)
{
  if(first >= last) { return 1; }
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
  static const bool avx512 = avx2 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                             std::getenv("GCSA_B200_PACK_NO_AVX512") == nullptr;
  if(default_alphabet != 0 && avx512 && length > 0 && length % 32 == 0)
  {
    const u64 words = length / 32;                                // the patterns are one stream of 32-character words
    return packStreamAvx512(chars, first * words, last * words, out) ? 1 : 0;
  }
  if(default_alphabet != 0 && avx2) { return packRangeAvx2(chars, first, last, length, code, out) ? 1 : 0; }
#endif
  return packRangeScalar(chars, first, last, length, code, out) ? 1 : 0;
}

extern "C" int gcsa_b200_internal_pack_patterns(const uint8_t* chars, uint64_t n, uint64_t length, const uint8_t* code,
                                                 int default_alphabet, uint64_t* out, int threads)
{
  if(threads < 1) { threads = 1; }
  const u64 BLOCK = 8192;
  const u64 blocks = (n + BLOCK - 1) / BLOCK;
  int ok = 1;
  #pragma omp parallel for schedule(static) num_threads(threads) reduction(&:ok)
  for(u64 b = 0; b < blocks; b++)
  {
    u64 first = b * BLOCK, last = (first + BLOCK < n ? first + BLOCK : n);
    ok &= gcsa_b200_internal_pack_range(chars, first, last, length, code, default_alphabet, out);
  }
  return ok;
}
