/*
  pack.cpp -- host-side 2-bit packing of fixed-length ACGT patterns for the host entry points of
  find().  PCIe is the bottleneck of gcsa_b200_find_fixed_host (32 pattern bytes in, 16 result bytes
  out per query); with enough host cores it is cheaper to pack the patterns to 2 bits per character
  before they cross the bus.  Patterns with any other character are left to the byte path.

  Layout (what find_kernel<.., PACKED = true> reads): ceil(length / 32) 64-bit words per pattern,
  character p at bits [2 (p % 32), 2 (p % 32) + 2) of word p / 32, value = comp - 1.
*/
#include <cstdint>
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
__attribute__((target("avx2")))
inline bool pack32(const uint8_t* p, u64* out)
{
  const __m256i v = _mm256_loadu_si256((const __m256i*)p);
  const __m256i x = _mm256_and_si256(v, _mm256_set1_epi8((char)0xDF));
  const __m256i valid = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(x, _mm256_set1_epi8('A')), _mm256_cmpeq_epi8(x, _mm256_set1_epi8('C'))),
                                        _mm256_or_si256(_mm256_cmpeq_epi8(x, _mm256_set1_epi8('G')), _mm256_cmpeq_epi8(x, _mm256_set1_epi8('T'))));
  // (c >> 1) & 3 = A 0, C 1, T 2, G 3; xor with its own high bit swaps G and T: comp - 1
  const __m256i t = _mm256_and_si256(_mm256_srli_epi16(v, 1), _mm256_set1_epi8(3));
  const __m256i code = _mm256_xor_si256(t, _mm256_and_si256(_mm256_srli_epi16(t, 1), _mm256_set1_epi8(1)));
  const __m256i p16 = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x0401));          // 2 characters -> 4 bits
  const __m256i p32 = _mm256_madd_epi16(p16, _mm256_set1_epi32(0x00100001));          // 4 characters -> 8 bits
  const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                          0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
  const __m256i bytes = _mm256_shuffle_epi8(p32, gather);
  *out = (u64)(uint32_t)_mm256_extract_epi32(bytes, 0) | ((u64)(uint32_t)_mm256_extract_epi32(bytes, 4) << 32);
  return (uint32_t)_mm256_movemask_epi8(valid) == 0xFFFFFFFFu;
}

__attribute__((target("avx2")))
bool packRangeAvx2(const uint8_t* chars, u64 first, u64 last, u64 length, const uint8_t* code, u64* out)
{
  const u64 words = (length + 31) / 32, full = length / 32;
  bool ok = true;
  for(u64 q = first; q < last; q++)
  {
    const uint8_t* p = chars + q * length;
    u64* o = out + q * words;
    for(u64 w = 0; w < full; w++) { ok &= pack32(p + 32 * w, o + w); }
    if(full < words) { ok &= packScalar(p + 32 * full, length - 32 * full, code, o + full); }
  }
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

extern "C" int gcsa_b200_internal_pack_patterns(const uint8_t* chars, uint64_t n, uint64_t length, const uint8_t* code,
                                                 int default_alphabet, uint64_t* out, int threads)
{
  if(threads < 1) { threads = 1; }
  bool simd = false;
#if defined(__x86_64__)
  simd = (default_alphabet != 0) && __builtin_cpu_supports("avx2");
#endif
  const u64 BLOCK = 8192;
  const u64 blocks = (n + BLOCK - 1) / BLOCK;
  int ok = 1;
  #pragma omp parallel for schedule(static) num_threads(threads) reduction(&:ok)
  for(u64 b = 0; b < blocks; b++)
  {
    u64 first = b * BLOCK, last = (first + BLOCK < n ? first + BLOCK : n);
    bool good;
#if defined(__x86_64__)
    if(simd) { good = packRangeAvx2(chars, first, last, length, code, out); }
    else
#endif
    { good = packRangeScalar(chars, first, last, length, code, out); }
    ok &= (good ? 1 : 0);
  }
  return ok;
}
