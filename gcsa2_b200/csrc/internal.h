/* Shared between the translation units of libgcsa2_b200.so; not part of the C ABI. */
#ifndef GCSA2_B200_INTERNAL_H
#define GCSA2_B200_INTERNAL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Sets the calling thread's gcsa_b200_last_error() message (defined in engine.cu). */
void gcsa_b200_internal_set_error(const char* message);

/* pack.cpp: 2-bit packs n patterns of `length` bytes into ceil(length / 32) words each (character p of a
   pattern at bits [2 (p % 32), +2) of word p / 32).  code[256] maps a byte to comp - 1 for the four fast
   characters and to 0xFF otherwise; default_alphabet != 0 says the table is exactly ACGT/acgt (enables the
   AVX2 path).  Returns 1 if every byte was a fast character, else 0 (the output is then unusable). */
int gcsa_b200_internal_pack_patterns(const uint8_t* chars, uint64_t n, uint64_t length, const uint8_t* code,
                                     int default_alphabet, uint64_t* out, int threads);

/* The same for patterns [first, last) of the batch on the calling thread (no OpenMP inside): out is the base of the
   whole batch's output. */
int gcsa_b200_internal_pack_range(const uint8_t* chars, uint64_t first, uint64_t last, uint64_t length, const uint8_t* code,
                                  int default_alphabet, uint64_t* out);

/* Measurement hooks of engine.cu (bench.py and the tests read them; not part of the C ABI): the number of batches of
   this process that took the two-kernel k-mer form of find(), and how many chunks of the last host-buffer find() went
   over the link packed / in all. */
unsigned long long gcsa_b200_internal_fast_launches(void);
void gcsa_b200_internal_pack_share(unsigned long long* packed, unsigned long long* total);

#ifdef __cplusplus
}
#endif
#endif
