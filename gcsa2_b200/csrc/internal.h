/* Shared between the translation units of libgcsa2_b200.so; not part of the C ABI. */
#ifndef GCSA2_B200_INTERNAL_H
#define GCSA2_B200_INTERNAL_H

#ifdef __cplusplus
extern "C" {
#endif

/* Sets the calling thread's gcsa_b200_last_error() message (defined in engine.cu). */
void gcsa_b200_internal_set_error(const char* message);

#ifdef __cplusplus
}
#endif
#endif
