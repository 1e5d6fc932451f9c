/* Stand-in for <sodium.h> (absent): include/zkp/hash.hpp:223-272 defines a BLAKE2b hasher next to the SHA-256
 * one the prover uses; only its declarations have to parse.  Nothing here is ever called or linked.
 * Test infrastructure only (see tests/stubs/gmp.h). */
#ifndef LGR_TEST_SODIUM_STANDIN_H
#define LGR_TEST_SODIUM_STANDIN_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
#define crypto_generichash_BYTES 32U
typedef struct { unsigned char opaque[384]; } crypto_generichash_state;
int sodium_init(void);
int crypto_generichash_init(crypto_generichash_state *, const unsigned char *, size_t, size_t);
int crypto_generichash_update(crypto_generichash_state *, const unsigned char *, unsigned long long);
int crypto_generichash_final(crypto_generichash_state *, unsigned char *, size_t);
#ifdef __cplusplus
}
#endif
#endif
