/* Minimal C use of libanemoi_b200.so: Jive 2->1 compression of a few BLS12-381 digest pairs and a Merkle root.
 * Build:  gcc -std=c99 -Iinclude examples/compress_example.c -Lanemoi_rust_b200 -lanemoi_b200 \
 *             -Wl,-rpath,$PWD/anemoi_rust_b200 -o compress_example
 * Inputs are Montgomery-form limbs exactly as arkworks stores them (a * 2^384 mod p); 0 is 0 and R mod p is 1. */
#include <stdio.h>
#include <string.h>

#include "anemoi_b200.h"

int main(void) {
    /* R mod p for BLS12-381 Fq = Montgomery form of 1 */
    const uint64_t one[6] = {0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                             0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL};
    uint64_t pairs[4][2][6], out[4][6], root[6];
    uint8_t bytes[48];
    int i, rc;
    memset(pairs, 0, sizeof(pairs));
    /* the four inputs of the reference's test_anemoi_jive: [0,0], [1,1], [0,1], [1,0] */
    memcpy(pairs[1][0], one, sizeof(one));
    memcpy(pairs[1][1], one, sizeof(one));
    memcpy(pairs[2][1], one, sizeof(one));
    memcpy(pairs[3][0], one, sizeof(one));
    if (anemoi_b200_device_count() == 0) {
        printf("no CUDA device: %s\n", anemoi_b200_strerror(ANEMOI_B200_ERR_NO_DEVICE));
        return 0;
    }
    rc = anemoi_b200_compress(ANEMOI_FIELD_BLS12_381, ANEMOI_INST_2_1, 2, &pairs[0][0][0], &out[0][0], 4, 0);
    if (rc) { printf("compress: %s (%s)\n", anemoi_b200_strerror(rc), anemoi_b200_last_cuda_error()); return 1; }
    for (i = 0; i < 4; i++) {
        int b;
        anemoi_b200_digest_to_bytes(ANEMOI_FIELD_BLS12_381, out[i], bytes, 1, 0);
        printf("compress[%d] = 0x", i);
        for (b = 47; b >= 0; b--) printf("%02x", bytes[b]);
        printf("\n");
    }
    /* the 4 digests as leaves of an arity-2 Jive Merkle tree */
    rc = anemoi_b200_merkle_root(ANEMOI_FIELD_BLS12_381, ANEMOI_INST_2_1, 2, &out[0][0], 4, root, 1);
    if (rc) { printf("merkle_root: %s\n", anemoi_b200_strerror(rc)); return 1; }
    printf("root limb0 = 0x%016llx\n", (unsigned long long)root[0]);
    /* the reference panics on k = 4 for Anemoi-2-1; the ABI reports it */
    rc = anemoi_b200_compress(ANEMOI_FIELD_BLS12_381, ANEMOI_INST_2_1, 4, &pairs[0][0][0], &out[0][0], 4, 0);
    printf("k = 4 on Anemoi-2-1 -> %d (%s)\n", rc, anemoi_b200_strerror(rc));
    return 0;
}
