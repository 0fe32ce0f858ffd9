// Internal: kernel mode ids and the argument block shared by the C ABI (api.cu) and the kernels.
#pragma once
#include <cstdint>

namespace anemoi {

enum Mode : int {
    MODE_PERMUTE = 0,      // in place: state -> permutation(state)
    MODE_SBOX = 1,         // in place: state -> sbox_layer(state)       (diagnostic, pins test_sbox)
    MODE_COMPRESS = 2,     // Jive k = 2: W felts -> W/2 felts
    MODE_COMPRESS4 = 3,    // Jive k = 4 (4-3 only): 4 felts -> 1
    MODE_HASH = 4,         // sponge hash_field, fixed length per message
    MODE_HASH_RAGGED = 5,  // sponge hash_field, offsets[]
    MODE_HASH_BYTES = 6,   // sponge hash (bytes), fixed length per message
    MODE_MERGE43 = 7,      // 4-3 sponge merge: [d0, d0, 0, 0] (sic, reference ignores d1)
    MODE_TO_BYTES = 8,     // digest.to_bytes: de-Montgomery, one felt per unit (no permutation)
    MODE_HASH_BYTES_RAGGED = 9,  // sponge hash (bytes), byte offsets[]
    // diagnostic single-layer modes (anemoi_layer_kernel)
    MODE_LAYER_ARK = 10,    // ark_layer(state, round = len)
    MODE_LAYER_MDS = 11,    // mds_layer(state)
    MODE_LAYER_ROUND = 12,  // round(state, round = len) = ark -> mds -> sbox
};

struct KernelArgs {
    const uint32_t* in;
    uint32_t* out;
    const unsigned long long* offsets;  // MODE_HASH_RAGGED: n + 1 element offsets; MODE_HASH_BYTES_RAGGED: byte offsets
    unsigned long long n;               // states / messages / felts
    unsigned long long len;             // felts (MODE_HASH) or bytes (MODE_HASH_BYTES) per message
    int mode;
    int vec16;  // in/out are 16-byte aligned -> 128-bit accesses
};

}  // namespace anemoi
