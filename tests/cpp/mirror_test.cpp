// Exercises include/anemoi_b200.hpp (the C++ host mirror of the reference interface) against two of the
// reference's own known answers. Built and run by tests/test_cpp_mirror.py; needs a GPU to run.
//   zero digest / Jive KATs: src/bls12_381/anemoi_2_1/hasher.rs test_anemoi_jive (inputs [0,0] and [1,1]).
#include <cstdio>
#include <cstring>

#include "anemoi_b200.hpp"

using namespace anemoi_b200;

int main() {
    using H = AnemoiBls12_381_2_1;
    // Montgomery forms: 0 -> 0; 1 -> R mod p
    H::F zero{}, one{};
    const uint64_t one_limbs[6] = {0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                                   0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL};
    for (int i = 0; i < 6; i++) one.limbs[i] = one_limbs[i];
    auto a = H::compress({zero, zero});
    auto b = H::compress_k({zero, zero}, 2);
    auto m = H::merge({H::Digest{{zero}}, H::Digest{{zero}}});
    if (a.size() != 1 || a != b || m.e[0] != a[0]) { printf("FAIL: compress/compress_k/merge disagree\n"); return 1; }
    // canonical bytes of the digest, printed for the Python side to compare with the reference's decimal KAT
    auto bytes = H::Digest{{a[0]}}.to_bytes();
    printf("jive00=");
    for (int i = 47; i >= 0; i--) printf("%02x", bytes[i]);
    printf("\n");
    auto c = H::compress({one, one});
    bytes = H::Digest{{c[0]}}.to_bytes();
    printf("jive11=");
    for (int i = 47; i >= 0; i--) printf("%02x", bytes[i]);
    printf("\n");
    bool threw = false;
    try { H::compress_k({zero, zero}, 4); } catch (const std::invalid_argument&) { threw = true; }
    if (!threw) { printf("FAIL: compress_k(.,4) on 2-1 must fail like the reference's assert!\n"); return 1; }
    threw = false;
    try { H::compress({zero, zero, zero}); } catch (const std::invalid_argument&) { threw = true; }
    if (!threw) { printf("FAIL: compress of 3 elements must fail\n"); return 1; }
    using H4 = AnemoiPallas_4_3;
    std::vector<H4::F> leaves(16);
    for (int i = 0; i < 16; i++) leaves[i].limbs[0] = i + 1;
    auto root = H4::merkle_root(leaves);
    auto l1 = H4::compress_k_batch(leaves, 4);
    auto root2 = H4::compress_k(l1, 4);
    if (root != root2[0]) { printf("FAIL: merkle_root != iterated compress_k\n"); return 1; }
    std::vector<H4::F> paths;
    std::vector<uint64_t> idx = {0, 7, 15};
    auto root3 = H4::merkle_open(leaves, idx, paths);
    if (root3 != root || paths.size() != 3 * 2 * 3) { printf("FAIL: merkle_open\n"); return 1; }
    std::vector<H4::F> vals = {leaves[0], leaves[7], leaves[15]};
    auto roots = H4::merkle_verify(vals, idx, paths, 2);
    for (auto& r : roots) if (r != root) { printf("FAIL: merkle_verify\n"); return 1; }
    printf("OK\n");
    return 0;
}
