// Anemoi kernels for bls12_381 (src/bls12_381/ in the reference): Anemoi-2-1 and Anemoi-4-3.
#define ANEMOI_FIELD_TABLES_bls12_381 1
#include "fp.cuh"
#include "generated/fields.cuh"
#include "field_tu.cuh"
ANEMOI_DEFINE_LAUNCHER(bls12_381)
