// Anemoi kernels for jubjub (src/jubjub/ in the reference): Anemoi-2-1 and Anemoi-4-3.
#define ANEMOI_FIELD_TABLES_jubjub 1
#include "fp.cuh"
#include "generated/fields.cuh"
#include "field_tu.cuh"
ANEMOI_DEFINE_LAUNCHER(jubjub)
