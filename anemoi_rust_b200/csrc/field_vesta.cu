// Anemoi kernels for vesta (src/vesta/ in the reference): Anemoi-2-1 and Anemoi-4-3.
#define ANEMOI_FIELD_TABLES_vesta 1
#include "fp.cuh"
#include "generated/fields.cuh"
#include "field_tu.cuh"
ANEMOI_DEFINE_LAUNCHER(vesta)
