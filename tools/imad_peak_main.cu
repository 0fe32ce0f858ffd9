// Standalone driver for the IMAD issue-rate microbenchmark (anemoi_rust_b200/csrc/imad_peak.cu).
#include <cstdio>
#include "anemoi_b200.h"
int main() {
    const char* names[] = {"mad.lo.u32 (IMAD)", "mad.hi.u32 (IMAD.HI)", "mad.wide.u32 (IMAD.WIDE)",
                           "carry chain (IMAD.WIDE.X)", "IMAD.WIDE + one add.u32 each (MAC32 counted)", "fma.rn.f64 (DFMA)",
                           "IMAD.WIDE + one DFMA each (MAC32 counted)", "IMAD.WIDE + one IMAD lo each (MAC32 counted)",
                           "heterogeneous warps: half IMAD.WIDE-only, half DFMA-only (both counted)"};
    for (int rep = 0; rep < 2; rep++)
        for (int v = 0; v < 9; v++) {
            double ops = 0, mhz = 0;
            int rc = anemoi_b200_imad_peak(v, &ops, &mhz);
            printf("{\"variant\": %d, \"name\": \"%s\", \"rc\": %d, \"Tops_per_s\": %.4f, \"sm_mhz\": %.1f, \"ops_per_clk_per_sm\": %.2f}\n",
                   v, names[v], rc, ops * 1e-12, mhz, ops / (mhz * 1e6) / 148.0);
        }
    return 0;
}
