#!/bin/bash
# Developer tool: build alternative variants of libanemoi_b200.so (different ladder programs) into variants/<name>.so so
# that one gpurun call can A/B them (ANEMOI_B200_LIB=variants/<name>.so python tools/quick_bench.py ...).
# usage: tools/ab_variants.sh name "ANEMOI_CHAIN_SOURCE value" [chains.json]
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; json=${3:-tools/chains.json}
mkdir -p variants
cp anemoi_rust_b200/csrc/generated/fields.cuh /tmp/fields.cuh.keep
cp oracle/params_gen.h /tmp/params_gen.h.keep
ANEMOI_CHAIN_SOURCE="$src" ANEMOI_CHAINS_JSON="$json" python tools/gen_params.py > variants/$name.gen.log
make -B -j8 anemoi_rust_b200/libanemoi_b200.so EXTRA_NVFLAGS="$EXTRA_NVFLAGS" > /dev/null
cp anemoi_rust_b200/libanemoi_b200.so variants/$name.so
grep -h "registers\|spill" build/field_*.ptxas.log | grep -v "^$" | sort | uniq -c | sort -rn | head -5 > variants/$name.ptxas.txt || true
cp /tmp/fields.cuh.keep anemoi_rust_b200/csrc/generated/fields.cuh
cp /tmp/params_gen.h.keep oracle/params_gen.h
