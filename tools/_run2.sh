set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2b; mkdir -p $O
for v in v0 v1 v2 v3; do
  export ANEMOI_B200_LIB=$PWD/variants/$v.so
  ( timeout 600 python -m pytest tests/test_gpu_kat.py tests/test_gpu_golden_random.py -x -q > $O/kat_$v.log 2>&1; echo "rc=$?" >> $O/kat_$v.log )
  tail -2 $O/kat_$v.log
  timeout 600 python tools/quick_bench.py --log2n 20 --reps 3 > $O/qb_$v.jsonl 2>&1
done
for v in v0 v3; do
  export ANEMOI_B200_LIB=$PWD/variants/$v.so
  timeout 300 python tools/latency_probe.py pallas/anemoi_4_3 bls12_377/anemoi_2_1 bn_254/anemoi_2_1 > $O/lat_$v.jsonl 2>&1
done
for spec in "v1 bls12_381 2_1" "v2 bls12_381 2_1" "v1 bls12_377 2_1" "v1 bn_254 4_3"; do
  set -- $spec
  export ANEMOI_B200_LIB=$PWD/variants/$1.so
  timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 1 --launch-count 1 -k regex:anemoi_kernel -f -o $O/ncu_$1_$2_$3 python tools/profile_target.py $2 $3 20 2 compress > $O/ncu_$1_$2_$3.log 2>&1
done
ls -la $O
