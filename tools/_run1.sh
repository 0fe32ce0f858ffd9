set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2a
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log )
tail -5 gpurun_out/r2a/pytest.log
timeout 600 python bench.py > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a/bench_ref.json 2>&1
timeout 300 python tools/latency_probe.py > gpurun_out/r2a/latency.jsonl 2>&1
# ncu full captures (each after a plain run has exited 0)
for spec in "bn_254 4_3 20 2 compress" "bn_254 4_3 18 2 hash37" "bls12_377 2_1 20 2 compress" "bn_254 2_1 20 2 compress"; do
  set -- $spec
  tag=$1_$2_$5
  timeout 200 python tools/profile_target.py $1 $2 $3 $4 $5 > gpurun_out/r2a/plain_$tag.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 1 --launch-count 1 -k regex:anemoi_kernel -f -o gpurun_out/r2a/ncu_$tag python tools/profile_target.py $1 $2 $3 $4 $5 > gpurun_out/r2a/ncu_$tag.log 2>&1
done
ls -la gpurun_out/r2a
