#!/bin/bash
# One gpurun call: what the driver runs at round end (GPU tests, smoke, both bench arms) + ncu launch list, one full
# capture of the headline kernel and a compute-sanitizer pass over the parity tests.
# usage (from the repo root, on the GPU box): bash profiles/run_gpu_round.sh <tag> [skip-tests]
tag=${1:-r01x}
out=gpurun_out
mkdir -p $out
if [ "$2" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
  tail -5 $out/${tag}_pytest_gpu.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/${tag}_smoke.log
timeout 900 python bench.py --impl reference > $out/${tag}_human_bench_reference.json 2> $out/${tag}_human_bench_reference.log
echo "reference arm exit $?"; cut -c1-400 $out/${tag}_human_bench_reference.json
timeout 900 python bench.py > $out/${tag}_human_bench.json 2> $out/${tag}_human_bench.log
echo "bench exit $?"; cat $out/${tag}_human_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'query|memset' -c 400 --csv \
  --log-file $out/${tag}_human_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_launches_run.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fold_query -s 3 -c 1 -f -o $out/${tag}_human_fold \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1
echo "ncu full exit $?"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not pool and not midsize and not repetitive" > $out/${tag}_compute_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; tail -3 $out/${tag}_compute_sanitizer_memcheck.log
ls -la $out | head -40
