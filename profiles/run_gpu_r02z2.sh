#!/bin/bash
# Round 2, final code: both bench arms as the driver launches them (reference first, fresh data/).
tag=${1:-r02z}
out=gpurun_out
mkdir -p $out
rm -rf data/human
SECONDS=0
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $out/${tag}_human_bench_reference.json 2> $out/${tag}_human_bench_reference.log
echo "reference arm exit $? in ${SECONDS}s"; cut -c1-900 $out/${tag}_human_bench_reference.json; tail -3 $out/${tag}_human_bench_reference.log
SECONDS=0
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${tag}_human_bench.json 2> $out/${tag}_human_bench.log
echo "bench exit $? in ${SECONDS}s"; cut -c1-600 $out/${tag}_human_bench.json
ls -la data/human | head; du -sh data
