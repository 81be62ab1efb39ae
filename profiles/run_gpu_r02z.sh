#!/bin/bash
# Round 2, final-code validation on one GPU: what the driver runs at round end (GPU tests, smoke, both bench arms).
tag=${1:-r02z}
out=gpurun_out
mkdir -p $out
timeout 1700 python -m pytest tests -m gpu -q --maxfail=12 > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -6 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/${tag}_smoke.log
rm -rf data/human
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $out/${tag}_human_bench_reference.json 2> $out/${tag}_human_bench_reference.log
echo "reference arm exit $?"; cut -c1-700 $out/${tag}_human_bench_reference.json; grep "Elapsed\|bench\]" $out/${tag}_human_bench_reference.log | tail -5
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${tag}_human_bench.json 2> $out/${tag}_human_bench.log
echo "bench exit $?"; cut -c1-800 $out/${tag}_human_bench.json
FMSI_GPU_TIMING=1 timeout 400 python profiles/backward_ab.py --label fold_final --dict 2 > $out/${tag}_fold_build.json 2> $out/${tag}_fold_build.log
grep "fold build\|lookup ids" $out/${tag}_fold_build.log
ls -la $out | tail -8
