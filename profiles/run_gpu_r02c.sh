#!/bin/bash
# Round 2, third GPU call (1 GPU): compact strand-folded dictionary (overflow rows only, ids on demand, sort passes),
# >= 256 k-mers per warp grab; parity, full bench line, e2e reads timeline, ncu of the fold kernel, launch lists.
tag=${1:-r02c}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -15 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/${tag}_smoke.log
FMSI_GPU_TIMING=1 timeout 400 python profiles/backward_ab.py --label fold --dict 2 > $out/${tag}_fold_build.json 2> $out/${tag}_fold_build.log
echo "fold build exit $?"; grep "fold build\|lookup ids" $out/${tag}_fold_build.log; cat $out/${tag}_fold_build.json
nvidia-smi --query-gpu=memory.used,memory.total --format=csv,noheader
timeout 1200 python bench.py > $out/${tag}_human_bench.json 2> $out/${tag}_human_bench.log
echo "bench exit $?"; tail -12 $out/${tag}_human_bench.log; cut -c1-1500 $out/${tag}_human_bench.json
FMSI_GPU_TRACE=1 timeout 300 python profiles/reads_e2e_trace.py > $out/${tag}_reads_trace.json 2> $out/${tag}_reads_trace.log
cat $out/${tag}_reads_trace.json; grep "fmsi trace" $out/${tag}_reads_trace.log | tail -2
FMSI_GPU_TRACE=1 timeout 300 python profiles/reads_e2e_trace.py --dict 0 > $out/${tag}_reads_trace_backward.json 2> $out/${tag}_reads_trace_backward.log
cat $out/${tag}_reads_trace_backward.json; grep "fmsi trace" $out/${tag}_reads_trace_backward.log | tail -2
for w in "fold:" "backward:FMSI_GPU_DICT=0"; do
  label=${w%%:*}; envs=${w#*:}
  env $envs timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'query_k|fold_query|stream_kernel|extract_|pack_bases|pack_presence' -c 260 --csv \
    --log-file $out/${tag}_human_${label}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --modes none > $out/${tag}_launches_${label}.log 2>&1
  echo "ncu list $label exit $?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fold_query_kernel -s 3 -c 1 -f -o $out/${tag}_human_fold \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --modes none > $out/${tag}_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la $out | tail -12
