#!/bin/bash
# Round 2, first GPU call: parity first, then the full bench line, the backward-tier A/B and ncu evidence.
tag=${1:-r02a}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader | head -2
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -25 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/${tag}_smoke.log
timeout 1200 python bench.py > $out/${tag}_human_bench.json 2> $out/${tag}_human_bench.log
echo "bench exit $?"; tail -40 $out/${tag}_human_bench.log; cut -c1-3000 $out/${tag}_human_bench.json
for v in "ms2:" "ms0:FMSI_GPU_MULTISTEP=0" "ms3:FMSI_GPU_MULTISTEP=3" "ms2_t12:FMSI_GPU_PREFIX_T=12" "ms2_t12_persist:FMSI_GPU_PREFIX_T=12 FMSI_GPU_L2_PERSIST=table:128" \
         "ms0_plain:FMSI_GPU_MULTISTEP=0 FMSI_GPU_LIB=$PWD/fmsi_b200/variants/libfmsi_gpu_plain.so" "ms2_plain:FMSI_GPU_LIB=$PWD/fmsi_b200/variants/libfmsi_gpu_plain.so"; do
  label=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python profiles/backward_ab.py --label $label >> $out/${tag}_backward_ab.jsonl 2>> $out/${tag}_backward_ab.log
  echo "ab $label exit $?"
done
cat $out/${tag}_backward_ab.jsonl
FMSI_GPU_DICT=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
  --log-file $out/${tag}_human_backward_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --modes none > $out/${tag}_launches_run.log 2>&1
echo "ncu list exit $?"
FMSI_GPU_DICT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:query_kmers_kernel -s 3 -c 1 -f -o $out/${tag}_human_backward \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --modes none > $out/${tag}_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la $out | tail -20
