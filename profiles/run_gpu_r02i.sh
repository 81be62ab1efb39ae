#!/bin/bash
# Round 2, 1 GPU: memcheck over the new kernels, the L2-persisting A/B with ncu counters, fold build memory peak.
tag=${1:-r02i}
out=gpurun_out
mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/${tag}_smoke.log
FMSI_GPU_TIMING=1 timeout 400 python profiles/backward_ab.py --label fold_mem --dict 2 > $out/${tag}_fold_build.json 2> $out/${tag}_fold_build.log
grep "fold build\|lookup ids" $out/${tag}_fold_build.log
for v in "t12:FMSI_GPU_PREFIX_T=12" "t12_persist:FMSI_GPU_PREFIX_T=12 FMSI_GPU_L2_PERSIST=table:128"; do
  label=${v%%:*}; envs=${v#*:}
  env $envs FMSI_GPU_DICT=0 timeout 600 ncu --metrics gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none \
    -k regex:query_kmers_kernel -s 3 -c 1 --csv --log-file $out/${tag}_persist_${label}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --modes none > $out/${tag}_persist_${label}.log 2>&1
  echo "ncu $label exit $?"; grep -o '"query_kmers_kernel.*' $out/${tag}_persist_${label}.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "reads_api or bit_packed or fold_lookup or streaming_and_query_goldens or (device_matches_oracle and (syn_k9_min or quirks_k3 or syn_k5_min)) or (chunks_streaming and syn_k9_min)" \
  > $out/${tag}_compute_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; tail -6 $out/${tag}_compute_sanitizer_memcheck.log
ls -la $out | tail -8
