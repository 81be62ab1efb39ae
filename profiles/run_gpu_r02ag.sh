#!/bin/bash
# Round 2, final code: the ncu launch list of the bench command (gpu__time_duration per launch; cold-cache, serialised: the kernels'
# SHARES of a step are what to read) and one `--set full` capture of the headline kernel for roofline.traffic.
tag=${1:-r02ag}
out=gpurun_out
mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'query_k|fold_query|stream_kernel|extract_|pack_bases|pack_presence|expand_reads|read_counts' -c 300 --csv \
  --log-file $out/${tag}_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --modes human --cli-records 0 > $out/${tag}_ncu_list.log 2>&1
echo "ncu list exit $?"; grep -c fold_query $out/${tag}_bench_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fold_query_kernel -s 3 -c 1 -f -o $out/${tag}_human_fold \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --modes none --cli-records 0 > $out/${tag}_ncu_full.log 2>&1
echo "ncu full exit $?"
