#!/bin/bash
# Round 2, 1 GPU: (1) the wide-layout multi-step sectors (64-bit counters, 192 rows) — every test that loads a wide index,
# including the synthetic 4.43 G-row one; (2) A/B of cudaLimitMaxL2FetchGranularity (default vs 32 bytes) on the headline
# kernel at human scale: CUDA-event timing, then DRAM bytes per launch under ncu.
tag=${1:-r02t}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_wide.py tests/test_gpu_build.py::test_wide_layout_at_100mbp tests/test_gpu_parity.py -m gpu -x -q \
  -k "wide or long_k or test_device_matches_oracle or general" --durations=8 > $out/${tag}_pytest_wide.log 2>&1
echo "pytest exit $?"; tail -14 $out/${tag}_pytest_wide.log
for g in 0 32; do
  timeout 300 python profiles/l2fetch_ab.py --gran $g --dict 2 --label fold_gran$g >> $out/${tag}_l2fetch_ab.jsonl 2>> $out/${tag}_l2fetch_ab.log
  echo "ab $g exit $?"
done
for g in 0 32; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
    -k regex:fold_query_kernel -c 6 --csv --log-file $out/${tag}_l2fetch_ncu_gran$g.csv \
    python profiles/l2fetch_ab.py --gran $g --dict 2 --steps 2 --label ncu_gran$g > /dev/null 2>> $out/${tag}_l2fetch_ab.log
  echo "ncu $g exit $?"
done
cat $out/${tag}_l2fetch_ab.jsonl
grep -h "dram__bytes_read.sum\|gpu__time_duration" $out/${tag}_l2fetch_ncu_gran*.csv | tail -24
