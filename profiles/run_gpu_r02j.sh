#!/bin/bash
# Round 2, 1 GPU: ncu --set full of stream_kernel at human scale (150-bp reads, -O -S, backward tier), the bench line with the
# CLI like-for-like on reads and lookup.
tag=${1:-r02j}
out=gpurun_out
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 4 -c 1 -f -o $out/${tag}_human_stream \
  python profiles/reads_e2e_trace.py --genome 3100000000 --dict 0 > $out/${tag}_ncu_stream.log 2>&1
echo "ncu stream exit $?"
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${tag}_human_bench.json 2> $out/${tag}_human_bench.log
echo "bench exit $?"; tail -5 $out/${tag}_human_bench.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02j_human_bench.json"))
for k, v in d["cli"].items():
    print(k, {kk: vv for kk, vv in v.items() if kk not in ("note", "ours_stages")})
print("wall", d["details"]["bench_wall_s"])
PY
ls -la $out | tail -5
