#!/bin/bash
# Round 2, final code on 2 GPUs: the driver's torchrun launch (reference arm first, then ours), every mode.
tag=${1:-r02af}
out=gpurun_out
mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 \
  > $out/${tag}_bench_n2_reference.json 2> $out/${tag}_bench_n2_reference.log
echo "reference arm exit $?"; cut -c1-300 $out/${tag}_bench_n2_reference.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 \
  > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.log
echo "bench exit $?"; tail -3 $out/${tag}_bench_n2.log
python - <<PY
import json
d = json.load(open("$out/${tag}_bench_n2.json"))
print("N", d["n_gpus"], "value G", round(d["value"] / 1e9, 2), "e2e G", round(d["e2e"]["value"] / 1e9, 2), "ms", round(d["ms_per_step"], 4))
for k, v in d.get("modes", {}).items():
    print(k, round(v["value"] / 1e9, 2), round(v["e2e"]["value"] / 1e9, 2), [v[x] for x in v if "parity" in x])
for k in ("pool", "sharded"):
    print(k, str(d.get(k))[:300])
PY
