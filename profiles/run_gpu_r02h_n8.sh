#!/bin/bash
# Round 2, 8 GPUs: the driver's torchrun launch at N = 8 with every mode (what SCALE_r02 will run).
tag=${1:-r02h}
N=${2:-8}
out=gpurun_out
mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29510 bench.py --gpus $N --steps 20 --warmup 5 \
   > $out/${tag}_bench_n${N}.json 2> $out/${tag}_bench_n${N}.log
echo "full bench exit $?"; tail -6 $out/${tag}_bench_n${N}.log; cut -c1-600 $out/${tag}_bench_n${N}.json
nproc; free -g | head -2
