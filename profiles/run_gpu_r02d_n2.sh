#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus N): the driver's torchrun launch at N ranks with every mode, then the
# inter-process slowdown experiment of VERDICT item 6 on the same box: one rank alone, two independent processes, a gloo
# group, an NCCL group — device-resident headline kernel only, with clocks and power of every GPU sampled meanwhile.
tag=${1:-r02d}
N=${2:-2}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
nvidia-smi --query-gpu=timestamp,index,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv,noheader -lms 200 -f $out/${tag}_smi.csv &
SMI=$!
run_n() {  # label, nproc, extra env...
  label=$1; np=$2; shift 2
  echo "== $label ($(date +%T))" | tee -a $out/${tag}_smi_marks.txt
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np --steps 20 --warmup 5 \
     --modes none --no-cpu-baseline --no-parity > $out/${tag}_${label}.json 2> $out/${tag}_${label}.log
  echo "$label exit $?"; python -c "import json,sys; d=json.load(open('$out/${tag}_${label}.json')); print('$label', 'ms_per_step', round(d['ms_per_step'],4), 'value G', round(d['value']/1e9,2), 'e2e G', round(d['e2e']['value']/1e9,2))"
}
# 1. the driver's launch, every mode
echo "== full N=$N ($(date +%T))" | tee -a $out/${tag}_smi_marks.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29510 bench.py --gpus $N --steps 10 --warmup 3 \
   > $out/${tag}_bench_n${N}.json 2> $out/${tag}_bench_n${N}.log
echo "full bench exit $?"; tail -5 $out/${tag}_bench_n${N}.log; cut -c1-1200 $out/${tag}_bench_n${N}.json
# 2. slowdown experiment
echo "== alone ($(date +%T))" | tee -a $out/${tag}_smi_marks.txt
timeout 600 python bench.py --steps 20 --warmup 5 --modes none --no-cpu-baseline --no-parity > $out/${tag}_alone.json 2> $out/${tag}_alone.log
python -c "import json; d=json.load(open('$out/${tag}_alone.json')); print('alone ms_per_step', round(d['ms_per_step'],4))"
echo "== two independent processes ($(date +%T))" | tee -a $out/${tag}_smi_marks.txt
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 200 --warmup 5 --modes none --no-cpu-baseline --no-parity > $out/${tag}_indep0.json 2> $out/${tag}_indep0.log &
P0=$!
CUDA_VISIBLE_DEVICES=1 timeout 600 python bench.py --steps 200 --warmup 5 --modes none --no-cpu-baseline --no-parity > $out/${tag}_indep1.json 2> $out/${tag}_indep1.log
wait $P0
python -c "import json; [print('independent', i, 'ms_per_step', round(json.load(open('$out/${tag}_indep%d.json' % i))['ms_per_step'],4)) for i in (0,1)]"
run_n gloo 2 FMSI_BENCH_DIST_BACKEND=gloo
run_n nccl 2 FMSI_BENCH_DIST_BACKEND=nccl
echo "== alone again ($(date +%T))" | tee -a $out/${tag}_smi_marks.txt
timeout 600 python bench.py --steps 20 --warmup 5 --modes none --no-cpu-baseline --no-parity > $out/${tag}_alone2.json 2> $out/${tag}_alone2.log
python -c "import json; d=json.load(open('$out/${tag}_alone2.json')); print('alone again ms_per_step', round(d['ms_per_step'],4))"
kill $SMI
ls -la $out | tail -15
