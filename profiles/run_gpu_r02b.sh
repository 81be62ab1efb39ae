#!/bin/bash
# Round 2, second GPU call (1 GPU): host-side timeline of the e2e reads path, CLI stage times, launch lists of the timed kernels.
tag=${1:-r02b}
out=gpurun_out
mkdir -p $out
FMSI_GPU_TRACE=1 timeout 300 python profiles/reads_e2e_trace.py > $out/${tag}_reads_trace.json 2> $out/${tag}_reads_trace.log
echo "trace exit $?"; cat $out/${tag}_reads_trace.json; grep "fmsi trace" $out/${tag}_reads_trace.log | tail -4
FMSI_GPU_TRACE=1 timeout 300 python profiles/reads_e2e_trace.py --dict 0 > $out/${tag}_reads_trace_backward.json 2> $out/${tag}_reads_trace_backward.log
cat $out/${tag}_reads_trace_backward.json; grep "fmsi trace" $out/${tag}_reads_trace_backward.log | tail -3
# CLI stages on 10 M single-31-mer records against the E. coli-sized index
python - <<'PY' > $out/${tag}_cli_stages.log 2>&1
import os, subprocess, sys, time
sys.path.insert(0, ".")
import bench
from fmsi_b200 import synth
wl = bench.prepare_file_index("ecoli")
q = bench.host_kmer_sample(wl, 10_000_000, 5)
open("/tmp/q10m.fa", "wb").write(synth.packed_to_fasta(q, 31))
for env in ({}, {"FMSI_GPU_THREADS": "14"}, {"FMSI_GPU_STRANDS": "lazy"}, {"FMSI_GPU_BATCH_BASES": str(4 << 20)}, {"FMSI_GPU_BATCH_BASES": str(64 << 20)}):
    for rep in range(2):
        t0 = time.time()
        r = subprocess.run([bench.OUR_FMSI, "query", "-O", "-q", "/tmp/q10m.fa", wl["prefix"]], stdout=open("/tmp/o.txt", "wb"), stderr=subprocess.PIPE,
                           env=dict(os.environ, FMSI_GPU_TIMING="1", **env))
        print(env, "wall %.3f" % (time.time() - t0), " | ".join(l.replace("[fmsi timing] ", "") for l in r.stderr.decode().splitlines()))
t0 = time.time()
subprocess.run([bench.REF_FMSI, "query", "-O", "-q", "/tmp/q10m.fa", wl["prefix"]], stdout=open("/tmp/r.txt", "wb"))
print("reference, one process: %.2f s" % (time.time() - t0), "identical:", open("/tmp/o.txt", "rb").read() == open("/tmp/r.txt", "rb").read())
print("nproc", os.cpu_count())
PY
cat $out/${tag}_cli_stages.log
for w in "fold:" "backward:FMSI_GPU_DICT=0"; do
  label=${w%%:*}; envs=${w#*:}
  env $envs timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'query_kernel|stream_kernel|extract|pack_|memset' -c 200 --csv \
    --log-file $out/${tag}_human_${label}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --modes none > $out/${tag}_launches_${label}.log 2>&1
  echo "ncu list $label exit $?"
done
ls -la $out | tail
