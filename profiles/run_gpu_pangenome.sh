#!/bin/bash
# Pangenome-shaped indexes (BASELINE configs[4]) on one B200: the scaled 60-copy concatenation and the masked-superstring
# form at the config's named size (base genome + SNP windows, ~1.2 G distinct k-mers). usage: bash profiles/run_gpu_pangenome.sh <tag>
tag=${1:-r01x}
out=gpurun_out
mkdir -p $out
timeout 600 python profiles/modes_bench.py --copies 60 --k 31 --tiers fold,backward > $out/${tag}_modes_pangenome60x_k31.json 2> $out/${tag}_pg60_k31.log
echo "60x k31 exit $?"; tail -n 12 $out/${tag}_modes_pangenome60x_k31.json | head -11
timeout 900 python profiles/modes_bench.py --variants 39000000 --k 31 --tiers fold,backward > $out/${tag}_modes_pangenome_variants_k31.json 2> $out/${tag}_pgv_k31.log
echo "variants k31 exit $?"; tail -n 12 $out/${tag}_modes_pangenome_variants_k31.json | head -11; tail -3 $out/${tag}_pgv_k31.log
timeout 900 python profiles/modes_bench.py --variants 52000000 --k 23 --tiers fold > $out/${tag}_modes_pangenome_variants_k23.json 2> $out/${tag}_pgv_k23.log
echo "variants k23 exit $?"; tail -n 12 $out/${tag}_modes_pangenome_variants_k23.json | head -11; tail -3 $out/${tag}_pgv_k23.log
