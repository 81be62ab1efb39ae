#!/bin/bash
# Round 2, 1 GPU: parity after the pointer-doubling fold build + the synthetic 4.4 G-row (wide) index test; disk bandwidth
# of the box (for the cached-index question); fold build stage times.
tag=${1:-r02e}
out=gpurun_out
mkdir -p $out
timeout 1700 python -m pytest tests -m gpu -q --maxfail=12 --durations=8 > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -22 $out/${tag}_pytest_gpu.log
FMSI_GPU_TIMING=1 timeout 400 python profiles/backward_ab.py --label fold_doubling --dict 2 > $out/${tag}_fold_build.json 2> $out/${tag}_fold_build.log
grep "fold build\|lookup ids" $out/${tag}_fold_build.log; cut -c1-400 $out/${tag}_fold_build.json
FMSI_GPU_TIMING=1 FMSI_GPU_FOLD_WALK=1 timeout 400 python profiles/backward_ab.py --label fold_walk --dict 2 > $out/${tag}_fold_build_walk.json 2> $out/${tag}_fold_build_walk.log
grep "fold build" $out/${tag}_fold_build_walk.log | head -4
# what a cached converted index would cost to read: sequential read of the box's scratch disk, page cache bypassed
mkdir -p data; ( dd if=/dev/zero of=data/ddtest.bin bs=16M count=256 oflag=direct 2>&1 | tail -1; dd if=data/ddtest.bin of=/dev/null bs=16M iflag=direct 2>&1 | tail -1; rm -f data/ddtest.bin ) | tee $out/${tag}_disk_bw.log
df -h . | tail -1 | tee -a $out/${tag}_disk_bw.log
ls -la $out | tail -8
