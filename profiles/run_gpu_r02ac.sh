#!/bin/bash
# Round 2, 1 GPU: compute-sanitizer memcheck and initcheck over this session's new kernels — the minimizer-bucketed dictionary
# (build + tile kernel, every parked / slow path the small indexes reach) and the wide multi-step sectors.
tag=${1:-r02ac}
out=gpurun_out
mkdir -p $out
sel='(chunks_streaming and (loc or wide) and (syn_k9_min or quirks_k3 or syn_k5_min or syn_k32 or data_k13)) or (streaming_and_query_goldens and (loc or wide)) or (device_matches_oracle and wide and (syn_k9_min or quirks_k3)) or (long_k and wide and syn_k47_max)'
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$sel" > $out/${tag}_compute_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; tail -5 $out/${tag}_compute_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool initcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$sel" > $out/${tag}_initcheck.log 2>&1
echo "initcheck exit $?"; tail -5 $out/${tag}_initcheck.log
