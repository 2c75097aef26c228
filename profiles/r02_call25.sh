#!/bin/bash
# Round-2 GPU call 25: launch list of the Conformer step under ncu (per-kernel durations), compute-sanitizer memcheck on the new kernels' tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $O/conformer_launches.csv python bench.py --mode conformer --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_conformer_launches.log 2>&1; echo "launch list rc=$?"; wc -l $O/conformer_launches.csv
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_conformer_kernels.py -m gpu -q -x > $O/sanitizer_conformer.log 2>&1; echo "memcheck rc=$?"; tail -5 $O/sanitizer_conformer.log | cut -c1-200
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_streaming.py -m gpu -q -x -k "fast_form or push_decode" > $O/sanitizer_stream.log 2>&1; echo "memcheck stream rc=$?"; tail -4 $O/sanitizer_stream.log | cut -c1-200
