#!/bin/bash
# Round-2 GPU call 10: restructured streaming push (warp per unit, inputs staged in shared memory): tests, phase trace, latency
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_streaming.py -m gpu -x -q -s > $O/stream_tests.log 2>&1; echo "stream tests rc=$?"; grep -E "fast streaming|streaming vs|passed|failed|Error|error" $O/stream_tests.log | head -20
{ timeout 120 python scratch/stream_trace.py 1; timeout 120 python scratch/stream_trace.py 8; } > $O/stream_trace.log 2>&1; cat $O/stream_trace.log | grep -v Warn
timeout 300 python bench.py --mode stream > $O/stream.jsonl 2> $O/stream.err; echo "stream rc=$?"; cut -c1-420 $O/stream.jsonl; tail -3 $O/stream.err
