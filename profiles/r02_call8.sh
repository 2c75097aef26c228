#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_streaming.py -m gpu -x -q -s > $O/stream_tests.log 2>&1; echo "stream tests rc=$?"; grep -E "fast streaming|streaming vs|passed|failed|Error|error" $O/stream_tests.log | head -20
timeout 300 python bench.py --mode stream > $O/stream.jsonl 2> $O/stream.err; echo "stream rc=$?"; cat $O/stream.jsonl; tail -3 $O/stream.err
timeout 300 python bench.py --mode infer --steps 20 --no-cpu-baseline > $O/infer.jsonl 2> $O/infer.err; echo "infer rc=$?"; cat $O/infer.jsonl | cut -c1-330
