#!/bin/bash
# Round-2 GPU call 40: batched W_hh transposes (one launch per step instead of ten): tests + headline breakdown
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_gru_tc.py tests/test_gpu_streaming.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --breakdown > $O/bench_tr.json 2> $O/bench_tr.err; grep '^{' $O/bench_tr.json | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('gru', d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['samples_ms_per_step'])"; grep calls $O/bench_tr.err | head -14
