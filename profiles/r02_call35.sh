#!/bin/bash
# Round-2 GPU call 35: register-resident CTC recursions: CTC tests (oracle, fixtures, edge cases), model tests, headline breakdown
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_conformer.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --breakdown > $O/bench_ctc.json 2> $O/bench_ctc.err; grep '^{' $O/bench_ctc.json | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('gru', d['ms_per_step'], d['value'], d['e2e']['value'])"; grep calls $O/bench_ctc.err | head -12
