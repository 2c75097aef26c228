#!/bin/bash
# Round-2 GPU call 42: optimizer under the recurrence: bit-identity test, model tests, headline line with the clean value region
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
for v in 1 0; do
  NSD_STEP_IN_BACKWARD=$v timeout 300 python bench.py --no-cpu-baseline --breakdown > $O/bench_sib$v.json 2> $O/bench_sib$v.err; grep '^{' $O/bench_sib$v.json | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step_in_backward=$v', d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['samples_ms_per_step'], 'gemm frac', d['roofline']['frac'], d['roofline']['share_of_step'], 'launches', d['gpu_launches'])"; grep "calls" $O/bench_sib$v.err | head -4
done
