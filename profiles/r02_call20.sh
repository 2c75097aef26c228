#!/bin/bash
# Round-2 GPU call 20 (2 GPUs): layer-0 wgrad split per direction (first 100 MB released one GEMM earlier): equivalence + timeline + step time
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -s > $O/dp_nccl.log 2>&1; echo "dp rc=$?"; grep -E "passed|failed|worst|OK" $O/dp_nccl.log | head -6
for i in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --no-cpu-baseline --timeline $O/allreduce_timeline_2gpu.json > $O/scale_2.json 2>/dev/null
python - <<PY
import json
d=[json.loads(l) for l in open('$O/scale_2.json') if l.startswith('{')][-1]
t=json.load(open('$O/allreduce_timeline_2gpu.json'))
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], [(b['bucket'][:22], b['ready_ms'], b['done_ms']) for b in t['buckets'][5:]], 'past last wait', t['compute_stream_past_last_wait_ms'])
PY
done
