#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
for v in 0 1 0 1; do
if [ $v = 1 ]; then export NSD_EARLY_TAIL=1; else unset NSD_EARLY_TAIL; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --no-cpu-baseline --timeline $O/tl_early$v.json > $O/scale_2_early$v.json 2>/dev/null
python - <<PY
import json
d=[json.loads(l) for l in open('$O/scale_2_early$v.json') if l.startswith('{')][-1]
t=json.load(open('$O/tl_early$v.json'))
print('early_tail=$v ms/step', d['ms_per_step'], [(b['bucket'][13:24], b['ready_ms'], b['done_ms']) for b in t['buckets'][5:]], 'past last wait', t['compute_stream_past_last_wait_ms'])
PY
done
