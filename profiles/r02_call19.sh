#!/bin/bash
# Round-2 GPU call 19 (2 GPUs): targeted tail overlap (--tail-ctas): only the layer-0 dgrad GEMM leaves SMs to NCCL
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --no-cpu-baseline "$@" --timeline $O/tl_$tag.json > $O/scale_2_$tag.json 2>/dev/null; python - <<PY
import json
d=[json.loads(l) for l in open('$O/scale_2_$tag.json') if l.startswith('{')][-1]
t=json.load(open('$O/tl_$tag.json'))
b=t['buckets'][5]
print('$tag', 'ms/step', d['ms_per_step'], 'layer0 bucket ready', b['ready_ms'], 'done', b['done_ms'], 'past last wait', t['compute_stream_past_last_wait_ms'], 'day bucket ready', t['buckets'][6]['ready_ms'])
PY
}
run base
run tail16 --tail-ctas 16
run tail24 --tail-ctas 24
run tail32 --tail-ctas 32
run base2
