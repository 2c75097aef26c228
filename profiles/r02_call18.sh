#!/bin/bash
# Round-2 GPU call 18 (2 GPUs): per-bucket all-reduce timeline of one data-parallel step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --no-cpu-baseline --timeline $O/allreduce_timeline_2gpu.json > $O/scale_2_tl.json 2> $O/scale_2_tl.err; echo "rc=$?"; tail -2 $O/scale_2_tl.err | cut -c1-200
cat $O/allreduce_timeline_2gpu.json
