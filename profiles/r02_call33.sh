#!/bin/bash
# Round-2 GPU call 33: K3 with four 32-row chains in flight for B > 64 (<1,4>) against the 64-row-chain form (<2,2>, NSD_GRU_WPC=2): tests, timing; LossReader e2e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gru_tc.py tests/test_gpu_fullshape.py -m gpu -x -q 2>&1 | tail -3
for m in 0 2; do for b in 256 128 96; do echo -n "NSD_GRU_WPC=$m  "; B=$b TP=60 NSD_GRU_WPC=$m timeout 120 python scratch/gru_time.py 2>&1 | tail -1; done; done | tee $O/k3_four_chains.log
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "train_step" 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline > $O/bench_lossreader.json 2> $O/bench_lossreader.err; grep '^{' $O/bench_lossreader.json | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('gru', d['ms_per_step'], d['value'], d['e2e'])"
timeout 600 python bench.py --no-cpu-baseline --T 2000 --batch 256 > $O/long_1.json 2> $O/long_1.err; grep '^{' $O/long_1.json | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('long', d['ms_per_step'], d['value'], d['e2e'])"
