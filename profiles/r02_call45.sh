#!/bin/bash
# Round-2 GPU call 45: the secondary lines at HEAD: Conformer graph step, offline inference and streaming latency
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 100 python bench.py --mode conformer --graph --no-cpu-baseline > $O/final_conformer.json 2> $O/final_conformer.err; grep '^{' $O/final_conformer.json | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('conformer graph', d['ms_per_step'], d['value'], d.get('e2e',{}).get('value'))"
timeout 60 python bench.py --mode infer > $O/final_infer.jsonl 2> $O/final_infer.err; cut -c1-260 $O/final_infer.jsonl | grep '^{' | head -4
timeout 60 python bench.py --mode stream > $O/final_stream.jsonl 2> $O/final_stream.err; cut -c1-300 $O/final_stream.jsonl | grep '^{' | head -4
