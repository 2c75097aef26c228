#!/bin/bash
# Round-2 GPU call 32: programmatic dependent launch on every non-cooperative kernel: full GPU suite, then A/B of the headline step and the Conformer step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/gputests_pdl.log 2>&1; echo "tests rc=$?"; tail -3 $O/gputests_pdl.log
for v in 1 0; do
  echo "NSD_PDL=$v"
  NSD_PDL=$v timeout 300 python bench.py --no-cpu-baseline --breakdown > $O/bench_pdl$v.json 2> $O/bench_pdl$v.err; grep '^{' $O/bench_pdl$v.json | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('gru', d['ms_per_step'], d['value'], d['e2e']['value'], d.get('roofline'))"
  NSD_PDL=$v timeout 300 python bench.py --mode conformer --graph --no-cpu-baseline > $O/conf_pdl$v.json 2> $O/conf_pdl$v.err; grep '^{' $O/conf_pdl$v.json | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('conformer graph', d['ms_per_step'], d['value'])"; tail -2 $O/conf_pdl$v.err
done
