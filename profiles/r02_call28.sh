#!/bin/bash
# Round-2 GPU call 28: K3 publish variants: deferred stores held behind the publisher's fence (default now; bit 4 = old order), bulk-store publish (bit 5), + relaxed red (bit 1, unsafe, timing only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
for f in 0 16 32 48 34 2; do echo -n "NSD_GRU_DEBUG=$f  "; NSD_GRU_DEBUG=$f timeout 120 python scratch/gru_time.py 2>&1 | tail -1; done | tee $O/k3_publish_variants.log
for f in 0 32; do echo "tests with NSD_GRU_DEBUG=$f"; NSD_GRU_DEBUG=$f timeout 600 python -m pytest tests/test_gpu_gru_tc.py -m gpu -x -q 2>&1 | tail -3; done | tee -a $O/k3_publish_variants.log
