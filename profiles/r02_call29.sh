#!/bin/bash
# Round-2 GPU call 29: deferral at 64-row chains (B=256; bit 6 flips the WPC condition -> with bit 6, WPC=2 defers and WPC=1 does not), K3 tests, B=128
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
for f in 0 64; do for b in 256 128; do echo -n "NSD_GRU_DEBUG=$f  "; B=$b TP=60 NSD_GRU_DEBUG=$f timeout 120 python scratch/gru_time.py 2>&1 | tail -1; done; done | tee $O/k3_defer_wpc2.log
NSD_GRU_DEBUG=64 timeout 600 python -m pytest tests/test_gpu_gru_tc.py tests/test_gpu_fullshape.py -m gpu -x -q -k "256 or gru" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_gru_tc.py tests/test_gpu_fullshape.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
