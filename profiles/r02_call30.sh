#!/bin/bash
# Round-2 GPU call 30: fast front-end forward vs the generic kernel (bit-identity + time), front-end tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 600 python scratch/frontend_ab.py 2>&1 | tee $O/frontend_ab.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3
