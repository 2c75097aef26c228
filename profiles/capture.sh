#!/bin/bash
# Run on the GPU box (gpurun): launch list + one full capture of the dominant kernel, after a plain run exited 0.
#   gpurun --timeout 1500 -- 'bash profiles/capture.sh'
# The cooperative cluster kernels (gru_*_ts_kernel) cannot be launched under ncu's replay; they are excluded from
# profiling with a negative look-ahead and run natively (their time comes from CUDA events / in-kernel stamps).
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(?!.*gru_).*$' -c 600 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# dominant kernel: the K2 GEMM.  Skip the warm-up steps' GEMM launches (3 steps x 23) and take the first three of a timed step
# (layer-0 W_ih forward 7552x6144x8192, layer-1, layer-2).
ncu --set full --clock-control none --import-source on -k "regex:gemm_tc2?_kernel" -s 69 -c 3 -f -o gpurun_out/prof_gemm \
    $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
# K1 / Adam / CTC: one launch each, for the HBM-bound rooflines
ncu --set full --clock-control none --import-source on -k 'regex:frontend_fwd_kernel|frontend_bwd_tc_kernel|adam_kernel|ctc_kernel' -s 12 -c 4 -f \
    -o gpurun_out/prof_misc $CMD > gpurun_out/ncu_misc.log 2>&1
echo "misc capture rc=$?"
ls -la gpurun_out/
