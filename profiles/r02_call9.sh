#!/bin/bash
# Round-2 GPU call 9: fused streaming push (nsd_stream_push) tests + latency; K3 hardware counters of the final kernel form; ncu of the push kernel.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_streaming.py -m gpu -x -q -s > $O/stream_tests.log 2>&1; echo "stream tests rc=$?"; grep -E "fast streaming|streaming vs|passed|failed|Error|error" $O/stream_tests.log | head -20
timeout 300 python bench.py --mode stream > $O/stream.jsonl 2> $O/stream.err; echo "stream rc=$?"; cat $O/stream.jsonl; tail -3 $O/stream.err
# K3 counters: non-cooperative cluster launch, hardware-counter sections + the explicit north-star metrics (no SASS patching)
M=lts__t_sectors.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max
NSD_GRU_NO_COOP=1 timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis --metrics $M --clock-control none -k regex:gru_ -s 1 -c 2 -f -o $O/prof_k3 python tests/trace_gru.py > $O/ncu_k3.log 2>&1; echo "ncu k3 rc=$?"; tail -3 $O/ncu_k3.log
# the streaming push kernel under ncu (cooperative, no clusters): full set
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_push -s 20 -c 1 -f -o $O/prof_stream python bench.py --mode stream > $O/ncu_stream.log 2>&1; echo "ncu stream rc=$?"; tail -3 $O/ncu_stream.log
ls -la $O | tail -8
