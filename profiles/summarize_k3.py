#!/usr/bin/env python
"""K3 recurrence under Nsight Compute -> profiles/r02_gru_ncu.txt.

    NSD_GRU_NO_COOP=1 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats \
        --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis --clock-control none -k regex:gru_ ... -o gpurun_out/prof_k3 python tests/trace_gru.py
    python profiles/summarize_k3.py gpurun_out/prof_k3.ncu-rep > profiles/r02_gru_ncu.txt

(--set full needs SASS patching for its source counters, which pushes the 168-register, 1-CTA/SM kernels over the register file:
 LaunchFailed.  The hardware-counter sections above need no patching.)"""
import csv
import subprocess
import sys

KEEP = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__cluster_size", "cluster"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor (hmma subpipe) inst % of peak"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor-memory (TMEM) cycles active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (occupancy achieved)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("lts__t_sectors.sum", "L2 sectors"), ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__t_sectors_srcunit_tex.sum", "L2 sectors from SMs (tex/LSU)"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts (LSU)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("smsp__inst_executed_op_shared_ld.sum", "shared loads (inst)"), ("smsp__inst_executed_op_shared_st.sum", "shared stores (inst)"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["long_scoreboard", "short_scoreboard", "barrier", "membar", "wait", "sleeping", "lg_throttle", "mio_throttle", "math_pipe_throttle",
               "branch_resolving", "dispatch_stall", "drain", "no_instruction", "not_selected", "selected", "tex_throttle", "misc"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("=" * 100)
        print(r[col["Kernel Name"]])
        for key, label in KEEP:
            if key in col and r[col[key]] != "":
                print(f"  {label:48s} {r[col[key]]:>18s} {units[col[key]]}")
        print("  warp stall reasons (warps stalled per issue-active cycle; the largest say what the warps wait for):")
        st = []
        for n in STALL_NAMES:
            k = STALLS % n
            if k in col and r[col[k]] != "":
                st.append((float(r[col[k]].replace(",", "")), n))
        for v, n in sorted(st, reverse=True)[:8]:
            print(f"      {n:24s} {v:8.3f}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof_k3.ncu-rep")
