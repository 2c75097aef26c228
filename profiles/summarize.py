"""Turn gpurun_out/ ncu artefacts into the small tracked summaries under profiles/ (run in the build container).
  python profiles/summarize.py launches gpurun_out/launches_r1.csv profiles/r01_launches_summary.csv
  python profiles/summarize.py rep gpurun_out/prof_gemm_big_r1.ncu-rep profiles/r01_gemm_ncu.txt
"""
import csv, io, re, subprocess, sys

def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
    agg = {}
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "").strip()
        name = re.sub(r"<unnamed>::", "", name)
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + float(r[14]) / 1e3)
    tot = sum(t for _, t in agg.values())
    with open(dst, "w") as f:
        f.write("kernel,launches,total_us,share\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{t:.1f},{t / tot:.4f}\n")
        f.write(f"\"TOTAL ({len(rows)} launches listed; cooperative-cluster gru_*_tc kernels cannot be launched under ncu and are excluded)\",{len(rows)},{tot:.1f},1.0\n")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]

def rep(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none, source: {src}\n")
        for r in data:
            f.write(f"\n== {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
            for i, h in enumerate(hdr):
                base = h.split(".TriageCompute.")[-1]
                if base in KEYS:
                    f.write(f"{base:80s} {r[i]:>16s} {units[i]}\n")

def traffic(src, dst):
    """dram__bytes_read.sum + dram__bytes_write.sum of the first kernel in an `ncu --set full` report -> small JSON that
    bench.py quotes as roofline.traffic (the layer-0 W_ih forward GEMM, the largest launch of the step)."""
    import json
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, r = rows[0], rows[1], rows[2]
    def val(key):
        i = [j for j, h in enumerate(hdr) if h.endswith(key)][0]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}[units[i]]
        return float(r[i].replace(",", "")) * mult
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    json.dump({"kernel": r[hdr.index("Kernel Name")][:80], "launch": "layer-0 W_ih forward 7552x6144x8192 (largest K2 launch of the step)",
               "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
               "algorithmic_bytes": 7552 * 8192 * 2 + 6144 * 8192 * 2 + 7552 * 6144 * 4,
               "duration_us": val("gpu__time_duration.sum"),
               "source": src}, open(dst, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "rep": rep, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
