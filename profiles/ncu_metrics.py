"""Concise key-metric dump of an .ncu-rep (read here, no GPU needed):
    python profiles/ncu_metrics.py gpurun_out/prof_update_r1b.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]

def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units, rows = r[0], r[1], r[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print("kernel:", rows[0][col["Kernel Name"]][:80], f"({len(rows)} launches captured)")
    for w in WANT:
        if w in col:
            print(f"  {w:80s} [{units[col[w]]}] " + "  ".join(row[col[w]][:10] for row in rows))
    st = []
    for h, i in col.items():
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                st.append((float(rows[0][i]), h))
            except ValueError:
                pass
    print("  top stall reasons (warps per issue-active cycle, first launch):")
    for v, h in sorted(st, reverse=True)[:6]:
        print(f"    {v:6.2f}  " + h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))

if __name__ == "__main__":
    main(sys.argv[1])
