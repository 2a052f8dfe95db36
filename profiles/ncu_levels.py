"""Per-launch summary of an `ncu --set full` capture of the multifrontal kernels (one launch = one tree level):
    python profiles/ncu_levels.py <capture.ncu-rep> <n_cells> [out.json]
Prints a markdown table (duration, DRAM bytes, DRAM / L1TEX / issue utilisation, occupancy limiters, top stalls) and
optionally writes the DRAM bytes per cell (read by bench.py for roofline.traffic)."""
import csv
import json
import subprocess
import sys


def main(path, n_cells, out=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, rows = rows[0], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def f(r, name):
        try:
            return float(r[col[name]].replace(",", ""))
        except Exception:
            return float("nan")

    units = {h: u for h, u in zip(hdr, list(csv.reader(raw.splitlines()))[1])}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    ur, uw = scale.get(units["dram__bytes_read.sum"], 1.0), scale.get(units["dram__bytes_write.sum"], 1.0)
    tms = {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units["gpu__time_duration.sum"], 1.0)
    print("| launch | kernel | grid | block | regs | smem KB | ms | DRAM read GB | DRAM write GB | DRAM % | L1TEX % | issue % | warps % | DMMA pipe % | top stalls (warps per issue) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    tot_ms = tot_b = 0.0
    per_level = []
    for i, r in enumerate(rows):
        name = r[col["Kernel Name"]].split("(")[0].split("::")[-1]
        st = []
        for h, j in col.items():
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    st.append((float(r[j]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        st.sort(reverse=True)
        ms = f(r, "gpu__time_duration.sum") * tms
        rd, wr = f(r, "dram__bytes_read.sum") * ur / 1e9, f(r, "dram__bytes_write.sum") * uw / 1e9
        # units row says Gbyte / Mbyte: normalise through the units line of the csv
        per_level.append({"kernel": name, "ms": ms, "dram_read": rd, "dram_write": wr})
        tot_ms += ms
        print(f"| {i} | {name} | {r[col['launch__grid_size']]} | {r[col['launch__block_size']]} | {r[col['launch__registers_per_thread']]} | "
              f"{f(r, 'launch__shared_mem_per_block_dynamic'):.1f} | {ms:.3f} | {rd:.3f} | {wr:.3f} | "
              f"{f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {f(r, 'l1tex__throughput.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | {f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{f(r, 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | " + ", ".join(f"{h} {v:.1f}" for v, h in st[:4]) + " |")
    tot_b = sum(p["dram_read"] + p["dram_write"] for p in per_level) * 1e9
    print(f"\ntotal {tot_ms:.2f} ms, DRAM traffic {tot_b / 1e9:.2f} GB = {tot_b / n_cells / 1e6:.2f} MB per coarse cell ({n_cells} cells), "
          f"{tot_b / (tot_ms * 1e-3) / 1e9:.0f} GB/s under the profiler")
    if out:
        json.dump({"capture": f"{path.split('/')[-1]}: ncu --set full --clock-control none, {n_cells} cells, {len(rows)} launches (one per tree level)",
                   "n_cells": n_cells, "dram_bytes_per_cell": tot_b / n_cells, "ms_under_profiler": tot_ms,
                   "levels": [{"kernel": p["kernel"], "ms": p["ms"], "dram_bytes": (p["dram_read"] + p["dram_write"]) * 1e9} for p in per_level]},
                  open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else None)
